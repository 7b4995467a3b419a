"""GPU: the full path with the reference's DEFAULT appearance (per-layer SH neural textures, config/volsurfs/base_5.cfg) against the
oracle chain: C ray tracer -> hits per mesh -> texture coordinates (volsurfs.py:509-516) -> oracle SHNeuralTextures per layer
(oracle/shtex.py) -> alpha decay -> dense K-layer compositing in torch (volsurfs.py:601-640,708; fp32), gradients by torch autograd.

Bars: texture coordinates bit-exact; image within 2e-3 at p99 and 4e-2 max (8-bit quantisation flips, see tests/test_gpu_shtex.py);
parameter gradients <= 1e-1 under grad_err (fp16 dZ in the texture networks; ~400 hits per network here, so a single ReLU unit whose
pre-activation ties within summation noise — see the "no ReLU ties" rule of the shtex fixtures — moves an entry by several % of the rms;
observed worst 5.8e-2)."""
import numpy as np
import pytest
import torch

from conftest import grad_err
from oracle import shtex as O
from oracle.raytrace import OracleRayTracer

pytestmark = pytest.mark.gpu
K, NLAT, NLON = 3, 48, 48
DEG_RES = (64, 32, 16, 8)


def _oracle_nets(model):
    nets = []
    for nt in model.neural_textures:
        net = O.TextureNet(nt.model.n_out, seed=0)
        net.table = nt.model.table.detach().cpu().clone().requires_grad_()
        net.weights = [w.detach().cpu().clone().requires_grad_() for w in nt.model.weights]
        nets.append(net)
    return nets


def test_textured_path_against_oracle_chain():
    from volsurfs_b200.pipeline import make_synthetic_textured_renderer
    from volsurfs_b200.synthetic import camera_rays, shell_face_uvs

    renderer, meshes = make_synthetic_textured_renderer(K=K, n_lat=NLAT, n_lon=NLON, deg_res=DEG_RES, table_init=0.5)
    o, d = camera_rays(40, 40)
    N = o.shape[0]
    torch.manual_seed(3)
    gt = torch.rand(N, 3)
    out = renderer.render_fwd_bwd(o.cuda(), d.cuda(), gt.cuda())
    rsp = out["ray_samples_packed"]
    S = rsp.get_total_nr_samples()
    assert S > 500

    # ---- oracle: per-mesh trace, texture coordinates, per-layer models, dense compositing
    lay = OracleRayTracer(meshes).trace_layers(o.numpy(), d.numpy(), mode="bvh")
    fuv = torch.from_numpy(shell_face_uvs(NLAT, NLON))
    hit = torch.from_numpy(lay["is_hit"]).T.contiguous()                       # [N,K] mesh order
    rgb_d = torch.zeros(N, K, 3)
    alpha_d = torch.zeros(N, K, 1)
    nets_rgb = [_oracle_nets(m) for m in renderer.rgb_models]
    nets_alpha = [_oracle_nets(m) for m in renderer.alpha_models]
    uv_all = {}
    for k in range(K):
        h = hit[:, k]
        if not h.any():
            continue
        res = lay["per_mesh"][k]
        bary = torch.from_numpy(res["barycentric"])
        uvp = torch.sum(bary.unsqueeze(-1) * fuv[torch.from_numpy(res["triangles_id"]).clamp(min=0)], dim=1)[h]   # volsurfs.py:511-514
        uv_all[k] = uvp
        dirs = d[h]
        normals = torch.from_numpy(res["normals"])[h]
        kw = dict(sh_deg=3, sh_range=[15.0] * 4, deg_res=list(DEG_RES), anchor=False, lerp=True)
        c = O.sh_neural_textures_forward(nets_rgb[k], uvp.clone(), dirs, nr_channels=3, **kw)
        a = O.sh_neural_textures_forward(nets_alpha[k], uvp.clone(), dirs, nr_channels=1, **kw)
        with torch.no_grad():
            decay = torch.sigmoid(10.0 * torch.sum(-dirs * normals, dim=1, keepdim=True).clamp(0.0, 1.0)) * 2.0 - 1.0
        idx = torch.nonzero(h).flatten()
        rgb_d = rgb_d.index_put((idx, torch.full_like(idx, k)), c)
        alpha_d = alpha_d.index_put((idx, torch.full_like(idx, k)), a * decay)
    a_f, c_f = torch.flip(alpha_d, dims=[1]), torch.flip(rgb_d, dims=[1])      # volsurfs.py:601-603 (outer -> inner), fp32 variant
    Tc = torch.cumprod(1 - a_f, dim=1)
    T = torch.cat([torch.ones_like(Tc[:, :1]), Tc[:, :-1]], dim=1)
    w = T * a_f
    pred_o = (c_f * w).sum(dim=1) + Tc[:, -1] * 1.0                            # white background (volsurfs.py:708)
    loss_o = (pred_o - gt).abs().mean()
    loss_o.backward()

    # texture coordinates of the packed hits: bit-exact (packed order: ray-major, outer -> inner)
    layer = rsp.samples_layer.view(-1).cpu().numpy()
    se = rsp.ray_start_end_idx.cpu().numpy()
    ray_of = np.repeat(np.arange(N), np.maximum(se[:, 1] - se[:, 0], 0))
    tex_uv = rsp.samples_tex_uv.cpu().numpy()
    for k in uv_all:
        rows = np.nonzero(layer == k)[0]
        assert np.array_equal(ray_of[rows], np.nonzero(hit[:, k].numpy())[0])
        assert np.array_equal(tex_uv[rows], uv_all[k].numpy()), f"texture coordinates of layer {k} differ"

    err = np.abs(out["rgb"].detach().cpu().numpy() - pred_o.detach().numpy())
    print("textured path: %d hits, image |err| p99 %.2e max %.2e, loss %.6f vs %.6f" % (S, np.quantile(err, 0.99), err.max(), float(out["loss"]),
                                                                                        float(loss_o.detach())))
    assert np.quantile(err, 0.99) < 2e-3 and err.max() < 4e-2
    worst = 0.0
    for k in range(K):
        for models, nets in ((renderer.rgb_models, nets_rgb), (renderer.alpha_models, nets_alpha)):
            for nt, net in zip(models[k].neural_textures, nets[k]):
                for wg, wo in zip(nt.model.weights, net.weights):
                    if wo.grad is None:
                        continue
                    worst = max(worst, grad_err(wg.grad.cpu().numpy(), wo.grad.numpy()))
                if net.table.grad is not None:
                    worst = max(worst, grad_err(nt.model.table.grad.cpu().numpy(), net.table.grad.numpy()))
    print("textured path: worst parameter grad_err %.2e" % worst)
    assert worst < 1e-1
