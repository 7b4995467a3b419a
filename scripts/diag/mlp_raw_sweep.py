"""diagnostic: vs_mlp_forward_raw / vs_mlp_backward_stashed_raw against torch autograd of the fp16-rounded forward, over output widths
and row counts (partial tiles included)"""
import sys

import torch

sys.path.insert(0, ".")
from oracle import shtex as O  # noqa: E402  (diagnostic script, not product)
from volsurfs_b200.textures import TextureNetwork  # noqa: E402

torch.manual_seed(0)
bad = 0
for n_out in (1, 3, 7, 8, 9, 15, 16, 17, 21, 32):
    net = TextureNetwork(n_out).cuda()
    for rows in (1, 37, 120, 128, 129, 300, 800, 5000):
        for rep in range(2):
            feat = (torch.randn(rows, 32, device="cuda") * 0.5).half().float()
            g = torch.randn(rows, n_out, device="cuda").half().float()
            stash = net.new_stash(rows, "cuda")
            raw = net.mlp_raw(feat, stash)
            ws = [w.detach().clone().requires_grad_() for w in net.weights]
            f = feat.clone().requires_grad_()
            want = O.mlp_half_forward(f, ws)
            (want * g).sum().backward()
            L = __import__("volsurfs_b200._lib", fromlist=["lib"]).lib()
            import ctypes
            n = len(net.weights)
            dims = net._dims_c()
            n_params = int(L.vs_mlp_num_params(n, dims))
            wsb = int(L.vs_mlp_backward_workspace_bytes(n, dims, 32, -1, 0, rows))
            wsbuf = torch.empty(wsb, dtype=torch.uint8, device="cuda")
            flat = torch.full((n_params,), float("nan"), device="cuda")
            d_feat = torch.full((rows, 32), float("nan"), device="cuda")
            code = L.vs_mlp_backward_stashed_raw(n, dims, net.packed().data_ptr(), stash.data_ptr(), g.data_ptr(), d_feat.data_ptr(), flat.data_ptr(), 0,
                                                 wsbuf.data_ptr(), rows, None, torch.cuda.current_stream().cuda_stream)
            assert code == 0, code
            torch.cuda.synchronize()
            e_fwd = ((raw - want.detach()).abs() / want.detach().abs().clamp(min=1)).max().item()
            o, errs = 0, []
            for w, wr in zip(net.weights, ws):
                k = w.numel()
                got = flat[o:o + k].view_as(w)
                o += k + w.shape[0]
                rms = wr.grad.pow(2).mean().sqrt().item() or 1.0
                errs.append(((got - wr.grad).abs() / wr.grad.abs().clamp(min=rms)).max().item())
            rmsf = f.grad.pow(2).mean().sqrt().item() or 1.0
            e_in = ((d_feat - f.grad).abs() / f.grad.abs().clamp(min=rmsf)).max().item()
            flag = "" if (e_fwd < 2e-3 and max(errs) < 3e-2 and e_in < 3e-2) else "  <-- BAD"
            bad += bool(flag)
            print(f"n_out {n_out:2d} rows {rows:5d} rep {rep}: fwd {e_fwd:.1e} dW {[f'{e:.1e}' for e in errs]} d_in {e_in:.1e}{flag}")
print("bad:", bad)
