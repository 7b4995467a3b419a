"""End-to-end hot path (trace -> pack -> heads -> compositing fwd/bwd -> heads bwd): CUDA-graph replay vs eager launches."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scene(n_side=96):
    from volsurfs_b200.pipeline import make_synthetic_renderer
    from volsurfs_b200.synthetic import camera_rays

    renderer, _ = make_synthetic_renderer(K=5, n_lat=64, n_lon=64, hidden=(64, 64, 64))
    o, d = camera_rays(n_side, n_side)
    g = torch.Generator().manual_seed(7)
    feats = (torch.rand(o.shape[0] * 5, 51, generator=g) * 2 - 1).cuda()
    gt = torch.rand(o.shape[0], 3, generator=g).cuda()
    return renderer, o.cuda(), d.cuda(), feats, gt


def test_graphed_step_matches_eager_and_follows_input_updates(lib):
    from volsurfs_b200.pipeline import GraphedTrainingStep
    from volsurfs_b200.synthetic import camera_rays

    renderer, o, d, feats, gt = _scene()
    eager = renderer.render_fwd_bwd(o, d, feats, gt)
    want = {k: eager[k].clone() for k in ("rgb", "loss", "grad_rgb", "grad_alpha", "d_features_rgb")}
    before = lib.vs_launch_count()
    step = GraphedTrainingStep(renderer, o, d, feats, gt)
    captured = lib.vs_launch_count() - before
    out = step.replay()
    torch.cuda.synchronize()
    assert lib.vs_launch_count() - before == captured, "a replay must not go through the per-kernel host path"
    n = int(out["ray_samples_packed"].total_dev.item())
    assert n == int(eager["ray_samples_packed"].total_dev.item()) and n > 0
    for k, w in want.items():
        assert torch.equal(out[k], w), k  # same kernels, same order, deterministic reductions: bit-identical

    # new inputs are copied INTO the static buffers; the replay must see them
    o2, d2 = camera_rays(96, 96, azimuth_deg=75.0)
    o.copy_(o2)
    d.copy_(d2)
    gt.mul_(0.5)
    out = step.replay()
    got = {k: out[k].clone() for k in want}
    ref = renderer.render_fwd_bwd(o, d, feats, gt)
    torch.cuda.synchronize()
    assert not torch.equal(got["rgb"], want["rgb"])
    for k in want:
        assert torch.equal(got[k], ref[k]), k


def test_pipelined_host_fed_loop_matches_eager(lib):
    """every step's image / loss must come back through the pinned buffers one step later, for rays that change every step"""
    from volsurfs_b200.pipeline import PipelinedTrainingStep
    from volsurfs_b200.synthetic import camera_rays

    renderer, o, d, feats, gt = _scene()
    views = [camera_rays(96, 96, azimuth_deg=a) for a in (10.0, 50.0, 90.0, 130.0, 170.0)]
    want = []
    for vo, vd in views:
        ref = renderer.render_fwd_bwd(vo.cuda(), vd.cuda(), feats, gt)
        want.append((ref["rgb"].cpu(), float(ref["loss"]), ref["grad_rgb"].clone()))
    o_pin, d_pin = views[0][0].clone().pin_memory(), views[0][1].clone().pin_memory()
    img_pin = torch.empty((o.shape[0], 3), dtype=torch.float32).pin_memory()
    loss_pin = torch.empty((), dtype=torch.float32).pin_memory()
    loop = PipelinedTrainingStep(renderer, o_pin, d_pin, feats, gt, img_pin, loss_pin)
    loop.prime()
    torch.cuda.synchronize()
    for i in range(len(views)):
        if i + 1 < len(views):  # rays of the NEXT step go into the pinned buffers before step i is submitted
            o_pin.copy_(views[i + 1][0])
            d_pin.copy_(views[i + 1][1])
        out = loop.step(i)
        torch.cuda.synchronize()
        assert torch.equal(out["grad_rgb"], want[i][2]), f"step {i}: head gradients"
        if i > 0:
            assert torch.equal(img_pin, want[i - 1][0]), f"image of step {i - 1}"
            assert float(loss_pin) == want[i - 1][1]
    loop.drain(len(views))
    torch.cuda.synchronize()
    assert torch.equal(img_pin, want[-1][0]) and float(loss_pin) == want[-1][1]
