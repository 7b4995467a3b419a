"""GPU parity of packing (bit-exact: integer/index work and single-rounded fp32)."""
import numpy as np
import pytest
import torch

from oracle.packing import RaySamplesPackedNP, pack_layer_hits

pytestmark = pytest.mark.gpu

import ctypes
from pathlib import Path

REF_PATH = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "libsampler_ref.so"


@pytest.fixture(scope="module")
def ref():
    """the reference's own RaySamplesPackedGPU.cuh kernels (compiled where they lie, oracle/ref_sampler_harness.cu)"""
    if not REF_PATH.exists():
        pytest.skip("oracle/_ref/libsampler_ref.so not built (python -m oracle.build where /root/reference is mounted)")
    lib = ctypes.CDLL(str(REF_PATH))
    assert lib.ref_sampler_abi_version() == 2
    return lib


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def ref_compact(ref, rsp):
    """src/RaySamplesPacked.cu:188-273 around the reference's kernel: fresh packet (constructor fills of :13-48), torch prefix sum
    (:218-221), compact_to_valid_samples_gpu"""
    from volsurfs_b200.volsurfs import RaySamplesPacked

    n = rsp.get_nr_rays()
    cnt = (rsp.ray_start_end_idx[:, 1] - rsp.ray_start_end_idx[:, 0])
    total = int(cnt.sum().item())
    vd = rsp.samples_values.shape[1]
    out = RaySamplesPacked(n, total, 0, vd)
    if total == 0:
        return out
    start = cnt.cumsum(0).to(torch.int32)
    start = torch.cat([torch.zeros(1, dtype=torch.int32, device=start.device), start[:-1]]).contiguous()
    rc = ref.ref_compact_to_valid_samples(n, rsp.get_max_nr_samples(), total, vd, P(rsp.samples_idx), P(rsp.samples_3d), P(rsp.samples_dirs),
                                          P(rsp.samples_z), P(rsp.samples_dt), P(rsp.samples_values), P(rsp.ray_start_end_idx), P(start),
                                          P(out.samples_idx), P(out.samples_3d), P(out.samples_dirs), P(out.samples_z), P(out.samples_dt),
                                          P(out.samples_values), P(out.ray_start_end_idx))
    assert rc == 0
    return out


def _uncompacted(n_rays, M, seed, values_dim=2, p_empty=0.1):
    """slot layout of RaySamplerGPU.cuh:206-271: ray r owns slots [r*M, r*M+cnt_r)"""
    rng = np.random.default_rng(seed)
    ref = RaySamplesPackedNP(n_rays, n_rays * M, 0, values_dim)
    ref.is_compacted = False
    cnt = rng.integers(0, M + 1, n_rays)
    cnt[rng.random(n_rays) < p_empty] = 0
    has = cnt > 0
    ref.ray_start_end_idx[has, 0] = np.nonzero(has)[0] * M
    ref.ray_start_end_idx[has, 1] = np.nonzero(has)[0] * M + cnt[has]
    for name in ("samples_3d", "samples_dirs", "samples_z", "samples_dt", "samples_values", "ray_o", "ray_d", "ray_enter", "ray_exit",
                 "ray_max_dt"):
        arr = getattr(ref, name)
        arr[...] = rng.standard_normal(arr.shape).astype(np.float32)
    ref.has_samples_values = True
    return ref


def _to_gpu(ref):
    from volsurfs_b200.volsurfs import RaySamplesPacked

    rsp = RaySamplesPacked(ref.get_nr_rays(), ref.get_max_nr_samples(), 0, ref.samples_values.shape[1])
    for name in ("samples_idx", "samples_3d", "samples_dirs", "samples_z", "samples_dt", "samples_values", "ray_start_end_idx", "ray_o",
                 "ray_d", "ray_enter", "ray_exit", "ray_max_dt"):
        setattr(rsp, name, torch.from_numpy(getattr(ref, name)).cuda())
    rsp.is_compacted = ref.is_compacted
    rsp.has_samples_values = ref.has_samples_values
    return rsp


@pytest.mark.parametrize("n_rays,M", [(1, 4), (2048, 5), (5000, 9), (70001, 3), (300, 96)])
def test_compact_to_valid_samples_bit_exact(n_rays, M):
    ref = _uncompacted(n_rays, M, seed=n_rays + M)
    want = ref.compact_to_valid_samples()
    got = _to_gpu(ref).compact_to_valid_samples()
    assert got.is_compacted and got.get_total_nr_samples() == want.get_total_nr_samples() == got.get_max_nr_samples()
    for name in ("ray_start_end_idx", "samples_idx", "samples_3d", "samples_dirs", "samples_z", "samples_dt", "samples_values", "ray_o",
                 "ray_d", "ray_enter", "ray_exit", "ray_max_dt"):
        assert np.array_equal(getattr(got, name).cpu().numpy(), getattr(want, name)), name
    with pytest.raises(RuntimeError):
        got.compact_to_valid_samples()  # CHECK(!is_compacted), RaySamplesPacked.cu:193


@pytest.mark.parametrize("n_rays,M", [(1, 4), (2048, 5), (5000, 9), (70001, 3), (300, 96)])
def test_compact_equals_reference_kernel(ref, n_rays, M):
    """product == the reference's compact_to_valid_samples_gpu (RaySamplesPackedGPU.cuh:172-257), every array bit for bit"""
    src = _to_gpu(_uncompacted(n_rays, M, seed=n_rays + M))
    src.samples_idx = torch.arange(n_rays * M, dtype=torch.int32, device="cuda").reshape(-1, 1)  # RaySamplesPacked.cu:20
    want = ref_compact(ref, src)
    got = src.compact_to_valid_samples()
    assert got.get_max_nr_samples() == want.get_max_nr_samples()
    for name in ("ray_start_end_idx", "samples_idx", "samples_3d", "samples_dirs", "samples_z", "samples_dt", "samples_values"):
        assert torch.equal(getattr(got, name), getattr(want, name)), name
    empty = (src.ray_start_end_idx[:, 1] - src.ray_start_end_idx[:, 0]) == 0
    assert bool((got.ray_start_end_idx[empty] == -1).all())                                     # RaySamplesPackedGPU.cuh:211-212


@pytest.mark.parametrize("n_rays,M", [(1, 4), (2048, 5), (5000, 9), (70001, 3), (300, 96)])
@pytest.mark.parametrize("is_background", [False, True])
def test_update_dt_equals_reference_kernel(ref, n_rays, M, is_background):
    """product == the reference's update_dt_gpu (RaySamplesPackedGPU.cuh:14-88); a third of the rays carry the constructor's
    ray_max_dt = -1 (importance_sample / init_with_one_sample_per_ray packets): clamp(x, 0, -1) = fmaxf(0, fminf(x, -1)) = 0"""
    rsp = _to_gpu(_uncompacted(n_rays, M, seed=7 * n_rays + M)).compact_to_valid_samples()
    if rsp.get_total_nr_samples() == 0:
        pytest.skip("no samples")
    rsp.samples_z = torch.sort(rsp.samples_z.abs(), dim=0).values.contiguous()      # increasing along every ray
    rsp.ray_exit = rsp.ray_exit.abs() + rsp.samples_z.max() * torch.rand_like(rsp.ray_exit)
    rsp.ray_max_dt = (rsp.ray_max_dt.abs() * 0.02).contiguous()
    rsp.ray_max_dt[::3] = -1.0
    want = torch.full_like(rsp.samples_dt, -1.0)
    rc = ref.ref_update_dt(n_rays, rsp.get_total_nr_samples(), int(is_background), P(rsp.ray_max_dt), P(rsp.ray_exit), P(rsp.samples_z),
                           P(rsp.ray_start_end_idx), P(want))
    assert rc == 0
    rsp.update_dt(is_background)
    assert torch.equal(rsp.samples_dt, want)
    assert float(want.min()) >= 0.0


def test_compact_all_empty():
    ref = RaySamplesPackedNP(100, 500, 0, 1)
    ref.is_compacted = False
    got = _to_gpu(ref).compact_to_valid_samples()
    assert got.get_max_nr_samples() == 0 and got.is_empty()
    assert np.all(got.ray_start_end_idx.cpu().numpy() == -1)


def test_container_accessors_match_reference_semantics():
    ref = _uncompacted(64, 6, seed=3).compact_to_valid_samples()
    rsp = _to_gpu(ref)
    assert rsp.get_nr_rays() == 64 and rsp.get_values_dim() == 2
    assert np.array_equal(rsp.get_nr_samples_per_ray().cpu().numpy(), ref.get_nr_samples_per_ray())
    r = int(np.nonzero(ref.get_nr_samples_per_ray() > 1)[0][0])
    s, e = ref.ray_start_end_idx[r]
    assert np.array_equal(rsp.get_ray_samples_z(r).cpu().numpy(), ref.samples_z[s:e])
    assert np.array_equal(rsp.get_ray_samples_3d(r).cpu().numpy(), ref.samples_3d[s:e])
    assert np.array_equal(rsp.get_ray_samples_values(r).cpu().numpy(), ref.samples_values[s:e])
    assert np.array_equal(rsp.get_ray_o(r).cpu().numpy(), ref.ray_o[r])
    c = rsp.copy()
    c.samples_z += 1
    assert np.array_equal(rsp.samples_z.cpu().numpy(), ref.samples_z)
    rsp.remove_samples_values()
    assert not rsp.are_samples_values_set() and bool((rsp.samples_values == -1).all())
    rsp.set_samples_values(torch.zeros(rsp.get_max_nr_samples(), 3).cuda())
    assert rsp.get_values_dim() == 3
    with pytest.raises(RuntimeError):
        rsp.set_samples_values(torch.zeros(rsp.get_max_nr_samples(), 3).cuda())


# K <= 16: the shared-memory staged scatter; K = 20: the thread-per-ray scatter it falls back to
@pytest.mark.parametrize("n_rays,K", [(4096, 5), (1000, 9), (33, 1), (100000, 5), (777, 16), (500, 20)])
def test_pack_layer_hits_bit_exact(n_rays, K):
    from volsurfs_b200.raytracer import pack_layer_hits as pack_gpu

    rng = np.random.default_rng(n_rays * 31 + K)
    rays_o = rng.standard_normal((n_rays, 3)).astype(np.float32)
    rays_d = rng.standard_normal((n_rays, 3)).astype(np.float32)
    hit = rng.random((n_rays, K)) < 0.7
    hit[rng.random(n_rays) < 0.05] = False
    depth = np.where(hit, rng.random((n_rays, K)) * 3 + 0.5, 1e6).astype(np.float32)
    depth[0, 0] = 100.0  # exactly t_far counts as a hit (raytracer.py:100: depth <= t_far)
    hit[0, 0] = True
    tri = rng.integers(0, 100000, (n_rays, K)).astype(np.int32)
    u = rng.random((n_rays, K)).astype(np.float32)
    v = rng.random((n_rays, K)).astype(np.float32)
    want_unc, layer_of_slot = pack_layer_hits(rays_o, rays_d, hit, depth)
    want = want_unc.compact_to_valid_samples()
    got = pack_gpu(torch.from_numpy(rays_o).cuda(), torch.from_numpy(rays_d).cuda(),
                   torch.from_numpy(np.ascontiguousarray(depth.T)).cuda(), torch.from_numpy(np.ascontiguousarray(tri.T)).cuda(),
                   torch.from_numpy(np.ascontiguousarray(u.T)).cuda(), torch.from_numpy(np.ascontiguousarray(v.T)).cuda(), t_far=100.0)
    S = want.get_total_nr_samples()
    assert got.get_max_nr_samples() == S
    for name in ("ray_start_end_idx", "samples_idx", "samples_3d", "samples_dirs", "samples_z", "ray_o", "ray_d"):
        assert np.array_equal(getattr(got, name).cpu().numpy(), getattr(want, name)), name
    lay = layer_of_slot[want.samples_idx[:, 0]]
    assert np.array_equal(got.samples_layer.cpu().numpy(), lay)
    ray = want.samples_idx[:, 0] // K
    assert np.array_equal(got.samples_triangle.cpu().numpy(), tri[ray, lay])
    assert np.array_equal(got.samples_uv.cpu().numpy(), np.stack([u[ray, lay], v[ray, lay]], 1))
