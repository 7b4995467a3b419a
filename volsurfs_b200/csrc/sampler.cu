// volsurfs_b200 — ray samplers and the occupancy-grid queries they rest on: the producers of RaySamplesPacked on the
// NeRF / NeuS / background path, i.e. the stage right before packed compositing.
//
// Replaces
//   RaySampler::compute_samples_fg                            src/RaySampler.cu:159-245,  kernel kernels/volsurfs/RaySamplerGPU.cuh:141-271
//   RaySampler::compute_samples_fg_in_grid_occupied_regions   src/RaySampler.cu:247-345,  kernel RaySamplerGPU.cuh:273-488
//   RaySampler::compute_samples_bg                            src/RaySampler.cu:72-157,   kernel RaySamplerGPU.cuh:39-139
//   OccupancyGrid::get_rays_t_near_t_far / check_occupancy    kernels/volsurfs/OccupancyGridGPU.cuh:318-441
//   pos_to_lin_idx / distance_to_next_voxel / morton3D        kernels/volsurfs/occ_grid_helpers.h:13-33,55-79,126-190
//
// The reference marches every ray into an UNCOMPACTED packet (nr_rays x max_nr_samples_per_ray slots of 36 bytes: 23 GB of staging at
// 640k rays x 1024) and then gathers it (compact_to_valid_samples).  Here the foreground samplers are two launches around one prefix sum:
//   sampler_fg_march_kernel<GRID>   thread per ray (the marches are sequential: every step depends on the position the previous one
//                                   reached): length of the occupied stretch -> sample count and spacing (ray_max_dt), then the sampling
//                                   march, which records only the DEPTH of every sample in a 4-byte-per-slot staging array
//   (vs_segment_offsets)            exclusive scan of the counts -> compacted start of every ray + the total
//   sampler_fg_expand_kernel        warp per ray, lanes over its samples: depth -> o + z d, rows stored at their compacted position
// so the result equals the reference's compacted packet (samples_idx = the slot the sample would have had, samples_dt untouched).  The
// marches are instruction bound (a step is ~25 dependent fp32 operations + a Morton code); the step shares p / extent between the voxel
// index and the boundary distance, spreads Morton bits in 32-bit arithmetic (the 64-bit path only outside the grid), and divides by
// the voxel count with an exact multiply when it is a power of two.  The occupancy && roi mask (1 byte per voxel) is L2 resident.
//
// Arithmetic contract (bit parity with the reference kernels as nvcc compiles them, pinned on the GPU against oracle/_ref/
// libsampler_ref.so): IEEE fp32 with the contractions nvcc applies to the reference source — `ray_o + t * ray_d`, `t + c * rnd` and
// helper_math's lerp are fused multiply-adds; everything else is separate round-to-nearest operations; the background sampler's
// `1.0 / (s + eps) - 1.0` is evaluated in double like the reference's double literals make it.
#include <algorithm>

#include "vs_common.cuh"

namespace vs {

// hard cap on the voxel marches below: a march through an n^3 grid takes O(n) steps (a few thousand with the +1e-6 nudges); the cap only
// ends marches whose t no longer advances in fp32 (|origin| huge against the step, infinite t_exit) instead of hanging the stream
constexpr int kMaxMarchSteps = 1 << 20;

// ---- pcg32 (kernels/volsurfs/pcg32.h:32-34,60-70,84-95,158-180) ------------------------------------------------------------------
struct Pcg {
    uint64_t state, inc;
    __device__ __forceinline__ void advance(uint64_t delta) {
        uint64_t cur_mult = 0x5851f42d4c957f2dULL, cur_plus = inc, acc_mult = 1u, acc_plus = 0u;
        while (delta > 0) {
            if (delta & 1) {
                acc_mult *= cur_mult;
                acc_plus = acc_plus * cur_mult + cur_plus;
            }
            cur_plus = (cur_mult + 1) * cur_plus;
            cur_mult *= cur_mult;
            delta >>= 1;
        }
        state = acc_mult * state + acc_plus;
    }
    __device__ __forceinline__ float next_float() {
        const uint64_t old = state;
        state = old * 0x5851f42d4c957f2dULL + inc;
        const uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        const uint32_t rot = (uint32_t)(old >> 59u);
        const uint32_t r = (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
        return __uint_as_float((r >> 9) | 0x3f800000u) - 1.0f;
    }
};

// ---- grid addressing (occ_grid_helpers.h) -----------------------------------------------------------------------------------------
struct Grid {
    int n;             // voxels per dimension
    float ex, ey, ez;  // extent of the cuboid, centred at the origin
    float inv_n;       // 1/n when n is a power of two (x / n == x * (1/n) exactly), else 0: divide
    const uint8_t* occ;
    const uint8_t* roi;  // NULL: `occ` already holds occupancy && roi
};

// 21-bit Morton spreading; the reference keeps only the low 32 bits of the 64-bit spread (uint32_t xx = expand_bits(x))
__device__ __forceinline__ uint32_t spread_bits64(uint32_t v) {
    uint64_t w = v;
    w &= 0x00000000001fffffULL;
    w = (w | w << 32) & 0x001f00000000ffffULL;
    w = (w | w << 16) & 0x001f0000ff0000ffULL;
    w = (w | w << 8) & 0x010f00f00f00f00fULL;
    w = (w | w << 4) & 0x10c30c30c30c30c3ULL;
    w = (w | w << 2) & 0x1249249249249249ULL;
    return (uint32_t)w;
}
// the same bits for v < 1024 in 32-bit arithmetic (the marches spend most of their instructions here)
__device__ __forceinline__ uint32_t spread_bits10(uint32_t v) {
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// voxel-space coordinate p / extent, shared by the index and the step computation (the reference evaluates it in both)
__device__ __forceinline__ float unit_coord(float p, float e) { return __fdiv_rn(p, e); }

__device__ __forceinline__ int lin_idx_from_unit(float ux, float uy, float uz, const Grid& g) {
    const float n = (float)g.n;
    // float -> uint32_t as the hardware converts it: truncation, negatives and NaN to 0, saturation above
    const uint32_t x = __float2uint_rz(__fmul_rn(__fadd_rn(ux, 0.5f), n)), y = __float2uint_rz(__fmul_rn(__fadd_rn(uy, 0.5f), n)),
                   z = __float2uint_rz(__fmul_rn(__fadd_rn(uz, 0.5f), n));
    if ((x | y | z) < 1024u) return (int)(spread_bits10(x) | (spread_bits10(y) << 1) | (spread_bits10(z) << 2));
    return (int)(spread_bits64(x) | (spread_bits64(y) << 1) | (spread_bits64(z) << 2));
}

__device__ __forceinline__ int pos_to_lin_idx(float px, float py, float pz, const Grid& g) {
    return lin_idx_from_unit(unit_coord(px, g.ex), unit_coord(py, g.ey), unit_coord(pz, g.ez), g);
}

// the reference's "DDA like step": distance (measured along the AXIS, not along the ray) to the next voxel boundary, + 1e-6
__device__ __forceinline__ float step_from_unit(float ux, float uy, float uz, float dx, float dy, float dz, const Grid& g) {
    const float eps = 1e-6f, n = (float)g.n;
    if (fabsf(dx) < eps && fabsf(dy) < eps && fabsf(dz) < eps) return 1e10f;
    float t3[3];
    const float u[3] = {ux, uy, uz}, d[3] = {dx, dy, dz}, e[3] = {g.ex, g.ey, g.ez};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        t3[a] = 1e10f;
        if (fabsf(d[a]) > eps) {
            const float q = __fmul_rn(u[a], n);
            const float sgn = d[a] > 0.f ? 1.f : -1.f;  // |d| > eps here
            const float prime = floorf(__fadd_rn(q, sgn));
            const float dq = fabsf(__fsub_rn(prime, q));
            t3[a] = __fmul_rn(g.inv_n != 0.f ? __fmul_rn(dq, g.inv_n) : __fdiv_rn(dq, n), e[a]);
        }
    }
    return __fadd_rn(fminf(fminf(t3[0], t3[1]), t3[2]), eps);
}

__device__ __forceinline__ float distance_to_next_voxel(float px, float py, float pz, float dx, float dy, float dz, const Grid& g) {
    return step_from_unit(unit_coord(px, g.ex), unit_coord(py, g.ey), unit_coord(pz, g.ez), dx, dy, dz, g);
}

__device__ __forceinline__ bool in_grid(int idx_voxel, const Grid& g) { return idx_voxel >= 0 && idx_voxel < g.n * g.n * g.n; }
__device__ __forceinline__ bool occupied(int idx_voxel, const Grid& g) {
    if (g.roi == nullptr) return __ldg(g.occ + idx_voxel) != 0;
    return __ldg(g.roi + idx_voxel) && __ldg(g.occ + idx_voxel);
}
__device__ __forceinline__ float clampf(float v, float a, float b) { return fmaxf(a, fminf(b, v)); }

// one step of a march: voxel of the position and the distance to the next boundary from one evaluation of p / extent
struct Probe {
    int voxel;
    float px, py, pz, ux, uy, uz;
};
__device__ __forceinline__ Probe probe(float t, float ox, float oy, float oz, float dx, float dy, float dz, const Grid& g) {
    Probe p;
    p.px = __fmaf_rn(t, dx, ox);
    p.py = __fmaf_rn(t, dy, oy);
    p.pz = __fmaf_rn(t, dz, oz);
    p.ux = unit_coord(p.px, g.ex);
    p.uy = unit_coord(p.py, g.ey);
    p.uz = unit_coord(p.pz, g.ez);
    p.voxel = lin_idx_from_unit(p.ux, p.uy, p.uz, g);
    return p;
}

// ---- foreground samplers -----------------------------------------------------------------------------------------------------------
// Pass 1 (one thread per ray): length of the occupied stretch -> sample count and spacing, then the sampling march.  It writes the
// ray's virtual uncompacted segment (ray*max_nr, ray*max_nr + created) or (-1,-1), ray_max_dt, and the depth of every sample into the
// ray's slots of `z_stage` (4 bytes per slot; only the slots of real samples are touched).
template <bool GRID>
__global__ void __launch_bounds__(128) sampler_fg_march_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                               const float* __restrict__ t_entry, const float* __restrict__ t_exit_p,
                                                               float min_dist, int min_nr, int max_nr, Pcg rng, int jitter, Grid g,
                                                               int32_t* __restrict__ se_virtual, float* __restrict__ ray_max_dt,
                                                               float* __restrict__ z_stage, int64_t n_rays) {
    const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    const float eps = 1e-6f;
    const float t_start = __ldg(t_entry + ray), t_exit = __ldg(t_exit_p + ray);
    const float ox = __ldg(rays_o + 3 * ray), oy = __ldg(rays_o + 3 * ray + 1), oz = __ldg(rays_o + 3 * ray + 2);
    const float dx = __ldg(rays_d + 3 * ray), dy = __ldg(rays_d + 3 * ray + 1), dz = __ldg(rays_d + 3 * ray + 2);

    // ---- how far does the ray travel through occupied space -> number of samples and their spacing
    float dist = 0.f;
    if (GRID) {
        float t = t_start, step = 0.f;
        while (t < t_exit) {
            const Probe p = probe(t, ox, oy, oz, dx, dy, dz, g);
            if (!in_grid(p.voxel, g)) break;
            if (occupied(p.voxel, g)) dist = __fadd_rn(dist, step);  // the step that LED here (reference quirk: lags one voxel)
            step = step_from_unit(p.ux, p.uy, p.uz, dx, dy, dz, g);
            t = __fadd_rn(t, step);
        }
        dist = clampf(dist, 0.f, __fsub_rn(t_exit, t_start));
    } else {
        dist = __fsub_rn(t_exit, t_start);
    }
    int to_create = 0;
    float spacing = 0.f;
    if (dist <= 0.f) {
        to_create = 0;
    } else if (dist > min_dist) {
        to_create = (int)__fdiv_rn(dist, min_dist);
        to_create = max(0, min(to_create, max_nr));
        spacing = __fdiv_rn(dist, (float)to_create);
    } else {
        to_create = 1;
        spacing = dist;
    }

    // ---- the sampling march
    int created = 0;
    const int64_t slot0 = ray * (int64_t)max_nr;
    if (to_create > 0 && to_create >= min_nr) {
        float t = t_start, to_next = 0.f;
        if (jitter) {
            rng.advance((uint64_t)ray);
            const float rnd = rng.next_float();
            if (GRID)
                to_next = __fmul_rn(spacing, rnd);
            else
                t = __fmaf_rn(spacing, rnd, t);
        }
        while (t < t_exit) {
            t = clampf(t, t_start, t_exit);
            if (created >= to_create) break;
            if (GRID) {
                const Probe p = probe(t, ox, oy, oz, dx, dy, dz, g);
                if (!in_grid(p.voxel, g)) break;
                const bool occ = occupied(p.voxel, g);
                if (occ && to_next == 0.f) {
                    z_stage[slot0 + created] = t;
                    ++created;
                    to_next = spacing;
                }
                const float to_voxel = step_from_unit(p.ux, p.uy, p.uz, dx, dy, dz, g);
                float step = to_voxel;
                if (occ) {
                    step = fminf(to_voxel, to_next);
                    to_next = __fsub_rn(to_next, step);
                    if (to_next <= eps) to_next = 0.f;
                }
                t = __fadd_rn(t, step);
            } else {
                z_stage[slot0 + created] = t;
                ++created;
                t = __fadd_rn(t, spacing);
            }
        }
    }
    // fewer than min_nr samples: the ray keeps (-1,-1) and ray_max_dt = -1 (the RaySamplesPacked constructor fill); otherwise its
    // spacing is recorded — also for a ray with zero samples when min_nr == 0 (reference behaviour, RaySamplerGPU.cuh:253-262)
    if (created >= min_nr) {
        ray_max_dt[ray] = spacing;
        reinterpret_cast<int2*>(se_virtual)[ray] = make_int2((int)slot0, (int)slot0 + created);
    } else {
        reinterpret_cast<int2*>(se_virtual)[ray] = make_int2(-1, -1);
    }
}

// Pass 2 (one warp per ray, lanes over its samples): depth -> position, rows stored at the ray's compacted offset with unit stride.
__global__ void __launch_bounds__(256) sampler_fg_expand_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, int max_nr,
                                                                const int32_t* __restrict__ se_virtual, const float* __restrict__ z_stage,
                                                                const int32_t* __restrict__ out_start, int32_t* __restrict__ se_out,
                                                                int32_t* __restrict__ s_idx, float* __restrict__ s_3d, float* __restrict__ s_dirs,
                                                                float* __restrict__ s_z, int64_t n_rays) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t ray = warp0; ray < n_rays; ray += n_warps) {
        const int2 sv = __ldg(reinterpret_cast<const int2*>(se_virtual) + ray);
        const int cnt = sv.y - sv.x;
        if (cnt <= 0) {
            if (lane == 0) reinterpret_cast<int2*>(se_out)[ray] = make_int2(-1, -1);
            continue;
        }
        const int64_t dst = __ldg(out_start + ray), slot0 = ray * (int64_t)max_nr;
        if (lane == 0) reinterpret_cast<int2*>(se_out)[ray] = make_int2((int)dst, (int)dst + cnt);
        const float ox = __ldg(rays_o + 3 * ray), oy = __ldg(rays_o + 3 * ray + 1), oz = __ldg(rays_o + 3 * ray + 2);
        const float dx = __ldg(rays_d + 3 * ray), dy = __ldg(rays_d + 3 * ray + 1), dz = __ldg(rays_d + 3 * ray + 2);
        for (int i = lane; i < cnt; i += 32) {
            const float t = __ldcs(z_stage + slot0 + i);
            const int64_t o = dst + i;
            __stcs(s_idx + o, (int32_t)(slot0 + i));
            __stcs(s_z + o, t);
            __stcs(s_3d + 3 * o, __fmaf_rn(t, dx, ox));
            __stcs(s_3d + 3 * o + 1, __fmaf_rn(t, dy, oy));
            __stcs(s_3d + 3 * o + 2, __fmaf_rn(t, dz, oz));
            __stcs(s_dirs + 3 * o, dx);
            __stcs(s_dirs + 3 * o + 1, dy);
            __stcs(s_dirs + 3 * o + 2, dz);
        }
    }
}

// ---- background sampler: nr_samples_per_ray samples per ray, uniform in inverse depth (RaySamplerGPU.cuh:39-139) -----------------------
__global__ void __launch_bounds__(128) sampler_bg_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                         const float* __restrict__ t_start_p, float t_far, int nr, Pcg rng, int jitter,
                                                         float* __restrict__ ray_max_dt, float* __restrict__ s_3d, float* __restrict__ s_dirs,
                                                         float* __restrict__ s_z, int32_t* __restrict__ se, int64_t n_rays) {
    const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    const float eps = 1e-6f;
    const float t_start = __ldg(t_start_p + ray);
    const float ox = __ldg(rays_o + 3 * ray), oy = __ldg(rays_o + 3 * ray + 1), oz = __ldg(rays_o + 3 * ray + 2);
    const float dx = __ldg(rays_d + 3 * ray), dy = __ldg(rays_d + 3 * ray + 1), dz = __ldg(rays_d + 3 * ray + 2);
    const float delta_s = (float)(1.0 / (double)(nr - 1));
    float max_dt = 0.f, s = 1.f, t_prec = t_start;
    for (int i = 0; i < nr; ++i) {
        float t = (float)(1.0 / (double)__fadd_rn(s, eps) - 1.0);
        t = __fadd_rn(t, t_start);
        t = clampf(t, t_start, t_far);
        if (jitter && i != 0 && i != nr - 1) {
            rng.advance((uint64_t)ray);
            const float interp = rng.next_float();
            t = __fmaf_rn(interp, __fsub_rn(t, t_prec), t_prec);
        }
        const int64_t o = ray * nr + i;
        s_z[o] = t;
        s_3d[3 * o] = __fmaf_rn(t, dx, ox);
        s_3d[3 * o + 1] = __fmaf_rn(t, dy, oy);
        s_3d[3 * o + 2] = __fmaf_rn(t, dz, oz);
        s_dirs[3 * o] = dx;
        s_dirs[3 * o + 1] = dy;
        s_dirs[3 * o + 2] = dz;
        s = __fsub_rn(s, delta_s);
        max_dt = fmaxf(max_dt, __fsub_rn(t, t_prec));
        t_prec = t;
    }
    ray_max_dt[ray] = max_dt;
    reinterpret_cast<int2*>(se)[ray] = make_int2((int)(ray * nr), (int)(ray * nr + nr));
}

// ---- scene contraction of background samples (RaySamplerGPU.cuh:528-658) -----------------------------------------------------------
// x -> (2 - 1/|2x|) x / |2x| for |2x| > 1 (and its inverse), depth re-measured from the camera.  Pointwise per sample; a warp walks a
// ray's samples with unit stride.  length() is sqrt of the dot product as nvcc contracts it in the reference (read off its SASS): fma(z,z, fma(x,x, y*y)).
__device__ __forceinline__ float length3(float x, float y, float z) { return __fsqrt_rn(__fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)))); }

template <bool UNCONTRACT>
__global__ void __launch_bounds__(256, 6) sampler_contract_kernel(const float* __restrict__ ray_o, const int32_t* __restrict__ se,
                                                               const float* __restrict__ s_3d, const float* __restrict__ s_z,
                                                               float* __restrict__ out_3d, float* __restrict__ out_z, int64_t n_rays) {
    __shared__ float tiles[8][96];
    const int lane = threadIdx.x & 31;
    float* tile = tiles[threadIdx.x >> 5];
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // the next ray's segment and origin are fetched while the current ray's samples are in flight (one DRAM round trip per ray, not two)
    int start_nx = 0, n_nx = 0;
    float cx_nx = 0.f, cy_nx = 0.f, cz_nx = 0.f;
    if (warp0 < n_rays) {
        n_nx = load_segment(se, warp0, start_nx);
        cx_nx = __ldg(ray_o + 3 * warp0), cy_nx = __ldg(ray_o + 3 * warp0 + 1), cz_nx = __ldg(ray_o + 3 * warp0 + 2);
    }
    for (int64_t ray = warp0; ray < n_rays; ray += n_warps) {
        const int start = start_nx, n = n_nx;
        const float cx = cx_nx, cy = cy_nx, cz = cz_nx;
        const int64_t nx = ray + n_warps;
        if (nx < n_rays) {
            n_nx = load_segment(se, nx, start_nx);
            cx_nx = __ldg(ray_o + 3 * nx), cy_nx = __ldg(ray_o + 3 * nx + 1), cz_nx = __ldg(ray_o + 3 * nx + 2);
        }
        for (int base = 0; base < n; base += 32) {
            // 32 samples = 96 consecutive floats of the [S,3] array: moved with unit stride, transposed through shared memory (reads at
            // stride 3 are conflict-free)
            const int cnt = min(32, n - base), i = base + lane;
            const int64_t s0 = (int64_t)start + base, s = s0 + lane;
            const float* src = s_3d + 3 * s0;
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (lane + 32 * k < 3 * cnt) tile[lane + 32 * k] = __ldcs(src + lane + 32 * k);
            float zz = i < n ? __ldcs(s_z + s) : 0.f;
            __syncwarp();
            float px = tile[3 * lane], py = tile[3 * lane + 1], pz = tile[3 * lane + 2];
            const float norm = length3(__fmul_rn(px, 2.f), __fmul_rn(py, 2.f), __fmul_rn(pz, 2.f));
            if (i < n && norm > 1.0f) {
                const float factor = UNCONTRACT ? __fdiv_rn(1.0f, __fsub_rn(2.0f, norm)) : __fsub_rn(2.0f, __fdiv_rn(1.0f, norm));
                px = __fdiv_rn(__fmul_rn(factor, px), norm);
                py = __fdiv_rn(__fmul_rn(factor, py), norm);
                pz = __fdiv_rn(__fmul_rn(factor, pz), norm);
                zz = length3(__fsub_rn(px, cx), __fsub_rn(py, cy), __fsub_rn(pz, cz));
            }
            __syncwarp();
            tile[3 * lane] = px, tile[3 * lane + 1] = py, tile[3 * lane + 2] = pz;
            __syncwarp();
            float* dst = out_3d + 3 * s0;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (lane + 32 * k < 3 * cnt) __stcs(dst + lane + 32 * k, tile[lane + 32 * k]);
            if (i < n) __stcs(out_z + s, zz);
        }
    }
}

// ---- occupancy-grid maintenance (OccupancyGridGPU.cuh:31-196) ------------------------------------------------------------------------
// Morton inverse of one axis (occ_grid_helpers.h:45-53)
__device__ __forceinline__ uint32_t compact_bits(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}
// lin_idx_to_3D (occ_grid_helpers.h:74-113) for one axis: index / n [- 0.5] [+ half a voxel], times the extent
__device__ __forceinline__ float voxel_axis(uint32_t c, int n, float extent, bool centre_grid, bool centre_of_voxel) {
    float x = __fdiv_rn((float)c, (float)n);
    if (centre_grid) x = __fsub_rn(x, 0.5f);
    if (centre_of_voxel) x = __fadd_rn(x, (float)(1.0 / (double)n) * 0.5f);
    return __fmul_rn(x, extent);
}

// get_grid_lower_left_voxels_vertices_gpu (centre = 0) / get_grid_samples_gpu (centre = 1, optional jitter inside the voxel)
__global__ void __launch_bounds__(256) occgrid_points_kernel(const int32_t* __restrict__ point_indices, int n, float ex, float ey, float ez,
                                                             int centre, Pcg rng, int jitter, float* __restrict__ out, int64_t n_points) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_points) return;
    const uint32_t v = (uint32_t)__ldg(point_indices + idx);
    float px = voxel_axis(compact_bits(v), n, ex, true, centre != 0);
    float py = voxel_axis(compact_bits(v >> 1), n, ey, true, centre != 0);
    float pz = voxel_axis(compact_bits(v >> 2), n, ez, true, centre != 0);
    if (centre && jitter) {
        const float nf = (float)n;
        const float sx = __fdiv_rn(ex, nf), sy = __fdiv_rn(ey, nf), sz = __fdiv_rn(ez, nf);
        rng.advance((uint64_t)(int64_t)((int)idx * 3));
        px = __fadd_rn(px, __fmaf_rn(sx, rng.next_float(), -(sx * 0.5f)));  // voxel_size * rand - half_voxel_size, contracted by nvcc
        py = __fadd_rn(py, __fmaf_rn(sy, rng.next_float(), -(sy * 0.5f)));
        pz = __fadd_rn(pz, __fmaf_rn(sz, rng.next_float(), -(sz * 0.5f)));
    }
    out[3 * idx] = px, out[3 * idx + 1] = py, out[3 * idx + 2] = pz;
}

// update_grid_values_gpu: grid[v] = max(new, grid[v] * decay) (duplicate indices race exactly as in the reference)
__global__ void __launch_bounds__(256) occgrid_update_values_kernel(const int32_t* __restrict__ point_indices, const float* __restrict__ values,
                                                                    float decay, float* __restrict__ grid_values, int64_t n_points) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_points) return;
    const int v = __ldg(point_indices + idx);
    grid_values[v] = fmaxf(__ldg(values + idx), __fmul_rn(grid_values[v], decay));
}

// update_grid_occupancy_with_density_values_gpu: occupied unless the voxel (or, with check_neighbours, its whole 3x3x3 neighbourhood)
// is at or below the threshold.  The neighbourhood is addressed as the reference does: the voxel's corner in UNIT-grid coordinates
// times the extent times n (so it is the voxel's integer coordinate only for a unit extent).
__global__ void __launch_bounds__(256) occgrid_update_occupancy_kernel(const int32_t* __restrict__ point_indices, int n, float ex, float ey,
                                                                       float ez, float thresh, int check_neighbours,
                                                                       const float* __restrict__ grid_values, uint8_t* __restrict__ occupancy,
                                                                       int64_t n_points) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_points) return;
    const int v = __ldg(point_indices + idx);
    bool is_empty = true;
    if (check_neighbours) {
        const float nf = (float)n, top = (float)(n - 1);
        const float px = __fmul_rn(voxel_axis(compact_bits((uint32_t)v), n, ex, false, false), nf);
        const float py = __fmul_rn(voxel_axis(compact_bits((uint32_t)v >> 1), n, ey, false, false), nf);
        const float pz = __fmul_rn(voxel_axis(compact_bits((uint32_t)v >> 2), n, ez, false, false), nf);
        for (int i = -1; i <= 1; ++i) {
            const float qx = __fadd_rn(px, (float)i);
            if (qx < 0.f || qx > top) continue;
            for (int j = -1; j <= 1; ++j) {
                const float qy = __fadd_rn(py, (float)j);
                if (qy < 0.f || qy > top) continue;
                for (int k = -1; k <= 1; ++k) {
                    const float qz = __fadd_rn(pz, (float)k);
                    if (qz < 0.f || qz > top) continue;
                    const uint32_t nb = spread_bits64(__float2uint_rz(qx)) | (spread_bits64(__float2uint_rz(qy)) << 1) |
                                        (spread_bits64(__float2uint_rz(qz)) << 2);
                    is_empty = is_empty && __ldg(grid_values + nb) <= thresh;
                }
            }
        }
    } else {
        is_empty = __ldg(grid_values + v) <= thresh;
    }
    occupancy[v] = is_empty ? 0 : 1;
}

// update_grid_occupancy_with_sdf_values_gpu (OccupancyGridGPU.cuh:220-316): a voxel stays occupied when the logistic density
// beta e^(-beta d) / (1 + e^(-beta d))^2 at the smallest distance d the surface can have inside the voxel (|sdf| minus half the voxel
// diagonal, at least 0) exceeds the threshold.  expf / powf / sqrtf as in the reference, so the same libdevice code decides.
__global__ void __launch_bounds__(256) occgrid_update_occupancy_sdf_kernel(const int32_t* __restrict__ point_indices, int n, float ex, float ey,
                                                                           float ez, const float* __restrict__ logistic_beta, float thresh,
                                                                           const float* __restrict__ grid_values,
                                                                           uint8_t* __restrict__ occupancy, int64_t n_points) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_points) return;
    const int v = __ldg(point_indices + idx);
    const float df = fabsf(__ldg(grid_values + v));
    const float nf = (float)n;
    const float sx = __fdiv_rn(ex, nf), sy = __fdiv_rn(ey, nf), sz = __fdiv_rn(ez, nf);
    // max_distance_in_cuboid: the longest vertex-to-vertex distance of the voxel is its diagonal
    const float diagonal = sqrtf(__fadd_rn(__fadd_rn(powf(sx, 2), powf(sy, 2)), powf(sz, 2)));
    const float d = fmaxf(0.0f, fminf(__fsub_rn(df, diagonal * 0.5f), 1e10f));
    const float beta = __ldg(logistic_beta + idx);
    const float e = fmaxf(-1e6f, fminf(expf(__fmul_rn(-beta, d)), 1e6f));
    const float weight = __fdiv_rn(__fmul_rn(beta, e), powf(__fadd_rn(1.0f, e), 2));
    occupancy[v] = weight > thresh ? 1 : 0;
}

// get_first_rays_sample_start_of_grid_occupied_regions_gpu (OccupancyGridGPU.cuh:505-582): the sphere tracer's starting point — the
// position at which each ray first probes an occupied voxel of the region of interest (one sample per hit ray at row = ray index, depth =
// the t AFTER that voxel's step, as the reference stores it); rays without one get the segment (0,0) and keep the packet's fill values
__global__ void __launch_bounds__(128) occgrid_first_sample_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                                   const float* __restrict__ t_entry, const float* __restrict__ t_exit_p,
                                                                   Grid g, float* __restrict__ s_3d, float* __restrict__ s_dirs,
                                                                   float* __restrict__ s_z, float* __restrict__ s_dt, int32_t* __restrict__ se,
                                                                   int64_t n_rays) {
    const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    const float t_exit = __ldg(t_exit_p + ray);
    const float ox = __ldg(rays_o + 3 * ray), oy = __ldg(rays_o + 3 * ray + 1), oz = __ldg(rays_o + 3 * ray + 2);
    const float dx = __ldg(rays_d + 3 * ray), dy = __ldg(rays_d + 3 * ray + 1), dz = __ldg(rays_d + 3 * ray + 2);
    float t = __ldg(t_entry + ray);
    for (int guard = 0; t < t_exit && guard < kMaxMarchSteps; ++guard) {
        const Probe p = probe(t, ox, oy, oz, dx, dy, dz, g);
        if (!in_grid(p.voxel, g)) break;
        t = __fadd_rn(__fadd_rn(t, step_from_unit(p.ux, p.uy, p.uz, dx, dy, dz, g)), 1e-6f);
        if (occupied(p.voxel, g)) {
            reinterpret_cast<int2*>(se)[ray] = make_int2((int)ray, (int)ray + 1);
            s_3d[3 * ray] = p.px, s_3d[3 * ray + 1] = p.py, s_3d[3 * ray + 2] = p.pz;
            s_dirs[3 * ray] = dx, s_dirs[3 * ray + 1] = dy, s_dirs[3 * ray + 2] = dz;
            s_z[ray] = t;
            s_dt[ray] = 0.f;
            return;
        }
    }
    reinterpret_cast<int2*>(se)[ray] = make_int2(0, 0);
}

// advance_ray_sample_to_next_occupied_voxel_gpu (OccupancyGridGPU.cuh:443-503): each point marches along its direction to the first
// occupied voxel of the roi (position where that voxel was probed) or, when it leaves the grid, to the last position probed inside it
// (within = false).  out may alias points (the reference writes in place).
__global__ void __launch_bounds__(128) occgrid_advance_kernel(const float* __restrict__ dirs, const float* points, Grid g, float* out,
                                                              uint8_t* __restrict__ within, int64_t n_points) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_points) return;
    const float ox = points[3 * i], oy = points[3 * i + 1], oz = points[3 * i + 2];
    const float dx = __ldg(dirs + 3 * i), dy = __ldg(dirs + 3 * i + 1), dz = __ldg(dirs + 3 * i + 2);
    float t = 0.f, prec_t = 0.f;
    bool inside = true;
    for (int guard = 0;; ++guard) {
        const Probe p = probe(t, ox, oy, oz, dx, dy, dz, g);
        // Deviation: the reference's index clamps coordinates below the grid (and NaN) to voxel 0 of that axis, so a point leaving through
        // a lower face is marched on for ever (its kernel does not return; the call site, utils/sphere_tracing.py:131, is commented
        // out).  Here a position below the grid ends the march like one above it.
        const bool below = !(__fadd_rn(p.ux, 0.5f) >= 0.f) || !(__fadd_rn(p.uy, 0.5f) >= 0.f) || !(__fadd_rn(p.uz, 0.5f) >= 0.f);
        if (below || !in_grid(p.voxel, g) || guard >= kMaxMarchSteps) {  // the cap: t stopped advancing in fp32 (huge |origin|, tiny step)
            inside = false;
            out[3 * i] = __fmaf_rn(prec_t, dx, ox), out[3 * i + 1] = __fmaf_rn(prec_t, dy, oy), out[3 * i + 2] = __fmaf_rn(prec_t, dz, oz);
            break;
        }
        prec_t = t;
        t = __fadd_rn(__fadd_rn(t, step_from_unit(p.ux, p.uy, p.uz, dx, dy, dz, g)), 1e-6f);
        if (occupied(p.voxel, g)) {
            out[3 * i] = p.px, out[3 * i + 1] = p.py, out[3 * i + 2] = p.pz;
            break;
        }
    }
    within[i] = inside ? 1 : 0;
}

// ---- occupancy-grid queries ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) occgrid_t_near_t_far_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                                   const float* __restrict__ t_entry, const float* __restrict__ t_exit_p,
                                                                   Grid g, float* __restrict__ t_near, float* __restrict__ t_far,
                                                                   int64_t n_rays) {
    const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    const float t_start = __ldg(t_entry + ray), t_exit = __ldg(t_exit_p + ray);
    const float ox = __ldg(rays_o + 3 * ray), oy = __ldg(rays_o + 3 * ray + 1), oz = __ldg(rays_o + 3 * ray + 2);
    const float dx = __ldg(rays_d + 3 * ray), dy = __ldg(rays_d + 3 * ray + 1), dz = __ldg(rays_d + 3 * ray + 2);
    float near = t_start, far = t_start, t = t_start;
    bool first = true;
    for (int guard = 0; t < t_exit && guard < kMaxMarchSteps; ++guard) {
        const Probe p = probe(t, ox, oy, oz, dx, dy, dz, g);
        if (!in_grid(p.voxel, g)) break;
        const bool occ = occupied(p.voxel, g);
        if (occ && first) {
            near = t;
            first = false;
        }
        t = __fadd_rn(t, step_from_unit(p.ux, p.uy, p.uz, dx, dy, dz, g));
        if (occ) far = clampf(t, t_start, t_exit);
    }
    t_near[ray] = near;
    t_far[ray] = far;
}

__global__ void __launch_bounds__(256) occgrid_check_kernel(const float* __restrict__ points, Grid g, const float* __restrict__ values,
                                                            uint8_t* __restrict__ out_occ, float* __restrict__ out_val, int64_t n_points) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_points) return;
    const int v = pos_to_lin_idx(__ldg(points + 3 * i), __ldg(points + 3 * i + 1), __ldg(points + 3 * i + 2), g);
    if (in_grid(v, g)) {
        out_occ[i] = occupied(v, g) ? 1 : 0;
        out_val[i] = __ldg(values + v);
    } else {
        out_occ[i] = 0;
        out_val[i] = 0.f;
    }
}

}  // namespace vs

using namespace vs;

extern "C" {

static int make_grid(int nr_voxels_per_dim, const float* extent, const uint8_t* occ, const uint8_t* roi, Grid* g) {
    if (nr_voxels_per_dim < 1 || nr_voxels_per_dim > 1024 || !extent || !occ) return VS_ERR_INVALID_ARG;
    if (!(extent[0] > 0.f && extent[1] > 0.f && extent[2] > 0.f)) return VS_ERR_INVALID_ARG;
    const bool pow2 = (nr_voxels_per_dim & (nr_voxels_per_dim - 1)) == 0;
    *g = Grid{nr_voxels_per_dim, extent[0], extent[1], extent[2], pow2 ? 1.0f / (float)nr_voxels_per_dim : 0.f, occ, roi};
    return VS_OK;
}

// Pass 1 of compute_samples_fg (nr_voxels_per_dim == 0: no grid) / compute_samples_fg_in_grid_occupied_regions.
//   rays_o, rays_d [n,3], t_entry, t_exit [n,1] f32 · extent: HOST float[3] · occupancy, roi: DEVICE u8 [nr_voxels_per_dim^3] (torch.bool);
//   roi == NULL: `occupancy` already holds occupancy && roi
//   se_virtual [n,2] i32: the segment each ray WOULD own in the reference's uncompacted packet, (-1,-1) for rays without samples
//   ray_max_dt [n,1] f32: pre-filled -1 by the caller (constructor fill), written for rays with samples
//   z_stage [n * max_nr] f32: depth of sample i of ray r at r*max_nr + i (slots of real samples only; no initialisation needed)
int vs_sampler_fg_count(const float* rays_o, const float* rays_d, const float* t_entry, const float* t_exit, float min_dist, int min_nr,
                        int max_nr, uint64_t rng_state, uint64_t rng_inc, int jitter, int nr_voxels_per_dim, const float* extent,
                        const uint8_t* occupancy, const uint8_t* roi, int32_t* se_virtual, float* ray_max_dt, float* z_stage, int64_t n_rays,
                        void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && max_nr >= 0 && min_nr >= 0 && n_rays * (int64_t)max_nr <= 0x7fffffffLL);
    if (n_rays == 0) return VS_OK;
    VS_CHECK_ARG(rays_o && rays_d && t_entry && t_exit && se_virtual && ray_max_dt && (z_stage || max_nr == 0));
    Grid g{0, 1.f, 1.f, 1.f, 0.f, nullptr, nullptr};
    const Pcg rng{rng_state, rng_inc};
    const unsigned grid = (unsigned)div_up(n_rays, 128);
    if (nr_voxels_per_dim > 0) {
        int e = make_grid(nr_voxels_per_dim, extent, occupancy, roi, &g);
        if (e != VS_OK) return e;
        sampler_fg_march_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, t_entry, t_exit, min_dist, min_nr, max_nr, rng, jitter,
                                                                               g, se_virtual, ray_max_dt, z_stage, n_rays);
    } else {
        sampler_fg_march_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, t_entry, t_exit, min_dist, min_nr, max_nr, rng,
                                                                                jitter, g, se_virtual, ray_max_dt, z_stage, n_rays);
    }
    return launched(1);
}

// Pass 2: out_start [n] i32 from vs_segment_offsets(se_virtual); writes se_out [n,2] and, for every sample, samples_idx (the slot of the
// reference's uncompacted packet), samples_3d = o + z d, samples_dirs, samples_z at its compacted position.
int vs_sampler_fg_write(const float* rays_o, const float* rays_d, int max_nr, const int32_t* se_virtual, const float* z_stage,
                        const int32_t* out_start, int32_t* se_out, int32_t* samples_idx, float* samples_3d, float* samples_dirs,
                        float* samples_z, int64_t n_rays, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && max_nr >= 0);
    if (n_rays == 0) return VS_OK;
    VS_CHECK_ARG(rays_o && rays_d && se_virtual && out_start && se_out);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)std::min<int64_t>(div_up(n_rays, 8), (int64_t)sms * 8);
    sampler_fg_expand_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, max_nr, se_virtual, z_stage, out_start, se_out, samples_idx,
                                                                     samples_3d, samples_dirs, samples_z, n_rays);
    return launched(1);
}

// compute_samples_bg: samples_* are [n_rays * nr_samples_per_ray, .] (always compacted), ray_max_dt [n,1], se [n,2]
int vs_sampler_bg(const float* rays_o, const float* rays_d, const float* t_start, float t_far, int nr_samples_per_ray, uint64_t rng_state,
                  uint64_t rng_inc, int jitter, float* ray_max_dt, float* samples_3d, float* samples_dirs, float* samples_z, int32_t* se,
                  int64_t n_rays, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && nr_samples_per_ray >= 1 && n_rays * (int64_t)nr_samples_per_ray <= 0x7fffffffLL);
    if (n_rays == 0) return VS_OK;
    VS_CHECK_ARG(rays_o && rays_d && t_start && ray_max_dt && samples_3d && samples_dirs && samples_z && se);
    sampler_bg_kernel<<<(unsigned)div_up(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, t_start, t_far, nr_samples_per_ray,
                                                                                       Pcg{rng_state, rng_inc}, jitter, ray_max_dt, samples_3d,
                                                                                       samples_dirs, samples_z, se, n_rays);
    return launched(1);
}

// RaySampler::contract_samples / uncontract_samples without the closing update_dt: samples_3d, samples_z of a compacted packet ->
// out_3d, out_z (same shapes; may alias the inputs)
int vs_sampler_contract(const float* ray_o, const int32_t* se, const float* samples_3d, const float* samples_z, float* out_3d, float* out_z,
                        int uncontract, int64_t n_rays, void* stream) {
    VS_CHECK_ARG(n_rays >= 0);
    if (n_rays == 0) return VS_OK;
    VS_CHECK_ARG(ray_o && se && samples_3d && samples_z && out_3d && out_z);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // persistent grid: exactly the blocks that are resident at once (a grid sized past residency runs a second, partly empty wave)
    static int resident[2] = {0, 0};
    if (!resident[uncontract != 0]) {
        int b = 0;
        cudaError_t e = uncontract ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, sampler_contract_kernel<true>, 256, 0)
                                   : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, sampler_contract_kernel<false>, 256, 0);
        resident[uncontract != 0] = (e == cudaSuccess && b > 0) ? b : 4;
    }
    const unsigned grid = (unsigned)std::min<int64_t>(div_up(n_rays, 8), (int64_t)sms * resident[uncontract != 0]);
    if (uncontract)
        sampler_contract_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(ray_o, se, samples_3d, samples_z, out_3d, out_z, n_rays);
    else
        sampler_contract_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(ray_o, se, samples_3d, samples_z, out_3d, out_z, n_rays);
    return launched(1);
}

// OccupancyGrid::get_grid_lower_left_voxels_vertices (centre = 0; src/OccupancyGrid.cu:206-234) and get_grid_samples /
// get_random_grid_samples[_in_roi] (centre = 1; :236-347): world position of voxels point_indices [n] (Morton order) -> out [n,3]
int vs_occgrid_points(const int32_t* point_indices, int nr_voxels_per_dim, const float* extent, int centre, uint64_t rng_state, uint64_t rng_inc,
                      int jitter, float* out, int64_t n_points, void* stream) {
    VS_CHECK_ARG(n_points >= 0 && nr_voxels_per_dim > 0);
    if (n_points == 0) return VS_OK;
    VS_CHECK_ARG(point_indices && extent && out);
    Pcg rng{rng_state, rng_inc};
    occgrid_points_kernel<<<(unsigned)div_up(n_points, 256), 256, 0, (cudaStream_t)stream>>>(point_indices, nr_voxels_per_dim, extent[0], extent[1],
                                                                                           extent[2], centre, rng, jitter, out, n_points);
    return launched(1);
}

// OccupancyGrid::update_grid_values (src/OccupancyGrid.cu:446-474): grid_values[point_indices[i]] = max(values[i], decay * old)
int vs_occgrid_update_values(const int32_t* point_indices, const float* values, float decay, float* grid_values, int64_t n_points, void* stream) {
    VS_CHECK_ARG(n_points >= 0);
    if (n_points == 0) return VS_OK;
    VS_CHECK_ARG(point_indices && values && grid_values);
    occgrid_update_values_kernel<<<(unsigned)div_up(n_points, 256), 256, 0, (cudaStream_t)stream>>>(point_indices, values, decay, grid_values,
                                                                                                  n_points);
    return launched(1);
}

// OccupancyGrid::update_grid_occupancy_with_density_values (src/OccupancyGrid.cu:476-503)
int vs_occgrid_update_occupancy_density(const int32_t* point_indices, int nr_voxels_per_dim, const float* extent, float occupancy_thresh,
                                        int check_neighbours, const float* grid_values, uint8_t* occupancy, int64_t n_points, void* stream) {
    VS_CHECK_ARG(n_points >= 0 && nr_voxels_per_dim > 0);
    if (n_points == 0) return VS_OK;
    VS_CHECK_ARG(point_indices && extent && grid_values && occupancy);
    occgrid_update_occupancy_kernel<<<(unsigned)div_up(n_points, 256), 256, 0, (cudaStream_t)stream>>>(
        point_indices, nr_voxels_per_dim, extent[0], extent[1], extent[2], occupancy_thresh, check_neighbours, grid_values, occupancy, n_points);
    return launched(1);
}

// OccupancyGrid::update_grid_occupancy_with_sdf_values (src/OccupancyGrid.cu:505-533); logistic_beta [n_points,1]
int vs_occgrid_update_occupancy_sdf(const int32_t* point_indices, int nr_voxels_per_dim, const float* extent, const float* logistic_beta,
                                    float occupancy_thresh, const float* grid_values, uint8_t* occupancy, int64_t n_points, void* stream) {
    VS_CHECK_ARG(n_points >= 0 && nr_voxels_per_dim > 0);
    if (n_points == 0) return VS_OK;
    VS_CHECK_ARG(point_indices && extent && logistic_beta && grid_values && occupancy);
    occgrid_update_occupancy_sdf_kernel<<<(unsigned)div_up(n_points, 256), 256, 0, (cudaStream_t)stream>>>(
        point_indices, nr_voxels_per_dim, extent[0], extent[1], extent[2], logistic_beta, occupancy_thresh, grid_values, occupancy, n_points);
    return launched(1);
}

// OccupancyGrid::get_rays_t_near_t_far: first / last t inside occupied voxels of the region of interest along every ray
int vs_occgrid_rays_t_near_t_far(const float* rays_o, const float* rays_d, const float* t_entry, const float* t_exit, int nr_voxels_per_dim,
                                 const float* extent, const uint8_t* occupancy, const uint8_t* roi, float* t_near, float* t_far, int64_t n_rays,
                                 void* stream) {
    VS_CHECK_ARG(n_rays >= 0);
    Grid g;
    int e = make_grid(nr_voxels_per_dim, extent, occupancy, roi, &g);
    if (e != VS_OK) return e;
    if (n_rays == 0) return VS_OK;
    VS_CHECK_ARG(rays_o && rays_d && t_entry && t_exit && t_near && t_far);
    occgrid_t_near_t_far_kernel<<<(unsigned)div_up(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, t_entry, t_exit, g, t_near, t_far,
                                                                                                 n_rays);
    return launched(1);
}

// OccupancyGrid::get_first_rays_sample_start_of_grid_occupied_regions (src/OccupancyGrid.cu:536-573): one-sample-per-ray packet of the
// first occupied voxel each ray probes; the caller presets the packet's fill values (rows of rays without a hit are left untouched)
int vs_occgrid_first_sample_start(const float* rays_o, const float* rays_d, const float* t_entry, const float* t_exit, int nr_voxels_per_dim,
                                  const float* extent, const uint8_t* occupancy, const uint8_t* roi, float* samples_3d, float* samples_dirs,
                                  float* samples_z, float* samples_dt, int32_t* se, int64_t n_rays, void* stream) {
    VS_CHECK_ARG(n_rays >= 0);
    Grid g;
    int e = make_grid(nr_voxels_per_dim, extent, occupancy, roi, &g);
    if (e != VS_OK) return e;
    if (n_rays == 0) return VS_OK;
    VS_CHECK_ARG(rays_o && rays_d && t_entry && t_exit && samples_3d && samples_dirs && samples_z && samples_dt && se);
    occgrid_first_sample_kernel<<<(unsigned)div_up(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, t_entry, t_exit, g, samples_3d,
                                                                                                 samples_dirs, samples_z, samples_dt, se, n_rays);
    return launched(1);
}

// OccupancyGrid::advance_ray_sample_to_next_occupied_voxel (src/OccupancyGrid.cu:575-607); new_samples_3d may alias samples_3d
int vs_occgrid_advance_to_next_occupied(const float* samples_dirs, const float* samples_3d, int nr_voxels_per_dim, const float* extent,
                                        const uint8_t* occupancy, const uint8_t* roi, float* new_samples_3d, uint8_t* is_within_bounds,
                                        int64_t n_points, void* stream) {
    VS_CHECK_ARG(n_points >= 0);
    Grid g;
    int e = make_grid(nr_voxels_per_dim, extent, occupancy, roi, &g);
    if (e != VS_OK) return e;
    if (n_points == 0) return VS_OK;
    VS_CHECK_ARG(samples_dirs && samples_3d && new_samples_3d && is_within_bounds);
    occgrid_advance_kernel<<<(unsigned)div_up(n_points, 128), 128, 0, (cudaStream_t)stream>>>(samples_dirs, samples_3d, g, new_samples_3d,
                                                                                            is_within_bounds, n_points);
    return launched(1);
}

// OccupancyGrid::check_occupancy: per point (occupied && in roi, grid value); points outside the grid -> (false, 0)
int vs_occgrid_check_occupancy(const float* points, int nr_voxels_per_dim, const float* extent, const float* values, const uint8_t* occupancy,
                               const uint8_t* roi, uint8_t* out_occupancy, float* out_values, int64_t n_points, void* stream) {
    VS_CHECK_ARG(n_points >= 0);
    Grid g;
    int e = make_grid(nr_voxels_per_dim, extent, occupancy, roi, &g);
    if (e != VS_OK) return e;
    if (n_points == 0) return VS_OK;
    VS_CHECK_ARG(points && values && out_occupancy && out_values);
    occgrid_check_kernel<<<(unsigned)div_up(n_points, 256), 256, 0, (cudaStream_t)stream>>>(points, g, values, out_occupancy, out_values, n_points);
    return launched(1);
}

}  // extern "C"
