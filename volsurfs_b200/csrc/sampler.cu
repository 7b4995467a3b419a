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
// The reference marches every ray into an UNCOMPACTED packet (nr_rays x max_nr_samples_per_ray slots: 23 GB of staging at 640k rays x
// 1024) and then gathers it (compact_to_valid_samples).  Here the foreground samplers run in two launches around one prefix sum:
//   sampler_fg_kernel<GRID, false>   per ray: length of the occupied stretch -> sample count and spacing (ray_max_dt), then a dry run
//                                    of the sampling march: the count the reference would have created (0 below min_nr_samples_per_ray)
//   (vs_segment_offsets)             exclusive scan of the counts -> compacted start of every ray + the total
//   sampler_fg_kernel<GRID, true>    the sampling march again, storing samples straight at their compacted position
// so the result equals the reference's compacted packet (samples_idx = the slot the sample would have had, samples_dt untouched) and
// nothing of size nr_rays x max_nr_samples_per_ray ever exists.  The marches are sequential per ray (every step depends on the
// position the previous one reached), so a thread owns a ray; the occupancy and region-of-interest masks (1 byte per voxel, Morton
// order) are L2 resident.
//
// Arithmetic contract (bit parity with the reference kernels as nvcc compiles them, pinned on the GPU against oracle/_ref/
// libsampler_ref.so): IEEE fp32 with the contractions nvcc applies to the reference source — `ray_o + t * ray_d`, `t + c * rnd` and
// helper_math's lerp are fused multiply-adds; everything else is separate round-to-nearest operations; the background sampler's
// `1.0 / (s + eps) - 1.0` is evaluated in double like the reference's double literals make it.
#include "vs_common.cuh"

namespace vs {

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
    const uint8_t* occ;
    const uint8_t* roi;
};

// 21-bit Morton spreading; the reference keeps only the low 32 bits of the 64-bit spread (uint32_t xx = expand_bits(x))
__device__ __forceinline__ uint32_t spread_bits(uint32_t v) {
    uint64_t w = v;
    w &= 0x00000000001fffffULL;
    w = (w | w << 32) & 0x001f00000000ffffULL;
    w = (w | w << 16) & 0x001f0000ff0000ffULL;
    w = (w | w << 8) & 0x010f00f00f00f00fULL;
    w = (w | w << 4) & 0x10c30c30c30c30c3ULL;
    w = (w | w << 2) & 0x1249249249249249ULL;
    return (uint32_t)w;
}

__device__ __forceinline__ int pos_to_lin_idx(float px, float py, float pz, const Grid& g) {
    const float n = (float)g.n;
    const float x = __fmul_rn(__fadd_rn(__fdiv_rn(px, g.ex), 0.5f), n);
    const float y = __fmul_rn(__fadd_rn(__fdiv_rn(py, g.ey), 0.5f), n);
    const float z = __fmul_rn(__fadd_rn(__fdiv_rn(pz, g.ez), 0.5f), n);
    // float -> uint32_t as the hardware converts it: truncation, negatives and NaN to 0, saturation above
    const uint32_t xx = spread_bits(__float2uint_rz(x)), yy = spread_bits(__float2uint_rz(y)), zz = spread_bits(__float2uint_rz(z));
    return (int)(xx | (yy << 1) | (zz << 2));
}

__device__ __forceinline__ int sign_of(float x) { return x > 0.f ? 1 : (x < 0.f ? -1 : 0); }

// the reference's "DDA like step": distance (measured along the AXIS, not along the ray) to the next voxel boundary, + 1e-6
__device__ __forceinline__ float distance_to_next_voxel(float px, float py, float pz, float dx, float dy, float dz, const Grid& g) {
    const float eps = 1e-6f, n = (float)g.n;
    if (fabsf(dx) < eps && fabsf(dy) < eps && fabsf(dz) < eps) return 1e10f;
    float t3[3];
    const float p[3] = {px, py, pz}, d[3] = {dx, dy, dz}, e[3] = {g.ex, g.ey, g.ez};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        t3[a] = 1e10f;
        if (fabsf(d[a]) > eps) {
            const float q = __fmul_rn(__fdiv_rn(p[a], e[a]), n);
            const float prime = floorf(__fadd_rn(q, (float)sign_of(d[a])));
            t3[a] = __fmul_rn(__fdiv_rn(fabsf(__fsub_rn(prime, q)), n), e[a]);
        }
    }
    return __fadd_rn(fminf(fminf(t3[0], t3[1]), t3[2]), eps);
}

__device__ __forceinline__ bool in_grid(int idx_voxel, const Grid& g) { return idx_voxel >= 0 && idx_voxel < g.n * g.n * g.n; }
__device__ __forceinline__ bool occupied(int idx_voxel, const Grid& g) { return __ldg(g.roi + idx_voxel) && __ldg(g.occ + idx_voxel); }
__device__ __forceinline__ float clampf(float v, float a, float b) { return fmaxf(a, fminf(b, v)); }

// ---- foreground samplers -----------------------------------------------------------------------------------------------------------
// WRITE = false: writes the ray's virtual uncompacted segment (ray*max_nr, ray*max_nr + created) or (-1,-1), ray_max_dt and n_create.
// WRITE = true : stores the samples of rays with a non-empty segment at out_start[ray] + i and the compacted segment.
template <bool GRID, bool WRITE>
__global__ void __launch_bounds__(128) sampler_fg_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                         const float* __restrict__ t_entry, const float* __restrict__ t_exit_p, float min_dist,
                                                         int min_nr, int max_nr, Pcg rng, int jitter, Grid g, int32_t* __restrict__ se_virtual,
                                                         float* __restrict__ ray_max_dt, int32_t* __restrict__ n_create,
                                                         const int32_t* __restrict__ out_start, int32_t* __restrict__ se_out,
                                                         int32_t* __restrict__ s_idx, float* __restrict__ s_3d, float* __restrict__ s_dirs,
                                                         float* __restrict__ s_z, int64_t n_rays) {
    const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    const float eps = 1e-6f;
    const float t_start = __ldg(t_entry + ray), t_exit = __ldg(t_exit_p + ray);
    const float ox = __ldg(rays_o + 3 * ray), oy = __ldg(rays_o + 3 * ray + 1), oz = __ldg(rays_o + 3 * ray + 2);
    const float dx = __ldg(rays_d + 3 * ray), dy = __ldg(rays_d + 3 * ray + 1), dz = __ldg(rays_d + 3 * ray + 2);

    int to_create = 0;
    float spacing = 0.f;
    int64_t dst = 0;
    if (WRITE) {
        const int2 sv = __ldg(reinterpret_cast<const int2*>(se_virtual) + ray);
        const int cnt = sv.y - sv.x;
        if (cnt <= 0) {
            reinterpret_cast<int2*>(se_out)[ray] = make_int2(-1, -1);
            return;
        }
        dst = __ldg(out_start + ray);
        reinterpret_cast<int2*>(se_out)[ray] = make_int2((int)dst, (int)dst + cnt);
        to_create = __ldg(n_create + ray);
        spacing = __ldg(ray_max_dt + ray);
    } else {
        // ---- how far does the ray travel through occupied space -> number of samples and their spacing
        float dist = 0.f;
        if (GRID) {
            float t = t_start, step = 0.f;
            while (t < t_exit) {
                const float px = __fmaf_rn(t, dx, ox), py = __fmaf_rn(t, dy, oy), pz = __fmaf_rn(t, dz, oz);
                const int v = pos_to_lin_idx(px, py, pz, g);
                if (!in_grid(v, g)) break;
                if (occupied(v, g)) dist = __fadd_rn(dist, step);  // the step that LED here (reference quirk: lags one voxel)
                step = distance_to_next_voxel(px, py, pz, dx, dy, dz, g);
                t = __fadd_rn(t, step);
            }
            dist = clampf(dist, 0.f, __fsub_rn(t_exit, t_start));
        } else {
            dist = __fsub_rn(t_exit, t_start);
        }
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
    }

    // ---- the sampling march (dry when !WRITE)
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
            const float px = __fmaf_rn(t, dx, ox), py = __fmaf_rn(t, dy, oy), pz = __fmaf_rn(t, dz, oz);
            if (created >= to_create) break;
            bool emit = true, occ = true;
            int v = 0;
            if (GRID) {
                v = pos_to_lin_idx(px, py, pz, g);
                if (!in_grid(v, g)) break;
                occ = occupied(v, g);
                emit = occ && to_next == 0.f;
            }
            if (emit) {
                if (WRITE) {
                    const int64_t o = dst + created;
                    s_idx[o] = (int32_t)(slot0 + created);
                    s_3d[3 * o] = px;
                    s_3d[3 * o + 1] = py;
                    s_3d[3 * o + 2] = pz;
                    s_dirs[3 * o] = dx;
                    s_dirs[3 * o + 1] = dy;
                    s_dirs[3 * o + 2] = dz;
                    s_z[o] = t;
                }
                ++created;
                if (GRID) to_next = spacing;
            }
            if (GRID) {
                const float to_voxel = distance_to_next_voxel(px, py, pz, dx, dy, dz, g);
                float step;
                if (occ) {
                    step = fminf(to_voxel, to_next);
                    to_next = __fsub_rn(to_next, step);
                    if (to_next <= eps) to_next = 0.f;
                } else {
                    step = to_voxel;
                }
                t = __fadd_rn(t, step);
            } else {
                t = __fadd_rn(t, spacing);
            }
        }
    }
    if (!WRITE) {
        // fewer than min_nr samples: the ray keeps (-1,-1) and ray_max_dt = -1 (the RaySamplesPacked constructor fill); otherwise its
        // spacing is recorded — also for a ray with zero samples when min_nr == 0 (reference behaviour, RaySamplerGPU.cuh:253-262)
        if (created >= min_nr) {
            ray_max_dt[ray] = spacing;
            reinterpret_cast<int2*>(se_virtual)[ray] = make_int2((int)slot0, (int)slot0 + created);
        } else {
            reinterpret_cast<int2*>(se_virtual)[ray] = make_int2(-1, -1);
        }
        n_create[ray] = to_create;
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
    while (t < t_exit) {
        const float px = __fmaf_rn(t, dx, ox), py = __fmaf_rn(t, dy, oy), pz = __fmaf_rn(t, dz, oz);
        const int v = pos_to_lin_idx(px, py, pz, g);
        if (!in_grid(v, g)) break;
        const bool occ = occupied(v, g);
        if (occ && first) {
            near = t;
            first = false;
        }
        t = __fadd_rn(t, distance_to_next_voxel(px, py, pz, dx, dy, dz, g));
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
    if (nr_voxels_per_dim < 1 || nr_voxels_per_dim > 1024 || !extent || !occ || !roi) return VS_ERR_INVALID_ARG;
    if (!(extent[0] > 0.f && extent[1] > 0.f && extent[2] > 0.f)) return VS_ERR_INVALID_ARG;
    *g = Grid{nr_voxels_per_dim, extent[0], extent[1], extent[2], occ, roi};
    return VS_OK;
}

// Pass 1 of compute_samples_fg (nr_voxels_per_dim == 0: no grid) / compute_samples_fg_in_grid_occupied_regions.
//   rays_o, rays_d [n,3], t_entry, t_exit [n,1] f32 · extent: HOST float[3] · occupancy, roi: DEVICE u8 [nr_voxels_per_dim^3] (torch.bool)
//   se_virtual [n,2] i32: the segment each ray WOULD own in the reference's uncompacted packet, (-1,-1) for rays without samples
//   ray_max_dt [n,1] f32: pre-filled -1 by the caller (constructor fill), written for rays with samples · n_create [n] i32 scratch
int vs_sampler_fg_count(const float* rays_o, const float* rays_d, const float* t_entry, const float* t_exit, float min_dist, int min_nr,
                        int max_nr, uint64_t rng_state, uint64_t rng_inc, int jitter, int nr_voxels_per_dim, const float* extent,
                        const uint8_t* occupancy, const uint8_t* roi, int32_t* se_virtual, float* ray_max_dt, int32_t* n_create, int64_t n_rays,
                        void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && max_nr >= 0 && min_nr >= 0 && n_rays * (int64_t)max_nr <= 0x7fffffffLL);
    if (n_rays == 0) return VS_OK;
    VS_CHECK_ARG(rays_o && rays_d && t_entry && t_exit && se_virtual && ray_max_dt && n_create);
    Grid g{0, 1.f, 1.f, 1.f, nullptr, nullptr};
    const Pcg rng{rng_state, rng_inc};
    const unsigned grid = (unsigned)div_up(n_rays, 128);
    if (nr_voxels_per_dim > 0) {
        int e = make_grid(nr_voxels_per_dim, extent, occupancy, roi, &g);
        if (e != VS_OK) return e;
        sampler_fg_kernel<true, false><<<grid, 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, t_entry, t_exit, min_dist, min_nr, max_nr, rng,
                                                                                jitter, g, se_virtual, ray_max_dt, n_create, nullptr, nullptr,
                                                                                nullptr, nullptr, nullptr, nullptr, n_rays);
    } else {
        sampler_fg_kernel<false, false><<<grid, 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, t_entry, t_exit, min_dist, min_nr, max_nr, rng,
                                                                                 jitter, g, se_virtual, ray_max_dt, n_create, nullptr, nullptr,
                                                                                 nullptr, nullptr, nullptr, nullptr, n_rays);
    }
    return launched(1);
}

// Pass 2: out_start [n] i32 from vs_segment_offsets(se_virtual); writes se_out [n,2] and, for every sample, samples_idx (the slot of the
// reference's uncompacted packet), samples_3d, samples_dirs, samples_z at its compacted position.
int vs_sampler_fg_write(const float* rays_o, const float* rays_d, const float* t_entry, const float* t_exit, float min_dist, int min_nr,
                        int max_nr, uint64_t rng_state, uint64_t rng_inc, int jitter, int nr_voxels_per_dim, const float* extent,
                        const uint8_t* occupancy, const uint8_t* roi, const int32_t* se_virtual, const float* ray_max_dt,
                        const int32_t* n_create, const int32_t* out_start, int32_t* se_out, int32_t* samples_idx, float* samples_3d,
                        float* samples_dirs, float* samples_z, int64_t n_rays, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && max_nr >= 0 && min_nr >= 0);
    if (n_rays == 0) return VS_OK;
    VS_CHECK_ARG(rays_o && rays_d && t_entry && t_exit && se_virtual && ray_max_dt && n_create && out_start && se_out);
    Grid g{0, 1.f, 1.f, 1.f, nullptr, nullptr};
    const Pcg rng{rng_state, rng_inc};
    const unsigned grid = (unsigned)div_up(n_rays, 128);
    int32_t* sv = const_cast<int32_t*>(se_virtual);
    float* md = const_cast<float*>(ray_max_dt);
    int32_t* nc = const_cast<int32_t*>(n_create);
    if (nr_voxels_per_dim > 0) {
        int e = make_grid(nr_voxels_per_dim, extent, occupancy, roi, &g);
        if (e != VS_OK) return e;
        sampler_fg_kernel<true, true><<<grid, 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, t_entry, t_exit, min_dist, min_nr, max_nr, rng,
                                                                               jitter, g, sv, md, nc, out_start, se_out, samples_idx, samples_3d,
                                                                               samples_dirs, samples_z, n_rays);
    } else {
        sampler_fg_kernel<false, true><<<grid, 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, t_entry, t_exit, min_dist, min_nr, max_nr, rng,
                                                                                jitter, g, sv, md, nc, out_start, se_out, samples_idx,
                                                                                samples_3d, samples_dirs, samples_z, n_rays);
    }
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
