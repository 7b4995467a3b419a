// volsurfs_b200 — importance sampling chain on packed ray samples.
//
// Replaces
//   VolumeRendering::importance_sample            src/VolumeRendering.cu:466-548, kernel kernels/volsurfs/VolumeRenderingGPU.cuh:507-678
//   VolumeRendering::combine_ray_samples_packets  src/VolumeRendering.cu:550-669, kernel kernels/volsurfs/VolumeRenderingGPU.cuh:680-894
// (callers: volsurfs_py/utils/nerf_utils.py:10-92, utils/sdf_utils.py:97-179).  Both kernels produce the UNCOMPACTED packet the
// reference produces before its compact_to_valid_samples(); the host shim compacts with vs_compact_*.
//
// importance_sample: the reference walks a ray's n_imp samples in one thread (lane r touches rows r*n_imp+i: stride n_imp*12
// bytes between lanes).  Here one thread owns one (ray, i) pair, so a warp writes 32 consecutive sample rows; the pcg32 stream
// position the reference reaches sequentially (advance(ray) before every draw, VolumeRenderingGPU.cuh:583) is computed directly
// with the O(log n) jump, so the jittered samples are the same numbers.  The arithmetic keeps the reference's expression shapes
// (double literals folded to float where the reference's float assignments do it) so results agree to the last bit or ulp.
//
// combine: sequential 2-way merge per ray with the reference's min-distance filter (a sample is dropped when it is closer than
// min_dist to the last KEPT sample — a true sequential dependency), one thread per ray.
#include "vs_common.cuh"

namespace vs {

// ---- pcg32 (kernels/volsurfs/pcg32.h:32-34,60-70,158-180): state advance by `delta` steps, then one draw ------------------------
constexpr uint64_t kPcgMult = 0x5851f42d4c957f2dULL;

__device__ __forceinline__ uint64_t pcg_advance(uint64_t state, uint64_t inc, uint64_t delta) {
    uint64_t cur_mult = kPcgMult, cur_plus = inc, acc_mult = 1u, acc_plus = 0u;
    while (delta > 0) {
        if (delta & 1) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
    return acc_mult * state + acc_plus;
}
__device__ __forceinline__ float pcg_float_at(uint64_t state) {
    const uint32_t xorshifted = (uint32_t)(((state >> 18u) ^ state) >> 27u);
    const uint32_t rot = (uint32_t)(state >> 59u);
    const uint32_t r = (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
    return __uint_as_float((r >> 9) | 0x3f800000u) - 1.0f;
}

// VolumeRenderingGPU.cuh:15-21
__device__ __forceinline__ float map_range_val(float v, float in_start, float in_end, float out_start, float out_end) {
    const float clamped = fmaxf(in_start, fminf(in_end, v));
    if (in_start >= in_end) return out_end;
    return out_start + ((out_end - out_start) / (in_end - in_start)) * (clamped - in_start);
}

// VolumeRenderingGPU.cuh:481-505: index of the first cdf entry above `val` in [imin, imax]
__device__ __forceinline__ int cdf_search(const float* __restrict__ cdf, float val, int imin, int imax) {
    if (imax <= imin) return imax;  // single-sample segment (the reference loop would not terminate)
    while (imax >= imin) {
        const int imid = imin + (imax - imin) / 2;
        if (__ldg(cdf + imid) > val)
            imax = imid;
        else
            imin = imid;
        if (imax - imin == 1) return imax;
    }
    return imax;
}

__global__ void importance_sample_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const int32_t* __restrict__ se,
                                         const float* __restrict__ z, const float* __restrict__ cdf, int64_t n_rays, int n_imp,
                                         uint64_t rng_state, uint64_t rng_inc, int jitter, float* __restrict__ out_3d,
                                         float* __restrict__ out_dirs, float* __restrict__ out_z, int32_t* __restrict__ out_se) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rays * n_imp) return;
    const int64_t ray = t / n_imp;
    const int i = (int)(t - ray * n_imp);
    int start;
    const int count = load_segment(se, ray, start);
    if (count == 0) return;  // slots and ray_start_end_idx keep the constructor's -1 fill
    const int end = start + count;

    const float dist = (float)(1.0 / (double)(n_imp + 1));
    float u = dist + i * dist;
    if (jitter) {
        // the reference thread has called advance(ray) (i+1) times and drawn i numbers before this draw
        const uint64_t st = pcg_advance(rng_state, rng_inc, (uint64_t)ray * (uint64_t)(i + 1) + (uint64_t)i);
        const float rnd = pcg_float_at(st);
        const float mov = (float)((double)dist / 2.0);
        u += map_range_val(rnd, 0.0f, 1.0f, -mov, +mov);
    }
    u = fmaxf((float)(0.0 + 1e-6), fminf((float)(1.0 - 1e-6), u));

    const int imax = cdf_search(cdf, u, start, end - 1);
    const int imin = max(imax - 1, 0);
    const float z_imp = map_range_val(u, __ldg(cdf + imin), __ldg(cdf + imax), __ldg(z + imin), __ldg(z + imax));

    const float ox = __ldg(rays_o + 3 * ray), oy = __ldg(rays_o + 3 * ray + 1), oz = __ldg(rays_o + 3 * ray + 2);
    const float dx = __ldg(rays_d + 3 * ray), dy = __ldg(rays_d + 3 * ray + 1), dz = __ldg(rays_d + 3 * ray + 2);
    out_3d[3 * t] = ox + z_imp * dx;
    out_3d[3 * t + 1] = oy + z_imp * dy;
    out_3d[3 * t + 2] = oz + z_imp * dz;
    out_dirs[3 * t] = dx;
    out_dirs[3 * t + 1] = dy;
    out_dirs[3 * t + 2] = dz;
    out_z[t] = z_imp;
    if (i == 0) {
        out_se[2 * ray] = (int32_t)(ray * n_imp);
        out_se[2 * ray + 1] = (int32_t)(ray * n_imp + n_imp);
    }
}

__global__ void combine_kernel(int64_t n_rays, float min_dist, int values_dim, const int32_t* __restrict__ se1, const int32_t* __restrict__ idx1,
                               const float* __restrict__ p1, const float* __restrict__ d1, const float* __restrict__ z1,
                               const float* __restrict__ v1, const int32_t* __restrict__ se2, const int32_t* __restrict__ idx2,
                               const float* __restrict__ p2, const float* __restrict__ d2, const float* __restrict__ z2,
                               const float* __restrict__ v2, const int32_t* __restrict__ out_start, int32_t* __restrict__ c_idx,
                               float* __restrict__ c_3d, float* __restrict__ c_dirs, float* __restrict__ c_z, float* __restrict__ c_val,
                               int32_t* __restrict__ c_se) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    int s1, s2;
    const int n1 = load_segment(se1, r, s1), n2 = load_segment(se2, r, s2);
    if (n1 == 0 && n2 == 0) return;
    const int base = out_start[r];
    int a = 0, b = 0, written = 0;
    // a packet without samples for this ray counts as exhausted (the reference would read row start-1 = -2 here)
    bool done1 = n1 == 0, done2 = n2 == 0;
    float prec_z = 0.0f;
    for (int it = 0; it < n1 + n2; ++it) {
        if (done1 && done2) break;
        const float za = done1 ? 1e10f : __ldg(z1 + s1 + a);
        const float zb = done2 ? 1e10f : __ldg(z2 + s2 + b);
        const bool take1 = za < zb;
        const float zz = take1 ? za : zb;
        const int src = take1 ? s1 + a : s2 + b;
        if (!(zz - prec_z < min_dist)) {
            const int dst = base + written;
            const int32_t* si = take1 ? idx1 : idx2;
            const float* sp = take1 ? p1 : p2;
            const float* sd = take1 ? d1 : d2;
            const float* sv = take1 ? v1 : v2;
            c_idx[dst] = __ldg(si + src);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                c_3d[3 * dst + c] = __ldg(sp + 3 * src + c);
                c_dirs[3 * dst + c] = __ldg(sd + 3 * src + c);
            }
            c_z[dst] = zz;
            prec_z = zz;
            for (int c = 0; c < values_dim; ++c) c_val[(int64_t)dst * values_dim + c] = __ldg(sv + (int64_t)src * values_dim + c);
            ++written;
        }
        if (take1) {
            if (a + 1 >= n1)
                done1 = true;
            else
                ++a;
        } else {
            if (b + 1 >= n2)
                done2 = true;
            else
                ++b;
        }
    }
    c_se[2 * r] = base;
    c_se[2 * r + 1] = base + written;
}

}  // namespace vs

using namespace vs;

extern "C" {

// Uncompacted importance samples: ray r owns rows [r*n_imp, (r+1)*n_imp) of out_* (rows and out_se of rays without uniform samples
// are left untouched: the caller pre-fills -1 like the RaySamplesPacked constructor).  rng_state/rng_inc: the pcg32 generator the
// reference passes by value (default: state 0x853c49e6748fea9b, inc 0xda3e39cb94b95bdb); the caller advances it by 2^32 after a
// jittered call (VolumeRendering.cu:520-523).
int vs_importance_sample(const float* rays_o, const float* rays_d, const int32_t* se, const float* z, const float* cdf, int64_t n_rays,
                         int64_t n_samples, int n_imp, uint64_t rng_state, uint64_t rng_inc, int jitter, float* out_3d, float* out_dirs,
                         float* out_z, int32_t* out_se, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && n_samples >= 0 && n_imp > 0);
    if (n_rays == 0) return VS_OK;
    VS_CHECK_ARG(rays_o && rays_d && se && z && cdf && out_3d && out_dirs && out_z && out_se);
    VS_CHECK_ARG(n_rays * (int64_t)n_imp <= 0x7fffffffLL);
    const int threads = 256;
    const int64_t total = n_rays * n_imp;
    importance_sample_kernel<<<(unsigned)div_up(total, threads), threads, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, se, z, cdf, n_rays, n_imp, rng_state, rng_inc, jitter, out_3d, out_dirs, out_z, out_se);
    return launched(1);
}

// z-ordered merge of two compacted packets into the uncompacted combined packet (rows [out_start[r], out_start[r]+written_r),
// out_start from vs_combine_offsets);
// c_se of rays without samples is left untouched (pre-filled -1).
int vs_combine_merge(int64_t n_rays, float min_dist, int values_dim, const int32_t* se1, const int32_t* idx1, const float* p1, const float* d1,
                     const float* z1, const float* v1, const int32_t* se2, const int32_t* idx2, const float* p2, const float* d2,
                     const float* z2, const float* v2, const int32_t* out_start, int32_t* c_idx, float* c_3d, float* c_dirs, float* c_z,
                     float* c_val, int32_t* c_se, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && values_dim >= 0);
    if (n_rays == 0) return VS_OK;
    VS_CHECK_ARG(se1 && se2 && out_start && c_idx && c_3d && c_dirs && c_z && c_se && (values_dim == 0 || (v1 && v2 && c_val)));
    combine_kernel<<<(unsigned)div_up(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(n_rays, min_dist, values_dim, se1, idx1, p1, d1, z1, v1, se2,
                                                                                     idx2, p2, d2, z2, v2, out_start, c_idx, c_3d, c_dirs, c_z,
                                                                                     c_val, c_se);
    return launched(1);
}

}  // extern "C"
