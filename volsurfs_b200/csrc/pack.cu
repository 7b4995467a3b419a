// volsurfs_b200 — packing: per-ray counts -> exclusive scan -> dense RaySamplesPacked arrays.
//
// Replaces (a) RaySamplesPacked::compact_to_valid_samples (src/RaySamplesPacked.cu:188-273, kernel
// kernels/volsurfs/RaySamplesPackedGPU.cuh:172-257; the reference takes the offsets from torch.cumsum and copies one
// ray per thread) and (b) the K-layer hit bookkeeping of the volsurfs method (volsurfs.py:476-518, ~10 boolean-mask
// kernels per layer) by writing the sorted hits of all K layers straight into packed form.
//
// All index arithmetic is int32 like the reference's (ray_start_end_idx is int32): totals above 2^31-1 are refused.
// The scan is a three-launch reduce / scan-of-block-sums / apply scheme (deterministic, no spinning).
#include <cstdlib>
#include "vs_common.cuh"

namespace vs {

constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanBlock * kScanItems;  // 2048 counts per block

__device__ __forceinline__ int warp_incl_scan_i32(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(VS_FULL_MASK, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// exclusive scan of one int per thread over a 256-thread block; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ int block_excl_scan_i32(int v, int* total, int* warp_sums /* >= 8 ints */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = warp_incl_scan_i32(v, lane);
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    int wprefix = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < kScanBlock / 32; ++k) {
        int s = warp_sums[k];
        if (k < wid) wprefix += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return wprefix + incl - v;
}

__global__ void __launch_bounds__(kScanBlock) counts_from_segments_kernel(const int32_t* __restrict__ se, int32_t* __restrict__ counts,
                                                                          int64_t n_rays) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    int start;
    counts[r] = load_segment(se, r, start);
}

__global__ void __launch_bounds__(kScanBlock) counts_from_two_segments_kernel(const int32_t* __restrict__ se1, const int32_t* __restrict__ se2,
                                                                              int32_t* __restrict__ counts, int64_t n_rays) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    int s1, s2;
    counts[r] = load_segment(se1, r, s1) + load_segment(se2, r, s2);
}

// hit counts of the K-layer trace: depth is layer-major [K, n_rays]; a layer is hit iff depth <= t_far
// (raytracelib/raytracer.py:100: is_hit = depth <= t_far)
__global__ void __launch_bounds__(kScanBlock) counts_from_hits_kernel(const float* __restrict__ depth, int32_t* __restrict__ counts,
                                                                      int64_t n_rays, int K, float t_far) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    int c = 0;
    for (int k = 0; k < K; ++k) c += (__ldg(depth + (int64_t)k * n_rays + r) <= t_far) ? 1 : 0;
    counts[r] = c;
}

__global__ void __launch_bounds__(kScanBlock) scan_reduce_kernel(const int32_t* __restrict__ counts, int64_t n,
                                                                 long long* __restrict__ block_sums) {
    __shared__ long long red[kScanBlock / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    long long s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        int64_t i = base + (int64_t)k * kScanBlock + threadIdx.x;
        if (i < n) s += counts[i];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(VS_FULL_MASK, s, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int k = 0; k < kScanBlock / 32; ++k) t += red[k];
        block_sums[blockIdx.x] = t;
    }
}

// single block: in-place exclusive scan of the block sums, total written to *total_out
__global__ void __launch_bounds__(1024) scan_block_sums_kernel(long long* __restrict__ block_sums, int64_t n_blocks,
                                                               long long* __restrict__ total_out) {
    __shared__ long long warp_sums[32];
    __shared__ long long carry_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_blocks; base += 1024) {
        const int64_t i = base + threadIdx.x;
        long long v = i < n_blocks ? block_sums[i] : 0;
        long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            long long o = __shfl_up_sync(VS_FULL_MASK, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        long long wprefix = 0, tot = 0;
        for (int k = 0; k < 32; ++k) {
            long long s = warp_sums[k];
            if (k < wid) wprefix += s;
            tot += s;
        }
        const long long carry = carry_s;
        if (i < n_blocks) block_sums[i] = carry + wprefix + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry_s;
}

// offsets[i] = exclusive prefix of counts (int32); items are taken in blocked order so the scan is a plain
// per-thread serial scan + block scan of thread sums
__global__ void __launch_bounds__(kScanBlock) scan_apply_kernel(const int32_t* __restrict__ counts, int64_t n,
                                                                const long long* __restrict__ block_offsets,
                                                                int32_t* __restrict__ offsets) {
    __shared__ int warp_sums[kScanBlock / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int v[kScanItems];
    int tsum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? counts[base + k] : 0;
        tsum += v[k];
    }
    int btotal;
    int tprefix = block_excl_scan_i32(tsum, &btotal, warp_sums);
    long long run = block_offsets[blockIdx.x] + tprefix;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) offsets[base + k] = (int32_t)run;
        run += v[k];
    }
}

// --------------------------------------------------------------------------------------------
// generic compaction gather: one group of W lanes per ray copies the ray's segment
// --------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256) compact_gather_kernel(const int32_t* __restrict__ se_in, const int32_t* __restrict__ offsets,
                                                             const int32_t* __restrict__ idx_in, const float* __restrict__ p3d_in,
                                                             const float* __restrict__ dirs_in, const float* __restrict__ z_in,
                                                             const float* __restrict__ dt_in, const float* __restrict__ val_in,
                                                             int values_dim, int32_t* __restrict__ se_out, int32_t* __restrict__ idx_out,
                                                             float* __restrict__ p3d_out, float* __restrict__ dirs_out,
                                                             float* __restrict__ z_out, float* __restrict__ dt_out,
                                                             float* __restrict__ val_out, int64_t n_rays) {
    const int gl = threadIdx.x & (W - 1);
    const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / W;
    if (ray >= n_rays) return;
    int start;
    const int n = load_segment(se_in, ray, start);
    const int o = __ldg(offsets + ray);
    if (gl == 0) {
        // rays without samples keep (-1,-1) (RaySamplesPackedGPU.cuh:211-212)
        int2 v = n > 0 ? make_int2(o, o + n) : make_int2(-1, -1);
        reinterpret_cast<int2*>(se_out)[ray] = v;
    }
    for (int i = gl; i < n; i += W) {
        const int64_t a = (int64_t)start + i, b = (int64_t)o + i;
        idx_out[b] = idx_in[a];
        z_out[b] = z_in[a];
        dt_out[b] = dt_in[a];
    }
    for (int e = gl; e < 3 * n; e += W) {
        const int64_t a = 3 * (int64_t)start + e, b = 3 * (int64_t)o + e;
        p3d_out[b] = p3d_in[a];
        dirs_out[b] = dirs_in[a];
    }
    if (val_in != nullptr && val_out != nullptr) {
        for (int e = gl; e < values_dim * n; e += W) val_out[(int64_t)values_dim * o + e] = val_in[(int64_t)values_dim * start + e];
    }
}

// --------------------------------------------------------------------------------------------
// K-layer hits -> packed samples.  Layer-major inputs [K, n_rays]; packed order per ray is outer -> inner, i.e.
// descending mesh index (volsurfs.py:601-603).  samples_idx = source slot r*K + rank (slot layout of
// RaySamplerGPU.cuh:206-271 with M = K), samples_z = t, samples_3d = o + t*d (separate fp32 multiply and add, the
// rounding contract shared with the oracle), samples_dirs = d.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_hits_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                        const float* __restrict__ depth, const int32_t* __restrict__ tri,
                                                        const float* __restrict__ bary_u, const float* __restrict__ bary_v,
                                                        const int32_t* __restrict__ offsets, int K, float t_far,
                                                        int32_t* __restrict__ se_out, int32_t* __restrict__ idx_out,
                                                        float* __restrict__ p3d_out, float* __restrict__ dirs_out,
                                                        float* __restrict__ z_out, int32_t* __restrict__ layer_out,
                                                        int32_t* __restrict__ tri_out, float* __restrict__ uv_out, int64_t n_rays) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
    const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
    const int o = offsets[r];
    int rank = 0;
    for (int k = K - 1; k >= 0; --k) {
        const int64_t src = (int64_t)k * n_rays + r;
        const float t = __ldg(depth + src);
        if (t <= t_far) {
            const int64_t b = (int64_t)o + rank;
            idx_out[b] = (int32_t)(r * K + rank);
            z_out[b] = t;
            p3d_out[3 * b] = __fmaf_rn(dx, t, ox);  // = the reference kernel's positions (bvh.cu:445, contracted to one FMA by nvcc)
            p3d_out[3 * b + 1] = __fmaf_rn(dy, t, oy);
            p3d_out[3 * b + 2] = __fmaf_rn(dz, t, oz);
            dirs_out[3 * b] = dx;
            dirs_out[3 * b + 1] = dy;
            dirs_out[3 * b + 2] = dz;
            if (layer_out) layer_out[b] = k;
            if (tri_out) tri_out[b] = tri ? __ldg(tri + src) : -1;
            if (uv_out) {
                uv_out[2 * b] = bary_u ? __ldg(bary_u + src) : 0.f;
                uv_out[2 * b + 1] = bary_v ? __ldg(bary_v + src) : 0.f;
            }
            ++rank;
        }
    }
    int2 v = rank > 0 ? make_int2(o, o + rank) : make_int2(-1, -1);
    reinterpret_cast<int2*>(se_out)[r] = v;
}

// The same scatter with unit-stride stores.  pack_hits_kernel writes a ray's hits from the ray's own thread: neighbouring lanes store K
// elements apart (12 scattered partial sectors per store instruction; 61 us per 640 k rays x 5 layers on B200 for 110 MB of traffic).
// Here a warp's 32 rays (one contiguous range of the packed arrays) are assembled in shared memory, field by field, and every array
// leaves as consecutive 4-byte words.  Values and order are identical (tests/test_gpu_packing.py is bit-exact on both kernels).
constexpr int kPackWarps = 4;
__global__ void __launch_bounds__(32 * kPackWarps) pack_hits_staged_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ depth, const int32_t* __restrict__ tri,
    const float* __restrict__ bary_u, const float* __restrict__ bary_v, const int32_t* __restrict__ offsets, int K, float t_far,
    int32_t* __restrict__ se_out, int32_t* __restrict__ idx_out, float* __restrict__ p3d_out, float* __restrict__ dirs_out,
    float* __restrict__ z_out, int32_t* __restrict__ layer_out, int32_t* __restrict__ tri_out, float* __restrict__ uv_out, int64_t n_rays) {
    extern __shared__ __align__(16) float pack_sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r0 = ((int64_t)blockIdx.x * kPackWarps + warp) * 32;
    if (r0 >= n_rays) return;
    const int64_t r = r0 + lane;
    const bool live = r < n_rays;
    const int cap = 32 * K;  // samples a warp can produce
    float* sm = pack_sm + (size_t)warp * cap * 12;
    int32_t* s_idx = reinterpret_cast<int32_t*>(sm);
    float* s_z = sm + cap;
    int32_t* s_layer = reinterpret_cast<int32_t*>(sm + 2 * cap);
    int32_t* s_tri = reinterpret_cast<int32_t*>(sm + 3 * cap);
    float* s_uv = sm + 4 * cap;
    float* s_p3d = sm + 6 * cap;
    float* s_dirs = sm + 9 * cap;

    float ox = 0.f, oy = 0.f, oz = 0.f, dx = 0.f, dy = 0.f, dz = 0.f;
    int o = 0;
    if (live) {
        ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
        dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
        o = offsets[r];
    }
    const int base = __shfl_sync(VS_FULL_MASK, o, 0);
    int rank = 0;
    if (live) {
        for (int k = K - 1; k >= 0; --k) {
            const int64_t src = (int64_t)k * n_rays + r;
            const float t = __ldg(depth + src);
            if (t <= t_far) {
                const int b = o - base + rank;
                s_idx[b] = (int32_t)(r * K + rank);
                s_z[b] = t;
                s_p3d[3 * b] = __fmaf_rn(dx, t, ox);  // = the reference kernel's positions (bvh.cu:445, contracted to one FMA by nvcc)
                s_p3d[3 * b + 1] = __fmaf_rn(dy, t, oy);
                s_p3d[3 * b + 2] = __fmaf_rn(dz, t, oz);
                s_dirs[3 * b] = dx;
                s_dirs[3 * b + 1] = dy;
                s_dirs[3 * b + 2] = dz;
                if (layer_out) s_layer[b] = k;
                if (tri_out) s_tri[b] = tri ? __ldg(tri + src) : -1;
                if (uv_out) {
                    s_uv[2 * b] = bary_u ? __ldg(bary_u + src) : 0.f;
                    s_uv[2 * b + 1] = bary_v ? __ldg(bary_v + src) : 0.f;
                }
                ++rank;
            }
        }
        reinterpret_cast<int2*>(se_out)[r] = rank > 0 ? make_int2(o, o + rank) : make_int2(-1, -1);
    }
    const int total = (int)__reduce_max_sync(VS_FULL_MASK, (unsigned)(live ? o - base + rank : 0));  // offsets are monotone
    __syncwarp();
    const int64_t b0 = base;
    for (int i = lane; i < total; i += 32) {
        idx_out[b0 + i] = s_idx[i];
        z_out[b0 + i] = s_z[i];
        if (layer_out) layer_out[b0 + i] = s_layer[i];
        if (tri_out) tri_out[b0 + i] = s_tri[i];
    }
    if (uv_out)
        for (int i = lane; i < 2 * total; i += 32) uv_out[2 * b0 + i] = s_uv[i];
    for (int i = lane; i < 3 * total; i += 32) {
        p3d_out[3 * b0 + i] = s_p3d[i];
        dirs_out[3 * b0 + i] = s_dirs[i];
    }
}

__global__ void __launch_bounds__(256) count_total_kernel(const int32_t* __restrict__ se, int64_t n_rays,
                                                          unsigned long long* __restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    long long c = 0;
    if (r < n_rays) {
        int start;
        c = load_segment(se, r, start);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(VS_FULL_MASK, c, d);
    if ((threadIdx.x & 31) == 0 && c != 0) atomicAdd(out, (unsigned long long)c);
}

static inline int64_t scan_blocks(int64_t n) { return div_up(n, kScanTile); }

// counts (int32[n]) -> offsets (int32[n]) + total (int64 on device). scratch: >= vs_scan_scratch_bytes(n)
static int launch_scan(const int32_t* counts, int64_t n, int32_t* offsets, long long* total_dev, void* scratch, cudaStream_t st) {
    const int64_t nb = scan_blocks(n);
    long long* block_sums = reinterpret_cast<long long*>(scratch);
    scan_reduce_kernel<<<(unsigned)nb, kScanBlock, 0, st>>>(counts, n, block_sums);
    scan_block_sums_kernel<<<1, 1024, 0, st>>>(block_sums, nb, total_dev);
    scan_apply_kernel<<<(unsigned)nb, kScanBlock, 0, st>>>(counts, n, block_sums, offsets);
    return launched(3);
}

}  // namespace vs

using namespace vs;

extern "C" {

// bytes of device scratch the packing entry points need for n_rays rays:
//   [block sums: 8*ceil(n/2048)] [counts: 4*n] [offsets: 4*n], each region 256-byte aligned
int64_t vs_pack_scratch_bytes(int64_t n_rays) {
    if (n_rays < 0) return 0;
    auto al = [](int64_t b) { return (b + 255) / 256 * 256; };
    return al(8 * (scan_blocks(n_rays) + 1)) + 2 * al(4 * n_rays) + 256;
}

static void carve_scratch(void* scratch, int64_t n_rays, long long** block_sums, int32_t** counts, int32_t** offsets) {
    auto al = [](int64_t b) { return (b + 255) / 256 * 256; };
    char* p = reinterpret_cast<char*>(scratch);
    *block_sums = reinterpret_cast<long long*>(p);
    p += al(8 * (scan_blocks(n_rays) + 1));
    *counts = reinterpret_cast<int32_t*>(p);
    p += al(4 * n_rays);
    *offsets = reinterpret_cast<int32_t*>(p);
}

int vs_count_total(const int32_t* se, int64_t n_rays, int64_t* total_dev, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && total_dev);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(total_dev, 0, sizeof(int64_t), st);
    if (e != cudaSuccess) return (int)e;
    if (n_rays == 0) return VS_OK;
    VS_CHECK_ARG(se);
    count_total_kernel<<<(unsigned)div_up(n_rays, 256), 256, 0, st>>>(se, n_rays, reinterpret_cast<unsigned long long*>(total_dev));
    return launched(1);
}

// Phase 1 of compact_to_valid_samples: offsets + total.  The caller reads *total_dev (one D2H copy) to size the
// output tensors exactly like the reference does, then calls vs_compact_gather.
int vs_compact_offsets(const int32_t* se_in, int64_t n_rays, int64_t* total_dev, void* scratch, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && total_dev && scratch);
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rays == 0) return (int)cudaMemsetAsync(total_dev, 0, sizeof(int64_t), st);
    VS_CHECK_ARG(se_in);
    long long* bs;
    int32_t *counts, *offsets;
    carve_scratch(scratch, n_rays, &bs, &counts, &offsets);
    counts_from_segments_kernel<<<(unsigned)div_up(n_rays, kScanBlock), kScanBlock, 0, st>>>(se_in, counts, n_rays);
    launched(1);
    return launch_scan(counts, n_rays, offsets, reinterpret_cast<long long*>(total_dev), bs, st);
}

// Compacted start of every ray's segment for producers that write straight into compacted form (csrc/sampler.cu): out_start[r] =
// exclusive prefix sum of (end - start) over se_in, *total_dev = the grand total.  scratch: vs_pack_scratch_bytes(n_rays).
int vs_segment_offsets(const int32_t* se_in, int64_t n_rays, int32_t* out_start, int64_t* total_dev, void* scratch, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && total_dev && scratch);
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rays == 0) return (int)cudaMemsetAsync(total_dev, 0, sizeof(int64_t), st);
    VS_CHECK_ARG(se_in && out_start);
    long long* bs;
    int32_t *counts, *offsets;
    carve_scratch(scratch, n_rays, &bs, &counts, &offsets);
    counts_from_segments_kernel<<<(unsigned)div_up(n_rays, kScanBlock), kScanBlock, 0, st>>>(se_in, counts, n_rays);
    launched(1);
    int e = launch_scan(counts, n_rays, offsets, reinterpret_cast<long long*>(total_dev), bs, st);
    if (e != VS_OK) return e;
    return (int)cudaMemcpyAsync(out_start, offsets, sizeof(int32_t) * n_rays, cudaMemcpyDeviceToDevice, st);
}

// Start offsets of combine_ray_samples_packets (src/VolumeRendering.cu:595-603: cumsum of count1+count2, shifted): out_start[r] =
// exclusive prefix sum of the two packets' per-ray counts, *total_dev = their grand total.  scratch: vs_pack_scratch_bytes(n_rays).
int vs_combine_offsets(const int32_t* se1, const int32_t* se2, int64_t n_rays, int32_t* out_start, int64_t* total_dev, void* scratch,
                       void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && total_dev && scratch);
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rays == 0) return (int)cudaMemsetAsync(total_dev, 0, sizeof(int64_t), st);
    VS_CHECK_ARG(se1 && se2 && out_start);
    long long* bs;
    int32_t *counts, *offsets;
    carve_scratch(scratch, n_rays, &bs, &counts, &offsets);
    counts_from_two_segments_kernel<<<(unsigned)div_up(n_rays, kScanBlock), kScanBlock, 0, st>>>(se1, se2, counts, n_rays);
    launched(1);
    int e = launch_scan(counts, n_rays, offsets, reinterpret_cast<long long*>(total_dev), bs, st);
    if (e != VS_OK) return e;
    return (int)cudaMemcpyAsync(out_start, offsets, sizeof(int32_t) * n_rays, cudaMemcpyDeviceToDevice, st);
}

int vs_compact_gather(const int32_t* se_in, const void* scratch, const int32_t* idx_in, const float* p3d_in, const float* dirs_in,
                      const float* z_in, const float* dt_in, const float* val_in, int values_dim, int32_t* se_out, int32_t* idx_out,
                      float* p3d_out, float* dirs_out, float* z_out, float* dt_out, float* val_out, int64_t n_rays, int64_t n_samples_out,
                      void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && n_samples_out >= 0 && values_dim >= 0);
    if (n_rays == 0) return VS_OK;
    VS_CHECK_ARG(se_in && scratch && se_out);
    VS_CHECK_ARG(n_samples_out == 0 || (idx_in && p3d_in && dirs_in && z_in && dt_in && idx_out && p3d_out && dirs_out && z_out && dt_out));
    cudaStream_t st = (cudaStream_t)stream;
    long long* bs;
    int32_t *counts, *offsets;
    carve_scratch(const_cast<void*>(scratch), n_rays, &bs, &counts, &offsets);
    int Wsel = pick_group_width(n_rays, n_samples_out);
    const unsigned grid = (unsigned)div_up(n_rays * Wsel, 256);
#define VS_GATHER(W)                                                                                                                   \
    compact_gather_kernel<W><<<grid, 256, 0, st>>>(se_in, offsets, idx_in, p3d_in, dirs_in, z_in, dt_in, val_in, values_dim, se_out,   \
                                                    idx_out, p3d_out, dirs_out, z_out, dt_out, val_out, n_rays)
    switch (Wsel) {
        case 4: VS_GATHER(4); break;
        case 8: VS_GATHER(8); break;
        case 16: VS_GATHER(16); break;
        default: VS_GATHER(32); break;
    }
#undef VS_GATHER
    return launched(1);
}

// K-layer hit packing, phase 1: counts from depth <= t_far, offsets, total
int vs_pack_hits_offsets(const float* depth /*[K,n_rays]*/, int K, float t_far, int64_t n_rays, int64_t* total_dev, void* scratch,
                         void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && K > 0 && total_dev && scratch);
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rays == 0) return (int)cudaMemsetAsync(total_dev, 0, sizeof(int64_t), st);
    VS_CHECK_ARG(depth);
    if ((double)n_rays * K > 2147483647.0) return VS_ERR_UNSUPPORTED;
    long long* bs;
    int32_t *counts, *offsets;
    carve_scratch(scratch, n_rays, &bs, &counts, &offsets);
    counts_from_hits_kernel<<<(unsigned)div_up(n_rays, kScanBlock), kScanBlock, 0, st>>>(depth, counts, n_rays, K, t_far);
    launched(1);
    return launch_scan(counts, n_rays, offsets, reinterpret_cast<long long*>(total_dev), bs, st);
}

// phase 2: scatter.  Output arrays must hold at least total samples (n_rays*K always suffices).
int vs_pack_hits_scatter(const float* rays_o, const float* rays_d, const float* depth, const int32_t* tri, const float* bary_u,
                         const float* bary_v, const void* scratch, int K, float t_far, int32_t* se_out, int32_t* idx_out, float* p3d_out,
                         float* dirs_out, float* z_out, int32_t* layer_out, int32_t* tri_out, float* uv_out, int64_t n_rays,
                         void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && K > 0);
    if (n_rays == 0) return VS_OK;
    VS_CHECK_ARG(rays_o && rays_d && depth && scratch && se_out && idx_out && p3d_out && dirs_out && z_out);
    cudaStream_t st = (cudaStream_t)stream;
    long long* bs;
    int32_t *counts, *offsets;
    carve_scratch(const_cast<void*>(scratch), n_rays, &bs, &counts, &offsets);
    const size_t smem = (size_t)kPackWarps * 32 * K * 12 * sizeof(float);
    static const bool direct = std::getenv("VS_PACK_DIRECT") != nullptr;  // A/B knob: the thread-per-ray scatter
    if (smem <= 96 * 1024 && !direct) {
        cudaError_t e = cudaFuncSetAttribute(pack_hits_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        pack_hits_staged_kernel<<<(unsigned)div_up(n_rays, 32 * kPackWarps), 32 * kPackWarps, smem, st>>>(
            rays_o, rays_d, depth, tri, bary_u, bary_v, offsets, K, t_far, se_out, idx_out, p3d_out, dirs_out, z_out, layer_out, tri_out, uv_out,
            n_rays);
    } else {
        pack_hits_kernel<<<(unsigned)div_up(n_rays, 256), 256, 0, st>>>(rays_o, rays_d, depth, tri, bary_u, bary_v, offsets, K, t_far, se_out,
                                                                       idx_out, p3d_out, dirs_out, z_out, layer_out, tri_out, uv_out, n_rays);
    }
    return launched(1);
}

}  // extern "C"
