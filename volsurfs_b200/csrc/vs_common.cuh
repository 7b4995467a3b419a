// volsurfs_b200 — shared device helpers (sm_100a only).
//
// Conventions used by every kernel in this directory
//   * a "ray segment" is ray_start_end_idx[r] = (start, end), int32 pairs; empty
//     rays carry (-1,-1) and are recognised by end-start == 0, exactly like the
//     reference kernels (kernels/volsurfs/VolumeRenderingGPU.cuh:47-52).
//   * "group-per-ray": a group of W lanes (W = 4,8,16,32; 32/W rays per warp)
//     walks one ray in W-wide chunks, carrying the running product / sum between
//     chunks.  Segmented scans are shuffle based (__shfl_up_sync with width W).
//   * nothing here synchronises the device or allocates memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define VS_FULL_MASK 0xffffffffu

// error codes of the C ABI: 0 ok, >0 cudaError_t, <0 argument errors
#define VS_OK 0
#define VS_ERR_INVALID_ARG (-1)
#define VS_ERR_UNSUPPORTED (-2)
#define VS_ERR_ALLOC (-3)

#define VS_CHECK_ARG(cond)              \
    do {                                \
        if (!(cond)) return VS_ERR_INVALID_ARG; \
    } while (0)

namespace vs {

// number of kernels this library has enqueued (read by vs_launch_count; bench.py reports it as gpu_launches)
extern long long g_launches;
inline int launched(int n_kernels) {
    __atomic_fetch_add(&g_launches, (long long)n_kernels, __ATOMIC_RELAXED);
    return (int)cudaGetLastError();
}

__host__ __device__ __forceinline__ int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- streaming loads/stores (data touched once: keep it out of L1) ----------
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ float4 ld_stream4(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream4(float4* p, float4 v) { __stcs(p, v); }

// ---- group (sub-warp) primitives ------------------------------------------
// inclusive product scan inside a W-lane group; gl = lane index in the group
template <int W>
__device__ __forceinline__ float group_scan_mul(float v, int gl) {
#pragma unroll
    for (int d = 1; d < W; d <<= 1) {
        float o = __shfl_up_sync(VS_FULL_MASK, v, d, W);
        if (gl >= d) v *= o;
    }
    return v;
}

template <int W>
__device__ __forceinline__ float group_scan_add(float v, int gl) {
#pragma unroll
    for (int d = 1; d < W; d <<= 1) {
        float o = __shfl_up_sync(VS_FULL_MASK, v, d, W);
        if (gl >= d) v += o;
    }
    return v;
}

template <int W>
__device__ __forceinline__ float group_reduce_add(float v) {
#pragma unroll
    for (int d = W / 2; d > 0; d >>= 1) v += __shfl_xor_sync(VS_FULL_MASK, v, d, W);
    return v;
}

// exclusive value from an inclusive scan: lane gl gets lane gl-1's value, lane 0 gets `first`
template <int W>
__device__ __forceinline__ float group_shift_up(float incl, int gl, float first) {
    float o = __shfl_up_sync(VS_FULL_MASK, incl, 1, W);
    return gl == 0 ? first : o;
}

template <int W>
__device__ __forceinline__ float group_bcast(float v, int src) {
    return __shfl_sync(VS_FULL_MASK, v, src, W);
}

// Reverse (suffix) inclusive scan of affine maps F_i(x) = A_i x + B_i inside a W-lane group:
// on return lane gl holds F_gl o F_{gl+1} o ... o F_{W-1}.
template <int W>
__device__ __forceinline__ void group_rscan_affine(float& A, float& B, int gl) {
#pragma unroll
    for (int d = 1; d < W; d <<= 1) {
        float Ao = __shfl_down_sync(VS_FULL_MASK, A, d, W);
        float Bo = __shfl_down_sync(VS_FULL_MASK, B, d, W);
        if (gl + d < W) {
            B = fmaf(A, Bo, B);
            A = A * Ao;
        }
    }
}

// ---- TMA bulk copies (cp.async.bulk, 1-D) + mbarrier: SASS UBLKCP / SYNCS ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared, completion counted in bytes on `bar`.  dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global (bulk async-group).  Generic-proxy writes to the source must be fenced with fence_proxy_async().
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
// pull `bytes` (multiple of 16) at `src` (16-byte aligned) into L2 ahead of use: one instruction, no destination, no completion
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// largest value of v over the warp (all lanes must call)
__device__ __forceinline__ int warp_max_i32(int v) { return __reduce_max_sync(VS_FULL_MASK, v); }

// segment of ray r; returns count (0 for empty / (-1,-1) rays)
__device__ __forceinline__ int load_segment(const int32_t* __restrict__ se, int64_t ray, int& start) {
    int2 v = __ldg(reinterpret_cast<const int2*>(se) + ray);
    start = v.x;
    return v.y - v.x;
}

// choose the group width from the mean segment length (host side)
inline int pick_group_width(int64_t n_rays, int64_t n_samples) {
    if (n_rays <= 0) return 32;
    double mean = (double)n_samples / (double)n_rays;
    if (mean <= 4.0) return 4;
    if (mean <= 8.0) return 8;
    if (mean <= 16.0) return 16;
    return 32;
}

}  // namespace vs
