// volsurfs_b200 — the tail of the per-ray path in training: background blend + L1 photometric loss + its gradient, one launch.
//
// Replaces, for a step that composites with csrc/composite.cu, the torch glue between the compositing forward and backward:
//   pred  = rgb_fg + bgT * bg                                  volsurfs_py/methods/volsurfs.py:708
//   loss  = (gt - pred).abs().mean()                           volsurfs_py/utils/losses.py:14-19 (volsurfs.py:804-806, mask = None)
//   g_pred = d loss / d pred = sign(pred - gt) / numel         (autograd of the two lines above)
//   g_bgT  = sum_c g_pred_c * bg_c                             (the bgT branch of line 708)
// which are nine small launches (addcmul, sub, abs, mean, sign, div, mul, sum, zeros) and ~45 us of the 3 ms benchmark step.
#include "vs_common.cuh"

namespace vs {

constexpr int kLossThreads = 256;

__global__ void __launch_bounds__(kLossThreads) blend_l1_kernel(const float* __restrict__ rgb_fg, const float* __restrict__ bgT,
                                                                const float* __restrict__ gt, float bg_r, float bg_g, float bg_b,
                                                                float inv_numel, float* __restrict__ pred, float* __restrict__ g_pred,
                                                                float* __restrict__ g_bgT, float* __restrict__ loss,
                                                                unsigned long long* __restrict__ scratch, int64_t n_rays) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float part = 0.f;
    if (r < n_rays) {
        const float T = __ldg(bgT + r);
        const float bg[3] = {bg_r, bg_g, bg_b};
        float gb = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float p = __fadd_rn(__ldg(rgb_fg + 3 * r + c), __fmul_rn(T, bg[c]));  // torch.addcmul: self + t1 * t2, two roundings
            const float d = __fsub_rn(p, __ldg(gt + 3 * r + c));
            part += fabsf(d);
            const float g = d > 0.f ? inv_numel : (d < 0.f ? -inv_numel : 0.f);           // torch.sign(0) = 0
            pred[3 * r + c] = p;
            g_pred[3 * r + c] = g;
            gb = __fadd_rn(gb, __fmul_rn(g, bg[c]));                                      // (g_pred * bg).sum(dim=1), left to right
        }
        g_bgT[r] = gb;
    }
    // block sum of |diff| (fixed shuffle order) -> 2^-32 fixed point -> one INTEGER atomic per block: the total does not depend on the order
    // the blocks arrive in, so a replayed CUDA graph reproduces the eager step's loss bit for bit.  The last block to arrive (ticket in
    // scratch[1]) turns the total into the fp32 mean and clears the scratch for the next call.
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(VS_FULL_MASK, part, d);
    __shared__ float s_part[kLossThreads / 32];
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < kLossThreads / 32 ? s_part[threadIdx.x] : 0.f;
#pragma unroll
        for (int d = 4; d > 0; d >>= 1) v += __shfl_xor_sync(VS_FULL_MASK, v, d);
        if (threadIdx.x == 0) {
            // 2^-32 fixed point: a block's sum (768 terms) of colour differences stays far below 2^20, 2^12 x that many blocks below 2^64
            atomicAdd(scratch, (unsigned long long)fmin((double)v * 4294967296.0, 1.8e19));
            __threadfence();
            const unsigned long long ticket = atomicAdd(scratch + 1, 1ull);
            if (ticket == (unsigned long long)gridDim.x - 1) {
                __threadfence();
                const unsigned long long total = atomicExch(scratch, 0ull);
                atomicExch(scratch + 1, 0ull);
                *loss = (float)((double)total * (1.0 / 4294967296.0) * (double)inv_numel);
            }
        }
    }
}

}  // namespace vs

using namespace vs;

extern "C" {

// rgb_fg [n,3], bgT [n,1], gt [n,3] device; bg_rgb HOST float[3]; outputs pred [n,3], g_pred [n,3], g_bgT [n,1], loss: one device
// float; scratch: 16 bytes of device memory, zero when first used (the kernel leaves it zero), not shared by calls that may overlap
int vs_blend_l1_loss(const float* rgb_fg, const float* bgT, const float* gt, const float* bg_rgb, float* pred, float* g_pred, float* g_bgT,
                     float* loss, void* scratch, int64_t n_rays, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && loss && bg_rgb && scratch);
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rays == 0) return (int)cudaMemsetAsync(loss, 0, sizeof(float), st);
    VS_CHECK_ARG(rgb_fg && bgT && gt && pred && g_pred && g_bgT);
    VS_CHECK_ARG(div_up(n_rays, kLossThreads) < (1ll << 31));
    const float inv_numel = 1.0f / (float)(3 * n_rays);
    blend_l1_kernel<<<(unsigned)div_up(n_rays, kLossThreads), kLossThreads, 0, st>>>(rgb_fg, bgT, gt, bg_rgb[0], bg_rgb[1], bg_rgb[2], inv_numel,
                                                                                    pred, g_pred, g_bgT, loss,
                                                                                    reinterpret_cast<unsigned long long*>(scratch), n_rays);
    return launched(1);
}

}  // extern "C"
