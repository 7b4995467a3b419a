// volsurfs_b200 — declarations shared by the tensor-core appearance-head kernels (mlp.cu forward, mlp_bwd.cu backward):
// layer layout of the packed weight blob, tcgen05 / TMEM wrappers, UMMA descriptors, activation and SH helpers.
#pragma once
#include <cuda_fp16.h>

#include <algorithm>
#include <cstring>

#include "vs_common.cuh"

namespace vs {

constexpr int kMlpThreads = 512;   // 16 warps: 4 column groups x 4 lane quarters work on one 128-sample tile
constexpr int kTileM = 128;
constexpr int kMaxLayers = 6;
constexpr int kMaxWidth = 128;     // widest layer (TMEM columns allocated, A1 buffer)
constexpr int kExtraStride = 21;   // per-row scratch: 16 SH + 3 normal (+ pad to an odd stride: conflict-free)

struct MlpConfig {
    int n_layers;            // Linear layers (hidden + output)
    int k_pad[kMaxLayers];   // padded fan-in  (multiple of 16)
    int n_pad[kMaxLayers];   // padded fan-out (multiple of 16)
    int w_off[kMaxLayers];   // byte offsets into the packed blob
    int b_off[kMaxLayers];
    int blob_bytes;
    int pos_dim, n_sh, normal_dep, in_dim, out_dim;
    int activation;          // 0 relu, 1 gelu
    int alpha_decay;
    int out_linear;          // 1: the last layer's output is written as it is (tiny-cuda-nn "output_activation": "None"); 0: sigmoid
    int variant;             // debug: bit0 swaps LBO/SBO in the shared-memory descriptors
    int a1_width;            // widest hidden layer (size of the activation buffer), tmem_cols: power of two >= widest layer
    int tmem_cols;
    int h_bytes;             // per-slot hidden-activation buffer (also the TMA landing zone of the tile's fp32 features)
    int n_slots;             // tiles a CTA keeps in flight: 2 (MMA of one slot under the epilogue of the other) or 1 (large blobs)
};

// Activation stash written by the training-mode forward and read by the backward: per 128-sample tile, item 0 = the fp16 input operand
// A_0 of the first layer and item l >= 1 = the fp16 PRE-activations Z_{l-1} (bias included) of hidden layer l-1, each stored as the exact
// shared-memory image the tensor core consumes ([width/8][128][8] halves), so the backward fetches an item with one bulk copy.  The
// backward turns Z_{l-1} into the operand A_l = act(Z_{l-1}) in place and keeps act'(Z_{l-1}) in registers: one 2-byte value per hidden
// unit is all a layer costs (832 B per sample for [67,128,128,64,c]; the first version kept A_l and act' separately: 1472 B).
struct MlpStash {
    int a_off[kMaxLayers];
    int a_bytes[kMaxLayers];  // 128 * (k_pad + 8 * fold) * 2
    int fold[kMaxLayers];     // fan-in < 128: the item is followed by a chunk whose channel 0 is 1 (bias gradient = one more dW row)
    int tile_bytes;
};

static inline int pad16(int x) { return (x + 15) / 16 * 16; }

// dims = [in, h1, ..., h_{L-1}, out]
static inline int mlp_layout(int n_layers, const int* dims, MlpConfig* c) {
    if (n_layers < 1 || n_layers > kMaxLayers) return VS_ERR_UNSUPPORTED;
    std::memset(c, 0, sizeof(*c));
    c->n_layers = n_layers;
    int off = 0;
    for (int l = 0; l < n_layers; ++l) {
        if (dims[l] <= 0 || dims[l + 1] <= 0) return VS_ERR_INVALID_ARG;
        c->k_pad[l] = pad16(dims[l]);
        c->n_pad[l] = pad16(dims[l + 1]);
        if (c->k_pad[l] > kMaxWidth || c->n_pad[l] > kMaxWidth) return VS_ERR_UNSUPPORTED;
        if (l > 0 && dims[l] % 16 != 0) return VS_ERR_UNSUPPORTED;  // hidden widths must be multiples of 16
        c->w_off[l] = off;
        off += c->k_pad[l] * c->n_pad[l] * 2;
    }
    for (int l = 0; l < n_layers; ++l) {
        c->b_off[l] = off;
        off += c->n_pad[l] * 4;
    }
    c->blob_bytes = (off + 15) / 16 * 16;
    int widest = 16, widest_hidden = 16;
    for (int l = 0; l < n_layers; ++l) {
        widest = std::max(widest, c->n_pad[l]);
        if (l + 1 < n_layers) widest_hidden = std::max(widest_hidden, c->n_pad[l]);
    }
    c->a1_width = widest_hidden;
    c->tmem_cols = widest <= 32 ? 32 : (widest <= 64 ? 64 : 128);
    c->in_dim = dims[0];
    c->out_dim = dims[n_layers];
    if (c->out_dim > 32) return VS_ERR_UNSUPPORTED;  // sigmoid heads: <= 8 (checked by the callers); linear outputs: <= 32
    return VS_OK;
}

static inline void mlp_stash_layout(const MlpConfig& c, MlpStash* s) {
    std::memset(s, 0, sizeof(*s));
    int off = 0;
    for (int l = 0; l < c.n_layers; ++l) {
        s->a_off[l] = off;
        s->fold[l] = c.k_pad[l] + 8 <= kMaxWidth ? 1 : 0;
        s->a_bytes[l] = kTileM * (c.k_pad[l] + 8 * s->fold[l]) * 2;
        off += s->a_bytes[l];
    }
    s->tile_bytes = off;
}

// ---- tcgen05 wrappers -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T ; fp16 inputs, fp32 accumulate; one instruction covers K = 16
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 16 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// the same read split into issue and wait, so that several reads can be in flight
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, float* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]), "=f"(v[10]),
          "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle: element (row, k) of a [rows x K] fp16 operand lives at
//   (k/8) * LBO + (row/8) * SBO + (row%8) * 16 + (k%8) * 2     with LBO = rows*16 bytes, SBO = 128 bytes
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version of sm_100
    return d;                // layout_type (bits 61-63) = 0: no swizzle
}
// The same descriptor split into 32-bit halves: a GEMM's K-steps only advance the address field of the low word
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr, uint32_t lbo) { return ((saddr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16); }
__device__ __forceinline__ uint32_t umma_desc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3FFFu) | (1u << 14); }
__device__ __forceinline__ uint64_t umma_desc_join(uint32_t lo, uint32_t hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1,%2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}
// D[tmem] (+)= A * B over nk K-steps of 16 (one elected thread); the operand descriptors advance by a_step / b_step BYTES per K-step
__device__ __forceinline__ void umma_gemm_f16(uint32_t tmem_d, uint32_t a_addr, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_step, uint32_t b_addr,
                                              uint32_t b_lbo, uint32_t b_sbo, uint32_t b_step, uint32_t idesc, int nk, bool accumulate);
__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4)                      // D format: f32
           | (0u << 7) | (0u << 10)       // A, B format: f16
           | (0u << 15) | (0u << 16)      // A, B K-major
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_gemm_f16(uint32_t tmem_d, uint32_t a_addr, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_step, uint32_t b_addr,
                                              uint32_t b_lbo, uint32_t b_sbo, uint32_t b_step, uint32_t idesc, int nk, bool accumulate) {
    uint32_t a_lo = umma_desc_lo(a_addr, a_lbo), b_lo = umma_desc_lo(b_addr, b_lbo);
    const uint32_t a_hi = umma_desc_hi(a_sbo), b_hi = umma_desc_hi(b_sbo);
    const uint32_t a_inc = a_step >> 4, b_inc = b_step >> 4;
    uint32_t acc = accumulate ? 1u : 0u;
    for (int ks = 0; ks < nk; ++ks) {
        tc_mma_f16(tmem_d, umma_desc_join(a_lo, a_hi), umma_desc_join(b_lo, b_hi), idesc, acc);
        a_lo += a_inc;  // operand buffers never cross the 256 KB window of the 14-bit address field
        b_lo += b_inc;
        acc = 1u;
    }
}

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2: two fp32 lanes per issue slot) -------------------------------
__device__ __forceinline__ uint64_t f2_pack(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_splat(float a) { return f2_pack(a, a); }
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float tanh_approx(float x) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// GELU(x) = x Phi(x) (torch.nn.GELU(), exact-erf form: models/mlp.py:36) evaluated as
//     Phi(x) = 0.5 (1 + tanh(u(x))),   u(x) = x (c1 + c2 x^2 + c3 x^4)
// with (c1, c2, c3) fitted to atanh(erf(x / sqrt 2)) on |x| <= 8 (scripts/fit_gelu.py): max |GELU error| 3.7e-5, max |GELU' error| 9.3e-5
// — below the fp16 rounding of the activations — for ONE MUFU (tanh.approx, rel. error 2^-11) and ~5 packed-FMA issue slots per
// activation; the A&S 7.1.28 erf used before took 1 MUFU + ~9.5 slots and a second MUFU (ex2) for the derivative, and the MUFU pipe
// (16 lanes per SM and clock) was what bounded the forward kernel.  x^2 is clamped at 64: beyond |x| = 8 tanh has saturated and the fitted
// polynomial (c3 < 0) must not be followed any further.  With GRAD also d/dx = Phi + 0.5 x (1 - t^2) u'(x), from the same tanh.
constexpr float kGelu1 = 0.797422805f, kGelu2 = 0.0370039313f, kGelu3 = -3.47585310e-4f;

template <bool GRAD>
__device__ __forceinline__ void gelu_pair(float x0, float x1, float& a0, float& a1, float& g0, float& g1) {
    const uint64_t x = f2_pack(x0, x1);
    const uint64_t s = f2_pack(fminf(x0 * x0, 64.f), fminf(x1 * x1, 64.f));
    uint64_t p = f2_fma(s, f2_splat(kGelu3), f2_splat(kGelu2));
    p = f2_fma(p, s, f2_splat(kGelu1));
    float u0, u1;
    f2_unpack(f2_mul(x, p), u0, u1);
    const uint64_t t = f2_pack(tanh_approx(u0), tanh_approx(u1));
    const uint64_t h = f2_mul(x, f2_splat(0.5f));
    f2_unpack(f2_fma(h, t, h), a0, a1);
    if (GRAD) {
        uint64_t up = f2_fma(s, f2_splat(5.f * kGelu3), f2_splat(3.f * kGelu2));
        up = f2_fma(up, s, f2_splat(kGelu1));
        const uint64_t sech2 = f2_fma(f2_mul(t, t), f2_splat(-1.f), f2_splat(1.f));
        const uint64_t Phi = f2_fma(t, f2_splat(0.5f), f2_splat(0.5f));
        f2_unpack(f2_fma(f2_mul(h, sech2), up, Phi), g0, g1);
    }
}
// ---- the same evaluation in packed half precision: what the head kernels' epilogues run ---------------------------------------------
// The activations are rounded to fp16 anyway (they are the next GEMM's operand) and the backward re-evaluates act / act' from the fp16
// pre-activation it finds in the stash, so the whole evaluation can stay in half2: ~9 issue slots and ONE packed MUFU (tanh.approx.f16x2) per
// PAIR of activations, no fp32 <-> fp16 conversions around it.  Error against the fp32 evaluation: a few fp16 ulp of the result (measured
// in tests/test_gpu_mlp.py against torch's exact-erf GELU through the whole head: the 2e-4 bar on the outputs holds with > 3x margin).
__device__ __forceinline__ __half2 tanh_h2(__half2 x) {
    uint32_t r, a = *reinterpret_cast<uint32_t*>(&x);
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(a));
    return *reinterpret_cast<__half2*>(&r);
}
template <bool GRAD>
__device__ __forceinline__ void gelu_h2(__half2 z, __half2& a, __half2& g) {
    // x^2 clamped at 64 (one min: z^2 may round to +inf, min brings it back): beyond |z| = 8 the polynomial keeps its value at 8, the
    // tanh argument z p only grows in magnitude (p(64) > 0) and tanh stays saturated, sech^2 = 0
    const __half2 s = __hmin2(__hmul2(z, z), __float2half2_rn(64.f));
    __half2 p = __hfma2(s, __float2half2_rn(kGelu3), __float2half2_rn(kGelu2));
    p = __hfma2(p, s, __float2half2_rn(kGelu1));
    const __half2 t = tanh_h2(__hmul2(z, p));
    const __half2 h = __hmul2(z, __float2half2_rn(0.5f));
    a = __hfma2(h, t, h);
    if (GRAD) {
        __half2 up = __hfma2(s, __float2half2_rn(5.f * kGelu3), __float2half2_rn(3.f * kGelu2));
        up = __hfma2(up, s, __float2half2_rn(kGelu1));
        const __half2 sech2 = __hfma2(__hneg2(t), t, __float2half2_rn(1.f));
        const __half2 Phi = __hfma2(t, __float2half2_rn(0.5f), __float2half2_rn(0.5f));
        g = __hfma2(__hmul2(h, sech2), up, Phi);
    }
}
// ReLU twin: a = max(z, 0), g = z > 0
template <bool GRAD>
__device__ __forceinline__ void relu_h2(__half2 z, __half2& a, __half2& g) {
    a = __hmax2(z, __float2half2_rn(0.f));
    if (GRAD) g = __hgt2(z, __float2half2_rn(0.f));
}
// (lo, hi) -> half2 with saturation to the largest finite fp16 (a later product with an exact 0 must not meet an infinity)
__device__ __forceinline__ __half2 floats2half2_sat(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return *reinterpret_cast<__half2*>(&r);
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

__device__ __forceinline__ float gelu_fwd(float x) {
    float a0, a1, g0, g1;
    gelu_pair<false>(x, x, a0, a1, g0, g1);
    return a0;
}
__device__ __forceinline__ float gelu_grad(float x) {
    float a0, a1, g0, g1;
    gelu_pair<true>(x, x, a0, a1, g0, g1);
    return g0;
}

__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

__device__ __forceinline__ void sh_eval(float x, float y, float z, int n_sh, float* o) {
    // sphericalharmonics.py:103-150
    o[0] = 0.28209479177387814f;
    if (n_sh > 1) {
        o[1] = -0.4886025119029199f * y;
        o[2] = 0.4886025119029199f * z;
        o[3] = -0.4886025119029199f * x;
    }
    if (n_sh > 4) {
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        o[4] = 1.0925484305920792f * xy;
        o[5] = -1.0925484305920792f * yz;
        o[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
        o[7] = -1.0925484305920792f * xz;
        o[8] = 0.5462742152960396f * (xx - yy);
        if (n_sh > 9) {
            o[9] = -0.5900435899266435f * y * (3 * xx - yy);
            o[10] = 2.890611442640554f * xy * z;
            o[11] = -0.4570457994644658f * y * (4 * zz - xx - yy);
            o[12] = 0.3731763325901154f * z * (2 * zz - 3 * xx - 3 * yy);
            o[13] = -0.4570457994644658f * x * (4 * zz - xx - yy);
            o[14] = 1.445305721320277f * z * (xx - yy);
            o[15] = -0.5900435899266435f * x * (xx - 3 * yy);
        }
    }
}

}  // namespace vs
