// volsurfs_b200 — SH neural textures: the reference's DEFAULT appearance (config/volsurfs/base_5.cfg:11-20), SURVEY 8a row a6'.
//
// Replaces, per layer hit,
//   volsurfs_py/models/sh_neural_textures.py:64-97   one texture network per SH degree -> coefficient tensor [S,C,16] -> fp16 ->
//                                                    SHEncoder.eval (encodings/sphericalharmonics.py:156-229) -> sigmoid
//   volsurfs_py/models/neural_texture.py:81-197      uv -> texel corners (mvdatasets/utils/images.py:46-117), align_to_webgl rotation,
//                                                    4 network queries, sigmoid, 8-bit quantisation (STE), fp16 re-expansion, lerp
//   tiny-cuda-nn HashGrid encoding (neural_texture.py:54-63; 2-D, 16 levels x 2 features, 2^15 entries, base 16, scale 1.5)
// with three kinds of kernels around the tensor-core network (vs_mlp_forward_raw / vs_mlp_backward_stashed_raw, csrc/mlp*.cu):
//   hashgrid_encode_kernel   thread per query row: texel-corner uv computed on the fly from the hit's uv, 16 levels x 4 gathers from
//                            the L2-resident table (<= 2.8 MB per network) in flight, fp32 interpolation, fp16-rounded feature row
//                            [32] written as the thread's own 128-byte line
//   hashgrid_backward_*      hashed levels: vector atomics (red.global.add.v2.f32) into the table gradient; coarse dense levels: one
//                            shared-memory image of the levels per CTA (shared atomics), flushed once
//   shtex_combine_*          thread per hit: the 4 corner outputs of every degree -> sigmoid -> quantise -> fp16 expansion -> lerp ->
//                            fp16 coefficients -> mixed-precision SH evaluation -> sigmoid; the backward replays the same chain
//                            including the fp16 roundings torch autograd applies to the gradients of fp16 tensors.
// Arithmetic contract: every fp32 operation of the glue is written with round-to-nearest intrinsics (no FMA contraction) in the
// reference's operand order, so the coefficient tensor is bit-exact against the CPU restatement (oracle/shtex.py), up to the rare
// quantisation flips an ulp of sigmoid() can cause.
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "vs_common.cuh"

namespace vs {

constexpr int kHgMaxLevels = 16;
constexpr int kShMaxDeg = 3;

struct HgLevels {
    int n_levels;
    float scale[kHgMaxLevels];
    uint32_t res[kHgMaxLevels], size[kHgMaxLevels], offset[kHgMaxLevels], hashed[kHgMaxLevels];
    uint32_t total;
};

struct TexGeom {
    int mode;   // 0 anchor (texel centre), 1 lerp (4 texel corners), 2 direct (bake: uv as given; encode / backward kernels only)
    int align;  // align_to_webgl rotation (neural_texture.py:103-109, 124-130)
    int res_h, res_w;
};

// tiny-cuda-nn grid.h: grid_scale / grid_resolution / the per-level parameter count.  The scale is evaluated in double on the host and
// rounded to fp32 once (oracle/shtex.py:hashgrid_levels evaluates the same expression).
static int hg_layout(int n_levels, int log2_hashmap_size, int base_resolution, float per_level_scale, HgLevels* lv) {
    if (n_levels < 1 || n_levels > kHgMaxLevels || log2_hashmap_size < 3 || log2_hashmap_size > 28 || base_resolution < 1) return VS_ERR_INVALID_ARG;
    lv->n_levels = n_levels;
    uint64_t offset = 0;
    for (int l = 0; l < n_levels; ++l) {
        const float scale = (float)(std::pow((double)per_level_scale, (double)l) * (double)base_resolution - 1.0);
        const uint32_t res = (uint32_t)std::ceil((double)scale) + 1u;
        uint64_t dense = (uint64_t)res * res;
        uint64_t size = std::min<uint64_t>(dense, 0x7fffffffu);
        size = (size + 7) / 8 * 8;
        size = std::min<uint64_t>(size, 1ull << log2_hashmap_size);
        lv->scale[l] = scale;
        lv->res[l] = res;
        lv->size[l] = (uint32_t)size;
        lv->offset[l] = (uint32_t)offset;
        lv->hashed[l] = dense > size ? 1u : 0u;
        offset += size;
        if (offset > 0x7fffffffu) return VS_ERR_UNSUPPORTED;
    }
    lv->total = (uint32_t)offset;
    return VS_OK;
}

// uv of query row (sample s, corner k) in the texture's normalised space, exactly as torch evaluates the reference's expressions:
// images.py:54-61 (uv * flip(res)), :70-81 (floor(uv - 0.5) + corner + 0.5), :46-51 (/ flip(res)); neural_texture.py:98-111 (anchor)
__device__ __forceinline__ float2 texel_query(const TexGeom& tg, float u, float v, int k, float* lerp_w) {
    const float W = (float)tg.res_w, H = (float)tg.res_h;
    if (tg.mode == 2) {  // bake: uv is used as it is (neural_texture.py:83-86)
        if (lerp_w) *lerp_w = 1.f;
        return make_float2(u, v);
    }
    if (tg.mode == 0) {
        long long px = (long long)floorf(__fmul_rn(u, W)), py = (long long)floorf(__fmul_rn(v, H));
        if (tg.align) {
            const long long t = px;
            px = (long long)(tg.res_w - 1) - py;
            py = t;
        }
        if (lerp_w) *lerp_w = 1.f;
        return make_float2(__fdiv_rn(__fadd_rn((float)px, 0.5f), W), __fdiv_rn(__fadd_rn((float)py, 0.5f), H));
    }
    float a = __fmul_rn(u, W), b = __fmul_rn(v, H);
    if (tg.align) {
        const float t = a;
        a = __fsub_rn(W, b);  // `width - uv_nn[:, 1]` with width = res[1]
        b = t;
    }
    const float cx = __fadd_rn(floorf(__fsub_rn(a, 0.5f)), 0.5f), cy = __fadd_rn(floorf(__fsub_rn(b, 0.5f)), 0.5f);  // corner 0
    if (lerp_w) {
        const float dx = __fsub_rn(a, cx), dy = __fsub_rn(b, cy);
        const float wx = (k & 1) ? dx : __fsub_rn(1.f, dx), wy = (k & 2) ? dy : __fsub_rn(1.f, dy);
        *lerp_w = __fmul_rn(wx, wy);
    }
    // corners are (floor + offset) + 0.5 with the integer offset added BEFORE the half texel (images.py:30-43, 79)
    const float qx = __fadd_rn(__fadd_rn(floorf(__fsub_rn(a, 0.5f)), (float)(k & 1)), 0.5f);
    const float qy = __fadd_rn(__fadd_rn(floorf(__fsub_rn(b, 0.5f)), (float)((k >> 1) & 1)), 0.5f);
    return make_float2(__fdiv_rn(qx, W), __fdiv_rn(qy, H));
}

// the 4 interpolation corners of one level at x: entry index (level-relative + offset) and fp32 weight (grid.h: kernel_grid, grid_index)
__device__ __forceinline__ void hg_corners(const HgLevels& lv, int level, float2 x, uint32_t idx[4], float w[4]) {
    const float scale = lv.scale[level];
    const float px = __fadd_rn(__fmul_rn(x.x, scale), 0.5f), py = __fadd_rn(__fmul_rn(x.y, scale), 0.5f);
    const float fx = floorf(px), fy = floorf(py);
    const uint32_t gx = (uint32_t)(int)fx, gy = (uint32_t)(int)fy;
    const float rx = __fsub_rn(px, fx), ry = __fsub_rn(py, fy);
    const uint32_t res = lv.res[level], size = lv.size[level], off = lv.offset[level];
    const bool hashed = lv.hashed[level] != 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t cx = gx + (uint32_t)(c & 1), cy = gy + (uint32_t)((c >> 1) & 1);
        const float wx = (c & 1) ? rx : __fsub_rn(1.f, rx), wy = (c & 2) ? ry : __fsub_rn(1.f, ry);
        w[c] = __fmul_rn(wx, wy);
        // hashed levels own exactly 2^log2_hashmap_size entries; a dense index leaves [0, size) only at the texture border
        uint32_t h;
        if (hashed) {
            h = (cx ^ (cy * 2654435761u)) & (size - 1u);
        } else {
            h = cx + cy * res;
            if (h >= size) h %= size;
        }
        idx[c] = off + h;
    }
}

__device__ __forceinline__ float round_half(float v) { return __half2float(__float2half_rn(v)); }

// One thread per query row walks the levels: the texel query is evaluated once, the 4 x n_levels gathers of a row are independent
// loads in flight, and the thread writes its own 128-byte feature row.
__global__ void __launch_bounds__(256) hashgrid_encode_kernel(const HgLevels lv, const TexGeom tg, const float* __restrict__ uv,
                                                              const float2* __restrict__ table, float* __restrict__ feat, int64_t n_samples,
                                                              const int64_t* __restrict__ n_valid_dev) {
    int64_t n = n_samples;
    if (n_valid_dev != nullptr) n = min(n, *n_valid_dev);
    const int csh = tg.mode == 1 ? 2 : 0;
    const int L = lv.n_levels;
    const int64_t rows = n << csh;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < rows; row += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = row >> csh;
        const int k = (int)(row & ((1 << csh) - 1));
        const float2 q = texel_query(tg, __ldg(uv + 2 * s), __ldg(uv + 2 * s + 1), k, nullptr);
        float2* dst = reinterpret_cast<float2*>(feat + row * (2 * L));
#pragma unroll 4
        for (int level = 0; level < L; ++level) {
            uint32_t idx[4];
            float w[4];
            hg_corners(lv, level, q, idx, w);
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float2 e = __ldg(table + idx[c]);
                const float t0 = __fmul_rn(w[c], round_half(e.x)), t1 = __fmul_rn(w[c], round_half(e.y));
                a0 = c == 0 ? t0 : __fadd_rn(a0, t0);
                a1 = c == 0 ? t1 : __fadd_rn(a1, t1);
            }
            __stcs(dst + level, make_float2(round_half(a0), round_half(a1)));
        }
    }
}

// Backward, levels [lv_begin, lv_end) with atomics straight into the table gradient (the hashed levels: 2^15 entries each, collisions
// between the threads of a warp are rare).
__global__ void __launch_bounds__(256) hashgrid_backward_kernel(const HgLevels lv, const TexGeom tg, const float* __restrict__ uv,
                                                                const float* __restrict__ d_feat, float2* __restrict__ d_table,
                                                                int64_t n_samples, const int64_t* __restrict__ n_valid_dev, int lv_begin,
                                                                int lv_end) {
    int64_t n = n_samples;
    if (n_valid_dev != nullptr) n = min(n, *n_valid_dev);
    const int csh = tg.mode == 1 ? 2 : 0;
    const int L = lv.n_levels;
    const int64_t rows = n << csh;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < rows; row += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = row >> csh;
        const int k = (int)(row & ((1 << csh) - 1));
        const float2 q = texel_query(tg, __ldg(uv + 2 * s), __ldg(uv + 2 * s + 1), k, nullptr);
        const float2* src = reinterpret_cast<const float2*>(d_feat + row * (2 * L));
#pragma unroll 2
        for (int level = lv_begin; level < lv_end; ++level) {
            const float2 g = __ldg(src + level);
            if (g.x == 0.f && g.y == 0.f) continue;
            uint32_t idx[4];
            float w[4];
            hg_corners(lv, level, q, idx, w);
#pragma unroll
            for (int c = 0; c < 4; ++c) atomicAdd(d_table + idx[c], make_float2(w[c] * g.x, w[c] * g.y));
        }
    }
}

// Backward, the coarse (dense) levels [0, lv_end): a level with 256 entries receives rows*4 contributions, i.e. millions of atomics per
// address if they go to global memory.  One persistent CTA per SM accumulates its share of the rows in a shared-memory image of those
// levels (26 504 entries x 8 B = 207 KB for the reference's grid: levels 0-5) and flushes the image once.
__global__ void __launch_bounds__(1024, 1) hashgrid_backward_dense_kernel(const HgLevels lv, const TexGeom tg, const float* __restrict__ uv,
                                                                          const float* __restrict__ d_feat, float* __restrict__ d_table,
                                                                          int64_t n_samples, const int64_t* __restrict__ n_valid_dev,
                                                                          int lv_end, int n_entries) {
    extern __shared__ float acc[];  // [n_entries][2]
    for (int i = threadIdx.x; i < 2 * n_entries; i += blockDim.x) acc[i] = 0.f;
    __syncthreads();
    int64_t n = n_samples;
    if (n_valid_dev != nullptr) n = min(n, *n_valid_dev);
    const int csh = tg.mode == 1 ? 2 : 0;
    const int L = lv.n_levels;
    const int64_t rows = n << csh;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < rows; row += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = row >> csh;
        const int k = (int)(row & ((1 << csh) - 1));
        const float2 q = texel_query(tg, __ldg(uv + 2 * s), __ldg(uv + 2 * s + 1), k, nullptr);
        const float2* src = reinterpret_cast<const float2*>(d_feat + row * (2 * L));
        for (int level = 0; level < lv_end; ++level) {
            const float2 g = __ldg(src + level);
            if (g.x == 0.f && g.y == 0.f) continue;
            uint32_t idx[4];
            float w[4];
            hg_corners(lv, level, q, idx, w);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                atomicAdd(acc + 2 * idx[c], w[c] * g.x);
                atomicAdd(acc + 2 * idx[c] + 1, w[c] * g.y);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * n_entries; i += blockDim.x) {
        const float v = acc[i];
        if (v != 0.f) atomicAdd(d_table + i, v);
    }
}

// Dense levels in lerp mode, one thread per HIT: the 4 texel corners of a hit are neighbouring texels, and on the coarse (dense) levels
// they fall into the same grid cell most of the time, i.e. they update the SAME 4 entries — the thread adds ONE combined contribution
// per entry: up to 4x fewer shared-memory atomics on exactly the levels where they collide most.
struct HgCell {
    uint32_t gx, gy;
    float rx, ry;
};
__device__ __forceinline__ HgCell hg_cell(const HgLevels& lv, int level, float2 x) {
    const float scale = lv.scale[level];
    const float px = __fadd_rn(__fmul_rn(x.x, scale), 0.5f), py = __fadd_rn(__fmul_rn(x.y, scale), 0.5f);
    const float fx = floorf(px), fy = floorf(py);
    return HgCell{(uint32_t)(int)fx, (uint32_t)(int)fy, __fsub_rn(px, fx), __fsub_rn(py, fy)};
}
__device__ __forceinline__ void hg_dense_entries(const HgLevels& lv, int level, const HgCell& c, uint32_t idx[4]) {
    const uint32_t res = lv.res[level], size = lv.size[level], off = lv.offset[level];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t h = (c.gx + (uint32_t)(k & 1)) + (c.gy + (uint32_t)((k >> 1) & 1)) * res;
        if (h >= size) h %= size;
        idx[k] = off + h;
    }
}
__device__ __forceinline__ void hg_cell_weights(const HgCell& c, float w[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float wx = (k & 1) ? c.rx : 1.f - c.rx, wy = (k & 2) ? c.ry : 1.f - c.ry;
        w[k] = wx * wy;
    }
}

__global__ void __launch_bounds__(1024, 1) hashgrid_backward_dense_lerp_kernel(const HgLevels lv, const TexGeom tg, const float* __restrict__ uv,
                                                                               const float* __restrict__ d_feat, float* __restrict__ d_table,
                                                                               int64_t n_samples, const int64_t* __restrict__ n_valid_dev,
                                                                               int lv_end, int n_entries) {
    extern __shared__ float acc[];  // [n_entries][2]
    for (int i = threadIdx.x; i < 2 * n_entries; i += blockDim.x) acc[i] = 0.f;
    __syncthreads();
    int64_t n = n_samples;
    if (n_valid_dev != nullptr) n = min(n, *n_valid_dev);
    const int L = lv.n_levels;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x) {
        const float u = __ldg(uv + 2 * s), v = __ldg(uv + 2 * s + 1);
        float2 q[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) q[k] = texel_query(tg, u, v, k, nullptr);
        const float2* src = reinterpret_cast<const float2*>(d_feat + (4 * s) * (2 * L));
        for (int level = 0; level < lv_end; ++level) {
            HgCell c[4];
            float2 g[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                c[k] = hg_cell(lv, level, q[k]);
                g[k] = __ldg(src + k * L + level);
            }
            const bool same = c[1].gx == c[0].gx && c[1].gy == c[0].gy && c[2].gx == c[0].gx && c[2].gy == c[0].gy && c[3].gx == c[0].gx &&
                              c[3].gy == c[0].gy;
            uint32_t idx[4];
            if (same) {
                float2 sum[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float w[4];
                    hg_cell_weights(c[k], w);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        sum[j].x = fmaf(w[j], g[k].x, sum[j].x);
                        sum[j].y = fmaf(w[j], g[k].y, sum[j].y);
                    }
                }
                hg_dense_entries(lv, level, c[0], idx);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (sum[j].x != 0.f) atomicAdd(acc + 2 * idx[j], sum[j].x);
                    if (sum[j].y != 0.f) atomicAdd(acc + 2 * idx[j] + 1, sum[j].y);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (g[k].x == 0.f && g[k].y == 0.f) continue;
                    float w[4];
                    hg_cell_weights(c[k], w);
                    hg_dense_entries(lv, level, c[k], idx);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        atomicAdd(acc + 2 * idx[j], w[j] * g[k].x);
                        atomicAdd(acc + 2 * idx[j] + 1, w[j] * g[k].y);
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * n_entries; i += blockDim.x) {
        const float v = acc[i];
        if (v != 0.f) atomicAdd(d_table + i, v);
    }
}

// ---- coefficient assembly + SH evaluation ----------------------------------------------------------------------------------------
struct ShTexConfig {
    int sh_deg, nr_channels, n_coeffs;
    int squeeze, quantize;
    TexGeom geom[kShMaxDeg + 1];
    float range_mul[kShMaxDeg + 1];  // (hi - lo) as torch passes it to mul(half, Scalar): kept in fp32
    float range_lo[kShMaxDeg + 1];   // lo as torch passes it to add(half, Scalar): rounded to fp16 first
    const float* raw[kShMaxDeg + 1]; // network outputs [rows, C*(2g+1)] fp32
    float* d_raw[kShMaxDeg + 1];
};

__device__ __forceinline__ float sigmoid_precise(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }

// SH basis factors P_k of sphericalharmonics.py:183-215 in torch's evaluation order (python scalars are fp32 factors)
__device__ __forceinline__ void sh_factors(float x, float y, float z, int deg, float* P) {
    P[0] = 0.28209479177387814f;
    if (deg > 0) {
        const float C1 = 0.4886025119029199f;
        P[1] = __fmul_rn(C1, y);
        P[2] = __fmul_rn(C1, z);
        P[3] = __fmul_rn(C1, x);
    }
    if (deg > 1) {
        const float xx = __fmul_rn(x, x), yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
        const float xy = __fmul_rn(x, y), yz = __fmul_rn(y, z), xz = __fmul_rn(x, z);
        P[4] = __fmul_rn(1.0925484305920792f, xy);
        P[5] = __fmul_rn(-1.0925484305920792f, yz);
        P[6] = __fmul_rn(0.31539156525252005f, __fsub_rn(__fsub_rn(__fmul_rn(2.0f, zz), xx), yy));
        P[7] = __fmul_rn(-1.0925484305920792f, xz);
        P[8] = __fmul_rn(0.5462742152960396f, __fsub_rn(xx, yy));
        if (deg > 2) {
            P[9] = __fmul_rn(__fmul_rn(-0.5900435899266435f, y), __fsub_rn(__fmul_rn(3.f, xx), yy));
            P[10] = __fmul_rn(__fmul_rn(2.890611442640554f, xy), z);
            P[11] = __fmul_rn(__fmul_rn(-0.4570457994644658f, y), __fsub_rn(__fsub_rn(__fmul_rn(4.f, zz), xx), yy));
            P[12] = __fmul_rn(__fmul_rn(0.3731763325901154f, z), __fsub_rn(__fsub_rn(__fmul_rn(2.f, zz), __fmul_rn(3.f, xx)), __fmul_rn(3.f, yy)));
            P[13] = __fmul_rn(__fmul_rn(-0.4570457994644658f, x), __fsub_rn(__fsub_rn(__fmul_rn(4.f, zz), xx), yy));
            P[14] = __fmul_rn(__fmul_rn(1.445305721320277f, z), __fsub_rn(xx, yy));
            P[15] = __fmul_rn(__fmul_rn(-0.5900435899266435f, x), __fsub_rn(xx, __fmul_rn(3.f, yy)));
        }
    }
}

// one texel value: network output -> (sigmoid -> quantise) -> fp16 -> expansion to [lo, hi] in fp16 (neural_texture.py:152-181)
__device__ __forceinline__ float texel_value(const ShTexConfig& c, int g, float raw, float* s_out) {
    float o = round_half(raw);  // tiny-cuda-nn returns fp16
    float s = 0.f;
    if (c.squeeze) {
        s = sigmoid_precise(o);
        o = s;
        if (c.quantize) o = __fdiv_rn(rintf(__fmul_rn(o, 255.0f)), 255.0f);
    }
    if (s_out) *s_out = s;
    float h = round_half(o);
    if (c.squeeze) {
        h = round_half(__fmul_rn(c.range_mul[g], h));
        h = round_half(__fadd_rn(c.range_lo[g], h));
    }
    return h;
}

// One warp owns 32 consecutive hits.  Their network outputs are ONE contiguous chunk of every raw[g] (rows 4*s0 .. 4*s0+127, all
// columns), so the warp copies the chunks into shared memory with unit-stride loads, every lane then works on its own hit out of shared
// memory, and (backward) the gradients — written over the values they were computed from — leave as unit-stride stores.
constexpr int kCombineWarps = 2;

template <bool BWD>
__global__ void __launch_bounds__(32 * kCombineWarps) shtex_combine_kernel(const ShTexConfig cfg, const float* __restrict__ uv,
                                                                           const float* __restrict__ dirs, float* __restrict__ coeffs,
                                                                           float* __restrict__ out, const float* __restrict__ g_out,
                                                                           const float* __restrict__ g_coeffs, int64_t n_samples,
                                                                           const int64_t* __restrict__ n_valid_dev, int warp_floats) {
    extern __shared__ __align__(16) float sm_all[];
    int64_t n = n_samples;
    if (n_valid_dev != nullptr) n = min(n, *n_valid_dev);
    const int C = cfg.nr_channels, NC = cfg.n_coeffs, deg = cfg.sh_deg;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* sm = sm_all + (size_t)warp * warp_floats;
    int corners[kShMaxDeg + 1], chunk[kShMaxDeg + 1], soff[kShMaxDeg + 1];  // per hit: floats of degree g; start of the degree's image
    {
        int off = 0;
#pragma unroll
        for (int g = 0; g <= kShMaxDeg; ++g) {
            corners[g] = chunk[g] = soff[g] = 0;
            if (g > deg) continue;
            corners[g] = cfg.geom[g].mode == 1 ? 4 : 1;
            chunk[g] = corners[g] * C * (2 * g + 1);
            soff[g] = off;
            off += 32 * chunk[g];
        }
    }
    const int64_t warp_stride = (int64_t)gridDim.x * kCombineWarps * 32;
    for (int64_t s0 = ((int64_t)blockIdx.x * kCombineWarps + warp) * 32; s0 < n; s0 += warp_stride) {
        const int cnt = (int)min((int64_t)32, n - s0);
        // ---- network outputs of the warp's hits -> shared memory (unit stride)
#pragma unroll
        for (int g = 0; g <= kShMaxDeg; ++g) {
            if (g > deg) continue;
            const float* src = cfg.raw[g] + s0 * chunk[g];
            const int total = cnt * chunk[g];
            if ((chunk[g] & 3) == 0) {
                const float4* src4 = reinterpret_cast<const float4*>(src);
                float4* dst4 = reinterpret_cast<float4*>(sm + soff[g]);
                for (int i = lane; i < total / 4; i += 32) dst4[i] = __ldcs(src4 + i);
            } else {
                for (int i = lane; i < total; i += 32) sm[soff[g] + i] = __ldcs(src + i);
            }
        }
        __syncwarp();
        const int64_t s = s0 + lane;
        if (lane < cnt) {
            const float u = __ldg(uv + 2 * s), v = __ldg(uv + 2 * s + 1);
            float lw[kShMaxDeg + 1][4];
#pragma unroll
            for (int g = 0; g <= kShMaxDeg; ++g) {
                if (g > deg) continue;
                for (int k = 0; k < corners[g]; ++k) texel_query(cfg.geom[g], u, v, k, &lw[g][k]);
            }
            float P[16];
            if (dirs != nullptr) sh_factors(__ldg(dirs + 3 * s), __ldg(dirs + 3 * s + 1), __ldg(dirs + 3 * s + 2), deg, P);

            for (int ch = 0; ch < C; ++ch) {
                // ---- forward: coefficients of this channel
                float co[16];
#pragma unroll
                for (int g = 0; g <= kShMaxDeg; ++g) {
                    if (g > deg) continue;
                    const int nm = 2 * g + 1, width = C * nm;
                    const float* raw = sm + soff[g] + lane * chunk[g] + ch * nm;
#pragma unroll
                    for (int m = 0; m < 2 * kShMaxDeg + 1; ++m) {
                        if (m >= nm) continue;
                        float acc = 0.f;
                        if (corners[g] == 4) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float t = __fmul_rn(texel_value(cfg, g, raw[k * width + m], nullptr), lw[g][k]);
                                acc = k == 0 ? t : __fadd_rn(acc, t);
                            }
                        } else {
                            acc = texel_value(cfg, g, raw[m], nullptr);
                        }
                        co[g * g + m] = acc;
                    }
                }
                if (!BWD && coeffs != nullptr) {
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        if (k < NC) coeffs[(s * C + ch) * NC + k] = co[k];
                }
                if (dirs == nullptr && !BWD) continue;

                float d_co[16];  // gradient of the fp32 coefficient tensor
                if (dirs != nullptr) {
                    // SHEncoder.eval on fp16 coefficients: the degree-0 term is an fp16 product, everything after it is fp32
                    float r = round_half(__fmul_rn(round_half(co[0]), P[0]));
                    if (deg > 0) {
                        r = __fsub_rn(r, __fmul_rn(P[1], round_half(co[1])));
                        r = __fadd_rn(r, __fmul_rn(P[2], round_half(co[2])));
                        r = __fsub_rn(r, __fmul_rn(P[3], round_half(co[3])));
                    }
#pragma unroll
                    for (int k = 4; k < 16; ++k)
                        if (k < NC) r = __fadd_rn(r, __fmul_rn(P[k], round_half(co[k])));
                    if (!BWD) {
                        // sh_deg == 0: the result is still an fp16 tensor when it reaches torch.sigmoid
                        const float o = sigmoid_precise(r);
                        out[s * C + ch] = deg == 0 ? round_half(o) : o;
                        continue;
                    }
                    // sigmoid backward (grad * (1 - y)) * y (in fp16 for sh_deg == 0), then the fp16 gradients of the fp16 coefficients
                    const float y = __ldg(out + s * C + ch);
                    float gr = __ldg(g_out + s * C + ch);
                    if (deg == 0) gr = round_half(gr);
                    gr = __fmul_rn(__fmul_rn(gr, __fsub_rn(1.f, y)), y);
                    if (deg == 0) gr = round_half(gr);
                    d_co[0] = round_half(__fmul_rn(round_half(gr), P[0]));
#pragma unroll
                    for (int k = 1; k < 16; ++k)
                        if (k < NC) {
                            const float t = round_half(__fmul_rn(gr, P[k]));
                            d_co[k] = (k == 1 || k == 3) ? -t : t;
                        }
                } else {
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        if (k < NC) d_co[k] = __ldg(g_coeffs + (s * C + ch) * NC + k);
                }
                // ---- backward through lerp / expansion / quantisation (STE) / sigmoid, per texel; the gradient replaces the value
#pragma unroll
                for (int g = 0; g <= kShMaxDeg; ++g) {
                    if (g > deg) continue;
                    const int nm = 2 * g + 1, width = C * nm;
                    float* raw = sm + soff[g] + lane * chunk[g] + ch * nm;
#pragma unroll
                    for (int m = 0; m < 2 * kShMaxDeg + 1; ++m) {
                        if (m >= nm) continue;
                        const float dc = d_co[g * g + m];
                        for (int k = 0; k < corners[g]; ++k) {
                            float gh = corners[g] == 4 ? round_half(__fmul_rn(dc, lw[g][k])) : round_half(dc);
                            float sg = 0.f;
                            if (cfg.squeeze) {
                                texel_value(cfg, g, raw[k * width + m], &sg);
                                gh = round_half(__fmul_rn(gh, cfg.range_mul[g]));
                                if (cfg.quantize) gh = __fmul_rn(__fdiv_rn(gh, 255.0f), 255.0f);
                                gh = __fmul_rn(__fmul_rn(gh, __fsub_rn(1.f, sg)), sg);
                            }
                            raw[k * width + m] = round_half(gh);  // the network's output is fp16: so is its gradient
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (BWD) {
#pragma unroll
            for (int g = 0; g <= kShMaxDeg; ++g) {
                if (g > deg) continue;
                float* dst = cfg.d_raw[g] + s0 * chunk[g];
                const int total = cnt * chunk[g];
                if ((chunk[g] & 3) == 0) {
                    float4* dst4 = reinterpret_cast<float4*>(dst);
                    const float4* src4 = reinterpret_cast<const float4*>(sm + soff[g]);
                    for (int i = lane; i < total / 4; i += 32) __stcs(dst4 + i, src4[i]);
                } else {
                    for (int i = lane; i < total; i += 32) __stcs(dst + i, sm[soff[g] + i]);
                }
            }
            __syncwarp();
        }
    }
}

// shared memory of one warp: the network outputs of its 32 hits, all degrees
static int combine_warp_floats(const ShTexConfig& c) {
    int f = 0;
    for (int g = 0; g <= c.sh_deg; ++g) f += 32 * (c.geom[g].mode == 1 ? 4 : 1) * c.nr_channels * (2 * g + 1);
    return (f + 3) & ~3;
}

static int grid_for(int64_t work, int threads) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (int)std::min<int64_t>(std::max<int64_t>(div_up(work, threads), 1), (int64_t)sms * 16);
}

}  // namespace vs

using namespace vs;

extern "C" {

// Level table of the tiny-cuda-nn HashGrid (2-D): HOST arrays of n_levels entries (any may be NULL); returns the total number of
// table entries (each n_features = 2 floats) or < 0.
int64_t vs_hashgrid_levels(int n_levels, int log2_hashmap_size, int base_resolution, float per_level_scale, float* scale, int32_t* res,
                           int32_t* size, int32_t* offset) {
    HgLevels lv;
    int e = hg_layout(n_levels, log2_hashmap_size, base_resolution, per_level_scale, &lv);
    if (e != VS_OK) return e;
    for (int l = 0; l < n_levels; ++l) {
        if (scale) scale[l] = lv.scale[l];
        if (res) res[l] = (int32_t)lv.res[l];
        if (size) size[l] = (int32_t)lv.size[l];
        if (offset) offset[l] = (int32_t)lv.offset[l];
    }
    return lv.total;
}

// features [rows, 2*n_levels] fp32 (fp16-representable), rows = n_samples * (mode == 1 ? 4 : 1), row = sample*corners + corner:
// the hash-grid encoding of every texel query of NeuralTexture.forward (neural_texture.py:96-150).
//   mode 0: anchor (texel centre)  1: lerp (4 corners)  2: uv as given (bake) · align: align_to_webgl · res_h/res_w: texture resolution
//   uv [n_samples,2] fp32 · table [total_entries,2] fp32 master parameters (rounded to fp16 on the fly)
int vs_hashgrid_forward(int n_levels, int log2_hashmap_size, int base_resolution, float per_level_scale, int mode, int align, int res_h,
                        int res_w, const float* uv, const float* table, float* features, int64_t n_samples, const int64_t* n_valid_dev,
                        void* stream) {
    VS_CHECK_ARG(n_samples >= 0 && mode >= 0 && mode <= 2 && res_h > 0 && res_w > 0);
    HgLevels lv;
    int e = hg_layout(n_levels, log2_hashmap_size, base_resolution, per_level_scale, &lv);
    if (e != VS_OK) return e;
    if (n_samples == 0) return VS_OK;
    VS_CHECK_ARG(uv && table && features);
    VS_CHECK_ARG((reinterpret_cast<uintptr_t>(table) & 7) == 0 && (reinterpret_cast<uintptr_t>(features) & 7) == 0);
    const TexGeom tg{mode, align ? 1 : 0, res_h, res_w};
    const int64_t work = n_samples * (mode == 1 ? 4 : 1);
    hashgrid_encode_kernel<<<grid_for(work, 256), 256, 0, (cudaStream_t)stream>>>(lv, tg, uv, reinterpret_cast<const float2*>(table), features,
                                                                                  n_samples, n_valid_dev);
    return launched(1);
}

// d_table [total_entries,2] fp32 += scatter of d_features (ACCUMULATES: zero it first).  Same arguments as vs_hashgrid_forward.
int vs_hashgrid_backward(int n_levels, int log2_hashmap_size, int base_resolution, float per_level_scale, int mode, int align, int res_h,
                         int res_w, const float* uv, const float* d_features, float* d_table, int64_t n_samples, const int64_t* n_valid_dev,
                         void* stream) {
    VS_CHECK_ARG(n_samples >= 0 && mode >= 0 && mode <= 2 && res_h > 0 && res_w > 0);
    HgLevels lv;
    int e = hg_layout(n_levels, log2_hashmap_size, base_resolution, per_level_scale, &lv);
    if (e != VS_OK) return e;
    if (n_samples == 0) return VS_OK;
    VS_CHECK_ARG(uv && d_features && d_table);
    VS_CHECK_ARG((reinterpret_cast<uintptr_t>(d_table) & 7) == 0 && (reinterpret_cast<uintptr_t>(d_features) & 7) == 0);
    const TexGeom tg{mode, align ? 1 : 0, res_h, res_w};
    const int corners = mode == 1 ? 4 : 1;
    // the leading dense levels whose shared-memory image fits one CTA go through the accumulating kernel
    int dense_end = 0, dense_entries = 0;
    while (dense_end < n_levels && !lv.hashed[dense_end] && (int64_t)(dense_entries + lv.size[dense_end]) * 8 <= 216 * 1024) {
        dense_entries += (int)lv.size[dense_end];
        ++dense_end;
    }
    if (const char* env = std::getenv("VS_HASHGRID_BWD_DENSE")) {  // A/B knob: 0 = everything through global atomics
        if (std::atoi(env) == 0) dense_end = dense_entries = 0;
    }
    int launches = 0;
    if (dense_end > 0) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int smem = dense_entries * 8;
        cudaError_t ce = cudaFuncSetAttribute(hashgrid_backward_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (ce != cudaSuccess) return (int)ce;
        static const bool per_row = std::getenv("VS_HASHGRID_DENSE_PER_ROW") != nullptr;  // A/B knob
        if (mode == 1 && !per_row) {
            ce = cudaFuncSetAttribute(hashgrid_backward_dense_lerp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (ce != cudaSuccess) return (int)ce;
            const int grid = (int)std::min<int64_t>(std::max<int64_t>(div_up(n_samples, 1024), 1), sms);
            hashgrid_backward_dense_lerp_kernel<<<grid, 1024, smem, (cudaStream_t)stream>>>(lv, tg, uv, d_features, d_table, n_samples,
                                                                                            n_valid_dev, dense_end, dense_entries);
        } else {
            const int64_t work = n_samples * corners;
            const int grid = (int)std::min<int64_t>(std::max<int64_t>(div_up(work, 1024), 1), sms);
            hashgrid_backward_dense_kernel<<<grid, 1024, smem, (cudaStream_t)stream>>>(lv, tg, uv, d_features, d_table, n_samples, n_valid_dev,
                                                                                       dense_end, dense_entries);
        }
        ++launches;
    }
    if (dense_end < n_levels) {
        const int64_t work = n_samples * corners;
        hashgrid_backward_kernel<<<grid_for(work, 256), 256, 0, (cudaStream_t)stream>>>(lv, tg, uv, d_features, reinterpret_cast<float2*>(d_table),
                                                                                        n_samples, n_valid_dev, dense_end, n_levels);
        ++launches;
    }
    return launched(launches);
}

static int shtex_config(int sh_deg, int nr_channels, int mode, int align, const int* res_hw, const float* sh_range_lo,
                        const float* sh_range_hi, int squeeze, int quantize, ShTexConfig* c) {
    if (sh_deg < 0 || sh_deg > kShMaxDeg || nr_channels < 1 || nr_channels * (2 * sh_deg + 1) > 32 || !(mode == 0 || mode == 1) || !res_hw) return VS_ERR_INVALID_ARG;
    if (squeeze && (!sh_range_lo || !sh_range_hi)) return VS_ERR_INVALID_ARG;
    if (quantize && !squeeze) return VS_ERR_INVALID_ARG;  // sh_neural_textures.py:33-37
    std::memset(c, 0, sizeof(*c));
    c->sh_deg = sh_deg;
    c->nr_channels = nr_channels;
    c->n_coeffs = (sh_deg + 1) * (sh_deg + 1);
    c->squeeze = squeeze ? 1 : 0;
    c->quantize = quantize ? 1 : 0;
    for (int g = 0; g <= sh_deg; ++g) {
        if (res_hw[2 * g] <= 0 || res_hw[2 * g + 1] <= 0) return VS_ERR_INVALID_ARG;
        c->geom[g] = TexGeom{mode, align ? 1 : 0, res_hw[2 * g], res_hw[2 * g + 1]};
        if (squeeze) {
            // val_range[0] + (val_range[1] - val_range[0]) * output on an fp16 tensor: python evaluates hi - lo in double; torch keeps the
            // multiplier in fp32 and rounds the addend to fp16 (measured on torch 2.11 CPU; oracle/shtex.py runs the very expression)
            c->range_mul[g] = (float)((double)sh_range_hi[g] - (double)sh_range_lo[g]);
            c->range_lo[g] = __half2float(__float2half_rn(sh_range_lo[g]));
        }
    }
    return VS_OK;
}

// SHNeuralTextures.forward after the networks (sh_neural_textures.py:71-97 + neural_texture.py:152-195).
//   raw: HOST array of sh_deg+1 DEVICE pointers, raw[g] = [n_samples*corners, C*(2g+1)] fp32 network outputs (row = sample*corners + k)
//   res_hw: HOST [sh_deg+1][2] (height, width) · range_lo/hi: HOST [sh_deg+1] val_range of every degree (NULL without squeeze)
//   coeffs: [n_samples, C, (sh_deg+1)^2] fp32 or NULL · dirs [n_samples,3] + out [n_samples,C] or both NULL (view_dirs=None)
int vs_shtex_combine_forward(int sh_deg, int nr_channels, int mode, int align, const int* res_hw, const float* range_lo, const float* range_hi,
                             int squeeze, int quantize, const float* uv, const float* dirs, const float* const* raw, float* coeffs, float* out,
                             int64_t n_samples, const int64_t* n_valid_dev, void* stream) {
    VS_CHECK_ARG(n_samples >= 0 && raw);
    ShTexConfig c;
    int e = shtex_config(sh_deg, nr_channels, mode, align, res_hw, range_lo, range_hi, squeeze, quantize, &c);
    if (e != VS_OK) return e;
    if (n_samples == 0) return VS_OK;
    VS_CHECK_ARG(uv && (coeffs || out) && ((dirs == nullptr) == (out == nullptr)));
    for (int g = 0; g <= sh_deg; ++g) {
        VS_CHECK_ARG(raw[g]);
        c.raw[g] = raw[g];
    }
    const int wf = combine_warp_floats(c), smem = wf * 4 * kCombineWarps;
    cudaError_t ce = cudaFuncSetAttribute(shtex_combine_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (ce != cudaSuccess) return (int)ce;
    shtex_combine_kernel<false><<<grid_for(n_samples, 32 * kCombineWarps), 32 * kCombineWarps, smem, (cudaStream_t)stream>>>(
        c, uv, dirs, coeffs, out, nullptr, nullptr, n_samples, n_valid_dev, wf);
    return launched(1);
}

// Backward of vs_shtex_combine_forward: d_raw[g] (same shapes as raw[g], fully overwritten for live rows) from g_out [n_samples,C] (with
// dirs + the forward's `out`) or, when dirs == NULL, from g_coeffs [n_samples,C,(sh_deg+1)^2].  Gradients of fp16 tensors are rounded
// to fp16 where torch autograd rounds them.
int vs_shtex_combine_backward(int sh_deg, int nr_channels, int mode, int align, const int* res_hw, const float* range_lo, const float* range_hi,
                              int squeeze, int quantize, const float* uv, const float* dirs, const float* const* raw, const float* out,
                              const float* g_out, const float* g_coeffs, float* const* d_raw, int64_t n_samples, const int64_t* n_valid_dev,
                              void* stream) {
    VS_CHECK_ARG(n_samples >= 0 && raw && d_raw);
    ShTexConfig c;
    int e = shtex_config(sh_deg, nr_channels, mode, align, res_hw, range_lo, range_hi, squeeze, quantize, &c);
    if (e != VS_OK) return e;
    if (n_samples == 0) return VS_OK;
    VS_CHECK_ARG(uv && (dirs ? (out && g_out) : g_coeffs != nullptr));
    for (int g = 0; g <= sh_deg; ++g) {
        VS_CHECK_ARG(raw[g] && d_raw[g]);
        c.raw[g] = raw[g];
        c.d_raw[g] = d_raw[g];
    }
    const int wf = combine_warp_floats(c), smem = wf * 4 * kCombineWarps;
    cudaError_t ce = cudaFuncSetAttribute(shtex_combine_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (ce != cudaSuccess) return (int)ce;
    shtex_combine_kernel<true><<<grid_for(n_samples, 32 * kCombineWarps), 32 * kCombineWarps, smem, (cudaStream_t)stream>>>(
        c, uv, dirs, nullptr, const_cast<float*>(out), g_out, g_coeffs, n_samples, n_valid_dev, wf);
    return launched(1);
}

}  // extern "C"

// =====================================================================================================================================
// Baked-texture mesh renderer (second half): volsurfs_py/renderers/mesh_renderer.py:112-201 (render_rays) + :62-110 (shade)
// after the mesh trace — texture coordinates from the barycentrics, bilinear lookup of the baked SH-coefficient texture
// (mvdatasets/utils/tensor_texture.py:66-96 with lerp=True: zero-padded image, corners / weights of images.py:30-43,70-81,97-118), fp16
// coefficients, mixed-precision SH evaluation (encodings/sphericalharmonics.py:156-229), sigmoid, and the shaded output buffers — in ONE
// launch, thread per ray (the reference: ~40 torch kernels and boolean-mask gathers).  Same evaluation order as torch, operation by
// operation; the bilinear sum runs k = 0..3 in sequence.
// =====================================================================================================================================
namespace vs {

__global__ void __launch_bounds__(256) baked_shade_kernel(const uint8_t* __restrict__ is_hit, const int64_t* __restrict__ tri_id,
                                                          const float* __restrict__ bary, const float* __restrict__ dirs,
                                                          const float* __restrict__ normals, const float* __restrict__ face_uvs,
                                                          const float* __restrict__ tex, int res_h, int res_w, int nr_coeffs, float bg_r,
                                                          float bg_g, float bg_b, float* __restrict__ o_hit, float* __restrict__ o_normals,
                                                          float* __restrict__ o_uvs, float* __restrict__ o_rgb, float* __restrict__ o_alpha,
                                                          float* __restrict__ o_dirs, int64_t n_rays) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const float bg[3] = {bg_r, bg_g, bg_b};
    if (!is_hit[r]) {
        o_hit[r] = 0.f;
        o_alpha[r] = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) o_normals[3 * r + k] = o_uvs[3 * r + k] = o_rgb[3 * r + k] = o_dirs[3 * r + k] = bg[k];
        return;
    }
    const float dx = __ldg(dirs + 3 * r), dy = __ldg(dirs + 3 * r + 1), dz = __ldg(dirs + 3 * r + 2);
    o_hit[r] = 1.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {  // (x + 1) * 0.5 (mesh_renderer.py:76,103)
        o_normals[3 * r + k] = __fmul_rn(__fadd_rn(__ldg(normals + 3 * r + k), 1.f), 0.5f);
        o_dirs[3 * r + k] = __fmul_rn(__fadd_rn(__ldg(dirs + 3 * r + k), 1.f), 0.5f);
    }
    // uv = sum_j barycentric_j * face_uv_j (mesh_renderer.py:147-150), torch's order ((b0 uv0) + (b1 uv1)) + (b2 uv2)
    const float b0 = __ldg(bary + 3 * r), b1 = __ldg(bary + 3 * r + 1), b2 = __ldg(bary + 3 * r + 2);
    const float2* f = reinterpret_cast<const float2*>(face_uvs) + 3 * __ldg(tri_id + r);
    const float2 u0 = __ldg(f), u1 = __ldg(f + 1), u2 = __ldg(f + 2);
    const float u = __fadd_rn(__fadd_rn(__fmul_rn(b0, u0.x), __fmul_rn(b1, u1.x)), __fmul_rn(b2, u2.x));
    const float v = __fadd_rn(__fadd_rn(__fmul_rn(b0, u0.y), __fmul_rn(b1, u1.y)), __fmul_rn(b2, u2.y));
    o_uvs[3 * r] = u;
    o_uvs[3 * r + 1] = v;
    o_uvs[3 * r + 2] = 0.f;
    // bilinear corners and weights in the texture's non-normalised space (tensor_texture.py:70-86)
    const float a = __fmul_rn(u, (float)res_w), b = __fmul_rn(v, (float)res_h);
    const float fx = floorf(__fsub_rn(a, 0.5f)), fy = floorf(__fsub_rn(b, 0.5f));
    const float ddx = __fsub_rn(a, __fadd_rn(fx, 0.5f)), ddy = __fsub_rn(b, __fadd_rn(fy, 0.5f));
    const float w[4] = {__fmul_rn(__fsub_rn(1.f, ddx), __fsub_rn(1.f, ddy)), __fmul_rn(ddx, __fsub_rn(1.f, ddy)),
                        __fmul_rn(__fsub_rn(1.f, ddx), ddy), __fmul_rn(ddx, ddy)};
    const int C4 = 4 * nr_coeffs, pw = res_w + 2;
    const float* texel[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // floor(corner) + 1 in the zero-padded image (tensor_texture.py:88-92)
        long long px = (long long)floorf(__fadd_rn(__fadd_rn(fx, (float)(k & 1)), 0.5f)) + 1;
        long long py = (long long)floorf(__fadd_rn(__fadd_rn(fy, (float)(k >> 1)), 0.5f)) + 1;
        px = min(max(px, 0LL), (long long)res_w + 1);
        py = min(max(py, 0LL), (long long)res_h + 1);
        texel[k] = tex + ((size_t)py * pw + (size_t)px) * C4;
    }
    const int deg = nr_coeffs == 1 ? 0 : (nr_coeffs == 4 ? 1 : (nr_coeffs == 9 ? 2 : 3));
    float P[16];
    sh_factors(dx, dy, dz, deg, P);
    float rgba[4];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        float co[16];
        for (int k = 0; k < nr_coeffs; ++k) {
            const int c = ch * nr_coeffs + k;  // sh_coeffs.view(-1, 4, nr_coeffs)
            float acc = __fmul_rn(__ldg(texel[0] + c), w[0]);
            acc = __fadd_rn(acc, __fmul_rn(__ldg(texel[1] + c), w[1]));
            acc = __fadd_rn(acc, __fmul_rn(__ldg(texel[2] + c), w[2]));
            acc = __fadd_rn(acc, __fmul_rn(__ldg(texel[3] + c), w[3]));
            co[k] = round_half(acc);  // .half()
        }
        float res = round_half(__fmul_rn(co[0], P[0]));  // degree 0: an fp16 product; every later term promotes to fp32
        if (deg > 0) {
            res = __fsub_rn(res, __fmul_rn(P[1], co[1]));
            res = __fadd_rn(res, __fmul_rn(P[2], co[2]));
            res = __fsub_rn(res, __fmul_rn(P[3], co[3]));
        }
        for (int k = 4; k < nr_coeffs; ++k) res = __fadd_rn(res, __fmul_rn(P[k], co[k]));
        const float o = sigmoid_precise(res);
        rgba[ch] = deg == 0 ? round_half(o) : o;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) o_rgb[3 * r + k] = fminf(fmaxf(rgba[k], 0.f), 1.f);
    o_alpha[r] = fminf(fmaxf(rgba[3], 0.f), 1.f);
}

}  // namespace vs

extern "C" {

// Shaded buffers of MeshRenderer.render_rays (volsurfs_py/renderers/mesh_renderer.py:112-201) from one mesh trace: is_hit [N] u8,
// triangles_id [N] i64, barycentric / view dirs / unit normals [N,3] f32 as RayTracer.trace returns them; face_uvs [F,3,2];
// tex: the baked SH-coefficient texture ZERO-PADDED by one texel on every side, [(res_h+2), (res_w+2), 4*nr_coeffs] f32 (what
// TensorTexture(lerp=True) keeps, tensor_texture.py:55-64); nr_coeffs in {1, 4, 9, 16}.  Outputs: is_hit [N,1], normals [N,3],
// uvs [N,3], rgb [N,3], alpha [N,1], view_dirs [N,3] f32 (the "ray_traced" dict of the reference).  All pointers DEVICE.
int vs_baked_texture_shade(const uint8_t* is_hit, const int64_t* tri_id, const float* bary, const float* dirs, const float* normals,
                           const float* face_uvs, const float* tex, int res_h, int res_w, int nr_coeffs, const float* bg_rgb_host,
                           float* o_hit, float* o_normals, float* o_uvs, float* o_rgb, float* o_alpha, float* o_dirs, int64_t n_rays,
                           void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && res_h > 0 && res_w > 0 && bg_rgb_host);
    VS_CHECK_ARG(nr_coeffs == 1 || nr_coeffs == 4 || nr_coeffs == 9 || nr_coeffs == 16);
    if (n_rays == 0) return VS_OK;
    VS_CHECK_ARG(is_hit && tri_id && bary && dirs && normals && face_uvs && tex && o_hit && o_normals && o_uvs && o_rgb && o_alpha && o_dirs);
    vs::baked_shade_kernel<<<(unsigned)vs::div_up(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(
        is_hit, tri_id, bary, dirs, normals, face_uvs, tex, res_h, res_w, nr_coeffs, bg_rgb_host[0], bg_rgb_host[1], bg_rgb_host[2], o_hit,
        o_normals, o_uvs, o_rgb, o_alpha, o_dirs, n_rays);
    return vs::launched(1);
}

}  // extern "C"
