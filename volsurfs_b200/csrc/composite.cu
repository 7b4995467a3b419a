// volsurfs_b200 — fused alpha compositing over packed samples, forward and backward.
//
// One launch per direction replaces the reference's chain
//   cumprod -> (alpha*T) -> sum_over_rays -> integrate_3d -> integrate_1d      (fwd, nerf.py:308-334)
//   + the 5 backward kernels + torch glue                                       (bwd, volume_rendering_funcs.py:105-272)
// and equals the dense K-layer torch path of the volsurfs method (volsurfs.py:601-640,708) when each ray's
// samples are its layer hits in outer->inner order.
//
//   T_i = prod_{j<i}(1-a_j)   w_i = T_i a_i
//   rgb = sum w c   depth = sum w z   acc = sum w   bgT = prod_all(1-a_j)       (full product, dense-path semantics)
//
// backward (nothing saved by the forward; T is recomputed, division free):
//   g_i = g_rgb.c_i + g_depth z_i + g_acc
//   R_{s-1} = g_bgT ,  R_{i-1} = a_i g_i + (1-a_i) R_i          (reverse scan of affine maps)
//   d_alpha_i = T_i (g_i - R_i) ,  d_rgb_i = g_rgb w_i
//
// Algorithmic HBM bytes per ray with s samples: fwd 8+20s read, 24 written; bwd 8+24+20s read, 16s written:
// B(s) = 64 + 56 s   (SURVEY.md section 8d).
//
// Two kernel families, chosen on the host from the mean segment length S/N:
//   * "tile" (mean <= 8, the K-layer shells case): a CTA owns 256 consecutive rays, stages their contiguous sample
//     range into shared memory with 16-byte coalesced loads, each thread then walks its own ray out of shared
//     memory (segments of ~5 samples are too short for lanes to share), per-sample gradients go back through shared
//     memory as coalesced 16-byte stores.
//   * "scan" (longer rays, NeRF-style packets): a group of W lanes per ray (W = 16/32), W-wide chunks read straight
//     from global memory (contiguous per chunk), shuffle-based exclusive cumprod forward, shuffle-based reverse affine
//     scan backward, running values carried between chunks.
#include <algorithm>
#include <cstdlib>

#include "vs_common.cuh"

namespace vs {

constexpr int kScanThreads = 256;
constexpr int kTileRays = 256;  // rays per CTA == threads per CTA in the tile kernels

// =============================================================================================
// scan family
// =============================================================================================
// One chunk of a ray as the scan kernels hold it in registers: alpha, z of sample base+gl and three CONTIGUOUS floats of the
// chunk's colour window (element gl + W*k of the 3W floats), so that colour loads/stores are unit-stride across the group.
struct ScanChunk {
    float a, z, c0, c1, c2;
};

template <int W>
__device__ __forceinline__ ScanChunk scan_load_chunk(const float* __restrict__ alpha, const float* __restrict__ rgb,
                                                     const float* __restrict__ z, int64_t start, int base, int n, int gl) {
    ScanChunk ch;
    const int i = base + gl;
    const int64_t s = start + i;
    const bool valid = i < n;
    ch.a = valid ? ld_stream(alpha + s) : 0.f;
    ch.z = valid ? ld_stream(z + s) : 0.f;
    const int64_t w0 = 3 * (start + base);  // first float of the chunk's colour window
    const int lim = 3 * (n - base);         // floats of the window that belong to the ray
    ch.c0 = gl < lim ? ld_stream(rgb + w0 + gl) : 0.f;
    ch.c1 = gl + W < lim ? ld_stream(rgb + w0 + gl + W) : 0.f;
    ch.c2 = gl + 2 * W < lim ? ld_stream(rgb + w0 + gl + 2 * W) : 0.f;
    return ch;
}

// window layout (element gl + W*k in register k) -> per-sample layout (sample gl owns elements 3gl..3gl+2), through the warp's
// shared-memory scratch; `my` points at this group's 3W floats
template <int W>
__device__ __forceinline__ void window_to_samples(float* my, int gl, float w0, float w1, float w2, float& r, float& g, float& b) {
    my[gl] = w0;
    my[gl + W] = w1;
    my[gl + 2 * W] = w2;
    __syncwarp();
    r = my[3 * gl];
    g = my[3 * gl + 1];
    b = my[3 * gl + 2];
    __syncwarp();
}
template <int W>
__device__ __forceinline__ void samples_to_window(float* my, int gl, float r, float g, float b, float& w0, float& w1, float& w2) {
    my[3 * gl] = r;
    my[3 * gl + 1] = g;
    my[3 * gl + 2] = b;
    __syncwarp();
    w0 = my[gl];
    w1 = my[gl + W];
    w2 = my[gl + 2 * W];
    __syncwarp();
}

template <int W>
__global__ void __launch_bounds__(kScanThreads) composite_fwd_scan_kernel(
    const int32_t* __restrict__ se, const float* __restrict__ alpha, const float* __restrict__ rgb, const float* __restrict__ z,
    float* __restrict__ out_rgb, float* __restrict__ out_depth, float* __restrict__ out_acc, float* __restrict__ out_bgT,
    float* __restrict__ out_w, float* __restrict__ out_T, int64_t n_rays) {
    __shared__ float s_win[kScanThreads / 32][96];
    const int lane = threadIdx.x & 31;
    const int gl = lane & (W - 1);
    float* my = s_win[threadIdx.x >> 5] + (lane / W) * 3 * W;
    const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / W;
    int start = 0, n = 0;
    if (ray < n_rays) n = load_segment(se, ray, start);
    const int n_max = warp_max_i32(n);

    float carry = 1.f;
    float ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aa = 0.f;
    ScanChunk cur = scan_load_chunk<W>(alpha, rgb, z, start, 0, n, gl);
    for (int base = 0; base < n_max; base += W) {
        ScanChunk nxt;
        if (base + W < n_max) nxt = scan_load_chunk<W>(alpha, rgb, z, start, base + W, n, gl);  // in flight during the scan below
        const int i = base + gl;
        const bool valid = i < n;
        const int64_t s = (int64_t)start + i;
        float cr, cg, cb;
        window_to_samples<W>(my, gl, cur.c0, cur.c1, cur.c2, cr, cg, cb);
        const float a = cur.a;
        float incl = group_scan_mul<W>(1.f - a, gl);
        float Ti = carry * group_shift_up<W>(incl, gl, 1.f);
        float w = Ti * a;
        ar = fmaf(w, cr, ar);
        ag = fmaf(w, cg, ag);
        ab = fmaf(w, cb, ab);
        ad = fmaf(w, cur.z, ad);
        aa += w;
        if (valid) {
            if (out_w) st_stream(out_w + s, w);
            if (out_T) st_stream(out_T + s, Ti);
        }
        carry *= group_bcast<W>(incl, W - 1);
        cur = nxt;
    }
    ar = group_reduce_add<W>(ar);
    ag = group_reduce_add<W>(ag);
    ab = group_reduce_add<W>(ab);
    ad = group_reduce_add<W>(ad);
    aa = group_reduce_add<W>(aa);
    if (ray < n_rays && gl == 0) {
        out_rgb[3 * ray] = ar;
        out_rgb[3 * ray + 1] = ag;
        out_rgb[3 * ray + 2] = ab;
        out_depth[ray] = ad;
        out_acc[ray] = aa;
        out_bgT[ray] = carry;
    }
}

template <int W>
__global__ void __launch_bounds__(kScanThreads) composite_bwd_scan_kernel(
    const int32_t* __restrict__ se, const float* __restrict__ alpha, const float* __restrict__ rgb, const float* __restrict__ z,
    const float* __restrict__ g_rgb, const float* __restrict__ g_depth, const float* __restrict__ g_acc, const float* __restrict__ g_bgT,
    float* __restrict__ d_alpha, float* __restrict__ d_rgb, float* __restrict__ d_z, int64_t n_rays) {
    __shared__ float s_win[kScanThreads / 32][96];
    const int lane = threadIdx.x & 31;
    const int gl = lane & (W - 1);
    float* my = s_win[threadIdx.x >> 5] + (lane / W) * 3 * W;
    const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / W;
    int start = 0, n = 0;
    if (ray < n_rays) n = load_segment(se, ray, start);
    const int n_max = warp_max_i32(n);
    if (n_max == 0) return;

    float gr = 0.f, gg = 0.f, gb = 0.f, gd = 0.f, ga = 0.f, gT = 0.f;
    if (n > 0) {
        gr = __ldg(g_rgb + 3 * ray);
        gg = __ldg(g_rgb + 3 * ray + 1);
        gb = __ldg(g_rgb + 3 * ray + 2);
        gd = __ldg(g_depth + ray);
        ga = __ldg(g_acc + ray);
        gT = __ldg(g_bgT + ray);
    }

    // pass 1 (left to right): transmittance at the start of every chunk.  Lane c keeps chunk c's value, which covers
    // W*W samples; beyond that the per-sample T is parked in d_alpha (scratch, overwritten in pass 2).
    const int n_chunks_max = (n_max + W - 1) / W;
    const bool spill = n_chunks_max > W;
    float my_chunk_T = 1.f;
    {
        float carry = 1.f;
        float a_cur = gl < n ? __ldg(alpha + start + gl) : 0.f;
        for (int c = 0; c < n_chunks_max; ++c) {
            const int i = c * W + gl;
            const bool valid = i < n;
            float a_nxt = 0.f;
            if (c + 1 < n_chunks_max && i + W < n) a_nxt = __ldg(alpha + start + i + W);
            float incl = group_scan_mul<W>(1.f - a_cur, gl);
            if (spill) {
                float Ti = carry * group_shift_up<W>(incl, gl, 1.f);
                if (valid) d_alpha[(int64_t)start + i] = Ti;
            } else if (gl == c) {
                my_chunk_T = carry;
            }
            carry *= group_bcast<W>(incl, W - 1);
            a_cur = a_nxt;
        }
    }

    // pass 2 (right to left): reverse affine scan
    float Rcarry = gT;
    ScanChunk cur = scan_load_chunk<W>(alpha, rgb, z, start, (n_chunks_max - 1) * W, n, gl);
    for (int c = n_chunks_max - 1; c >= 0; --c) {
        ScanChunk nxt;
        if (c > 0) nxt = scan_load_chunk<W>(alpha, rgb, z, start, (c - 1) * W, n, gl);
        const int i = c * W + gl;
        const bool valid = i < n;
        const int64_t s = (int64_t)start + i;
        float cr, cg, cb;
        window_to_samples<W>(my, gl, cur.c0, cur.c1, cur.c2, cr, cg, cb);
        const float a = cur.a, zz = cur.z;
        float Ti;
        if (spill) {
            Ti = valid ? d_alpha[s] : 0.f;
        } else {
            float incl = group_scan_mul<W>(1.f - a, gl);
            float chunkT = group_bcast<W>(my_chunk_T, c & (W - 1));
            Ti = chunkT * group_shift_up<W>(incl, gl, 1.f);
        }
        const float gi = fmaf(gr, cr, fmaf(gg, cg, fmaf(gb, cb, fmaf(gd, zz, ga))));
        // F_i(x) = (1-a) x + a g_i ; identity for lanes past the end
        float A = valid ? (1.f - a) : 1.f;
        float B = valid ? a * gi : 0.f;
        group_rscan_affine<W>(A, B, gl);
        // R_{i-1} = A*Rcarry + B ; R_i is the next lane's value (Rcarry for the last lane)
        float Rprev = fmaf(A, Rcarry, B);
        float Ri = __shfl_down_sync(VS_FULL_MASK, Rprev, 1, W);
        if (gl == W - 1) Ri = Rcarry;
        const float w = Ti * a;
        float o0, o1, o2;
        samples_to_window<W>(my, gl, gr * w, gg * w, gb * w, o0, o1, o2);
        {
            const int64_t w0 = 3 * ((int64_t)start + c * W);
            const int lim = 3 * (n - c * W);
            if (gl < lim) st_stream(d_rgb + w0 + gl, o0);
            if (gl + W < lim) st_stream(d_rgb + w0 + gl + W, o1);
            if (gl + 2 * W < lim) st_stream(d_rgb + w0 + gl + 2 * W, o2);
        }
        if (valid) {
            st_stream(d_alpha + s, Ti * (gi - Ri));
            if (d_z) st_stream(d_z + s, gd * w);
        }
        Rcarry = group_bcast<W>(Rprev, 0);
        cur = nxt;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// scan family, coarsened: every lane owns an ALIGNED QUAD of samples per chunk (absolute sample indices 4q..4q+3), so alpha / z
// travel as one 16-byte load and the quad's colours as three; the exclusive cumprod is a 4-element serial product per lane plus ONE
// shuffle scan over the lanes' totals per 4W samples, the backward recurrence a serial composition of the lane's four affine maps
// plus ONE reverse shuffle scan.  Chunks are anchored at the ray's start rounded down to a multiple of four; samples outside
// [start, end) are masked to the identity (alpha = 0).
// ---------------------------------------------------------------------------------------------------------------------------
struct Quad {
    float a[4], z[4], c[12];
};

__device__ __forceinline__ void load_quad(Quad& q, const float* __restrict__ alpha, const float* __restrict__ rgb,
                                          const float* __restrict__ z, int64_t q0, int64_t start, int64_t end, int64_t n_samples,
                                          bool active) {
#pragma unroll
    for (int j = 0; j < 4; ++j) q.a[j] = 0.f, q.z[j] = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) q.c[j] = 0.f;
    if (!active) return;
    if (q0 + 4 <= n_samples) {  // the whole quad lies inside the arrays: 16-byte loads (bases are 16-byte aligned, q0 % 4 == 0)
        const float4 a4 = ld_stream4(reinterpret_cast<const float4*>(alpha + q0));
        const float4 z4 = ld_stream4(reinterpret_cast<const float4*>(z + q0));
        const float4 c0 = ld_stream4(reinterpret_cast<const float4*>(rgb + 3 * q0));
        const float4 c1 = ld_stream4(reinterpret_cast<const float4*>(rgb + 3 * q0 + 4));
        const float4 c2 = ld_stream4(reinterpret_cast<const float4*>(rgb + 3 * q0 + 8));
        q.a[0] = a4.x, q.a[1] = a4.y, q.a[2] = a4.z, q.a[3] = a4.w;
        q.z[0] = z4.x, q.z[1] = z4.y, q.z[2] = z4.z, q.z[3] = z4.w;
        q.c[0] = c0.x, q.c[1] = c0.y, q.c[2] = c0.z, q.c[3] = c0.w, q.c[4] = c1.x, q.c[5] = c1.y;
        q.c[6] = c1.z, q.c[7] = c1.w, q.c[8] = c2.x, q.c[9] = c2.y, q.c[10] = c2.z, q.c[11] = c2.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (q0 + j < n_samples) {
                q.a[j] = alpha[q0 + j];
                q.z[j] = z[q0 + j];
                q.c[3 * j] = rgb[3 * (q0 + j)];
                q.c[3 * j + 1] = rgb[3 * (q0 + j) + 1];
                q.c[3 * j + 2] = rgb[3 * (q0 + j) + 2];
            }
        }
    }
    // samples of the quad that belong to other rays (before start / at or after end) become identities
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (q0 + j < start || q0 + j >= end) q.a[j] = 0.f;
}

template <int W>
__global__ void __launch_bounds__(kScanThreads) composite_fwd_scan4_kernel(
    const int32_t* __restrict__ se, const float* __restrict__ alpha, const float* __restrict__ rgb, const float* __restrict__ z,
    float* __restrict__ out_rgb, float* __restrict__ out_depth, float* __restrict__ out_acc, float* __restrict__ out_bgT,
    float* __restrict__ out_w, float* __restrict__ out_T, int64_t n_rays, int64_t n_samples) {
    const int gl = threadIdx.x & (W - 1);
    const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / W;
    int start = 0, n = 0;
    if (ray < n_rays) n = load_segment(se, ray, start);
    const int64_t end = (int64_t)start + n;
    const int64_t A = (int64_t)start & ~(int64_t)3;
    const int n_chunks = n > 0 ? (int)((end - A + 4 * W - 1) / (4 * W)) : 0;
    const int n_chunks_max = warp_max_i32(n_chunks);

    float carry = 1.f;
    float ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aa = 0.f;
    Quad cur;
    load_quad(cur, alpha, rgb, z, A + 4 * gl, start, end, n_samples, n > 0 && A + 4 * gl < end);
    for (int c = 0; c < n_chunks_max; ++c) {
        const int64_t q0 = A + (int64_t)c * 4 * W + 4 * gl;
        Quad nxt;
        if (c + 1 < n_chunks_max) load_quad(nxt, alpha, rgb, z, q0 + 4 * W, start, end, n_samples, n > 0 && q0 + 4 * W < end);
        float tl[4];
        float p = 1.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            tl[j] = p;
            p *= (1.f - cur.a[j]);
        }
        const float incl = group_scan_mul<W>(p, gl);
        const float baseT = carry * group_shift_up<W>(incl, gl, 1.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float Tj = baseT * tl[j];
            const float w = Tj * cur.a[j];
            ar = fmaf(w, cur.c[3 * j], ar);
            ag = fmaf(w, cur.c[3 * j + 1], ag);
            ab = fmaf(w, cur.c[3 * j + 2], ab);
            ad = fmaf(w, cur.z[j], ad);
            aa += w;
            if ((out_w || out_T) && q0 + j >= start && q0 + j < end) {
                if (out_w) out_w[q0 + j] = w;
                if (out_T) out_T[q0 + j] = Tj;
            }
        }
        carry *= group_bcast<W>(incl, W - 1);
        cur = nxt;
    }
    ar = group_reduce_add<W>(ar);
    ag = group_reduce_add<W>(ag);
    ab = group_reduce_add<W>(ab);
    ad = group_reduce_add<W>(ad);
    aa = group_reduce_add<W>(aa);
    if (ray < n_rays && gl == 0) {
        out_rgb[3 * ray] = ar;
        out_rgb[3 * ray + 1] = ag;
        out_rgb[3 * ray + 2] = ab;
        out_depth[ray] = ad;
        out_acc[ray] = aa;
        out_bgT[ray] = carry;
    }
}

template <int W>
__global__ void __launch_bounds__(kScanThreads) composite_bwd_scan4_kernel(
    const int32_t* __restrict__ se, const float* __restrict__ alpha, const float* __restrict__ rgb, const float* __restrict__ z,
    const float* __restrict__ g_rgb, const float* __restrict__ g_depth, const float* __restrict__ g_acc, const float* __restrict__ g_bgT,
    float* __restrict__ d_alpha, float* __restrict__ d_rgb, float* __restrict__ d_z, int64_t n_rays, int64_t n_samples) {
    const int gl = threadIdx.x & (W - 1);
    const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / W;
    int start = 0, n = 0;
    if (ray < n_rays) n = load_segment(se, ray, start);
    const int64_t end = (int64_t)start + n;
    const int64_t A = (int64_t)start & ~(int64_t)3;
    const int n_chunks = n > 0 ? (int)((end - A + 4 * W - 1) / (4 * W)) : 0;
    const int n_chunks_max = warp_max_i32(n_chunks);
    if (n_chunks_max == 0) return;

    float gr = 0.f, gg = 0.f, gb = 0.f, gd = 0.f, ga = 0.f, gT = 0.f;
    if (n > 0) {
        gr = __ldg(g_rgb + 3 * ray);
        gg = __ldg(g_rgb + 3 * ray + 1);
        gb = __ldg(g_rgb + 3 * ray + 2);
        gd = __ldg(g_depth + ray);
        ga = __ldg(g_acc + ray);
        gT = __ldg(g_bgT + ray);
    }

    // pass 1 (left to right): transmittance at the start of every chunk; lane c keeps chunk c's value (covers 4*W*W samples),
    // beyond that the per-sample T is parked in d_alpha and overwritten in pass 2
    const bool spill = n_chunks_max > W;
    float my_chunk_T = 1.f;
    {
        float carry = 1.f;
        for (int c = 0; c < n_chunks_max; ++c) {
            const int64_t q0 = A + (int64_t)c * 4 * W + 4 * gl;
            float a[4] = {0.f, 0.f, 0.f, 0.f};
            if (n > 0 && q0 < end) {
                if (q0 + 4 <= n_samples) {
                    const float4 a4 = __ldg(reinterpret_cast<const float4*>(alpha + q0));
                    a[0] = a4.x, a[1] = a4.y, a[2] = a4.z, a[3] = a4.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (q0 + j < n_samples) a[j] = alpha[q0 + j];
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (q0 + j < start || q0 + j >= end) a[j] = 0.f;
            }
            float tl[4];
            float p = 1.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                tl[j] = p;
                p *= (1.f - a[j]);
            }
            const float incl = group_scan_mul<W>(p, gl);
            if (spill) {
                const float baseT = carry * group_shift_up<W>(incl, gl, 1.f);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (q0 + j >= start && q0 + j < end) d_alpha[q0 + j] = baseT * tl[j];
            } else if (gl == c) {
                my_chunk_T = carry;
            }
            carry *= group_bcast<W>(incl, W - 1);
        }
    }

    // pass 2 (right to left)
    float Rcarry = gT;
    Quad cur;
    {
        const int64_t q0 = A + (int64_t)(n_chunks_max - 1) * 4 * W + 4 * gl;
        load_quad(cur, alpha, rgb, z, q0, start, end, n_samples, n > 0 && q0 < end);
    }
    for (int c = n_chunks_max - 1; c >= 0; --c) {
        const int64_t q0 = A + (int64_t)c * 4 * W + 4 * gl;
        Quad nxt;
        if (c > 0) load_quad(nxt, alpha, rgb, z, q0 - 4 * W, start, end, n_samples, n > 0 && q0 - 4 * W < end);
        bool valid[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) valid[j] = q0 + j >= start && q0 + j < end;
        float T[4];
        if (spill) {
#pragma unroll
            for (int j = 0; j < 4; ++j) T[j] = valid[j] ? d_alpha[q0 + j] : 0.f;
        } else {
            float p = 1.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                T[j] = p;
                p *= (1.f - cur.a[j]);
            }
            const float incl = group_scan_mul<W>(p, gl);
            const float baseT = group_bcast<W>(my_chunk_T, c & (W - 1)) * group_shift_up<W>(incl, gl, 1.f);
#pragma unroll
            for (int j = 0; j < 4; ++j) T[j] *= baseT;
        }
        float g[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            g[j] = fmaf(gr, cur.c[3 * j], fmaf(gg, cur.c[3 * j + 1], fmaf(gb, cur.c[3 * j + 2], fmaf(gd, cur.z[j], ga))));
        // the lane's map M = F_0 o F_1 o F_2 o F_3 with F_j(x) = (1-a_j) x + a_j g_j (identity where alpha was masked to 0)
        float MA = 1.f, MB = 0.f;
#pragma unroll
        for (int j = 3; j >= 0; --j) {
            const float Aj = 1.f - cur.a[j], Bj = cur.a[j] * g[j];
            MB = fmaf(Aj, MB, Bj);
            MA = Aj * MA;
        }
        group_rscan_affine<W>(MA, MB, gl);
        const float V = fmaf(MA, Rcarry, MB);  // R just left of this lane's quad
        float R = __shfl_down_sync(VS_FULL_MASK, V, 1, W);
        if (gl == W - 1) R = Rcarry;           // R just right of this lane's quad
        float da[4], dc[12], dz[4];
#pragma unroll
        for (int j = 3; j >= 0; --j) {
            const float w = T[j] * cur.a[j];
            da[j] = T[j] * (g[j] - R);
            dc[3 * j] = gr * w;
            dc[3 * j + 1] = gg * w;
            dc[3 * j + 2] = gb * w;
            dz[j] = gd * w;
            R = fmaf(1.f - cur.a[j], R, cur.a[j] * g[j]);
        }
        if (valid[0] && valid[3] && q0 + 4 <= n_samples) {  // whole quad owned by this ray: 16-byte stores
            st_stream4(reinterpret_cast<float4*>(d_alpha + q0), make_float4(da[0], da[1], da[2], da[3]));
            st_stream4(reinterpret_cast<float4*>(d_rgb + 3 * q0), make_float4(dc[0], dc[1], dc[2], dc[3]));
            st_stream4(reinterpret_cast<float4*>(d_rgb + 3 * q0 + 4), make_float4(dc[4], dc[5], dc[6], dc[7]));
            st_stream4(reinterpret_cast<float4*>(d_rgb + 3 * q0 + 8), make_float4(dc[8], dc[9], dc[10], dc[11]));
            if (d_z) st_stream4(reinterpret_cast<float4*>(d_z + q0), make_float4(dz[0], dz[1], dz[2], dz[3]));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (valid[j]) {
                    d_alpha[q0 + j] = da[j];
                    d_rgb[3 * (q0 + j)] = dc[3 * j];
                    d_rgb[3 * (q0 + j) + 1] = dc[3 * j + 1];
                    d_rgb[3 * (q0 + j) + 2] = dc[3 * j + 2];
                    if (d_z) d_z[q0 + j] = dz[j];
                }
            }
        }
        Rcarry = group_bcast<W>(V, 0);
        cur = nxt;
    }
}

// =============================================================================================
// ring family (long, variable-length rays: NeRF-style packets)
// =============================================================================================
// What bounds the scan kernels above on long rays is bytes in flight: a warp waits a full memory latency for every chunk it has
// prefetched one iteration earlier, and most rays are one or two chunks long, so nothing is in flight while a ray starts or ends.
// Here a warp owns 32 consecutive rays and streams ALL their chunks (128 samples: an aligned quad per lane, as in the coarsened scan
// kernels) through its own shared-memory ring with cp.async — kRingDepth chunks of 2.5 KB in flight per warp, across ray boundaries,
// at no register cost.  Every lane copies exactly the bytes it later reads, so the ring needs no barrier: cp.async.wait_group is
// per thread.  The arithmetic is that of the coarsened scan kernels (serial 4-sample products + one shuffle scan per chunk).
constexpr int kRingWarps = 8;
constexpr int kRingDepthDefault = 3;
constexpr int kRingStageFloats = 640;  // 128 alpha | 128 z | 384 rgb

__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_full(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// group-masked shuffle primitives: the W-lane groups of a warp walk different rays and diverge from each other, so every shuffle
// names only its own group's lanes
template <int W>
__device__ __forceinline__ unsigned ring_group_mask(int lane) {
    return W == 32 ? 0xffffffffu : (((1u << W) - 1u) << (lane & ~(W - 1)));
}
template <int W>
__device__ __forceinline__ float ring_scan_mul(unsigned m, float v, int gl) {
#pragma unroll
    for (int d = 1; d < W; d <<= 1) {
        const float o = __shfl_up_sync(m, v, d, W);
        if (gl >= d) v *= o;
    }
    return v;
}
template <int W>
__device__ __forceinline__ float ring_reduce_add(unsigned m, float v) {
#pragma unroll
    for (int d = W / 2; d > 0; d >>= 1) v += __shfl_xor_sync(m, v, d, W);
    return v;
}
template <int W>
__device__ __forceinline__ void ring_rscan_affine(unsigned m, float& A, float& B, int gl) {
#pragma unroll
    for (int d = 1; d < W; d <<= 1) {
        const float Ao = __shfl_down_sync(m, A, d, W);
        const float Bo = __shfl_down_sync(m, B, d, W);
        if (gl + d < W) {
            B = fmaf(A, Bo, B);
            A = A * Ao;
        }
    }
}

// A warp's batch is 32 consecutive rays, one per lane; group g (W lanes) walks rays g*W .. g*W+W-1.  Chunks are 4W samples (an aligned
// quad per lane), anchored at the ray's start rounded down to a multiple of four samples (sample indices fit 32 bits: int32 segments).
struct RingRays {
    int start, end, nch;
};
template <int W>
__device__ __forceinline__ RingRays ring_load_rays(const int32_t* __restrict__ se, int64_t ray, int64_t n_rays) {
    RingRays r{0, 0, 0};
    if (ray < n_rays) {
        int st = 0;
        const int n = load_segment(se, ray, st);
        if (n > 0) {
            r.start = st;
            r.end = st + n;
            r.nch = (r.end - (st & ~3) + 4 * W - 1) / (4 * W);
        }
    }
    return r;
}

// group-uniform cursor over the (ray, item) pairs of a group: the current ray's segment is cached and re-fetched (three shuffles) only
// when the cursor moves to another ray.  r == W: the group is done (start = end = 0, so every quad reads as empty).
struct RingCursor2 {
    int r, k, items, start, end, nch;
};
template <bool BWD, int W>
__device__ __forceinline__ void ring_cursor_load(RingCursor2& c, const RingRays& mine, unsigned m, int lane) {
    // first ray at or after c.r that has items (BWD: one reverse item for single-chunk rays, else nch alpha-only + nch reverse items)
    const unsigned has = (__ballot_sync(m, mine.nch > 0) >> (lane & ~(W - 1))) & (W == 32 ? 0xffffffffu : ((1u << W) - 1u));
    const unsigned rest = c.r < W ? (has >> c.r) : 0u;
    if (rest == 0u) {
        c.r = W;
        c.k = c.items = c.start = c.end = c.nch = 0;
        return;
    }
    c.r += __ffs(rest) - 1;
    c.k = 0;
    c.start = __shfl_sync(m, mine.start, c.r, W);
    c.end = __shfl_sync(m, mine.end, c.r, W);
    c.nch = __shfl_sync(m, mine.nch, c.r, W);
    c.items = BWD ? (c.nch == 1 ? 1 : 2 * c.nch) : c.nch;
}
template <bool BWD, int W>
__device__ __forceinline__ void ring_cursor_next(RingCursor2& c, const RingRays& mine, unsigned m, int lane) {
    if (++c.k >= c.items) {
        ++c.r;
        ring_cursor_load<BWD, W>(c, mine, m, lane);
    }
}

// this lane's quad of chunk c of the ray [start, end): 16-byte async copies (zero fill past the end of the arrays); ALPHA_ONLY for
// the backward's transmittance pass.  `stage` is the GROUP's stage: alpha [4W] | z [4W] | rgb [12W]
template <bool ALPHA_ONLY, int W>
__device__ __forceinline__ void ring_issue(float* stage, const float* __restrict__ alpha, const float* __restrict__ rgb,
                                           const float* __restrict__ z, int start, int end, int c, int gl, int n_samples) {
    const int q0 = (start & ~3) + c * 4 * W + 4 * gl;
    if (q0 >= end) return;  // nothing of this ray in the lane's quad: the consumer masks it without reading
    float* sa = stage + 4 * gl;
    float* sz = stage + 4 * W + 4 * gl;
    float* sc = stage + 8 * W + 12 * gl;
    const float* c3 = rgb + 3 * (size_t)q0;
    if (q0 + 4 <= n_samples) {
        cp_async16_full(sa, alpha + q0);
        if (!ALPHA_ONLY) {
            cp_async16_full(sz, z + q0);
            cp_async16_full(sc, c3);
            cp_async16_full(sc + 4, c3 + 4);
            cp_async16_full(sc + 8, c3 + 8);
        }
        return;
    }
    // the last quad of the arrays (1..3 samples inside them): element-wise copies, the rest of the quad zeroed — no byte past the end
    // of an array is ever addressed
    const int valid = n_samples - q0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j < valid) {
            cp_async4(sa + j, alpha + q0 + j);
            if (!ALPHA_ONLY) {
                cp_async4(sz + j, z + q0 + j);
                cp_async4(sc + 3 * j, c3 + 3 * j);
                cp_async4(sc + 3 * j + 1, c3 + 3 * j + 1);
                cp_async4(sc + 3 * j + 2, c3 + 3 * j + 2);
            }
        } else {
            sa[j] = 0.f;
            if (!ALPHA_ONLY) {
                sz[j] = 0.f;
                sc[3 * j] = sc[3 * j + 1] = sc[3 * j + 2] = 0.f;
            }
        }
    }
}

// quad out of the ring; samples outside [start, end) become identities (alpha = 0).  Their colour / depth are the neighbouring rays'
// (finite) values and meet a zero weight.
template <bool ALPHA_ONLY, int W>
__device__ __forceinline__ void ring_read(Quad& q, const float* stage, int q0, int start, int end, int gl) {
    float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f), z4 = a4, c0 = a4, c1 = a4, c2 = a4;
    if (q0 < end) {
        a4 = *reinterpret_cast<const float4*>(stage + 4 * gl);
        if (!ALPHA_ONLY) {
            z4 = *reinterpret_cast<const float4*>(stage + 4 * W + 4 * gl);
            c0 = *reinterpret_cast<const float4*>(stage + 8 * W + 12 * gl);
            c1 = *reinterpret_cast<const float4*>(stage + 8 * W + 12 * gl + 4);
            c2 = *reinterpret_cast<const float4*>(stage + 8 * W + 12 * gl + 8);
        }
    }
    const int lo = start - q0, hi = end - q0;  // sample j of the quad belongs to the ray iff lo <= j < hi
    q.a[0] = (0 >= lo && 0 < hi) ? a4.x : 0.f;
    q.a[1] = (1 >= lo && 1 < hi) ? a4.y : 0.f;
    q.a[2] = (2 >= lo && 2 < hi) ? a4.z : 0.f;
    q.a[3] = (3 >= lo && 3 < hi) ? a4.w : 0.f;
    if (!ALPHA_ONLY) {
        q.z[0] = z4.x, q.z[1] = z4.y, q.z[2] = z4.z, q.z[3] = z4.w;
        q.c[0] = c0.x, q.c[1] = c0.y, q.c[2] = c0.z, q.c[3] = c0.w, q.c[4] = c1.x, q.c[5] = c1.y;
        q.c[6] = c1.z, q.c[7] = c1.w, q.c[8] = c2.x, q.c[9] = c2.y, q.c[10] = c2.z, q.c[11] = c2.w;
    }
}

// WT: also store per-sample weights / transmittance (out_w, out_T may each be NULL)
template <bool WT, int D, int W>
__global__ void __launch_bounds__(32 * kRingWarps) composite_fwd_ring_kernel(
    const int32_t* __restrict__ se, const float* __restrict__ alpha, const float* __restrict__ rgb, const float* __restrict__ z,
    float* __restrict__ out_rgb, float* __restrict__ out_depth, float* __restrict__ out_acc, float* __restrict__ out_bgT,
    float* __restrict__ out_w, float* __restrict__ out_T, int64_t n_rays, int n_samples) {
    extern __shared__ __align__(16) float ring_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane & (W - 1), gbase = lane & ~(W - 1);
    const unsigned m = ring_group_mask<W>(lane);
    float* ring = ring_smem + (size_t)warp * D * kRingStageFloats + (gbase / W) * 20 * W;  // this group's slice of every stage
    const int64_t n_batches = (n_rays + 31) / 32;
    for (int64_t b = (int64_t)blockIdx.x * kRingWarps + warp; b < n_batches; b += (int64_t)gridDim.x * kRingWarps) {
        const int64_t ray0 = b * 32;
        const RingRays mine = ring_load_rays<W>(se, ray0 + lane, n_rays);
        if (ray0 + lane < n_rays && mine.nch == 0) {  // empty ray: nothing composited, full transmittance
            const int64_t r = ray0 + lane;
            out_rgb[3 * r] = out_rgb[3 * r + 1] = out_rgb[3 * r + 2] = 0.f;
            out_depth[r] = 0.f;
            out_acc[r] = 0.f;
            out_bgT[r] = 1.f;
        }
        RingCursor2 pc{0, 0, 0, 0, 0, 0};
        ring_cursor_load<false, W>(pc, mine, m, lane);
        RingCursor2 cc = pc;
        int ps = 0, cs = 0;  // producer / consumer stage
        auto issue_one = [&]() {
            if (pc.r < W) {
                ring_issue<false, W>(ring + ps * kRingStageFloats, alpha, rgb, z, pc.start, pc.end, pc.k, gl, n_samples);
                ps = ps + 1 == D ? 0 : ps + 1;
                ring_cursor_next<false, W>(pc, mine, m, lane);
            }
            cp_async_commit();
        };
#pragma unroll
        for (int i = 0; i < D - 1; ++i) issue_one();

        float carry = 1.f, ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aa = 0.f;
        while (cc.r < W) {
            issue_one();
            cp_async_wait<D - 1>();
            const int q0 = (cc.start & ~3) + cc.k * 4 * W + 4 * gl;
            Quad cur;
            ring_read<false, W>(cur, ring + cs * kRingStageFloats, q0, cc.start, cc.end, gl);
            cs = cs + 1 == D ? 0 : cs + 1;
            float tl[4];
            float p = 1.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                tl[j] = p;
                p *= (1.f - cur.a[j]);
            }
            const float incl = ring_scan_mul<W>(m, p, gl);
            float excl = __shfl_up_sync(m, incl, 1, W);
            if (gl == 0) excl = 1.f;
            const float baseT = carry * excl;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float Tj = baseT * tl[j];
                const float w = Tj * cur.a[j];
                ar = fmaf(w, cur.c[3 * j], ar);
                ag = fmaf(w, cur.c[3 * j + 1], ag);
                ab = fmaf(w, cur.c[3 * j + 2], ab);
                ad = fmaf(w, cur.z[j], ad);
                aa += w;
                if (WT && q0 + j >= cc.start && q0 + j < cc.end) {
                    if (out_w) out_w[q0 + j] = w;
                    if (out_T) out_T[q0 + j] = Tj;
                }
            }
            carry *= __shfl_sync(m, incl, W - 1, W);
            if (cc.k == cc.nch - 1) {  // the ray is complete
                ar = ring_reduce_add<W>(m, ar);
                ag = ring_reduce_add<W>(m, ag);
                ab = ring_reduce_add<W>(m, ab);
                ad = ring_reduce_add<W>(m, ad);
                aa = ring_reduce_add<W>(m, aa);
                if (gl == 0) {
                    const int64_t r = ray0 + gbase + cc.r;
                    out_rgb[3 * r] = ar;
                    out_rgb[3 * r + 1] = ag;
                    out_rgb[3 * r + 2] = ab;
                    out_depth[r] = ad;
                    out_acc[r] = aa;
                    out_bgT[r] = carry;
                }
                carry = 1.f;
                ar = ag = ab = ad = aa = 0.f;
            }
            ring_cursor_next<false, W>(cc, mine, m, lane);
        }
        cp_async_wait<0>();
        __syncwarp();
    }
}

// Backward: a ray with one chunk is a single item (transmittance and reverse recurrence from the same quad); a longer ray is a
// left-to-right pass over alpha (transmittance at every chunk start, kept by lane c for chunk c) followed by the right-to-left pass.
template <bool DZ, int D, int W>
__global__ void __launch_bounds__(32 * kRingWarps) composite_bwd_ring_kernel(
    const int32_t* __restrict__ se, const float* __restrict__ alpha, const float* __restrict__ rgb, const float* __restrict__ z,
    const float* __restrict__ g_rgb, const float* __restrict__ g_depth, const float* __restrict__ g_acc, const float* __restrict__ g_bgT,
    float* __restrict__ d_alpha, float* __restrict__ d_rgb, float* __restrict__ d_z, int64_t n_rays, int n_samples) {
    extern __shared__ __align__(16) float ring_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane & (W - 1), gbase = lane & ~(W - 1);
    const unsigned m = ring_group_mask<W>(lane);
    float* ring = ring_smem + (size_t)warp * D * kRingStageFloats + (gbase / W) * 20 * W;
    const int64_t n_batches = (n_rays + 31) / 32;
    for (int64_t b = (int64_t)blockIdx.x * kRingWarps + warp; b < n_batches; b += (int64_t)gridDim.x * kRingWarps) {
        const int64_t ray0 = b * 32;
        const RingRays mine = ring_load_rays<W>(se, ray0 + lane, n_rays);
        float m_gr = 0.f, m_gg = 0.f, m_gb = 0.f, m_gd = 0.f, m_ga = 0.f, m_gT = 0.f;
        if (mine.nch > 0) {
            const int64_t r = ray0 + lane;
            m_gr = __ldg(g_rgb + 3 * r);
            m_gg = __ldg(g_rgb + 3 * r + 1);
            m_gb = __ldg(g_rgb + 3 * r + 2);
            m_gd = __ldg(g_depth + r);
            m_ga = __ldg(g_acc + r);
            m_gT = __ldg(g_bgT + r);
        }
        RingCursor2 pc{0, 0, 0, 0, 0, 0};
        ring_cursor_load<true, W>(pc, mine, m, lane);
        RingCursor2 cc = pc;
        int ps = 0, cs = 0;
        auto issue_one = [&]() {
            if (pc.r < W) {
                float* stage = ring + ps * kRingStageFloats;
                if (pc.nch > 1 && pc.k < pc.nch)
                    ring_issue<true, W>(stage, alpha, rgb, z, pc.start, pc.end, pc.k, gl, n_samples);
                else
                    ring_issue<false, W>(stage, alpha, rgb, z, pc.start, pc.end, pc.nch > 1 ? 2 * pc.nch - 1 - pc.k : 0, gl, n_samples);
                ps = ps + 1 == D ? 0 : ps + 1;
                ring_cursor_next<true, W>(pc, mine, m, lane);
            }
            cp_async_commit();
        };
#pragma unroll
        for (int i = 0; i < D - 1; ++i) issue_one();

        float carry = 1.f, my_chunk_T = 1.f, Rcarry = 0.f;
        float gr = 0.f, gg = 0.f, gb = 0.f, gd = 0.f, ga = 0.f;
        while (cc.r < W) {
            issue_one();
            cp_async_wait<D - 1>();
            const int s = cc.start, e = cc.end, nch = cc.nch;
            const bool spill = nch > W;  // more chunk starts than lanes: per-sample T parked in d_alpha (overwritten by the reverse pass)
            const float* stage = ring + cs * kRingStageFloats;
            cs = cs + 1 == D ? 0 : cs + 1;
            if (cc.k == 0) {  // a new ray: its upstream gradients, fresh carries
                gr = __shfl_sync(m, m_gr, cc.r, W);
                gg = __shfl_sync(m, m_gg, cc.r, W);
                gb = __shfl_sync(m, m_gb, cc.r, W);
                gd = __shfl_sync(m, m_gd, cc.r, W);
                ga = __shfl_sync(m, m_ga, cc.r, W);
                Rcarry = __shfl_sync(m, m_gT, cc.r, W);
                carry = 1.f;
                my_chunk_T = 1.f;
            }
            if (nch > 1 && cc.k < nch) {
                // ---- transmittance pass, chunk cc.k
                const int c = cc.k;
                const int q0 = (s & ~3) + c * 4 * W + 4 * gl;
                Quad cur;
                ring_read<true, W>(cur, stage, q0, s, e, gl);
                float tl[4];
                float p = 1.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    tl[j] = p;
                    p *= (1.f - cur.a[j]);
                }
                const float incl = ring_scan_mul<W>(m, p, gl);
                if (spill) {
                    float excl = __shfl_up_sync(m, incl, 1, W);
                    if (gl == 0) excl = 1.f;
                    const float baseT = carry * excl;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (q0 + j >= s && q0 + j < e) d_alpha[q0 + j] = baseT * tl[j];
                } else if (gl == c) {
                    my_chunk_T = carry;
                }
                carry *= __shfl_sync(m, incl, W - 1, W);
            } else {
                // ---- reverse pass, chunk c
                const int c = nch > 1 ? 2 * nch - 1 - cc.k : 0;
                const int q0 = (s & ~3) + c * 4 * W + 4 * gl;
                Quad cur;
                ring_read<false, W>(cur, stage, q0, s, e, gl);
                bool valid[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) valid[j] = q0 + j >= s && q0 + j < e;
                float T[4];
                if (spill) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) T[j] = valid[j] ? d_alpha[q0 + j] : 0.f;
                } else {
                    float p = 1.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        T[j] = p;
                        p *= (1.f - cur.a[j]);
                    }
                    const float incl = ring_scan_mul<W>(m, p, gl);
                    float excl = __shfl_up_sync(m, incl, 1, W);
                    if (gl == 0) excl = 1.f;
                    const float baseT = __shfl_sync(m, my_chunk_T, c & (W - 1), W) * excl;
#pragma unroll
                    for (int j = 0; j < 4; ++j) T[j] *= baseT;
                }
                float g[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    g[j] = fmaf(gr, cur.c[3 * j], fmaf(gg, cur.c[3 * j + 1], fmaf(gb, cur.c[3 * j + 2], fmaf(gd, cur.z[j], ga))));
                float MA = 1.f, MB = 0.f;
#pragma unroll
                for (int j = 3; j >= 0; --j) {
                    const float Aj = 1.f - cur.a[j], Bj = cur.a[j] * g[j];
                    MB = fmaf(Aj, MB, Bj);
                    MA = Aj * MA;
                }
                ring_rscan_affine<W>(m, MA, MB, gl);
                const float V = fmaf(MA, Rcarry, MB);
                float R = __shfl_down_sync(m, V, 1, W);
                if (gl == W - 1) R = Rcarry;
                float da[4], dc[12], dz[4];
#pragma unroll
                for (int j = 3; j >= 0; --j) {
                    const float w = T[j] * cur.a[j];
                    da[j] = T[j] * (g[j] - R);
                    dc[3 * j] = gr * w;
                    dc[3 * j + 1] = gg * w;
                    dc[3 * j + 2] = gb * w;
                    dz[j] = gd * w;
                    R = fmaf(1.f - cur.a[j], R, cur.a[j] * g[j]);
                }
                if (valid[0] && valid[3] && q0 + 4 <= n_samples) {
                    st_stream4(reinterpret_cast<float4*>(d_alpha + q0), make_float4(da[0], da[1], da[2], da[3]));
                    float* c3 = d_rgb + 3 * (size_t)q0;
                    st_stream4(reinterpret_cast<float4*>(c3), make_float4(dc[0], dc[1], dc[2], dc[3]));
                    st_stream4(reinterpret_cast<float4*>(c3 + 4), make_float4(dc[4], dc[5], dc[6], dc[7]));
                    st_stream4(reinterpret_cast<float4*>(c3 + 8), make_float4(dc[8], dc[9], dc[10], dc[11]));
                    if (DZ) st_stream4(reinterpret_cast<float4*>(d_z + q0), make_float4(dz[0], dz[1], dz[2], dz[3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (valid[j]) {
                            d_alpha[q0 + j] = da[j];
                            d_rgb[3 * (size_t)(q0 + j)] = dc[3 * j];
                            d_rgb[3 * (size_t)(q0 + j) + 1] = dc[3 * j + 1];
                            d_rgb[3 * (size_t)(q0 + j) + 2] = dc[3 * j + 2];
                            if (DZ) d_z[q0 + j] = dz[j];
                        }
                    }
                }
                Rcarry = __shfl_sync(m, V, 0, W);
            }
            ring_cursor_next<true, W>(cc, mine, m, lane);
        }
        cp_async_wait<0>();
        __syncwarp();
    }
}

// =============================================================================================
// tile family (short segments)
// =============================================================================================
// copy src[first, first+count) -> dst[(first - align4(first)) ...] with 16-byte loads for the aligned body.
// dst must be 16-byte aligned; src base 16-byte aligned.
__device__ __forceinline__ void stage_in(float* __restrict__ dst, const float* __restrict__ src, int64_t first, int count, int tid,
                                         int nthreads) {
    const int64_t last = first + count;
    const int64_t first_al = first & ~(int64_t)3;
    const int64_t body0 = (first + 3) & ~(int64_t)3;
    const int64_t body1 = last & ~(int64_t)3;
    if (body1 > body0) {
        const int nvec = (int)((body1 - body0) >> 2);
        const float4* s4 = reinterpret_cast<const float4*>(src + body0);
        float4* d4 = reinterpret_cast<float4*>(dst + (body0 - first_al));
        for (int v = tid; v < nvec; v += nthreads) d4[v] = ld_stream4(s4 + v);
        // head [first, body0) and tail [body1, last): at most 3 elements each
        if (tid < 3) {
            int64_t e = first + tid;
            if (e < body0) dst[e - first_al] = ld_stream(src + e);
        } else if (tid < 6) {
            int64_t e = body1 + (tid - 3);
            if (e < last) dst[e - first_al] = ld_stream(src + e);
        }
    } else {
        for (int64_t e = first + tid; e < last; e += nthreads) dst[e - first_al] = ld_stream(src + e);
    }
}

__device__ __forceinline__ void stage_out(float* __restrict__ dst, const float* __restrict__ src_smem, int64_t first, int count, int tid,
                                          int nthreads) {
    const int64_t last = first + count;
    const int64_t first_al = first & ~(int64_t)3;
    const int64_t body0 = (first + 3) & ~(int64_t)3;
    const int64_t body1 = last & ~(int64_t)3;
    if (body1 > body0) {
        const int nvec = (int)((body1 - body0) >> 2);
        float4* d4 = reinterpret_cast<float4*>(dst + body0);
        const float4* s4 = reinterpret_cast<const float4*>(src_smem + (body0 - first_al));
        for (int v = tid; v < nvec; v += nthreads) st_stream4(d4 + v, s4[v]);
        if (tid < 3) {
            int64_t e = first + tid;
            if (e < body0) st_stream(dst + e, src_smem[e - first_al]);
        } else if (tid < 6) {
            int64_t e = body1 + (tid - 3);
            if (e < last) st_stream(dst + e, src_smem[e - first_al]);
        }
    } else {
        for (int64_t e = first + tid; e < last; e += nthreads) st_stream(dst + e, src_smem[e - first_al]);
    }
}

// TMA flavour of stage_in: the 16-byte aligned span [first & ~3, (first+count) & ~3) travels as ONE bulk copy issued by the
// calling thread (completion on `bar`); tma_span_bytes() is what the caller adds to the barrier's expected byte count.
// The <= 3 elements behind the span are plain loads (stage_in_tail, any 3 threads).
__device__ __forceinline__ uint32_t tma_span_bytes(int64_t first, int count) {
    return (uint32_t)((((first + count) & ~(int64_t)3) - (first & ~(int64_t)3)) * 4);
}
__device__ __forceinline__ void tma_stage_in(float* dst, const float* src, int64_t first, int count, uint64_t* bar) {
    const uint32_t bytes = tma_span_bytes(first, count);
    if (bytes) bulk_g2s(dst, src + (first & ~(int64_t)3), bytes, bar);
}
__device__ __forceinline__ void stage_in_tail(float* dst, const float* src, int64_t first, int count, int k /*0..2*/) {
    const int64_t last = first + count;
    const int64_t e = (last & ~(int64_t)3) + k;
    if (e < last && e >= first) dst[e - (first & ~(int64_t)3)] = ld_stream(src + e);
}
// TMA flavour of stage_out: aligned body [roundup4(first), rounddown4(last)) as one bulk store by the calling thread
__device__ __forceinline__ void tma_stage_out(float* dst, const float* src_smem, int64_t first, int count) {
    const int64_t last = first + count;
    const int64_t body0 = (first + 3) & ~(int64_t)3, body1 = last & ~(int64_t)3;
    if (body1 > body0) bulk_s2g(dst + body0, src_smem + (body0 - (first & ~(int64_t)3)), (uint32_t)((body1 - body0) * 4));
}
// head [first, roundup4(first)) and tail [rounddown4(last), last) of an output range: k = 0..5 (any 6 threads)
__device__ __forceinline__ void stage_out_edges(float* dst, const float* src_smem, int64_t first, int count, int k) {
    const int64_t last = first + count;
    const int64_t first_al = first & ~(int64_t)3;
    const int64_t body0 = (first + 3) & ~(int64_t)3, body1 = last & ~(int64_t)3;
    if (body1 > body0) {
        const int64_t e = k < 3 ? first + k : body1 + (k - 3);
        const bool ok = k < 3 ? e < body0 : e < last;
        if (ok) st_stream(dst + e, src_smem[e - first_al]);
    } else {  // no aligned body: fewer than 8 elements in total
        for (int64_t e = first + k; e < last; e += 6) st_stream(dst + e, src_smem[e - first_al]);
    }
}

// block-wide: lowest start / highest end / sum of counts over the tile's non-empty rays
struct TileRange {
    int lo, hi, total;
};

__device__ __forceinline__ TileRange tile_range(int start, int n, int* red /* 3*8 ints */) {
    int lo = n > 0 ? start : 0x7fffffff;
    int hi = n > 0 ? start + n : -1;
    int tot = n;
    lo = __reduce_min_sync(VS_FULL_MASK, lo);
    hi = __reduce_max_sync(VS_FULL_MASK, hi);
    tot = __reduce_add_sync(VS_FULL_MASK, tot);
    const int wid = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        red[wid] = lo;
        red[8 + wid] = hi;
        red[16 + wid] = tot;
    }
    __syncthreads();
    TileRange r;
    r.lo = red[0];
    r.hi = red[8];
    r.total = red[16];
#pragma unroll
    for (int k = 1; k < kTileRays / 32; ++k) {
        r.lo = min(r.lo, red[k]);
        r.hi = max(r.hi, red[8 + k]);
        r.total += red[16 + k];
    }
    return r;
}

template <bool TMA>
__global__ void __launch_bounds__(kTileRays) composite_fwd_tile_kernel(
    const int32_t* __restrict__ se, const float* __restrict__ alpha, const float* __restrict__ rgb, const float* __restrict__ z,
    float* __restrict__ out_rgb, float* __restrict__ out_depth, float* __restrict__ out_acc, float* __restrict__ out_bgT,
    float* __restrict__ out_w, float* __restrict__ out_T, int64_t n_rays, int cap) {
    extern __shared__ __align__(16) float smem[];
    __shared__ int red[24];
    __shared__ __align__(8) uint64_t bar;
    float* s_alpha = smem;                 // cap+4
    float* s_z = s_alpha + (cap + 4);      // cap+4
    float* s_rgb = s_z + (cap + 4);        // 3*cap+4
    float* s_out = s_rgb + (3 * cap + 4);  // 3*kTileRays: per-ray colours, written back coalesced

    const int tid = threadIdx.x;
    if (TMA && tid == 0) mbar_init(&bar, 1);
    const int64_t ray0 = (int64_t)blockIdx.x * kTileRays;
    const int64_t ray = ray0 + tid;
    int start = 0, n = 0;
    if (ray < n_rays) n = load_segment(se, ray, start);
    const TileRange tr = tile_range(start, n, red);
    const int count = tr.hi - tr.lo;
    const bool staged = tr.total > 0 && count <= cap && count == tr.total;  // contiguous and fits

    // base pointers + element offsets of this ray's first sample (shared-memory tile or global memory)
    const float* pa = alpha;
    const float* pz = z;
    const float* pc = rgb;
    int64_t ia = start, ic = 3 * (int64_t)start;
    if (staged) {
        if (TMA) {
            if (tid == 0) {
                mbar_arrive_expect_tx(&bar, 2 * tma_span_bytes(tr.lo, count) + tma_span_bytes(3 * (int64_t)tr.lo, 3 * count));
                tma_stage_in(s_alpha, alpha, tr.lo, count, &bar);
                tma_stage_in(s_rgb, rgb, 3 * (int64_t)tr.lo, 3 * count, &bar);
                tma_stage_in(s_z, z, tr.lo, count, &bar);
            } else if (tid >= 32 && tid < 35) {
                stage_in_tail(s_alpha, alpha, tr.lo, count, tid - 32);
            } else if (tid >= 64 && tid < 67) {
                stage_in_tail(s_z, z, tr.lo, count, tid - 64);
            } else if (tid >= 96 && tid < 99) {
                stage_in_tail(s_rgb, rgb, 3 * (int64_t)tr.lo, 3 * count, tid - 96);
            }
            __syncthreads();
            mbar_wait(&bar, 0);
        } else {
            stage_in(s_alpha, alpha, tr.lo, count, tid, kTileRays);
            stage_in(s_z, z, tr.lo, count, tid, kTileRays);
            stage_in(s_rgb, rgb, 3 * (int64_t)tr.lo, 3 * count, tid, kTileRays);
            __syncthreads();
        }
        pa = s_alpha;
        pz = s_z;
        pc = s_rgb;
        ia = start - (tr.lo & ~3);
        ic = 3 * (int64_t)start - ((3 * (int64_t)tr.lo) & ~(int64_t)3);
    }

    float T = 1.f, ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aa = 0.f;
    for (int i = 0; i < n; ++i) {
        const float a = pa[ia + i];
        const float w = T * a;
        ar = fmaf(w, pc[ic + 3 * i], ar);
        ag = fmaf(w, pc[ic + 3 * i + 1], ag);
        ab = fmaf(w, pc[ic + 3 * i + 2], ab);
        ad = fmaf(w, pz[ia + i], ad);
        aa += w;
        if (out_w) out_w[(int64_t)start + i] = w;
        if (out_T) out_T[(int64_t)start + i] = T;
        T *= (1.f - a);
    }
    // per-ray outputs through shared memory so that the [N,3] colour rows leave as full 128-byte lines
    s_out[3 * tid] = ar;
    s_out[3 * tid + 1] = ag;
    s_out[3 * tid + 2] = ab;
    __syncthreads();
    const int64_t rays_here = min((int64_t)kTileRays, n_rays - ray0);
    for (int e = tid; e < 3 * rays_here; e += kTileRays) out_rgb[3 * ray0 + e] = s_out[e];
    if (ray < n_rays) {
        out_depth[ray] = ad;
        out_acc[ray] = aa;
        out_bgT[ray] = T;
    }
}

template <bool TMA>
__global__ void __launch_bounds__(kTileRays) composite_bwd_tile_kernel(
    const int32_t* __restrict__ se, const float* __restrict__ alpha, const float* __restrict__ rgb, const float* __restrict__ z,
    const float* __restrict__ g_rgb, const float* __restrict__ g_depth, const float* __restrict__ g_acc, const float* __restrict__ g_bgT,
    float* __restrict__ d_alpha, float* __restrict__ d_rgb, float* __restrict__ d_z, int64_t n_rays, int cap) {
    extern __shared__ __align__(16) float smem[];
    __shared__ int red[24];
    float* s_alpha = smem;                 // cap+4  (becomes d_alpha)
    float* s_z = s_alpha + (cap + 4);      // cap+4  (becomes d_z when requested)
    float* s_rgb = s_z + (cap + 4);        // 3*cap+4 (becomes d_rgb)
    float* s_T = s_rgb + (3 * cap + 4);    // cap+4
    float* s_g = s_T + (cap + 4);          // 3*kTileRays staged g_rgb
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    if (TMA && tid == 0) mbar_init(&bar, 1);
    const int64_t ray0 = (int64_t)blockIdx.x * kTileRays;
    const int64_t ray = ray0 + tid;
    int start = 0, n = 0;
    if (ray < n_rays) n = load_segment(se, ray, start);
    const TileRange tr = tile_range(start, n, red);
    if (tr.total == 0) return;
    const int count = tr.hi - tr.lo;
    const bool staged = count <= cap && count == tr.total;

    const int64_t rays_here = min((int64_t)kTileRays, n_rays - ray0);
    for (int e = tid; e < 3 * rays_here; e += kTileRays) s_g[e] = __ldg(g_rgb + 3 * ray0 + e);
    float gd = 0.f, ga = 0.f, R = 0.f;
    if (ray < n_rays) {
        gd = __ldg(g_depth + ray);
        ga = __ldg(g_acc + ray);
        R = __ldg(g_bgT + ray);
    }

    if (staged) {
        if (TMA) {
            if (tid == 0) {
                mbar_arrive_expect_tx(&bar, 2 * tma_span_bytes(tr.lo, count) + tma_span_bytes(3 * (int64_t)tr.lo, 3 * count));
                tma_stage_in(s_alpha, alpha, tr.lo, count, &bar);
                tma_stage_in(s_rgb, rgb, 3 * (int64_t)tr.lo, 3 * count, &bar);
                tma_stage_in(s_z, z, tr.lo, count, &bar);
            } else if (tid >= 32 && tid < 35) {
                stage_in_tail(s_alpha, alpha, tr.lo, count, tid - 32);
            } else if (tid >= 64 && tid < 67) {
                stage_in_tail(s_z, z, tr.lo, count, tid - 64);
            } else if (tid >= 96 && tid < 99) {
                stage_in_tail(s_rgb, rgb, 3 * (int64_t)tr.lo, 3 * count, tid - 96);
            }
            __syncthreads();
            mbar_wait(&bar, 0);
        } else {
            stage_in(s_alpha, alpha, tr.lo, count, tid, kTileRays);
            stage_in(s_z, z, tr.lo, count, tid, kTileRays);
            stage_in(s_rgb, rgb, 3 * (int64_t)tr.lo, 3 * count, tid, kTileRays);
            __syncthreads();
        }
        const int ia = start - (tr.lo & ~3);
        const int ic = (int)(3 * (int64_t)start - ((3 * (int64_t)tr.lo) & ~(int64_t)3));
        const float gr = s_g[3 * tid], gg = s_g[3 * tid + 1], gb = s_g[3 * tid + 2];
        float T = 1.f;
        for (int i = 0; i < n; ++i) {
            s_T[ia + i] = T;
            T *= (1.f - s_alpha[ia + i]);
        }
        for (int i = n - 1; i >= 0; --i) {
            const float a = s_alpha[ia + i];
            const float Ti = s_T[ia + i];
            const float cr = s_rgb[ic + 3 * i], cg = s_rgb[ic + 3 * i + 1], cb = s_rgb[ic + 3 * i + 2];
            const float zz = s_z[ia + i];
            const float gi = fmaf(gr, cr, fmaf(gg, cg, fmaf(gb, cb, fmaf(gd, zz, ga))));
            const float w = Ti * a;
            s_alpha[ia + i] = Ti * (gi - R);
            s_rgb[ic + 3 * i] = gr * w;
            s_rgb[ic + 3 * i + 1] = gg * w;
            s_rgb[ic + 3 * i + 2] = gb * w;
            s_z[ia + i] = gd * w;
            R = fmaf(1.f - a, R, a * gi);
        }
        if (TMA) {
            fence_proxy_async();  // make this thread's shared-memory writes visible to the bulk-copy engine
            __syncthreads();
            if (tid == 0) {
                tma_stage_out(d_alpha, s_alpha, tr.lo, count);
                tma_stage_out(d_rgb, s_rgb, 3 * (int64_t)tr.lo, 3 * count);
                if (d_z) tma_stage_out(d_z, s_z, tr.lo, count);
                bulk_commit();
                bulk_wait_read_all();  // shared memory must stay alive until the engine has read it
            } else if (tid >= 32 && tid < 38) {
                stage_out_edges(d_alpha, s_alpha, tr.lo, count, tid - 32);
            } else if (tid >= 64 && tid < 70) {
                stage_out_edges(d_rgb, s_rgb, 3 * (int64_t)tr.lo, 3 * count, tid - 64);
            } else if (tid >= 96 && tid < 102) {
                if (d_z) stage_out_edges(d_z, s_z, tr.lo, count, tid - 96);
            }
        } else {
            __syncthreads();
            stage_out(d_alpha, s_alpha, tr.lo, count, tid, kTileRays);
            stage_out(d_rgb, s_rgb, 3 * (int64_t)tr.lo, 3 * count, tid, kTileRays);
            if (d_z) stage_out(d_z, s_z, tr.lo, count, tid, kTileRays);
        }
    } else {
        __syncthreads();
        const float gr = s_g[3 * tid], gg = s_g[3 * tid + 1], gb = s_g[3 * tid + 2];
        // direct: T parked in d_alpha, then overwritten right to left
        float T = 1.f;
        for (int i = 0; i < n; ++i) {
            d_alpha[(int64_t)start + i] = T;
            T *= (1.f - alpha[(int64_t)start + i]);
        }
        for (int i = n - 1; i >= 0; --i) {
            const int64_t s = (int64_t)start + i;
            const float a = alpha[s];
            const float Ti = d_alpha[s];
            const float gi = fmaf(gr, rgb[3 * s], fmaf(gg, rgb[3 * s + 1], fmaf(gb, rgb[3 * s + 2], fmaf(gd, z[s], ga))));
            const float w = Ti * a;
            d_alpha[s] = Ti * (gi - R);
            d_rgb[3 * s] = gr * w;
            d_rgb[3 * s + 1] = gg * w;
            d_rgb[3 * s + 2] = gb * w;
            if (d_z) d_z[s] = gd * w;
            R = fmaf(1.f - a, R, a * gi);
        }
    }
}

// auto mode (measured on B200, profiles/r01_bench_composite_ring.jsonl): tile family up to a mean of 16 samples per ray (K = 9 shells:
// 75 % of the HBM peak against 26 % for the scan family), ring family beyond (C3: 61 % against 47 %), except the FORWARD of very long
// rays (mean > 256), where the one-sample-per-lane scan keeps more loads in flight per instruction (4.8 against 4.3 TB/s at mean 400)
constexpr double kTileMaxMean = 16.0;     // backward
constexpr double kTileMaxMeanFwd = 16.0;  // forward (mean 24: tile 0.128 ms, ring with 8-lane groups 0.120 ms)
constexpr double kRingFwdMaxMean = 256.0;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// lanes per ray of the coarsened scan kernels (4 samples per lane): a chunk of 4W samples should be of the order of the mean ray
static inline int pick_scan4_width(int64_t n_rays, int64_t n_samples) {
    const double mean = n_rays > 0 ? (double)n_samples / (double)n_rays : 0.0;
    return mean <= 24.0 ? 8 : (mean <= 96.0 ? 16 : 32);
}

// shared-memory capacity (in samples) of a tile: 1.25x the mean load of 256 rays, at least 256
// ring family: one warp per 32 rays; a few CTAs per SM, the rest of the batches are walked grid-stride
static inline unsigned ring_grid(int64_t n_rays) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t batches = (n_rays + 31) / 32;
    static const int mult = [] {
        const char* env = std::getenv("VS_RING_GRID_MULT");  // A/B knob: CTAs per SM in the grid (3 are resident at a time)
        const int v = env ? std::atoi(env) : 0;
        return v > 0 ? v : 12;  // measured on config C3: 3 -> 62.3 %, 6 -> 64.2 %, 12 -> 65.0 %, 24 -> 64.7 % of the HBM peak
    }();
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>((batches + kRingWarps - 1) / kRingWarps, (int64_t)sms * mult));
}
// lanes per ray of the ring kernels: a chunk (4W samples) should not be much longer than the typical ray, or most lanes of a chunk are
// masked; VS_RING_WIDTH overrides (A/B knob)
static inline int ring_width(int64_t n_rays, int64_t n_samples, bool backward) {
    static const int forced = [] {
        const char* env = std::getenv("VS_RING_WIDTH");
        const int v = env ? std::atoi(env) : 0;
        return (v == 8 || v == 16 || v == 32) ? v : 0;
    }();
    if (forced) return forced;
    // measured (profiles/r01_bench_composite_ring.jsonl): narrow groups pay off in the forward pass of short rays (mean 24: 0.120 ms with
    // W = 8, 0.167 ms with W = 32); in the backward pass the groups of a warp sit in different passes (transmittance / reverse) most of
    // the time and serialise (mean 24: 0.264 ms with W = 8, 0.232 ms with W = 32; C3: 0.649 ms with W = 16, 0.533 ms with W = 32)
    if (backward) return 32;
    const double mean = n_rays > 0 ? (double)n_samples / (double)n_rays : 0.0;
    return mean <= 40.0 ? 8 : 32;
}
static inline bool ring_in_auto() {
    static const bool off = std::getenv("VS_COMPOSITE_NO_RING") != nullptr;  // A/B knob: auto falls back to the scan family
    return !off;
}

static inline int tile_cap(int64_t n_rays, int64_t n_samples) {
    double mean = n_rays > 0 ? (double)n_samples / (double)n_rays : 0.0;
    int cap = (int)(mean * kTileRays * 1.25) + 64;
    cap = (cap + 3) & ~3;
    if (cap < 256) cap = 256;
    return cap;
}

}  // namespace vs

using namespace vs;

extern "C" {

// mode: 0 = auto (rules above kTileMaxMean), 1 = tile family (TMA bulk staging), 2 = scan family (one sample per lane, W from the mean ray length),
// 3 = coarsened scan family (aligned quad per lane, float4 loads; W auto), 4 = tile family with LDG/STS staging,
// 5/6/7 = coarsened scan with W = 8/16/32, 8 = ring family (cp.async-pipelined quads, a warp per 32 rays; what auto picks for
// mean ray lengths above 8).  3..7 exist for A/B measurements: on B200 the coarsened kernels measured SLOWER
// than the one-sample-per-lane ones (profiles/r01_bench_composite_scan_variants.jsonl), so auto never picks them.
int vs_composite_fwd(const int32_t* se, const float* alpha, const float* rgb, const float* z, float* out_rgb, float* out_depth,
                     float* out_acc, float* out_bgT, float* out_w, float* out_T, int64_t n_rays, int64_t n_samples, int mode,
                     void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && n_samples >= 0 && mode >= 0 && mode <= 8);
    if (n_rays == 0) return VS_OK;
    VS_CHECK_ARG(se && out_rgb && out_depth && out_acc && out_bgT);
    VS_CHECK_ARG(n_samples == 0 || (alpha && rgb && z));
    cudaStream_t st = (cudaStream_t)stream;
    const double mean = (double)n_samples / (double)n_rays;
    bool tile = (mode == 1) || (mode == 4) || (mode == 0 && mean <= kTileMaxMeanFwd);
    if (tile && !(aligned16(alpha) && aligned16(rgb) && aligned16(z))) tile = false;
    if (tile) {
        const int cap = tile_cap(n_rays, n_samples);
        const size_t smem = sizeof(float) * ((size_t)5 * cap + 12 + 3 * kTileRays);
        if (smem > 200 * 1024) {
            tile = false;
        } else {
            auto kern = mode == 4 ? composite_fwd_tile_kernel<false> : composite_fwd_tile_kernel<true>;
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            kern<<<(unsigned)div_up(n_rays, kTileRays), kTileRays, smem, st>>>(se, alpha, rgb, z, out_rgb, out_depth, out_acc, out_bgT, out_w,
                                                                              out_T, n_rays, cap);
            return launched(1);
        }
    }
    const bool a16 = aligned16(alpha) && aligned16(rgb) && aligned16(z);
    // ring family: cp.async-pipelined quads, 32 rays per warp.  Forward: on very long rays the one-sample-per-lane scan streams better
    if ((mode == 8 || (mode == 0 && ring_in_auto() && mean <= kRingFwdMaxMean)) && a16 && n_samples > 0 && n_samples <= 0x7fffffffLL) {
        const size_t smem = sizeof(float) * kRingWarps * kRingDepthDefault * kRingStageFloats;
        const bool wt = out_w || out_T;
        const int W = ring_width(n_rays, n_samples, false);
        constexpr int D = kRingDepthDefault;
        auto kern = W == 8    ? (wt ? composite_fwd_ring_kernel<true, D, 8> : composite_fwd_ring_kernel<false, D, 8>)
                    : W == 16 ? (wt ? composite_fwd_ring_kernel<true, D, 16> : composite_fwd_ring_kernel<false, D, 16>)
                              : (wt ? composite_fwd_ring_kernel<true, D, 32> : composite_fwd_ring_kernel<false, D, 32>);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<ring_grid(n_rays), 32 * kRingWarps, smem, st>>>(se, alpha, rgb, z, out_rgb, out_depth, out_acc, out_bgT, out_w, out_T, n_rays,
                                                              (int)n_samples);
        return launched(1);
    }
    if ((mode == 3 || (mode >= 5 && mode <= 7)) && a16) {  // coarsened scan kernels (quad per lane)
        const int W4 = mode == 5 ? 8 : mode == 6 ? 16 : mode == 7 ? 32 : pick_scan4_width(n_rays, n_samples);
        const unsigned grid4 = (unsigned)div_up(n_rays * W4, kScanThreads);
        switch (W4) {
            case 8:
                composite_fwd_scan4_kernel<8><<<grid4, kScanThreads, 0, st>>>(se, alpha, rgb, z, out_rgb, out_depth, out_acc, out_bgT, out_w, out_T, n_rays, n_samples);
                break;
            case 16:
                composite_fwd_scan4_kernel<16><<<grid4, kScanThreads, 0, st>>>(se, alpha, rgb, z, out_rgb, out_depth, out_acc, out_bgT, out_w, out_T, n_rays, n_samples);
                break;
            default:
                composite_fwd_scan4_kernel<32><<<grid4, kScanThreads, 0, st>>>(se, alpha, rgb, z, out_rgb, out_depth, out_acc, out_bgT, out_w, out_T, n_rays, n_samples);
                break;
        }
        return launched(1);
    }
    int Wsel = pick_group_width(n_rays, n_samples);
    if (Wsel < 8) Wsel = 8;
    const unsigned grid = (unsigned)div_up(n_rays * Wsel, kScanThreads);
    switch (Wsel) {
        case 8:
            composite_fwd_scan_kernel<8><<<grid, kScanThreads, 0, st>>>(se, alpha, rgb, z, out_rgb, out_depth, out_acc, out_bgT, out_w, out_T, n_rays);
            break;
        case 16:
            composite_fwd_scan_kernel<16><<<grid, kScanThreads, 0, st>>>(se, alpha, rgb, z, out_rgb, out_depth, out_acc, out_bgT, out_w, out_T, n_rays);
            break;
        default:
            composite_fwd_scan_kernel<32><<<grid, kScanThreads, 0, st>>>(se, alpha, rgb, z, out_rgb, out_depth, out_acc, out_bgT, out_w, out_T, n_rays);
            break;
    }
    return launched(1);
}

int vs_composite_bwd(const int32_t* se, const float* alpha, const float* rgb, const float* z, const float* g_rgb, const float* g_depth,
                     const float* g_acc, const float* g_bgT, float* d_alpha, float* d_rgb, float* d_z, int64_t n_rays, int64_t n_samples,
                     int mode, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && n_samples >= 0 && mode >= 0 && mode <= 8);
    if (n_rays == 0 || n_samples == 0) return VS_OK;
    VS_CHECK_ARG(se && alpha && rgb && z && g_rgb && g_depth && g_acc && g_bgT && d_alpha && d_rgb);
    cudaStream_t st = (cudaStream_t)stream;
    const double mean = (double)n_samples / (double)n_rays;
    bool tile = (mode == 1) || (mode == 4) || (mode == 0 && mean <= kTileMaxMean);
    if (tile && !(aligned16(alpha) && aligned16(rgb) && aligned16(z) && aligned16(d_alpha) && aligned16(d_rgb) && (!d_z || aligned16(d_z))))
        tile = false;
    if (tile) {
        const int cap = tile_cap(n_rays, n_samples);
        const size_t smem = sizeof(float) * ((size_t)6 * cap + 16 + 3 * kTileRays);
        if (smem > 200 * 1024) {
            tile = false;
        } else {
            auto kern = mode == 4 ? composite_bwd_tile_kernel<false> : composite_bwd_tile_kernel<true>;
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            kern<<<(unsigned)div_up(n_rays, kTileRays), kTileRays, smem, st>>>(se, alpha, rgb, z, g_rgb, g_depth, g_acc, g_bgT, d_alpha, d_rgb,
                                                                              d_z, n_rays, cap);
            return launched(1);
        }
    }
    const bool a16 = aligned16(alpha) && aligned16(rgb) && aligned16(z) && aligned16(d_alpha) && aligned16(d_rgb) && (!d_z || aligned16(d_z));
    if ((mode == 8 || (mode == 0 && ring_in_auto())) && a16 && n_samples <= 0x7fffffffLL) {
        const size_t smem = sizeof(float) * kRingWarps * kRingDepthDefault * kRingStageFloats;
        const int W = ring_width(n_rays, n_samples, true);
        constexpr int D = kRingDepthDefault;
        auto kern = W == 8    ? (d_z ? composite_bwd_ring_kernel<true, D, 8> : composite_bwd_ring_kernel<false, D, 8>)
                    : W == 16 ? (d_z ? composite_bwd_ring_kernel<true, D, 16> : composite_bwd_ring_kernel<false, D, 16>)
                              : (d_z ? composite_bwd_ring_kernel<true, D, 32> : composite_bwd_ring_kernel<false, D, 32>);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<ring_grid(n_rays), 32 * kRingWarps, smem, st>>>(se, alpha, rgb, z, g_rgb, g_depth, g_acc, g_bgT, d_alpha, d_rgb, d_z, n_rays,
                                                              (int)n_samples);
        return launched(1);
    }
    if ((mode == 3 || (mode >= 5 && mode <= 7)) && a16) {
        const int W4 = mode == 5 ? 8 : mode == 6 ? 16 : mode == 7 ? 32 : pick_scan4_width(n_rays, n_samples);
        const unsigned grid4 = (unsigned)div_up(n_rays * W4, kScanThreads);
        switch (W4) {
            case 8:
                composite_bwd_scan4_kernel<8><<<grid4, kScanThreads, 0, st>>>(se, alpha, rgb, z, g_rgb, g_depth, g_acc, g_bgT, d_alpha, d_rgb, d_z, n_rays, n_samples);
                break;
            case 16:
                composite_bwd_scan4_kernel<16><<<grid4, kScanThreads, 0, st>>>(se, alpha, rgb, z, g_rgb, g_depth, g_acc, g_bgT, d_alpha, d_rgb, d_z, n_rays, n_samples);
                break;
            default:
                composite_bwd_scan4_kernel<32><<<grid4, kScanThreads, 0, st>>>(se, alpha, rgb, z, g_rgb, g_depth, g_acc, g_bgT, d_alpha, d_rgb, d_z, n_rays, n_samples);
                break;
        }
        return launched(1);
    }
    int Wsel = pick_group_width(n_rays, n_samples);
    if (Wsel < 8) Wsel = 8;
    const unsigned grid = (unsigned)div_up(n_rays * Wsel, kScanThreads);
    switch (Wsel) {
        case 8:
            composite_bwd_scan_kernel<8><<<grid, kScanThreads, 0, st>>>(se, alpha, rgb, z, g_rgb, g_depth, g_acc, g_bgT, d_alpha, d_rgb, d_z, n_rays);
            break;
        case 16:
            composite_bwd_scan_kernel<16><<<grid, kScanThreads, 0, st>>>(se, alpha, rgb, z, g_rgb, g_depth, g_acc, g_bgT, d_alpha, d_rgb, d_z, n_rays);
            break;
        default:
            composite_bwd_scan_kernel<32><<<grid, kScanThreads, 0, st>>>(se, alpha, rgb, z, g_rgb, g_depth, g_acc, g_bgT, d_alpha, d_rgb, d_z, n_rays);
            break;
    }
    return launched(1);
}

}  // extern "C"
