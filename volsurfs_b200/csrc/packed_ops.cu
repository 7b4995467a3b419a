// volsurfs_b200 — op-by-op packed VolumeRendering operators (drop-in semantics of the
// reference's static methods, src/VolumeRendering.cu + kernels/volsurfs/VolumeRenderingGPU.cuh).
//
// The reference runs one THREAD per ray with a serial loop, so neighbouring lanes read
// addresses s_r*4 bytes apart.  Here a GROUP of W lanes owns one ray (W picked from the
// mean segment length, 32/W rays per warp): every load/store of a chunk is contiguous,
// and the per-ray recurrences are shuffle-based segmented scans whose running value is
// carried from chunk to chunk.  HBM-bound streaming kernels; no shared memory needed.
//
// Every kernel writes its per-ray outputs for EVERY ray (defaults for empty rays), so
// callers may pass uninitialised per-ray buffers.  Per-sample outputs are written for every
// sample that belongs to a ray.
#include "vs_common.cuh"

namespace vs {

constexpr int kThreads = 256;
constexpr int kAhead = 4;  // chunks of a ray requested ahead in the scan kernels

#define VS_GROUP_SETUP(W)                                                        \
    const int lane = threadIdx.x & 31;                                           \
    const int gl = lane & (W - 1);                                               \
    const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / W;    \
    int start = 0;                                                               \
    int n = 0;                                                                   \
    if (ray < n_rays) n = load_segment(se, ray, start);                          \
    const int n_max = warp_max_i32(n);                                           \
    (void)lane;

// ---------------------------------------------------------------------------------------------
// cumprod_one_minus_alpha_to_transmittance  (VolumeRenderingGPU.cuh:28-78)
//   T_i = prod_{j<i} x_j,  bg = T_{s-1}  (last x not applied; empty ray: bg = 1)
// ---------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(kThreads) cumprod_fwd_kernel(const int32_t* __restrict__ se, const float* __restrict__ x,
                                                               float* __restrict__ T, float* __restrict__ bg, int64_t n_rays) {
    VS_GROUP_SETUP(W)
    float carry = 1.f;
    // kAhead chunks of the ray are requested before the first is scanned: with one 128-byte line in flight per warp a full SM keeps
    // 8 KB outstanding, a third of what the HBM latency-bandwidth product asks for (round 1: 0.22 of the peak on config[2] packets)
    for (int base0 = 0; base0 < n_max; base0 += kAhead * W) {
        float xv[kAhead];
#pragma unroll
        for (int u = 0; u < kAhead; ++u) {
            const int i = base0 + u * W + gl;
            xv[u] = i < n ? ld_stream(x + start + i) : 1.f;
        }
#pragma unroll
        for (int u = 0; u < kAhead; ++u) {
            const int base = base0 + u * W;
            if (base >= n_max) break;  // warp-uniform
            const int i = base + gl;
            const bool valid = i < n;
            float incl = group_scan_mul<W>(xv[u], gl);
            float excl = group_shift_up<W>(incl, gl, 1.f);
            float Ti = carry * excl;
            if (valid) {
                st_stream(T + start + i, Ti);
                if (i == n - 1) bg[ray] = Ti;
            }
            carry *= group_bcast<W>(incl, W - 1);
        }
    }
    if (ray < n_rays && n == 0 && gl == 0) bg[ray] = 1.f;
}

// ---------------------------------------------------------------------------------------------
// integrate_with_weights_{1d,3d}  (VolumeRenderingGPU.cuh:80-177): out_r = sum_i w_i v_i
// ---------------------------------------------------------------------------------------------
template <int W, int D>
__global__ void __launch_bounds__(kThreads) integrate_fwd_kernel(const int32_t* __restrict__ se, const float* __restrict__ v,
                                                                 const float* __restrict__ w, float* __restrict__ out, int64_t n_rays) {
    VS_GROUP_SETUP(W)
    float acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = 0.f;
    for (int base = 0; base < n_max; base += W) {
        const int i = base + gl;
        if (i < n) {
            const int64_t s = (int64_t)start + i;
            float wi = ld_stream(w + s);
#pragma unroll
            for (int c = 0; c < D; ++c) acc[c] = fmaf(wi, ld_stream(v + s * D + c), acc[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = group_reduce_add<W>(acc[c]);
    if (ray < n_rays && gl == 0) {
#pragma unroll
        for (int c = 0; c < D; ++c) out[ray * D + c] = acc[c];
    }
}

// backward (VolumeRenderingGPU.cuh:945-1033): dv_i = g_r w_i ; dw_i = g_r . v_i
// ref_bug != 0 reproduces the reference's 3-D kernel reading channel [1] for z (:1021)
template <int W, int D>
__global__ void __launch_bounds__(kThreads) integrate_bwd_kernel(const int32_t* __restrict__ se, const float* __restrict__ g,
                                                                 const float* __restrict__ v, const float* __restrict__ w,
                                                                 float* __restrict__ dv, float* __restrict__ dw, int64_t n_rays,
                                                                 int ref_bug) {
    VS_GROUP_SETUP(W)
    float gr[D];
#pragma unroll
    for (int c = 0; c < D; ++c) gr[c] = (ray < n_rays) ? __ldg(g + ray * D + c) : 0.f;
    for (int base = 0; base < n_max; base += W) {
        const int i = base + gl;
        if (i < n) {
            const int64_t s = (int64_t)start + i;
            float wi = ld_stream(w + s);
            float vv[D];
#pragma unroll
            for (int c = 0; c < D; ++c) vv[c] = ld_stream(v + s * D + c);
            if (D == 3 && ref_bug) vv[D - 1] = vv[1];
            float dwi = __fmul_rn(gr[0], vv[0]);
#pragma unroll
            for (int c = 1; c < D; ++c) dwi = __fadd_rn(dwi, __fmul_rn(gr[c], vv[c]));
#pragma unroll
            for (int c = 0; c < D; ++c) st_stream(dv + s * D + c, gr[c] * wi);
            st_stream(dw + s, dwi);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// sum_over_rays<val_dim>  (VolumeRenderingGPU.cuh:245-303) and backward (:1035-1079)
// ---------------------------------------------------------------------------------------------
template <int W, int D>
__global__ void __launch_bounds__(kThreads) sum_fwd_kernel(const int32_t* __restrict__ se, const float* __restrict__ v,
                                                           float* __restrict__ sum_ray, float* __restrict__ sum_sample, int64_t n_rays) {
    VS_GROUP_SETUP(W)
    float acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = 0.f;
    for (int base = 0; base < n_max; base += W) {
        const int i = base + gl;
        if (i < n) {
            const int64_t s = (int64_t)start + i;
#pragma unroll
            for (int c = 0; c < D; ++c) acc[c] += ld_stream(v + s * D + c);
        }
    }
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = group_reduce_add<W>(acc[c]);
    if (ray < n_rays && gl == 0) {
#pragma unroll
        for (int c = 0; c < D; ++c) sum_ray[ray * D + c] = acc[c];
    }
    if (sum_sample != nullptr) {
        for (int base = 0; base < n_max; base += W) {
            const int i = base + gl;
            if (i < n) {
                const int64_t s = (int64_t)start + i;
#pragma unroll
                for (int c = 0; c < D; ++c) st_stream(sum_sample + s * D + c, acc[c]);
            }
        }
    }
}

// val_dim == 32: one warp per ray, lane == channel (row of 32 floats is one 128-byte line)
__global__ void __launch_bounds__(kThreads) sum_fwd32_kernel(const int32_t* __restrict__ se, const float* __restrict__ v,
                                                             float* __restrict__ sum_ray, float* __restrict__ sum_sample, int64_t n_rays) {
    const int lane = threadIdx.x & 31;
    const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray >= n_rays) return;
    int start;
    const int n = load_segment(se, ray, start);
    float acc = 0.f;
    for (int i = 0; i < n; ++i) acc += ld_stream(v + ((int64_t)start + i) * 32 + lane);
    sum_ray[ray * 32 + lane] = acc;
    if (sum_sample != nullptr)
        for (int i = 0; i < n; ++i) st_stream(sum_sample + ((int64_t)start + i) * 32 + lane, acc);
}

template <int W, int D>
__global__ void __launch_bounds__(kThreads) sum_bwd_kernel(const int32_t* __restrict__ se, const float* __restrict__ g_ray,
                                                           const float* __restrict__ g_sample, float* __restrict__ dv, int64_t n_rays) {
    VS_GROUP_SETUP(W)
    float gr[D];
#pragma unroll
    for (int c = 0; c < D; ++c) gr[c] = (ray < n_rays) ? __ldg(g_ray + ray * D + c) : 0.f;
    for (int base = 0; base < n_max; base += W) {
        const int i = base + gl;
        if (i < n) {
            const int64_t s = (int64_t)start + i;
#pragma unroll
            for (int c = 0; c < D; ++c) st_stream(dv + s * D + c, gr[c] + ld_stream(g_sample + s * D + c));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// cumsum_over_rays  (VolumeRenderingGPU.cuh:305-361): inclusive, optionally from the ray's end
// ---------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(kThreads) cumsum_kernel(const int32_t* __restrict__ se, const float* __restrict__ v,
                                                          float* __restrict__ out, int64_t n_rays, int inverse) {
    VS_GROUP_SETUP(W)
    float carry = 0.f;
    for (int base0 = 0; base0 < n_max; base0 += kAhead * W) {  // kAhead chunks in flight (see cumprod_fwd_kernel)
        float vv[kAhead];
#pragma unroll
        for (int u = 0; u < kAhead; ++u) {
            const int i = base0 + u * W + gl;
            const int64_t s = inverse ? ((int64_t)start + n - 1 - i) : ((int64_t)start + i);
            vv[u] = i < n ? ld_stream(v + s) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kAhead; ++u) {
            const int base = base0 + u * W;
            if (base >= n_max) break;  // warp-uniform
            const int i = base + gl;
            const bool valid = i < n;
            const int64_t s = inverse ? ((int64_t)start + n - 1 - i) : ((int64_t)start + i);
            float incl = group_scan_add<W>(vv[u], gl);
            if (valid) st_stream(out + s, carry + incl);
            carry += group_bcast<W>(incl, W - 1);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// cumprod backward  (VolumeRenderingGPU.cuh:896-943): needs cumsumLV from the caller
// ---------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(kThreads) cumprod_bwd_kernel(const int32_t* __restrict__ se, const float* __restrict__ g_bg,
                                                               const float* __restrict__ x, const float* __restrict__ bg,
                                                               const float* __restrict__ cumsumLV, float* __restrict__ dx, int64_t n_rays) {
    VS_GROUP_SETUP(W)
    float gb = 0.f, b = 0.f;
    if (ray < n_rays && n > 0) {
        gb = __ldg(g_bg + ray);
        b = __ldg(bg + ray);
    }
    const float gbb = __fmul_rn(gb, b);
    for (int base = 0; base < n_max; base += W) {
        const int i = base + gl;
        if (i < n) {
            const int64_t s = (int64_t)start + i;
            float d = 0.f;
            if (i < n - 1) {
                float den = fmaxf(ld_stream(x + s), 1e-6f);
                d = __fdiv_rn(ld_stream(cumsumLV + s + 1), den);
                d = __fadd_rn(d, __fdiv_rn(gbb, den));
            }
            st_stream(dx + s, d);
        }
    }
}

// fused variant of the python half + kernel (volume_rendering_funcs.py:105-179):
// LV = gT*T, reverse inclusive cumsum P, dx_i = (P_{i+1} + g_bg*bg)/max(x_i,1e-6)
template <int W>
__global__ void __launch_bounds__(kThreads) cumprod_bwd_fused_kernel(const int32_t* __restrict__ se, const float* __restrict__ g_T,
                                                                     const float* __restrict__ g_bg, const float* __restrict__ x,
                                                                     const float* __restrict__ T, const float* __restrict__ bg,
                                                                     float* __restrict__ dx, int64_t n_rays) {
    VS_GROUP_SETUP(W)
    float gb = 0.f, b = 0.f;
    if (ray < n_rays && n > 0) {
        gb = __ldg(g_bg + ray);
        b = __ldg(bg + ray);
    }
    const float gbb = gb * b;
    float carry = 0.f;  // sum of LV over all samples to the right of the current chunk
    for (int base = 0; base < n_max; base += W) {
        const int i = base + gl;            // distance from the ray's end
        const bool valid = i < n;
        const int64_t s = (int64_t)start + n - 1 - i;
        float lv = valid ? ld_stream(g_T + s) * ld_stream(T + s) : 0.f;
        float incl = group_scan_add<W>(lv, gl);      // P_s - carry
        float excl = group_shift_up<W>(incl, gl, 0.f);  // P_{s+1} - carry
        if (valid) {
            float d = 0.f;
            if (i > 0) {  // not the ray's last sample
                float den = fmaxf(ld_stream(x + s), 1e-6f);
                d = (carry + excl) / den + gbb / den;
            }
            st_stream(dx + s, d);
        }
        carry += group_bcast<W>(incl, W - 1);
    }
}

// ---------------------------------------------------------------------------------------------
// update_dt  (RaySamplesPackedGPU.cuh:14-88)
// ---------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(kThreads) update_dt_kernel(const int32_t* __restrict__ se, const float* __restrict__ z,
                                                             const float* __restrict__ ray_exit, const float* __restrict__ ray_max_dt,
                                                             float* __restrict__ dt, int64_t n_rays, int is_background) {
    VS_GROUP_SETUP(W)
    float max_dt = 0.f, t_exit = 0.f;
    if (ray < n_rays && n > 0) {
        max_dt = __ldg(ray_max_dt + ray);
        t_exit = __ldg(ray_exit + ray);
    }
    for (int base = 0; base < n_max; base += W) {
        const int i = base + gl;
        if (i < n) {
            const int64_t s = (int64_t)start + i;
            float cur = ld_stream(z + s);
            float out;
            if (i < n - 1) {
                out = fmaxf(0.f, fminf(__fsub_rn(__ldg(z + s + 1), cur), max_dt));  // helper_math clamp(f, a, b) = fmaxf(a, fminf(f, b))
            } else if (is_background) {
                out = 1e10f;
            } else {
                out = fmaxf(0.f, fminf(__fsub_rn(t_exit, cur), max_dt));
            }
            st_stream(dt + s, out);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// "next" operators of the same container: NeuS alpha, median depth, cdf.  Their outputs depend on
// discrete, order-sensitive decisions (first sample whose running sum crosses a threshold, snapping of the last cdf entry), so
// they keep the reference's one-thread-per-ray left-to-right order instead of a re-associated scan.
// ---------------------------------------------------------------------------------------------
// VolumeRenderingGPU.cuh:185-243.  The reference mixes float variables with double literals; C's promotion rules are kept.  Every sample's
// alpha depends on its own dt / beta and on sdf[s], sdf[s+1] only (no running state), so a W-lane group walks the ray with unit stride
// and the per-sample arithmetic — hence the result — is the one of the reference's thread-per-ray loop, bit for bit.
template <int W>
__global__ void __launch_bounds__(kThreads) sdf2alpha_kernel(const int32_t* __restrict__ se, const float* __restrict__ dt,
                                                             const float* __restrict__ sdf, const float* __restrict__ beta,
                                                             float* __restrict__ alpha, int64_t n_rays) {
    VS_GROUP_SETUP(W)
    for (int base = 0; base < n_max; base += W) {
        const int i = base + gl;
        if (i + 1 < n) {
            const int64_t s = (int64_t)start + i;
            const float d = ld_stream(dt + s), prev = ld_stream(sdf + s), next = __ldg(sdf + s + 1);
            const float mid = (float)((double)(prev + next) * 0.5);
            float c = (float)((double)(next - prev) / ((double)d + 1e-6));
            c = fminf(fmaxf(c, -1e3f), 0.0f);
            const double half_step = (double)(c * d) * 0.5;
            const float prev_e = (float)((double)mid - half_step), next_e = (float)((double)mid + half_step);
            const float b = ld_stream(beta + s);
            const float pc = (float)(1.0 / (1.0 + (double)expf(-(prev_e * b))));
            const float nc = (float)(1.0 / (1.0 + (double)expf(-(next_e * b))));
            st_stream(alpha + s, (float)(((double)(pc - nc) + 1e-6) / ((double)pc + 1e-6)));
        }
    }
}

// VolumeRenderingGPU.cuh:364-409; ref_bug keeps the fallback index samples_z[nr_samples-1] (no idx_start, :407)
__global__ void __launch_bounds__(kThreads) median_depth_kernel(const int32_t* __restrict__ se, const float* __restrict__ z,
                                                                const float* __restrict__ w, float threshold, float* __restrict__ out,
                                                                int64_t n_rays, int ref_bug) {
    const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    int start;
    const int n = load_segment(se, ray, start);
    if (n == 0) return;  // output keeps its zero initialisation
    // the running sum is the reference's (one fp32 add per sample, in order); the loads of the next eight weights do not depend on it and
    // are issued together, so the thread pays one memory latency per eight samples instead of one per sample
    const float* __restrict__ wr = w + start;
    float run = 0.f;
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        float wv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) wv[j] = __ldg(wr + i + j);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            run = __fadd_rn(run, wv[j]);
            if (run >= threshold) {
                out[ray] = z[(int64_t)start + i + j];
                return;
            }
        }
    }
    for (; i < n; ++i) {
        run = __fadd_rn(run, __ldg(wr + i));
        if (run >= threshold) {
            out[ray] = z[(int64_t)start + i];
            return;
        }
    }
    out[ray] = ref_bug ? z[n - 1] : z[(int64_t)start + n - 1];
}

// VolumeRenderingGPU.cuh:412-471: exclusive running sum; rays with < 2 samples are skipped; last entry snapped to 1 when the
// weights sum to ~1 but the last cdf value is not ~1
__global__ void __launch_bounds__(kThreads) compute_cdf_kernel(const int32_t* __restrict__ se, const float* __restrict__ w,
                                                               float* __restrict__ cdf, int64_t n_rays) {
    const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    int start;
    const int n = load_segment(se, ray, start);
    if (n < 2) return;
    // the reference's running sum (one fp32 add per sample, in order); eight weights are loaded together ahead of the adds
    const float* __restrict__ wr = w + start;
    float* __restrict__ cr = cdf + start;
    float run = 0.f, last = 0.f;
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        float wv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) wv[j] = __ldg(wr + i + j);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            last = run;
            cr[i + j] = run;
            run = __fadd_rn(run, wv[j]);
        }
    }
    for (; i < n; ++i) {
        last = run;
        cr[i] = run;
        run = __fadd_rn(run, __ldg(wr + i));
    }
    if (fabs((double)run - 1.0) < 1e-3 && fabs((double)last - 1.0) > 1e-3) cdf[(int64_t)start + n - 1] = 1.0f;
}

static inline dim3 grid_for(int64_t n_rays, int W) { return dim3((unsigned)div_up(n_rays * W, kThreads)); }

}  // namespace vs

using namespace vs;

#define VS_DISPATCH_W(W_, ...)                 \
    switch (W_) {                              \
        case 4: { constexpr int W = 4; __VA_ARGS__; } break;   \
        case 8: { constexpr int W = 8; __VA_ARGS__; } break;   \
        case 16: { constexpr int W = 16; __VA_ARGS__; } break; \
        default: { constexpr int W = 32; __VA_ARGS__; } break; \
    }

extern "C" {

int vs_cumprod_fwd(const int32_t* se, const float* x, float* T, float* bgT, int64_t n_rays, int64_t n_samples, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && n_samples >= 0);
    if (n_rays == 0) return VS_OK;
    VS_CHECK_ARG(se && bgT && (n_samples == 0 || (x && T)));
    cudaStream_t st = (cudaStream_t)stream;
    int Wsel = pick_group_width(n_rays, n_samples);
    VS_DISPATCH_W(Wsel, cumprod_fwd_kernel<W><<<grid_for(n_rays, W), kThreads, 0, st>>>(se, x, T, bgT, n_rays));
    return launched(1);
}

int vs_integrate_fwd(const int32_t* se, const float* values, const float* weights, float* out, int dim, int64_t n_rays,
                     int64_t n_samples, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && n_samples >= 0);
    if (dim != 1 && dim != 3) return VS_ERR_UNSUPPORTED;
    if (n_rays == 0) return VS_OK;
    VS_CHECK_ARG(se && out && (n_samples == 0 || (values && weights)));
    cudaStream_t st = (cudaStream_t)stream;
    int Wsel = pick_group_width(n_rays, n_samples);
    if (dim == 1) {
        VS_DISPATCH_W(Wsel, integrate_fwd_kernel<W, 1><<<grid_for(n_rays, W), kThreads, 0, st>>>(se, values, weights, out, n_rays));
    } else {
        VS_DISPATCH_W(Wsel, integrate_fwd_kernel<W, 3><<<grid_for(n_rays, W), kThreads, 0, st>>>(se, values, weights, out, n_rays));
    }
    return launched(1);
}

int vs_integrate_bwd(const int32_t* se, const float* grad_out, const float* values, const float* weights, float* d_values,
                     float* d_weights, int dim, int64_t n_rays, int64_t n_samples, int ref_bug, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && n_samples >= 0);
    if (dim != 1 && dim != 3) return VS_ERR_UNSUPPORTED;
    if (n_rays == 0 || n_samples == 0) return VS_OK;
    VS_CHECK_ARG(se && grad_out && values && weights && d_values && d_weights);
    cudaStream_t st = (cudaStream_t)stream;
    int Wsel = pick_group_width(n_rays, n_samples);
    if (dim == 1) {
        VS_DISPATCH_W(Wsel, integrate_bwd_kernel<W, 1><<<grid_for(n_rays, W), kThreads, 0, st>>>(se, grad_out, values, weights, d_values,
                                                                                              d_weights, n_rays, ref_bug));
    } else {
        VS_DISPATCH_W(Wsel, integrate_bwd_kernel<W, 3><<<grid_for(n_rays, W), kThreads, 0, st>>>(se, grad_out, values, weights, d_values,
                                                                                              d_weights, n_rays, ref_bug));
    }
    return launched(1);
}

int vs_sum_fwd(const int32_t* se, const float* values, float* sum_ray, float* sum_sample, int dim, int64_t n_rays, int64_t n_samples,
               void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && n_samples >= 0);
    if (!(dim == 1 || dim == 2 || dim == 3 || dim == 32)) return VS_ERR_UNSUPPORTED;  // VolumeRendering.cu:243,307
    if (n_rays == 0) return VS_OK;
    VS_CHECK_ARG(se && sum_ray && (n_samples == 0 || values));
    cudaStream_t st = (cudaStream_t)stream;
    int Wsel = pick_group_width(n_rays, n_samples);
    if (dim == 32) {
        sum_fwd32_kernel<<<grid_for(n_rays, 32), kThreads, 0, st>>>(se, values, sum_ray, sum_sample, n_rays);
    } else if (dim == 1) {
        VS_DISPATCH_W(Wsel, sum_fwd_kernel<W, 1><<<grid_for(n_rays, W), kThreads, 0, st>>>(se, values, sum_ray, sum_sample, n_rays));
    } else if (dim == 2) {
        VS_DISPATCH_W(Wsel, sum_fwd_kernel<W, 2><<<grid_for(n_rays, W), kThreads, 0, st>>>(se, values, sum_ray, sum_sample, n_rays));
    } else {
        VS_DISPATCH_W(Wsel, sum_fwd_kernel<W, 3><<<grid_for(n_rays, W), kThreads, 0, st>>>(se, values, sum_ray, sum_sample, n_rays));
    }
    return launched(1);
}

int vs_sum_bwd(const int32_t* se, const float* grad_sum_ray, const float* grad_sum_sample, float* d_values, int dim, int64_t n_rays,
               int64_t n_samples, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && n_samples >= 0);
    if (!(dim == 1 || dim == 2 || dim == 3)) return VS_ERR_UNSUPPORTED;  // VolumeRendering.cu:882
    if (n_rays == 0 || n_samples == 0) return VS_OK;
    VS_CHECK_ARG(se && grad_sum_ray && grad_sum_sample && d_values);
    cudaStream_t st = (cudaStream_t)stream;
    int Wsel = pick_group_width(n_rays, n_samples);
    if (dim == 1) {
        VS_DISPATCH_W(Wsel, sum_bwd_kernel<W, 1><<<grid_for(n_rays, W), kThreads, 0, st>>>(se, grad_sum_ray, grad_sum_sample, d_values, n_rays));
    } else if (dim == 2) {
        VS_DISPATCH_W(Wsel, sum_bwd_kernel<W, 2><<<grid_for(n_rays, W), kThreads, 0, st>>>(se, grad_sum_ray, grad_sum_sample, d_values, n_rays));
    } else {
        VS_DISPATCH_W(Wsel, sum_bwd_kernel<W, 3><<<grid_for(n_rays, W), kThreads, 0, st>>>(se, grad_sum_ray, grad_sum_sample, d_values, n_rays));
    }
    return launched(1);
}

int vs_cumsum(const int32_t* se, const float* values, float* out, int inverse, int64_t n_rays, int64_t n_samples, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && n_samples >= 0);
    if (n_rays == 0 || n_samples == 0) return VS_OK;
    VS_CHECK_ARG(se && values && out);
    cudaStream_t st = (cudaStream_t)stream;
    int Wsel = pick_group_width(n_rays, n_samples);
    VS_DISPATCH_W(Wsel, cumsum_kernel<W><<<grid_for(n_rays, W), kThreads, 0, st>>>(se, values, out, n_rays, inverse));
    return launched(1);
}

int vs_cumprod_bwd(const int32_t* se, const float* grad_bgT, const float* x, const float* bgT, const float* cumsumLV, float* dx,
                   int64_t n_rays, int64_t n_samples, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && n_samples >= 0);
    if (n_rays == 0 || n_samples == 0) return VS_OK;
    VS_CHECK_ARG(se && grad_bgT && x && bgT && cumsumLV && dx);
    cudaStream_t st = (cudaStream_t)stream;
    int Wsel = pick_group_width(n_rays, n_samples);
    VS_DISPATCH_W(Wsel, cumprod_bwd_kernel<W><<<grid_for(n_rays, W), kThreads, 0, st>>>(se, grad_bgT, x, bgT, cumsumLV, dx, n_rays));
    return launched(1);
}

int vs_cumprod_bwd_fused(const int32_t* se, const float* grad_T, const float* grad_bgT, const float* x, const float* T, const float* bgT,
                         float* dx, int64_t n_rays, int64_t n_samples, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && n_samples >= 0);
    if (n_rays == 0 || n_samples == 0) return VS_OK;
    VS_CHECK_ARG(se && grad_T && grad_bgT && x && T && bgT && dx);
    cudaStream_t st = (cudaStream_t)stream;
    int Wsel = pick_group_width(n_rays, n_samples);
    VS_DISPATCH_W(Wsel,
                  cumprod_bwd_fused_kernel<W><<<grid_for(n_rays, W), kThreads, 0, st>>>(se, grad_T, grad_bgT, x, T, bgT, dx, n_rays));
    return launched(1);
}

int vs_update_dt(const int32_t* se, const float* samples_z, const float* ray_exit, const float* ray_max_dt, float* samples_dt,
                 int is_background, int64_t n_rays, int64_t n_samples, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && n_samples >= 0);
    if (n_rays == 0 || n_samples == 0) return VS_OK;
    VS_CHECK_ARG(se && samples_z && ray_exit && ray_max_dt && samples_dt);
    cudaStream_t st = (cudaStream_t)stream;
    int Wsel = pick_group_width(n_rays, n_samples);
    VS_DISPATCH_W(Wsel, update_dt_kernel<W><<<grid_for(n_rays, W), kThreads, 0, st>>>(se, samples_z, ray_exit, ray_max_dt, samples_dt,
                                                                                     n_rays, is_background));
    return launched(1);
}

// replaces VolumeRendering::sdf2alpha (src/VolumeRendering.cu:178-229); alpha must be zero-filled (last sample of a ray keeps 0)
int vs_sdf2alpha(const int32_t* se, const float* samples_dt, const float* samples_sdf, const float* logistic_beta, float* alpha,
                 int64_t n_rays, int64_t n_samples, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && n_samples >= 0);
    if (n_rays == 0 || n_samples == 0) return VS_OK;
    VS_CHECK_ARG(se && samples_dt && samples_sdf && logistic_beta && alpha);
    cudaStream_t st = (cudaStream_t)stream;
    int Wsel = pick_group_width(n_rays, n_samples);
    VS_DISPATCH_W(Wsel, sdf2alpha_kernel<W><<<grid_for(n_rays, W), kThreads, 0, st>>>(se, samples_dt, samples_sdf, logistic_beta, alpha, n_rays));
    return launched(1);
}

// replaces VolumeRendering::median_depth_over_rays (src/VolumeRendering.cu:372-416); out must be zero-filled
int vs_median_depth(const int32_t* se, const float* samples_z, const float* weights, float threshold, float* out, int64_t n_rays,
                    int64_t n_samples, int ref_bug, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && n_samples >= 0);
    if (n_rays == 0 || n_samples == 0) return VS_OK;
    VS_CHECK_ARG(se && samples_z && weights && out);
    median_depth_kernel<<<(unsigned)div_up(n_rays, kThreads), kThreads, 0, (cudaStream_t)stream>>>(se, samples_z, weights, threshold, out,
                                                                                                 n_rays, ref_bug);
    return launched(1);
}

// replaces VolumeRendering::compute_cdf (src/VolumeRendering.cu:418-465); cdf must be zero-filled
int vs_compute_cdf(const int32_t* se, const float* weights, float* cdf, int64_t n_rays, int64_t n_samples, void* stream) {
    VS_CHECK_ARG(n_rays >= 0 && n_samples >= 0);
    if (n_rays == 0 || n_samples == 0) return VS_OK;
    VS_CHECK_ARG(se && weights && cdf);
    // (round 2 tried a W-lane group per ray that loads coalesced chunks and replays the running sum in order with one shuffle + add per
    // sample, to keep the reference's fp32 rounding sequence: 0.42 ms against 0.23 ms for this thread-per-ray kernel on config[2] packets —
    // the 32-deep dependent shuffle chain per chunk is worse than the strided loads it removes)
    compute_cdf_kernel<<<(unsigned)div_up(n_rays, kThreads), kThreads, 0, (cudaStream_t)stream>>>(se, weights, cdf, n_rays);
    return launched(1);
}

}  // extern "C"
