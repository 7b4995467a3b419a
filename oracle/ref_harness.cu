// ORACLE (test infrastructure only — never linked into or loaded by the product).
//
// A thin C harness around the REFERENCE's own packed volume-rendering kernels: it #includes
//   /root/reference/kernels/volsurfs/VolumeRenderingGPU.cuh          (the kernels, unmodified, where they lie)
//   /root/reference/kernels/volsurfs/pcg32.h                          (the RNG of importance_sample)
// and launches every __global__ function in that header with the launch shape the reference host code uses
// (src/VolumeRendering.cu: blocks = div_round_up(nr_rays, 256), 256 threads, legacy stream, device synchronise).
// Built by oracle/build.py into oracle/_ref/libvolsurfs_ref.so (git-ignored; travels to the GPU box with the snapshot).
// tests/test_gpu_reference_kernels.py compares the product's kernels with these on identical inputs, which pins the CPU
// restatement (oracle/compositing.py) AND the CUDA path to the reference implementation itself.
//
// No reference source is copied here: this file only builds accessors over raw device pointers and forwards them.
#include <cstdint>

#include "volsurfs/pcg32.h"
#include "volsurfs/VolumeRenderingGPU.cuh"

namespace {

template <typename T, int N>
using Acc = torch::PackedTensorAccessor32<T, N, torch::RestrictPtrTraits>;

template <typename T>
Acc<T, 2> acc2(const T* p, int64_t rows, int64_t cols) {
    const int64_t sizes[2] = {rows, cols};
    const int64_t strides[2] = {cols, 1};
    return Acc<T, 2>(const_cast<T*>(p), sizes, strides);
}
template <typename T>
Acc<T, 1> acc1(const T* p, int64_t n) {
    const int64_t sizes[1] = {n};
    const int64_t strides[1] = {1};
    return Acc<T, 1>(const_cast<T*>(p), sizes, strides);
}

inline dim3 grid_for(int nr_rays) { return dim3((unsigned)((nr_rays + BLOCK_SIZE - 1) / BLOCK_SIZE), 1, 1); }

inline int finish() {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    return (int)e;
}

}  // namespace

extern "C" {

int ref_abi_version() { return 1; }

// VolumeRendering.cu:30-78
int ref_cumprod_fwd(const int* se, const float* x, float* T, float* bgT, int nr_rays, int nr_samples) {
    VolumeRenderingGPU::cumprod_one_minus_alpha_to_transmittance_gpu<<<grid_for(nr_rays), BLOCK_SIZE>>>(
        nr_rays, nr_samples, acc2(se, nr_rays, 2), acc2(x, nr_samples, 1), acc2(T, nr_samples, 1), acc2(bgT, nr_rays, 1));
    return finish();
}

// VolumeRendering.cu:80-176
int ref_integrate_fwd(const int* se, const float* v, const float* w, float* out, int dim, int nr_rays, int nr_samples) {
    if (dim == 1)
        VolumeRenderingGPU::integrate_with_weights_1d_gpu<<<grid_for(nr_rays), BLOCK_SIZE>>>(
            nr_rays, nr_samples, acc2(se, nr_rays, 2), acc2(v, nr_samples, 1), acc2(w, nr_samples, 1), acc2(out, nr_rays, 1));
    else if (dim == 3)
        VolumeRenderingGPU::integrate_with_weights_3d_gpu<<<grid_for(nr_rays), BLOCK_SIZE>>>(
            nr_rays, nr_samples, acc2(se, nr_rays, 2), acc2(v, nr_samples, 3), acc2(w, nr_samples, 1), acc2(out, nr_rays, 3));
    else
        return -1;
    return finish();
}

// VolumeRendering.cu:178-229
int ref_sdf2alpha(const int* se, const float* dt, const float* sdf, const float* beta, float* alpha, int nr_rays, int nr_samples) {
    VolumeRenderingGPU::sdf2alpha_gpu<<<grid_for(nr_rays), BLOCK_SIZE>>>(nr_rays, nr_samples, acc2(se, nr_rays, 2), acc2(dt, nr_samples, 1),
                                                                         acc2(sdf, nr_samples, 1), acc2(beta, nr_samples, 1),
                                                                         acc2(alpha, nr_samples, 1));
    return finish();
}

// VolumeRendering.cu:231-324
int ref_sum_fwd(const int* se, const float* v, float* per_ray, float* per_sample, int dim, int nr_rays, int nr_samples) {
    auto a_se = acc2(se, nr_rays, 2);
    auto a_v = acc2(v, nr_samples, dim);
    auto a_r = acc2(per_ray, nr_rays, dim);
    auto a_s = acc2(per_sample, nr_samples, dim);
    const dim3 g = grid_for(nr_rays);
    if (dim == 1)
        VolumeRenderingGPU::sum_over_rays_gpu<1><<<g, BLOCK_SIZE>>>(nr_rays, nr_samples, a_se, a_v, a_r, a_s);
    else if (dim == 2)
        VolumeRenderingGPU::sum_over_rays_gpu<2><<<g, BLOCK_SIZE>>>(nr_rays, nr_samples, a_se, a_v, a_r, a_s);
    else if (dim == 3)
        VolumeRenderingGPU::sum_over_rays_gpu<3><<<g, BLOCK_SIZE>>>(nr_rays, nr_samples, a_se, a_v, a_r, a_s);
    else if (dim == 32)
        VolumeRenderingGPU::sum_over_rays_gpu<32><<<g, BLOCK_SIZE>>>(nr_rays, nr_samples, a_se, a_v, a_r, a_s);
    else
        return -1;
    return finish();
}

// VolumeRendering.cu:326-370
int ref_cumsum(const int* se, const float* v, int inverse, float* out, int nr_rays, int nr_samples) {
    VolumeRenderingGPU::cumsum_over_rays_gpu<<<grid_for(nr_rays), BLOCK_SIZE>>>(nr_rays, nr_samples, acc2(se, nr_rays, 2), acc2(v, nr_samples, 1),
                                                                                inverse != 0, acc2(out, nr_samples, 1));
    return finish();
}

// VolumeRendering.cu:372-414
int ref_median_depth(const int* se, const float* z, const float* w, float threshold, float* depth, int nr_rays, int nr_samples) {
    VolumeRenderingGPU::median_depth_over_rays_gpu<<<grid_for(nr_rays), BLOCK_SIZE>>>(
        nr_rays, nr_samples, threshold, acc2(se, nr_rays, 2), acc2(z, nr_samples, 1), acc2(w, nr_samples, 1), acc2(depth, nr_rays, 1));
    return finish();
}

// VolumeRendering.cu:416-464
int ref_compute_cdf(const int* se, const float* w, float* cdf, int nr_rays, int nr_samples) {
    VolumeRenderingGPU::compute_cdf_gpu<<<grid_for(nr_rays), BLOCK_SIZE>>>(nr_rays, nr_samples, acc2(se, nr_rays, 2), acc2(w, nr_samples, 1),
                                                                           acc2(cdf, nr_samples, 1));
    return finish();
}

// VolumeRendering.cu:466-548 (kernel launch only; the reference compacts the result afterwards)
int ref_importance_sample(const float* rays_o, const float* rays_d, const int* se, const float* z, const float* cdf, int nr_rays,
                          int uniform_nr_samples, int n_imp, uint64_t rng_state, uint64_t rng_inc, int jitter, int* imp_idx, float* imp_3d,
                          float* imp_dirs, float* imp_z, int* imp_se) {
    pcg32 rng;
    rng.state = rng_state;
    rng.inc = rng_inc;
    const int imp_max = nr_rays * n_imp;
    VolumeRenderingGPU::importance_sample_gpu<<<grid_for(nr_rays), BLOCK_SIZE>>>(
        nr_rays, acc2(rays_o, nr_rays, 3), acc2(rays_d, nr_rays, 3), uniform_nr_samples, imp_max, acc2(se, nr_rays, 2), acc2(z, uniform_nr_samples, 1),
        acc2(cdf, uniform_nr_samples, 1), n_imp, rng, jitter != 0, acc2(imp_idx, imp_max, 1), acc2(imp_3d, imp_max, 3), acc2(imp_dirs, imp_max, 3),
        acc2(imp_z, imp_max, 1), acc2(imp_se, nr_rays, 2));
    return finish();
}

// VolumeRendering.cu:550-669 (kernel launch only)
int ref_combine(int nr_rays, float min_dist, int values_dim, int n1, const int* se1, const int* idx1, const float* p1, const float* d1,
                const float* z1, const float* v1, int n2, const int* se2, const int* idx2, const float* p2, const float* d2, const float* z2,
                const float* v2, int n_comb, const int* out_start, int* c_idx, float* c_3d, float* c_dirs, float* c_z, float* c_val, int* c_se) {
    VolumeRenderingGPU::combine_ray_samples_packets_gpu<<<grid_for(nr_rays), BLOCK_SIZE>>>(
        nr_rays, min_dist, values_dim, n1, acc2(se1, nr_rays, 2), acc2(idx1, n1, 1), acc2(p1, n1, 3), acc2(d1, n1, 3), acc2(z1, n1, 1),
        acc2(v1, n1, values_dim), n2, acc2(se2, nr_rays, 2), acc2(idx2, n2, 1), acc2(p2, n2, 3), acc2(d2, n2, 3), acc2(z2, n2, 1),
        acc2(v2, n2, values_dim), n_comb, acc1(out_start, nr_rays), acc2(c_idx, n_comb, 1), acc2(c_3d, n_comb, 3), acc2(c_dirs, n_comb, 3),
        acc2(c_z, n_comb, 1), acc2(c_val, n_comb, values_dim), acc2(c_se, nr_rays, 2));
    return finish();
}

// VolumeRendering.cu:671-718
int ref_cumprod_bwd(const int* se, const float* gT, const float* gbg, const float* x, const float* T, const float* bgT, const float* cumsumLV,
                    float* dx, int nr_rays, int nr_samples) {
    VolumeRenderingGPU::cumprod_one_minus_alpha_to_transmittance_backward_gpu<<<grid_for(nr_rays), BLOCK_SIZE>>>(
        nr_rays, nr_samples, acc2(se, nr_rays, 2), acc2(gT, nr_samples, 1), acc2(gbg, nr_rays, 1), acc2(x, nr_samples, 1), acc2(T, nr_samples, 1),
        acc2(bgT, nr_rays, 1), acc2(cumsumLV, nr_samples, 1), acc2(dx, nr_samples, 1));
    return finish();
}

// VolumeRendering.cu:720-818
int ref_integrate_bwd(const int* se, const float* g, const float* v, const float* w, const float* result, float* dv, float* dw, int dim,
                      int nr_rays, int nr_samples) {
    if (dim == 1)
        VolumeRenderingGPU::integrate_with_weights_1d_backward_gpu<<<grid_for(nr_rays), BLOCK_SIZE>>>(
            nr_rays, nr_samples, acc2(se, nr_rays, 2), acc2(g, nr_rays, 1), acc2(v, nr_samples, 1), acc2(w, nr_samples, 1), acc2(result, nr_rays, 1),
            acc2(dv, nr_samples, 1), acc2(dw, nr_samples, 1));
    else if (dim == 3)
        VolumeRenderingGPU::integrate_with_weights_3d_backward_gpu<<<grid_for(nr_rays), BLOCK_SIZE>>>(
            nr_rays, nr_samples, acc2(se, nr_rays, 2), acc2(g, nr_rays, 3), acc2(v, nr_samples, 3), acc2(w, nr_samples, 1), acc2(result, nr_rays, 3),
            acc2(dv, nr_samples, 3), acc2(dw, nr_samples, 1));
    else
        return -1;
    return finish();
}

// VolumeRendering.cu:820-899
int ref_sum_bwd(const int* se, const float* g_ray, const float* g_sample, const float* v, float* dv, int dim, int nr_rays, int nr_samples) {
    auto a_se = acc2(se, nr_rays, 2);
    auto a_gr = acc2(g_ray, nr_rays, dim);
    auto a_gs = acc2(g_sample, nr_samples, dim);
    auto a_v = acc2(v, nr_samples, dim);
    auto a_dv = acc2(dv, nr_samples, dim);
    const dim3 g = grid_for(nr_rays);
    if (dim == 1)
        VolumeRenderingGPU::sum_over_rays_backward_gpu<1><<<g, BLOCK_SIZE>>>(nr_rays, nr_samples, a_se, a_gr, a_gs, a_v, a_dv);
    else if (dim == 2)
        VolumeRenderingGPU::sum_over_rays_backward_gpu<2><<<g, BLOCK_SIZE>>>(nr_rays, nr_samples, a_se, a_gr, a_gs, a_v, a_dv);
    else if (dim == 3)
        VolumeRenderingGPU::sum_over_rays_backward_gpu<3><<<g, BLOCK_SIZE>>>(nr_rays, nr_samples, a_se, a_gr, a_gs, a_v, a_dv);
    else
        return -1;
    return finish();
}

}  // extern "C"
