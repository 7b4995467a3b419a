// ORACLE (test infrastructure only — never linked into or loaded by the product).
//
// A thin C harness around the REFERENCE's own permutohedral-encoding kernels: it #includes
//   /root/reference/submodules/permutohedral_encoding/kernels/permutohedral_encoding/EncodingGPU.cuh   (unmodified, where it lies)
// and launches forward_gpu / backward_gpu / backward_gpu_only_pos for pos_dim 3, 2 features per level with the launch shape of the
// reference host code (src/Encoding.cu:84-101,166-197: grid (ceil(N/128), levels), 128 threads, then a device synchronise).
// Built by oracle/build.py:build_ref_permuto into oracle/_ref/libpermuto_ref.so (git-ignored; travels to the GPU box).
// tests/test_gpu_permuto.py compares the product's kernels and the numpy restatement (oracle/permuto.py) with these.
//
// No reference source is copied here: this file only builds accessors over raw device pointers and forwards them.
#include <cstdint>

#include "permutohedral_encoding/EncodingGPU.cuh"

namespace {

template <typename T, int N>
using Acc = torch::PackedTensorAccessor32<T, N, torch::RestrictPtrTraits>;

Acc<float, 1> acc1(const float* p, int64_t n) {
    const int64_t sizes[1] = {n};
    const int64_t strides[1] = {1};
    return Acc<float, 1>(const_cast<float*>(p), sizes, strides);
}
Acc<float, 2> acc2(const float* p, int64_t a, int64_t b) {
    const int64_t sizes[2] = {a, b};
    const int64_t strides[2] = {b, 1};
    return Acc<float, 2>(const_cast<float*>(p), sizes, strides);
}
Acc<float, 3> acc3(const float* p, int64_t a, int64_t b, int64_t c) {
    const int64_t sizes[3] = {a, b, c};
    const int64_t strides[3] = {b * c, c, 1};
    return Acc<float, 3>(const_cast<float*>(p), sizes, strides);
}

inline int finish() {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    return (int)e;
}

}  // namespace

extern "C" {

int ref_permuto_abi_version() { return 1; }

// Encoding.cu:55-113.  sliced: [levels + extra, 2, n] (torch::empty in the reference — every element is written).
int ref_permuto_forward(const float* positions, const float* lattice_values, const float* scale_factor, const float* random_shift,
                        const float* anneal_window, float* sliced, int n, int capacity, int levels, int concat_points, float points_scaling) {
    const int extra = concat_points ? 2 : 0;  // ceil(3 / 2)
    const dim3 blocks((unsigned)((n + BLOCK_SIZE - 1) / BLOCK_SIZE), (unsigned)(levels + extra), 1);
    forward_gpu<3, 2><<<blocks, BLOCK_SIZE>>>(n, capacity, levels, extra, acc2(positions, n, 3), acc3(lattice_values, levels, capacity, 2),
                                              acc2(scale_factor, levels, 3), acc2(random_shift, levels, 3), acc1(anneal_window, levels),
                                              acc3(sliced, levels + extra, 2, n), concat_points != 0, points_scaling, true, true);
    return finish();
}

// Encoding.cu:116-217.  lattice_grad: [levels, 2, capacity] zero-filled by the caller (the reference permutes it to [levels, capacity, 2]
// afterwards); positions_grad: [3, n] zero-filled by the caller (transposed afterwards), or NULL.
int ref_permuto_backward(const float* positions, const float* lattice_values, const float* scale_factor, const float* random_shift,
                         const float* anneal_window, const float* grad_sliced, float* lattice_grad, float* positions_grad, int n,
                         int capacity, int levels, int concat_points) {
    const int extra = concat_points ? 2 : 0;
    const dim3 blocks((unsigned)((n + BLOCK_SIZE_BACK - 1) / BLOCK_SIZE_BACK), (unsigned)levels, 1);
    const bool want_pos = positions_grad != nullptr;
    auto a_pos_grad = want_pos ? acc2(positions_grad, 3, n) : acc2(lattice_grad, 1, 1);
    backward_gpu<3, 2><<<blocks, BLOCK_SIZE_BACK>>>(n, capacity, acc3(lattice_values, levels, capacity, 2), acc2(positions, n, 3),
                                                   acc2(scale_factor, levels, 3), acc2(random_shift, levels, 3), acc1(anneal_window, levels),
                                                   acc3(grad_sliced, levels + extra, 2, n), acc3(lattice_grad, levels, 2, capacity),
                                                   a_pos_grad, concat_points != 0, true, want_pos);
    if (want_pos)
        backward_gpu_only_pos<3, 2><<<blocks, BLOCK_SIZE_BACK>>>(n, capacity, acc3(lattice_values, levels, capacity, 2), acc2(positions, n, 3),
                                                                acc2(scale_factor, levels, 3), acc2(random_shift, levels, 3),
                                                                acc1(anneal_window, levels), acc3(grad_sliced, levels + extra, 2, n),
                                                                acc3(lattice_grad, levels, 2, capacity), a_pos_grad, concat_points != 0,
                                                                true, true);
    return finish();
}

}  // extern "C"
