// ORACLE (test infrastructure only — never linked into or loaded by the product).
//
// A thin C harness around the REFERENCE's own ray-sampling and occupancy-grid kernels: it #includes
//   /root/reference/kernels/volsurfs/RaySamplerGPU.cuh        (compute_samples_bg / _fg / _fg_in_grid_occupied_regions)
//   /root/reference/kernels/volsurfs/OccupancyGridGPU.cuh     (get_rays_t_near_t_far, check_occupancy)
//   /root/reference/kernels/volsurfs/occ_grid_helpers.h       (pos_to_lin_idx, distance_to_next_voxel; pulled in by the two above)
//   /root/reference/kernels/volsurfs/RaySamplesPackedGPU.cuh  (update_dt_gpu, compact_to_valid_samples_gpu)
// unmodified, where they lie, and launches them with the launch shape of src/RaySampler.cu / src/OccupancyGrid.cu
// (blocks = div_round_up(n, 256), 256 threads, legacy stream, device synchronise).
// Those headers use Eigen::Vector3f only as a 3-float parameter type (x(), y(), z()); Eigen is not installed in this image and cannot
// be downloaded, so a 3-float stand-in with those three accessors is declared below — it is the harness's own code, not a copy.
// Built by oracle/build.py into oracle/_ref/libsampler_ref.so (git-ignored; travels to the GPU box with the snapshot), with nvcc's
// default floating-point flags (FMA contraction on), as the reference's own CMake build does.
#include <cstdint>

namespace Eigen {
struct Vector3f {
    float v[3];
    __host__ __device__ Vector3f() : v{0.f, 0.f, 0.f} {}
    __host__ __device__ Vector3f(float a, float b, float c) : v{a, b, c} {}
    __host__ __device__ float x() const { return v[0]; }
    __host__ __device__ float y() const { return v[1]; }
    __host__ __device__ float z() const { return v[2]; }
    // grid_extent / nr_voxels_per_dim in update_grid_occupancy_with_sdf_values: Eigen's vector / scalar is a per-component division
    __host__ __device__ Vector3f operator/(float s) const { return Vector3f(v[0] / s, v[1] / s, v[2] / s); }
};
}  // namespace Eigen

#include "volsurfs/RaySamplerGPU.cuh"
#undef BLOCK_SIZE
#include "volsurfs/OccupancyGridGPU.cuh"
#undef BLOCK_SIZE
#include "volsurfs/RaySamplesPackedGPU.cuh"  // update_dt_gpu (:14-88), compact_to_valid_samples_gpu (:172-257)

namespace {

template <typename T, int N>
using Acc = torch::PackedTensorAccessor32<T, N, torch::RestrictPtrTraits>;

template <typename T>
Acc<T, 1> acc1(const T* p, int64_t n) {
    const int64_t sizes[1] = {n}, strides[1] = {1};
    return Acc<T, 1>(const_cast<T*>(p), sizes, strides);
}
template <typename T>
Acc<T, 2> acc2(const T* p, int64_t rows, int64_t cols) {
    const int64_t sizes[2] = {rows, cols}, strides[2] = {cols, 1};
    return Acc<T, 2>(const_cast<T*>(p), sizes, strides);
}
template <typename T>
Acc<T, 3> acc3(const T* p, int64_t a, int64_t b, int64_t c) {
    const int64_t sizes[3] = {a, b, c}, strides[3] = {b * c, c, 1};
    return Acc<T, 3>(const_cast<T*>(p), sizes, strides);
}

inline dim3 grid_for(int n) { return dim3((unsigned)((n + 255) / 256), 1, 1); }

inline int finish() {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    return (int)e;
}

inline pcg32 make_rng(uint64_t state, uint64_t inc) {
    pcg32 rng;
    rng.state = state;
    rng.inc = inc;
    return rng;
}

}  // namespace

extern "C" {

int ref_sampler_abi_version() { return 2; }

// RaySamplesPacked.cu:188-273 (compact_to_valid_samples): the kernel launch; `out_indices_start` is the exclusive prefix sum the reference
// forms with torch (`nr_samples_per_ray.cumsum(0).to(int32)` shifted by one, :218-221) — the caller passes it.  Outputs preset by the caller
// as the RaySamplesPacked constructor does (-1 everywhere, RaySamplesPacked.cu:13-48): rays without samples keep (-1,-1).
int ref_compact_to_valid_samples(int nr_rays, int max_in, int max_out, int values_dim, const int* sidx, const float* s3d, const float* sdirs,
                                 const float* sz, const float* sdt, const float* sval, const int* se, const int* out_start, int* o_sidx,
                                 float* o_s3d, float* o_sdirs, float* o_sz, float* o_sdt, float* o_sval, int* o_se) {
    RaySamplesPackedGPU::compact_to_valid_samples_gpu<<<grid_for(nr_rays), 256>>>(
        nr_rays, max_in, max_out, values_dim, acc2(sidx, max_in, 1), acc2(s3d, max_in, 3), acc2(sdirs, max_in, 3), acc2(sz, max_in, 1),
        acc2(sdt, max_in, 1), acc2(sval, max_in, values_dim), acc2(se, nr_rays, 2), acc1(out_start, nr_rays), acc2(o_sidx, max_out, 1),
        acc2(o_s3d, max_out, 3), acc2(o_sdirs, max_out, 3), acc2(o_sz, max_out, 1), acc2(o_sdt, max_out, 1), acc2(o_sval, max_out, values_dim),
        acc2(o_se, nr_rays, 2));
    return finish();
}

// RaySamplesPacked.cu:396-461 (update_dt): writes samples_dt in place
int ref_update_dt(int nr_rays, int nr_samples, int is_background, const float* ray_max_dt, const float* ray_exit, const float* sz, const int* se,
                  float* sdt) {
    RaySamplesPackedGPU::update_dt_gpu<<<grid_for(nr_rays), 256>>>(nr_rays, nr_samples, is_background != 0, acc2(ray_max_dt, nr_rays, 1),
                                                                  acc2(ray_exit, nr_rays, 1), acc2(sz, nr_samples, 1), acc2(se, nr_rays, 2),
                                                                  acc2(sdt, nr_samples, 1));
    return finish();
}

// RaySampler.cu:72-157
int ref_samples_bg(const float* rays_o, const float* rays_d, const float* t_start, float t_far, int nr_samples_per_ray, uint64_t rng_state,
                   uint64_t rng_inc, int jitter, float* ray_max_dt, float* s3d, float* sdirs, float* sz, float* sdt, int* se, int nr_rays) {
    RaySamplerGPU::compute_samples_bg_gpu<<<grid_for(nr_rays), 256>>>(
        nr_rays, nr_samples_per_ray, acc2(rays_o, nr_rays, 3), acc2(rays_d, nr_rays, 3), acc2(t_start, nr_rays, 1), t_far,
        make_rng(rng_state, rng_inc), jitter != 0, acc2(ray_max_dt, nr_rays, 1), acc3(s3d, nr_rays, nr_samples_per_ray, 3),
        acc3(sdirs, nr_rays, nr_samples_per_ray, 3), acc2(sz, nr_rays, nr_samples_per_ray), acc2(sdt, nr_rays, nr_samples_per_ray),
        acc2(se, nr_rays, 2));
    return finish();
}

// RaySampler.cu:159-245 (before compact_to_valid_samples)
int ref_samples_fg(const float* rays_o, const float* rays_d, const float* t_entry, const float* t_exit, float min_dist, int min_nr, int max_nr,
                   uint64_t rng_state, uint64_t rng_inc, int jitter, float* ray_max_dt, int* sidx, float* s3d, float* sdirs, float* sz,
                   float* sdt, int* se, int nr_rays) {
    const int64_t cap = (int64_t)nr_rays * max_nr;
    RaySamplerGPU::compute_samples_fg_gpu<<<grid_for(nr_rays), 256>>>(
        nr_rays, acc2(rays_o, nr_rays, 3), acc2(rays_d, nr_rays, 3), acc2(t_entry, nr_rays, 1), acc2(t_exit, nr_rays, 1), min_dist, min_nr,
        max_nr, (int)cap, make_rng(rng_state, rng_inc), jitter != 0, acc2(ray_max_dt, nr_rays, 1), acc2(sidx, cap, 1), acc2(s3d, cap, 3),
        acc2(sdirs, cap, 3), acc2(sz, cap, 1), acc2(sdt, cap, 1), acc2(se, nr_rays, 2));
    return finish();
}

// RaySampler.cu:247-340 (before compact_to_valid_samples)
int ref_samples_fg_occupied(const float* rays_o, const float* rays_d, const float* t_entry, const float* t_exit, float min_dist, int min_nr,
                            int max_nr, uint64_t rng_state, uint64_t rng_inc, int jitter, int nr_voxels_per_dim, const float* extent,
                            const bool* occupancy, const bool* roi, float* ray_max_dt, int* sidx, float* s3d, float* sdirs, float* sz,
                            float* sdt, int* se, int nr_rays) {
    const int64_t cap = (int64_t)nr_rays * max_nr;
    const int64_t nv = (int64_t)nr_voxels_per_dim * nr_voxels_per_dim * nr_voxels_per_dim;
    RaySamplerGPU::compute_samples_fg_in_grid_occupied_regions_gpu<<<grid_for(nr_rays), 256>>>(
        nr_rays, acc2(rays_o, nr_rays, 3), acc2(rays_d, nr_rays, 3), acc2(t_entry, nr_rays, 1), acc2(t_exit, nr_rays, 1), min_dist, min_nr,
        max_nr, (int)cap, make_rng(rng_state, rng_inc), jitter != 0, nr_voxels_per_dim, Eigen::Vector3f(extent[0], extent[1], extent[2]),
        acc1(occupancy, nv), acc1(roi, nv), acc2(ray_max_dt, nr_rays, 1), acc2(sidx, cap, 1), acc2(s3d, cap, 3), acc2(sdirs, cap, 3),
        acc2(sz, cap, 1), acc2(sdt, cap, 1), acc2(se, nr_rays, 2));
    return finish();
}

// RaySampler.cu:336-427: contract_samples / uncontract_samples (before the closing update_dt)
int ref_contract_samples(const float* ray_o, const int* se, const float* s3d, const float* sz, float* out_3d, float* out_z, int nr_rays,
                         int nr_samples, int uncontract) {
    if (uncontract)
        RaySamplerGPU::uncontract_samples_gpu<<<grid_for(nr_rays), 256>>>(nr_rays, acc2(ray_o, nr_rays, 3), acc2(se, nr_rays, 2),
                                                                          acc2(s3d, nr_samples, 3), acc2(sz, nr_samples, 1),
                                                                          acc2(out_3d, nr_samples, 3), acc2(out_z, nr_samples, 1));
    else
        RaySamplerGPU::contract_samples_gpu<<<grid_for(nr_rays), 256>>>(nr_rays, acc2(ray_o, nr_rays, 3), acc2(se, nr_rays, 2),
                                                                        acc2(s3d, nr_samples, 3), acc2(sz, nr_samples, 1),
                                                                        acc2(out_3d, nr_samples, 3), acc2(out_z, nr_samples, 1));
    return finish();
}

// OccupancyGrid.cu:206-347: get_grid_lower_left_voxels_vertices (centre = 0) / get_grid_samples, get_random_grid_samples[_in_roi] (centre = 1)
int ref_grid_points(const int* point_indices, int nr_voxels_per_dim, const float* extent, int centre, uint64_t rng_state, uint64_t rng_inc,
                    int jitter, float* out, int nr_points) {
    const Eigen::Vector3f e(extent[0], extent[1], extent[2]);
    if (centre)
        OccupancyGridGPU::get_grid_samples_gpu<<<grid_for(nr_points), 256>>>(nr_points, nr_voxels_per_dim, e, acc1(point_indices, nr_points),
                                                                            make_rng(rng_state, rng_inc), jitter != 0, acc2(out, nr_points, 3));
    else
        OccupancyGridGPU::get_grid_lower_left_voxels_vertices_gpu<<<grid_for(nr_points), 256>>>(nr_points, nr_voxels_per_dim, e,
                                                                                               acc1(point_indices, nr_points), acc2(out, nr_points, 3));
    return finish();
}

// OccupancyGrid.cu:446-474: update_grid_values
int ref_update_grid_values(const int* point_indices, const float* values, float decay, int nr_voxels_per_dim, float* grid_values, int nr_points) {
    const int64_t nv = (int64_t)nr_voxels_per_dim * nr_voxels_per_dim * nr_voxels_per_dim;
    OccupancyGridGPU::update_grid_values_gpu<<<grid_for(nr_points), 256>>>(nr_points, acc2(values, nr_points, 1), nr_voxels_per_dim,
                                                                          acc1(point_indices, nr_points), decay, acc1(grid_values, nv));
    return finish();
}

// OccupancyGrid.cu:476-503: update_grid_occupancy_with_density_values
int ref_update_grid_occupancy_density(const int* point_indices, int nr_voxels_per_dim, const float* extent, float occupancy_thresh,
                                      int check_neighbours, const float* grid_values, bool* grid_occupancy, int nr_points) {
    const int64_t nv = (int64_t)nr_voxels_per_dim * nr_voxels_per_dim * nr_voxels_per_dim;
    OccupancyGridGPU::update_grid_occupancy_with_density_values_gpu<<<grid_for(nr_points), 256>>>(
        nr_points, nr_voxels_per_dim, Eigen::Vector3f(extent[0], extent[1], extent[2]), acc1(point_indices, nr_points), occupancy_thresh,
        check_neighbours != 0, acc1(grid_values, nv), acc1(grid_occupancy, nv));
    return finish();
}

// OccupancyGrid.cu:505-533: update_grid_occupancy_with_sdf_values
int ref_update_grid_occupancy_sdf(const int* point_indices, int nr_voxels_per_dim, const float* extent, const float* logistic_beta,
                                  float occupancy_thresh, int check_neighbours, const float* grid_values, bool* grid_occupancy, int nr_points) {
    const int64_t nv = (int64_t)nr_voxels_per_dim * nr_voxels_per_dim * nr_voxels_per_dim;
    OccupancyGridGPU::update_grid_occupancy_with_sdf_values_gpu<<<grid_for(nr_points), 256>>>(
        nr_points, Eigen::Vector3f(extent[0], extent[1], extent[2]), nr_voxels_per_dim, acc1(point_indices, nr_points),
        acc2(logistic_beta, nr_points, 1), occupancy_thresh, check_neighbours != 0, acc1(grid_values, nv), acc1(grid_occupancy, nv));
    return finish();
}

// OccupancyGrid.cu:536-573: get_first_rays_sample_start_of_grid_occupied_regions (outputs preset by the caller as the RaySamplesPacked
// constructor does: the kernel leaves the rows of rays without a hit untouched)
int ref_first_sample_start(const float* rays_o, const float* rays_d, const float* t_entry, const float* t_exit, int nr_voxels_per_dim,
                           const float* extent, const bool* occupancy, const bool* roi, float* samples_3d, float* samples_dirs, float* samples_z,
                           float* samples_dt, int* se, int nr_rays) {
    const int64_t nv = (int64_t)nr_voxels_per_dim * nr_voxels_per_dim * nr_voxels_per_dim;
    OccupancyGridGPU::get_first_rays_sample_start_of_grid_occupied_regions_gpu<<<grid_for(nr_rays), 256>>>(
        nr_rays, nr_voxels_per_dim, Eigen::Vector3f(extent[0], extent[1], extent[2]), acc2(rays_o, nr_rays, 3), acc2(rays_d, nr_rays, 3),
        acc2(t_entry, nr_rays, 1), acc2(t_exit, nr_rays, 1), acc1(occupancy, nv), acc1(roi, nv), acc2(samples_3d, nr_rays, 3),
        acc2(samples_dirs, nr_rays, 3), acc2(samples_z, nr_rays, 1), acc2(samples_dt, nr_rays, 1), acc2(se, nr_rays, 2));
    return finish();
}

// OccupancyGrid.cu:575-607: advance_ray_sample_to_next_occupied_voxel (the reference passes the input tensor as the output too)
int ref_advance_to_next_occupied(const float* samples_dirs, const float* samples_3d, int nr_voxels_per_dim, const float* extent,
                                 const bool* occupancy, const bool* roi, float* new_samples_3d, bool* is_within_bounds, int nr_points) {
    const int64_t nv = (int64_t)nr_voxels_per_dim * nr_voxels_per_dim * nr_voxels_per_dim;
    OccupancyGridGPU::advance_ray_sample_to_next_occupied_voxel_gpu<<<grid_for(nr_points), 256>>>(
        nr_points, nr_voxels_per_dim, Eigen::Vector3f(extent[0], extent[1], extent[2]), acc2(samples_dirs, nr_points, 3),
        acc2(samples_3d, nr_points, 3), acc1(occupancy, nv), acc1(roi, nv), acc2(new_samples_3d, nr_points, 3), acc2(is_within_bounds, nr_points, 1));
    return finish();
}

// OccupancyGrid.cu: get_rays_t_near_t_far
int ref_rays_t_near_t_far(const float* rays_o, const float* rays_d, const float* t_entry, const float* t_exit, int nr_voxels_per_dim,
                          const float* extent, const bool* occupancy, const bool* roi, float* t_near, float* t_far, int nr_rays) {
    const int64_t nv = (int64_t)nr_voxels_per_dim * nr_voxels_per_dim * nr_voxels_per_dim;
    OccupancyGridGPU::get_rays_t_near_t_far_gpu<<<grid_for(nr_rays), 256>>>(
        nr_rays, acc2(rays_o, nr_rays, 3), acc2(rays_d, nr_rays, 3), acc2(t_entry, nr_rays, 1), acc2(t_exit, nr_rays, 1), nr_voxels_per_dim,
        Eigen::Vector3f(extent[0], extent[1], extent[2]), acc1(occupancy, nv), acc1(roi, nv), acc2(t_near, nr_rays, 1), acc2(t_far, nr_rays, 1));
    return finish();
}

// OccupancyGrid.cu: check_occupancy
int ref_check_occupancy(const float* points, int nr_voxels_per_dim, const float* extent, const float* values, const bool* occupancy,
                        const bool* roi, bool* out_occ, float* out_val, int nr_points) {
    const int64_t nv = (int64_t)nr_voxels_per_dim * nr_voxels_per_dim * nr_voxels_per_dim;
    OccupancyGridGPU::check_occupancy_gpu<<<grid_for(nr_points), 256>>>(
        nr_points, nr_voxels_per_dim, Eigen::Vector3f(extent[0], extent[1], extent[2]), acc1(values, nv), acc1(occupancy, nv), acc1(roi, nv),
        acc2(points, nr_points, 3), acc2(out_occ, nr_points, 1), acc2(out_val, nr_points, 1));
    return finish();
}

}  // extern "C"
