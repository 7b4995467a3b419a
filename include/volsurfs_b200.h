/* volsurfs_b200 — C ABI of the B200-native per-ray rendering hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every pointer is a DEVICE pointer on the
 * current CUDA device unless stated otherwise; `stream` is a cudaStream_t passed as void* (NULL = legacy default
 * stream).  Calls only enqueue work: they never synchronise the device, never allocate device memory (callers pass
 * outputs and scratch) and never abort the process.
 *
 * Return value: 0 = ok, >0 = cudaError_t of the launch, <0 = argument error
 *   VS_ERR_INVALID_ARG (-1), VS_ERR_UNSUPPORTED (-2), VS_ERR_ALLOC (-3).
 * (The reference instead CHECK()-aborts on bad arguments and prints-and-ignores CUDA errors,
 *  src/VolumeRendering.cu:35,66-71.)
 *
 * Tensor conventions (identical to the reference's pybind surface, src/PyBridge.cxx:70-129):
 *   ray_start_end_idx  int32 [n_rays,2]  (start,end) into the packed sample arrays; empty rays are (-1,-1)
 *   per-sample arrays  float32 [n_samples,d] row-major contiguous
 *   per-ray arrays     float32 [n_rays,d]
 */
#ifndef VOLSURFS_B200_H
#define VOLSURFS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VS_OK 0
#define VS_ERR_INVALID_ARG (-1)
#define VS_ERR_UNSUPPORTED (-2)
#define VS_ERR_ALLOC (-3)

/* ---- library info ------------------------------------------------------------------------------------------------ */
int vs_abi_version(void);
/* static string for a return code of this ABI (host pointer, never NULL) */
const char* vs_error_string(int code);
/* number of CUDA kernels this library has enqueued since it was loaded (process-wide, monotonic) */
long long vs_launch_count(void);

/* ---- packed VolumeRendering operators (op-by-op, reference semantics incl. quirks) -------------------------------- */

/* replaces VolumeRendering::cumprod_one_minus_alpha_to_transmittance, src/VolumeRendering.cu:30-78
 * (kernel kernels/volsurfs/VolumeRenderingGPU.cuh:28-78).  T_i = prod_{j<i} x_j ; bgT = T_{s-1} (last x not applied);
 * bgT = 1 for empty rays.  T must be zero-filled by the caller if the layout has unowned slots. */
int vs_cumprod_fwd(const int32_t* ray_start_end_idx, const float* x, float* T, float* bgT, int64_t n_rays, int64_t n_samples,
                   void* stream);

/* replaces VolumeRendering::cumprod_one_minus_alpha_to_transmittance_backward, src/VolumeRendering.cu:671-718
 * (kernel VolumeRenderingGPU.cuh:896-943): dx_i = cumsumLV_{i+1}/max(x_i,1e-6) + g_bgT*bgT/max(x_i,1e-6), dx_{s-1}=0 */
int vs_cumprod_bwd(const int32_t* ray_start_end_idx, const float* grad_bgT, const float* x, const float* bgT, const float* cumsumLV,
                   float* dx, int64_t n_rays, int64_t n_samples, void* stream);

/* the python half (volume_rendering_funcs.py:105-179: LV = gT*T, reverse cumsum) fused with the kernel above */
int vs_cumprod_bwd_fused(const int32_t* ray_start_end_idx, const float* grad_T, const float* grad_bgT, const float* x, const float* T,
                         const float* bgT, float* dx, int64_t n_rays, int64_t n_samples, void* stream);

/* replaces VolumeRendering::cumsum_over_rays, src/VolumeRendering.cu:326-370 (kernel :305-361) */
int vs_cumsum(const int32_t* ray_start_end_idx, const float* values, float* out, int inverse, int64_t n_rays, int64_t n_samples,
              void* stream);

/* replaces VolumeRendering::integrate_with_weights_{1d,3d}, src/VolumeRendering.cu:80-176 (kernels :80-177); dim in {1,3} */
int vs_integrate_fwd(const int32_t* ray_start_end_idx, const float* values, const float* weights, float* out, int dim, int64_t n_rays,
                     int64_t n_samples, void* stream);

/* replaces VolumeRendering::integrate_with_weights_{1d,3d}_backward, src/VolumeRendering.cu:720-818 (kernels :945-1033).
 * ref_bug != 0 reproduces the reference's 3-D kernel reading values[.][1] for the z channel (:1021). */
int vs_integrate_bwd(const int32_t* ray_start_end_idx, const float* grad_out, const float* values, const float* weights,
                     float* d_values, float* d_weights, int dim, int64_t n_rays, int64_t n_samples, int ref_bug, void* stream);

/* replaces VolumeRendering::sum_over_rays, src/VolumeRendering.cu:231-324 (kernel :245-303); dim in {1,2,3,32};
 * sum_sample may be NULL */
int vs_sum_fwd(const int32_t* ray_start_end_idx, const float* values, float* sum_ray, float* sum_sample, int dim, int64_t n_rays,
               int64_t n_samples, void* stream);

/* replaces VolumeRendering::sum_over_rays_backward, src/VolumeRendering.cu:820-899 (kernel :1035-1079); dim in {1,2,3} */
int vs_sum_bwd(const int32_t* ray_start_end_idx, const float* grad_sum_ray, const float* grad_sum_sample, float* d_values, int dim,
               int64_t n_rays, int64_t n_samples, void* stream);

/* replaces RaySamplesPacked::update_dt, src/RaySamplesPacked.cu:396-461 (kernel RaySamplesPackedGPU.cuh:14-88) */
int vs_update_dt(const int32_t* ray_start_end_idx, const float* samples_z, const float* ray_exit, const float* ray_max_dt,
                 float* samples_dt, int is_background, int64_t n_rays, int64_t n_samples, void* stream);

/* "next" operators of the same container.  Outputs must be zero-filled by the caller (the reference
 * returns torch::zeros and skips rays/samples it does not own). */
/* replaces VolumeRendering::sdf2alpha, src/VolumeRendering.cu:178-229 (kernel VolumeRenderingGPU.cuh:185-243) */
int vs_sdf2alpha(const int32_t* ray_start_end_idx, const float* samples_dt, const float* samples_sdf, const float* logistic_beta,
                 float* alpha, int64_t n_rays, int64_t n_samples, void* stream);
/* replaces VolumeRendering::median_depth_over_rays, src/VolumeRendering.cu:372-416 (kernel :364-409); ref_bug != 0 keeps the
 * reference's fallback index samples_z[nr_samples-1] (:407) */
int vs_median_depth(const int32_t* ray_start_end_idx, const float* samples_z, const float* weights, float threshold, float* out,
                    int64_t n_rays, int64_t n_samples, int ref_bug, void* stream);
/* replaces VolumeRendering::compute_cdf, src/VolumeRendering.cu:418-465 (kernel :412-471) */
int vs_compute_cdf(const int32_t* ray_start_end_idx, const float* weights, float* cdf, int64_t n_rays, int64_t n_samples, void* stream);

/* ---- fused compositing (one launch per direction) ------------------------------------------------------------------
 * Replaces the chain cumprod -> alpha*T -> sum_over_rays -> integrate_3d -> integrate_1d of nerf.py:308-334 and equals
 * the dense K-layer torch path volsurfs_py/methods/volsurfs.py:601-640,708.  bgT is the FULL product of (1-alpha).
 * out_w / out_T (per-sample weights and transmittance) may be NULL.  d_z may be NULL.
 * mode: 0 auto (tile kernels up to a mean of 16 samples per ray, cp.async ring kernels beyond, scan kernels for the forward pass above
 * 256); 1 tile kernels (TMA staging), 2 scan kernels (one sample per lane), 3 / 5 / 6 / 7 coarsened scan kernels (auto / 8 / 16 / 32 lanes
 * per ray), 4 tile kernels with LDG staging, 8 ring kernels — 1..8 exist for A/B measurements. */
int vs_composite_fwd(const int32_t* ray_start_end_idx, const float* alpha, const float* rgb, const float* z, float* out_rgb,
                     float* out_depth, float* out_acc, float* out_bgT, float* out_w, float* out_T, int64_t n_rays, int64_t n_samples,
                     int mode, void* stream);
int vs_composite_bwd(const int32_t* ray_start_end_idx, const float* alpha, const float* rgb, const float* z, const float* g_rgb,
                     const float* g_depth, const float* g_acc, const float* g_bgT, float* d_alpha, float* d_rgb, float* d_z,
                     int64_t n_rays, int64_t n_samples, int mode, void* stream);

/* ---- packing ------------------------------------------------------------------------------------------------------- */
/* bytes of device scratch needed by the packing entry points for n_rays rays */
int64_t vs_pack_scratch_bytes(int64_t n_rays);

/* sum of (end-start) over all rays into *total_dev (device int64); replaces RaySamplesPacked::get_total_nr_samples,
 * src/RaySamplesPacked.cu:170-175 (which syncs through .item()) */
int vs_count_total(const int32_t* ray_start_end_idx, int64_t n_rays, int64_t* total_dev, void* stream);

/* replaces RaySamplesPacked::compact_to_valid_samples, src/RaySamplesPacked.cu:188-273
 * (kernel kernels/volsurfs/RaySamplesPackedGPU.cuh:172-257), in two phases so the caller can size the outputs:
 *   1. vs_compact_offsets: exclusive scan of the per-ray counts into `scratch`, total into *total_dev
 *   2. vs_compact_gather: copies every ray's segment to its offset; empty rays get (-1,-1) */
int vs_compact_offsets(const int32_t* se_in, int64_t n_rays, int64_t* total_dev, void* scratch, void* stream);
int vs_compact_gather(const int32_t* se_in, const void* scratch, const int32_t* samples_idx_in, const float* samples_3d_in,
                      const float* samples_dirs_in, const float* samples_z_in, const float* samples_dt_in, const float* values_in,
                      int values_dim, int32_t* se_out, int32_t* samples_idx_out, float* samples_3d_out, float* samples_dirs_out,
                      float* samples_z_out, float* samples_dt_out, float* values_out, int64_t n_rays, int64_t n_samples_out,
                      void* stream);

/* K-layer hit bookkeeping of volsurfs_py/methods/volsurfs.py:476-518,601-603 written straight into packed form.
 * depth/tri/bary are layer-major [K,n_rays] (mesh 0 = innermost); a layer is hit iff depth <= t_far
 * (raytracelib/raytracer.py:100).  Packed order per ray: descending mesh index (outer -> inner). */
int vs_pack_hits_offsets(const float* depth, int K, float t_far, int64_t n_rays, int64_t* total_dev, void* scratch, void* stream);
int vs_pack_hits_scatter(const float* rays_o, const float* rays_d, const float* depth, const int32_t* tri, const float* bary_u,
                         const float* bary_v, const void* scratch, int K, float t_far, int32_t* se_out, int32_t* samples_idx_out,
                         float* samples_3d_out, float* samples_dirs_out, float* samples_z_out, int32_t* layer_out, int32_t* tri_out,
                         float* uv_out, int64_t n_rays, void* stream);

/* ---- K-layer shell intersection ------------------------------------------------------------------------------------
 * Replaces, for this path, raytracelib's create_raytracer + RayTracer.trace called once per mesh
 * (submodules/raytracelib/src/raytracer.cu:23-69, src/bvh.cu:186-263,420-469, raytracelib/raytracer.py:7-113;
 * call site volsurfs_py/methods/volsurfs.py:476-485). */

/* Build the per-layer BVHs on the host from HOST arrays (verts[k]: float32 [n_verts[k],3], faces[k]: int32 [n_faces[k],3];
 * mesh 0 = innermost) and upload them to the current device.  The handle owns device memory.  Synchronous. */
int vs_shells_build(int K, const float* const* verts, const int64_t* n_verts, const int32_t* const* faces, const int64_t* n_faces,
                    void** handle_out);
int vs_shells_free(void* handle);
int vs_shells_num_layers(const void* handle);
int vs_shells_info(const void* handle, int layer, int64_t* n_nodes, int64_t* n_tris);
/* 1 if a traversal stack ever overflowed since the build (results unreliable); synchronises the device */
int vs_shells_overflowed(const void* handle);

/* Nearest hit (t > 0, first along the ray) with layers [layer_first, layer_first+layer_count) in one launch.
 * Outputs layer-major [layer_count, n_rays]: depth (1e6 on a miss, include/raytracing/common.h:21), original face index
 * (-1 on a miss), barycentric u and v of triangle.cuh:53-55 (0 on a miss).
 * layer_count <= 64.  The launch uses a small set of per-layer work counters owned by the handle (eight sets, handed out in turn and
 * cleared by a memset node on `stream` in front of the kernel): launches on different streams may overlap, and the call can be captured
 * into a CUDA graph (a graph keeps the set it was captured with). */
int vs_shells_trace(const void* handle, const float* rays_o, const float* rays_d, int64_t n_rays, int layer_first, int layer_count,
                    float* depth_out, int32_t* tri_out, float* u_out, float* v_out, void* stream);

/* The reference's per-mesh result buffers (bvh.cu:440-468) from one layer's compact record: positions = o + depth*d,
 * unit face normals, triangles_mesh_id / triangles_id (int64), barycentric (1-u-v, u, v). */
int vs_shells_expand(const void* handle, int layer, const float* rays_o, const float* rays_d, const float* depth, const int32_t* tri,
                     const float* u, const float* v, int64_t n_rays, float* positions, float* normals, int64_t* tri_mesh_id,
                     int64_t* tri_id, float* barycentric, void* stream);

/* Unit face normals of packed hits (volsurfs.py:503: surfs_normals[hits, i] = normals[hits]).  n_valid_dev may be NULL. */
int vs_shells_sample_normals(const void* handle, const int32_t* layer_of, const int32_t* tri, int64_t n_samples,
                             const int64_t* n_valid_dev, float* normals, void* stream);

/* Texture coordinates of packed hits (volsurfs_py/methods/volsurfs.py:509-516: uv = sum_j barycentric_j * face_uv_j, barycentric =
 * (1-u-v, u, v)).  bary_uv [n,2]: the packer's samples_uv; face_uvs [sum_l F_l, 3, 2]: per-face-vertex uvs of all layers back to back;
 * face_offset [n_layers] i32: first face row of every layer.  All DEVICE pointers.  n_valid_dev may be NULL. */
int vs_shells_sample_uvs(const int32_t* layer_of, const int32_t* tri, const float* bary_uv, const float* face_uvs, const int32_t* face_offset,
                         int64_t n_samples, const int64_t* n_valid_dev, float* out_uv, void* stream);

/* ---- fused appearance head (tensor cores) ----------------------------------------------------------------------------
 * Replaces RGB.forward (volsurfs_py/models/rgb.py:104-149: [pos features | SH(dirs) | normals?] -> MLP -> sigmoid), MLP.forward
 * (models/mlp.py:8-52), SHEncoder.__call__ (encodings/sphericalharmonics.py:84-153) and the alpha decay of
 * volsurfs_py/methods/volsurfs.py:583-594 at their call sites volsurfs.py:544-549,575-594.
 * dims = [in, h1, ..., out] (n_layers + 1 entries): in <= 128, hidden widths multiples of 16 and <= 128, out <= 8
 * (sigmoid heads; the linear-output variants below take out <= 32). */
int64_t vs_mlp_blob_bytes(int n_layers, const int* dims);
/* weights[l]: DEVICE fp32 [dims[l+1], dims[l]] (torch.nn.Linear layout), biases[l]: DEVICE fp32 [dims[l+1]] or NULL; the pointer
 * arrays themselves are HOST arrays.  Writes the fp16 tensor-core layout + fp32 biases into `blob` (device, 16-byte aligned). */
int vs_mlp_pack(int n_layers, const int* dims, const float* const* weights, const float* const* biases, void* blob, void* stream);
/* out[s,:out] = sigmoid(MLP([pos[s] | SH_deg(dirs[s]) | normals[s] if normal_dep])) * (alpha_decay ? 2*sigmoid(10*clamp(-d.n,0,1))-1 : 1)
 * activation: 0 ReLU, 1 GELU (torch.nn.GELU(): x*Phi(x), evaluated to 4e-5 absolute).  n_valid_dev: optional device int64 capping n_samples.  variant: 0 (debug knob). */
int vs_mlp_forward(int n_layers, const int* dims, const void* blob, int pos_dim, int sh_degree, int normal_dep, int activation,
                   int alpha_decay, const float* pos, const float* dirs, const float* normals, float* out, void* stash, int64_t n_samples,
                   const int64_t* n_valid_dev, int variant, void* stream);
/* Training mode: pass `stash` (vs_mlp_stash_bytes(n_samples) bytes, 16-byte aligned) to vs_mlp_forward and it also keeps, per
 * 128-sample tile, the first layer's fp16 input operand and every hidden layer's fp16 pre-activations for vs_mlp_backward_stashed
 * (what torch autograd saves for backward in the reference; 2 bytes per hidden unit and sample).  NULL: inference. */
int64_t vs_mlp_stash_bytes(int n_layers, const int* dims, int64_t n_samples);

/* ---- importance sampling chain ------------------------------------------------------------------------------
 * VolumeRendering::importance_sample (src/VolumeRendering.cu:466-548; kernel VolumeRenderingGPU.cuh:507-678): n_imp samples per ray
 * drawn from the per-ray cdf (compute_cdf) by inverse-transform sampling, optionally jittered with the reference's pcg32 stream
 * (rng_state / rng_inc = the generator passed by value; default state 0x853c49e6748fea9b, inc 0xda3e39cb94b95bdb; the caller advances
 * it by 2^32 after a jittered call, VolumeRendering.cu:520-523).  Output = the UNCOMPACTED packet: ray r owns rows
 * [r*n_imp, (r+1)*n_imp); rows and out_se of rays without uniform samples are left untouched (pre-fill -1). */
int vs_importance_sample(const float* rays_o, const float* rays_d, const int32_t* se, const float* z, const float* cdf, int64_t n_rays,
                         int64_t n_samples, int n_imp, uint64_t rng_state, uint64_t rng_inc, int jitter, float* out_3d, float* out_dirs,
                         float* out_z, int32_t* out_se, void* stream);
/* VolumeRendering::combine_ray_samples_packets (src/VolumeRendering.cu:550-669; kernel VolumeRenderingGPU.cuh:680-894).
 * vs_combine_offsets: out_start[r] = exclusive prefix sum of count1[r]+count2[r] (VolumeRendering.cu:595-603), *total_dev = grand total.
 * vs_combine_merge: z-ordered merge with the min-distance filter into rows [out_start[r], out_start[r]+written_r) of the combined
 * (uncompacted) packet; c_se of rays without samples is left untouched (pre-fill -1). */
int vs_combine_offsets(const int32_t* se1, const int32_t* se2, int64_t n_rays, int32_t* out_start, int64_t* total_dev, void* scratch,
                       void* stream);
int vs_combine_merge(int64_t n_rays, float min_dist, int values_dim, const int32_t* se1, const int32_t* idx1, const float* p1, const float* d1,
                     const float* z1, const float* v1, const int32_t* se2, const int32_t* idx2, const float* p2, const float* d2,
                     const float* z2, const float* v2, const int32_t* out_start, int32_t* c_idx, float* c_3d, float* c_dirs, float* c_z,
                     float* c_val, int32_t* c_se, void* stream);

/* ---- appearance head, backward (training) -------------------------------------------------------------------------------------
 * Replaces torch autograd through RGB.forward / MLP.forward (models/rgb.py:104-149, models/mlp.py:38-52) in the training step of
 * volsurfs_py/methods/volsurfs.py (loss.backward() in trainer.py): gradients of the Linear weights/biases and of the positional
 * features (the permutohedral encoder's output).  SH features, normals and the alpha decay carry no gradient (no_grad in the
 * reference).  The forward pass is recomputed inside the kernel; nothing has to be saved by vs_mlp_forward. */
/* number of fp32 parameter gradients, laid out [W_0 | b_0 | W_1 | b_1 | ...] with torch.nn.Linear layouts; < 0 on error */
int64_t vs_mlp_num_params(int n_layers, const int* dims);
/* bytes of 16-byte aligned device scratch for vs_mlp_backward on n_samples samples; < 0 on error / unsupported widths */
int64_t vs_mlp_backward_workspace_bytes(int n_layers, const int* dims, int pos_dim, int sh_degree, int normal_dep, int64_t n_samples);
/* d_out [n_samples,out] upstream gradient of vs_mlp_forward's output; d_pos [n_samples,pos_dim] or NULL; d_params flat fp32
 * (overwritten, or added to when accumulate != 0).  Deterministic (fixed summation order for a given grid). */
int vs_mlp_backward(int n_layers, const int* dims, const void* blob, int pos_dim, int sh_degree, int normal_dep, int activation,
                    int alpha_decay, const float* pos, const float* dirs, const float* normals, const float* d_out, float* d_pos,
                    float* d_params, int accumulate, void* workspace, int64_t n_samples, const int64_t* n_valid_dev, int variant,
                    void* stream);
/* Backward from the stash of a training-mode vs_mlp_forward (no GEMM recomputation; HBM-bound).  activation: the forward's.
 * fwd_out: the output that forward wrote. */
int vs_mlp_backward_stashed(int n_layers, const int* dims, const void* blob, const void* stash, int pos_dim, int sh_degree, int normal_dep,
                            int activation, int alpha_decay, const float* dirs, const float* normals, const float* fwd_out, const float* d_out, float* d_pos,
                            float* d_params, int accumulate, void* workspace, int64_t n_samples, const int64_t* n_valid_dev, void* stream);

/* ---- permutohedral-lattice hash encoding ---------------------------------------------------------------------
 * Replaces permutohedral_encoding's Encoding<POS_DIM,2>::forward / ::backward (submodules/permutohedral_encoding/src/Encoding.cu:55-113,
 * 116-217; kernels forward_gpu / backward_gpu / backward_gpu_only_pos, kernels/permutohedral_encoding/EncodingGPU.cuh:68-261,264-416,
 * 534-700), the row permute of PermutoEncoding.forward (src/pytorch_modules/modules.py:85) and — when bb_sides is given — the point
 * normalisation and out-of-bounds mask of volsurfs_py/encodings/permutohash.py:77-86.
 *   lattice [n_levels, capacity, 2] f32 · scale, shift [n_levels, pos_dim] f32 · window [n_levels] f32: DEVICE pointers
 *   bb_sides: HOST pointer to pos_dim floats (points are mapped from [-bb/2, bb/2] onto [0,1]) or NULL (positions used as they are)
 *   out [n, out_stride] f32: the first out_cols of the 2*(n_levels + extra) encoded columns of every row, column = level*2 + feature
 *     (out_cols = output_dims - 1 is the reference's remove_last_element); out_of_bounds [n] u8 or NULL
 *   n_valid_dev: device int64 — only rows < min(n, *n_valid_dev) are processed — or NULL.
 * pos_dim 2..4, 2 features per level, n_levels <= 32. */
int vs_permuto_output_dims(int pos_dim, int n_levels, int concat_points);
int vs_permuto_forward(int pos_dim, int n_levels, int64_t capacity, int concat_points, float points_scaling, const float* bb_sides,
                       const float* positions, const float* lattice, const float* scale, const float* shift, const float* window, float* out,
                       int out_cols, int64_t out_stride, uint8_t* out_of_bounds, int64_t n, const int64_t* n_valid_dev, void* stream);
/* d_out [n, in_stride]: upstream gradient of the first in_cols encoded columns.  d_lattice [n_levels, capacity, 2] is ADDED to (zero it
 * for a fresh gradient; the reference returns zeros + atomics permuted to this layout) or NULL; d_positions [n, pos_dim] is overwritten,
 * or NULL.  As in the reference the concat-points columns pass no gradient to the positions. */
int vs_permuto_backward(int pos_dim, int n_levels, int64_t capacity, int concat_points, const float* bb_sides, const float* positions,
                        const float* lattice, const float* scale, const float* shift, const float* window, const float* d_out, int in_cols,
                        int64_t in_stride, float* d_lattice, float* d_positions, int64_t n, const int64_t* n_valid_dev, void* stream);
/* Same, with a per-position ordering hint order_key [n] (int32, low 5 bits used) or NULL: every block of 128 positions is walked in key
 * order, which makes positions with equal keys neighbours in a warp.  For a packed K-layer ray packet key = samples_layer turns the
 * same-layer hits of neighbouring rays (which share lattice vertices) into adjacent lanes, whose reductions are merged before they leave
 * the SM.  A pure performance hint: the sums are the same contributions in another order (the reference's atomics are unordered too,
 * EncodingGPU.cuh:560-612). */
int vs_permuto_backward_keyed(int pos_dim, int n_levels, int64_t capacity, int concat_points, const float* bb_sides, const float* positions,
                              const int32_t* order_key, const float* lattice, const float* scale, const float* shift, const float* window,
                              const float* d_out, int in_cols, int64_t in_stride, float* d_lattice, float* d_positions, int64_t n,
                              const int64_t* n_valid_dev, void* stream);

/* ---- SH neural textures: the default appearance (SURVEY 8a row a6' / 8f row 4) ------------------------------------------------------
 * Replaces SHNeuralTextures.forward (volsurfs_py/models/sh_neural_textures.py:64-97), NeuralTexture.forward
 * (volsurfs_py/models/neural_texture.py:81-197), the uv helpers it calls (submodules/mvdatasets/mvdatasets/utils/images.py:30-117),
 * SHEncoder.eval (volsurfs_py/encodings/sphericalharmonics.py:156-229) and the tiny-cuda-nn modules built at neural_texture.py:54-79
 * (HashGrid encoding + FullyFusedMLP; un-vendored dependency, restated from its published algorithm: "parity unpinned") at their call
 * sites volsurfs_py/methods/volsurfs.py:539-541,570-572.  Pipeline per SH degree g: vs_hashgrid_forward (texel queries -> features) ->
 * vs_mlp_forward_raw (features -> C*(2g+1) raw outputs) ; then one vs_shtex_combine_forward over all degrees. */
/* tiny-cuda-nn HashGrid level table (2-D): HOST arrays of n_levels entries (any may be NULL); returns the total entry count or < 0 */
int64_t vs_hashgrid_levels(int n_levels, int log2_hashmap_size, int base_resolution, float per_level_scale, float* scale, int32_t* res,
                           int32_t* size, int32_t* offset);
/* features [rows, 2*n_levels] f32 (fp16-representable), rows = n_samples * (mode == 1 ? 4 : 1), row = sample*corners + corner.
 * mode 0: anchor (texel centre), 1: lerp (4 texel corners), 2: uv as given (bake); align: align_to_webgl; uv [n_samples,2]; table [entries,2] f32 (DEVICE). */
int vs_hashgrid_forward(int n_levels, int log2_hashmap_size, int base_resolution, float per_level_scale, int mode, int align, int res_h,
                        int res_w, const float* uv, const float* table, float* features, int64_t n_samples, const int64_t* n_valid_dev,
                        void* stream);
/* d_table [entries,2] f32 is ADDED to (zero it for a fresh gradient) */
int vs_hashgrid_backward(int n_levels, int log2_hashmap_size, int base_resolution, float per_level_scale, int mode, int align, int res_h,
                         int res_w, const float* uv, const float* d_features, float* d_table, int64_t n_samples, const int64_t* n_valid_dev,
                         void* stream);
/* MLP with a LINEAR last layer (tiny-cuda-nn "output_activation": "None"), out <= 32, no SH / normal columns: out[r,:] = MLP(in[r,:]).
 * dims, blob, activation, stash as vs_mlp_forward. */
int vs_mlp_forward_raw(int n_layers, const int* dims, const void* blob, int activation, const float* in, float* out, void* stash,
                       int64_t n_rows, const int64_t* n_valid_dev, void* stream);
/* backward of a training-mode vs_mlp_forward_raw; workspace: vs_mlp_backward_workspace_bytes(n_layers, dims, dims[0], -1, 0, n_rows) */
int vs_mlp_backward_stashed_raw(int n_layers, const int* dims, const void* blob, const void* stash, int activation, const float* d_out, float* d_in,
                                float* d_params, int accumulate, void* workspace, int64_t n_rows, const int64_t* n_valid_dev, void* stream);
/* raw: HOST array of sh_deg+1 DEVICE pointers, raw[g] = [n_samples*corners, C*(2g+1)] f32; res_hw: HOST [sh_deg+1][2] (height, width);
 * range_lo / range_hi: HOST [sh_deg+1] val_range per degree (needed with squeeze); coeffs [n_samples,C,(sh_deg+1)^2] f32 or NULL;
 * dirs [n_samples,3] and out [n_samples,C], or both NULL (view_dirs=None: coefficients only). */
int vs_shtex_combine_forward(int sh_deg, int nr_channels, int mode, int align, const int* res_hw, const float* range_lo, const float* range_hi,
                             int squeeze, int quantize, const float* uv, const float* dirs, const float* const* raw, float* coeffs, float* out,
                             int64_t n_samples, const int64_t* n_valid_dev, void* stream);
/* d_raw: HOST array of DEVICE pointers shaped like raw (overwritten) from g_out [n_samples,C] (+ dirs and the forward's out) or, with
 * dirs == NULL, from g_coeffs [n_samples,C,(sh_deg+1)^2]; gradients of fp16 tensors are rounded to fp16 where torch autograd does. */
int vs_shtex_combine_backward(int sh_deg, int nr_channels, int mode, int align, const int* res_hw, const float* range_lo, const float* range_hi,
                              int squeeze, int quantize, const float* uv, const float* dirs, const float* const* raw, const float* out,
                              const float* g_out, const float* g_coeffs, float* const* d_raw, int64_t n_samples, const int64_t* n_valid_dev,
                              void* stream);

/* ---- ray samplers + occupancy-grid queries ----------------------------------------------------------------------
 * Replace RaySampler::compute_samples_fg / compute_samples_fg_in_grid_occupied_regions / compute_samples_bg (src/RaySampler.cu:72-345;
 * kernels kernels/volsurfs/RaySamplerGPU.cuh:39-488) and OccupancyGrid::get_rays_t_near_t_far / check_occupancy
 * (kernels/volsurfs/OccupancyGridGPU.cuh:318-441; helpers kernels/volsurfs/occ_grid_helpers.h).  The foreground samplers produce the
 * COMPACTED packet directly (what the reference returns after compact_to_valid_samples) in two launches around one prefix sum:
 *   vs_sampler_fg_count (marches; stages 4 bytes of depth per sample) -> vs_segment_offsets -> (host reads the total, sizes the
 *   outputs) -> vs_sampler_fg_write (depth -> rows at their compacted position).
 * nr_voxels_per_dim == 0 selects compute_samples_fg (no grid; extent / occupancy / roi ignored).  extent: HOST float[3];
 * occupancy, roi: DEVICE u8 [nr_voxels_per_dim^3] in Morton order (torch.bool storage).  rng_state / rng_inc: the class's static pcg32
 * (the caller advances it by 2^32 after a jittered call, RaySampler.cu:228-231). */
int vs_segment_offsets(const int32_t* se_in, int64_t n_rays, int32_t* out_start, int64_t* total_dev, void* scratch, void* stream);
/* se_virtual [n,2]: the segment each ray would own in the reference's uncompacted packet ((-1,-1): none); ray_max_dt [n,1] pre-filled
 * with -1 by the caller; z_stage [n * max_nr] f32: depth of sample i of ray r at r*max_nr + i (handed to vs_sampler_fg_write; only the
 * slots of real samples are touched).  roi == NULL: `occupancy` already holds occupancy && roi. */
int vs_sampler_fg_count(const float* rays_o, const float* rays_d, const float* t_entry, const float* t_exit, float min_dist, int min_nr,
                        int max_nr, uint64_t rng_state, uint64_t rng_inc, int jitter, int nr_voxels_per_dim, const float* extent,
                        const uint8_t* occupancy, const uint8_t* roi, int32_t* se_virtual, float* ray_max_dt, float* z_stage, int64_t n_rays,
                        void* stream);
/* out_start [n] from vs_segment_offsets(se_virtual): rows of every sample (samples_idx = its slot in the reference's uncompacted packet,
 * samples_3d = o + z d, samples_dirs, samples_z) at their compacted position + the compacted se_out [n,2] */
int vs_sampler_fg_write(const float* rays_o, const float* rays_d, int max_nr, const int32_t* se_virtual, const float* z_stage,
                        const int32_t* out_start, int32_t* se_out, int32_t* samples_idx, float* samples_3d, float* samples_dirs,
                        float* samples_z, int64_t n_rays, void* stream);
/* compute_samples_bg: nr_samples_per_ray samples per ray, uniform in inverse depth; samples_* hold n_rays*nr_samples_per_ray rows */
int vs_sampler_bg(const float* rays_o, const float* rays_d, const float* t_start, float t_far, int nr_samples_per_ray, uint64_t rng_state,
                  uint64_t rng_inc, int jitter, float* ray_max_dt, float* samples_3d, float* samples_dirs, float* samples_z, int32_t* se,
                  int64_t n_rays, void* stream);
/* RaySampler::contract_samples / uncontract_samples (src/RaySampler.cu:336-427; kernels RaySamplerGPU.cuh:528-658) without the closing
 * update_dt: scene contraction x -> (2 - 1/|2x|) x/|2x| for |2x| > 1 (uncontract: its inverse), depth re-measured from ray_o */
int vs_sampler_contract(const float* ray_o, const int32_t* se, const float* samples_3d, const float* samples_z, float* out_3d, float* out_z,
                        int uncontract, int64_t n_rays, void* stream);
/* OccupancyGrid::get_grid_lower_left_voxels_vertices (centre = 0, src/OccupancyGrid.cu:206-234; kernel OccupancyGridGPU.cuh:31-60) and
 * get_grid_samples / get_random_grid_samples / get_random_grid_samples_in_roi (centre = 1, :236-347; kernel :62-120): world position of
 * the voxels point_indices [n] (Morton order) -> out [n,3]; jitter draws pcg32 floats as the reference (advance(3 * i), 3 draws) */
int vs_occgrid_points(const int32_t* point_indices, int nr_voxels_per_dim, const float* extent, int centre, uint64_t rng_state, uint64_t rng_inc,
                      int jitter, float* out, int64_t n_points, void* stream);
/* OccupancyGrid::update_grid_values (src/OccupancyGrid.cu:446-474; kernel OccupancyGridGPU.cuh:122-147) */
int vs_occgrid_update_values(const int32_t* point_indices, const float* values, float decay, float* grid_values, int64_t n_points, void* stream);
/* OccupancyGrid::update_grid_occupancy_with_density_values (src/OccupancyGrid.cu:476-503; kernel OccupancyGridGPU.cuh:149-218) */
int vs_occgrid_update_occupancy_density(const int32_t* point_indices, int nr_voxels_per_dim, const float* extent, float occupancy_thresh,
                                        int check_neighbours, const float* grid_values, uint8_t* occupancy, int64_t n_points, void* stream);
/* OccupancyGrid::update_grid_occupancy_with_sdf_values (src/OccupancyGrid.cu:505-533; kernel OccupancyGridGPU.cuh:220-316);
 * logistic_beta [n_points,1]; the reference's check_neighbours argument is unused by its kernel and has no counterpart here */
int vs_occgrid_update_occupancy_sdf(const int32_t* point_indices, int nr_voxels_per_dim, const float* extent, const float* logistic_beta,
                                    float occupancy_thresh, const float* grid_values, uint8_t* occupancy, int64_t n_points, void* stream);
/* OccupancyGrid::get_first_rays_sample_start_of_grid_occupied_regions (src/OccupancyGrid.cu:536-573; kernel OccupancyGridGPU.cuh:505-582):
 * samples_* [n_rays, .] and se [n_rays,2] of a RaySamplesPacked(n_rays, n_rays, 0, 1) preset by the caller */
int vs_occgrid_first_sample_start(const float* rays_o, const float* rays_d, const float* t_entry, const float* t_exit, int nr_voxels_per_dim,
                                  const float* extent, const uint8_t* occupancy, const uint8_t* roi, float* samples_3d, float* samples_dirs,
                                  float* samples_z, float* samples_dt, int32_t* se, int64_t n_rays, void* stream);
/* OccupancyGrid::advance_ray_sample_to_next_occupied_voxel (src/OccupancyGrid.cu:575-607; kernel OccupancyGridGPU.cuh:443-503):
 * new_samples_3d [n,3] may alias samples_3d (the reference updates its input in place); is_within_bounds [n,1] bool */
int vs_occgrid_advance_to_next_occupied(const float* samples_dirs, const float* samples_3d, int nr_voxels_per_dim, const float* extent,
                                        const uint8_t* occupancy, const uint8_t* roi, float* new_samples_3d, uint8_t* is_within_bounds,
                                        int64_t n_points, void* stream);
int vs_occgrid_rays_t_near_t_far(const float* rays_o, const float* rays_d, const float* t_entry, const float* t_exit, int nr_voxels_per_dim,
                                 const float* extent, const uint8_t* occupancy, const uint8_t* roi, float* t_near, float* t_far, int64_t n_rays,
                                 void* stream);
int vs_occgrid_check_occupancy(const float* points, int nr_voxels_per_dim, const float* extent, const float* values, const uint8_t* occupancy,
                               const uint8_t* roi, uint8_t* out_occupancy, float* out_values, int64_t n_points, void* stream);

/* ---- baked-texture mesh renderer ------------------------------------------------------------------------------
 * Replaces everything MeshRenderer.render_rays + shade (volsurfs_py/renderers/mesh_renderer.py:62-201) run after the mesh trace: texture
 * coordinates from the barycentrics, TensorTexture(lerp=True) bilinear lookup of the baked SH-coefficient texture
 * (mvdatasets/utils/tensor_texture.py:66-96), fp16 coefficients, SHEncoder.eval (encodings/sphericalharmonics.py:156-229), sigmoid and
 * the shaded output buffers, in one launch.  Inputs as RayTracer.trace returns them (is_hit as u8, triangles_id i64); tex is the
 * ZERO-PADDED texture [(res_h+2), (res_w+2), 4*nr_coeffs] f32 that TensorTexture keeps; nr_coeffs in {1,4,9,16}; bg_rgb: HOST float[3].
 * Outputs ("ray_traced" dict): is_hit [N,1], normals [N,3], uvs [N,3], rgb [N,3], alpha [N,1], view_dirs [N,3] f32. */
int vs_baked_texture_shade(const uint8_t* is_hit, const int64_t* tri_id, const float* bary, const float* dirs, const float* normals,
                           const float* face_uvs, const float* tex, int res_h, int res_w, int nr_coeffs, const float* bg_rgb,
                           float* o_hit, float* o_normals, float* o_uvs, float* o_rgb, float* o_alpha, float* o_dirs, int64_t n_rays,
                           void* stream);

/* ---- training tail of the per-ray path: background blend + L1 loss + its gradient in one launch ------------------------------------
 * Replaces the torch glue between the compositing forward and backward of a training step: pred = rgb_fg + bgT * bg
 * (volsurfs_py/methods/volsurfs.py:708), loss = (gt - pred).abs().mean() (volsurfs_py/utils/losses.py:14-19 as called at
 * volsurfs.py:804-806 without a mask) and what autograd returns for them: g_pred = sign(pred - gt) / numel, g_bgT = sum_c g_pred_c bg_c.
 * rgb_fg, gt, pred, g_pred [n,3]; bgT, g_bgT [n,1]; bg_rgb: HOST float[3]; loss: one device float; scratch: 16 bytes of device memory,
 * zero before the first call (the kernel leaves it zero), not shared between calls that may run concurrently.  The mean is accumulated in
 * fixed point with integer atomics: its value does not depend on the order the blocks finish in (a replayed CUDA graph returns the same bits). */
int vs_blend_l1_loss(const float* rgb_fg, const float* bgT, const float* gt, const float* bg_rgb, float* pred, float* g_pred, float* g_bgT,
                     float* loss, void* scratch, int64_t n_rays, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VOLSURFS_B200_H */
