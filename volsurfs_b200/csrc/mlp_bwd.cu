// volsurfs_b200 — fused appearance head, BACKWARD, on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Replaces what torch autograd does for the reference's legacy RGB / alpha head in a training step
//   volsurfs_py/models/rgb.py:104-149, models/mlp.py:8-52 (Linear + GELU stack, sigmoid output), alpha decay volsurfs.py:583-594 (no_grad)
// i.e. per hidden layer: dZ = dA * act'(Z); dW += dZ^T A; db += sum dZ; dA_prev = dZ W   (cuBLAS SGEMMs + elementwise kernels there)
// with ONE persistent kernel.  A CTA owns tiles of 128 samples; per tile
//   1. the forward pass is recomputed exactly like mlp.cu (nothing was saved by the forward kernel), but every layer's fp16
//      input activations A_l stay resident in shared memory;
//   2. the output epilogue turns dOut into dZ_last = dOut * decay * s(1-s) * scale (fp16, loss-scaled by a power of two chosen from
//      max|dOut| on the device);
//   3. walking the layers backwards, one elected thread issues
//        dW_l^T (+)= A_l^T dZ_l      M = fan-in (TMEM lanes), N = fan-out, K = the tile's 128 samples — both operands are read
//                                     MN-major from the very buffers the forward GEMMs read K-major (no transposes in shared memory);
//                                     the accumulators stay in TMEM for ALL tiles of the CTA (persistent split over CTAs);
//        db_l: a column of ones appended to A_l (fan-in < 128: the bias gradient is lane K_l of dW_l^T) or a 16-column side GEMM;
//        Z_{l-1} = A_{l-1} W_{l-1}^T  (recomputed pre-activations, fp32 in TMEM — shared memory cannot hold them next to A_l);
//      the epilogue reads Z_{l-1}, keeps act'(Z_{l-1}) in registers, then
//        dA_l = dZ_l W_l             W_l read MN-major from the forward blob (no transposed copy of the weights);
//      and dZ_{l-1} = dA_l * act'(Z_{l-1}) goes back to shared memory as fp16;
//   4. the input gradient (first pos_dim columns of dA_0, un-scaled) leaves through shared memory as one TMA bulk store per tile.
// After its last tile a CTA writes its dW/db accumulators to a per-CTA slice of the workspace; mlp_bwd_reduce_kernel sums the slices
// in a fixed order (deterministic), removes the loss scale and writes / accumulates the fp32 parameter gradients.
#include <type_traits>
#include <cstdlib>

#include "mlp_common.cuh"

namespace vs {

constexpr int kOnesBytes = kTileM * 16;  // one 8-channel chunk: [128 samples][8 halves], channel 0 == 1
constexpr int kChunkBytes = kTileM * 16;

struct MlpBwdPlan {
    int dims[kMaxLayers + 1];  // true layer widths
    int a_off[kMaxLayers];     // byte offset of A_l (fp16 [k_pad/8][128][8]) from the start of the activation area
    int fold_bias[kMaxLayers]; // 1: a ones chunk follows A_l, db_l = lane k_pad[l] of dW_l^T; 0: side GEMM into db_col
    int dw_col[kMaxLayers];    // TMEM column of dW_l^T (n_pad[l] columns)
    int db_col[kMaxLayers];    // TMEM column of the bias side GEMM (16 columns) or -1
    int p_off_w[kMaxLayers];   // offsets into the flat parameter-gradient vector [W_0 | b_0 | W_1 | b_1 ...] (torch layouts)
    int p_off_b[kMaxLayers];
    int n_params;
    int act_bytes;             // activation area
    int dz_bytes;              // dZ buffer (also the TMA landing zone of the features and the staging of the input gradient)
    int work_cols, tmem_cols;
    int smem_bytes;
};

static inline int mlp_bwd_plan(const MlpConfig& c, const int* dims, int pos_dim, MlpBwdPlan* p) {
    std::memset(p, 0, sizeof(*p));
    const int L = c.n_layers;
    for (int l = 0; l <= L; ++l) p->dims[l] = dims[l];
    int off = 0, col = 0, work = 16, po = 0;
    for (int l = 0; l < L; ++l) {
        work = std::max(work, std::max(c.k_pad[l], c.n_pad[l]));
        p->a_off[l] = off;
        off += kTileM * c.k_pad[l] * 2;
        p->fold_bias[l] = c.k_pad[l] + 8 <= kMaxWidth ? 1 : 0;
        if (p->fold_bias[l]) off += kOnesBytes;
        p->p_off_w[l] = po;
        po += dims[l] * dims[l + 1];
        p->p_off_b[l] = po;
        po += dims[l + 1];
    }
    p->n_params = po;
    p->act_bytes = off;
    p->work_cols = work;
    col = work;
    for (int l = 0; l < L; ++l) {
        p->dw_col[l] = col;
        col += c.n_pad[l];
        p->db_col[l] = -1;
        if (!p->fold_bias[l]) {
            p->db_col[l] = col;
            col += 16;
        }
    }
    if (col > 512) return VS_ERR_UNSUPPORTED;
    p->tmem_cols = col <= 32 ? 32 : (col <= 64 ? 64 : (col <= 128 ? 128 : (col <= 256 ? 256 : 512)));
    int widest = 16;
    for (int l = 0; l < L; ++l) widest = std::max(widest, c.n_pad[l]);
    const int stage = ((kTileM * pos_dim + 3) & ~3) * 4;
    p->dz_bytes = std::max(kTileM * widest * 2, stage);
    p->dz_bytes = (p->dz_bytes + 127) & ~127;
    // blob | activations | ones (side GEMM operand) | dZ | extras.  An M = 128 (or N = 16) MN-major operand reads a fixed window
    // of 16 (2) chunks from its base whatever the true width is: the windows of every buffer must stay inside the allocation.
    int total = c.blob_bytes + p->act_bytes + kOnesBytes + p->dz_bytes + kTileM * kExtraStride * 4;
    for (int l = 0; l < L; ++l) total = std::max(total, c.blob_bytes + p->a_off[l] + 16 * kChunkBytes);
    total = std::max(total, c.blob_bytes + p->act_bytes + kOnesBytes + 16 * kChunkBytes);
    p->smem_bytes = total + 128;
    return VS_OK;
}

// D[tmem] (+)= A * B over nk K-steps of 16; descriptors advance by a_step / b_step bytes per K-step
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_addr, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_step, uint32_t b_addr,
                                           uint32_t b_lbo, uint32_t b_sbo, uint32_t b_step, uint32_t idesc, int nk, bool accumulate) {
    for (int ks = 0; ks < nk; ++ks) {
        const uint64_t ad = umma_desc(a_addr + (uint32_t)ks * a_step, a_lbo, a_sbo);
        const uint64_t bd = umma_desc(b_addr + (uint32_t)ks * b_step, b_lbo, b_sbo);
        tc_mma_f16(tmem_d, ad, bd, idesc, (accumulate || ks > 0) ? 1u : 0u);
    }
}
constexpr uint32_t kIdescAMn = 1u << 15, kIdescBMn = 1u << 16;  // operand is MN-major (read "transposed")

// power-of-two loss scale from max|dOut|: max * scale lands in [512, 1024) (fp16 keeps 24 binades below that)
__device__ __forceinline__ int scale_exponent(float amax) {
    if (!(amax > 0.f)) return 0;
    const int e = (int)((__float_as_uint(amax) >> 23) & 0xffu) - 127;
    return max(-100, min(100, 9 - e));
}
__device__ __forceinline__ float pow2i(int e) { return __uint_as_float((uint32_t)(127 + e) << 23); }

__global__ void mlp_absmax_kernel(const float* __restrict__ x, int64_t n_rows, int width, const int64_t* __restrict__ n_valid_dev,
                                  float* __restrict__ out) {
    int64_t n = n_rows;
    if (n_valid_dev != nullptr) n = min(n, *n_valid_dev);
    n *= width;
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(__ldg(x + i)));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(VS_FULL_MASK, m, d));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(m));  // non-negative floats order as uints
}

template <int ACT>
__global__ void __launch_bounds__(kMlpThreads) mlp_bwd_kernel(const MlpConfig cfg, const MlpBwdPlan plan, const uint8_t* __restrict__ blob,
                                                              const float* __restrict__ pos, const float* __restrict__ dirs,
                                                              const float* __restrict__ normals, const float* __restrict__ d_out,
                                                              const float* __restrict__ absmax_dev, float* __restrict__ d_pos,
                                                              float* __restrict__ partials, int64_t n_samples,
                                                              const int64_t* __restrict__ n_valid_dev) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_w, bar_in, bar_mma;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int row = tid & (kTileM - 1);
    const int cg = tid >> 7;
    const int F = cfg.pos_dim;
    const int L = cfg.n_layers;
    const int k0 = cfg.k_pad[0];

    uint8_t* s_blob = smem;
    uint8_t* s_act = s_blob + cfg.blob_bytes;
    uint8_t* s_ones = s_act + plan.act_bytes;
    uint8_t* s_dz = s_ones + kOnesBytes;
    float* s_stage = reinterpret_cast<float*>(s_dz);
    float* s_extra = reinterpret_cast<float*>(s_dz + plan.dz_bytes);

    int64_t n = n_samples;
    if (n_valid_dev != nullptr) n = min(n, *n_valid_dev);
    const int64_t n_tiles = (n + kTileM - 1) / kTileM;
    const int sexp = scale_exponent(*absmax_dev);
    const float scale = pow2i(sexp), inv_scale = pow2i(-sexp);

    if (tid == 0) {
        mbar_init(&bar_w, 1);
        mbar_init(&bar_in, 1);
        mbar_init(&bar_mma, 1);
    }
    if (warp == 0) tmem_alloc(&tmem_slot, (uint32_t)plan.tmem_cols);
    // constant "ones" chunks: channel 0 of every sample is 1
    if (tid < kTileM) {
        const uint4 one = make_uint4(0x00003C00u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(s_ones + tid * 16) = one;
        for (int l = 0; l < L; ++l)
            if (plan.fold_bias[l]) *reinterpret_cast<uint4*>(s_act + plan.a_off[l] + kTileM * cfg.k_pad[l] * 2 + tid * 16) = one;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tmem_work = tmem_base + lane_off;  // this warp's lanes of the working region (columns 0..work_cols)

    if ((int64_t)blockIdx.x < n_tiles && tid == 0) {
        mbar_arrive_expect_tx(&bar_w, (uint32_t)cfg.blob_bytes);
        bulk_g2s(s_blob, blob, (uint32_t)cfg.blob_bytes, &bar_w);
    }
    bool weights_ready = false;
    uint32_t par_in = 0, par_mma = 0;
    const bool mn_swap = (cfg.variant & 2) != 0;  // debug: swap LBO/SBO of the MN-major descriptors
    const uint32_t mn_lbo = mn_swap ? (uint32_t)kChunkBytes : 128u;  // K 8-group stride (8 samples x 16 bytes)
    const uint32_t mn_sbo = mn_swap ? 128u : (uint32_t)kChunkBytes;  // MN 8-group stride (one chunk of 8 channels)
    bool first_tile = true;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * kTileM;
        const int rows = (int)min((int64_t)kTileM, n - row0);
        const bool full = rows == kTileM;
        const int64_t r = row0 + row;
        const bool live = row < rows;

        // ---- forward recomputation -------------------------------------------------------------------------------------
        if (full) {
            if (tid == 0) {
                bulk_wait_read_all();  // the previous tile's input-gradient store has finished reading this buffer
                mbar_arrive_expect_tx(&bar_in, (uint32_t)(kTileM * F * 4));
                bulk_g2s(s_stage, pos + row0 * F, (uint32_t)(kTileM * F * 4), &bar_in);
            }
        } else {
            if (tid == 0) bulk_wait_read_all();
            __syncthreads();
            for (int e = tid; e < rows * F; e += kMlpThreads) s_stage[e] = __ldg(pos + row0 * F + e);
        }
        float dx = 0.f, dy = 0.f, dz_ = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
        if (live) {
            if (dirs != nullptr) {
                dx = __ldg(dirs + 3 * r);
                dy = __ldg(dirs + 3 * r + 1);
                dz_ = __ldg(dirs + 3 * r + 2);
            }
            if (normals != nullptr) {
                nx = __ldg(normals + 3 * r);
                ny = __ldg(normals + 3 * r + 1);
                nz = __ldg(normals + 3 * r + 2);
            }
        }
        if (cg == 0) {
            float sh[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) sh[i] = 0.f;
            sh_eval(dx, dy, dz_, cfg.n_sh, sh);
            float* ex = s_extra + row * kExtraStride;
#pragma unroll
            for (int i = 0; i < 16; ++i) ex[i] = sh[i];
            if (cfg.normal_dep) {
                ex[cfg.n_sh] = nx;
                ex[cfg.n_sh + 1] = ny;
                ex[cfg.n_sh + 2] = nz;
            }
        }
        __syncthreads();
        if (full) {
            mbar_wait(&bar_in, par_in);
            par_in ^= 1;
        }
        {
            __half* s_a0 = reinterpret_cast<__half*>(s_act + plan.a_off[0]);
            const float* srow = s_stage + row * F;
            const float* ex = s_extra + row * kExtraStride;
            const int in_dim = cfg.in_dim;
            for (int kc = cg; kc < k0 / 8; kc += 4) {
                __half2 h[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v2[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int c = kc * 8 + 2 * j + q;
                        float v = 0.f;
                        if (live && c < in_dim) v = c < F ? srow[c] : ex[c - F];
                        v2[q] = v;
                    }
                    h[j] = __floats2half2_rn(v2[0], v2[1]);
                }
                *reinterpret_cast<uint4*>(s_a0 + ((size_t)kc * kTileM + row) * 8) = *reinterpret_cast<const uint4*>(h);
            }
        }
        if (!weights_ready) {
            mbar_wait(&bar_w, 0);
            weights_ready = true;
        }

        for (int l = 0; l < L; ++l) {
            const int K = cfg.k_pad[l], N = cfg.n_pad[l];
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                issue_gemm(tmem_base, smem_u32(s_act + plan.a_off[l]), kChunkBytes, 128, 2 * kChunkBytes, smem_u32(s_blob + cfg.w_off[l]),
                           (uint32_t)N * 16, 128, 2 * (uint32_t)N * 16, umma_idesc_f16(kTileM, N), K / 16, false);
                tc_commit(&bar_mma);
            }
            mbar_wait(&bar_mma, par_mma);
            par_mma ^= 1;
            tc_fence_after();
            const float* bias = reinterpret_cast<const float*>(s_blob + cfg.b_off[l]);
            if (l + 1 < L) {
                __half* s_next = reinterpret_cast<__half*>(s_act + plan.a_off[l + 1]);
                for (int c0 = cg * 16; c0 < N; c0 += 64) {
                    float v[16];
                    tmem_ld16(tmem_work + (uint32_t)c0, v);
                    const float4* b4 = reinterpret_cast<const float4*>(bias + c0);
                    __half2 h[8];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 bb = b4[q];
                        float a0 = v[4 * q] + bb.x, a1 = v[4 * q + 1] + bb.y, a2 = v[4 * q + 2] + bb.z, a3 = v[4 * q + 3] + bb.w;
                        if (ACT == 1) {
                            a0 = gelu_fwd(a0);
                            a1 = gelu_fwd(a1);
                            a2 = gelu_fwd(a2);
                            a3 = gelu_fwd(a3);
                        } else {
                            a0 = fmaxf(a0, 0.f);
                            a1 = fmaxf(a1, 0.f);
                            a2 = fmaxf(a2, 0.f);
                            a3 = fmaxf(a3, 0.f);
                        }
                        h[2 * q] = __floats2half2_rn(a0, a1);
                        h[2 * q + 1] = __floats2half2_rn(a2, a3);
                    }
                    uint4* dst = reinterpret_cast<uint4*>(s_next + ((size_t)(c0 / 8) * kTileM + row) * 8);
                    dst[0] = *reinterpret_cast<const uint4*>(&h[0]);
                    dst[kTileM] = *reinterpret_cast<const uint4*>(&h[4]);
                }
            } else if (cg == 0) {
                // output layer: dZ = dOut * decay * s (1 - s) * scale, fp16, columns >= out_dim are zero (N == 16: two chunks)
                float v[16];
                tmem_ld16(tmem_work, v);
                float decay = 1.f;
                if (cfg.alpha_decay) {
                    const float dot = fminf(fmaxf(-(dx * nx + dy * ny + dz_ * nz), 0.f), 1.f);
                    decay = 2.f * sigmoid_f(10.f * dot) - 1.f;
                }
                float g[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    g[j] = 0.f;
                    if (live && j < cfg.out_dim) {
                        const float s = sigmoid_f(v[j] + bias[j]);
                        g[j] = __ldg(d_out + r * cfg.out_dim + j) * decay * s * (1.f - s) * scale;
                    }
                }
                __half2 h[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) h[q] = __floats2half2_rn(g[2 * q], g[2 * q + 1]);
                uint4* dst = reinterpret_cast<uint4*>(s_dz + (size_t)row * 16);
                dst[0] = *reinterpret_cast<const uint4*>(h);
                dst[kTileM] = make_uint4(0u, 0u, 0u, 0u);
            }
        }

        // ---- backward ----------------------------------------------------------------------------------------------------
        for (int l = L - 1; l >= 0; --l) {
            const int K = cfg.k_pad[l], N = cfg.n_pad[l];
            const uint32_t a_addr = smem_u32(s_act + plan.a_off[l]);
            const uint32_t dz_addr = smem_u32(s_dz);
            const bool want_dx = (l == 0) && d_pos != nullptr;
            // W_l read MN-major from the forward blob ([k/8][n_pad][8] halves): fan-in groups are n_pad*16 bytes apart, groups of
            // 8 fan-out rows 128 bytes
            const uint32_t w_lbo = mn_swap ? (uint32_t)N * 16 : 128u, w_sbo = mn_swap ? 128u : (uint32_t)N * 16;
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                // dW_l^T (+)= A_l^T dZ_l: lanes = fan-in (+ the ones chunk: bias gradient), columns = fan-out
                issue_gemm(tmem_base + (uint32_t)plan.dw_col[l], a_addr, mn_lbo, mn_sbo, 256, dz_addr, mn_lbo, mn_sbo, 256,
                           umma_idesc_f16(kTileM, N) | kIdescAMn | kIdescBMn, kTileM / 16, !first_tile);
                if (!plan.fold_bias[l])  // db_l (+)= dZ_l^T 1: lanes = fan-out, column 0
                    issue_gemm(tmem_base + (uint32_t)plan.db_col[l], dz_addr, mn_lbo, mn_sbo, 256, smem_u32(s_ones), mn_lbo, mn_sbo, 256,
                               umma_idesc_f16(kTileM, 16) | kIdescAMn | kIdescBMn, kTileM / 16, !first_tile);
                if (l > 0) {  // Z_{l-1} again (pre-activations of the layer below)
                    const int Kp = cfg.k_pad[l - 1], Np = cfg.n_pad[l - 1];
                    issue_gemm(tmem_base, smem_u32(s_act + plan.a_off[l - 1]), kChunkBytes, 128, 2 * kChunkBytes,
                               smem_u32(s_blob + cfg.w_off[l - 1]), (uint32_t)Np * 16, 128, 2 * (uint32_t)Np * 16, umma_idesc_f16(kTileM, Np),
                               Kp / 16, false);
                } else if (want_dx) {  // dA_0 = dZ_0 W_0 (input gradient)
                    issue_gemm(tmem_base, dz_addr, kChunkBytes, 128, 2 * kChunkBytes, smem_u32(s_blob + cfg.w_off[0]), w_lbo, w_sbo, 256,
                               umma_idesc_f16(kTileM, K) | kIdescBMn, N / 16, false);
                }
                tc_commit(&bar_mma);
            }
            mbar_wait(&bar_mma, par_mma);
            par_mma ^= 1;
            tc_fence_after();

            if (l > 0) {
                // act'(Z_{l-1}) into registers (this thread's row, its 16-column slices)
                const float* bias = reinterpret_cast<const float*>(s_blob + cfg.b_off[l - 1]);
                __half2 gq[2][8];
#pragma unroll
                for (int it = 0; it < 2; ++it) {
                    const int c0 = cg * 16 + 64 * it;
                    if (c0 < K) {
                        float v[16];
                        tmem_ld16(tmem_work + (uint32_t)c0, v);
                        const float4* b4 = reinterpret_cast<const float4*>(bias + c0);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 bb = b4[q];
                            float a0 = v[4 * q] + bb.x, a1 = v[4 * q + 1] + bb.y, a2 = v[4 * q + 2] + bb.z, a3 = v[4 * q + 3] + bb.w;
                            if (ACT == 1) {
                                a0 = gelu_grad(a0);
                                a1 = gelu_grad(a1);
                                a2 = gelu_grad(a2);
                                a3 = gelu_grad(a3);
                            } else {
                                a0 = a0 > 0.f ? 1.f : 0.f;
                                a1 = a1 > 0.f ? 1.f : 0.f;
                                a2 = a2 > 0.f ? 1.f : 0.f;
                                a3 = a3 > 0.f ? 1.f : 0.f;
                            }
                            gq[it][2 * q] = __floats2half2_rn(a0, a1);
                            gq[it][2 * q + 1] = __floats2half2_rn(a2, a3);
                        }
                    }
                }
                tc_fence_before();
                __syncthreads();  // everyone has read Z: the working columns may be overwritten
                if (tid == 0) {
                    tc_fence_after();
                    // dA_l = dZ_l W_l: A = dZ_l K-major (K = fan-out), B = W_l read MN-major (N = fan-in)
                    issue_gemm(tmem_base, dz_addr, kChunkBytes, 128, 2 * kChunkBytes, smem_u32(s_blob + cfg.w_off[l]), w_lbo, w_sbo, 256,
                               umma_idesc_f16(kTileM, K) | kIdescBMn, N / 16, false);
                    tc_commit(&bar_mma);
                }
                mbar_wait(&bar_mma, par_mma);
                par_mma ^= 1;
                tc_fence_after();
                // dZ_{l-1} = dA_l * act'(Z_{l-1}) -> fp16 (every MMA that read dZ_l has completed)
#pragma unroll
                for (int it = 0; it < 2; ++it) {
                    const int c0 = cg * 16 + 64 * it;
                    if (c0 < K) {
                        float v[16];
                        tmem_ld16(tmem_work + (uint32_t)c0, v);
                        __half2 h[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float2 gg = __half22float2(gq[it][q]);
                            h[q] = __floats2half2_rn(v[2 * q] * gg.x, v[2 * q + 1] * gg.y);
                        }
                        uint4* dst = reinterpret_cast<uint4*>(s_dz + ((size_t)(c0 / 8) * kTileM + row) * 16);
                        dst[0] = *reinterpret_cast<const uint4*>(&h[0]);
                        dst[kTileM] = *reinterpret_cast<const uint4*>(&h[4]);
                    }
                }
            } else if (want_dx) {
                // input gradient: first pos_dim columns of dA_0, un-scaled, through shared memory (dZ_0 is dead now)
#pragma unroll
                for (int it = 0; it < 2; ++it) {
                    const int c0 = cg * 16 + 64 * it;
                    if (c0 < F) {
                        float v[16];
                        tmem_ld16(tmem_work + (uint32_t)c0, v);
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < F) s_stage[row * F + c0 + j] = v[j] * inv_scale;
                    }
                }
                fence_proxy_async();
                __syncthreads();
                if (full) {
                    if (tid == 0) {
                        bulk_s2g(d_pos + row0 * F, s_stage, (uint32_t)(kTileM * F * 4));
                        bulk_commit();
                    }
                } else {
                    for (int e = tid; e < rows * F; e += kMlpThreads) d_pos[row0 * F + e] = s_stage[e];
                }
            }
        }
        first_tile = false;
        fence_proxy_async();  // generic writes to the dZ / staging buffer before the next tile's TMA load lands there
        tc_fence_before();
        __syncthreads();
    }

    // ---- this CTA's parameter-gradient accumulators -> its slice of the workspace ---------------------------------------
    if ((int64_t)blockIdx.x < n_tiles) {
        tc_fence_after();
        float* mine = partials + (size_t)blockIdx.x * plan.n_params;
        for (int l = 0; l < L; ++l) {
            const int N = cfg.n_pad[l], Kt = plan.dims[l], Nt = plan.dims[l + 1];
            for (int c0 = cg * 16; c0 < N; c0 += 64) {
                float v[16];
                tmem_ld16(tmem_work + (uint32_t)(plan.dw_col[l] + c0), v);
                if (row < Kt) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < Nt) mine[plan.p_off_w[l] + (c0 + j) * Kt + row] = v[j];
                } else if (plan.fold_bias[l] && row == cfg.k_pad[l]) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < Nt) mine[plan.p_off_b[l] + c0 + j] = v[j];
                }
            }
            if (!plan.fold_bias[l] && cg == 0) {
                float v[16];
                tmem_ld16(tmem_work + (uint32_t)plan.db_col[l], v);
                if (row < Nt) mine[plan.p_off_b[l] + row] = v[0];
            }
        }
        if (tid == 0) bulk_wait_read_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)plan.tmem_cols);
}

// d_params[p] (+)= 2^-scale_exp * sum over CTA slices, in slice order
__global__ void mlp_bwd_reduce_kernel(const float* __restrict__ partials, int n_slices, int n_params, const float* __restrict__ absmax_dev,
                                      const int64_t* __restrict__ n_valid_dev, int64_t n_samples, float* __restrict__ d_params, int accumulate) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_params) return;
    int64_t n = n_samples;
    if (n_valid_dev != nullptr) n = min(n, *n_valid_dev);
    const int64_t tiles = (n + kTileM - 1) / kTileM;
    const int live = (int)min((int64_t)n_slices, tiles);  // CTAs without a tile wrote nothing
    float s = 0.f;
    for (int i = 0; i < live; ++i) s += partials[(size_t)i * n_params + p];
    s *= pow2i(-scale_exponent(*absmax_dev));
    d_params[p] = accumulate ? d_params[p] + s : s;
}

// =====================================================================================================================
// Backward from the activation stash (training-mode vs_mlp_forward): no recomputation, no activation-derivative math.
//
// Per 128-sample tile the kernel needs, in this order: dZ_last (built from dOut and the forward output), A_{L-1}, dZ_{L-2}, A_{L-2},
// ..., dZ_0, A_0 and a staging area for the input gradient.  Lifetimes are first-in-first-out, so the operands live in ONE
// shared-memory ring.  The CTA is warp-specialised:
//   * the control warp (one lane) walks the item sequence ahead of everybody else (across tile boundaries): it reserves ring space,
//     fetches every A_l with a TMA bulk copy as soon as there is room, and per layer issues
//         dA_l  = dZ_l W_l        (commit -> bar_da: the epilogue warps wait for this one only)
//         dW_l^T (+)= A_l^T dZ_l  (+ bias side GEMM; commit -> bar_dw: runs under the epilogue; then dZ_l and A_l leave the ring)
//   * the 16 epilogue warps fetch their own elements of G_{l-1} (the saved activation derivative) straight from global memory into
//     registers BEFORE they wait for dA_l, multiply and write dZ_{l-1} (fp16) into its ring slot, and announce it on `ready`.
// Every thread replays the same allocator arithmetic, so ring offsets need no communication; an item's mbarrier says "landed" (TMA
// items) or "this space is yours" (items the kernel produces itself).
struct MlpBwd2Plan {
    int n_items;                        // items per tile: 2L (+1 with the input gradient)
    int item_bytes[2 * kMaxLayers + 2];
    int item_src[2 * kMaxLayers + 2];   // byte offset inside the tile's stash image, -1: produced by the kernel
    int ring_bytes;
    int want_dx;
    int smem_bytes;
    int prefetch_tiles;                 // whole stash tiles the control thread pulls into L2 ahead of the ring (0: off)
};

constexpr int kItemBars = 16;
constexpr int kBwdThreads = kMlpThreads + 96;  // 16 epilogue warps + the MMA issuer warp + the ring warp + the output-gradient warp

static inline int mlp_bwd2_plan(const MlpConfig& c, const MlpStash& st, int pos_dim, int want_dx, MlpBwd2Plan* p) {
    std::memset(p, 0, sizeof(*p));
    const int L = c.n_layers;
    int i = 0;
    p->item_bytes[i] = kTileM * c.n_pad[L - 1] * 2;  // dZ of the output layer
    p->item_src[i++] = -1;
    for (int l = L - 1; l >= 0; --l) {
        p->item_bytes[i] = st.a_bytes[l];
        p->item_src[i++] = st.a_off[l];
        if (l >= 1) {
            p->item_bytes[i] = kTileM * c.n_pad[l - 1] * 2;  // dZ_{l-1}
            p->item_src[i++] = -1;
        }
    }
    if (want_dx) {
        p->item_bytes[i] = (kTileM * pos_dim * 4 + 127) & ~127;
        p->item_src[i++] = -1;
    }
    p->n_items = i;
    p->want_dx = want_dx;
    // ring | ones (bias side-GEMM operand) | blob.  MN-major operands read a fixed 16-chunk (32 KB) window from their base: what
    // lies behind the ring must cover it.
    const int tail = kOnesBytes + std::max(c.blob_bytes, 16 * kChunkBytes);
    int ring = (227 * 1024 - 2048 - tail) / 2048 * 2048;  // 2 KB: the kernel's static shared memory (barriers, table copies) + alignment
    int top3[3] = {0, 0, 0};
    for (int k = 0; k < p->n_items; ++k) {
        int b = p->item_bytes[k];
        for (int t = 0; t < 3; ++t)
            if (b > top3[t]) std::swap(b, top3[t]);
    }
    const int need = 2 * (top3[0] + top3[1] + top3[2]);  // three live operands + the next layer's prefetch + wrap fragments
    ring = std::min(ring, std::max(need, 64 * 1024));
    if (ring < top3[0] + top3[1] + top3[2] + top3[0]) return VS_ERR_UNSUPPORTED;
    p->ring_bytes = ring;
    p->smem_bytes = ring + tail + 128;
    p->prefetch_tiles = 1;  // measured on B200 (893k samples, [128,128,64]): 0 -> 0.391 ms, 1 -> 0.387, 2 -> 0.404, 4 -> 0.506
    return VS_OK;
}

struct RingCursor {
    int head;
    __device__ __forceinline__ int alloc(int bytes, int ring_bytes) {
        if (head + bytes > ring_bytes) head = 0;
        const int off = head;
        head += bytes;
        return off;
    }
};

// Opt-in clock64 event trace of CTA 0 (scripts/trace_mlp_bwd.py builds a private library with -DVS_KERNEL_TRACE; the product
// build compiles the macros to nothing).  who: 0 MMA issuer, 1 epilogue thread 0, 2 ring lane.  An event is one clock read and one
// fire-and-forget store (the event counter lives in a register), a few tens of cycles.
#ifdef VS_KERNEL_TRACE
constexpr int kTraceCap = 256;
__device__ long long g_trace[3][2 * kTraceCap];
__device__ int g_trace_n[3];
#define VS_TR_DECL int tr_n_ = 0
#define VS_TR(who, id)                                                             \
    do {                                                                           \
        if (blockIdx.x == 0 && k >= 2 && k < 4 && tr_n_ < kTraceCap) {             \
            g_trace[who][2 * tr_n_] = (id);                                        \
            g_trace[who][2 * tr_n_ + 1] = clock64();                               \
            ++tr_n_;                                                               \
        }                                                                          \
    } while (0)
#define VS_TR_END(who)                             \
    do {                                           \
        if (blockIdx.x == 0) g_trace_n[who] = tr_n_; \
    } while (0)
#else
#define VS_TR_DECL ((void)0)
#define VS_TR(who, id) ((void)0)
#define VS_TR_END(who) ((void)0)
#endif
#define VS_TR_E(id)               \
    do {                          \
        if (tid == 0) VS_TR(1, id); \
    } while (0)

template <int ACT>  // 0 ReLU, 1 GELU: the backward recomputes act / act' from the stashed pre-activations
__global__ void __launch_bounds__(kBwdThreads, 1) mlp_bwd_stashed_kernel(const __grid_constant__ MlpConfig cfg_param,
                                                                         const __grid_constant__ MlpBwdPlan plan_param,
                                                                         const __grid_constant__ MlpBwd2Plan p2_param,
                                                                         const __grid_constant__ MlpStash st_param, const uint8_t* __restrict__ blob,
                                                                      const uint8_t* __restrict__ stash, const float* __restrict__ dirs,
                                                                      const float* __restrict__ normals, const float* __restrict__ fwd_out,
                                                                      const float* __restrict__ d_out, const float* __restrict__ absmax_dev,
                                                                      float* __restrict__ d_pos, float* __restrict__ partials,
                                                                      int64_t n_samples, const int64_t* __restrict__ n_valid_dev) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_w, bar_da, bar_dw[4], bar_ready, bar_x, bar_act[2], bar_item[kItemBars], bar_fill[kItemBars];
    __shared__ uint32_t tmem_slot;
    // per-layer tables indexed with the run-time layer number: shared-memory copies (see mlp_fwd_kernel)
    __shared__ MlpConfig cfg;
    __shared__ MlpBwdPlan plan;
    __shared__ MlpBwd2Plan p2;
    __shared__ MlpStash st;
    {
        auto copy_words = [&](void* dst, const void* src, int words) {
            for (int i = threadIdx.x; i < words; i += blockDim.x) reinterpret_cast<int*>(dst)[i] = reinterpret_cast<const int*>(src)[i];
        };
        copy_words(&cfg, &cfg_param, (int)(sizeof(MlpConfig) / 4));
        copy_words(&plan, &plan_param, (int)(sizeof(MlpBwdPlan) / 4));
        copy_words(&p2, &p2_param, (int)(sizeof(MlpBwd2Plan) / 4));
        copy_words(&st, &st_param, (int)(sizeof(MlpStash) / 4));
    }
    __syncthreads();

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int row = tid & (kTileM - 1);
    const int cg = (tid >> 7) & 3;
    const int F = cfg.pos_dim;
    const int L = cfg.n_layers;
    const int R = p2.ring_bytes;

    uint8_t* s_ring = smem;
    uint8_t* s_ones = s_ring + R;
    uint8_t* s_blob = s_ones + kOnesBytes;

    int64_t n = n_samples;
    if (n_valid_dev != nullptr) n = min(n, *n_valid_dev);
    const int64_t n_tiles = (n + kTileM - 1) / kTileM;
    const int64_t my_tiles = (int64_t)blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int sexp = scale_exponent(*absmax_dev);
    const float scale = pow2i(sexp), inv_scale = pow2i(-sexp);

    if (tid == 0) {
        mbar_init(&bar_w, 1);
        mbar_init(&bar_da, 1);
        for (int i = 0; i < 4; ++i) mbar_init(&bar_dw[i], 1);
        mbar_init(&bar_ready, kMlpThreads / 32);
        mbar_init(&bar_x, kMlpThreads / 32);
        mbar_init(&bar_act[0], kMlpThreads / 32);
        mbar_init(&bar_act[1], kMlpThreads / 32);
        for (int i = 0; i < kItemBars; ++i) mbar_init(&bar_item[i], 1);
        for (int i = 0; i < kItemBars; ++i) mbar_init(&bar_fill[i], 1);
    }
    if (warp == 0) tmem_alloc(&tmem_slot, (uint32_t)plan.tmem_cols);
    if (tid < kTileM) *reinterpret_cast<uint4*>(s_ones + tid * 16) = make_uint4(0x00003C00u, 0u, 0u, 0u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t mn_lbo = 128u, mn_sbo = (uint32_t)kChunkBytes;
    auto item_bar = [&](uint32_t seq) { return &bar_item[seq & (kItemBars - 1)]; };
    auto item_parity = [&](uint32_t seq) { return (seq / kItemBars) & 1u; };

    if (warp == kMlpThreads / 32 + 2) {
        // ================= output-gradient warp: fills item 0 (dZ of the output layer) of every tile =================
        // dZ_last = dOut * d(out)/dz * scale, d(out)/dz = out (1 - out/decay)  (out = decay * sigmoid(z)); linear heads: dOut * scale.
        // A lane owns four rows.  This used to be the first thing the epilogue warps did in a tile: ~1.7k cycles of dependent loads
        // and scalar code with the tensor core and every other warp waiting (clock64 trace).  Here it runs as far ahead of them as
        // the ring reserves the item (most of a tile).  bar_fill uses the item barriers' slot scheme, so a slot's next phase cannot
        // start before the issuer has consumed this one (the item is released after the GEMMs that read it).
        if (my_tiles > 0) {
            RingCursor cur{0};
            uint32_t seq = 0;
            const int n_items = p2.n_items;
            const int n_chunks_last = cfg.n_pad[L - 1] / 8;
            const int od = cfg.out_dim;
            const bool linear = cfg.out_linear != 0, use_decay = cfg.alpha_decay != 0;
            for (int64_t k = 0; k < my_tiles; ++k) {
                const int64_t row0 = (blockIdx.x + k * gridDim.x) * kTileM;
                const int off_last = cur.alloc(p2.item_bytes[0], R);
                for (int i = 1; i < n_items; ++i) cur.alloc(p2.item_bytes[i], R);
                const uint32_t seq0 = seq;
                seq += (uint32_t)n_items;
                // the values are computed into registers BEFORE the wait for the item's ring space (that reservation comes late: the
                // item sits behind the previous tile's input-gradient item), all loads of a batch issued together
                if (!linear) {  // sigmoid heads: at most 8 outputs, chunk 0 carries them, the other chunks are zero
                    // W outputs wide, ROWS rows of the lane in flight at a time (the first batch before the wait for the ring space)
                    auto fill = [&](auto w_tag, auto rows_tag) {
                        constexpr int W = decltype(w_tag)::value, ROWS = decltype(rows_tag)::value;
#pragma unroll 1
                        for (int q0 = 0; q0 < kTileM / 32; q0 += ROWS) {
                            uint4 hv[ROWS];
#pragma unroll
                            for (int q = 0; q < ROWS; ++q) {
                                const int64_t r_raw = row0 + lane + 32 * (q0 + q);
                                const int64_t r = min(r_raw, n - 1);  // clamped: the loads carry no branch; dead rows are zeroed below
                                float d[W], o[W], dn[6];
#pragma unroll
                                for (int j = 0; j < W; ++j) {
                                    const int col = min(j, od - 1);
                                    d[j] = __ldg(d_out + r * od + col);
                                    o[j] = __ldg(fwd_out + r * od + col);
                                }
                                float decay = 1.f;
                                if (use_decay) {
#pragma unroll
                                    for (int j = 0; j < 3; ++j) {
                                        dn[j] = __ldg(dirs + 3 * r + j);
                                        dn[3 + j] = __ldg(normals + 3 * r + j);
                                    }
                                    const float dot = fminf(fmaxf(-(dn[0] * dn[3] + dn[1] * dn[4] + dn[2] * dn[5]), 0.f), 1.f);
                                    decay = 2.f * sigmoid_f(10.f * dot) - 1.f;
                                }
                                float g[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) g[j] = 0.f;
#pragma unroll
                                for (int j = 0; j < W; ++j) {
                                    const float ds = decay != 0.f ? o[j] * (1.f - __fdividef(o[j], decay)) : 0.f;
                                    if (r_raw < n && j < od) g[j] = d[j] * ds * scale;
                                }
                                __half2 h[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(g[2 * j], g[2 * j + 1]);
                                hv[q] = *reinterpret_cast<const uint4*>(h);
                            }
                            if (q0 == 0) mbar_wait(item_bar(seq0), item_parity(seq0));  // the ring space is reserved
#pragma unroll
                            for (int q = 0; q < ROWS; ++q) {
                                uint4* dst = reinterpret_cast<uint4*>(s_ring + off_last + (size_t)(lane + 32 * (q0 + q)) * 16);
                                dst[0] = hv[q];
                                for (int c = 1; c < n_chunks_last; ++c) dst[c * kTileM] = make_uint4(0u, 0u, 0u, 0u);
                            }
                        }
                    };
                    if (od <= 4) fill(std::integral_constant<int, 4>{}, std::integral_constant<int, 4>{});
                    else fill(std::integral_constant<int, 8>{}, std::integral_constant<int, 2>{});
                } else {  // linear heads (texture nets, up to 32 outputs): two rows per lane in registers at a time
                    constexpr int kMaxChunks = 4;  // the output layer is at most 32 columns wide (mlp_layout)
#pragma unroll 1
                    for (int q0 = 0; q0 < kTileM / 32; q0 += 2) {
                        uint4 hv[2][kMaxChunks];
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int64_t r = min(row0 + lane + 32 * (q0 + q), n - 1);
                            const bool live = row0 + lane + 32 * (q0 + q) < n;
#pragma unroll
                            for (int c = 0; c < kMaxChunks; ++c) {
                                float d[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) d[j] = __ldg(d_out + r * od + min(c * 8 + j, od - 1));
                                __half2 h[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    h[j] = __floats2half2_rn((live && c * 8 + 2 * j < od) ? d[2 * j] * scale : 0.f,
                                                             (live && c * 8 + 2 * j + 1 < od) ? d[2 * j + 1] * scale : 0.f);
                                hv[q][c] = *reinterpret_cast<const uint4*>(h);
                            }
                        }
                        if (q0 == 0) mbar_wait(item_bar(seq0), item_parity(seq0));
#pragma unroll
                        for (int q = 0; q < 2; ++q)
#pragma unroll
                            for (int c = 0; c < kMaxChunks; ++c)
                                if (c < n_chunks_last)
                                    *reinterpret_cast<uint4*>(s_ring + off_last + ((size_t)c * kTileM + lane + 32 * (q0 + q)) * 16) = hv[q][c];
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_fill[seq0 & (kItemBars - 1)]);
            }
        }
    } else if (warp == kMlpThreads / 32 + 1) {
        // ================= ring warp: TMA producer, ring-space accounting, input-gradient store =================
        // Kept off the MMA issuer's lane: a release + produce round is several hundred cycles of dependent scalar work, which sat
        // between every layer's dW commit and the next layer's dA issue while one lane did both jobs (clock64 trace, round 2).
        if (lane == 0 && my_tiles > 0) {
            mbar_arrive_expect_tx(&bar_w, (uint32_t)cfg.blob_bytes);
            bulk_g2s(s_blob, blob, (uint32_t)cfg.blob_bytes, &bar_w);
            // producer / release cursors replay the allocator; counters are kept incrementally (no division on this lane's path)
            const int n_items = p2.n_items;
            int64_t prod_seq = 0, free_seq = 0;  // items [free_seq, prod_seq) are live
            const int64_t total_items = my_tiles * n_items;
            int prod_item = 0, free_item = 0;    // position inside the tile's item list
            int64_t prod_tile = blockIdx.x;
            RingCursor prod{0}, freed{0};
            int used = 0;                         // ring bytes held by live items (including wrap fragments)
            auto produce = [&]() {
                while (prod_seq < total_items && prod_seq - free_seq < kItemBars - 1) {
                    const int bytes = p2.item_bytes[prod_item];
                    const int skip = (prod.head + bytes > R) ? R - prod.head : 0;
                    if (used + skip + bytes > R) break;
                    const int off = prod.alloc(bytes, R);
                    used += skip + bytes;
                    uint64_t* bar = &bar_item[prod_seq & (kItemBars - 1)];
                    const int src = p2.item_src[prod_item];
                    if (src >= 0) {
                        mbar_arrive_expect_tx(bar, (uint32_t)bytes);
                        bulk_g2s(s_ring + off, stash + prod_tile * (int64_t)st.tile_bytes + src, (uint32_t)bytes, bar);
                    } else {
                        mbar_arrive(bar);  // produced by the kernel itself: "this ring space is yours"
                    }
                    ++prod_seq;
                    if (++prod_item == n_items) {
                        prod_item = 0;
                        prod_tile += gridDim.x;
                    }
                }
            };
            auto release = [&](int count) {
                for (int k = 0; k < count; ++k) {
                    const int bytes = p2.item_bytes[free_item];
                    const int skip = (freed.head + bytes > R) ? R - freed.head : 0;
                    freed.alloc(bytes, R);
                    used -= skip + bytes;
                    ++free_seq;
                    if (++free_item == n_items) free_item = 0;
                }
            };
            produce();
            const int pf_depth = p2.prefetch_tiles;
            // a future tile's image, and the rows of the output-layer gradient the epilogue warps read with plain loads at the top of
            // that tile (a DRAM miss there stalls the whole CTA for 2.5 - 3.8k cycles under this kernel's traffic: clock64 trace)
            auto prefetch_rows = [&](const float* rows, int width, int64_t tile) {  // bulk prefetches want 16-byte aligned addresses
                const float* p = rows + tile * kTileM * width;
                if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) bulk_prefetch_l2(p, (uint32_t)(kTileM * width * 4));
            };
            auto prefetch_tile = [&](int64_t tile) {
                bulk_prefetch_l2(stash + tile * (int64_t)st.tile_bytes, (uint32_t)st.tile_bytes);
                if ((tile + 1) * kTileM <= n) {
                    prefetch_rows(d_out, cfg.out_dim, tile);
                    if (!cfg.out_linear) {
                        prefetch_rows(fwd_out, cfg.out_dim, tile);
                        if (cfg.alpha_decay) {
                            prefetch_rows(dirs, 3, tile);
                            prefetch_rows(normals, 3, tile);
                        }
                    }
                }
            };
            for (int64_t j = 0; j < pf_depth && j < my_tiles; ++j) prefetch_tile(blockIdx.x + j * gridDim.x);
            RingCursor cur{0};  // only the input-gradient item's offset is needed here
            VS_TR_DECL;
            uint32_t n_dw = 0, par_x = 0;
            for (int64_t k = 0; k < my_tiles; ++k) {
                const int64_t tile = blockIdx.x + k * gridDim.x;
                const bool full = (tile + 1) * kTileM <= n;
                // the ring holds less than one tile: the DRAM latency of the saved operands (TMA items) is taken off the critical
                // path by pulling whole tile images into L2 `pf_depth` tiles ahead
                if (pf_depth > 0 && k + pf_depth < my_tiles) prefetch_tile(tile + pf_depth * gridDim.x);
                cur.alloc(p2.item_bytes[0], R);
                for (int l = L - 1; l >= 0; --l) {
                    cur.alloc(st.a_bytes[l], R);
                    if (l >= 1) cur.alloc(kTileM * cfg.n_pad[l - 1] * 2, R);
                    // once the dW / db GEMMs of layer l are done, dZ_l and A_l leave the ring.  Four barriers in turn: this lane would
                    // have to fall four commits (a whole tile of operands, more than the ring holds) behind to miss a phase.
                    VS_TR(2, 10 + l);
                    mbar_wait(&bar_dw[n_dw & 3u], (n_dw >> 2) & 1u);
                    VS_TR(2, 20 + l);
                    ++n_dw;
                    release(2);
                    produce();
                    VS_TR(2, 30 + l);
                    if (l == 0 && p2.want_dx) {
                        const int off_x = cur.alloc(p2.item_bytes[n_items - 1], R);
                        mbar_wait(&bar_x, par_x);  // every epilogue warp has staged its part of the input gradient
                        par_x ^= 1;
                        if (full) {
                            bulk_s2g(d_pos + tile * kTileM * F, s_ring + off_x, (uint32_t)(kTileM * F * 4));
                            bulk_commit();
                            bulk_wait_read_all();
                        }
                        release(1);
                        produce();
                        VS_TR(2, 40);
                    }
                }
            }
            VS_TR_END(2);
        }
    } else if (warp == kMlpThreads / 32) {
        // ================= MMA issuer warp =================
        if (lane == 0 && my_tiles > 0) {
            mbar_wait(&bar_w, 0);
            RingCursor cur{0};
            uint32_t seq = 0;  // only its low bits matter (barrier slot and phase parity)
            uint32_t par_ready = 0, n_act = 0, n_dw = 0, fill_parity = 0;
            VS_TR_DECL;
            const uint32_t blob_addr = smem_u32(s_blob), ring_addr = smem_u32(s_ring), ones_addr = smem_u32(s_ones);
            for (int64_t k = 0; k < my_tiles; ++k) {
                int off_dz = cur.alloc(p2.item_bytes[0], R);
                const uint32_t seq0 = seq++;
                for (int l = L - 1; l >= 0; --l) {
                    const int K = cfg.k_pad[l], N = cfg.n_pad[l];
                    const int off_a = cur.alloc(st.a_bytes[l], R);
                    const uint32_t seq_a = seq++;
                    int off_next = 0;
                    if (l >= 1) {
                        off_next = cur.alloc(kTileM * cfg.n_pad[l - 1] * 2, R);
                        ++seq;
                    }
                    const bool want_dx = (l == 0) && p2.want_dx;
                    const uint32_t a_addr = ring_addr + (uint32_t)off_a, dz_addr = ring_addr + (uint32_t)off_dz;
                    VS_TR(0, 10 + l);
                    mbar_wait(&bar_ready, par_ready);  // dZ_l is complete (and the work columns have been read)
                    VS_TR(0, 20 + l);
                    par_ready ^= 1;
                    if (l == L - 1) {  // dZ of the output layer is there (a fill slot only sees the item-0 arrivals: its parity is kept here)
                        const uint32_t s0 = seq0 & (kItemBars - 1);
                        mbar_wait(&bar_fill[s0], (fill_parity >> s0) & 1u);
                        fill_parity ^= 1u << s0;
                    }
                    tc_fence_after();
                    if (l >= 1 || want_dx) {  // dA_l = dZ_l W_l: A = dZ_l K-major (K = fan-out), B = W_l read MN-major from the blob
                        umma_gemm_f16(tmem_base, dz_addr, kChunkBytes, 128, 2 * kChunkBytes, blob_addr + (uint32_t)cfg.w_off[l], 128u,
                                      (uint32_t)N * 16, 256, umma_idesc_f16(kTileM, K) | kIdescBMn, N / 16, false);
                        tc_commit(&bar_da);
                    }
                    VS_TR(0, 30 + l);
                    if (l >= 1) {
                        // the epilogue warps have turned the landed Z_{l-1} into A_l = act(Z_{l-1}) in place
                        mbar_wait(&bar_act[n_act & 1u], (n_act >> 1) & 1u);
                        ++n_act;
                        tc_fence_after();
                    } else {
                        mbar_wait(&bar_item[seq_a & (kItemBars - 1)], (seq_a / kItemBars) & 1u);  // A_0 has landed
                    }
                    VS_TR(0, 40 + l);
                    umma_gemm_f16(tmem_base + (uint32_t)plan.dw_col[l], a_addr, mn_lbo, mn_sbo, 256, dz_addr, mn_lbo, mn_sbo, 256,
                                  umma_idesc_f16(kTileM, N) | kIdescAMn | kIdescBMn, kTileM / 16, k > 0);
                    if (!plan.fold_bias[l])
                        umma_gemm_f16(tmem_base + (uint32_t)plan.db_col[l], dz_addr, mn_lbo, mn_sbo, 256, ones_addr, mn_lbo, mn_sbo, 256,
                                      umma_idesc_f16(kTileM, 16) | kIdescAMn | kIdescBMn, kTileM / 16, k > 0);
                    tc_commit(&bar_dw[n_dw & 3u]);  // the ring warp frees dZ_l and A_l when these GEMMs have read them
                    ++n_dw;
                    VS_TR(0, 50 + l);
                    if (want_dx) {
                        cur.alloc(p2.item_bytes[p2.n_items - 1], R);
                        ++seq;
                    }
                    off_dz = off_next;
                }
            }
            VS_TR_END(0);
        }
    } else {
        // ================= epilogue warps =================
        const uint32_t tmem_work = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        RingCursor cur{0};
        uint32_t seq = 0;  // only its low bits matter (barrier slot and phase parity)
        uint32_t par_da = 0;
        uint32_t n_act = 0;  // conversions this warp has announced on bar_act
        VS_TR_DECL;
        auto announce = [&](uint64_t* bar) {
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar);
        };
        for (int64_t k = 0; k < my_tiles; ++k) {
            const int64_t tile = blockIdx.x + k * gridDim.x;
            const int64_t row0 = tile * kTileM;
            const int rows = (int)min((int64_t)kTileM, n - row0);
            const bool full = rows == kTileM;
            const int64_t r = row0 + row;
            const bool live = row < rows;

            VS_TR_E(1);
            // item 0 (dZ of the output layer) is filled by the output-gradient warp; this announcement only says that the warp has read
            // the previous tile's last accumulator out of the work columns
            cur.alloc(p2.item_bytes[0], R);
            ++seq;
            announce(&bar_ready);

            for (int l = L - 1; l >= 0; --l) {
                const int K = cfg.k_pad[l];
                const int off_a = cur.alloc(st.a_bytes[l], R);  // item l: A_0 (consumed by the tensor core only) or Z_{l-1}
                const uint32_t seq_a = seq++;
                if (l >= 1) {
                    // Z_{l-1} (fp16 pre-activations, landed by TMA) -> A_l = act(Z_{l-1}) in place (the dW_l GEMM's operand) and this
                    // thread's elements of act'(Z_{l-1}) in registers, while the tensor core computes dA_l
                    VS_TR_E(10 + l);
                    mbar_wait(item_bar(seq_a), item_parity(seq_a));
                    VS_TR_E(20 + l);
                    // conversions are announced on two barriers in turn: nothing orders a warp's next conversion after the completion of
                    // the phase it arrived on last (a fast warp would arrive twice in one phase of a single barrier and release the dW GEMM
                    // before a slow warp has converted its part), but its conversion after that lies behind a dA GEMM that every warp's
                    // previous announcement precedes
                    VS_TR_E(30 + l);
                    uint4 gq[2][2];
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        const int c0 = cg * 16 + 64 * it;
                        if (c0 < K) {
                            uint4* zp = reinterpret_cast<uint4*>(s_ring + off_a + ((size_t)(c0 / 8) * kTileM + row) * 16);
#pragma unroll
                            for (int half = 0; half < 2; ++half) {
                                uint4 zq = zp[half * kTileM];
                                __half2* zh = reinterpret_cast<__half2*>(&zq);
                                __half2* gh = reinterpret_cast<__half2*>(&gq[it][half]);
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    __half2 a;
                                    if (ACT == 1) gelu_h2<true>(zh[q], a, gh[q]);
                                    else relu_h2<true>(zh[q], a, gh[q]);
                                    zh[q] = a;
                                }
                                zp[half * kTileM] = zq;
                            }
                        }
                    }
                    announce(&bar_act[n_act & 1u]);
                    ++n_act;
                    VS_TR_E(40 + l);
                    const int off_next = cur.alloc(kTileM * cfg.n_pad[l - 1] * 2, R);
                    mbar_wait(item_bar(seq), item_parity(seq));
                    VS_TR_E(50 + l);
                    ++seq;
                    mbar_wait(&bar_da, par_da);
                    VS_TR_E(60 + l);
                    par_da ^= 1;
                    tc_fence_after();
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        const int c0 = cg * 16 + 64 * it;
                        if (c0 < K) {
                            float v[16];
                            tmem_ld16(tmem_work + (uint32_t)c0, v);
                            const __half2* gl = reinterpret_cast<const __half2*>(&gq[it][0]);
                            const __half2* gh = reinterpret_cast<const __half2*>(&gq[it][1]);
                            __half2 h[8];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {  // dZ = fp16(dA) * act' in packed half precision (dZ is an fp16 operand anyway)
                                h[q] = __hmul2(floats2half2_sat(v[2 * q], v[2 * q + 1]), gl[q]);
                                h[4 + q] = __hmul2(floats2half2_sat(v[8 + 2 * q], v[8 + 2 * q + 1]), gh[q]);
                            }
                            uint4* slot = reinterpret_cast<uint4*>(s_ring + off_next + ((size_t)(c0 / 8) * kTileM + row) * 16);
                            slot[0] = *reinterpret_cast<const uint4*>(&h[0]);
                            slot[kTileM] = *reinterpret_cast<const uint4*>(&h[4]);
                        }
                    }
                    announce(&bar_ready);
                    VS_TR_E(70 + l);
                } else if (p2.want_dx) {
                    const int off_x = cur.alloc(p2.item_bytes[p2.n_items - 1], R);
                    VS_TR_E(80);
                    mbar_wait(item_bar(seq), item_parity(seq));
                    VS_TR_E(81);
                    ++seq;
                    float* xs = reinterpret_cast<float*>(s_ring + off_x);
                    mbar_wait(&bar_da, par_da);
                    VS_TR_E(82);
                    par_da ^= 1;
                    tc_fence_after();
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        const int c0 = cg * 16 + 64 * it;
                        if (c0 < F) {
                            float v[16];
                            tmem_ld16(tmem_work + (uint32_t)c0, v);
                            if (full) {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    if (c0 + j < F) xs[row * F + c0 + j] = v[j] * inv_scale;
                            } else if (live) {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    if (c0 + j < F) d_pos[r * F + c0 + j] = v[j] * inv_scale;
                            }
                        }
                    }
                    announce(&bar_x);
                    VS_TR_E(83);
                }
            }
        }
        if (tid == 0) VS_TR_END(1);
    }

    // ---- this CTA's parameter-gradient accumulators -> its slice of the workspace ---------------------------------------
    tc_fence_before();
    __syncthreads();
    if (my_tiles > 0 && warp < kMlpThreads / 32) {
        tc_fence_after();
        const uint32_t tmem_work = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        float* mine = partials + (size_t)blockIdx.x * plan.n_params;
        for (int l = 0; l < L; ++l) {
            const int N = cfg.n_pad[l], Kt = plan.dims[l], Nt = plan.dims[l + 1];
            for (int c0 = cg * 16; c0 < N; c0 += 64) {
                float v[16];
                tmem_ld16(tmem_work + (uint32_t)(plan.dw_col[l] + c0), v);
                if (row < Kt) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < Nt) mine[plan.p_off_w[l] + (c0 + j) * Kt + row] = v[j];
                } else if (plan.fold_bias[l] && row == cfg.k_pad[l]) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < Nt) mine[plan.p_off_b[l] + c0 + j] = v[j];
                }
            }
            if (!plan.fold_bias[l] && cg == 0) {
                float v[16];
                tmem_ld16(tmem_work + (uint32_t)plan.db_col[l], v);
                if (row < Nt) mine[plan.p_off_b[l] + row] = v[0];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)plan.tmem_cols);
}

}  // namespace vs

using namespace vs;

extern "C" {

#ifdef VS_KERNEL_TRACE
// copies the trace out (out_host: 3 * 2 * 256 int64, n_host: 3 counters)
int vs_debug_trace(long long* out_host, int* n_host) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out_host, g_trace, sizeof(g_trace));
    cudaMemcpyFromSymbol(n_host, g_trace_n, sizeof(g_trace_n));
    return 0;
}
#endif

static int bwd_setup(int n_layers, const int* dims, int pos_dim, int sh_degree, int normal_dep, MlpConfig* c, MlpBwdPlan* p,
                     bool check_recompute_smem = true) {
    if (!dims || pos_dim < 0 || sh_degree > 3) return VS_ERR_INVALID_ARG;
    int e = mlp_layout(n_layers, dims, c);
    if (e != VS_OK) return e;
    const int n_sh = sh_degree < 0 ? 0 : (sh_degree + 1) * (sh_degree + 1);
    if (dims[0] != pos_dim + n_sh + 3 * (normal_dep ? 1 : 0)) return VS_ERR_INVALID_ARG;
    c->pos_dim = pos_dim;
    c->n_sh = n_sh;
    c->normal_dep = normal_dep ? 1 : 0;
    e = mlp_bwd_plan(*c, dims, pos_dim, p);
    if (e != VS_OK) return e;
    if (check_recompute_smem && p->smem_bytes > 227 * 1024) return VS_ERR_UNSUPPORTED;
    return VS_OK;
}

static int bwd_grid(int64_t n_samples) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (int)std::min<int64_t>(std::max<int64_t>(div_up(n_samples, kTileM), 1), sms);
}

// number of fp32 parameter gradients: sum_l dims[l]*dims[l+1] + dims[l+1], laid out [W_0 | b_0 | W_1 | b_1 | ...]; < 0 on error
int64_t vs_mlp_num_params(int n_layers, const int* dims) {
    if (!dims || n_layers < 1 || n_layers > kMaxLayers) return VS_ERR_INVALID_ARG;
    int64_t n = 0;
    for (int l = 0; l < n_layers; ++l) n += (int64_t)dims[l] * dims[l + 1] + dims[l + 1];
    return n;
}

// bytes of device scratch vs_mlp_backward needs for n_samples samples (per-CTA gradient slices + the loss-scale cell); < 0 on error
int64_t vs_mlp_backward_workspace_bytes(int n_layers, const int* dims, int pos_dim, int sh_degree, int normal_dep, int64_t n_samples) {
    MlpConfig c;
    MlpBwdPlan p;
    int e = bwd_setup(n_layers, dims, pos_dim, sh_degree, normal_dep, &c, &p, false);
    if (e != VS_OK) return e;
    return 256 + (int64_t)bwd_grid(n_samples) * p.n_params * 4;
}

// Backward of vs_mlp_forward for the same arguments.  d_out: [n_samples, out_dim] upstream gradient.  d_pos: [n_samples, pos_dim]
// gradient of the positional features (NULL: not needed).  d_params: flat fp32 [vs_mlp_num_params] (overwritten, or added to
// when accumulate != 0).  workspace: vs_mlp_backward_workspace_bytes bytes, 16-byte aligned.  The alpha decay, the SH features and
// the normals carry no gradient (the reference evaluates them under no_grad, rgb.py:123-124, volsurfs.py:583-594).
int vs_mlp_backward(int n_layers, const int* dims, const void* blob, int pos_dim, int sh_degree, int normal_dep, int activation,
                    int alpha_decay, const float* pos, const float* dirs, const float* normals, const float* d_out, float* d_pos,
                    float* d_params, int accumulate, void* workspace, int64_t n_samples, const int64_t* n_valid_dev, int variant,
                    void* stream) {
    VS_CHECK_ARG(blob && n_samples >= 0 && d_params && workspace);
    MlpConfig c;
    MlpBwdPlan p;
    int e = bwd_setup(n_layers, dims, pos_dim, sh_degree, normal_dep, &c, &p);
    if (e != VS_OK) return e;
    if (c.out_dim > 8) return VS_ERR_UNSUPPORTED;  // sigmoid heads only; linear outputs go through vs_mlp_backward_stashed_raw
    VS_CHECK_ARG((c.n_sh == 0 || dirs) && (!(normal_dep || alpha_decay) || normals) && (!alpha_decay || dirs));
    VS_CHECK_ARG(pos_dim == 0 || pos);
    VS_CHECK_ARG(n_samples == 0 || d_out);
    VS_CHECK_ARG((reinterpret_cast<uintptr_t>(blob) & 15) == 0 && (reinterpret_cast<uintptr_t>(pos) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(d_pos) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0);
    c.activation = activation;
    c.alpha_decay = alpha_decay ? 1 : 0;
    c.variant = variant;
    cudaStream_t st = (cudaStream_t)stream;
    float* absmax = reinterpret_cast<float*>(workspace);
    float* partials = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + 256);
    const int grid = bwd_grid(n_samples);
    cudaError_t ce = cudaMemsetAsync(absmax, 0, 256, st);
    if (ce != cudaSuccess) return (int)ce;
    int launches = 0;
    if (n_samples > 0) {
        mlp_absmax_kernel<<<296, 256, 0, st>>>(d_out, n_samples, c.out_dim, n_valid_dev, absmax);
        auto kern = activation == 1 ? mlp_bwd_kernel<1> : mlp_bwd_kernel<0>;
        ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_bytes);
        if (ce != cudaSuccess) return (int)ce;
        kern<<<grid, kMlpThreads, p.smem_bytes, st>>>(c, p, reinterpret_cast<const uint8_t*>(blob), pos, dirs, normals, d_out, absmax, d_pos,
                                                      partials, n_samples, n_valid_dev);
        launches += 2;
    }
    mlp_bwd_reduce_kernel<<<(p.n_params + 255) / 256, 256, 0, st>>>(partials, n_samples > 0 ? grid : 0, p.n_params, absmax, n_valid_dev, n_samples,
                                                                    d_params, accumulate);
    return launched(launches + 1);
}

}  // extern "C"

// Backward of a training-mode vs_mlp_forward from its activation stash (no GEMM is recomputed; act / act' are re-evaluated from the stashed
// pre-activations: `activation` must be the forward's).  fwd_out: the [n_samples,out] output
// that forward wrote; other arguments as vs_mlp_backward.  workspace: vs_mlp_backward_workspace_bytes bytes.
static int mlp_backward_stashed_impl(int n_layers, const int* dims, const void* blob, const void* stash, int pos_dim, int sh_degree,
                                     int normal_dep, int activation, int alpha_decay, int out_linear, const float* dirs, const float* normals,
                                     const float* fwd_out, const float* d_out, float* d_pos, float* d_params, int accumulate,
                                     void* workspace, int64_t n_samples, const int64_t* n_valid_dev, void* stream) {
    VS_CHECK_ARG(blob && stash && n_samples >= 0 && d_params && workspace);
    MlpConfig c;
    MlpBwdPlan p;
    int e = bwd_setup(n_layers, dims, pos_dim, sh_degree, normal_dep, &c, &p, false);
    if (e != VS_OK) return e;
    if (!out_linear && c.out_dim > 8) return VS_ERR_UNSUPPORTED;
    c.out_linear = out_linear ? 1 : 0;
    VS_CHECK_ARG(!alpha_decay || (dirs && normals));
    VS_CHECK_ARG(n_samples == 0 || (d_out && (fwd_out || out_linear)));
    VS_CHECK_ARG((reinterpret_cast<uintptr_t>(blob) & 15) == 0 && (reinterpret_cast<uintptr_t>(stash) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(d_pos) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0);
    c.alpha_decay = alpha_decay ? 1 : 0;
    c.activation = activation;
    MlpStash st;
    mlp_stash_layout(c, &st);
    MlpBwd2Plan p2;
    e = mlp_bwd2_plan(c, st, pos_dim, d_pos != nullptr && pos_dim > 0, &p2);
    if (e != VS_OK) return e;
    if (p2.smem_bytes > 227 * 1024) return VS_ERR_UNSUPPORTED;
    if (const char* env = std::getenv("VS_MLP_BWD_PREFETCH")) p2.prefetch_tiles = std::max(0, std::atoi(env));  // A/B knob
    cudaStream_t s = (cudaStream_t)stream;
    float* absmax = reinterpret_cast<float*>(workspace);
    float* partials = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + 256);
    const int grid = bwd_grid(n_samples);
    cudaError_t ce = cudaMemsetAsync(absmax, 0, 256, s);
    if (ce != cudaSuccess) return (int)ce;
    int launches = 0;
    if (n_samples > 0) {
        mlp_absmax_kernel<<<296, 256, 0, s>>>(d_out, n_samples, c.out_dim, n_valid_dev, absmax);
        auto kern = activation == 1 ? mlp_bwd_stashed_kernel<1> : mlp_bwd_stashed_kernel<0>;
        ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p2.smem_bytes);
        if (ce != cudaSuccess) return (int)ce;
        kern<<<grid, kBwdThreads, p2.smem_bytes, s>>>(c, p, p2, st, reinterpret_cast<const uint8_t*>(blob),
                                                                        reinterpret_cast<const uint8_t*>(stash), dirs, normals, fwd_out, d_out,
                                                                        absmax, d_pos, partials, n_samples, n_valid_dev);
        launches += 2;
    }
    mlp_bwd_reduce_kernel<<<(p.n_params + 255) / 256, 256, 0, s>>>(partials, n_samples > 0 ? grid : 0, p.n_params, absmax, n_valid_dev, n_samples,
                                                                    d_params, accumulate);
    return launched(launches + 1);
}

extern "C" {

int vs_mlp_backward_stashed(int n_layers, const int* dims, const void* blob, const void* stash, int pos_dim, int sh_degree, int normal_dep,
                            int activation, int alpha_decay, const float* dirs, const float* normals, const float* fwd_out, const float* d_out, float* d_pos,
                            float* d_params, int accumulate, void* workspace, int64_t n_samples, const int64_t* n_valid_dev, void* stream) {
    return mlp_backward_stashed_impl(n_layers, dims, blob, stash, pos_dim, sh_degree, normal_dep, activation, alpha_decay, 0, dirs, normals, fwd_out, d_out,
                                     d_pos, d_params, accumulate, workspace, n_samples, n_valid_dev, stream);
}

// Backward of a training-mode vs_mlp_forward_raw (linear last layer): d_out [n_rows,out] -> d_in [n_rows,dims[0]] (or NULL) and the flat
// parameter gradients; workspace: vs_mlp_backward_workspace_bytes(n_layers, dims, dims[0], -1, 0, n_rows) bytes.
int vs_mlp_backward_stashed_raw(int n_layers, const int* dims, const void* blob, const void* stash, int activation, const float* d_out, float* d_in,
                                float* d_params, int accumulate, void* workspace, int64_t n_rows, const int64_t* n_valid_dev, void* stream) {
    VS_CHECK_ARG(dims);
    return mlp_backward_stashed_impl(n_layers, dims, blob, stash, dims[0], -1, 0, activation, 0, 1, nullptr, nullptr, nullptr, d_out, d_in, d_params,
                                     accumulate, workspace, n_rows, n_valid_dev, stream);
}

}  // extern "C"
