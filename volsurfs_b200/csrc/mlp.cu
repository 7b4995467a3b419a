// volsurfs_b200 — fused appearance head on the 5th-generation tensor cores (tcgen05 + TMEM), forward.
//
// Replaces, per layer hit, the reference's legacy RGB / alpha head
//   volsurfs_py/models/rgb.py:104-149   x = [pos_features | SH(dirs) | normals?] -> MLP -> sigmoid
//   volsurfs_py/models/mlp.py:8-52      Linear + exact-erf GELU per hidden layer, linear last layer (cuBLAS SGEMMs + ~20 elementwise kernels)
//   volsurfs_py/encodings/sphericalharmonics.py:84-153   hard-coded real SH polynomials, directions not normalised
//   volsurfs_py/methods/volsurfs.py:583-594              alpha *= 2*sigmoid(10*clamp(-d.n,0,1)) - 1
// with ONE kernel: a CTA owns tiles of 128 samples (= the 128 TMEM lanes).  Per tile
//   1. the [128, F] fp32 feature block arrives in shared memory by one TMA bulk copy (cp.async.bulk + mbarrier);
//   2. each thread converts its own row (features | SH evaluated in registers | normals, zero padded) to fp16 and writes it in
//      the UMMA canonical K-major layout (8x16-byte core matrices, no swizzle);
//   3. for every Linear layer one elected thread issues K/16 tcgen05.mma (M=128, N=layer width, fp16 x fp16 -> fp32 in TMEM),
//      commits them to an mbarrier;
//   4. all 4 warps read their 32 accumulator lanes back with tcgen05.ld, add the bias, apply GELU / ReLU and write the fp16
//      activations as the next layer's A operand — hidden activations never leave the SM;
//   5. the last layer's epilogue applies sigmoid (+ alpha decay) and stores fp32 outputs.
// All layers' fp16 weights (pre-packed into the UMMA layout by mlp_pack_kernel) stay resident in shared memory.
//
// Precision: fp16 operands, fp32 accumulation (the reference's default appearance path runs tiny-cuda-nn's fp16 FullyFusedMLP;
// the legacy torch path is fp32) — parity tolerance for this stage is stated in tests/test_gpu_mlp.py (abs 4e-3 on sigmoid outputs).
#include "mlp_common.cuh"

namespace vs {

// ---- packing: torch.nn.Linear weight [N,K] fp32 row-major -> fp16 [K_pad/8][N_pad][8] (UMMA K-major core matrices) ------
__global__ void mlp_pack_kernel(const float* __restrict__ W, const float* __restrict__ b, int N, int K, int n_pad, int k_pad,
                                __half* __restrict__ w_out, float* __restrict__ b_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = (k_pad / 8) * n_pad;
    if (i < total) {
        const int kc = i / n_pad, n = i % n_pad;
        __half v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = kc * 8 + j;
            v[j] = __float2half_rn((n < N && k < K) ? W[(int64_t)n * K + k] : 0.f);
        }
        *reinterpret_cast<uint4*>(w_out + (int64_t)i * 8) = *reinterpret_cast<const uint4*>(v);
    }
    if (i < n_pad) b_out[i] = (i < N && b != nullptr) ? b[i] : 0.f;
}


template <int ACT>  // 0 ReLU, 1 GELU: compile-time so the epilogue loop carries no branch
__global__ void __launch_bounds__(kMlpThreads) mlp_fwd_kernel(const MlpConfig cfg, const uint8_t* __restrict__ blob,
                                                              const float* __restrict__ pos, const float* __restrict__ dirs,
                                                              const float* __restrict__ normals, float* __restrict__ out,
                                                              int64_t n_samples, const int64_t* __restrict__ n_valid_dev) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_w, bar_in[2], bar_mma;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int row = tid & (kTileM - 1);  // sample of the tile this thread works on (== its TMEM lane)
    const int cg = tid >> 7;             // column group 0..3: which quarter of the columns / K-chunks this thread handles
    const int F = cfg.pos_dim;
    const int k0 = cfg.k_pad[0];

    // carve shared memory
    uint8_t* s_blob = smem;                                                         // blob_bytes (multiple of 16)
    __half* s_a0 = reinterpret_cast<__half*>(s_blob + cfg.blob_bytes);              // 128 x k0 fp16 (layer-0 operand)
    __half* s_a1 = s_a0 + kTileM * k0;                                              // 128 x a1_width fp16 (hidden activations)
    const int stage_floats = (kTileM * F + 3) & ~3;
    float* s_stage0 = reinterpret_cast<float*>(s_a1 + kTileM * cfg.a1_width);       // 128 x F fp32 (TMA landing zone) x 1 or 2
    float* s_extra = s_stage0 + (cfg.prefetch ? 2 : 1) * stage_floats;              // 128 x kExtraStride

    int64_t n = n_samples;
    if (n_valid_dev != nullptr) n = min(n, *n_valid_dev);
    const int64_t n_tiles = (n + kTileM - 1) / kTileM;

    if (tid == 0) {
        mbar_init(&bar_w, 1);
        mbar_init(&bar_in[0], 1);
        mbar_init(&bar_in[1], 1);
        mbar_init(&bar_mma, 1);
    }
    if (warp == 0) tmem_alloc(&tmem_slot, (uint32_t)cfg.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);  // a warp may only touch lanes 32*(warp%4)..+31

    if ((int64_t)blockIdx.x < n_tiles && tid == 0) {  // weights + biases: resident for the whole kernel
        mbar_arrive_expect_tx(&bar_w, (uint32_t)cfg.blob_bytes);
        bulk_g2s(s_blob, blob, (uint32_t)cfg.blob_bytes, &bar_w);
    }
    bool weights_ready = false;
    uint32_t par_in_bits = 0, par_mma = 0;  // bit b of par_in_bits: phase parity of bar_in[b]
    int buf = 0;
    bool have_prefetched = false;  // the current tile's features were requested during the previous iteration
    const uint32_t lbo_sel = (cfg.variant & 1);

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * kTileM;
        const int rows = (int)min((int64_t)kTileM, n - row0);
        const bool full = rows == kTileM;
        const int64_t r = row0 + row;
        const bool live = row < rows;

        // ---- 1. features -> shared memory
        float* s_stage = s_stage0 + buf * stage_floats;
        if (full) {
            if (tid == 0 && !have_prefetched) {
                mbar_arrive_expect_tx(&bar_in[buf], (uint32_t)(kTileM * F * 4));
                bulk_g2s(s_stage, pos + row0 * F, (uint32_t)(kTileM * F * 4), &bar_in[buf]);
            }
        } else {
            for (int e = tid; e < rows * F; e += kMlpThreads) s_stage[e] = __ldg(pos + row0 * F + e);
        }
        // prefetch the next tile of this CTA into the other buffer (its previous contents were consumed an iteration ago)
        bool next_prefetched = false;
        if (cfg.prefetch) {
            const int64_t nt = tile + gridDim.x;
            if (nt < n_tiles && (nt + 1) * kTileM <= n) {
                next_prefetched = true;
                if (tid == 0) {
                    mbar_arrive_expect_tx(&bar_in[buf ^ 1], (uint32_t)(kTileM * F * 4));
                    bulk_g2s(s_stage0 + (buf ^ 1) * stage_floats, pos + nt * kTileM * F, (uint32_t)(kTileM * F * 4), &bar_in[buf ^ 1]);
                }
            }
        }
        // ---- 2. per-row extras: SH(dir), normal
        float dx = 0.f, dy = 0.f, dz = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
        if (live) {
            if (dirs != nullptr) {
                dx = __ldg(dirs + 3 * r);
                dy = __ldg(dirs + 3 * r + 1);
                dz = __ldg(dirs + 3 * r + 2);
            }
            if (normals != nullptr) {
                nx = __ldg(normals + 3 * r);
                ny = __ldg(normals + 3 * r + 1);
                nz = __ldg(normals + 3 * r + 2);
            }
        }
        {
            float sh[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) sh[i] = 0.f;
            sh_eval(dx, dy, dz, cfg.n_sh, sh);
            if (cg == 0) {
                float* ex = s_extra + row * kExtraStride;
#pragma unroll
                for (int i = 0; i < 16; ++i) ex[i] = sh[i];
                // with n_sh < 16 the normal follows the SH block directly
                if (cfg.normal_dep) {
                    ex[cfg.n_sh] = nx;
                    ex[cfg.n_sh + 1] = ny;
                    ex[cfg.n_sh + 2] = nz;
                }
            }
        }
        __syncthreads();
        if (full) {
            mbar_wait(&bar_in[buf], (par_in_bits >> buf) & 1u);
            par_in_bits ^= (1u << buf);
        }
        // ---- 3. row -> fp16 A operand (K-major core matrices): chunk kc of row r at (kc*128 + r) * 16 bytes; the four column
        //         groups split the K-chunks of a row
        {
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

        // ---- 4. layers
        for (int l = 0; l < cfg.n_layers; ++l) {
            const int K = cfg.k_pad[l], N = cfg.n_pad[l];
            fence_proxy_async();  // this thread's operand writes -> visible to the tensor core (async proxy)
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                const uint32_t a_base = smem_u32(l == 0 ? s_a0 : s_a1);
                const uint32_t b_base = smem_u32(s_blob + cfg.w_off[l]);
                const uint32_t a_lbo = kTileM * 16, b_lbo = (uint32_t)N * 16, sbo = 128;
                const uint32_t idesc = umma_idesc_f16(kTileM, N);
                for (int ks = 0; ks < K / 16; ++ks) {
                    const uint32_t a_addr = a_base + (uint32_t)ks * 2 * a_lbo;
                    const uint32_t b_addr = b_base + (uint32_t)ks * 2 * b_lbo;
                    const uint64_t ad = lbo_sel ? umma_desc(a_addr, sbo, a_lbo) : umma_desc(a_addr, a_lbo, sbo);
                    const uint64_t bd = lbo_sel ? umma_desc(b_addr, sbo, b_lbo) : umma_desc(b_addr, b_lbo, sbo);
                    tc_mma_f16(tmem_base, ad, bd, idesc, ks > 0 ? 1u : 0u);
                }
                tc_commit(&bar_mma);  // arrives when the MMAs above have completed
            }
            mbar_wait(&bar_mma, par_mma);
            par_mma ^= 1;
            tc_fence_after();

            const float* bias = reinterpret_cast<const float*>(s_blob + cfg.b_off[l]);
            if (l + 1 < cfg.n_layers) {
                // hidden layer epilogue: bias + activation -> fp16 -> next A operand (row = tid)
                for (int c0 = cg * 16; c0 < N; c0 += 64) {
                    float v[16];
                    tmem_ld16(tmem_lane + (uint32_t)c0, v);
                    const float4* b4 = reinterpret_cast<const float4*>(bias + c0);
                    __half2 h[8];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 bb = b4[q];
                        float a0 = v[4 * q] + bb.x, a1 = v[4 * q + 1] + bb.y, a2 = v[4 * q + 2] + bb.z, a3 = v[4 * q + 3] + bb.w;
                        if (ACT == 1) {
                            a0 = gelu_erf(a0);
                            a1 = gelu_erf(a1);
                            a2 = gelu_erf(a2);
                            a3 = gelu_erf(a3);
                        } else {
                            a0 = fmaxf(a0, 0.f);
                            a1 = fmaxf(a1, 0.f);
                            a2 = fmaxf(a2, 0.f);
                            a3 = fmaxf(a3, 0.f);
                        }
                        h[2 * q] = __floats2half2_rn(a0, a1);
                        h[2 * q + 1] = __floats2half2_rn(a2, a3);
                    }
                    uint4* dst = reinterpret_cast<uint4*>(s_a1 + ((size_t)(c0 / 8) * kTileM + row) * 8);
                    dst[0] = *reinterpret_cast<const uint4*>(&h[0]);
                    dst[kTileM] = *reinterpret_cast<const uint4*>(&h[4]);  // next 8-column chunk: +128 rows * 16 bytes
                }
            } else if (cg == 0) {
                float v[16];
                tmem_ld16(tmem_lane, v);
                if (live) {
                    float decay = 1.f;
                    if (cfg.alpha_decay) {
                        const float dot = fminf(fmaxf(-(dx * nx + dy * ny + dz * nz), 0.f), 1.f);
                        decay = 2.f * sigmoid_f(10.f * dot) - 1.f;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j < cfg.out_dim) out[r * cfg.out_dim + j] = sigmoid_f(v[j] + bias[j]) * decay;
                }
            }
        }
        // the next tile overwrites s_stage / s_extra / s_a0 and TMEM: everyone must be done reading them
        tc_fence_before();
        __syncthreads();
        have_prefetched = next_prefetched;
        if (cfg.prefetch) buf ^= 1;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)cfg.tmem_cols);
}

static size_t mlp_smem_bytes(const MlpConfig& c) {
    size_t b = (size_t)c.blob_bytes;
    b += (size_t)kTileM * c.k_pad[0] * 2;
    b += (size_t)kTileM * c.a1_width * 2;
    b += (size_t)((kTileM * c.pos_dim + 3) & ~3) * 4 * (c.prefetch ? 2 : 1);
    b += (size_t)kTileM * kExtraStride * 4;
    return b + 128;
}

}  // namespace vs

using namespace vs;

extern "C" {

// bytes of the packed weight blob for an MLP with dims = [in, h1, ..., out] (n_layers + 1 entries); < 0 on error
int64_t vs_mlp_blob_bytes(int n_layers, const int* dims) {
    MlpConfig c;
    if (!dims) return VS_ERR_INVALID_ARG;
    int e = mlp_layout(n_layers, dims, &c);
    return e != VS_OK ? e : c.blob_bytes;
}

// Pack torch.nn.Linear parameters (weights[l]: fp32 [dims[l+1], dims[l]] row-major, biases[l]: fp32 [dims[l+1]] or NULL; DEVICE
// pointers listed in HOST arrays) into the fp16 tensor-core layout.  Call again whenever the parameters change.
int vs_mlp_pack(int n_layers, const int* dims, const float* const* weights, const float* const* biases, void* blob, void* stream) {
    VS_CHECK_ARG(dims && weights && biases && blob);
    MlpConfig c;
    int e = mlp_layout(n_layers, dims, &c);
    if (e != VS_OK) return e;
    for (int l = 0; l < n_layers; ++l) {
        VS_CHECK_ARG(weights[l]);
        const int total = (c.k_pad[l] / 8) * c.n_pad[l];
        const int threads = 128, blocks = (std::max(total, c.n_pad[l]) + threads - 1) / threads;
        mlp_pack_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(weights[l], biases[l], dims[l + 1], dims[l], c.n_pad[l], c.k_pad[l],
                                                                       reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(blob) + c.w_off[l]),
                                                                       reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(blob) + c.b_off[l]));
    }
    return launched(n_layers);
}

// out[s, :out_dim] = sigmoid(MLP([pos[s] | SH_deg(dirs[s]) | normals[s]?])) (* alpha decay).  dims[0] must equal
// pos_dim + (sh_degree+1)^2 (0 if sh_degree < 0) + 3*normal_dep.  n_valid_dev (optional device int64) caps the sample count.
int vs_mlp_forward(int n_layers, const int* dims, const void* blob, int pos_dim, int sh_degree, int normal_dep, int activation,
                   int alpha_decay, const float* pos, const float* dirs, const float* normals, float* out, int64_t n_samples,
                   const int64_t* n_valid_dev, int variant, void* stream) {
    VS_CHECK_ARG(dims && blob && n_samples >= 0 && pos_dim >= 0 && sh_degree <= 3);
    MlpConfig c;
    int e = mlp_layout(n_layers, dims, &c);
    if (e != VS_OK) return e;
    const int n_sh = sh_degree < 0 ? 0 : (sh_degree + 1) * (sh_degree + 1);
    VS_CHECK_ARG(dims[0] == pos_dim + n_sh + 3 * (normal_dep ? 1 : 0));
    VS_CHECK_ARG((n_sh == 0 || dirs) && (!(normal_dep || alpha_decay) || normals) && (!alpha_decay || dirs));
    VS_CHECK_ARG(pos_dim == 0 || pos);
    if (n_samples == 0) return VS_OK;
    VS_CHECK_ARG(out);
    VS_CHECK_ARG((reinterpret_cast<uintptr_t>(blob) & 15) == 0 && (reinterpret_cast<uintptr_t>(pos) & 15) == 0);
    c.pos_dim = pos_dim;
    c.n_sh = n_sh;
    c.normal_dep = normal_dep ? 1 : 0;
    c.activation = activation;
    c.alpha_decay = alpha_decay ? 1 : 0;
    c.variant = variant;
    size_t smem = mlp_smem_bytes(c);
    if (smem > 227 * 1024) return VS_ERR_UNSUPPORTED;
    if (smem > 113 * 1024) {  // one CTA per SM anyway: spend the spare shared memory on a second feature buffer
        c.prefetch = 1;
        if (mlp_smem_bytes(c) > 227 * 1024) c.prefetch = 0;
        smem = mlp_smem_bytes(c);
    }
    auto kern = activation == 1 ? mlp_fwd_kernel<1> : mlp_fwd_kernel<0>;
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return (int)ce;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ctas_per_sm = smem <= 75 * 1024 ? 3 : (smem <= 113 * 1024 ? 2 : 1);
    const int64_t tiles = div_up(n_samples, kTileM);
    const unsigned grid = (unsigned)std::min<int64_t>(tiles, (int64_t)sms * ctas_per_sm);
    kern<<<grid, kMlpThreads, smem, (cudaStream_t)stream>>>(c, reinterpret_cast<const uint8_t*>(blob), pos, dirs, normals, out,
                                                                     n_samples, n_valid_dev);
    return launched(1);
}

}  // extern "C"
