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
#include <cstdlib>

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


// One CTA keeps TWO 128-sample tiles in flight (slots 0 and 1, each with its own operand buffers and TMEM accumulator) and is
// warp-specialised:
//   * warps 0..15 (512 threads) are epilogue warps: they build the layer-0 operand and, per layer, read the accumulator back
//     (tcgen05.ld), apply bias + activation and write the next operand.  They never synchronise with each other: a warp that is
//     done announces it on the slot's `ready` mbarrier and moves on to the other slot;
//   * warp 16 is the control warp: one lane waits for `ready`, issues the layer's tcgen05.mma sequence, commits it to the slot's
//     `full` mbarrier (which the epilogue warps wait on) and requests the features of the slot's next tile by TMA as soon as the
//     landing zone is free.
// While the epilogue warps work on one slot the tensor core works on the other, so neither the MMA latency nor the hand-over is
// exposed.  STASH: training mode — the first layer's input operand and every hidden layer's fp16 pre-activations also go to global
// memory (16-byte stores, a warp covers 512 contiguous bytes) for the backward (layout: MlpStash).
constexpr int kFwdThreads = kMlpThreads + 32;

// Opt-in clock64 event trace of CTA 0's third and fourth tile pairs (scripts/trace_mlp_fwd.py, -DVS_KERNEL_TRACE; see mlp_bwd.cu)
#ifdef VS_KERNEL_TRACE
constexpr int kFwdTraceCap = 256;
__device__ long long g_trace_fwd[2][2 * kFwdTraceCap];
__device__ int g_trace_fwd_n[2];
#define VS_TRF_DECL int tr_n_ = 0, tr_k_ = 0
#define VS_TRF_NEXT ++tr_k_
#define VS_TRF(who, id)                                                                  \
    do {                                                                                 \
        if (blockIdx.x == 0 && tr_k_ >= 2 && tr_k_ < 4 && tr_n_ < kFwdTraceCap) {        \
            g_trace_fwd[who][2 * tr_n_] = (id);                                          \
            g_trace_fwd[who][2 * tr_n_ + 1] = clock64();                                 \
            ++tr_n_;                                                                     \
        }                                                                                \
    } while (0)
#define VS_TRF_END(who)                                  \
    do {                                                 \
        if (blockIdx.x == 0) g_trace_fwd_n[who] = tr_n_; \
    } while (0)
#else
#define VS_TRF_DECL ((void)0)
#define VS_TRF_NEXT ((void)0)
#define VS_TRF(who, id) ((void)0)
#define VS_TRF_END(who) ((void)0)
#endif
#define VS_TRF_E(id)                  \
    do {                              \
        if (tid == 0) VS_TRF(1, id);  \
    } while (0)

template <int ACT, bool STASH>  // ACT: 0 ReLU, 1 GELU (compile-time: the epilogue loop carries no branch)
__global__ void __launch_bounds__(kFwdThreads, 1) mlp_fwd_kernel(const __grid_constant__ MlpConfig cfg_param,
                                                                 const __grid_constant__ MlpStash stash_param, const uint8_t* __restrict__ blob,
                                                              const float* __restrict__ pos, const float* __restrict__ dirs,
                                                              const float* __restrict__ normals, float* __restrict__ out,
                                                              uint8_t* __restrict__ stash, int64_t n_samples,
                                                              const int64_t* __restrict__ n_valid_dev) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_w, bar_in[2], bar_full[2], bar_ready[2];
    __shared__ uint32_t tmem_slot;
    // the per-layer tables are indexed with the (run-time) layer number in the hot loops: from shared memory that is a ~30-cycle LDS; out
    // of the kernel-parameter bank it is a dependent constant load per use (measured: 14 % of all stall samples on one such line)
    __shared__ MlpConfig cfg;
    __shared__ MlpStash stash_cfg;
    for (int i = threadIdx.x; i < (int)(sizeof(MlpConfig) / 4); i += blockDim.x)
        reinterpret_cast<int*>(&cfg)[i] = reinterpret_cast<const int*>(&cfg_param)[i];
    for (int i = threadIdx.x; i < (int)(sizeof(MlpStash) / 4); i += blockDim.x)
        reinterpret_cast<int*>(&stash_cfg)[i] = reinterpret_cast<const int*>(&stash_param)[i];
    __syncthreads();

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int row = tid & (kTileM - 1);  // sample of the tile this thread works on (== its TMEM lane)
    const int cg = (tid >> 7) & 3;       // column group 0..3: which quarter of the columns this thread handles
    const int F = cfg.pos_dim;
    const int L = cfg.n_layers;
    const int k0 = cfg.k_pad[0];

    // carve shared memory: blob | per slot: A0 (layer-0 operand) | H (hidden activations; also the TMA landing zone of the fp32
    // features, which are consumed before the first hidden activation is written) | extras (per-row scratch of the row's owner)
    const int a0_bytes = kTileM * k0 * 2;
    const int h_bytes = cfg.h_bytes;
    const int slot_bytes = a0_bytes + h_bytes + kTileM * kExtraStride * 4;
    uint8_t* s_blob = smem;
    uint8_t* s_slots = s_blob + cfg.blob_bytes;

    int64_t n = n_samples;
    if (n_valid_dev != nullptr) n = min(n, *n_valid_dev);
    const int64_t n_tiles = (n + kTileM - 1) / kTileM;

    if (tid == 0) {
        mbar_init(&bar_w, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bar_in[s], 1);
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_ready[s], kMlpThreads / 32);  // one arrival per epilogue warp
        }
    }
    if (warp == 0) tmem_alloc(&tmem_slot, (uint32_t)(cfg.n_slots * cfg.tmem_cols));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    auto slot_a0 = [&](int s) { return reinterpret_cast<__half*>(s_slots + (size_t)s * slot_bytes); };
    auto slot_h = [&](int s) { return reinterpret_cast<__half*>(s_slots + (size_t)s * slot_bytes + a0_bytes); };
    auto slot_extra = [&](int s) { return reinterpret_cast<float*>(s_slots + (size_t)s * slot_bytes + a0_bytes + h_bytes); };

    const int64_t stride = gridDim.x;
    const bool two = cfg.n_slots == 2;
    const int64_t step = two ? 2 * stride : stride;

    if (warp == kMlpThreads / 32) {
        // ================= control warp =================
        if (lane == 0 && (int64_t)blockIdx.x < n_tiles) {
            mbar_arrive_expect_tx(&bar_w, (uint32_t)cfg.blob_bytes);  // weights + biases: resident for the whole kernel
            bulk_g2s(s_blob, blob, (uint32_t)cfg.blob_bytes, &bar_w);
            // features of a full tile -> the slot's landing zone (one TMA bulk copy); partial tiles are read from global memory
            auto request_features = [&](int s, int64_t tile) {
                if (tile < n_tiles && (tile + 1) * kTileM <= n && F > 0) {
                    mbar_arrive_expect_tx(&bar_in[s], (uint32_t)(kTileM * F * 4));
                    bulk_g2s(slot_h(s), pos + tile * kTileM * F, (uint32_t)(kTileM * F * 4), &bar_in[s]);
                    // the landing zone is the slot's hidden-activation buffer, so this copy can only be requested one output epilogue
                    // ahead of its use: pull the slot's FOLLOWING tile into L2 now, a whole tile time ahead, and the copy above finds
                    // its lines there instead of in DRAM
                    if ((tile + step + 1) * kTileM <= n) bulk_prefetch_l2(pos + (tile + step) * kTileM * F, (uint32_t)(kTileM * F * 4));
                }
                // the rows' directions / normals are read with plain loads at the top of build_a0: a DRAM miss there (queued behind this
                // kernel's stash writes) is 2 - 3 k cycles on the tile's critical path; from L2 it is a few hundred
                if (tile < n_tiles && (tile + 1) * kTileM <= n) {
                    const float* dp = dirs + tile * kTileM * 3;
                    const float* np = normals + tile * kTileM * 3;
                    if (dirs != nullptr && (reinterpret_cast<uintptr_t>(dp) & 15) == 0) bulk_prefetch_l2(dp, (uint32_t)(kTileM * 12));
                    if (normals != nullptr && (reinterpret_cast<uintptr_t>(np) & 15) == 0) bulk_prefetch_l2(np, (uint32_t)(kTileM * 12));
                }
            };
            int64_t t[2] = {(int64_t)blockIdx.x, two ? (int64_t)blockIdx.x + stride : n_tiles};
            request_features(0, t[0]);
            request_features(1, t[1]);
            mbar_wait(&bar_w, 0);
            uint32_t par_ready[2] = {0, 0}, par_full[2] = {0, 0};
            VS_TRF_DECL;
            while (t[0] < n_tiles) {
                for (int l = 0; l < L; ++l) {
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        if (t[s] >= n_tiles) continue;
                        VS_TRF(0, 100 + 10 * l + s);
                        mbar_wait(&bar_ready[s], par_ready[s]);  // every epilogue warp has written its part of the operand
                        VS_TRF(0, 200 + 10 * l + s);
                        par_ready[s] ^= 1;
                        tc_fence_after();
                        const int K = cfg.k_pad[l], N = cfg.n_pad[l];
                        const uint32_t a_base = smem_u32(l == 0 ? slot_a0(s) : slot_h(s));
                        const uint32_t b_base = smem_u32(s_blob + cfg.w_off[l]);
                        const uint32_t a_lbo = kTileM * 16, b_lbo = (uint32_t)N * 16;
                        umma_gemm_f16(tmem_base + (uint32_t)(s * cfg.tmem_cols), a_base, a_lbo, 128, 2 * a_lbo, b_base, b_lbo, 128, 2 * b_lbo,
                                      umma_idesc_f16(kTileM, N), K / 16, false);
                        tc_commit(&bar_full[s]);  // arrives when the MMAs above have completed
                        VS_TRF(0, 300 + 10 * l + s);
                        if (l == L - 1) {
                            // the (short) output-layer GEMM was the last reader of H: fetch the slot's next features under the
                            // other slot's epilogue
                            mbar_wait(&bar_full[s], par_full[s]);
                            request_features(s, t[s] + step);
                            VS_TRF(0, 400 + s);
                        }
                        par_full[s] ^= 1;
                    }
                }
                t[0] += step;
                if (two) t[1] += step;
                VS_TRF_NEXT;
            }
            VS_TRF_END(0);
        }
    } else {
        // ================= epilogue warps =================
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;  // a warp may only touch TMEM lanes 32*(warp%4)..+31
        uint32_t par_in[2] = {0, 0}, par_full[2] = {0, 0};
        // Slot s's row specials (SH / normal columns and alpha decay in build_a0, the sigmoid outputs in the last epilogue) belong to
        // column group s: the two slots' serial single-column-group stretches run side by side instead of queueing on column group 0.
        float decay[2] = {1.f, 1.f};  // alpha decay of this thread's row (held by the row's owner thread of the slot)
        VS_TRF_DECL;

        auto announce = [&](int s) {  // this warp's writes to the slot's next operand are done
            fence_proxy_async();      // generic-proxy writes -> visible to the tensor core (async proxy)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_ready[s]);
        };

        auto build_a0 = [&](int s, int64_t tile) {
            const int64_t row0 = tile * kTileM;
            const int rows = (int)min((int64_t)kTileM, n - row0);
            const bool full = rows == kTileM;
            const int64_t r = row0 + row;
            const bool live = row < rows;
            // directions / normals of the row come straight from global memory: in flight while the feature tile lands
            float dx = 0.f, dy = 0.f, dz = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
            const int q = (cg - s + 3) & 3;  // 3: the slot's owner column group; 0..2: the others in turn
            if (q == 3 && live) {
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
            VS_TRF_E(10 + s);
            if (full && F > 0) {
                mbar_wait(&bar_in[s], par_in[s]);
                par_in[s] ^= 1;
            }
            VS_TRF_E(20 + s);
            const float* srow = full ? reinterpret_cast<const float*>(slot_h(s)) + row * F : pos + r * F;
            __half* a0 = slot_a0(s);
            uint8_t* st_a0 = STASH ? stash + tile * (int64_t)stash_cfg.tile_bytes + stash_cfg.a_off[0] : nullptr;
            const int in_dim = cfg.in_dim;
            const int tail0 = F / 8;  // first chunk that holds SH / normal / padding columns
            float* ex = slot_extra(s) + row * kExtraStride;
            if (q == 3) {
                decay[s] = 1.f;
                if (cfg.alpha_decay) {
                    const float dot = fminf(fmaxf(-(dx * nx + dy * ny + dz * nz), 0.f), 1.f);
                    decay[s] = 2.f * sigmoid_f(10.f * dot) - 1.f;
                }
                float sh[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) sh[i] = 0.f;
                sh_eval(dx, dy, dz, cfg.n_sh, sh);
#pragma unroll
                for (int i = 0; i < 16; ++i) ex[i] = sh[i];  // private scratch row: written and read back by this thread only
                if (cfg.normal_dep) {                        // with n_sh < 16 the normal follows the SH block directly
                    ex[cfg.n_sh] = nx;
                    ex[cfg.n_sh + 1] = ny;
                    ex[cfg.n_sh + 2] = nz;
                }
            }
            // row -> fp16 A operand (K-major core matrices): chunk kc of row r at (kc*128 + r) * 16 bytes.  Pure feature chunks are
            // split over column groups 1..3 (no per-element decisions on a full tile) while column group 0 evaluates the row's SH /
            // normal columns into its scratch row; then the four warps that share these 32 rows meet on a named barrier and the tail
            // chunks (features' last columns, SH, normal, padding: a per-element select each) are split over all four column groups.
            // (One column group doing the whole tail was 3 - 4 k cycles per tile with the other twelve warps and the tensor core
            // waiting: clock64 trace, round 2.)
            auto tail_chunk = [&](int kc) {
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
                const uint4 pk = *reinterpret_cast<const uint4*>(h);
                *reinterpret_cast<uint4*>(a0 + ((size_t)kc * kTileM + row) * 8) = pk;
                if (STASH) *reinterpret_cast<uint4*>(st_a0 + ((size_t)kc * kTileM + row) * 16) = pk;
            };
            if (q != 3) {
                if (full) {
                    for (int kc = q; kc < tail0; kc += 3) {
                        const float* sp = srow + kc * 8;
                        __half2 h[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(sp[2 * j], sp[2 * j + 1]);
                        const uint4 pk = *reinterpret_cast<const uint4*>(h);
                        *reinterpret_cast<uint4*>(a0 + ((size_t)kc * kTileM + row) * 8) = pk;
                        if (STASH) *reinterpret_cast<uint4*>(st_a0 + ((size_t)kc * kTileM + row) * 16) = pk;
                    }
                } else {
                    for (int kc = q; kc < tail0; kc += 3) tail_chunk(kc);
                }
            }
            asm volatile("bar.sync %0, %1;" ::"r"(1 + (warp & 3)), "r"(4 * 32) : "memory");  // the row group's scratch rows are written
            for (int kc = tail0 + q; kc < k0 / 8; kc += 4) tail_chunk(kc);
            if (STASH && stash_cfg.fold[0] && cg == 1)  // the "ones" chunk behind A_0 (bias gradient row of the backward's dW GEMM)
                *reinterpret_cast<uint4*>(st_a0 + ((size_t)(k0 / 8) * kTileM + row) * 16) = make_uint4(0x00003C00u, 0u, 0u, 0u);
            announce(s);
            VS_TRF_E(30 + s);
        };

        // layer l of slot s has landed in TMEM: bias + activation -> next operand (hidden) or sigmoid -> out (last layer)
        auto epilogue = [&](int s, int l, int64_t tile) {
            VS_TRF_E(100 + 10 * l + s);
            mbar_wait(&bar_full[s], par_full[s]);
            VS_TRF_E(200 + 10 * l + s);
            par_full[s] ^= 1;
            tc_fence_after();
            const int N = cfg.n_pad[l];
            const uint32_t tmem_lane = tmem_base + lane_off + (uint32_t)(s * cfg.tmem_cols);
            const float* bias = reinterpret_cast<const float*>(s_blob + cfg.b_off[l]);
            const int64_t r = tile * kTileM + row;
            if (l + 1 < L) {
                uint8_t* hbuf = reinterpret_cast<uint8_t*>(slot_h(s));
                uint8_t* st_a = STASH ? stash + tile * (int64_t)stash_cfg.tile_bytes + stash_cfg.a_off[l + 1] : nullptr;
                if (STASH && stash_cfg.fold[l + 1] && cg == 1)
                    *reinterpret_cast<uint4*>(st_a + ((size_t)(N / 8) * kTileM + row) * 16) = make_uint4(0x00003C00u, 0u, 0u, 0u);
                for (int c0 = cg * 16; c0 < N; c0 += 64) {
                    float v[16];
                    tmem_ld16(tmem_lane + (uint32_t)c0, v);
                    const float4* b4 = reinterpret_cast<const float4*>(bias + c0);
                    __half2 zh[8], h[8];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 bb = b4[q];
                        zh[2 * q] = __floats2half2_rn(v[4 * q] + bb.x, v[4 * q + 1] + bb.y);
                        zh[2 * q + 1] = __floats2half2_rn(v[4 * q + 2] + bb.z, v[4 * q + 3] + bb.w);
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        __half2 g;
                        if (ACT == 1) gelu_h2<false>(zh[q], h[q], g);
                        else relu_h2<false>(zh[q], h[q], g);
                    }
                    const size_t off = ((size_t)(c0 / 8) * kTileM + row) * 16;  // next 8-column chunk: +128 rows * 16 bytes
                    *reinterpret_cast<uint4*>(hbuf + off) = *reinterpret_cast<const uint4*>(&h[0]);
                    *reinterpret_cast<uint4*>(hbuf + off + kTileM * 16) = *reinterpret_cast<const uint4*>(&h[4]);
                    if (STASH) {  // the backward re-evaluates act and act' from the fp16 pre-activation
                        *reinterpret_cast<uint4*>(st_a + off) = *reinterpret_cast<const uint4*>(&zh[0]);
                        *reinterpret_cast<uint4*>(st_a + off + kTileM * 16) = *reinterpret_cast<const uint4*>(&zh[4]);
                    }
                }
                announce(s);
                VS_TRF_E(300 + 10 * l + s);
            } else if (cfg.out_linear ? cg * 16 < N : cg == s) {  // output layer: N is 16 (sigmoid heads, <= 8 outputs: the slot's owner
                                                                    // column group) or up to 32 (linear texture nets: 16 columns per group)
                const int c0 = cfg.out_linear ? cg * 16 : 0;
                float v[16];
                tmem_ld16(tmem_lane + (uint32_t)c0, v);
                if (r < n) {
                    if (cfg.out_linear) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < cfg.out_dim) out[r * cfg.out_dim + c0 + j] = v[j] + bias[c0 + j];
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j < cfg.out_dim) out[r * cfg.out_dim + j] = sigmoid_f(v[j] + bias[j]) * decay[s];
                    }
                }
            }
        };

        int64_t tA = blockIdx.x, tB = two ? tA + stride : n_tiles;
        if (tA < n_tiles) mbar_wait(&bar_w, 0);  // biases live in the blob
        while (tA < n_tiles) {
            const bool hasB = tB < n_tiles;
            build_a0(0, tA);
            if (hasB) build_a0(1, tB);
            for (int l = 0; l < L; ++l) {
                epilogue(0, l, tA);
                if (hasB) epilogue(1, l, tB);
            }
            tA += step;
            if (two) tB += step;
            VS_TRF_E(900);
            VS_TRF_NEXT;
        }
        if (tid == 0) VS_TRF_END(1);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)(cfg.n_slots * cfg.tmem_cols));
}

static size_t mlp_smem_bytes(const MlpConfig& c) {
    const size_t slot = (size_t)kTileM * c.k_pad[0] * 2 + (size_t)c.h_bytes + (size_t)kTileM * kExtraStride * 4;
    return (size_t)c.blob_bytes + c.n_slots * slot + 128;
}

}  // namespace vs

using namespace vs;

extern "C" {

#ifdef VS_KERNEL_TRACE
int vs_debug_trace_fwd(long long* out_host, int* n_host) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out_host, g_trace_fwd, sizeof(g_trace_fwd));
    cudaMemcpyFromSymbol(n_host, g_trace_fwd_n, sizeof(g_trace_fwd_n));
    return 0;
}
#endif

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

// bytes of the activation stash a training-mode vs_mlp_forward writes for n_samples samples (0 samples -> 0); < 0 on error
int64_t vs_mlp_stash_bytes(int n_layers, const int* dims, int64_t n_samples) {
    MlpConfig c;
    if (!dims || n_samples < 0) return VS_ERR_INVALID_ARG;
    int e = mlp_layout(n_layers, dims, &c);
    if (e != VS_OK) return e;
    MlpStash s;
    mlp_stash_layout(c, &s);
    return div_up(n_samples, kTileM) * (int64_t)s.tile_bytes;
}

}  // extern "C"

// out[s, :out_dim] = sigmoid(MLP([pos[s] | SH_deg(dirs[s]) | normals[s]?])) (* alpha decay).  dims[0] must equal
// pos_dim + (sh_degree+1)^2 (0 if sh_degree < 0) + 3*normal_dep.  n_valid_dev (optional device int64) caps the sample count.
// stash (optional, vs_mlp_stash_bytes bytes, 16-byte aligned): training mode, activations are kept for vs_mlp_backward_stashed.
static int mlp_forward_impl(int n_layers, const int* dims, const void* blob, int pos_dim, int sh_degree, int normal_dep, int activation,
                            int alpha_decay, int out_linear, const float* pos, const float* dirs, const float* normals, float* out,
                            void* stash, int64_t n_samples, const int64_t* n_valid_dev, int variant, void* stream) {
    VS_CHECK_ARG(dims && blob && n_samples >= 0 && pos_dim >= 0 && sh_degree <= 3);
    MlpConfig c;
    int e = mlp_layout(n_layers, dims, &c);
    if (e != VS_OK) return e;
    if (!out_linear && c.out_dim > 8) return VS_ERR_UNSUPPORTED;
    c.out_linear = out_linear ? 1 : 0;
    const int n_sh = sh_degree < 0 ? 0 : (sh_degree + 1) * (sh_degree + 1);
    VS_CHECK_ARG(dims[0] == pos_dim + n_sh + 3 * (normal_dep ? 1 : 0));
    VS_CHECK_ARG((n_sh == 0 || dirs) && (!(normal_dep || alpha_decay) || normals) && (!alpha_decay || dirs));
    VS_CHECK_ARG(pos_dim == 0 || pos);
    if (n_samples == 0) return VS_OK;
    VS_CHECK_ARG(out);
    VS_CHECK_ARG((reinterpret_cast<uintptr_t>(blob) & 15) == 0 && (reinterpret_cast<uintptr_t>(pos) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(stash) & 15) == 0);
    c.pos_dim = pos_dim;
    c.n_sh = n_sh;
    c.normal_dep = normal_dep ? 1 : 0;
    c.activation = activation;
    c.alpha_decay = alpha_decay ? 1 : 0;
    c.variant = variant;
    c.h_bytes = (std::max(kTileM * c.a1_width * 2, kTileM * pos_dim * 4) + 127) & ~127;
    MlpStash sc;
    mlp_stash_layout(c, &sc);
    c.n_slots = 2;
    if (mlp_smem_bytes(c) > 227 * 1024) c.n_slots = 1;
    const size_t smem = mlp_smem_bytes(c);
    if (smem > 227 * 1024) return VS_ERR_UNSUPPORTED;
    void (*kern)(MlpConfig, MlpStash, const uint8_t*, const float*, const float*, const float*, float*, uint8_t*, int64_t, const int64_t*);
    if (stash)
        kern = activation == 1 ? mlp_fwd_kernel<1, true> : mlp_fwd_kernel<0, true>;
    else
        kern = activation == 1 ? mlp_fwd_kernel<1, false> : mlp_fwd_kernel<0, false>;
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return (int)ce;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t tiles = div_up(n_samples, kTileM);
    const unsigned grid = (unsigned)std::min<int64_t>(std::max<int64_t>((tiles + c.n_slots - 1) / c.n_slots, 1), (int64_t)sms);
    kern<<<grid, kFwdThreads, smem, (cudaStream_t)stream>>>(c, sc, reinterpret_cast<const uint8_t*>(blob), pos, dirs, normals, out,
                                                             reinterpret_cast<uint8_t*>(stash), n_samples, n_valid_dev);
    return launched(1);
}

extern "C" {

int vs_mlp_forward(int n_layers, const int* dims, const void* blob, int pos_dim, int sh_degree, int normal_dep, int activation,
                   int alpha_decay, const float* pos, const float* dirs, const float* normals, float* out, void* stash, int64_t n_samples,
                   const int64_t* n_valid_dev, int variant, void* stream) {
    return mlp_forward_impl(n_layers, dims, blob, pos_dim, sh_degree, normal_dep, activation, alpha_decay, 0, pos, dirs, normals, out, stash,
                            n_samples, n_valid_dev, variant, stream);
}

// out[r, :out] = MLP(in[r, :dims[0]]) with a LINEAR last layer (no sigmoid), out <= 32: the texture networks of the default appearance
// (tiny-cuda-nn FullyFusedMLP with "output_activation": "None", volsurfs_py/models/neural_texture.py:65-77).  stash as vs_mlp_forward.
int vs_mlp_forward_raw(int n_layers, const int* dims, const void* blob, int activation, const float* in, float* out, void* stash,
                       int64_t n_rows, const int64_t* n_valid_dev, void* stream) {
    VS_CHECK_ARG(dims);
    return mlp_forward_impl(n_layers, dims, blob, dims[0], -1, 0, activation, 0, 1, in, nullptr, nullptr, out, stash, n_rows, n_valid_dev, 0,
                            stream);
}

}  // extern "C"
