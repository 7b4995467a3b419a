// Permutohedral-lattice hash encoding, forward and backward.
//
// Replaces (reference, relative to /root/reference/submodules/permutohedral_encoding/):
//   forward_gpu              kernels/permutohedral_encoding/EncodingGPU.cuh:68-261   (host src/Encoding.cu:55-113)
//   backward_gpu             kernels/permutohedral_encoding/EncodingGPU.cuh:264-416  (host src/Encoding.cu:116-217)
//   backward_gpu_only_pos    kernels/permutohedral_encoding/EncodingGPU.cuh:534-700
// and the point normalisation + out-of-bounds mask of volsurfs_py/encodings/permutohash.py:77-86.
//
// Layout / mapping (B200): the reference launches one thread per (position, level) with the level in blockIdx.y and writes a
// level-major [levels, 2, N] tensor that Python then permutes into rows.  Here a CTA owns a tile of 128 consecutive positions for ALL
// levels: 256 threads = 128 positions x 2 level-halves, a warp = 32 consecutive positions of one level (same 2 MB table, coherent
// simplices on the coarse levels -> broadcast loads / aggregated atomics).  The tile is staged in shared memory (column-major, padded)
// and leaves as complete rows [N, out_cols] — exactly the operand layout of the appearance MLP (vs_mlp_forward `pos`), so no permute,
// slice or concat kernel runs in between.  The lattice tables (levels x 2 MB) live in the 126 MB L2: the kernel is bound by L2
// gathers (4 x 8 B per position and level), not by HBM.
//
// Arithmetic follows the reference expression by expression (same association, nvcc contracts the same a*b+c pairs), so the
// forward output agrees bit for bit with the reference kernels wherever the discrete simplex choice agrees.  The lattice
// gradient is accumulated with fp32 atomics (as in the reference: order, hence last-ulp rounding, is unspecified).
#include "volsurfs_b200.h"
#include <cstdlib>

#include "vs_common.cuh"

namespace vs {

constexpr int PM_TILE = 128;       // positions per CTA
constexpr int PM_THREADS = 256;    // 2 level-halves
constexpr int PM_PAD = PM_TILE + 1;
constexpr int PM_MAX_COLS = 72;    // 2 * (levels + extra) <= 72  (levels <= 32)

struct PermutoArgs {
    int n_levels, n_extra, concat_points, pow2;
    int agg_max_heads;   // red_add_runs aggregates runs of equal slots when a warp has at most this many runs
    uint32_t capacity;
    float points_scaling;
    int has_bb;
    float bb_half[8], bb_inv_half[8];   // permutohash.py:77-86
};

template <int D>
struct Simplex {
    int rem0[D + 1];
    int rank[D + 1];
    float bary[D + 1];   // weight of the vertex with remainder r
};

// EncodingGPU.cuh:128-205 (same in the three reference kernels)
template <int D>
__device__ __forceinline__ void locate(const float (&pos)[D], const float* __restrict__ shift, const float* __restrict__ scale, Simplex<D>& s) {
    float elevated[D + 1];
    float sm = 0;
#pragma unroll
    for (int i = D; i > 0; i--) {
        float cf = (pos[i - 1] + __ldg(shift + i - 1)) * __ldg(scale + i - 1);
        elevated[i] = sm - i * cf;
        sm += cf;
    }
    elevated[0] = sm;

    int sum = 0;
#pragma unroll
    for (int i = 0; i <= D; i++) {
        float v = elevated[i] * (1.0f / (D + 1));
        float up = ceilf(v) * (D + 1);
        float down = floorf(v) * (D + 1);
        s.rem0[i] = (up - elevated[i] < elevated[i] - down) ? (int)up : (int)down;
        sum += s.rem0[i];
        s.rank[i] = 0;
    }
    sum /= D + 1;

    float diff[D + 1];
#pragma unroll
    for (int i = 0; i <= D; i++) diff[i] = elevated[i] - s.rem0[i];
#pragma unroll
    for (int i = 0; i < D; i++) {
#pragma unroll
        for (int j = i + 1; j <= D; j++) {
            if (diff[i] < diff[j])
                s.rank[i]++;
            else
                s.rank[j]++;
        }
    }
#pragma unroll
    for (int i = 0; i <= D; i++) {
        s.rank[i] += sum;
        if (s.rank[i] < 0) {
            s.rank[i] += D + 1;
            s.rem0[i] += D + 1;
        } else if (s.rank[i] > D) {
            s.rank[i] -= D + 1;
            s.rem0[i] -= D + 1;
        }
    }
    // barycentric weights.  The reference scatters +delta_i into slot D-rank_i and -delta_i into slot D+1-rank_i of a zeroed array;
    // rank is a permutation, so slot t holds delta[rank == D-t] - delta[rank == D+1-t], one rounding whichever term came first.
    // (gathered per slot — "the delta whose rank is k" — not scattered per vertex: the scatter form is compiled into a dynamically indexed
    // local-memory array, 16 STL + 1 LDL per level in the SASS of round 1)
    float delta[D + 1], by_rank[D + 1];
#pragma unroll
    for (int i = 0; i <= D; i++) delta[i] = (elevated[i] - s.rem0[i]) * (1.0f / (D + 1));
#pragma unroll
    for (int k = 0; k <= D; k++) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i <= D; i++) v = (s.rank[i] == k) ? delta[i] : v;
        by_rank[k] = v;
    }
#pragma unroll
    for (int t = 1; t <= D; t++) s.bary[t] = by_rank[D - t] - by_rank[D + 1 - t];
    s.bary[0] = by_rank[D] + (1.0f + (0.0f - by_rank[0]));
}

// EncodingGPU.cuh:22-45,216-227: hash slot of the simplex vertex with this remainder.  The reference folds the D keys
// key_i = rem0_i + remainder - (rank_i > D - remainder ? D + 1 : 0) as k = (k + key_i) * M, i = 0..D-1: in uint32 ring arithmetic that is
// the LINEAR form  sum_i key_i M^(D-i)  =  [sum_i rem0_i M^(D-i)]  +  remainder [sum_i M^(D-i)]  -  (D+1) sum_i [rank_i + remainder > D] M^(D-i),
// bit-identical (wrap-around arithmetic is exact), with the first bracket computed once per simplex (HashBase) instead of once per vertex.
// a == k ? x : y as one setp + selp the compiler cannot re-interpret as an indexed array read
__device__ __forceinline__ float select_eq(int a, int k, float x, float y) {
    float r;
    asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, %2;\n\tselp.f32 %0, %3, %4, p;\n\t}" : "=f"(r) : "r"(a), "r"(k), "f"(x), "f"(y));
    return r;
}

template <int D>
struct HashBase {
    uint32_t base;
};
template <int D>
__device__ __forceinline__ constexpr uint32_t hash_pow(int e) {   // M^e mod 2^32
    uint32_t p = 1u;
    for (int i = 0; i < e; i++) p *= 2531011u;
    return p;
}
template <int D>
__device__ __forceinline__ HashBase<D> hash_base(const Simplex<D>& s) {
    HashBase<D> h;
    uint32_t k = 0;
#pragma unroll
    for (int i = 0; i < D; i++) {
        k += (uint32_t)s.rem0[i];
        k *= 2531011u;
    }
    h.base = k;
    return h;
}
// All D+1 slots of a simplex.  From remainder r-1 to r the sum gains sum_i M^(D-i) and loses (D+1) M^(D-i) for the ONE coordinate whose
// rank is D+1-r (rank is a permutation; nothing if that coordinate is the last one, which the hash does not read): two adds per vertex
// after a select chain over the ranks, instead of D compare-and-subtract steps per vertex.  Same uint32 ring arithmetic, same bits.
// POW2 is a template parameter: as a run-time flag both reductions were compiled in, and the general 32-bit modulo was a quarter of the
// forward kernel's instructions (23 per vertex, ncu source page, round 2) although every table of the benchmark is 2^18 entries.
template <int D, bool POW2>
__device__ __forceinline__ void vertex_slots(const Simplex<D>& s, const HashBase<D>& h, uint32_t capacity, uint32_t (&slot)[D + 1]) {
    uint32_t sum_pow = 0;
#pragma unroll
    for (int i = 0; i < D; i++) sum_pow += hash_pow<D>(D - i);
    uint32_t lose[D + 1];   // by rank
#pragma unroll
    for (int k = 0; k <= D; k++) {
        uint32_t c = 0u;
#pragma unroll
        for (int i = 0; i < D; i++) c = (s.rank[i] == k) ? (uint32_t)(D + 1) * hash_pow<D>(D - i) : c;
        lose[k] = c;
    }
    uint32_t k = h.base;
#pragma unroll
    for (int r = 0; r <= D; r++) {
        if (r > 0) k += sum_pow - lose[D + 1 - r];
        slot[r] = POW2 ? (k & (capacity - 1u)) : (k % capacity);
    }
}

template <int D>
__device__ __forceinline__ bool load_position(const PermutoArgs& a, const float* __restrict__ positions, int64_t idx, bool valid, float (&pos)[D]) {
    bool oob = false;
#pragma unroll
    for (int i = 0; i < D; i++) {
        float p = valid ? __ldg(positions + idx * D + i) : 0.f;
        if (a.has_bb) {
            oob = oob || (p <= -a.bb_half[i]) || (p >= a.bb_half[i]);
            p = __fdiv_rn(__fadd_rn(__fmul_rn(p, a.bb_inv_half[i]), 1.0f), 2.0f);   // torch: (points * scaling + 1) / 2, op by op
        }
        pos[i] = p;
    }
    return oob;
}

__device__ __forceinline__ void red_add_f32x2(float2* addr, float x, float y) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(x), "f"(y) : "memory");
}

// One (x, y) contribution per lane into table[slot].  Lanes of a warp are consecutive positions: on the coarse levels long runs of
// lanes hit the same vertex, and 32 same-address atomics serialise in L2 — so runs of equal slots are summed with shuffles first and
// the run's first lane issues one vector reduction.  When most lanes differ (fine levels) the atomics go out directly.
__device__ __forceinline__ void red_add_runs(float2* table, uint32_t slot, float x, float y, bool valid, int lane, int agg_max_heads) {
    uint32_t prev = __shfl_up_sync(VS_FULL_MASK, slot, 1);
    bool head = (lane == 0) || (prev != slot);
    uint32_t heads = __ballot_sync(VS_FULL_MASK, head);
    if (__popc(heads) > agg_max_heads) {
        if (valid) red_add_f32x2(table + slot, x, y);
        return;
    }
    uint32_t above = (lane == 31) ? 0u : (heads & ~((2u << lane) - 1u));
    int end = above ? (__ffs(above) - 1) : 32;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        float ox = __shfl_down_sync(VS_FULL_MASK, x, d);
        float oy = __shfl_down_sync(VS_FULL_MASK, y, d);
        if (lane + d < end) {
            x += ox;
            y += oy;
        }
    }
    if (head && valid) red_add_f32x2(table + slot, x, y);
}

template <int D, bool POW2>
__global__ void __launch_bounds__(PM_THREADS) permuto_fwd_kernel(PermutoArgs a, const float* __restrict__ positions,
                                                                 const float2* __restrict__ lattice, const float* __restrict__ scale,
                                                                 const float* __restrict__ shift, const float* __restrict__ window,
                                                                 float* __restrict__ out, int out_cols, int64_t out_stride,
                                                                 uint8_t* __restrict__ oob_out, int64_t n, const int64_t* __restrict__ n_valid_dev) {
    extern __shared__ __align__(128) float tile[];   // [cols][PM_PAD], or the tile's output rows as they lie in memory (bulk path)
    const int64_t n_eff = n_valid_dev ? min(n, *n_valid_dev) : n;
    const int64_t base = (int64_t)blockIdx.x * PM_TILE;
    if (base >= n_eff) return;
    const int p = threadIdx.x & (PM_TILE - 1), half = threadIdx.x >> 7;
    const int64_t idx = base + p;
    const bool valid = idx < n_eff;
    // A full tile of densely packed rows leaves as ONE bulk copy: the threads write their columns where the row-major image wants them
    // (row stride out_cols: odd strides are bank-conflict free) instead of a transposed tile that a second loop walks out to memory.
    const bool bulk = out_stride == out_cols && base + PM_TILE <= n_eff && (reinterpret_cast<uintptr_t>(out + base * out_stride) & 15) == 0;
    const int t_row = bulk ? out_cols : 1, t_col = bulk ? 1 : PM_PAD;   // tile[row * t_row + col * t_col]

    float pos[D];
    const bool oob = load_position<D>(a, positions, idx, valid, pos);
    if (oob_out && half == 0 && valid) oob_out[idx] = oob ? 1 : 0;

    for (int lvl = half; lvl < a.n_levels; lvl += 2) {
        Simplex<D> s;
        locate<D>(pos, shift + lvl * D, scale + lvl * D, s);
        const float w_lvl = __ldg(window + lvl);
        const float2* table = lattice + (size_t)lvl * a.capacity;
        const HashBase<D> hb = hash_base<D>(s);
        uint32_t slot[D + 1];
        vertex_slots<D, POW2>(s, hb, a.capacity, slot);
        float2 v[D + 1];
#pragma unroll
        for (int r = 0; r <= D; r++) v[r] = __ldg(table + slot[r]);
        float ax = 0.f, ay = 0.f;
#pragma unroll
        for (int r = 0; r <= D; r++) {
            float w = s.bary[r] * w_lvl;
            ax = ax + v[r].x * w;
            ay = ay + v[r].y * w;
        }
        if (2 * lvl < out_cols) tile[p * t_row + (2 * lvl) * t_col] = ax;
        if (2 * lvl + 1 < out_cols) tile[p * t_row + (2 * lvl + 1) * t_col] = ay;
    }
    // concat-points levels (EncodingGPU.cuh:104-124): raw (normalised) coordinates, zero padded
    for (int e = half; e < a.n_extra; e += 2) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
            int src = i + e * 2;
            float val = 0.f;
#pragma unroll
            for (int k = 0; k < D; k++)
                if (k == src) val = pos[k] * a.points_scaling;
            const int c = 2 * (a.n_levels + e) + i;
            if (c < out_cols) tile[p * t_row + c * t_col] = val;
        }
    }
    if (bulk) {
        fence_proxy_async();   // the tile was written through the generic proxy, the bulk copy reads it through the async proxy
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_s2g(out + base * out_stride, tile, (uint32_t)(PM_TILE * out_cols * sizeof(float)));
            bulk_commit();
            bulk_wait_read_all();   // shared memory must outlive the read
        }
        return;
    }
    __syncthreads();
    const int rows = (int)min((int64_t)PM_TILE, n_eff - base);
    // flat walk over the tile's [rows, out_cols] block: (r, c) advance incrementally (one division per thread, not one per element)
    const int total = rows * out_cols;
    const int step_r = PM_THREADS / out_cols, step_c = PM_THREADS - step_r * out_cols;
    int r = threadIdx.x / out_cols, c = threadIdx.x - r * out_cols;
    for (int i = threadIdx.x; i < total; i += PM_THREADS) {
        out[(base + r) * out_stride + c] = tile[c * PM_PAD + r];
        r += step_r;
        c += step_c;
        if (c >= out_cols) {
            c -= out_cols;
            ++r;
        }
    }
}

// DPOS: also the position gradient (the lattice-only instantiation, the training step's, carries none of that code or its registers).
// KEYED: the block walks its 128 positions in the order of a small per-position key (a stable counting sort inside the tile).  The
// lattice gradient is bound by the number of reduction lanes that leave the SM, and red_add_runs only merges ADJACENT lanes that hit
// one vertex: in a packed K-layer ray packet adjacent samples are the K hits of one ray, while the samples that share vertices on the
// mid and fine levels are the same layer's hits of the neighbouring rays, K lanes apart.  With key = layer they become neighbours
// (B200, 893 k hits of the 5-shell benchmark scene: 0.604 ms in packed order, 0.416 ms layer-major, 1.33 ms shuffled).
template <int D, bool DPOS, bool KEYED, bool POW2>
__global__ void __launch_bounds__(PM_THREADS) permuto_bwd_kernel(PermutoArgs a, const float* __restrict__ positions,
                                                                 const int* __restrict__ order_key,
                                                                 const float2* __restrict__ lattice, const float* __restrict__ scale,
                                                                 const float* __restrict__ shift, const float* __restrict__ window,
                                                                 const float* __restrict__ d_out, int in_cols, int64_t in_stride,
                                                                 float2* __restrict__ d_lattice, float* __restrict__ d_positions, int64_t n,
                                                                 const int64_t* __restrict__ n_valid_dev) {
    extern __shared__ __align__(128) float tile[];   // upstream gradient [cols][PM_PAD] (or row-major as in memory: bulk path), then [D][PM_PAD]
    __shared__ __align__(8) uint64_t bar;
    const int64_t n_eff = n_valid_dev ? min(n, *n_valid_dev) : n;
    const int64_t base = (int64_t)blockIdx.x * PM_TILE;
    if (base >= n_eff) return;
    const int half = threadIdx.x >> 7, lane = threadIdx.x & 31;
    int p = threadIdx.x & (PM_TILE - 1);   // row of the tile this thread works on
    const int rows = (int)min((int64_t)PM_TILE, n_eff - base);
    const int n_cols = 2 * a.n_levels;   // the concat-points columns pass no gradient (Encoding.cu:135-139,163)
    // a full tile of densely packed gradient rows arrives as ONE bulk copy (row-major, read with the odd row stride in_cols) while the
    // threads sort the tile / load their positions, instead of a transposing loop whose loads every thread waits for (16 % of all stall
    // samples of the round-2 capture sat on that loop's stores)
    const bool bulk = in_stride == in_cols && rows == PM_TILE && (reinterpret_cast<uintptr_t>(d_out + base * in_stride) & 15) == 0;
    if (bulk && threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_proxy_async();
        mbar_arrive_expect_tx(&bar, (uint32_t)(PM_TILE * in_cols * sizeof(float)));
        bulk_g2s(tile, d_out + base * in_stride, (uint32_t)(PM_TILE * in_cols * sizeof(float)), &bar);
    }
    if (KEYED) {   // stable counting sort of the tile's rows by key (32 bins; rows beyond the end sort last)
        __shared__ int s_count[PM_TILE / 32][32];
        __shared__ unsigned char s_perm[PM_TILE];
        const int w = p >> 5;
        const int key = (p < rows) ? (__ldg(order_key + base + p) & 31) : 31;
        const uint32_t peers = __match_any_sync(VS_FULL_MASK, key);
        const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (half == 0) {
            s_count[w][lane] = 0;
            __syncwarp();
            if (rank_in_warp == 0) s_count[w][key] = __popc(peers);
        }
        __syncthreads();
        if (threadIdx.x < 32) {   // lane k: rows with key k per warp -> start of (warp, key) in the sorted order
            int c[PM_TILE / 32], tot = 0;
#pragma unroll
            for (int i = 0; i < PM_TILE / 32; i++) {
                c[i] = s_count[i][lane];
                tot += c[i];
            }
            int incl = tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int o = __shfl_up_sync(VS_FULL_MASK, incl, d);
                if (lane >= d) incl += o;
            }
            int start = incl - tot;
#pragma unroll
            for (int i = 0; i < PM_TILE / 32; i++) {
                s_count[i][lane] = start;
                start += c[i];
            }
        }
        __syncthreads();
        if (half == 0) s_perm[s_count[w][key] + rank_in_warp] = (unsigned char)p;
        __syncthreads();
        p = s_perm[p];
    }
    const int64_t idx = base + p;
    const bool valid = idx < n_eff;

    if (!bulk) {   // flat walk, (r, c) advanced incrementally (see permuto_fwd_kernel)
        const int step_r = PM_THREADS / in_cols, step_c = PM_THREADS - step_r * in_cols;
        int r = threadIdx.x / in_cols, c = threadIdx.x - r * in_cols;
        for (int i = threadIdx.x; i < PM_TILE * in_cols; i += PM_THREADS) {
            if (c < n_cols) tile[c * PM_PAD + r] = (r < rows) ? __ldg(d_out + (base + r) * in_stride + c) : 0.f;
            r += step_r;
            c += step_c;
            if (c >= in_cols) {
                c -= in_cols;
                ++r;
            }
        }
    }
    float pos[D];
    load_position<D>(a, positions, idx, valid, pos);
    __syncthreads();
    if (bulk) mbar_wait(&bar, 0);
    const int t_row = bulk ? in_cols : 1, t_col = bulk ? 1 : PM_PAD;   // tile[row * t_row + col * t_col]

    float d_pos[D];
#pragma unroll
    for (int i = 0; i < D; i++) d_pos[i] = 0.f;

    for (int lvl = half; lvl < a.n_levels; lvl += 2) {
        // columns beyond in_cols (a caller that dropped trailing columns, e.g. remove_last_element) carry zero gradient
        const float gx = (2 * lvl < in_cols) ? tile[p * t_row + (2 * lvl) * t_col] : 0.f;
        const float gy = (2 * lvl + 1 < in_cols) ? tile[p * t_row + (2 * lvl + 1) * t_col] : 0.f;
        Simplex<D> s;
        locate<D>(pos, shift + lvl * D, scale + lvl * D, s);
        const float w_lvl = __ldg(window + lvl);
        const HashBase<D> hb = hash_base<D>(s);
        uint32_t slot[D + 1];
        vertex_slots<D, POW2>(s, hb, a.capacity, slot);
        if (d_lattice) {
            float2* table = d_lattice + (size_t)lvl * a.capacity;
#pragma unroll
            for (int r = 0; r <= D; r++) {
                float w = s.bary[r] * w_lvl;
                red_add_runs(table, slot[r], gx * w, gy * w, valid, lane, a.agg_max_heads);
            }
        }
        if (DPOS) {   // EncodingGPU.cuh:630-690
            const float2* table = lattice + (size_t)lvl * a.capacity;
            float dl_db[D + 2];
#pragma unroll
            for (int r = 0; r <= D; r++) {
                float2 v = __ldg(table + slot[r]);
                float t = 0.f;
                t += v.x * w_lvl * gx;
                t += v.y * w_lvl * gy;
                dl_db[r] = t;
            }
            dl_db[D + 1] = 0.f + dl_db[0];
            // dl_de_i = dl_db[D - rank_i] / (D+1) - dl_db[D + 1 - rank_i] / (D+1): evaluated once per RANK (static indices), then each
            // vertex picks the value of its rank with a chain of selects spelled in PTX (any C++ form of `x = table[rank_i]`, select chains
            // included, is compiled into a dynamically indexed local-memory array: 5 STL + 40 LDL per level in the SASS of round 2)
            float e_by_rank[D + 1];
#pragma unroll
            for (int k = 0; k <= D; k++) {
                float e = 0.f;
                e += dl_db[D - k] * (1.0f / (D + 1));
                e -= dl_db[D + 1 - k] * (1.0f / (D + 1));
                e_by_rank[k] = e;
            }
            float dl_de[D + 1];
#pragma unroll
            for (int i = 0; i <= D; i++) {
                float e = 0.f;
#pragma unroll
                for (int k = 0; k <= D; k++) e = select_eq(s.rank[i], k, e_by_rank[k], e);
                dl_de[i] = e;
            }
#pragma unroll
            for (int i = 0; i < D; i++) {
                const float sc = __ldg(scale + lvl * D + i);
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j <= i; j++) acc += dl_de[j] * sc;
                acc -= dl_de[i + 1] * sc * (i + 1);
                d_pos[i] += acc;
            }
        }
    }
    if (DPOS) {
        __syncthreads();   // every thread is done reading the gradient tile
        if (half == 1) {
#pragma unroll
            for (int i = 0; i < D; i++) tile[i * PM_PAD + p] = d_pos[i];
        }
        __syncthreads();
        if (half == 0 && valid) {
#pragma unroll
            for (int i = 0; i < D; i++) {
                float g = d_pos[i] + tile[i * PM_PAD + p];
                if (a.has_bb) g = g * a.bb_inv_half[i] * 0.5f;   // chain rule through (p * inv_half + 1) / 2
                d_positions[idx * D + i] = g;
            }
        }
    }
}

static int fill_args(PermutoArgs& a, int pos_dim, int n_levels, int64_t capacity, int concat_points, float points_scaling, const float* bb_sides) {
    if (pos_dim < 2 || pos_dim > 4) return VS_ERR_UNSUPPORTED;
    if (n_levels < 1 || n_levels > 32 || capacity < 1 || capacity > (int64_t)0x7fffffff) return VS_ERR_INVALID_ARG;
    a.n_levels = n_levels;
    a.concat_points = concat_points != 0;
    a.n_extra = concat_points ? (pos_dim + 1) / 2 : 0;   // ceil(pos_dim / 2), Encoding.cu:72-75
    a.capacity = (uint32_t)capacity;
    a.pow2 = (capacity & (capacity - 1)) == 0;
    a.points_scaling = points_scaling;
    a.has_bb = bb_sides != nullptr;
    static const int agg_env = getenv("VS_PERMUTO_AGG") ? atoi(getenv("VS_PERMUTO_AGG")) : 4;   // measured on B200: 0 -> 2.35 ms, 4 -> 0.494, 20 -> 0.512, 32 -> 0.524
    a.agg_max_heads = agg_env;
    for (int i = 0; i < 8; i++) {
        a.bb_half[i] = 1.f;
        a.bb_inv_half[i] = 1.f;
    }
    if (bb_sides)
        for (int i = 0; i < pos_dim; i++) {
            a.bb_half[i] = bb_sides[i] / 2.0f;
            a.bb_inv_half[i] = 1.0f / a.bb_half[i];
        }
    return VS_OK;
}

}  // namespace vs

using namespace vs;

extern "C" {

int vs_permuto_output_dims(int pos_dim, int n_levels, int concat_points) {
    return 2 * (n_levels + (concat_points ? (pos_dim + 1) / 2 : 0));
}

int vs_permuto_forward(int pos_dim, int n_levels, int64_t capacity, int concat_points, float points_scaling, const float* bb_sides,
                       const float* positions, const float* lattice, const float* scale, const float* shift, const float* window, float* out,
                       int out_cols, int64_t out_stride, uint8_t* out_of_bounds, int64_t n, const int64_t* n_valid_dev, void* stream) {
    PermutoArgs a;
    int rc = fill_args(a, pos_dim, n_levels, capacity, concat_points, points_scaling, bb_sides);
    if (rc != VS_OK) return rc;
    const int cols = 2 * (a.n_levels + a.n_extra);
    VS_CHECK_ARG(n >= 0 && out_cols >= 1 && out_cols <= cols && out_stride >= out_cols && cols <= PM_MAX_COLS);
    if (n == 0) return VS_OK;
    VS_CHECK_ARG(positions && lattice && scale && shift && window && out);
    VS_CHECK_ARG(div_up(n, PM_TILE) <= 0x7fffffff);
    const dim3 grid((unsigned)div_up(n, PM_TILE));
    const size_t smem = (size_t)cols * PM_PAD * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
#define VS_PM_FWD_(D, POW2)                                                                                                                \
    permuto_fwd_kernel<D, POW2><<<grid, PM_THREADS, smem, st>>>(a, positions, (const float2*)lattice, scale, shift, window, out, out_cols, \
                                                                 out_stride, out_of_bounds, n, n_valid_dev)
#define VS_PM_FWD(D)           \
    if (a.pow2)                \
        VS_PM_FWD_(D, true);   \
    else                       \
        VS_PM_FWD_(D, false)
    if (pos_dim == 2) {
        VS_PM_FWD(2);
    } else if (pos_dim == 3) {
        VS_PM_FWD(3);
    } else {
        VS_PM_FWD(4);
    }
#undef VS_PM_FWD
#undef VS_PM_FWD_
    return launched(1);
}

int vs_permuto_backward_keyed(int pos_dim, int n_levels, int64_t capacity, int concat_points, const float* bb_sides, const float* positions,
                              const int32_t* order_key, const float* lattice, const float* scale, const float* shift, const float* window,
                              const float* d_out, int in_cols, int64_t in_stride, float* d_lattice, float* d_positions, int64_t n,
                              const int64_t* n_valid_dev, void* stream) {
    PermutoArgs a;
    int rc = fill_args(a, pos_dim, n_levels, capacity, concat_points, 1.0f, bb_sides);
    if (rc != VS_OK) return rc;
    const int cols = 2 * (a.n_levels + a.n_extra);
    VS_CHECK_ARG(n >= 0 && in_cols >= 1 && in_cols <= cols && in_stride >= in_cols && cols <= PM_MAX_COLS);
    if (n == 0 || (!d_lattice && !d_positions)) return VS_OK;
    VS_CHECK_ARG(positions && lattice && scale && shift && window && d_out);
    VS_CHECK_ARG(div_up(n, PM_TILE) <= 0x7fffffff);
    const dim3 grid((unsigned)div_up(n, PM_TILE));
    const size_t smem = (size_t)cols * PM_PAD * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
#define VS_PM_BWD__(D, DPOS, KEYED, POW2)                                                                                                  \
    permuto_bwd_kernel<D, DPOS, KEYED, POW2><<<grid, PM_THREADS, smem, st>>>(a, positions, order_key, (const float2*)lattice, scale,       \
                                                                              shift, window, d_out, in_cols, in_stride,                    \
                                                                              (float2*)d_lattice, d_positions, n, n_valid_dev)
#define VS_PM_BWD_(D, DPOS, KEYED)           \
    if (a.pow2)                              \
        VS_PM_BWD__(D, DPOS, KEYED, true);   \
    else                                     \
        VS_PM_BWD__(D, DPOS, KEYED, false)
#define VS_PM_BWD(D)                      \
    if (d_positions && order_key) {       \
        VS_PM_BWD_(D, true, true);        \
    } else if (d_positions) {             \
        VS_PM_BWD_(D, true, false);       \
    } else if (order_key) {               \
        VS_PM_BWD_(D, false, true);       \
    } else {                              \
        VS_PM_BWD_(D, false, false);      \
    }
    if (pos_dim == 2) {
        VS_PM_BWD(2);
    } else if (pos_dim == 3) {
        VS_PM_BWD(3);
    } else {
        VS_PM_BWD(4);
    }
#undef VS_PM_BWD
#undef VS_PM_BWD_
#undef VS_PM_BWD__
    return launched(1);
}

int vs_permuto_backward(int pos_dim, int n_levels, int64_t capacity, int concat_points, const float* bb_sides, const float* positions,
                        const float* lattice, const float* scale, const float* shift, const float* window, const float* d_out, int in_cols,
                        int64_t in_stride, float* d_lattice, float* d_positions, int64_t n, const int64_t* n_valid_dev, void* stream) {
    return vs_permuto_backward_keyed(pos_dim, n_levels, capacity, concat_points, bb_sides, positions, nullptr, lattice, scale, shift, window,
                                     d_out, in_cols, in_stride, d_lattice, d_positions, n, n_valid_dev, stream);
}

}  // extern "C"
