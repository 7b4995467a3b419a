/* ORACLE (test infrastructure only — never linked into or called by the product).
 *
 * Plain-C twin of oracle/permuto.py (the numpy restatement of the reference's permutohedral-lattice hash encoding,
 * submodules/permutohedral_encoding/kernels/permutohedral_encoding/EncodingGPU.cuh:22-45,68-261,264-416) for the CPU baseline of
 * bench.py: the reference's encoder has no CPU path at all, the numpy restatement runs on one core, this one runs the levels on all
 * host threads (OpenMP).  Arithmetic: fp32, separately rounded operations in the numpy oracle's order with fma=False (compile with
 * -ffp-contract=off); tests/test_oracle_permuto.py holds it to the numpy oracle bit for bit (forward) and to its position-ordered
 * accumulation (backward). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXD 8

/* elevation, closest remainder-0 point, ranks, barycentric weights of one position on one level (permuto.py:_simplex, fma=False) */
static void simplex(const float* pos, const float* shift, const float* scale, int d, int* rem0, int* rank, float* bary) {
    float elevated[MAXD + 1];
    float sm = 0.0f;
    for (int i = d; i > 0; --i) {
        float ps = pos[i - 1] + shift[i - 1];
        float cf = ps * scale[i - 1];
        float icf = (float)i * cf;
        elevated[i] = sm - icf;
        sm = sm + cf;
    }
    elevated[0] = sm;
    const float inv = 1.0f / (float)(d + 1);
    int sum = 0;
    for (int i = 0; i <= d; ++i) {
        float v = elevated[i] * inv;
        float up = ceilf(v) * (float)(d + 1);
        float down = floorf(v) * (float)(d + 1);
        float du = up - elevated[i], dd = elevated[i] - down;
        rem0[i] = (int)(du < dd ? up : down);
        sum += rem0[i];
    }
    sum /= (d + 1); /* C integer division (truncation), as the reference */
    float diff[MAXD + 1];
    for (int i = 0; i <= d; ++i) {
        diff[i] = elevated[i] - (float)rem0[i];
        rank[i] = 0;
    }
    for (int i = 0; i < d; ++i)
        for (int j = i + 1; j <= d; ++j) {
            if (diff[i] < diff[j]) rank[i]++;
            else rank[j]++;
        }
    for (int i = 0; i <= d; ++i) {
        rank[i] += sum;
        if (rank[i] < 0) {
            rank[i] += d + 1;
            rem0[i] += d + 1;
        } else if (rank[i] > d) {
            rank[i] -= d + 1;
            rem0[i] -= d + 1;
        }
    }
    for (int i = 0; i <= d + 1; ++i) bary[i] = 0.0f;
    for (int i = 0; i <= d; ++i) {
        float delta = (elevated[i] - (float)rem0[i]) * inv;
        bary[d - rank[i]] += delta;
        bary[d + 1 - rank[i]] -= delta;
    }
    bary[0] = bary[0] + (1.0f + bary[d + 1]);
}

static inline int64_t vertex_index(const int* rem0, const int* rank, int remainder, int d, uint32_t capacity) {
    uint32_t k = 0;
    for (int i = 0; i < d; ++i) {
        int key = rem0[i] + remainder;
        if (rank[i] > d - remainder) key -= d + 1;
        k += (uint32_t)key;
        k *= 2531011u;
    }
    return (int64_t)(k % capacity);
}

/* out: rows [n, 2*(L+extra)] (the layout modules.py:85 hands to the network) */
void vpo_forward(const float* positions, int64_t n, int d, const float* lattice, int L, int64_t capacity, const float* scale,
                 const float* shift, const float* window, int concat_points, float points_scaling, float* out_rows) {
    const int extra = concat_points ? (d + 1) / 2 : 0;
    const int width = 2 * (L + extra);
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n; ++p) {
        const float* pos = positions + p * d;
        float* o = out_rows + p * width;
        for (int lvl = 0; lvl < L; ++lvl) {
            int rem0[MAXD + 1], rank[MAXD + 1];
            float bary[MAXD + 2];
            simplex(pos, shift + lvl * d, scale + lvl * d, d, rem0, rank, bary);
            float a0 = 0.0f, a1 = 0.0f;
            for (int r = 0; r <= d; ++r) {
                const int64_t idx = vertex_index(rem0, rank, r, d, (uint32_t)capacity);
                const float w = bary[r] * window[lvl];
                const float* val = lattice + ((int64_t)lvl * capacity + idx) * 2;
                float t0 = val[0] * w, t1 = val[1] * w;
                a0 = t0 + a0;
                a1 = t1 + a1;
            }
            o[2 * lvl] = a0;
            o[2 * lvl + 1] = a1;
        }
        for (int e = 0; e < extra; ++e)
            for (int i = 0; i < 2; ++i) {
                const int src = i + e * 2;
                o[2 * (L + e) + i] = src < d ? pos[src] * points_scaling : 0.0f;
            }
    }
}

/* lattice gradient only (what the training step needs): g_lat [L, capacity, 2] zeroed by the caller; one level per thread, positions in
 * order (the numpy oracle's accumulation order) */
void vpo_backward_lattice(const float* positions, int64_t n, int d, int L, int64_t capacity, const float* scale, const float* shift,
                          const float* window, const float* grad_rows, int width, float* g_lat) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int lvl = 0; lvl < L; ++lvl) {
        float* gl = g_lat + (int64_t)lvl * capacity * 2;
        for (int64_t p = 0; p < n; ++p) {
            int rem0[MAXD + 1], rank[MAXD + 1];
            float bary[MAXD + 2];
            simplex(positions + p * d, shift + lvl * d, scale + lvl * d, d, rem0, rank, bary);
            const float gx = grad_rows[p * width + 2 * lvl], gy = grad_rows[p * width + 2 * lvl + 1];
            for (int r = 0; r <= d; ++r) {
                const int64_t idx = vertex_index(rem0, rank, r, d, (uint32_t)capacity);
                const float w = bary[r] * window[lvl];
                float t0 = gx * w, t1 = gy * w;
                gl[2 * idx] += t0;
                gl[2 * idx + 1] += t1;
            }
        }
    }
}
