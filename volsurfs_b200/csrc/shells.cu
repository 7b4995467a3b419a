// volsurfs_b200 — K-layer shell intersector: nearest hit of every ray against each of K nested meshes in ONE launch.
//
// Replaces, for the per-ray rendering path, the K python-loop calls of raytracelib's RayTracer.trace
// (volsurfs_py/methods/volsurfs.py:476-485 -> raytracelib/raytracer.py:35-113 -> src/bvh.cu:186-263,420-469): there, one thread
// per ray walks a 4-ary median-split BVH with 32-byte AoS nodes and 56-byte AoS triangles out of global memory, one mesh per
// launch, with a device synchronisation and a host-visible any_hit flag after every mesh.
//
// Here:
//   * the BVH of each layer is built on the host with a binned surface-area heuristic, collapsed to 4-wide nodes whose four
//     child boxes sit in one 128-byte line (SoA: 6 x float4 + 4 child refs), emitted in breadth-first order;
//   * one launch covers all (layer, ray) pairs with persistent CTAs (kTraceCtasPerSm per SM, split over the layers): every WARP draws
//     its next 32 consecutive rays from the layer's work counter, so a warp whose rays miss everything is not parked until the slowest
//     warp of its CTA is done (B200, C2 scene: 0.695 ms with a static 512-rays-per-CTA mapping, achieved occupancy 33 % of a 50 %
//     limit -> 0.616 ms); a warp whose layer has run dry moves on to the layers that still have work;
//   * nodes and triangles are read through L1 / L2.  (Rounds 1-2 staged the top four levels of the CTA's layer in shared memory by a TMA
//     bulk copy: with persistent CTAs that copy is a one-off and the 87 KB of shared memory per SM it took are worth more as L1 —
//     0.546 ms staged, 0.535 ms without, B200, C2 scene);
//   * triangles are 3 x float4 (vertex + original face index), read with 16-byte loads.
//
// Parity contract (bit-exact hits, pinned by the reference's own kernel): the ray/triangle arithmetic is the reference's
// (include/raytracing/triangle.cuh:42-70, Eigen 3.3.7 evaluation order) with exactly the FMA contractions nvcc -O3 applies to the
// reference's raytrace_kernel for sm_100a — read off the SASS of oracle/_ref/libraytrace_ref.so (the reference's src/bvh.cu compiled where
// it lies) and spelled out with __fmaf_rn/__fmul_rn so that no compiler decision is left:
//   cross(a,b)_i = fma(a_j, b_k, -rn(a_k*b_j));   n = (b-a)x(c-a), q = (o-a)xd
//   d.n, n.(o-a), n.n = fma(x0,y0, fma(x2,y2, rn(x1*y1)));   q.(c-a), q.(b-a) = fma(x0,y0, fma(x1,y1, rn(x2*y2)))
//   D = 1/(d.n) (IEEE), u = rn(D*-(q.(c-a))), v = rn(D*(q.(b-a))), t = rn(D*-(n.(o-a)));  position = fma(t, d, o)
//   miss if u<0 || u>1 || v<0 || u+v>1 || t<0;  accept if t > 0 && t < best.
// tests/test_gpu_raytrace.py holds this kernel, the reference's kernel and oracle/raytrace_oracle.c (contract "device") to the same bits.
// The nearest hit does not depend on the BVH as long as no box containing the winning triangle is culled, so child boxes are
// padded by a relative epsilon and children are visited whenever t_near <= best (ties in t resolve to the lowest original
// face index, which is what a brute-force pass in index order returns).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "vs_common.cuh"

namespace vs {

constexpr float kMaxDist = 1e6f;  // include/raytracing/common.h:21
constexpr int kTraceThreads = 128;
constexpr int kTraceCtasPerSm = 8;   // 63 registers x 128 threads: 8 CTAs fit (measured: 4 -> 0.74 ms, 8 -> 0.616, 12 -> 0.646)
constexpr int kTraceMaxLayers = 64;
constexpr int kCounterSets = 8;      // work-counter sets handed out in turn, so that launches in flight on different streams do not share one
constexpr int kStack = 48;
constexpr int kLeafMax = 2;  // measured on the C2 scene (B200, persistent kernel): 1 -> 0.580 ms, 2 -> 0.569, 4 -> 0.593, 6 -> 0.607, 8 -> 0.640

struct __align__(16) Node4 {
    float lox[4], loy[4], loz[4], hix[4], hiy[4], hiz[4];
    int32_t child[4];  // > 0 inner node index, 0 empty slot, < 0 leaf: ~child = first_tri * 8 + (count - 1)
    int32_t pad[4];
};
static_assert(sizeof(Node4) == 128, "Node4 must be one 128-byte line");

struct Layer {
    Node4* nodes = nullptr;      // device
    float4* tris = nullptr;      // device, 3 float4 per triangle in BVH order; .w of the first = original face index
    float4* pre = nullptr;       // device, 4 float4 per triangle: (a, idx) (e1 = b-a) (e2 = c-a) (n = e1 x e2), oracle arithmetic
    int32_t* orig_to_bvh = nullptr;  // device
    int32_t n_nodes = 0;
    int64_t n_tris = 0;
};

struct Shells {
    int K = 0;
    std::vector<Layer> layers;
    Layer* layers_dev = nullptr;  // device copy of the table
    int* overflow_dev = nullptr;  // [0] set to 1 if a traversal stack ever overflowed; then kCounterSets x kTraceMaxLayers work counters
    unsigned launches = 0;        // picks the counter set
};

// =====================================================================================================================
// host: binned-SAH BVH2 -> BVH4 collapse -> BFS layout
// =====================================================================================================================
struct Box {
    float lo[3], hi[3];
    void reset() {
        for (int k = 0; k < 3; ++k) lo[k] = FLT_MAX, hi[k] = -FLT_MAX;
    }
    void grow(const float* p) {
        for (int k = 0; k < 3; ++k) lo[k] = std::min(lo[k], p[k]), hi[k] = std::max(hi[k], p[k]);
    }
    void grow(const Box& b) {
        for (int k = 0; k < 3; ++k) lo[k] = std::min(lo[k], b.lo[k]), hi[k] = std::max(hi[k], b.hi[k]);
    }
    float area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0.f;
        return 2.f * (dx * dy + dy * dz + dz * dx);
    }
};

struct B2Node {
    Box box;
    int left = -1, right = -1;  // children (inner) ...
    int first = 0, count = 0;   // ... or primitive range (leaf, count > 0)
};

struct Builder {
    const float* verts;
    const int32_t* faces;
    int64_t n_faces;
    std::vector<Box> pbox;
    std::vector<float> pcen;  // 3 per prim
    std::vector<int32_t> order;
    std::vector<B2Node> nodes;

    void prim_setup() {
        pbox.resize(n_faces);
        pcen.resize(3 * n_faces);
        order.resize(n_faces);
        for (int64_t i = 0; i < n_faces; ++i) {
            Box b;
            b.reset();
            for (int c = 0; c < 3; ++c) b.grow(verts + 3 * (int64_t)faces[3 * i + c]);
            pbox[i] = b;
            for (int k = 0; k < 3; ++k) pcen[3 * i + k] = 0.5f * (b.lo[k] + b.hi[k]);
            order[i] = (int32_t)i;
        }
    }

    void build() {
        prim_setup();
        nodes.reserve(2 * n_faces / 2 + 16);
        nodes.emplace_back();
        struct Item {
            int node;
            int64_t lo, hi;
        };
        std::vector<Item> stack;
        stack.push_back({0, 0, n_faces});
        constexpr int NB = 16;
        while (!stack.empty()) {
            Item it = stack.back();
            stack.pop_back();
            Box nb, cb;
            nb.reset();
            cb.reset();
            for (int64_t i = it.lo; i < it.hi; ++i) {
                nb.grow(pbox[order[i]]);
                cb.grow(&pcen[3 * (int64_t)order[i]]);
            }
            nodes[it.node].box = nb;
            const int64_t cnt = it.hi - it.lo;
            if (cnt <= kLeafMax) {
                nodes[it.node].first = (int)it.lo;
                nodes[it.node].count = (int)cnt;
                continue;
            }
            int axis = 0;
            float ext[3] = {cb.hi[0] - cb.lo[0], cb.hi[1] - cb.lo[1], cb.hi[2] - cb.lo[2]};
            if (ext[1] > ext[axis]) axis = 1;
            if (ext[2] > ext[axis]) axis = 2;
            int64_t mid = -1;
            if (ext[axis] > 0.f) {
                Box bb[NB];
                int bc[NB];
                for (int b = 0; b < NB; ++b) bb[b].reset(), bc[b] = 0;
                const float scale = NB / ext[axis];
                auto bin_of = [&](int32_t p) {
                    int b = (int)((pcen[3 * (int64_t)p + axis] - cb.lo[axis]) * scale);
                    return std::min(std::max(b, 0), NB - 1);
                };
                for (int64_t i = it.lo; i < it.hi; ++i) {
                    int b = bin_of(order[i]);
                    bb[b].grow(pbox[order[i]]);
                    bc[b]++;
                }
                float right_area[NB];
                int right_cnt[NB];
                Box acc;
                acc.reset();
                int c = 0;
                for (int b = NB - 1; b > 0; --b) {
                    acc.grow(bb[b]);
                    c += bc[b];
                    right_area[b] = acc.area();
                    right_cnt[b] = c;
                }
                acc.reset();
                c = 0;
                float best = FLT_MAX;
                int best_split = -1;
                for (int b = 0; b < NB - 1; ++b) {
                    acc.grow(bb[b]);
                    c += bc[b];
                    if (c == 0 || right_cnt[b + 1] == 0) continue;
                    float cost = acc.area() * c + right_area[b + 1] * right_cnt[b + 1];
                    if (cost < best) best = cost, best_split = b;
                }
                if (best_split >= 0) {
                    auto first_right = std::partition(order.begin() + it.lo, order.begin() + it.hi,
                                                      [&](int32_t p) { return bin_of(p) <= best_split; });
                    mid = first_right - order.begin();
                }
            }
            if (mid <= it.lo || mid >= it.hi) {  // degenerate: median split in index order along the axis
                mid = it.lo + cnt / 2;
                std::nth_element(order.begin() + it.lo, order.begin() + mid, order.begin() + it.hi, [&](int32_t a, int32_t b) {
                    return pcen[3 * (int64_t)a + axis] < pcen[3 * (int64_t)b + axis];
                });
            }
            int l = (int)nodes.size();
            nodes.emplace_back();
            nodes.emplace_back();
            nodes[it.node].left = l;
            nodes[it.node].right = l + 1;
            stack.push_back({l, it.lo, mid});
            stack.push_back({l + 1, mid, it.hi});
        }
    }
};

// collapse + BFS numbering. Returns wide nodes and the triangle order.
static void collapse_bfs(const Builder& B, std::vector<Node4>& out) {
    struct Wide {
        int kids[4];
        int n;
    };
    // BFS over "wide" nodes: each is identified by the BVH2 node it was grown from
    std::vector<int> queue;   // BVH2 root of each wide node, in BFS order
    std::vector<Wide> wides;
    queue.push_back(0);
    size_t head = 0;
    std::vector<int> wide_index_of(B.nodes.size(), -1);
    while (head < queue.size()) {
        int root = queue[head];
        wide_index_of[root] = (int)head;
        ++head;
        Wide w;
        w.n = 0;
        const B2Node& r = B.nodes[root];
        if (r.count > 0) {  // a single-leaf tree: the root itself becomes the only child
            w.kids[w.n++] = root;
        } else {
            w.kids[w.n++] = r.left;
            w.kids[w.n++] = r.right;
            while (w.n < 4) {
                int pick = -1;
                float best = -1.f;
                for (int i = 0; i < w.n; ++i) {
                    const B2Node& c = B.nodes[w.kids[i]];
                    if (c.count == 0 && c.box.area() > best) best = c.box.area(), pick = i;
                }
                if (pick < 0) break;
                const B2Node& c = B.nodes[w.kids[pick]];
                w.kids[pick] = c.left;
                w.kids[w.n++] = c.right;
            }
        }
        for (int i = 0; i < w.n; ++i)
            if (B.nodes[w.kids[i]].count == 0) queue.push_back(w.kids[i]);
        wides.push_back(w);
    }
    out.resize(wides.size());
    for (size_t wi = 0; wi < wides.size(); ++wi) {
        Node4 nd;
        std::memset(&nd, 0, sizeof(nd));
        for (int i = 0; i < 4; ++i) {
            if (i < wides[wi].n) {
                const B2Node& c = B.nodes[wides[wi].kids[i]];
                float pad[3];
                for (int k = 0; k < 3; ++k) {
                    float m = std::max(std::fabs(c.box.lo[k]), std::fabs(c.box.hi[k]));
                    pad[k] = 4e-6f * m + 1e-7f * (c.box.hi[k] - c.box.lo[k]) + 1e-30f;
                }
                nd.lox[i] = c.box.lo[0] - pad[0];
                nd.loy[i] = c.box.lo[1] - pad[1];
                nd.loz[i] = c.box.lo[2] - pad[2];
                nd.hix[i] = c.box.hi[0] + pad[0];
                nd.hiy[i] = c.box.hi[1] + pad[1];
                nd.hiz[i] = c.box.hi[2] + pad[2];
                if (c.count > 0) nd.child[i] = ~(c.first * 8 + (c.count - 1));
                else nd.child[i] = wide_index_of[wides[wi].kids[i]];
            } else {
                nd.lox[i] = nd.loy[i] = nd.loz[i] = FLT_MAX;
                nd.hix[i] = nd.hiy[i] = nd.hiz[i] = -FLT_MAX;
                nd.child[i] = 0;
            }
        }
        out[wi] = nd;
    }
}

// =====================================================================================================================
// device
// =====================================================================================================================
// dots that involve the face normal (d.n, n.rov0, n.n): the middle product is the separately rounded one
__device__ __forceinline__ float dot_n(float x0, float x1, float x2, float y0, float y1, float y2) {
    return __fmaf_rn(x0, y0, __fmaf_rn(x2, y2, __fmul_rn(x1, y1)));
}
// dots of q with a triangle edge: the last product is the separately rounded one
__device__ __forceinline__ float dot_q(float x0, float x1, float x2, float y0, float y1, float y2) {
    return __fmaf_rn(x0, y0, __fmaf_rn(x1, y1, __fmul_rn(x2, y2)));
}
__device__ __forceinline__ void cross_ref(float a0, float a1, float a2, float b0, float b1, float b2, float& r0, float& r1, float& r2) {
    r0 = __fmaf_rn(a1, b2, -__fmul_rn(a2, b1));
    r1 = __fmaf_rn(a2, b0, -__fmul_rn(a0, b2));
    r2 = __fmaf_rn(a0, b1, -__fmul_rn(a1, b0));
}
// Eigen normalized() as the reference kernel evaluates it: v / sqrt(v.v) with IEEE sqrt and divisions, unchanged for the zero vector
__device__ __forceinline__ void normalize_ref(float& nx, float& ny, float& nz) {
    const float zz = dot_n(nx, ny, nz, nx, ny, nz);
    if (zz > 0.0f) {
        const float s = __fsqrt_rn(zz);
        nx = __fdiv_rn(nx, s);
        ny = __fdiv_rn(ny, s);
        nz = __fdiv_rn(nz, s);
    }
}

// triangle.cuh:42-70 in the reference kernel's arithmetic; returns is_hit, leaves t/u/v as computed.  e1 = b-a, e2 = c-a and n = e1 x e2
// do not depend on the ray: they are computed once per triangle by precompute_tri_kernel with the very same rounded operations.
__device__ __forceinline__ bool tri_test(const float4 A, const float4 E1, const float4 E2, const float4 Nn, float ox, float oy, float oz,
                                         float dx, float dy, float dz, float& t, float& u, float& v) {
    const float rx = __fsub_rn(ox, A.x), ry = __fsub_rn(oy, A.y), rz = __fsub_rn(oz, A.z);
    float qx, qy, qz;
    cross_ref(rx, ry, rz, dx, dy, dz, qx, qy, qz);
    const float D = __fdiv_rn(1.0f, dot_n(dx, dy, dz, Nn.x, Nn.y, Nn.z));
    u = __fmul_rn(D, -dot_q(qx, qy, qz, E2.x, E2.y, E2.z));
    v = __fmul_rn(D, dot_q(qx, qy, qz, E1.x, E1.y, E1.z));
    t = __fmul_rn(D, -dot_n(Nn.x, Nn.y, Nn.z, rx, ry, rz));
    return !(u < 0.0f || u > 1.0f || v < 0.0f || __fadd_rn(u, v) > 1.0f || t < 0.0f);
}

// the same from raw vertices
__device__ __forceinline__ bool tri_test_raw(const float4 A, const float4 Bv, const float4 C, float ox, float oy, float oz, float dx, float dy,
                                         float dz, float& t, float& u, float& v) {
    const float e1x = __fsub_rn(Bv.x, A.x), e1y = __fsub_rn(Bv.y, A.y), e1z = __fsub_rn(Bv.z, A.z);
    const float e2x = __fsub_rn(C.x, A.x), e2y = __fsub_rn(C.y, A.y), e2z = __fsub_rn(C.z, A.z);
    float nx, ny, nz;
    cross_ref(e1x, e1y, e1z, e2x, e2y, e2z, nx, ny, nz);
    return tri_test(A, make_float4(e1x, e1y, e1z, 0.f), make_float4(e2x, e2y, e2z, 0.f), make_float4(nx, ny, nz, 0.f), ox, oy, oz, dx, dy, dz,
                    t, u, v);
}

__global__ void __launch_bounds__(256) precompute_tri_kernel(const float4* __restrict__ tris, float4* __restrict__ pre, int64_t n_tris) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_tris) return;
    const float4 A = tris[3 * p], Bv = tris[3 * p + 1], C = tris[3 * p + 2];
    const float e1x = __fsub_rn(Bv.x, A.x), e1y = __fsub_rn(Bv.y, A.y), e1z = __fsub_rn(Bv.z, A.z);
    const float e2x = __fsub_rn(C.x, A.x), e2y = __fsub_rn(C.y, A.y), e2z = __fsub_rn(C.z, A.z);
    float nx, ny, nz;
    cross_ref(e1x, e1y, e1z, e2x, e2y, e2z, nx, ny, nz);
    pre[4 * p] = A;
    pre[4 * p + 1] = make_float4(e1x, e1y, e1z, 0.f);
    pre[4 * p + 2] = make_float4(e2x, e2y, e2z, 0.f);
    pre[4 * p + 3] = make_float4(nx, ny, nz, 0.f);
}

__device__ __forceinline__ void cswap_desc(float& ka, int& va, float& kb, int& vb) {  // larger key first
    if (ka < kb) {
        float tk = ka;
        ka = kb;
        kb = tk;
        int tv = va;
        va = vb;
        vb = tv;
    }
}

template <bool WW, bool PRE>
__global__ void __launch_bounds__(kTraceThreads) shells_trace_kernel(const Layer* __restrict__ layers, int layer_first,
                                                                     const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                                     int64_t n_rays, float* __restrict__ depth_out,
                                                                     int32_t* __restrict__ tri_out, float* __restrict__ u_out,
                                                                     float* __restrict__ v_out, int* __restrict__ overflow, int layer_count,
                                                                     int* __restrict__ work_counter) {
    // the CTA's own layer first; a WARP whose layer has run dry moves on to the others by itself: no block-wide hand-over, nobody waits
    // for a neighbour's last chunk.  No shared memory: the whole tree comes through L1 (see the header comment)
    for (int a = 0; a < layer_count; ++a) {
    const int li = (int)((blockIdx.y + a) % layer_count);
    if (a > 0 && (int64_t)(*reinterpret_cast<volatile int*>(work_counter + li)) * 32 >= n_rays) continue;
    const Layer L = layers[layer_first + li];

    const Node4* __restrict__ gnodes = L.nodes;
    const float4* __restrict__ pre = L.pre;
    const float4* __restrict__ tris = L.tris;
    const int64_t out_base = (int64_t)li * n_rays;  // outputs are [n_layers_traced, n_rays]

    for (;;) {
        int c = 0;
        if ((threadIdx.x & 31) == 0) c = atomicAdd(work_counter + li, 1);
        c = __shfl_sync(VS_FULL_MASK, c, 0);
        if ((int64_t)c * 32 >= n_rays) break;
        const int64_t r = (int64_t)c * 32 + (threadIdx.x & 31);
        if (r >= n_rays) continue;
        const float ox = __ldg(rays_o + 3 * r), oy = __ldg(rays_o + 3 * r + 1), oz = __ldg(rays_o + 3 * r + 2);
        const float dx = __ldg(rays_d + 3 * r), dy = __ldg(rays_d + 3 * r + 1), dz = __ldg(rays_d + 3 * r + 2);
        const float idx_ = 1.0f / dx, idy = 1.0f / dy, idz = 1.0f / dz;

        float best_t = kMaxDist, best_u = 0.f, best_v = 0.f;
        int best_tri = -1;

        if (WW) {
        int stack_ref[kStack];
            float stack_t[kStack];
            int sp = 0;
            // "while-while" traversal: every lane first descends through inner nodes until its current entry is a leaf, then the
            // lanes of the warp test their leaves together (the triangle test is the expensive, otherwise badly diverged part)
            int ref = 0;            // current entry: >= 0 wide node index (0 = root), < 0 leaf
            float ref_t = -FLT_MAX;  // its entry distance
            bool alive = true;
            auto pop = [&]() {
                while (sp > 0) {
                    --sp;
                    if (stack_t[sp] <= best_t) {
                        ref = stack_ref[sp];
                        ref_t = stack_t[sp];
                        return true;
                    }
                }
                return false;
            };
            while (alive) {
                while (alive && ref >= 0) {
                    const Node4* nd = &gnodes[ref];
                    const float4 lox = *reinterpret_cast<const float4*>(nd->lox), loy = *reinterpret_cast<const float4*>(nd->loy);
                    const float4 loz = *reinterpret_cast<const float4*>(nd->loz), hix = *reinterpret_cast<const float4*>(nd->hix);
                    const float4 hiy = *reinterpret_cast<const float4*>(nd->hiy), hiz = *reinterpret_cast<const float4*>(nd->hiz);
                    const int4 ch = *reinterpret_cast<const int4*>(nd->child);
                    float tn[4];
                    int cr[4] = {ch.x, ch.y, ch.z, ch.w};
                    const float lx[4] = {lox.x, lox.y, lox.z, lox.w}, ly[4] = {loy.x, loy.y, loy.z, loy.w}, lz[4] = {loz.x, loz.y, loz.z, loz.w};
                    const float hx[4] = {hix.x, hix.y, hix.z, hix.w}, hy[4] = {hiy.x, hiy.y, hiy.z, hiy.w}, hz[4] = {hiz.x, hiz.y, hiz.z, hiz.w};
    #pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float t0x = (lx[i] - ox) * idx_, t1x = (hx[i] - ox) * idx_;
                        const float t0y = (ly[i] - oy) * idy, t1y = (hy[i] - oy) * idy;
                        const float t0z = (lz[i] - oz) * idz, t1z = (hz[i] - oz) * idz;
                        const float tnear = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
                        const float tfar = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
                        // padded boxes + slack on the far plane keep the test conservative w.r.t. the triangle arithmetic
                        const bool hit = cr[i] != 0 && tnear <= tfar * 1.0000004f + 1e-30f && tfar >= 0.0f && tnear <= best_t;
                        tn[i] = hit ? tnear : FLT_MAX;
                    }
                    // descending by entry distance: invalid (FLT_MAX) first, the nearest child last
                    cswap_desc(tn[0], cr[0], tn[2], cr[2]);
                    cswap_desc(tn[1], cr[1], tn[3], cr[3]);
                    cswap_desc(tn[0], cr[0], tn[1], cr[1]);
                    cswap_desc(tn[2], cr[2], tn[3], cr[3]);
                    cswap_desc(tn[1], cr[1], tn[2], cr[2]);
                    if (tn[3] == FLT_MAX) {
                        alive = pop();
                    } else {
    #pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            if (tn[i] != FLT_MAX) {
                                if (sp < kStack) {
                                    stack_ref[sp] = cr[i];
                                    stack_t[sp] = tn[i];
                                    ++sp;
                                } else {
                                    atomicOr(overflow, 1);  // never expected (depth*3 << kStack); reported by vs_shells_overflowed
                                }
                            }
                        }
                        ref = cr[3];  // continue with the nearest child without a round trip through the stack
                        ref_t = tn[3];
                    }
                }
                if (!alive) break;
                if (ref_t <= best_t) {
                    const int enc = ~ref;
                    const int first = enc >> 3, cnt = (enc & 7) + 1;
                    for (int k = 0; k < cnt; ++k) {
                        float4 A;
                        float t, u, v;
                        bool ok;
                        if (PRE) {
                            const float4* tp = pre + 4 * (int64_t)(first + k);
                            A = __ldg(tp);
                            ok = tri_test(A, __ldg(tp + 1), __ldg(tp + 2), __ldg(tp + 3), ox, oy, oz, dx, dy, dz, t, u, v);
                        } else {
                            const float4* tp = tris + 3 * (int64_t)(first + k);
                            A = __ldg(tp);
                            ok = tri_test_raw(A, __ldg(tp + 1), __ldg(tp + 2), ox, oy, oz, dx, dy, dz, t, u, v);
                        }
                        if (ok) {
                            const int oi = __float_as_int(A.w);
                            if (t > 0.0f && (t < best_t || (t == best_t && best_tri >= 0 && oi < best_tri))) {
                                best_t = t;
                                best_u = u;
                                best_v = v;
                                best_tri = oi;
                            }
                        }
                    }
                }
                alive = pop();
            }
    
        } else {
            // one 8-byte entry (node reference, entry distance): a push or a pop is ONE 64-bit local-memory access, not two 32-bit ones
            int2 stack[kStack];
            int sp = 0;
            // the entry to work on lives in registers: the nearest child of a node is taken up directly, only its siblings go through
            // the stack
            int ref = 0;  // the root is wide node 0
            float ref_t = -FLT_MAX;
            bool have = true;
            for (;;) {
                if (!have) {
                    if (sp == 0) break;
                    --sp;
                    const int2 top = stack[sp];
                    ref = top.x;
                    ref_t = __int_as_float(top.y);
                }
                have = false;
                if (ref_t > best_t) continue;
                if (ref < 0) {
                    const int enc = ~ref;
                    const int first = enc >> 3, cnt = (enc & 7) + 1;
                    for (int k = 0; k < cnt; ++k) {
                        float4 A;
                        float t, u, v;
                        bool ok;
                        if (PRE) {
                            const float4* tp = pre + 4 * (int64_t)(first + k);
                            A = __ldg(tp);
                            ok = tri_test(A, __ldg(tp + 1), __ldg(tp + 2), __ldg(tp + 3), ox, oy, oz, dx, dy, dz, t, u, v);
                        } else {
                            const float4* tp = tris + 3 * (int64_t)(first + k);
                            A = __ldg(tp);
                            ok = tri_test_raw(A, __ldg(tp + 1), __ldg(tp + 2), ox, oy, oz, dx, dy, dz, t, u, v);
                        }
                        if (ok) {
                            const int oi = __float_as_int(A.w);
                            if (t > 0.0f && (t < best_t || (t == best_t && best_tri >= 0 && oi < best_tri))) {
                                best_t = t;
                                best_u = u;
                                best_v = v;
                                best_tri = oi;
                            }
                        }
                    }
                    continue;
                }
                const Node4* nd = &gnodes[ref];
                const float4 lox = *reinterpret_cast<const float4*>(nd->lox), loy = *reinterpret_cast<const float4*>(nd->loy);
                const float4 loz = *reinterpret_cast<const float4*>(nd->loz), hix = *reinterpret_cast<const float4*>(nd->hix);
                const float4 hiy = *reinterpret_cast<const float4*>(nd->hiy), hiz = *reinterpret_cast<const float4*>(nd->hiz);
                const int4 ch = *reinterpret_cast<const int4*>(nd->child);
                float tn[4];
                int cr[4] = {ch.x, ch.y, ch.z, ch.w};
                const float lx[4] = {lox.x, lox.y, lox.z, lox.w}, ly[4] = {loy.x, loy.y, loy.z, loy.w}, lz[4] = {loz.x, loz.y, loz.z, loz.w};
                const float hx[4] = {hix.x, hix.y, hix.z, hix.w}, hy[4] = {hiy.x, hiy.y, hiy.z, hiy.w}, hz[4] = {hiz.x, hiz.y, hiz.z, hiz.w};
    #pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float t0x = (lx[i] - ox) * idx_, t1x = (hx[i] - ox) * idx_;
                    const float t0y = (ly[i] - oy) * idy, t1y = (hy[i] - oy) * idy;
                    const float t0z = (lz[i] - oz) * idz, t1z = (hz[i] - oz) * idz;
                    const float tnear = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
                    const float tfar = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
                    // padded boxes + slack on the far plane keep the test conservative w.r.t. the triangle arithmetic
                    const bool hit = cr[i] != 0 && tnear <= tfar * 1.0000004f + 1e-30f && tfar >= 0.0f && tnear <= best_t;
                    tn[i] = hit ? tnear : FLT_MAX;
                }
                // far -> near so that the nearest child is popped first
                cswap_desc(tn[0], cr[0], tn[2], cr[2]);
                cswap_desc(tn[1], cr[1], tn[3], cr[3]);
                cswap_desc(tn[0], cr[0], tn[1], cr[1]);
                cswap_desc(tn[2], cr[2], tn[3], cr[3]);
                cswap_desc(tn[1], cr[1], tn[2], cr[2]);
    #pragma unroll
                for (int i = 0; i < 3; ++i) {
                    if (tn[i] != FLT_MAX) {
                        if (sp < kStack) {
                            stack[sp] = make_int2(cr[i], __float_as_int(tn[i]));
                            ++sp;
                        } else {
                            atomicOr(overflow, 1);  // never expected (depth*3 << kStack); reported by vs_shells_overflowed
                        }
                    }
                }
                if (tn[3] != FLT_MAX) {  // (sorted far -> near: a node with any hit child has its nearest in slot 3)
                    ref = cr[3];
                    ref_t = tn[3];
                    have = true;
                }
            }
    
        }
        depth_out[out_base + r] = best_t;
        tri_out[out_base + r] = best_tri;
        u_out[out_base + r] = best_u;
        v_out[out_base + r] = best_v;
    }
    }
}

// reference-format outputs of RayTracer.trace(mesh_id) (bvh.cu:440-468, raytracer.py:103-113) from the compact hit record
__global__ void __launch_bounds__(256) shells_expand_kernel(const Layer* __restrict__ layers, int layer, const float* __restrict__ rays_o,
                                                            const float* __restrict__ rays_d, const float* __restrict__ depth,
                                                            const int32_t* __restrict__ tri, const float* __restrict__ u,
                                                            const float* __restrict__ v, int64_t n_rays, float* __restrict__ positions,
                                                            float* __restrict__ normals, int64_t* __restrict__ tri_mesh_id,
                                                            int64_t* __restrict__ tri_id, float* __restrict__ bary) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const Layer L = layers[layer];
    const float t = depth[r];
    positions[3 * r] = __fmaf_rn(rays_d[3 * r], t, rays_o[3 * r]);  // ray_o + depth * ray_d, contracted by nvcc in the reference kernel
    positions[3 * r + 1] = __fmaf_rn(rays_d[3 * r + 1], t, rays_o[3 * r + 1]);
    positions[3 * r + 2] = __fmaf_rn(rays_d[3 * r + 2], t, rays_o[3 * r + 2]);
    const int ti = tri[r];
    if (ti >= 0) {
        const int64_t p = L.orig_to_bvh[ti];
        const float4 A = L.tris[3 * p], Bv = L.tris[3 * p + 1], C = L.tris[3 * p + 2];
        const float e1x = __fsub_rn(Bv.x, A.x), e1y = __fsub_rn(Bv.y, A.y), e1z = __fsub_rn(Bv.z, A.z);
        const float e2x = __fsub_rn(C.x, A.x), e2y = __fsub_rn(C.y, A.y), e2z = __fsub_rn(C.z, A.z);
        float nx, ny, nz;
        cross_ref(e1x, e1y, e1z, e2x, e2y, e2z, nx, ny, nz);
        normalize_ref(nx, ny, nz);
        normals[3 * r] = nx;
        normals[3 * r + 1] = ny;
        normals[3 * r + 2] = nz;
        tri_mesh_id[r] = 0;
        tri_id[r] = ti;
        const float uu = u[r], vv = v[r];
        bary[3 * r] = __fsub_rn(1.0f, __fadd_rn(uu, vv));
        bary[3 * r + 1] = uu;
        bary[3 * r + 2] = vv;
    } else {
        normals[3 * r] = normals[3 * r + 1] = normals[3 * r + 2] = 0.f;
        bary[3 * r] = bary[3 * r + 1] = bary[3 * r + 2] = 0.f;
        tri_mesh_id[r] = -1;
        tri_id[r] = -1;
    }
}

// per-sample unit face normals for packed hits (appearance stage): normal of triangle `tri[s]` of layer `layer[s]`
__global__ void __launch_bounds__(256) shells_normals_kernel(const Layer* __restrict__ layers, const int32_t* __restrict__ layer_of,
                                                             const int32_t* __restrict__ tri, int64_t n_samples,
                                                             const int64_t* __restrict__ n_valid_dev, float* __restrict__ normals) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_samples) return;
    if (n_valid_dev != nullptr && s >= *n_valid_dev) return;
    const Layer L = layers[layer_of[s]];
    const int64_t p = L.orig_to_bvh[tri[s]];
    const float4 A = L.tris[3 * p], Bv = L.tris[3 * p + 1], C = L.tris[3 * p + 2];
    const float e1x = __fsub_rn(Bv.x, A.x), e1y = __fsub_rn(Bv.y, A.y), e1z = __fsub_rn(Bv.z, A.z);
    const float e2x = __fsub_rn(C.x, A.x), e2y = __fsub_rn(C.y, A.y), e2z = __fsub_rn(C.z, A.z);
    float nx, ny, nz;
    cross_ref(e1x, e1y, e1z, e2x, e2y, e2z, nx, ny, nz);
    normalize_ref(nx, ny, nz);
    normals[3 * s] = nx;
    normals[3 * s + 1] = ny;
    normals[3 * s + 2] = nz;
}

static void free_layer(Layer& L) {
    if (L.nodes) cudaFree(L.nodes);
    if (L.tris) cudaFree(L.tris);
    if (L.pre) cudaFree(L.pre);
    if (L.orig_to_bvh) cudaFree(L.orig_to_bvh);
    L = Layer();
}

// Texture coordinates of packed hits (volsurfs_py/methods/volsurfs.py:509-516: uv = sum_j barycentric_j * face_uv_j with barycentric =
// (1-(u+v), u, v) as raytracelib returns it), evaluated in torch's order: ((b0*uv0) + (b1*uv1)) + (b2*uv2) per component, separate
// roundings.  face_uvs: [sum_l F_l, 3, 2] per-face-vertex uvs of all layers back to back, face_offset[l]: first face row of layer l.
__global__ void __launch_bounds__(256) shells_uvs_kernel(const int32_t* __restrict__ layer_of, const int32_t* __restrict__ tri,
                                                         const float* __restrict__ bary_uv, const float* __restrict__ face_uvs,
                                                         const int32_t* __restrict__ face_offset, int64_t n_samples,
                                                         const int64_t* __restrict__ n_valid_dev, float* __restrict__ out_uv) {
    int64_t n = n_samples;
    if (n_valid_dev != nullptr) n = min(n, *n_valid_dev);
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const float2 b = __ldg(reinterpret_cast<const float2*>(bary_uv) + s);
    const float b0 = __fsub_rn(1.f, __fadd_rn(b.x, b.y));  // raytracelib: Vector3f(1 - (u + v), u, v), src/bvh.cu:460
    const int64_t row = (int64_t)__ldg(face_offset + __ldg(layer_of + s)) + __ldg(tri + s);
    const float2* f = reinterpret_cast<const float2*>(face_uvs) + 3 * row;
    const float2 u0 = __ldg(f), u1 = __ldg(f + 1), u2 = __ldg(f + 2);
    float2 o;
    o.x = __fadd_rn(__fadd_rn(__fmul_rn(b0, u0.x), __fmul_rn(b.x, u1.x)), __fmul_rn(b.y, u2.x));
    o.y = __fadd_rn(__fadd_rn(__fmul_rn(b0, u0.y), __fmul_rn(b.x, u1.y)), __fmul_rn(b.y, u2.y));
    reinterpret_cast<float2*>(out_uv)[s] = o;
}

}  // namespace vs

using namespace vs;

extern "C" {

// Builds the K per-layer BVHs on the HOST from HOST arrays (like raytracelib's create_raytracer, src/raytracer.cu:23-45, which
// takes numpy arrays) and uploads them to the current device.  verts[k]: float32 [n_verts[k],3]; faces[k]: int32 [n_faces[k],3].
// The handle owns device memory; free it with vs_shells_free.  Synchronous (cudaMemcpy).
int vs_shells_build(int K, const float* const* verts, const int64_t* n_verts, const int32_t* const* faces, const int64_t* n_faces,
                    void** handle_out) {
    VS_CHECK_ARG(K > 0 && verts && n_verts && faces && n_faces && handle_out);
    for (int k = 0; k < K; ++k) {
        VS_CHECK_ARG(verts[k] && faces[k] && n_verts[k] > 0 && n_faces[k] > 0);
        if (n_faces[k] >= (1 << 27)) return VS_ERR_UNSUPPORTED;
        for (int64_t i = 0; i < 3 * n_faces[k]; ++i) VS_CHECK_ARG(faces[k][i] >= 0 && faces[k][i] < n_verts[k]);
    }
    Shells* S = new Shells();
    S->K = K;
    S->layers.resize(K);
    cudaError_t err = cudaSuccess;
    for (int k = 0; k < K && err == cudaSuccess; ++k) {
        Builder B;
        B.verts = verts[k];
        B.faces = faces[k];
        B.n_faces = n_faces[k];
        B.build();
        std::vector<Node4> wide;
        collapse_bfs(B, wide);
        const int64_t T = n_faces[k];
        std::vector<float4> tris(3 * T);
        std::vector<int32_t> o2b(T);
        for (int64_t p = 0; p < T; ++p) {
            const int32_t oi = B.order[p];
            o2b[oi] = (int32_t)p;
            const float* a = verts[k] + 3 * (int64_t)faces[k][3 * oi];
            const float* b = verts[k] + 3 * (int64_t)faces[k][3 * oi + 1];
            const float* c = verts[k] + 3 * (int64_t)faces[k][3 * oi + 2];
            float w;
            std::memcpy(&w, &oi, 4);
            tris[3 * p] = make_float4(a[0], a[1], a[2], w);
            tris[3 * p + 1] = make_float4(b[0], b[1], b[2], 0.f);
            tris[3 * p + 2] = make_float4(c[0], c[1], c[2], 0.f);
        }
        Layer& L = S->layers[k];
        L.n_nodes = (int32_t)wide.size();
        L.n_tris = T;
        if ((err = cudaMalloc(&L.nodes, sizeof(Node4) * std::max<size_t>(wide.size(), 1))) != cudaSuccess) break;
        if ((err = cudaMalloc(&L.tris, sizeof(float4) * 3 * T)) != cudaSuccess) break;
        if ((err = cudaMalloc(&L.orig_to_bvh, sizeof(int32_t) * T)) != cudaSuccess) break;
        if ((err = cudaMemcpy(L.nodes, wide.data(), sizeof(Node4) * wide.size(), cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((err = cudaMemcpy(L.tris, tris.data(), sizeof(float4) * 3 * T, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((err = cudaMalloc(&L.pre, sizeof(float4) * 4 * T)) != cudaSuccess) break;
        precompute_tri_kernel<<<(unsigned)div_up(T, 256), 256>>>(L.tris, L.pre, T);
        if ((err = cudaDeviceSynchronize()) != cudaSuccess) break;
        if ((err = cudaMemcpy(L.orig_to_bvh, o2b.data(), sizeof(int32_t) * T, cudaMemcpyHostToDevice)) != cudaSuccess) break;
    }
    if (err == cudaSuccess) err = cudaMalloc(&S->overflow_dev, (1 + kCounterSets * kTraceMaxLayers) * sizeof(int));
    if (err == cudaSuccess) err = cudaMemset(S->overflow_dev, 0, (1 + kCounterSets * kTraceMaxLayers) * sizeof(int));
    if (err == cudaSuccess) err = cudaMalloc(&S->layers_dev, sizeof(Layer) * K);
    if (err == cudaSuccess) err = cudaMemcpy(S->layers_dev, S->layers.data(), sizeof(Layer) * K, cudaMemcpyHostToDevice);
    if (err != cudaSuccess) {
        for (auto& L : S->layers) free_layer(L);
        if (S->layers_dev) cudaFree(S->layers_dev);
        if (S->overflow_dev) cudaFree(S->overflow_dev);
        delete S;
        return (int)err;
    }
    *handle_out = S;
    return VS_OK;
}

int vs_shells_free(void* handle) {
    if (!handle) return VS_OK;
    Shells* S = reinterpret_cast<Shells*>(handle);
    for (auto& L : S->layers) free_layer(L);
    if (S->layers_dev) cudaFree(S->layers_dev);
    if (S->overflow_dev) cudaFree(S->overflow_dev);
    delete S;
    return VS_OK;
}

// 1 if any traversal since the build dropped a stack entry (results would be unreliable), else 0.  Synchronises the device.
int vs_shells_overflowed(const void* handle) {
    if (!handle) return 0;
    int v = 0;
    cudaMemcpy(&v, reinterpret_cast<const Shells*>(handle)->overflow_dev, sizeof(int), cudaMemcpyDeviceToHost);
    return v;
}

int vs_shells_info(const void* handle, int layer, int64_t* n_nodes, int64_t* n_tris) {
    VS_CHECK_ARG(handle);
    const Shells* S = reinterpret_cast<const Shells*>(handle);
    VS_CHECK_ARG(layer >= 0 && layer < S->K);
    if (n_nodes) *n_nodes = S->layers[layer].n_nodes;
    if (n_tris) *n_tris = S->layers[layer].n_tris;
    return VS_OK;
}

int vs_shells_num_layers(const void* handle) { return handle ? reinterpret_cast<const Shells*>(handle)->K : 0; }

// Nearest hit of every ray with layers [layer_first, layer_first+layer_count).  Outputs are layer-major
// [layer_count, n_rays]: depth (1e6 = miss), original face index (-1 = miss), barycentric u, v (0 on a miss).
int vs_shells_trace(const void* handle, const float* rays_o, const float* rays_d, int64_t n_rays, int layer_first, int layer_count,
                    float* depth_out, int32_t* tri_out, float* u_out, float* v_out, void* stream) {
    VS_CHECK_ARG(handle && n_rays >= 0);
    Shells* S = const_cast<Shells*>(reinterpret_cast<const Shells*>(handle));
    VS_CHECK_ARG(layer_first >= 0 && layer_count > 0 && layer_first + layer_count <= S->K && layer_count <= kTraceMaxLayers);
    if (n_rays == 0) return VS_OK;
    VS_CHECK_ARG(rays_o && rays_d && depth_out && tri_out && u_out && v_out);
    VS_CHECK_ARG(div_up(n_rays, 32) < 0x7fffff00);   // chunk numbers are ints (the counter runs a little past the last chunk)
    int sms = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int64_t chunks = div_up(n_rays, 32);
    const int ctas_per_layer = (int)std::max<int64_t>(1, std::min<int64_t>((sms * kTraceCtasPerSm + layer_count - 1) / layer_count,
                                                                             div_up(chunks, kTraceThreads / 32)));
    dim3 grid((unsigned)ctas_per_layer, (unsigned)layer_count);
    int* work_counter = S->overflow_dev + 1 + (S->launches++ % kCounterSets) * kTraceMaxLayers;
    if (cudaMemsetAsync(work_counter, 0, layer_count * sizeof(int), (cudaStream_t)stream) != cudaSuccess) return (int)cudaGetLastError();
    // traversal variants kept for A/B measurements (profiles/): 1 = default (per-entry pop, precomputed triangle edges/normal);
    // 0 = raw vertices; 2/3 = "while-while" traversal (measured slower on B200: 0.86 vs 0.72 ms on the C2 scene)
    static const int variant = getenv("VS_TRACE_VARIANT") ? atoi(getenv("VS_TRACE_VARIANT")) : 1;
#define VS_TRACE(WW, PRE)                                                                                                               \
    shells_trace_kernel<WW, PRE><<<grid, kTraceThreads, 0, (cudaStream_t)stream>>>(S->layers_dev, layer_first, rays_o, rays_d, n_rays, \
                                                                                    depth_out, tri_out, u_out, v_out, S->overflow_dev,    \
                                                                                    layer_count, work_counter)
    switch (variant) {
        case 0: VS_TRACE(false, false); break;
        case 2: VS_TRACE(true, false); break;
        case 3: VS_TRACE(true, true); break;
        default: VS_TRACE(false, true); break;
    }
#undef VS_TRACE
    return launched(1);
}

// Reference-format result of RayTracer.trace(mesh_id=layer): positions [N,3], unit face normals [N,3], triangles_mesh_id /
// triangles_id int64 [N], barycentric (1-u-v,u,v) [N,3], from one layer's compact hit record.
int vs_shells_expand(const void* handle, int layer, const float* rays_o, const float* rays_d, const float* depth, const int32_t* tri,
                     const float* u, const float* v, int64_t n_rays, float* positions, float* normals, int64_t* tri_mesh_id,
                     int64_t* tri_id, float* barycentric, void* stream) {
    VS_CHECK_ARG(handle && n_rays >= 0);
    const Shells* S = reinterpret_cast<const Shells*>(handle);
    VS_CHECK_ARG(layer >= 0 && layer < S->K);
    if (n_rays == 0) return VS_OK;
    VS_CHECK_ARG(rays_o && rays_d && depth && tri && u && v && positions && normals && tri_mesh_id && tri_id && barycentric);
    shells_expand_kernel<<<(unsigned)div_up(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(
        S->layers_dev, layer, rays_o, rays_d, depth, tri, u, v, n_rays, positions, normals, tri_mesh_id, tri_id, barycentric);
    return launched(1);
}

// Unit face normals of packed hits: normals[s] = normal of face tri[s] of layer layer_of[s].  n_valid_dev (optional, device
// int64) limits the work to the first *n_valid_dev samples of capacity-sized arrays.
int vs_shells_sample_normals(const void* handle, const int32_t* layer_of, const int32_t* tri, int64_t n_samples,
                             const int64_t* n_valid_dev, float* normals, void* stream) {
    VS_CHECK_ARG(handle && n_samples >= 0);
    if (n_samples == 0) return VS_OK;
    VS_CHECK_ARG(layer_of && tri && normals);
    const Shells* S = reinterpret_cast<const Shells*>(handle);
    shells_normals_kernel<<<(unsigned)div_up(n_samples, 256), 256, 0, (cudaStream_t)stream>>>(S->layers_dev, layer_of, tri, n_samples,
                                                                                              n_valid_dev, normals);
    return launched(1);
}

// Texture coordinates of packed hits from their barycentric (u, v): see shells_uvs_kernel.  All pointers DEVICE.
int vs_shells_sample_uvs(const int32_t* layer_of, const int32_t* tri, const float* bary_uv, const float* face_uvs, const int32_t* face_offset,
                         int64_t n_samples, const int64_t* n_valid_dev, float* out_uv, void* stream) {
    VS_CHECK_ARG(n_samples >= 0);
    if (n_samples == 0) return VS_OK;
    VS_CHECK_ARG(layer_of && tri && bary_uv && face_uvs && face_offset && out_uv);
    shells_uvs_kernel<<<(unsigned)div_up(n_samples, 256), 256, 0, (cudaStream_t)stream>>>(layer_of, tri, bary_uv, face_uvs, face_offset, n_samples,
                                                                                          n_valid_dev, out_uv);
    return launched(1);
}

}  // extern "C"
