/* ORACLE (test infrastructure only — never linked into or called by the product).
 *
 * Plain-C restatement of the reference's mesh ray tracer (submodules/raytracelib, vendored at a3a5e4e):
 *   triangle test ......... include/raytracing/triangle.cuh:42-70   (Triangle::ray_intersect)
 *   slab test ............. include/raytracing/bounding_box.cuh:151-198 (BoundingBox::ray_intersect)
 *   BVH4 build ............ src/bvh.cu:309-408  (median split on the max-variance centroid axis, <= 8 tris/leaf)
 *   BVH4 traversal ........ src/bvh.cu:186-263  (FixedStack<int,32>, children sorted far->near, pushed if t_near < t)
 *   per-ray outputs ....... src/bvh.cu:420-469  (depth, position, face normal, original triangle id, barycentric)
 *   hit flag .............. raytracelib/raytracer.py:100 (is_hit = depth <= t_far), min_depth = 0 (:70)
 *
 * PINNED by the reference's own sources: oracle/ref_raytrace_harness.cu compiles src/bvh.cu + the raytracing/ headers where they lie
 * (against oracle/eigen_standin: Eigen itself is downloaded by raytracelib's setup.py and absent here) into
 * oracle/_ref/libraytrace_ref.so, which runs the reference's traversal both on the host (gcc) and as its CUDA kernel.
 *
 * Two arithmetic contracts, selected at run time by vso_set_contract():
 *   0  "host": IEEE fp32, round-to-nearest, NO contraction (this file is compiled with -ffp-contract=off) in Eigen 3.3.7's
 *      evaluation order — bit-identical to the reference's __host__ code path built by gcc for x86-64 (tests/test_oracle_raytrace.py):
 *        cross(a,b) = (a1*b2 - a2*b1, a2*b0 - a0*b2, a0*b1 - a1*b0)
 *        dot(a,b)   = a0*b0 + (a1*b1 + a2*b2)        (redux_novec_unroller splits a length-3 reduction as 1 + 2)
 *        normalized = v / sqrt(dot(v,v)) component-wise division, v unchanged if dot(v,v) == 0
 *   1  "device": the same expressions with the FMA contractions nvcc 12.9 -O3 applies to raytrace_kernel for sm_100a, read off the SASS
 *      of libraytrace_ref.so (cuobjdump -sass) — bit-identical to the reference's CUDA kernel (tests/test_gpu_raytrace.py), and the
 *      contract volsurfs_b200/csrc/shells.cu implements:
 *        cross(a,b)_i = fma(a_j, b_k, -rn(a_k*b_j))
 *        d.n, n.rov0, n.n  = fma(x0,y0, fma(x2,y2, rn(x1*y1)))
 *        q.v2v0, q.v1v0    = fma(x0,y0, fma(x1,y1, rn(x2*y2)))
 *        D = 1/(d.n) correctly rounded; u, v, t = rn(D * .);  position = fma(t, d, o);  sqrt and divisions correctly rounded
 * Ties (two triangles with bit-identical t) are resolved by traversal order in the reference; the brute-force tracer
 * below visits triangles in original index order, so the lowest index wins — the rule the CUDA kernel implements.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAX_DIST 1e6f /* include/raytracing/common.h:21 */
#define BRANCH 4
#define LEAF_TRIS 8 /* src/raytracer.cu:42: build(triangles_cpu, 8) */
#define STACK_SIZE 32

typedef struct {
    float a[3], b[3], c[3];
    int64_t idx; /* index into the original face list (src/raytracer.cu:34) */
} Tri;

typedef struct {
    float bmin[3], bmax[3];
    int left, right; /* leaf: left = -first-1, right = -end-1 (bvh.cu:394-395); inner: [left,right) children */
} Node;

typedef struct {
    Tri* tris; /* reordered by the build */
    int64_t n_tris;
    Node* nodes;
    int n_nodes, cap_nodes;
} Bvh;

/* ---- 3-vector helpers in Eigen order -------------------------------------------------------------------------- */
static inline void sub3(const float* a, const float* b, float* r) {
    r[0] = a[0] - b[0];
    r[1] = a[1] - b[1];
    r[2] = a[2] - b[2];
}
static inline void cross3(const float* a, const float* b, float* r) {
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}
static inline float dot3(const float* a, const float* b) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }

/* ---- contract 1: the reference kernel's FMA contractions (see the header) ---------------------------------------------- */
static int g_contract = 0;
void vso_set_contract(int c) { g_contract = c; }
int vso_get_contract(void) { return g_contract; }

/* element-wise fmaf for the numpy side of the oracle (oracle/packing.py): out = a*b + c with one rounding */
void vso_fmaf_array(const float* a, const float* b, const float* c, float* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = fmaf(a[i], b[i], c[i]);
}

static inline void cross3_dev(const float* a, const float* b, float* r) {
    r[0] = fmaf(a[1], b[2], -(a[2] * b[1]));
    r[1] = fmaf(a[2], b[0], -(a[0] * b[2]));
    r[2] = fmaf(a[0], b[1], -(a[1] * b[0]));
}
/* dots that involve the face normal n: the middle product is the rounded one */
static inline float dot3_dev_n(const float* a, const float* b) { return fmaf(a[0], b[0], fmaf(a[2], b[2], a[1] * b[1])); }
/* dots of q with an edge: the last product is the rounded one */
static inline float dot3_dev_q(const float* a, const float* b) { return fmaf(a[0], b[0], fmaf(a[1], b[1], a[2] * b[2])); }

/* triangle.cuh:42-70.  Returns is_hit; t,u,v as the reference leaves them (t = -1 on a miss). */
static int tri_intersect(const Tri* tr, const float* o, const float* d, float* t_out, float* u_out, float* v_out) {
    float v1v0[3], v2v0[3], rov0[3], n[3], q[3];
    sub3(tr->b, tr->a, v1v0);
    sub3(tr->c, tr->a, v2v0);
    sub3(o, tr->a, rov0);
    float u, v, t;
    if (g_contract == 1) {
        cross3_dev(v1v0, v2v0, n);
        cross3_dev(rov0, d, q);
        float D = 1.0f / dot3_dev_n(d, n);
        u = D * -dot3_dev_q(q, v2v0);
        v = D * dot3_dev_q(q, v1v0);
        t = D * -dot3_dev_n(n, rov0);
    } else {
        cross3(v1v0, v2v0, n);
        cross3(rov0, d, q);
        float D = 1.0f / dot3(d, n);
        u = D * -dot3(q, v2v0);
        v = D * dot3(q, v1v0);
        t = D * -dot3(n, rov0);
    }
    int is_hit = 1;
    if (u < 0.0f || u > 1.0f || v < 0.0f || (u + v) > 1.0f || t < 0.0f) {
        is_hit = 0;
        t = -1.0f;
    }
    *t_out = t;
    *u_out = u;
    *v_out = v;
    return is_hit;
}

/* bounding_box.cuh:151-198, returns t_near (FLT_MAX on a miss) */
static float box_tnear(const Node* nd, const float* o, const float* d) {
    float tmin = (nd->bmin[0] - o[0]) / d[0];
    float tmax = (nd->bmax[0] - o[0]) / d[0];
    if (tmin > tmax) { float s = tmin; tmin = tmax; tmax = s; }
    float tymin = (nd->bmin[1] - o[1]) / d[1];
    float tymax = (nd->bmax[1] - o[1]) / d[1];
    if (tymin > tymax) { float s = tymin; tymin = tymax; tymax = s; }
    if (tmin > tymax || tymin > tmax) return FLT_MAX;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = (nd->bmin[2] - o[2]) / d[2];
    float tzmax = (nd->bmax[2] - o[2]) / d[2];
    if (tzmin > tzmax) { float s = tzmin; tzmin = tzmax; tzmax = s; }
    if (tmin > tzmax || tzmin > tmax) return FLT_MAX;
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    (void)tmax;
    return tmin;
}

/* ---- build (bvh.cu:309-408) -------------------------------------------------------------------------------------- */
static inline float centroid_axis(const Tri* t, int ax) { return (t->a[ax] + t->b[ax] + t->c[ax]) / 3; }

/* std::nth_element stand-in: after the call tris[m] is the element a full sort would put there and everything
 * before it compares <=.  (The permutation inside the halves is implementation-defined in libstdc++ too; the
 * nearest hit does not depend on it.) */
static void nth_element_axis(Tri* tris, int64_t lo, int64_t hi, int64_t m, int ax) {
    while (hi - lo > 1) {
        int64_t mid = lo + (hi - lo) / 2;
        float pv = centroid_axis(&tris[mid], ax);
        int64_t i = lo, j = hi - 1;
        while (i <= j) {
            while (centroid_axis(&tris[i], ax) < pv) ++i;
            while (centroid_axis(&tris[j], ax) > pv) --j;
            if (i <= j) {
                Tri tmp = tris[i];
                tris[i] = tris[j];
                tris[j] = tmp;
                ++i;
                --j;
            }
        }
        if (m <= j) hi = j + 1;
        else if (m >= i) lo = i;
        else return;
    }
}

static void range_bbox(const Tri* tris, int64_t lo, int64_t hi, Node* nd) {
    for (int k = 0; k < 3; ++k) nd->bmin[k] = nd->bmax[k] = tris[lo].a[k];
    for (int64_t i = lo; i < hi; ++i) {
        const float* vs[3] = {tris[i].a, tris[i].b, tris[i].c};
        for (int p = 0; p < 3; ++p)
            for (int k = 0; k < 3; ++k) {
                if (vs[p][k] < nd->bmin[k]) nd->bmin[k] = vs[p][k];
                if (vs[p][k] > nd->bmax[k]) nd->bmax[k] = vs[p][k];
            }
    }
}

static int new_node(Bvh* b) {
    if (b->n_nodes == b->cap_nodes) {
        b->cap_nodes = b->cap_nodes ? 2 * b->cap_nodes : 1024;
        b->nodes = (Node*)realloc(b->nodes, sizeof(Node) * (size_t)b->cap_nodes);
    }
    memset(&b->nodes[b->n_nodes], 0, sizeof(Node));
    return b->n_nodes++;
}

typedef struct {
    int node;
    int64_t lo, hi;
} BuildItem;

static int max_variance_axis(const Tri* tris, int64_t lo, int64_t hi) {
    float mean[3] = {0, 0, 0}, var[3] = {0, 0, 0};
    float cnt = (float)(hi - lo);
    for (int64_t i = lo; i < hi; ++i)
        for (int k = 0; k < 3; ++k) mean[k] += (tris[i].a[k] + tris[i].b[k] + tris[i].c[k]) / 3.0f;
    for (int k = 0; k < 3; ++k) mean[k] /= cnt;
    for (int64_t i = lo; i < hi; ++i)
        for (int k = 0; k < 3; ++k) {
            float df = (tris[i].a[k] + tris[i].b[k] + tris[i].c[k]) / 3.0f - mean[k];
            var[k] += df * df;
        }
    int ax = 0; /* Eigen maxCoeff: first maximum */
    if (var[1] > var[ax]) ax = 1;
    if (var[2] > var[ax]) ax = 2;
    return ax;
}

void* vso_bvh_build(const float* verts, const int32_t* faces, int64_t n_faces) {
    Bvh* b = (Bvh*)calloc(1, sizeof(Bvh));
    b->n_tris = n_faces;
    b->tris = (Tri*)malloc(sizeof(Tri) * (size_t)n_faces);
    for (int64_t i = 0; i < n_faces; ++i) {
        for (int k = 0; k < 3; ++k) {
            b->tris[i].a[k] = verts[3 * (int64_t)faces[3 * i] + k];
            b->tris[i].b[k] = verts[3 * (int64_t)faces[3 * i + 1] + k];
            b->tris[i].c[k] = verts[3 * (int64_t)faces[3 * i + 2] + k];
        }
        b->tris[i].idx = i;
    }
    int root = new_node(b);
    range_bbox(b->tris, 0, n_faces, &b->nodes[root]);
    int64_t cap = 64, top = 0;
    BuildItem* stack = (BuildItem*)malloc(sizeof(BuildItem) * (size_t)cap);
    stack[top++] = (BuildItem){root, 0, n_faces};
    while (top > 0) {
        BuildItem cur = stack[--top];
        int64_t lo[BRANCH], hi[BRANCH];
        lo[0] = cur.lo;
        hi[0] = cur.hi;
        int n_children = 1;
        while (n_children < BRANCH) {
            for (int i = n_children - 1; i >= 0; --i) {
                int ax = max_variance_axis(b->tris, lo[i], hi[i]);
                int64_t m = lo[i] + (hi[i] - lo[i]) / 2;
                nth_element_axis(b->tris, lo[i], hi[i], m, ax);
                int64_t l = lo[i], h = hi[i];
                lo[2 * i] = l;
                hi[2 * i + 1] = h;
                hi[2 * i] = lo[2 * i + 1] = m;
            }
            n_children *= 2;
        }
        int first_child = b->n_nodes;
        for (int i = 0; i < BRANCH; ++i) {
            int ci = new_node(b);
            range_bbox(b->tris, lo[i], hi[i], &b->nodes[ci]);
            if (hi[i] - lo[i] <= LEAF_TRIS) {
                b->nodes[ci].left = -(int)lo[i] - 1;
                b->nodes[ci].right = -(int)hi[i] - 1;
            } else {
                if (top + 1 >= cap) {
                    cap *= 2;
                    stack = (BuildItem*)realloc(stack, sizeof(BuildItem) * (size_t)cap);
                }
                stack[top++] = (BuildItem){ci, lo[i], hi[i]};
            }
        }
        b->nodes[cur.node].left = first_child;
        b->nodes[cur.node].right = b->n_nodes;
    }
    free(stack);
    return b;
}

void vso_bvh_free(void* h) {
    Bvh* b = (Bvh*)h;
    if (!b) return;
    free(b->tris);
    free(b->nodes);
    free(b);
}

int vso_bvh_num_nodes(const void* h) { return ((const Bvh*)h)->n_nodes; }

/* ---- traversal (bvh.cu:186-263) --------------------------------------------------------------------------------- */
typedef struct {
    float dist;
    int idx;
} DistIdx;

static inline void cas(DistIdx* x, DistIdx* y) { /* compare_and_swap: swaps when x < y => descending (bvh.cu:46-54) */
    if (x->dist < y->dist) {
        DistIdx t = *x;
        *x = *y;
        *y = t;
    }
}

static void bvh_intersect(const Bvh* b, const float* o, const float* d, float min_t, int* tri_out, float* t_out, float* u_out,
                          float* v_out, int* overflow) {
    int stack[STACK_SIZE];
    int top = 0;
    stack[top++] = 0;
    float curr_t = MAX_DIST, cu = 0.0f, cv = 0.0f;
    int tri_idx = -1;
    while (top > 0) {
        const Node* nd = &b->nodes[stack[--top]];
        if (nd->left < 0) {
            int end = -nd->right - 1;
            for (int i = -nd->left - 1; i < end; ++i) {
                float t, u, v;
                int is_hit = tri_intersect(&b->tris[i], o, d, &t, &u, &v);
                if (is_hit && t > min_t && t < curr_t) {
                    curr_t = t;
                    tri_idx = i;
                    cu = u;
                    cv = v;
                }
            }
        } else {
            DistIdx ch[BRANCH];
            for (int i = 0; i < BRANCH; ++i) {
                ch[i].dist = box_tnear(&b->nodes[nd->left + i], o, d);
                ch[i].idx = nd->left + i;
            }
            cas(&ch[0], &ch[2]); /* sorting_network<4>, bvh.cu:86-93 */
            cas(&ch[1], &ch[3]);
            cas(&ch[0], &ch[1]);
            cas(&ch[2], &ch[3]);
            cas(&ch[1], &ch[2]);
            for (int i = 0; i < BRANCH; ++i) {
                if (ch[i].dist < curr_t) {
                    if (top >= STACK_SIZE - 1) { /* FixedStack only warns (bvh.cuh:24-31); flag it, do not corrupt memory */
                        *overflow = 1;
                        if (top >= STACK_SIZE) continue;
                    }
                    stack[top++] = ch[i].idx;
                }
            }
        }
    }
    *tri_out = tri_idx;
    *t_out = curr_t;
    *u_out = cu;
    *v_out = cv;
}

static void brute_intersect(const Bvh* b, const Tri* tris_in_index_order, const float* o, const float* d, float min_t, int* tri_out,
                            float* t_out, float* u_out, float* v_out) {
    (void)b;
    float curr_t = MAX_DIST, cu = 0.0f, cv = 0.0f;
    int tri_idx = -1;
    const int64_t n = b->n_tris;
    for (int64_t i = 0; i < n; ++i) {
        float t, u, v;
        int is_hit = tri_intersect(&tris_in_index_order[i], o, d, &t, &u, &v);
        if (is_hit && t > min_t && t < curr_t) {
            curr_t = t;
            tri_idx = (int)i;
            cu = u;
            cv = v;
        }
    }
    *tri_out = tri_idx;
    *t_out = curr_t;
    *u_out = cu;
    *v_out = cv;
}

/* raytrace_kernel outputs (bvh.cu:420-469) for one ray given the winning triangle */
static void write_outputs(const Tri* tr, int has, const float* o, const float* d, float t, float u, float v, int64_t r, float* positions,
                          float* normals, float* depth, int64_t* tri_mesh_id, int64_t* tri_id, float* bary) {
    depth[r] = t;
    for (int k = 0; k < 3; ++k) positions[3 * r + k] = g_contract == 1 ? fmaf(d[k], t, o[k]) : o[k] + t * d[k]; /* ray_o + depth*ray_d */
    if (has) {
        float e1[3], e2[3], n[3];
        sub3(tr->b, tr->a, e1);
        sub3(tr->c, tr->a, e2);
        float z;
        if (g_contract == 1) {
            cross3_dev(e1, e2, n);
            z = dot3_dev_n(n, n);
        } else {
            cross3(e1, e2, n);
            z = dot3(n, n);
        }
        if (z > 0.0f) {
            float s = sqrtf(z);
            n[0] = n[0] / s;
            n[1] = n[1] / s;
            n[2] = n[2] / s;
        }
        for (int k = 0; k < 3; ++k) normals[3 * r + k] = n[k];
        tri_mesh_id[r] = 0;
        tri_id[r] = tr->idx;
        bary[3 * r] = 1 - (u + v);
        bary[3 * r + 1] = u;
        bary[3 * r + 2] = v;
    } else {
        for (int k = 0; k < 3; ++k) normals[3 * r + k] = 0.0f, bary[3 * r + k] = 0.0f;
        tri_mesh_id[r] = -1;
        tri_id[r] = -1;
    }
}

/* mode 0: reference-faithful BVH traversal; mode 1: brute force in original index order.
 * u_out / v_out (optional) receive the raw barycentric u, v of the winning triangle (0 on a miss).
 * Returns 1 if the reference's FixedStack<32> would have overflowed for some ray, else 0. */
int vso_trace(const void* h, int mode, const float* rays_o, const float* rays_d, const float* min_depth, int64_t n_rays, float* positions,
              float* normals, float* depth, int64_t* tri_mesh_id, int64_t* tri_id, float* bary, float* u_out, float* v_out) {
    const Bvh* b = (const Bvh*)h;
    Tri* by_index = NULL;
    if (mode == 1) {
        by_index = (Tri*)malloc(sizeof(Tri) * (size_t)b->n_tris);
        for (int64_t i = 0; i < b->n_tris; ++i) by_index[b->tris[i].idx] = b->tris[i];
    }
    int overflow = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(| : overflow)
    for (int64_t r = 0; r < n_rays; ++r) {
        const float* o = rays_o + 3 * r;
        const float* d = rays_d + 3 * r;
        float min_t = min_depth ? min_depth[r] : 0.0f;
        int ti;
        float t, u, v;
        int ov = 0;
        if (mode == 1) brute_intersect(b, by_index, o, d, min_t, &ti, &t, &u, &v);
        else bvh_intersect(b, o, d, min_t, &ti, &t, &u, &v, &ov);
        overflow |= ov;
        const Tri* tr = ti >= 0 ? (mode == 1 ? &by_index[ti] : &b->tris[ti]) : NULL;
        write_outputs(tr, ti >= 0, o, d, t, u, v, r, positions, normals, depth, tri_mesh_id, tri_id, bary);
        if (u_out) u_out[r] = ti >= 0 ? u : 0.0f;
        if (v_out) v_out[r] = ti >= 0 ? v : 0.0f;
    }
    free(by_index);
    return overflow;
}
