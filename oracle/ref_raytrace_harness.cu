// ORACLE (test infrastructure only — never linked into or loaded by the product).
//
// A thin C harness around the REFERENCE's own mesh ray tracer: it #includes
//   /root/reference/submodules/raytracelib/src/bvh.cu                         (BVH build, TriangleBvh4::ray_intersect, raytrace_kernel)
//   /root/reference/submodules/raytracelib/include/raytracing/{triangle,bounding_box,bvh}.cuh, common.h, gpu_memory.h  (via bvh.cu)
// unmodified, where they lie.  Those sources need Eigen 3.3.7 (downloaded by raytracelib's setup.py:19-63; no network here): the
// include path carries oracle/eigen_standin/Eigen/Dense instead — the harness's own minimal fixed-size vector types written in Eigen
// 3.3.7's evaluation order (see that header).  Compiled with nvcc's default floating-point flags (FMA contraction on), as
// raytracelib's own build does (setup.py passes -O3 only), for sm_100a.
//
// What is restated here (not included): the five lines of RayTracerImpl's constructor (src/raytracer.cu:23-45) that turn vertex / face
// arrays into `Triangle{a, b, c, idx = i, mesh_idx = 0}`, build the BVH with 8 primitives per leaf and upload the triangles —
// raytracer.cu itself needs libtorch and Eigen's dynamic matrices.
//
// Two ways in: the reference's CUDA kernel (`ref_rt_trace`, launch shape of common.h:38-46: 128 threads per block) and the very same
// `__host__ __device__` traversal run on the host by gcc (`ref_rt_trace_host`, no FMA contraction on x86-64 without -mfma) — the
// latter pins oracle/raytrace_oracle.c in the CPU-only test suite.
#include <algorithm>
#include <array>
#include <cassert>
#include <cstdio>
#include <limits>
#include <memory>
#include <tuple>
#include <vector>

#include "bvh.cu"  // found through -I /root/reference/submodules/raytracelib/src

namespace {

using raytracing::Triangle;
using raytracing::TriangleBvh4;
using raytracing::TriangleBvhNode;

struct Bvh4Access : public TriangleBvh4 {
    const std::vector<TriangleBvhNode>& host_nodes() const { return m_nodes; }
};

struct RefTracer {
    std::vector<Triangle> triangles_cpu;
    raytracing::GPUMemory<Triangle> triangles_gpu;
    Bvh4Access bvh;
    bool on_gpu = false;
};

}  // namespace

extern "C" {

int ref_rt_abi_version() { return 1; }

// src/raytracer.cu:23-45.  `want_gpu` = 0 builds the host copy only (the CPU-only test suite has no device).
void* ref_rt_create(const float* verts, int n_verts, const uint32_t* faces, int n_faces, int want_gpu) {
    (void)n_verts;
    auto* rt = new RefTracer();
    rt->triangles_cpu.resize((size_t)n_faces);
    for (int i = 0; i < n_faces; ++i) {
        const float* a = verts + 3 * (size_t)faces[3 * i + 0];
        const float* b = verts + 3 * (size_t)faces[3 * i + 1];
        const float* c = verts + 3 * (size_t)faces[3 * i + 2];
        rt->triangles_cpu[i] = {Eigen::Vector3f(a[0], a[1], a[2]), Eigen::Vector3f(b[0], b[1], b[2]), Eigen::Vector3f(c[0], c[1], c[2]),
                                (int64_t)i, 0};
    }
    try {
        rt->bvh.build(rt->triangles_cpu, 8);  // fills m_nodes, then uploads them (throws without a device, after the host build)
        rt->on_gpu = true;
    } catch (const std::exception& e) {
        if (want_gpu) {
            std::fprintf(stderr, "ref_rt_create: %s\n", e.what());
            delete rt;
            return nullptr;
        }
        cudaGetLastError();
    }
    if (want_gpu) {
        try {
            rt->triangles_gpu.resize_and_copy_from_host(rt->triangles_cpu);
        } catch (const std::exception& e) {
            std::fprintf(stderr, "ref_rt_create: %s\n", e.what());
            delete rt;
            return nullptr;
        }
    }
    return rt;
}

void ref_rt_free(void* h) { delete static_cast<RefTracer*>(h); }

int ref_rt_num_nodes(void* h) { return (int)static_cast<RefTracer*>(h)->bvh.host_nodes().size(); }

// raytracer.cu:48-69 -> bvh.cu:265-307 -> raytrace_kernel (bvh.cu:420-469).  All pointers are device pointers.
int ref_rt_trace(void* h, const float* rays_o, const float* rays_d, const float* min_depth, float* positions, float* normals, float* depth,
                 int64_t* triangles_mesh_id, int64_t* triangles_id, float* barycentric, int n_rays) {
    auto* rt = static_cast<RefTracer*>(h);
    if (!rt->on_gpu) return -1;
    rt->bvh.ray_trace_gpu((uint32_t)n_rays, rays_o, rays_d, min_depth, positions, normals, depth, triangles_mesh_id, triangles_id, barycentric,
                          rt->triangles_gpu.data(), 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    return (int)e;
}

// The body of raytrace_kernel (bvh.cu:438-468) around the same TriangleBvh4::ray_intersect, executed on the host.  Host pointers.
int ref_rt_trace_host(void* h, const float* rays_o, const float* rays_d, const float* min_depth, float* positions, float* normals, float* depth,
                      int64_t* triangles_mesh_id, int64_t* triangles_id, float* barycentric, int n_rays) {
    auto* rt = static_cast<RefTracer*>(h);
    const TriangleBvhNode* nodes = rt->bvh.host_nodes().data();
    const Triangle* tris = rt->triangles_cpu.data();
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n_rays; ++i) {
        Eigen::Vector3f ray_o(rays_o[3 * i], rays_o[3 * i + 1], rays_o[3 * i + 2]);
        Eigen::Vector3f ray_d(rays_d[3 * i], rays_d[3 * i + 1], rays_d[3 * i + 2]);
        auto res = TriangleBvh4::ray_intersect(ray_o, ray_d, min_depth[i], nodes, tris);
        depth[i] = std::get<2>(res);
        Eigen::Vector3f p = ray_o + depth[i] * ray_d;
        Eigen::Vector3f n, bc;
        if (std::get<0>(res) >= 0) {
            triangles_mesh_id[i] = tris[std::get<0>(res)].mesh_idx;
            triangles_id[i] = tris[std::get<1>(res)].idx;
            n = tris[std::get<1>(res)].normal();
            const float u = std::get<3>(res), v = std::get<4>(res);
            bc = Eigen::Vector3f(1 - (u + v), u, v);
        } else {
            triangles_mesh_id[i] = -1;
            triangles_id[i] = -1;
        }
        for (int k = 0; k < 3; ++k) {
            positions[3 * i + k] = p[k];
            normals[3 * i + k] = n[k];
            barycentric[3 * i + k] = bc[k];
        }
    }
    return 0;
}

}  // extern "C"
