// volsurfs_b200 — library-level entry points of the C ABI (include/volsurfs_b200.h).
#include "vs_common.cuh"

namespace vs {
long long g_launches = 0;
}

extern "C" {

int vs_abi_version(void) { return 1; }

long long vs_launch_count(void) { return __atomic_load_n(&vs::g_launches, __ATOMIC_RELAXED); }

const char* vs_error_string(int code) {
    switch (code) {
        case VS_OK: return "ok";
        case VS_ERR_INVALID_ARG: return "volsurfs_b200: invalid argument (null pointer, negative size or bad mode)";
        case VS_ERR_UNSUPPORTED: return "volsurfs_b200: unsupported configuration (e.g. value dimension)";
        case VS_ERR_ALLOC: return "volsurfs_b200: allocation failed";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "volsurfs_b200: unknown error";
}

}  // extern "C"
