"""ORACLE build recipe (test infrastructure): compiles raytrace_oracle.c with gcc into oracle/_build/ (IEEE fp32, -ffp-contract=off:
every FMA in it is an explicit fmaf), and the reference's own kernels behind C harnesses into oracle/_ref/ — see DESIGN.md §4."""
from __future__ import annotations

import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "raytrace_oracle.c"
OUT_DIR = HERE / "_build"
LIB = OUT_DIR / "libraytrace_oracle.so"


def build(force: bool = False) -> Path:
    if LIB.exists() and not force and LIB.stat().st_mtime >= SRC.stat().st_mtime:
        return LIB
    gcc = shutil.which("gcc")
    if gcc is None:
        raise RuntimeError("gcc not found")
    OUT_DIR.mkdir(exist_ok=True)
    cmd = [gcc, "-O2", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC", str(SRC), "-o", str(LIB), "-lm"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"gcc failed:\n{res.stdout}")
    return LIB


PERMUTO_C_SRC = HERE / "permuto_oracle.c"
PERMUTO_C_LIB = OUT_DIR / "libpermuto_oracle.so"


def build_permuto_c(force: bool = False) -> Path:
    """gcc-compile oracle/permuto_oracle.c (the C twin of oracle/permuto.py used by bench.py's CPU baseline)"""
    if PERMUTO_C_LIB.exists() and not force and PERMUTO_C_LIB.stat().st_mtime >= PERMUTO_C_SRC.stat().st_mtime:
        return PERMUTO_C_LIB
    gcc = shutil.which("gcc")
    if gcc is None:
        raise RuntimeError("gcc not found")
    OUT_DIR.mkdir(exist_ok=True)
    cmd = [gcc, "-O2", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC", str(PERMUTO_C_SRC), "-o",
           str(PERMUTO_C_LIB), "-lm"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"gcc failed:\n{res.stdout}")
    return PERMUTO_C_LIB


# ---- oracle/_ref: the reference's own packed volume-rendering kernels behind a C harness -------------------------------------
REFERENCE = Path("/root/reference")
REF_SRC = HERE / "ref_harness.cu"
REF_DIR = HERE / "_ref"
REF_LIB = REF_DIR / "libvolsurfs_ref.so"


def ref_available() -> bool:
    return REF_LIB.exists()


def build_ref(force: bool = False):
    """nvcc-compile oracle/ref_harness.cu, which #includes kernels/volsurfs/VolumeRenderingGPU.cuh + pcg32.h from the sources where
    they lie under /root/reference (nothing is copied), for sm_100a into oracle/_ref/ (git-ignored, travels to the GPU box).
    Needs only the torch HEADERS (PackedTensorAccessor32); the result links against cudart alone.  Returns None when the
    reference tree is not mounted (GPU box): the prebuilt .so is used as is.  Takes ~2.5 min (torch/torch.h)."""
    if not (REFERENCE / "kernels/volsurfs/VolumeRenderingGPU.cuh").exists():
        return REF_LIB if REF_LIB.exists() else None
    deps = [REF_SRC, REFERENCE / "kernels/volsurfs/VolumeRenderingGPU.cuh", REFERENCE / "kernels/volsurfs/pcg32.h"]
    if REF_LIB.exists() and not force and all(REF_LIB.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return REF_LIB
    from torch.utils.cpp_extension import include_paths

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    REF_DIR.mkdir(exist_ok=True)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
           "-shared", "-w", f"-I{REFERENCE / 'kernels'}", f"-I{REFERENCE / 'include'}", *[f"-I{p}" for p in include_paths()],
           "-D_GLIBCXX_USE_CXX11_ABI=1", str(REF_SRC), "-o", str(REF_LIB), "-lcudart"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on the reference harness:\n{res.stdout[-4000:]}")
    return REF_LIB


# ---- oracle/_ref: the reference's permutohedral-encoding kernels behind a C harness -----------------------------------------
PERMUTO_SRC = HERE / "ref_permuto_harness.cu"
PERMUTO_LIB = REF_DIR / "libpermuto_ref.so"
PERMUTO_HDR = REFERENCE / "submodules/permutohedral_encoding/kernels/permutohedral_encoding/EncodingGPU.cuh"


def permuto_ref_available() -> bool:
    return PERMUTO_LIB.exists()


def build_ref_permuto(force: bool = False):
    """nvcc-compile oracle/ref_permuto_harness.cu, which #includes the reference's EncodingGPU.cuh from where it lies under
    /root/reference/submodules/permutohedral_encoding (nothing is copied), for sm_100a into oracle/_ref/libpermuto_ref.so.
    Default nvcc floating-point flags (FMA contraction on), as in the reference's own build (setup.py passes only -O3-style flags).
    Returns None when the reference tree is not mounted (GPU box): the prebuilt .so is used as is."""
    if not PERMUTO_HDR.exists():
        return PERMUTO_LIB if PERMUTO_LIB.exists() else None
    deps = [PERMUTO_SRC, PERMUTO_HDR]
    if PERMUTO_LIB.exists() and not force and all(PERMUTO_LIB.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return PERMUTO_LIB
    from torch.utils.cpp_extension import include_paths

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    REF_DIR.mkdir(exist_ok=True)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
           "-shared", "-w", f"-I{PERMUTO_HDR.parent.parent}", *[f"-I{p}" for p in include_paths()], "-D_GLIBCXX_USE_CXX11_ABI=1",
           str(PERMUTO_SRC), "-o", str(PERMUTO_LIB), "-lcudart"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on the permutohedral reference harness:\n{res.stdout[-4000:]}")
    return PERMUTO_LIB


# ---- oracle/_ref: the reference's ray-sampler / occupancy-grid kernels behind a C harness -------------------------------------
SAMPLER_SRC = HERE / "ref_sampler_harness.cu"
SAMPLER_LIB = REF_DIR / "libsampler_ref.so"
SAMPLER_HDRS = [REFERENCE / "kernels/volsurfs/RaySamplerGPU.cuh", REFERENCE / "kernels/volsurfs/OccupancyGridGPU.cuh",
                REFERENCE / "kernels/volsurfs/occ_grid_helpers.h", REFERENCE / "kernels/volsurfs/RaySamplesPackedGPU.cuh"]


def sampler_ref_available() -> bool:
    return SAMPLER_LIB.exists()


def build_ref_sampler(force: bool = False):
    """nvcc-compile oracle/ref_sampler_harness.cu, which #includes the reference's RaySamplerGPU.cuh and OccupancyGridGPU.cuh from where
    they lie (nothing is copied; Eigen::Vector3f is replaced by a 3-float stand-in declared in the harness), for sm_100a into
    oracle/_ref/libsampler_ref.so.  Default nvcc floating-point flags (FMA contraction on), as in the reference's CMake build.
    Returns None when the reference tree is not mounted (GPU box): the prebuilt .so is used as is."""
    if not SAMPLER_HDRS[0].exists():
        return SAMPLER_LIB if SAMPLER_LIB.exists() else None
    deps = [SAMPLER_SRC, *SAMPLER_HDRS]
    if SAMPLER_LIB.exists() and not force and all(SAMPLER_LIB.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return SAMPLER_LIB
    from torch.utils.cpp_extension import include_paths

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    REF_DIR.mkdir(exist_ok=True)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
           "-shared", "-w", f"-I{REFERENCE / 'kernels'}", f"-I{REFERENCE / 'include'}", *[f"-I{p}" for p in include_paths()],
           "-D_GLIBCXX_USE_CXX11_ABI=1", str(SAMPLER_SRC), "-o", str(SAMPLER_LIB), "-lcudart"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on the sampler reference harness:\n{res.stdout[-4000:]}")
    return SAMPLER_LIB


# ---- oracle/_ref: the reference's mesh ray tracer (raytracelib) behind a C harness ----------------------------------------------
RAYTRACE_SRC = HERE / "ref_raytrace_harness.cu"
RAYTRACE_LIB = REF_DIR / "libraytrace_ref.so"
RAYTRACELIB = REFERENCE / "submodules/raytracelib"
EIGEN_STANDIN = HERE / "eigen_standin"


def raytrace_ref_available() -> bool:
    return RAYTRACE_LIB.exists()


def build_ref_raytrace(force: bool = False):
    """nvcc-compile oracle/ref_raytrace_harness.cu, which #includes the reference's src/bvh.cu (and through it triangle.cuh,
    bounding_box.cuh, bvh.cuh, common.h, gpu_memory.h) from where they lie under /root/reference/submodules/raytracelib (nothing is
    copied), for sm_100a into oracle/_ref/libraytrace_ref.so.  Eigen (downloaded by raytracelib's setup.py, absent here) is replaced on
    the include path by oracle/eigen_standin.  raytracelib's own flags: -O3, default floating-point mode (FMA contraction on in device
    code); the host part goes through gcc without -mfma, i.e. without contraction.  No libtorch involved (seconds, not minutes).
    Returns None when the reference tree is not mounted (GPU box): the prebuilt .so is used as is."""
    if not (RAYTRACELIB / "src/bvh.cu").exists():
        return RAYTRACE_LIB if RAYTRACE_LIB.exists() else None
    deps = [RAYTRACE_SRC, EIGEN_STANDIN / "Eigen/Dense", RAYTRACELIB / "src/bvh.cu", *sorted((RAYTRACELIB / "include/raytracing").glob("*"))]
    if RAYTRACE_LIB.exists() and not force and all(RAYTRACE_LIB.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return RAYTRACE_LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    REF_DIR.mkdir(exist_ok=True)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--expt-extended-lambda", "--expt-relaxed-constexpr",
           "-Xcompiler", "-fPIC,-fopenmp,-O3", "-shared", "-w", f"-I{EIGEN_STANDIN}", f"-I{RAYTRACELIB / 'include'}", f"-I{RAYTRACELIB / 'src'}",
           str(RAYTRACE_SRC), "-o", str(RAYTRACE_LIB), "-lcudart", "-lgomp"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on the ray-tracer reference harness:\n{res.stdout[-4000:]}")
    return RAYTRACE_LIB


if __name__ == "__main__":
    print(build(force=True))
    print(build_permuto_c(force=True))
    print(build_ref())
    print(build_ref_permuto())
    print(build_ref_sampler())
    print(build_ref_raytrace())
