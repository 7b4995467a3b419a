"""In-tree nvcc build of the CUDA kernels + C ABI into ``volsurfs_b200/libvolsurfs_b200.so``.

sm_100a only (``-gencode arch=compute_100a,code=sm_100a``); nvcc cross-compiles without a GPU.
The shared object is git-ignored but travels to the GPU box with the snapshot.

    python -m volsurfs_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libvolsurfs_b200.so"
STAMP = PKG_DIR / ".libvolsurfs_b200.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (looked at $NVCC, /usr/local/cuda/bin/nvcc, PATH)")


def sources():
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list((PKG_DIR.parent / "include").glob("*.h"))):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh() -> bool:
    return LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == _digest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every ``csrc/*.cu`` and link them into one shared object. Returns its path."""
    if not force and is_fresh():
        return LIB_PATH
    nvcc = _nvcc()
    obj_dir = PKG_DIR / "build"
    obj_dir.mkdir(exist_ok=True)
    include = PKG_DIR.parent / "include"
    procs = []
    objs = []
    for src in sources():
        obj = obj_dir / (src.stem + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(include), "-I", str(CSRC), "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"[volsurfs_b200.build] nvcc failed on {src.name}:\n{out}\n")
        elif verbose and out:
            print(out)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    link = [nvcc, "-shared", "-o", str(LIB_PATH), *map(str, objs), "-lcudart"]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}")
    STAMP.write_text(_digest())
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
