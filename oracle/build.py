"""ORACLE build recipe (test infrastructure): compiles raytrace_oracle.c with gcc into oracle/_build/.
IEEE fp32, no FMA contraction (-ffp-contract=off), OpenMP over rays.  The reference's own tracer cannot be built here
(needs Eigen, which raytracelib's setup.py downloads; no network) — see DESIGN.md."""
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


if __name__ == "__main__":
    print(build(force=True))
