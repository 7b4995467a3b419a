"""clock64 event trace of the head forward kernel (CTA 0, its tile pairs 2 and 3): where a tile pair's cycles go.

    python scripts/trace_mlp_bwd.py --build     # the private -DVS_KERNEL_TRACE library is shared with the backward's trace
    gpurun -- 'timeout 120 python scripts/trace_mlp_fwd.py'

Event ids (volsurfs_b200/csrc/mlp.cu, VS_TRF), l = layer, s = slot: control 1ls before / 2ls after the ready wait, 3ls GEMM issued and
committed, 40s next features requested; epilogue thread 0: 1s/2s around the feature wait of build_a0, 3s operand announced, 1ls / 2ls
around the accumulator wait, 3ls hidden epilogue announced, 900 pair done.
"""
import ctypes
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from volsurfs_b200 import build as vb  # noqa: E402

TRACE_DIR = vb.PKG_DIR / "build" / "trace"
TRACE_LIB = TRACE_DIR / "libvolsurfs_b200_trace.so"

if "--build" in sys.argv:
    TRACE_DIR.mkdir(parents=True, exist_ok=True)
    objs = []
    procs = []
    for src in vb.sources():
        obj = TRACE_DIR / (src.stem + ".o")
        objs.append(str(obj))
        procs.append(subprocess.Popen([vb._nvcc(), *vb.NVCC_FLAGS, "-DVS_KERNEL_TRACE", "-I", str(ROOT / "include"), "-I", str(vb.CSRC),
                                       "-c", str(src), "-o", str(obj)]))
    assert all(p.wait() == 0 for p in procs)
    subprocess.check_call([vb._nvcc(), "-shared", "-o", str(TRACE_LIB), *objs, "-lcudart"])
    print(TRACE_LIB)
    sys.exit(0)

import torch  # noqa: E402

from volsurfs_b200 import _lib  # noqa: E402

_lib.LIB_PATH = TRACE_LIB
from volsurfs_b200.appearance import AppearanceHead  # noqa: E402

n = 892741
torch.manual_seed(0)
head = AppearanceHead(51, (128, 128, 64), 3, 3, False, "gelu", False).cuda()
pos = torch.rand(n, 51, device="cuda") * 2 - 1
dirs = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda"), dim=1)
nrm = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda"), dim=1)
g = torch.randn(n, 3, device="cuda") / n
stash = head.new_stash(n)
flat = torch.zeros(head.num_params(), device="cuda")
dpos = torch.zeros_like(pos)
lib = ctypes.CDLL(str(TRACE_LIB))
CAP = 256
buf = (ctypes.c_longlong * (2 * 2 * CAP))()
cnt = (ctypes.c_int * 2)()
for _ in range(3):
    out, _ = head.forward_train(pos, dirs, nrm, stash=stash)
    torch.cuda.synchronize()
lib.vs_debug_trace_fwd(buf, cnt)
events = []
for who, name in ((0, "control"), (1, "epilogue")):
    for q in range(cnt[who]):
        events.append((buf[who * 2 * CAP + 2 * q + 1], name, buf[who * 2 * CAP + 2 * q]))
events.sort()
t0 = events[0][0] if events else 0
last = {}
for t, name, e in events:
    col = {"control": 0, "epilogue": 1}[name]
    dt = t - last.get(name, t)
    last[name] = t
    print(f"{t - t0:8d}  " + " " * (22 * col) + f"{name[:3]} {e:3d} (+{dt})")
