"""Same-box A/B of library builds: python tools/ab_bench.py [steps]  with P4B_LIB=<variant .so> (see p4pdes_b200/lib.py).
Prints one JSON line with ms per 513^3 solve and the finest-level kernel table; tolerates older builds of the library
(symbols they lack are dropped from the ctypes table)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from p4pdes_b200 import lib as L  # noqa: E402

raw = ctypes.CDLL(L.LIB_PATH)
for k in list(L._SIGS):
    if not hasattr(raw, k):
        L._SIGS.pop(k)
from p4pdes_b200.fish import Context, Multigrid, mg_options  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
side = torch.cuda.Stream()
torch.cuda.set_stream(side)
ctx = Context(0)
g = L.refined_grid(3, 8)
mg = Multigrid(ctx, g, mg_options(levels=7))
n = mg.nlocal
b, x = ctx.empty(n), ctx.empty(n)
mg.fish_setup("manuexp", True, b=b)
for _ in range(3):
    mg.cg_solve(b, x, rtol=1e-10)
mg.profile(True)
mg.profile_reset()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(ctx.stream)
for _ in range(steps):
    res = mg.cg_solve(b, x, rtol=1e-10)
e1.record(ctx.stream)
torch.cuda.synchronize()
st = mg.profile_stats()
print(json.dumps({"lib": os.path.basename(L.LIB_PATH), "ms_per_step": round(e0.elapsed_time(e1) / steps, 2), "its": res.its,
                  "kernels": {k: round(v["ms"] / v["launches"], 3) for k, v in st.items()}}))
