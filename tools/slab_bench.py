"""One rank's share of the 8-GPU (or N-GPU) 513^3 solve, alone on ONE GPU: a 513 x 513 x (512/N + 1) grid, so the
finest-level kernels see exactly the thin slab they see in the distributed run, without any exchange.  Separates
"thin slabs run the kernels below their roofline" from "the exchange costs time".

    python tools/slab_bench.py [N] [--march P,NT] [--force-mg 0|2] [--steps K]
Prints one JSON line: ms per solve, the finest-level kernel table (ms per launch, fraction of the HBM peak)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from p4pdes_b200 import lib as L  # noqa: E402
from p4pdes_b200.fish import Context, Multigrid, mg_options  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("nranks", type=int, nargs="?", default=8)
ap.add_argument("--march", default="")
ap.add_argument("--force-mg", type=int, default=0)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--no-graph", action="store_true")
ap.add_argument("--march-alt", type=int, default=1)
ap.add_argument("--full", action="store_true", help="the whole 513^3 grid instead of a slab")
ap.add_argument("--port-opts", type=int, default=0)
ap.add_argument("--trace", action="store_true", help="also one solve with every launch on every level bracketed")
a = ap.parse_args()
side = torch.cuda.Stream()
torch.cuda.set_stream(side)
if a.march:
    P, NT = (int(v) for v in a.march.split(","))
    L.tune("march_P", P)
    L.tune("march_NT", NT)
L.tune("force_mg", a.force_mg)
L.tune("march_alt", a.march_alt)
L.tune("port_opts", a.port_opts)
ctx = Context(0)
mz = 513 if a.full else 512 // a.nranks + 1
g = L.make_grid(3, (513, 513, mz), (1.0, 1.0, 1.0 if a.full else 1.0 / a.nranks), (1.0, 1.0, 1.0))
mg = Multigrid(ctx, g, mg_options(levels=0, use_graph=not a.no_graph))
n = mg.nlocal
b, x = ctx.empty(n), ctx.empty(n)
mg.fish_setup("manuexp", True, b=b)
for _ in range(3):
    res = mg.cg_solve(b, x, rtol=1e-10)
torch.cuda.synchronize()
f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
f0.record(ctx.stream)
for _ in range(a.steps):
    res = mg.cg_solve(b, x, rtol=1e-10)
f1.record(ctx.stream)
torch.cuda.synchronize()
ms_plain = f0.elapsed_time(f1) / a.steps
mg.profile(True)
mg.profile_reset()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(ctx.stream)
for _ in range(a.steps):
    res = mg.cg_solve(b, x, rtol=1e-10)
e1.record(ctx.stream)
torch.cuda.synchronize()
st = mg.profile_stats()
peak = 6553.9
trace = None
if a.trace:
    mg.profile(2)
    mg.profile_reset()
    mg.cg_solve(b, x, rtol=1e-10)
    tr = mg.profile_trace()
    mg.profile(False)
    trace = {str(l): {k: round(v["ms"] / v["launches"] * 1e3, 1) for k, v in d.items()} for l, d in tr.items()}
print(json.dumps({"port_opts": a.port_opts, "trace_us": trace, "slab_of": a.nranks, "grid": [513, 513, mz], "march": a.march or "default", "force_mg": a.force_mg,
                  "march_alt": a.march_alt, "levels": mg.nlevels, "ms_per_solve": round(ms_plain, 3),
                  "ms_per_solve_profiled": round(e0.elapsed_time(e1) / a.steps, 3), "its": res.its,
                  "kernels": {k: [round(v["ms"] / v["launches"], 4), round(v["bytes"] / v["ms"] / 1e6 / peak, 3)]
                              for k, v in st.items() if v["launches"]}}))
