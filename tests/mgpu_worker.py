"""Worker for the multi-GPU parity tests: one rank per GPU under torchrun.

Solves the fish 3-D problem on z-slabs (NCCL ghost planes + allreduce inside libp4b200) and, on rank 0,
compares the gathered solution and the residual history with the CPU oracle.  Prints one JSON line.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from p4pdes_b200 import lib as L  # noqa: E402
from p4pdes_b200.fish import Context, Multigrid, mg_options  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--refine", type=int, default=5)
    ap.add_argument("--levels", type=int, default=0)
    ap.add_argument("--rtol", type=float, default=1e-10)
    ap.add_argument("--cycle", default="v")
    ap.add_argument("--march-min-plane", type=int, default=16384)
    ap.add_argument("--rep-points", type=int, default=0)
    ap.add_argument("--comm-peer", type=int, default=1)
    ap.add_argument("--fused-halo", type=int, default=1)
    ap.add_argument("--repeat", type=int, default=1, help="solve this many times (CUDA-graph replay, exchange counters)")
    ap.add_argument("--compare-fused", action="store_true",
                    help="also solve with the other setting of fused_halo (same process, fresh hierarchy) and report "
                         "whether history and solution are bit-identical")
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--c-oracle", action="store_true",
                    help="check against oracle/fish_cpu.c (OpenMP; for the BASELINE sizes the NumPy oracle cannot run)")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L.tune("comm_peer", a.comm_peer)
    L.tune("fused_halo", a.fused_halo)
    ctx = Context(local, distributed=True)
    L.tune("march_min_plane", a.march_min_plane)
    if a.rep_points:
        L.tune("rep_points", a.rep_points)
    g = L.refined_grid(a.dim, a.refine)
    mg = Multigrid(ctx, g, mg_options(levels=a.levels, cycle=a.cycle))
    n = mg.nlocal
    b, x, u0 = ctx.empty(n), ctx.empty(n), ctx.empty(n)
    mg.fish_setup("manuexp", True, b=b, u0=u0)
    for _ in range(a.repeat):
        res = mg.cg_solve(b, x, rtol=a.rtol)
    if a.compare_fused:
        L.tune("fused_halo", 1 - a.fused_halo)
        mg2 = Multigrid(ctx, g, mg_options(levels=a.levels, cycle=a.cycle))
        x2 = ctx.empty(n)
        for _ in range(a.repeat):
            res2 = mg2.cg_solve(b, x2, rtol=a.rtol)
        same = torch.tensor([int(torch.equal(x, x2) and res.history == res2.history)], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        fused_equal = bool(same.item())
        mg2.close()
    ctx.axpy(-1.0, x, u0)          # u = u0 - y
    bnorm = ctx.norm2(b)           # allreduced inside the library
    # gather the slabs on rank 0
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([n], dtype=torch.int64, device="cuda"))
    sizes = [int(s.item()) for s in sizes]
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=torch.float64, device="cuda")
    buf[:n] = u0
    parts = [torch.zeros(pad, dtype=torch.float64, device="cuda") for _ in range(world)]
    dist.all_gather(parts, buf)
    out = {"world": world, "its": res.its, "reason": res.reason, "history": res.history, "bnorm": bnorm,
           "slab": [mg.zs, mg.zm], "nlevels": mg.nlevels}
    if a.compare_fused:
        out["fused_equal"] = fused_equal
    if rank == 0:
        u = torch.cat([p[:s] for p, s in zip(parts, sizes)]).cpu().numpy()
        assert u.size == g.n
        import hashlib
        out["sol_sha1"] = hashlib.sha1(u.tobytes()).hexdigest()
        if a.c_oracle:
            from oracle import fish_cpu as fc
            want = fc.solve(dim=a.dim, refine=a.refine, levels=a.levels, cycle=a.cycle, rtol=a.rtol, want_arrays=True)
            out["oracle_its"] = want["its"]
            out["sol_rel"] = float(np.linalg.norm(u - want["u"]) / np.linalg.norm(want["u"]))
            h, w = np.array(res.history), np.array(want["history"])
            out["hist_rel"] = float(np.max(np.abs(h - w) / (w + 4e-16 * w[0]))) if h.size == w.size else None
            out["bnorm_rel"] = abs(bnorm - want["fnorm0"]) / want["fnorm0"]
        elif not a.no_oracle:
            from oracle import fish_oracle as fo
            want = fo.fish(dim=a.dim, refine=a.refine, rtol=a.rtol,
                           mg=fo.MGOptions(levels=a.levels or None, cycle=a.cycle))
            out["oracle_its"] = want.its
            out["sol_rel"] = float(np.linalg.norm(u - want.u.ravel()) / np.linalg.norm(want.u))
            h = np.array(res.history)
            w = np.array(want.history)
            out["hist_rel"] = float(np.max(np.abs(h / w - 1.0))) if h.size == w.size else None
            out["bnorm_rel"] = abs(bnorm - want.fnorm0) / want.fnorm0
        print("MGPU_RESULT " + json.dumps(out), flush=True)
    mg.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
