"""Worker of the multi-GPU pattern.c tests: one rank per GPU under torchrun.  Runs the native time stepper on y-slabs
(p4b_pattern_solve with a communicator), gathers the final state on rank 0 and compares it with the SAME run on one GPU
(rank 0, a second context without a communicator).  Prints one JSON line."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from p4pdes_b200 import pattern as pp  # noqa: E402
from p4pdes_b200.fish import Context  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--argv", required=True)
    ap.add_argument("--time", action="store_true", help="also time a second run (after the first as warm-up)")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    ctx = Context(local, distributed=True)
    rep = pp.pattern_main(a.argv, ctx, native=True)
    seconds = None
    if a.time:
        torch.cuda.synchronize()
        dist.barrier()
        rep = pp.pattern_main(a.argv, ctx, native=True)
        torch.cuda.synchronize()
        t = torch.tensor([rep.seconds], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        seconds = float(t.item())
    parts = [torch.zeros_like(rep.Y) for _ in range(world)]
    dist.all_gather(parts, rep.Y)
    out = {"world": world, "m": rep.m, "steps": [[s[0], s[1], s[2]] for s in rep.steps], "seconds": seconds,
           "lines": rep.lines if rank == 0 else None}
    if rank == 0:
        Y = torch.cat(parts)
        one = pp.pattern_main(a.argv, Context(local), native=True)
        if a.time:
            one = pp.pattern_main(a.argv, Context(local), native=True)
        out["one_gpu_seconds"] = one.seconds
        # the all-reduced error norm rounds differently: times / steps agree to rounding, Newton counts exactly
        a, b = [[s[0], s[1], s[2]] for s in one.steps], out["steps"]
        out["same_steps"] = len(a) == len(b) and all(x[2] == y[2] and abs(x[0] - y[0]) <= 1e-9 * max(1.0, abs(x[0])) and
                                                     abs(x[1] - y[1]) <= 1e-9 * max(1.0, abs(x[1])) for x, y in zip(a, b))
        out["same_lines"] = one.lines == rep.lines
        out["one_gpu_lines"] = one.lines
        out["rel_diff"] = float(torch.linalg.vector_norm(Y - one.Y) / torch.linalg.vector_norm(one.Y))
        print("MGPU_RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
