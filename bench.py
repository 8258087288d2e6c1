#!/usr/bin/env python
"""bench.py -- fish 3-D Poisson CG + geometric-multigrid solve (BASELINE.json metric).

A "step" is one complete KSP solve of the fish.c 3-D manuexp problem (c/ch6/fish.c, options
`-fsh_dim 3 -da_refine 8 -pc_type mg -pc_mg_levels 7 -mg_levels_ksp_type chebyshev -mg_levels_ksp_max_it 2
-mg_levels_pc_type jacobi -ksp_rtol 1e-10`) on a 513^3 grid (135 M unknowns):
value = unknowns / solve seconds (MDOF/s), with b = F(u0) already resident in HBM.
`e2e` is the same solve through the C ABI's host-buffer entry point (p4b_cg_solve_host): b is
copied from pinned host memory, solved, and x copied back, all inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--refine R]

N > 1: launched by torchrun, one rank per GPU; the grid is split into z-slabs (strong scaling:
the total problem is fixed); ghost planes and dot products go through peer memory over NVLink, fused into the kernels
(--comm nccl: ncclSend/Recv/AllReduce instead, A/B).
--impl reference: the CPU restatement of the same algorithm (oracle/) on the host cores, on the same grid.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OPTIONS = ("-fsh_dim 3 -da_refine {refine} -pc_type mg -pc_mg_levels {levels} -mg_levels_ksp_type chebyshev "
           "-mg_levels_ksp_max_it 2 -mg_levels_pc_type jacobi -ksp_rtol 1e-10")
ALG_BYTES = {"apply_dot": "16N", "residual": "24N", "cheb_zero": "16N", "cheb_first": "24N", "cheb_next": "32N",
             "restrict": "8N+8Nc", "prolong_add": "16N+8Nc", "axpy2": "48N", "dot2": "16N", "aypx": "24N",
             "xp_update": "40N", "r_update": "24N"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled while the timed region runs (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.p = index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([s.strip() for s in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_cpus(index):
    """Pin this process to the CPUs that are local to GPU `index` (NVML's ideal affinity) BEFORE pinned host memory is
    allocated: under torchrun nothing binds the ranks, and eight ranks that pin their e2e buffers on one NUMA node share
    one memory controller and one inter-socket link for their PCIe copies (round 1: the 8-GPU e2e copies took 17 ms
    where 4.7 ms were possible).  Best effort: a restricted cpuset or a missing NVML leaves the affinity alone."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(index)
        if all(hasattr(pr, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):      # immune to CUDA_VISIBLE_DEVICES
            h = pynvml.nvmlDeviceGetHandleByPciBusId(("%08X:%02X:%02X.0" % (pr.pci_domain_id, pr.pci_bus_id,
                                                                              pr.pci_device_id)).encode())
        else:
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return {"cpus_before": before, "cpus_after": len(os.sched_getaffinity(0))}
    except Exception as exc:
        return {"error": repr(exc)[:120]}


def cpu_reference(refine, levels, threads=None, repeats=1):
    """Time the CPU restatement (oracle/) of the identical algorithm; returns (MDOF/s, info)."""
    from oracle import fish_cpu
    return fish_cpu.timed_solve(refine=refine, levels=levels, rtol=1e-10, threads=threads, repeats=repeats)


CPU_PORT_NOTE = ("OpenMP C restatement oracle/fish_cpu.c of the same algorithm and options (PETSc/MPI are not installable "
                 "here); matrix-free 7-point MatMult (16 B/node where PETSc's AIJ MatMult reads ~104 B/node), one loop per "
                 "PETSc operation (MatMult, PCApply, VecAXPY... unfused), so it moves ~2.7x the fused algorithmic bytes "
                 "the GPU roofline counts")


def cpu_stream_triad():
    """Host STREAM-triad GB/s (HARDWARE.md:14-21 of the reference: the CPU's own roofline), all cores."""
    try:
        from oracle import fish_cpu
        return round(fish_cpu.stream_triad_gbs(n=1 << 26, reps=5, threads=os.cpu_count()), 1)
    except Exception:
        return None


def workload_string(refine, n=None):
    m = 2 ** (refine + 1) + 1
    return ("fish.c 3-D Poisson manuexp %d^3 (%d unknowns), CG + V-cycle GMG, Chebyshev(2)/Jacobi, rtol 1e-10"
            % (m, n if n is not None else m ** 3))


def run_reference(args):
    """--impl reference: the CPU arm on the SAME workload as the GPU arm (same grid, same options).  PETSc cannot be
    built here, so the arm is the OpenMP C restatement (kind "port").  A 513^3 solve takes ~15 s on 16 cores, so the
    number of solves is clamped to a wall-clock budget and the line says how many ran."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    refine = args.cpu_refine if args.cpu_refine else args.refine
    levels = args.levels or max(2, refine - 1)
    cores = os.cpu_count()
    budget = args.cpu_budget_s
    t_start = time.perf_counter()
    warm = min(args.warmup, 1)            # a CPU solve needs one pass to fault its pages in, not three
    vals, ran_warm = [], 0
    for i in range(warm + args.steps):
        if i >= warm + 1 and vals:
            # stop when the next solve would overrun the budget
            if time.perf_counter() - t_start + vals[-1]["seconds"] * 1.3 > budget:
                break
        r = cpu_reference(refine, levels, threads=cores)
        if i >= warm:
            vals.append(r)
        else:
            ran_warm += 1
    secs = sum(v["seconds"] for v in vals) / len(vals)
    n = vals[0]["n"]
    mdofs = n / secs / 1e6
    m = 2 ** (refine + 1) + 1
    line = {
        "impl": "reference", "metric": "fish3d_cg_gmg_solve_mdof_per_s", "value": mdofs, "unit": "MDOF/s",
        "n_gpus": args.gpus, "steps": len(vals), "warmup": ran_warm, "steps_requested": args.steps,
        "warmup_requested": args.warmup, "ms_per_step": secs * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(refine, n), "options": OPTIONS.format(refine=refine, levels=levels),
                   "levels": vals[0]["nlevels"]},
        "cpu_baseline": {"value": mdofs, "unit": "MDOF/s", "cores": vals[0]["threads"], "kind": "port",
                         "stream_triad_gbs": cpu_stream_triad(),
                         "sample": "%d timed solve(s) of the full %d^3 workload (%d unknowns, %d KSP its, %.1f s each; "
                                   "solve count clamped to a %d s budget); %s"
                                   % (len(vals), m, n, vals[0]["its"], secs, budget, CPU_PORT_NOTE)},
        "e2e": {"value": mdofs, "unit": "MDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "ksp_its": vals[0]["its"], "errinf": vals[0]["errinf"],
    }
    print(json.dumps(line))
    return 0


def _time_kernel(fn, reps=30, warm=5):
    import torch
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run_secondary(args):
    """The other BASELINE.json configurations (one GPU): one JSON line each, same keys as the headline line.  A step =
    one complete run of the driver's solve (c1: KSP solve; c4: grid-sequenced Newton-Krylov-MG solve; c5: implicit
    time steps).  `roofline` is the dominant kernel of that driver timed ALONE in this process with CUDA events (these
    solves are launch-latency bound on 4 M unknowns: the kernel table is context, the solve time is the number)."""
    if args.impl == "reference":
        print(json.dumps({"impl": "reference", "config": {"workload": args.config},
                          "unavailable": "the CPU restatement that is timed as the reference arm exists for the fish "
                                         "configurations (c1, c2, c3); minimal.c / pattern.c need PETSc itself"}))
        return 0
    import torch
    from p4pdes_b200 import lib as L
    from p4pdes_b200 import minimal as pm, pattern as pp
    from p4pdes_b200.fish import Context, Multigrid, mg_options
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and args.config != "c5":
        if rank == 0:
            print(json.dumps({"config": {"workload": args.config}, "n_gpus": world,
                              "unavailable": "this configuration runs on one GPU (only c3 / c2 z-slabs and c5 y-slabs shard)"}))
        return 0
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    ctx = Context(local_rank, distributed=world > 1)
    lib = ctx.lib
    peak, peak_src = peaks()
    sampler = ClockSampler(0)
    roof, extra, cpu = None, {}, None
    if args.config == "c1":
        refine = 6
        g = L.refined_grid(2, refine)
        mg = Multigrid(ctx, g, mg_options(levels=refine + 1))
        n = mg.nlocal
        b, x, u0, ue = ctx.empty(n), ctx.empty(n), ctx.empty(n), ctx.empty(n)
        mg.fish_setup("manuexp", True, b=b, u0=u0, uexact=ue)
        for _ in range(args.warmup):
            res = mg.cg_solve(b, x, rtol=1e-10)
        torch.cuda.synchronize()
        sampler.start()
        l0 = lib.p4b_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ctx.stream)
        for _ in range(args.steps):
            res = mg.cg_solve(b, x, rtol=1e-10)
        e1.record(ctx.stream)
        torch.cuda.synchronize()
        launches = lib.p4b_launch_count() - l0
        ms = e0.elapsed_time(e1) / args.steps
        bh, xh = torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory()
        bh.copy_(b)
        mg.cg_solve_host(bh, xh, rtol=1e-10)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            mg.cg_solve_host(bh, xh, rtol=1e-10)
        torch.cuda.synchronize()
        ms_e2e = (time.perf_counter() - t0) / args.steps * 1e3
        ctx.axpy(-1.0, x, u0)
        ctx.axpy(-1.0, ue, u0)
        ndof = g.n
        workload = "fish.c 2-D Poisson manuexp %d^2 (-fsh_dim 2 -da_refine 6), CG + V-cycle GMG, Chebyshev(2)/Jacobi, rtol 1e-10" % g.mx
        options = "-fsh_dim 2 -da_refine 6 -pc_type mg -mg_levels_ksp_type chebyshev -mg_levels_ksp_max_it 2 -mg_levels_pc_type jacobi -ksp_rtol 1e-10"
        metric = "fish2d_cg_gmg_solve_mdof_per_s"
        extra = {"ksp_its": res.its, "errinf": ctx.norminf(u0), "note": "16 641 unknowns: every kernel is launch-latency bound"}
        h2d = d2h = 8 * n
        try:
            from oracle import fish_cpu
            r = fish_cpu.solve(dim=2, refine=6, rtol=1e-10, threads=os.cpu_count())
            cpu = {"value": r["n"] / r["seconds"] / 1e6, "unit": "MDOF/s", "cores": r["threads"], "kind": "port",
                   "solve_s": r["seconds"], "ksp_its": r["its"], "sample": "the same solve; " + CPU_PORT_NOTE}
        except Exception as exc:
            cpu = {"value": None, "unit": "MDOF/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (exc,)}
    elif args.config == "f3":
        from p4pdes_b200.bratu import bratu_main
        refine = args.refine if args.refine != 8 else 12          # --refine 12 (20481^2) unless given
        argv = ("-da_grid_x 6 -da_grid_y 6 -lb_exact -snes_rtol 1.0e-10 -snes_converged_reason -lb_showcounts -snes_type fas "
                "-snes_fas_type full -fas_levels_snes_type ngs -fas_levels_snes_ngs_sweeps 2 -fas_levels_snes_max_it 1 "
                "-fas_coarse_snes_type ngs -fas_coarse_snes_ngs_sweeps 2 -fas_coarse_snes_max_it 4 -da_refine %d" % refine)   # bratu2D.c:15
        for _ in range(max(1, min(args.warmup, 2))):
            rep = bratu_main(argv, ctx, keep_solution=True)
        uh = torch.empty(rep.mx * rep.my, dtype=torch.float64).pin_memory()      # pinned once, outside the timed region
        uh.copy_(rep.u)
        del rep.u
        torch.cuda.synchronize()
        sampler.start()
        l0 = lib.p4b_launch_count()
        dev_ms, t0 = 0.0, time.perf_counter()
        for _ in range(args.steps):
            rep = bratu_main(argv, ctx, keep_solution=True)
            dev_ms += rep.solve_ms
            uh.copy_(rep.u)                          # the solution reaches the host
            del rep.u
        torch.cuda.synchronize()
        ms_e2e = (time.perf_counter() - t0) / args.steps * 1e3
        ms = dev_ms / args.steps
        launches = lib.p4b_launch_count() - l0
        ndof = rep.mx * rep.my
        workload = ("bratu2D.c Liouville-Bratu %d^2 (%d unknowns), FAS full cycles + nonlinear Gauss-Seidel (red-black), "
                    "rtol 1e-10 (c/ch7/solns/bratu2D.c:15)" % (rep.mx, ndof))
        options = argv
        metric = "bratu2d_fas_ngs_solve_mdof_per_s"
        extra = {"fas_its": rep.its, "errinf": rep.errinf, "residual_calls": rep.residual_calls, "ngs_calls": rep.ngs_calls,
                 "reference_published": {"value": 12.7, "unit": "MDOF/s", "what": "4.0e8 unknowns in 31.50 s, mpiexec -n 20 on a "
                                         "40-core workstation, lexicographic NGS (c/ch7/solns/bratu2D.c:9-19, BASELINE.md 1 "
                                         "'adjacent'): other hardware, reported beside, not a same-box baseline"},
                 "note": "ms_per_step = CUDA events around the solve; e2e adds the (pooled) allocation of the hierarchy and the D2H "
                         "copy of the solution into pinned host memory (wall clock); the problem has no input vector (u0 = 0, "
                         "analytic boundary data)"}
        h2d, d2h = 0, 8 * ndof
        m = rep.mx
        uu, ff = ctx.zeros(m * m), ctx.empty(m * m)
        kms = _time_kernel(lambda: L.check(lib.p4b_bratu_ngs(ctx.h, m, m, 1.0, 1, 1, None, uu.data_ptr())), reps=10, warm=2)
        by = 2 * 16.0 * m * m                        # two half sweeps, each reads and writes u (the other colour's lines ride along)
        roof = {"bound": "hbm", "kernel": "bratu_ngs_kernel (one red-black sweep = two half-sweep launches)",
                "achieved": by / kms / 1e6, "peak": peak, "unit": "GB/s", "frac": by / kms / 1e6 / peak,
                "alg_bytes_per_launch": by, "ms_per_launch": kms, "traffic": None,
                "how": "one sweep (2 launches) timed alone, 10 repetitions, CUDA events; 3.4 GB per vector: exceeds L2"}
    elif args.config == "c4":
        argv = "-da_grid_x 33 -da_grid_y 33 -snes_grid_sequence 6 -snes_fd_color -pc_type mg"      # c/ch8/cluster.sh:70
        for _ in range(max(1, min(args.warmup, 2))):
            rep = pm.minimal_main(argv, ctx, native=True)
        torch.cuda.synchronize()
        sampler.start()
        l0 = lib.p4b_launch_count()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            rep = pm.minimal_main(argv, ctx, native=True, keep_solution=True)
            uh = rep.u.cpu()                       # the step's result reaches the host
        torch.cuda.synchronize()
        ms = ms_e2e = (time.perf_counter() - t0) / args.steps * 1e3
        launches = lib.p4b_launch_count() - l0
        ndof = 2049 * 2049
        workload = "minimal.c catenoid, grid-sequenced Newton-GMRES-MG to 2049^2 (c/ch8/cluster.sh:70), FD-coloured Jacobians"
        options = argv + " -mg_levels_pc_type jacobi"
        metric = "minimal2d_newton_krylov_mg_solve_mdof_per_s"
        extra = {"newton_its": [s.its for s in rep.stages], "errinf": rep.errinf,
                 "note": "whole run incl. the six coarser grids of the sequence; wall clock (host loop + kernels); the "
                         "initial iterate is generated on the device, the solution (33.6 MB) is copied to the host"}
        h2d, d2h = 0, 8 * ndof
        from p4pdes_b200 import callbacks as cb
        m = 2049
        gtab = cb.minimal_g(ctx, m, m, "catenoid", 1.0, 1.1)
        u = gtab.clone() * 0.9
        F0 = ctx.empty(m * m)
        cb.minimal_form_function(ctx, m, m, u, gtab, -0.5, out=F0)
        vals = ctx.empty(9 * m * m)
        L.check(lib.p4b_minimal_jacobian_fd(ctx.h, m, m, -0.5, u.data_ptr(), gtab.data_ptr(), F0.data_ptr(), vals.data_ptr()))
        xx, bb, out = ctx.empty(m * m), ctx.empty(m * m), ctx.empty(m * m)
        kms = _time_kernel(lambda: L.check(lib.p4b_stencil9_lin(ctx.h, m, m, vals.data_ptr(), xx.data_ptr(), bb.data_ptr(),
                                                                 None, 0.0, 1.0, 0.5, 1, out.data_ptr())))
        by = 96.0 * m * m
        roof = {"bound": "hbm", "kernel": "stencil9_lin (Chebyshev/Jacobi step on the assembled 9-point Jacobian)",
                "achieved": by / kms / 1e6, "peak": peak, "unit": "GB/s", "frac": by / kms / 1e6 / peak,
                "alg_bytes_per_launch": by, "ms_per_launch": kms, "traffic": None,
                "how": "kernel timed alone (30 launches, CUDA events) on the 2049^2 operator; 37 MB per operand: partly L2 resident"}
    else:
        argv = ("-da_grid_x 8 -da_grid_y 8 -da_refine 8 -ts_type beuler -ts_dt 5 -ts_max_time 50 -pc_type mg "
                "-p4b_mg_rscale 0.25")
        for _ in range(max(1, min(args.warmup, 2))):
            rep = pp.pattern_main(argv, ctx, native=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        if rank == 0:
            sampler.start()
        l0 = lib.p4b_launch_count()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            rep = pp.pattern_main(argv, ctx, native=True)
            yh = rep.Y.cpu()                       # each rank's rows reach its host
        torch.cuda.synchronize()
        nsteps = len(rep.steps)
        ms = ms_e2e = (time.perf_counter() - t0) / args.steps * 1e3
        if world > 1:                              # max over ranks (every rank sits in the same all-reduces)
            tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = ms_e2e = float(tt.item())
        launches = lib.p4b_launch_count() - l0
        m = 2048
        ndof = 2 * m * m * nsteps                  # unknowns advanced: 8.4 M per implicit step x steps per run
        workload = ("pattern.c Gray-Scott 2048^2 x 2 dof, %d backward-Euler steps (dt 5), Newton-GMRES-MG per step, "
                    "matrix-free stage Jacobian" % nsteps)
        options = argv + " -mg_levels_pc_type jacobi"
        metric = "pattern2d_implicit_steps_mdof_per_s"
        extra = {"ts_steps": nsteps, "ms_per_ts_step": ms / nsteps, "newton_per_step": [s[2] for s in rep.steps],
                 "parallelism": "y-slabs x%d (ring ghost rows + all-reduce over NCCL; levels below 16 rows replicated)" % world,
                 "note": "-p4b_mg_rscale 0.25 (averaging restriction) keeps GMRES counts mesh independent "
                         "(DESIGN.md section 8); the final state (67 MB) is copied to the host"}
        h2d, d2h = 0, 16 * m * m
        if rank != 0:
            dist.destroy_process_group()
            return 0
        from p4pdes_b200 import callbacks as cb
        Y = cb.pattern_initial_state(ctx, m, m)
        X, bb, out = Y.clone(), Y.clone(), ctx.empty(2 * m * m)
        kms = _time_kernel(lambda: L.check(lib.p4b_pattern_jac_lin(ctx.h, m, m, 2.5, 8.0e-5, 4.0e-5, 0.024, 0.06, 0.2,
                                                                    Y.data_ptr(), X.data_ptr(), bb.data_ptr(), None, 0.0, 1.0,
                                                                    0.5, 1, out.data_ptr())))
        by = 64.0 * m * m
        roof = {"bound": "hbm", "kernel": "pattern_jac_kernel (matrix-free stage-Jacobian smoother step)",
                "achieved": by / kms / 1e6, "peak": peak, "unit": "GB/s", "frac": by / kms / 1e6 / peak,
                "alg_bytes_per_launch": by, "ms_per_launch": kms, "traffic": None,
                "how": "kernel timed alone (30 launches, CUDA events) at 2048^2 x 2; 67 MB per operand: partly L2 resident"}
    clocks = sampler.stop()
    line = {"metric": metric, "value": ndof / (ms * 1e-3) / 1e6, "unit": "MDOF/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "solve_s": ms * 1e-3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "options": options, "baseline_config": args.config,
                       "l2": "vectors of this configuration fit the 126 MB L2; nothing is flushed between solves (the solve "
                             "itself streams every level many times)"},
            "e2e": {"value": ndof / (ms_e2e * 1e-3) / 1e6, "unit": "MDOF/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
    line.update(extra)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c1", "c2", "c3", "c4", "c5", "f3"],
                    help="BASELINE.json configuration: c3 (default) = fish 3-D 513^3, the headline; c2 = fish 3-D 257^3; "
                         "c1 = fish 2-D -da_refine 6; c4 = minimal.c 2049^2 (c/ch8/cluster.sh:70); c5 = pattern.c 2048^2 x 2; "
                         "f3 = c/ch7/solns/bratu2D.c FAS+NGS at 20481^2 (SURVEY 8 f3, the reference's one published throughput)")
    ap.add_argument("--refine", type=int, default=8, help="-da_refine (8 = 513^3, 7 = 257^3)")
    ap.add_argument("--levels", type=int, default=0, help="-pc_mg_levels (default refine-1: coarse grid 9^3)")
    ap.add_argument("--cpu-refine", type=int, default=0, help="grid of the CPU arm / cpu_baseline (default: --refine, "
                                                               "i.e. the same workload)")
    ap.add_argument("--cpu-budget-s", type=float, default=200.0, help="wall-clock budget of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-steps", type=int, default=1,
                    help="how many of the timed solves carry per-kernel CUDA-event brackets (the roofline source)")
    ap.add_argument("--port-stats", action="store_true", help="collect in-kernel wait / fence times (comm_stats)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fuse", action="store_true")
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl"], help="multi-GPU transport (A/B)")
    ap.add_argument("--no-fused-halo", action="store_true", help="one push kernel per ghost exchange (A/B)")
    ap.add_argument("--rep-points", type=int, default=0, help="replicate levels with at most this many nodes (A/B)")
    ap.add_argument("--force-mg", type=int, default=0, help="run the multi-GPU kernel variants on one GPU (A/B)")
    ap.add_argument("--port-opts", type=int, default=0, help="HaloPort experiments (p4b_tune port_opts)")
    ap.add_argument("--no-graph", action="store_true", help="launch the coarse levels kernel by kernel (A/B)")
    ap.add_argument("--trace", default="", help="after the timed steps, run one more solve with every launch on every "
                                                "level bracketed by CUDA events and write the table to this JSON file")
    args = ap.parse_args()
    if args.config == "c2":
        args.refine = 7
    if args.config in ("c1", "c4", "c5", "f3"):
        return run_secondary(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from p4pdes_b200 import lib as L
    from p4pdes_b200.fish import Context, Multigrid, mg_options

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    affinity = bind_to_gpu_cpus(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # everything runs on one explicit (capturable) stream: torch ops, the library's kernels, its CUDA graph
    side = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(side)
    L.tune("comm_peer", 1 if args.comm == "peer" else 0)
    L.tune("fused_halo", 0 if args.no_fused_halo else 1)
    L.tune("port_opts", args.port_opts | (4 if args.port_stats else 0))
    L.tune("force_mg", args.force_mg)
    if args.rep_points:
        L.tune("rep_points", args.rep_points)
    ctx = Context(local_rank, distributed=world > 1)
    lib = ctx.lib

    refine = args.refine
    levels = args.levels or max(2, refine - 1)
    g = L.refined_grid(3, refine)
    ndof = g.n
    mg = Multigrid(ctx, g, mg_options(levels=levels, fuse=not args.no_fuse, use_graph=not args.no_graph))
    nloc = mg.nlocal
    b = ctx.empty(nloc)
    x = ctx.empty(nloc)
    uex = ctx.empty(nloc)
    u0 = ctx.empty(nloc)
    mg.fish_setup("manuexp", True, b=b, u0=u0, uexact=uex)
    ctx.sync()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident solves -------------------------------------------------------------------
    res = None
    for _ in range(args.warmup):
        res = mg.cg_solve(b, x, rtol=1e-10)
    mg.profile(True)
    mg.profile_reset()
    # a timed region that saw a hardware / thermal slowdown is re-measured once (the recipe's rule); a power cap is
    # kept and reported
    BAD = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    remeasured = 0
    while True:
        mg.profile_reset()
        sampler = ClockSampler(local_rank)
        barrier()
        if rank == 0:
            sampler.start()
        launches0 = lib.p4b_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(ctx.stream)
        for i in range(args.steps):
            # kernel brackets (CUDA events around every finest-level launch) in the first --profile-steps timed
            # solves only: the event records cost GPU front-end time between kernels, which the thin slabs of an
            # 8-GPU run notice
            mg.profile(i < args.profile_steps)
            res = mg.cg_solve(b, x, rtol=1e-10)
        ev1.record(ctx.stream)
        barrier()
        launches = lib.p4b_launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
        again = 1.0 if (rank == 0 and remeasured == 0 and BAD & set(clocks["reasons"])) else 0.0
        if max_over_ranks(again) == 0.0:
            break
        remeasured += 1
    if rank == 0:
        clocks["remeasured"] = remeasured
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_step = ms_total / args.steps
    stats = mg.profile_stats()
    mg.profile(False)
    cs = (C.c_ulonglong * 5)()
    lib.p4b_comm_stats(ctx.h, C.byref(cs), 1)
    comm_stats = {"waits": cs[0], "wait_ms": cs[1] / 1e6, "wait_max_us": cs[2] / 1e3, "fences": cs[3],
                  "fence_ms": cs[4] / 1e6, "note": "sums over boundary CTAs (rank 0), all warm-up + timed steps; "
                                                   "collected only with --port-stats"}
    exchange = {k: stats.pop(k) for k in L.KERNEL_CLASSES[L.N_ROOFLINE_CLASSES:] if k in stats}
    if args.trace:
        mg.profile(2)
        mg.profile_reset()
        barrier()
        tres = mg.cg_solve(b, x, rtol=1e-10)
        barrier()
        tr = mg.profile_trace()
        mg.profile(False)
        with open("%s.rank%d" % (args.trace, rank) if world > 1 else args.trace, "w") as fh:
            json.dump({"n_gpus": world, "rank": rank, "solve_ms": tres.solve_ms, "its": tres.its,
                       "levels": {str(l): v for l, v in tr.items()}}, fh, indent=1)

    # correctness of what was timed: error norms of u = u0 - y against the exact solution (fish.c:248-280)
    ctx.axpy(-1.0, x, u0)
    ctx.axpy(-1.0, uex, u0)
    errinf = ctx.norminf(u0)
    err2h = ctx.norm2(u0) / ((g.mx - 1) * (g.my - 1) * (g.mz - 1)) ** 0.5

    # ---- end to end: pinned host buffers through p4b_cg_solve_host -----------------------------------
    e2e = None
    if not args.no_e2e:
        bh = torch.empty(nloc, dtype=torch.float64).pin_memory()
        xh = torch.empty(nloc, dtype=torch.float64).pin_memory()
        bh.copy_(b)
        torch.cuda.synchronize()
        mg.cg_solve_host(bh, xh, rtol=1e-10)
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ctx.stream)
        nst = max(1, min(args.steps, 3))
        for _ in range(nst):
            mg.cg_solve_host(bh, xh, rtol=1e-10)
        e1.record(ctx.stream)
        barrier()
        wall = (time.perf_counter() - t0) / nst
        ms_e2e = max_over_ranks(max(e0.elapsed_time(e1) / nst, 0.0))
        e2e = {"value": ndof / (ms_e2e * 1e-3) / 1e6, "unit": "MDOF/s", "h2d_bytes_per_step": 8 * nloc * world,
               "d2h_bytes_per_step": 8 * nloc * world, "ms_per_step": ms_e2e, "wall_ms_per_step": wall * 1e3,
               "api": "p4b_cg_solve_host (pinned host b -> device, solve, x -> pinned host)", "cpu_affinity": affinity}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant finest-level kernel ------------------------------------------------
    peak, peak_src = peaks()
    roof = None
    table = {}
    if stats:
        total_ms = sum(s["ms"] for s in stats.values())
        for name, s in sorted(stats.items(), key=lambda kv: -kv[1]["ms"]):
            table[name] = {"launches": s["launches"], "ms_per_launch": s["ms"] / s["launches"],
                           "GBs": s["bytes"] / s["ms"] / 1e6, "frac": s["bytes"] / s["ms"] / 1e6 / peak,
                           "share_of_fine_level_ms": s["ms"] / total_ms, "alg_bytes": ALG_BYTES[name]}
        top = max(stats, key=lambda k: stats[k]["ms"])
        s = stats[top]
        ach = s["bytes"] / s["ms"] / 1e6
        roof = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)", "traffic": None,
                "alg_bytes_per_launch": s["bytes"] / s["launches"], "launches": s["launches"],
                "ms_per_launch": s["ms"] / s["launches"],
                "profiled_steps": min(args.profile_steps, args.steps),
                "fine_level_kernel_ms_share_of_step": total_ms / (ms_step * min(args.profile_steps, args.steps))}
        # DRAM bytes per launch from the committed `ncu --set full` capture of these kernels on ONE GPU at this grid
        # (profiles/traffic.json; not re-measured by this run).  Slabs of a multi-GPU run launch on 1/N of the grid, so
        # the single-GPU capture does not describe them: null there.
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp) and world == 1 and refine == 8:
            tj = json.load(open(tp))
            roof["traffic"] = tj.get(top)
            roof["traffic_source"] = "committed ncu --set full capture (%s), dram__bytes_read.sum + dram__bytes_write.sum "\
                                     "per launch" % tj.get("_source", "profiles/traffic.json")

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cref = args.cpu_refine if args.cpu_refine else refine
            r = cpu_reference(cref, args.levels or max(2, cref - 1), threads=os.cpu_count())
            m = 2 ** (cref + 1) + 1
            cpu = {"value": r["n"] / r["seconds"] / 1e6, "unit": "MDOF/s", "cores": r["threads"], "kind": "port",
                   "solve_s": r["seconds"], "ksp_its": r["its"], "stream_triad_gbs": cpu_stream_triad(),
                   "sample": "one solve of the full %d^3 workload (%d unknowns); %s" % (m, r["n"], CPU_PORT_NOTE)}
        except Exception as exc:  # the baseline is a reported number, never a reason to lose the GPU line
            cpu = {"value": None, "unit": "MDOF/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (exc,)}

    m = g.mx
    line = {
        "metric": "fish3d_cg_gmg_solve_mdof_per_s", "value": ndof / (ms_step * 1e-3) / 1e6, "unit": "MDOF/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "solve_s": ms_step * 1e-3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(refine, ndof),
                   "options": OPTIONS.format(refine=refine, levels=levels), "levels": mg.nlevels,
                   "parallelism": "z-slabs x%d" % world, "transport": (args.comm if world > 1 else None), "l2": "inputs (%.2f GB per vector) exceed the 126 MB L2"
                   % (8 * ndof / 1e9), "fused": not args.no_fuse, "cuda_graph_coarse_levels": not args.no_graph,
                   "fused_halo": (not args.no_fused_halo) if world > 1 and args.comm == "peer" else None},
        "ksp_its": res.its, "ksp_reason": L.REASONS.get(res.reason), "rnorm0": res.rnorm0, "rnorm": res.rnorm,
        "errinf": errinf, "err2h": err2h,
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof, "kernels": table,
        "exchange_ms_per_step": {k: v["ms"] / max(1, min(args.profile_steps, args.steps)) for k, v in exchange.items()},
        "comm_stats": comm_stats,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
