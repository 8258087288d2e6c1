"""Host-side mirror of c/ch7/solns/bratu2D.c on top of the C ABI: `bratu_main(argv, ctx)` takes the
reference's command line for its FAS + NGS runs (c/ch7/solns/bratu2D.c:9-19, c/ch7/solns/makefile:12) and prints the
reference's lines (bratu2D.c:137-160).  The solve is ONE C-ABI call, p4b_bratu_solve (csrc/bratu.cu): device kernels for
the residual and the (red-black) nonlinear Gauss-Seidel sweeps, the FAS cycle's logic in C++ inside the library.
Newton-Krylov runs of bratu2D.c (-snes_type newtonls -pc_type mg ...) are not provided: its registered Jacobian is
Poisson's ("ONLY APPROXIMATE", bratu2D.c:126-129)."""
from __future__ import annotations

import ctypes as C
import shlex
import time
from dataclasses import dataclass, field

from . import lib as L


@dataclass
class BratuOptions:
    lam: float = 1.0
    exact: bool = False
    showcounts: bool = False
    grid_x: int = 3
    grid_y: int = 3
    refine: int = 0
    snes_type: str = "newtonls"
    fas_type: str = "multiplicative"
    fas_levels: int = 0
    levels_snes_type: str = ""
    coarse_snes_type: str = ""
    smooth_sweeps: int = 1
    smooth_its: int = 1
    coarse_sweeps: int = 1
    coarse_its: int = 50
    snes_rtol: float = 1.0e-8
    snes_max_it: int = 10000
    monitor: bool = False
    converged_reason: bool = False


@dataclass
class BratuReport:
    mx: int
    my: int
    its: int
    reason: int
    fnorm: list
    errinf: float | None
    residual_calls: int
    ngs_calls: int
    seconds: float
    solve_ms: float
    u: object = None
    lines: list = field(default_factory=list)


def parse_options(argv) -> BratuOptions:
    if isinstance(argv, str):
        argv = shlex.split(argv)
    o = BratuOptions()
    flags = {"-lb_exact": "exact", "-lb_showcounts": "showcounts", "-snes_monitor_short": "monitor", "-snes_monitor": "monitor",
             "-snes_converged_reason": "converged_reason"}
    valued = {"-lb_lambda": ("lam", float), "-da_grid_x": ("grid_x", int), "-da_grid_y": ("grid_y", int),
              "-da_refine": ("refine", int), "-snes_type": ("snes_type", str), "-snes_fas_type": ("fas_type", str),
              "-snes_fas_levels": ("fas_levels", int), "-fas_levels_snes_type": ("levels_snes_type", str),
              "-fas_coarse_snes_type": ("coarse_snes_type", str), "-fas_levels_snes_ngs_sweeps": ("smooth_sweeps", int),
              "-fas_levels_snes_max_it": ("smooth_its", int), "-fas_coarse_snes_ngs_sweeps": ("coarse_sweeps", int),
              "-fas_coarse_snes_max_it": ("coarse_its", int), "-snes_rtol": ("snes_rtol", float),
              "-snes_max_it": ("snes_max_it", int)}
    i = 0
    while i < len(argv):
        a = argv[i]
        if a in flags:
            setattr(o, flags[a], True)
            i += 1
        elif a in valued:
            name, typ = valued[a]
            setattr(o, name, typ(argv[i + 1]))
            i += 2
        else:
            raise L.P4BError("unknown or unsupported option %s" % a)
    if o.snes_type != "fas":
        raise L.P4BError("bratu2D on the device: -snes_type fas (the reference's headline runs, bratu2D.c:9-19); the "
                         "Newton-Krylov variants use Poisson's Jacobian as an approximation and are not provided")
    if o.levels_snes_type != "ngs" or o.coarse_snes_type != "ngs":
        raise L.P4BError("-fas_levels_snes_type ngs -fas_coarse_snes_type ngs are what the device path provides "
                         "(a Newton-Krylov coarse solve with CG+ICC/Cholesky is sequential PETSc machinery)")
    if o.fas_type not in ("full", "multiplicative"):
        raise L.P4BError("-snes_fas_type %s: full and multiplicative are provided" % o.fas_type)
    if o.exact and o.lam != 1.0:
        raise L.P4BError("Liouville exact solution only implemented for lambda = 1.0")          # bratu2D.c:99-101
    return o


def bratu_main(argv, ctx, echo=False, keep_solution=False) -> BratuReport:
    opt = parse_options(argv)
    lines = []

    def out(s):
        lines.append(s)
        if echo:
            print(s)

    o = L.BratuOpts()
    L.check(ctx.lib.p4b_bratu_default_opts(C.byref(o)))
    o.lam, o.exact = opt.lam, int(opt.exact)
    o.grid_x, o.grid_y, o.refine, o.levels = opt.grid_x, opt.grid_y, opt.refine, opt.fas_levels
    o.snes_rtol, o.snes_max_it = opt.snes_rtol, opt.snes_max_it
    o.smooth_sweeps, o.smooth_its = opt.smooth_sweeps, opt.smooth_its
    o.coarse_sweeps, o.coarse_its = opt.coarse_sweeps, opt.coarse_its
    o.full_cycle = int(opt.fas_type == "full")
    o.monitor, o.converged_reason = int(opt.monitor), int(opt.converged_reason)
    mx, my = opt.grid_x, opt.grid_y
    for _ in range(opt.refine):
        mx, my = 2 * mx - 1, 2 * my - 1
    u = ctx.empty(mx * my) if keep_solution else None
    res = L.BratuResult()
    cb = L.LINE_FN(lambda line, _ctx: out(line.decode()))
    t0 = time.perf_counter()
    L.check(ctx.lib.p4b_bratu_solve(ctx.h, C.byref(o), cb, None, u.data_ptr() if u is not None else None,
                                    u.numel() if u is not None else 0, C.byref(res)))
    seconds = time.perf_counter() - t0
    if opt.showcounts:          # bratu2D.c:137-142 (flops are PetscLogFlops bookkeeping of the host callbacks: not counted here)
        out("flops = (not counted on the device),  residual calls = %d,  NGS calls = %d" % (res.residual_calls, res.ngs_calls))
    if opt.exact:
        out("done on %d x %d grid:   error |u-uexact|_inf = %.3e" % (res.mx, res.my, res.errinf))      # :154-156
    else:
        out("done on %d x %d grid ..." % (res.mx, res.my))                                             # :158
    return BratuReport(res.mx, res.my, res.its, res.reason, [res.fnorm[i] for i in range(res.nnorm)],
                       res.errinf if opt.exact else None, res.residual_calls, res.ngs_calls, seconds, res.solve_ms, u, lines)
