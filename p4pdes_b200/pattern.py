"""Host-side mirror of c/ch5/pattern.c for its IMPLICIT runs, on top of the C ABI (include/p4b200.h).

`pattern_main(argv)` takes pattern.c's command line (`-ptn_*`, c/ch5/pattern.c:54-78; `-da_grid_x/_y -da_refine -ts_type
beuler -ts_dt -ts_max_time -pc_type mg -snes_rtol ...`: c/ch5/makefile:52-53), runs
pattern.c:main -- periodic 2-dof DMDA, banner, InitialState, TSSolve -- and prints the reference's lines.  Inside TSSolve:

  backward Euler           [PETSc] TSTHETA(theta = 1), fixed step, final step matched to -ts_max_time
  stage residual           p4b_pattern_ifunction (pattern.c:242-267) with Ydot = (Y - Y_n)/dt, minus p4b_pattern_rhsfunction
                           (:185-199)
  stage Jacobian           shift*I - C L9 - G'(Y), matrix-free: p4b_pattern_jac_apply (FormIJacobianLocal :274-318 and
                           FormRHSJacobianLocal :202-236; `-ptn_no_rhsjacobian` drops G')
  Newton / GMRES           the device loops of p4pdes_b200/minimal.py (SNESNEWTONLS + bt, KSPGMRES(30))
  -pc_type mg              V cycle on the rediscretised level operators at the injected iterate: p4b_pattern_jac_lin
                           (Chebyshev + Jacobi), p4b_pattern_restrict / _prolong_add / _inject (periodic Q1), dense inverse
                           of the base-grid operator x vector

  -ts_type cn              [PETSc] TSCN: the theta = 1/2 endpoint form of the same stage equation (c/ch5/makefile:56)
  -ts_type arkimex (pattern.c's default, :115): [PETSc] TSARKIMEX3 = ARK3(2)4L[2]SA with TSAdaptBasic and
                           MATCHSTEP.  The diffusion (IFunction) is implicit: each of the three implicit stages solves the
                           LINEAR system (shift*I - C L9) Y_i = shift*Z with the same GMRES + multigrid machinery (G' never
                           enters: the reaction is explicit); the embedded 2nd-order solution gives the error estimate,
                           p4b_vec_wrms2 its weighted norm ([PETSc] TSErrorWeightedNorm2).

`-ptn_noisy_init` uses the restated rander48 stream (unpinned, include/p4b200.h).  No CPU path: `ops` must be a device Context (tests/ substitutes a NumPy
stand-in to exercise this file's control flow without a GPU).
"""
from __future__ import annotations

import math
import shlex
import time
from dataclasses import dataclass, field

import numpy as np

from .minimal import KSPResult, SNESResult, gmres, linesearch_bt


@dataclass
class PatternOptions:
    L: float = 2.5
    Du: float = 8.0e-5
    Dv: float = 4.0e-5
    phi: float = 0.024
    kappa: float = 0.06
    no_rhsjacobian: bool = False
    call_back_report: bool = False
    noisy_init: float = 0.0         # -ptn_noisy_init (pattern.c:72,159-165): uniform noise on [0, level] from the
                                    # VecSetRandom stream ([PETSc] rander48, restated and unpinned: include/p4b200.h)
    grid_x: int = 3
    grid_y: int = 3
    refine: int = 0
    ts_type: str = "arkimex"
    ts_dt: float = 5.0
    ts_max_time: float = 200.0
    ts_max_steps: int = 5000
    ts_rtol: float = 1.0e-4
    ts_atol: float = 1.0e-4
    ts_monitor: bool = False
    ts_monitor_file: str = ""           # -ts_monitor binary:FILE            (c/ch5/MOVIES.md:44: PETSc binary Real records)
    ts_monitor_solution_file: str = ""  # -ts_monitor_solution binary:FILE   (PETSc binary Vec records; native host only)
    pc_type: str = "mg"
    smooth_its: int = 2
    mg_rscale: float = 1.0          # -p4b_mg_rscale: factor on the restricted residual.  1 = [PETSc] R = P^T.  pattern.c's
                                    # equations are scaled pointwise (no cell-volume factor, unlike fish.c), so P^T over-
                                    # weights the coarse correction 4x; 0.25 (averaging restriction) restores
                                    # mesh-independent Krylov counts on fine grids (2048^2: 362 -> single digits)
    snes_rtol: float = 1.0e-8
    snes_stol: float = 1.0e-8
    snes_atol: float = 1.0e-50
    snes_max_it: int = 50
    ksp_rtol: float = 1.0e-5
    ksp_max_it: int = 10000
    gmres_restart: int = 30
    snes_converged_reason: bool = False
    ksp_converged_reason: bool = False
    log_view: bool = False


def parse_options(argv) -> PatternOptions:
    if isinstance(argv, str):
        argv = shlex.split(argv)
    o = PatternOptions()
    flags = {"-ptn_no_rhsjacobian": "no_rhsjacobian", "-ptn_call_back_report": "call_back_report",
             "-ts_monitor": "ts_monitor", "-snes_converged_reason": "snes_converged_reason",
             "-ksp_converged_reason": "ksp_converged_reason", "-log_view": "log_view"}
    valued = {"-ptn_L": ("L", float), "-ptn_Du": ("Du", float), "-ptn_Dv": ("Dv", float), "-ptn_phi": ("phi", float),
              "-ptn_kappa": ("kappa", float), "-ptn_noisy_init": ("noisy_init", float), "-da_grid_x": ("grid_x", int), "-da_grid_y": ("grid_y", int),
              "-da_refine": ("refine", int), "-ts_type": ("ts_type", str), "-ts_dt": ("ts_dt", float),
              "-ts_max_time": ("ts_max_time", float), "-ts_max_steps": ("ts_max_steps", int), "-pc_type": ("pc_type", str),
              "-ts_rtol": ("ts_rtol", float), "-ts_atol": ("ts_atol", float),
              "-mg_levels_ksp_max_it": ("smooth_its", int), "-p4b_mg_rscale": ("mg_rscale", float),
              "-snes_rtol": ("snes_rtol", float),
              "-snes_max_it": ("snes_max_it", int), "-ksp_rtol": ("ksp_rtol", float), "-ksp_max_it": ("ksp_max_it", int),
              "-ksp_gmres_restart": ("gmres_restart", int)}
    accepted = {"-mg_levels_ksp_type": ("chebyshev",), "-mg_levels_pc_type": ("jacobi",), "-ksp_type": ("gmres",),
                "-ts_adapt_type": ("basic",), "-ts_arkimex_type": ("3",)}
    i = 0
    while i < len(argv):
        a = argv[i]
        if a in ("-ts_monitor", "-ts_monitor_solution") and i + 1 < len(argv) and argv[i + 1].startswith("binary:"):
            setattr(o, "ts_monitor_file" if a == "-ts_monitor" else "ts_monitor_solution_file", argv[i + 1][len("binary:"):])
            i += 2
        elif a == "-ts_monitor_solution":
            raise ValueError("-ts_monitor_solution: the binary viewer is provided (binary:FILE)")
        elif a in flags:
            setattr(o, flags[a], True)
            i += 1
        elif a in valued:
            name, typ = valued[a]
            setattr(o, name, typ(argv[i + 1]))
            i += 2
        elif a in accepted:
            if argv[i + 1] not in accepted[a]:
                raise ValueError("%s %s: the device path provides %s only" % (a, argv[i + 1], "|".join(accepted[a])))
            i += 2
        elif a == "-ptn_no_ijacobian":
            raise ValueError("%s is not provided by the device path (finite-difference IJacobian)" % a)
        else:
            raise ValueError("unknown or unsupported option %s" % a)
    if o.ts_type not in ("arkimex", "beuler", "cn", "bdf"):
        raise ValueError("-ts_type %s: the device path provides arkimex (pattern.c's default), beuler, cn and bdf" % o.ts_type)
    if o.pc_type not in ("mg", "none"):
        raise ValueError("-pc_type %s: the device path provides mg and none (ilu/sor are sequential)" % o.pc_type)
    return o


class Level:
    def __init__(self, ops, m, opt: PatternOptions):
        self.ops, self.m, self.n = ops, m, 2 * m * m
        self.Y = ops.empty(self.n)            # the iterate this level's operator is linearised at
        self.x, self.b, self.t = ops.empty(self.n), ops.empty(self.n), ops.empty(self.n)
        self.scale, self.omega = 0.0, []


class StageOperator:
    """J = shift*I - C L9 - G'(Y) on every level of the periodic hierarchy (levels[0] finest), with the V cycle."""

    def __init__(self, ops, levels, opt: PatternOptions, imex=False):
        self.ops, self.levels, self.opt = ops, levels, opt
        self.shift = 0.0
        self.Ainv = None
        self.par = (opt.L, opt.Du, opt.Dv, opt.phi, opt.kappa)
        self.no_rhs = opt.no_rhsjacobian or imex          # IMEX: the reaction is explicit, G' never enters the stage matrix

    def _Y(self, L):
        return None if self.no_rhs else L.Y

    def mult(self, x, y):
        L = self.levels[0]
        self.ops.pattern_jac_apply(L.m, *self.par, self.shift, self._Y(L), x, y)

    def setup(self, shift):
        """Linearise at levels[0].Y: injected iterates, Chebyshev targets, dense base-grid inverse."""
        ops, opt = self.ops, self.opt
        self.shift = shift
        for l, L in enumerate(self.levels):
            if l > 0 and not self.no_rhs:
                ops.pattern_inject(L.m, L.m, self.levels[l - 1].Y, L.Y)
            if l < len(self.levels) - 1:
                lam = ops.pattern_jac_gershgorin(L.m, *self.par, shift, self._Y(L), L.t)
                emin, emax = 0.1 * lam, 1.1 * lam
                L.scale = 2.0 / (emax + emin)                      # [PETSc] KSPSolve_Chebyshev
                alpha = 1.0 - L.scale * emin
                mu, omegaprod = 1.0 / alpha, 2.0 / alpha
                cm1, ck = 1.0, mu
                L.omega = []
                for _ in range(1, opt.smooth_its):
                    cp1 = 2.0 * mu * ck - cm1
                    L.omega.append(omegaprod * ck / cp1)
                    cm1, ck = ck, cp1
        if opt.pc_type != "mg":                                    # unpreconditioned: no base-grid solve to set up
            return
        C = self.levels[-1]
        if C.n > 2048:
            raise ValueError("base grid of the hierarchy has %d unknowns: use a coarser -da_grid_x/_y" % C.n)
        Yc = None if self.no_rhs else ops.to_host(C.Y).reshape(C.m, C.m, 2)
        self.Ainv = ops.from_host(np.linalg.inv(dense_stage_jacobian(C.m, shift, Yc, *self.par)).ravel())

    def _smooth(self, L, zero_guess):
        ops, its = self.ops, self.opt.smooth_its
        Y = self._Y(L)
        if its <= 0:
            if zero_guess:
                ops.set(0.0, L.x)
            return
        pm1, pk = L.x, L.t
        if zero_guess:
            ops.set(0.0, pm1)
        ops.pattern_jac_lin(L.m, *self.par, self.shift, Y, pm1, L.b, None, 0.0, 1.0, L.scale, True, pk)
        for i in range(1, its):
            w = L.omega[i - 1]
            ops.pattern_jac_lin(L.m, *self.par, self.shift, Y, pk, L.b, pm1, 1.0 - w, w, w * L.scale, True, pm1)
            pm1, pk = pk, pm1
        if pk is not L.x:
            L.x, L.t = L.t, L.x

    def _cycle(self, l, zero_guess):
        ops = self.ops
        L = self.levels[l]
        if l == len(self.levels) - 1:
            ops.dense_matvec(L.n, self.Ainv, L.b, L.x)
            return
        C = self.levels[l + 1]
        self._smooth(L, zero_guess)
        ops.pattern_jac_lin(L.m, *self.par, self.shift, self._Y(L), L.x, L.b, None, 0.0, 0.0, 1.0, False, L.t)   # b - J x
        ops.pattern_restrict(C.m, C.m, L.t, C.b)
        if self.opt.mg_rscale != 1.0:
            ops.axpby(self.opt.mg_rscale, C.b, 0.0, None, C.b)
        self._cycle(l + 1, True)
        ops.pattern_prolong_add(C.m, C.m, C.x, L.x)
        self._smooth(L, False)

    def precond(self, r, z):
        if self.opt.pc_type == "none":
            self.ops.copy(r, z)
            return
        L = self.levels[0]
        self.ops.copy(r, L.b)
        self._cycle(0, True)
        self.ops.copy(L.x, z)


def dense_stage_jacobian(m, shift, Y, L, Du, Dv, phi, kappa):
    """Host copy of the (small) base-grid operator as a dense array ([PETSc] PCLU on the coarsest level)."""
    h = L / m
    C = (Du / (6.0 * h * h), Dv / (6.0 * h * h))
    n = m * m
    A = np.zeros((2 * n, 2 * n))
    for j in range(m):
        for i in range(m):
            k = j * m + i
            for c in (0, 1):
                r = 2 * k + c
                A[r, r] += shift + 20.0 * C[c]
                for dj, di, w in ((0, -1, 4.0), (0, 1, 4.0), (-1, 0, 4.0), (1, 0, 4.0), (-1, -1, 1.0), (1, -1, 1.0),
                                  (-1, 1, 1.0), (1, 1, 1.0)):
                    A[r, 2 * (((j + dj) % m) * m + (i + di) % m) + c] += -w * C[c]
            if Y is not None:
                u, v = Y[j, i, 0], Y[j, i, 1]
                A[2 * k, 2 * k] -= -v * v - phi
                A[2 * k, 2 * k + 1] -= -2.0 * u * v
                A[2 * k + 1, 2 * k] -= v * v
                A[2 * k + 1, 2 * k + 1] -= 2.0 * u * v - (phi + kappa)
    return A


def fmt_g(v):
    """PETSc's %g: an integral value prints with a trailing '.' ("5.", "200.")."""
    s = "%g" % v
    return s + "." if s.lstrip("-").isdigit() else s


@dataclass
class PatternReport:
    m: int
    steps: list
    Y: object
    seconds: float
    lines: list


def _pattern_native(opt: PatternOptions, ctx, out) -> PatternReport:
    """The same run through ONE C-ABI call, p4b_pattern_solve (csrc/nk_device.cu + ts_solver.hpp)."""
    import ctypes as C

    from . import lib as L
    o = L.PatternOpts()
    L.check(ctx.lib.p4b_pattern_default_opts(C.byref(o)))
    o.L, o.Du, o.Dv, o.phi, o.kappa = opt.L, opt.Du, opt.Dv, opt.phi, opt.kappa
    o.no_rhsjacobian, o.call_back_report = int(opt.no_rhsjacobian), int(opt.call_back_report)
    o.grid_x, o.grid_y, o.refine = opt.grid_x, opt.grid_y, opt.refine
    o.ts_type = {"arkimex": 0, "beuler": 1, "cn": 2, "bdf": 3}[opt.ts_type]
    o.ts_dt, o.ts_max_time, o.ts_max_steps = opt.ts_dt, opt.ts_max_time, opt.ts_max_steps
    o.ts_rtol, o.ts_atol, o.ts_monitor = opt.ts_rtol, opt.ts_atol, int(opt.ts_monitor)
    o.pc_type, o.smooth_its, o.mg_rscale = {"none": 0, "mg": 1}[opt.pc_type], opt.smooth_its, opt.mg_rscale
    o.snes_rtol, o.snes_stol, o.snes_atol, o.snes_max_it = opt.snes_rtol, opt.snes_stol, opt.snes_atol, opt.snes_max_it
    o.ksp_rtol, o.ksp_max_it, o.gmres_restart = opt.ksp_rtol, opt.ksp_max_it, opt.gmres_restart
    o.snes_converged_reason, o.ksp_converged_reason = int(opt.snes_converged_reason), int(opt.ksp_converged_reason)
    m = opt.grid_x * 2 ** opt.refine
    # a context with a communicator (Context(distributed=True) under torchrun) runs on y-slabs: this rank's rows only
    nranks, rank = getattr(ctx, "nranks", 1), getattr(ctx, "rank", 0)
    rows, ys = m, 0
    if nranks > 1:
        if m % nranks or (m // nranks) % 2:
            raise ValueError("pattern on %d ranks: %d rows must split into an even number of rows per rank" % (nranks, m))
        rows, ys = m // nranks, rank * (m // nranks)
    Y = ctx.empty(2 * m * rows)
    res = L.PatternResult()
    say = out if rank == 0 else (lambda s: None)
    cb = L.LINE_FN(lambda line, _ctx: say(line.decode()))
    step_cb, files = None, []
    if opt.ts_monitor_file or opt.ts_monitor_solution_file:
        # [PETSc] binary viewers on the TS monitors: times and states as PETSc binary records (petscbin.py)
        import numpy as np
        from . import petscbin
        suffix = ".rank%d" % rank if nranks > 1 else ""          # on slabs every rank writes its rows
        ft = open(opt.ts_monitor_file, "wb") if (opt.ts_monitor_file and rank == 0) else None
        fu = open(opt.ts_monitor_solution_file + suffix, "wb") if opt.ts_monitor_solution_file else None
        files = [f for f in (ft, fu) if f]

        def _step(_user, _k, t, Yp, n):
            if ft:
                petscbin.write_real(ft, t)
            if fu:
                petscbin.write_vec(fu, np.ctypeslib.as_array(Yp, shape=(n,)))
            return 0

        step_cb = L.TS_STEP_FN(_step)
        L.check(ctx.lib.p4b_set_ts_step_monitor(step_cb, None))
    t0 = time.perf_counter()
    if opt.noisy_init > 0.0:
        # the caller's initial state (what the shim's TSSolve does with pattern.c's Vec); on slabs: rows [ys, ys + rows)
        say("running on %d x %d grid with square cells of side h = %.6f ..." % (m, m, opt.L / m))
        if nranks > 1:
            full = ctx.empty(2 * m * m)
            initial_state(ctx, opt, m, full)
            Y.copy_(full[2 * m * ys:2 * m * (ys + rows)])
            del full
        else:
            initial_state(ctx, opt, m, Y)
        L.check(ctx.lib.p4b_pattern_solve_from(ctx.h, C.byref(o), Y.data_ptr(), cb, None, Y.data_ptr(), Y.numel(),
                                               C.byref(res)))
    else:
        L.check(ctx.lib.p4b_pattern_solve(ctx.h, C.byref(o), cb, None, Y.data_ptr(), Y.numel(), C.byref(res)))
    seconds = time.perf_counter() - t0
    if step_cb is not None:
        L.check(ctx.lib.p4b_set_ts_step_monitor(C.cast(None, L.TS_STEP_FN), None))
        for f in files:
            f.close()
    steps = [(res.step_t[k], res.step_dt[k], res.step_newton[k]) for k in range(min(res.nsteps, 512))]
    if opt.log_view:
        out("TSSolve %.6f s (%d steps, %d rejected, %d GMRES iterations)" % (seconds, res.nsteps, res.rejected,
                                                                           res.ksp_its_total))
    rep = PatternReport(m=res.m, steps=steps, Y=Y, seconds=seconds, lines=None)
    rep.rejected = res.rejected
    return rep


def initial_state(ops, opt, m, Y):
    """InitialState (pattern.c:146-179), with -ptn_noisy_init the noise of VecSetRandom underneath the patch."""
    if opt.noisy_init > 0.0:
        ops.pattern_initial_state_noisy(m, m, opt.L, opt.noisy_init, Y)
    else:
        ops.pattern_initial_state(m, m, opt.L, Y)


def pattern_main(argv, ops, echo=False, native=False) -> PatternReport:
    """native=True: the whole run is one call of p4b_pattern_solve (host logic in C++ inside the library)."""
    opt = parse_options(argv)
    lines = []

    def out(s):
        lines.append(s)
        if echo:
            print(s)

    if native:
        rep = _pattern_native(opt, ops, out)
        rep.lines = lines
        return rep
    if opt.ts_monitor_file or opt.ts_monitor_solution_file:
        raise ValueError("-ts_monitor[_solution] binary:FILE: provided by the native host (pattern_main(..., native=True))")

    mx, my = opt.grid_x * 2 ** opt.refine, opt.grid_y * 2 ** opt.refine     # periodic: -da_refine doubles
    if mx != my:
        raise ValueError("pattern.c requires mx == my")                                                  # pattern.c:89
    m = mx
    out("running on %d x %d grid with square cells of side h = %.6f ..." % (m, m, opt.L / m))          # :94-96
    sizes = [m]
    if opt.pc_type == "mg":
        while sizes[-1] > opt.grid_x and sizes[-1] % 2 == 0:
            sizes.append(sizes[-1] // 2)
    levels = [Level(ops, s, opt) for s in sizes]
    if opt.ts_type == "arkimex":
        return _arkimex(ops, opt, levels, m, out, lines)
    if opt.ts_type == "bdf":
        return _bdf(ops, opt, levels, m, out, lines)
    A = StageOperator(ops, levels, opt)
    L0 = levels[0]
    n = L0.n
    Y, Y0, R, Ydot, G = L0.Y, ops.empty(n), ops.empty(n), ops.empty(n), ops.empty(n)
    theta = 0.5 if opt.ts_type == "cn" else 1.0
    affine = ops.empty(n) if theta != 1.0 else None
    y, Jy, w, gnew = ops.empty(n), ops.empty(n), ops.empty(n), ops.empty(n)
    work = [ops.empty(n) for _ in range(opt.gmres_restart + 1)]
    initial_state(ops, opt, m, Y)                                                                      # :146-179
    t0 = time.perf_counter()
    t, k, steps = 0.0, 0, []
    dt_last = opt.ts_dt
    while t < opt.ts_max_time - 1e-14 * max(1.0, abs(opt.ts_max_time)) and k < opt.ts_max_steps:
        dt = min(opt.ts_dt, opt.ts_max_time - t)                   # TS_EXACTFINALTIME_MATCHSTEP (:118)
        dt_last = dt
        if opt.ts_monitor:
            out("%d TS dt %s time %s" % (k, fmt_g(dt), fmt_g(t)))
        # [PETSc] TSTHETA: theta = 1 backward Euler; theta = 1/2 in endpoint form = Crank-Nicolson (TSCN):
        #   F(W, (W - Y0)/(theta dt)) - G(W) + (1 - theta)/theta [F(Y0, 0) - G(Y0)] = 0
        shift = 1.0 / (theta * dt)
        ops.copy(Y, Y0)
        if theta != 1.0:
            ops.set(0.0, Ydot)
            ops.pattern_ifunction(m, m, opt.L, opt.Du, opt.Dv, Y0, Ydot, affine)
            ops.pattern_rhsfunction(m, m, opt.phi, opt.kappa, Y0, G)
            ops.axpy(-1.0, G, affine)

        def F(W, f):
            ops.axpby(shift, W, -shift, Y0, Ydot)
            ops.pattern_ifunction(m, m, opt.L, opt.Du, opt.Dv, W, Ydot, f)
            ops.pattern_rhsfunction(m, m, opt.phi, opt.kappa, W, G)
            ops.axpy(-1.0, G, f)
            if theta != 1.0:
                ops.axpy((1.0 - theta) / theta, affine, f)

        F(Y, R)
        fnorm = ops.norm2(R)
        res = SNESResult(fnorms=[fnorm])
        ttol = opt.snes_rtol * fnorm
        if fnorm < opt.snes_atol:
            res.reason = "CONVERGED_FNORM_ABS"
        while not res.reason:
            if res.its >= opt.snes_max_it:
                res.reason = "DIVERGED_MAX_IT"
                break
            A.setup(shift)
            kr = gmres(ops, A.mult, R, y, A.precond, opt.ksp_rtol, restart=opt.gmres_restart, max_it=opt.ksp_max_it,
                       work=work)
            res.ksp_its.append(kr.its)
            if opt.ksp_converged_reason:
                out("      Linear solve %s due to %s iterations %d" % ("converged" if kr.reason.startswith("CONV")
                                                                       else "did not converge", kr.reason, kr.its))
            A.mult(y, Jy)
            gnorm, lam = linesearch_bt(ops, F, Y, R, fnorm, y, Jy, w, gnew)
            res.lambdas.append(lam)
            ops.axpby(1.0, w, -1.0, Y, y)
            snorm, xnorm = ops.norm2(y), ops.norm2(w)
            ops.copy(w, Y)
            ops.copy(gnew, R)
            fnorm = gnorm
            res.its += 1
            res.fnorms.append(fnorm)
            if not math.isfinite(fnorm):
                res.reason = "DIVERGED_FNORM_NAN"
            elif fnorm < opt.snes_atol:
                res.reason = "CONVERGED_FNORM_ABS"
            elif fnorm <= ttol:
                res.reason = "CONVERGED_FNORM_RELATIVE"
            elif snorm < opt.snes_stol * xnorm:
                res.reason = "CONVERGED_SNORM_RELATIVE"
        if opt.snes_converged_reason:
            out("    Nonlinear solve %s due to %s iterations %d" % ("converged" if res.reason.startswith("CONV")
                                                                    else "did not converge", res.reason, res.its))
        if not res.reason.startswith("CONV"):
            raise RuntimeError("TSSolve: nonlinear solve failed at step %d (%s)" % (k, res.reason))
        t += dt
        k += 1
        steps.append((t, dt, res))
    if opt.ts_monitor:
        out("%d TS dt %s time %s" % (k, fmt_g(dt_last), fmt_g(t)))
    ops.sync()
    seconds = time.perf_counter() - t0
    if opt.call_back_report:                                                                           # :127-135
        out("CALL-BACK REPORT")
        out("  solver type: %s" % opt.ts_type)
        out("  IFunction:   1  | IJacobian:   1")
        out("  RHSFunction: 1  | RHSJacobian: %d" % (0 if opt.no_rhsjacobian else 1))
    if opt.log_view:
        out("TSSolve %.6f s" % seconds)
    return PatternReport(m=m, steps=steps, Y=Y, seconds=seconds, lines=lines)


# ---------------------------------------------------------------------------------------------------------
# -ts_type bdf: [PETSc] TSBDF, order 2 (the statement of csrc/ts_solver.hpp and oracle/pattern_solver_oracle.py:pattern_bdf):
# backward-Euler half-step restart, stage derivative from the Lagrange basis over the history, extrapolated initial guess,
# LTE from the next-higher difference into TSAdaptBasic, MATCHSTEP.  c/ch5/output/pattern.test5 pins the restart step.
# ---------------------------------------------------------------------------------------------------------
def _lagrange_vals(t, T):
    v = [1.0] * len(T)
    for k in range(len(T)):
        for j in range(len(T)):
            if j != k:
                v[k] *= (t - T[j]) / (T[k] - T[j])
    return v


def _lagrange_ders(t, T):
    n = len(T)
    d = [0.0] * n
    for k in range(n):
        for j in range(n):
            if j == k:
                continue
            p = 1.0 / (T[k] - T[j])
            for l in range(n):
                if l not in (k, j):
                    p *= (t - T[l]) / (T[k] - T[l])
            d[k] += p
    return d


def _bdf(ops, opt: PatternOptions, levels, m, out, lines, order=2) -> PatternReport:
    A = StageOperator(ops, levels, opt)
    n = levels[0].n
    Y = levels[0].Y                                  # the stage operator linearises about this vector
    Yacc, R, Ydot, G, V0, lte = (ops.empty(n) for _ in range(6))
    y, Jy, w, gnew = (ops.empty(n) for _ in range(4))
    gwork = [ops.empty(n) for _ in range(opt.gmres_restart + 1)]
    tm, wk = [0.0] * 8, [ops.empty(n) for _ in range(8)]
    st = dict(k=0, n=0)
    initial_state(ops, opt, m, Yacc)
    t0 = time.perf_counter()
    tmax = opt.ts_max_time
    t, k, h, steps, rejected = 0.0, 0, min(opt.ts_dt, tmax), [], 0

    def advance(tt, X):                              # TSBDF_Advance
        tail = wk[7]
        for i in range(7, 1, -1):
            tm[i], wk[i] = tm[i - 1], wk[i - 1]
        st["n"] = min(st["n"] + 1, 7)
        tm[1], wk[1] = tt, tail
        ops.copy(X, tail)

    def stage(X):                                    # TSBDF_PreSolve + SNESSolve_NEWTONLS on X (in place)
        nn = max(st["k"], 1) + 1
        a = _lagrange_ders(tm[0], tm[:nn])
        ops.set(0.0, V0)
        for i in range(1, nn):
            ops.axpy(a[i], wk[i], V0)
        shift = a[0]

        def F(W, f):
            ops.axpby(shift, W, 1.0, V0, Ydot)
            ops.pattern_ifunction(m, m, opt.L, opt.Du, opt.Dv, W, Ydot, f)
            ops.pattern_rhsfunction(m, m, opt.phi, opt.kappa, W, G)
            ops.axpy(-1.0, G, f)

        F(X, R)
        fnorm = ops.norm2(R)
        res = SNESResult(fnorms=[fnorm])
        ttol = opt.snes_rtol * fnorm
        if fnorm < opt.snes_atol:
            res.reason = "CONVERGED_FNORM_ABS"
        while not res.reason:
            if res.its >= opt.snes_max_it:
                res.reason = "DIVERGED_MAX_IT"
                break
            ops.copy(X, Y)
            A.setup(shift)
            kr = gmres(ops, A.mult, R, y, A.precond, opt.ksp_rtol, restart=opt.gmres_restart, max_it=opt.ksp_max_it,
                       work=gwork)
            res.ksp_its.append(kr.its)
            if opt.ksp_converged_reason:
                out("      Linear solve %s due to %s iterations %d" % ("converged" if kr.reason.startswith("CONV")
                                                                       else "did not converge", kr.reason, kr.its))
            A.mult(y, Jy)
            gnorm, lam = linesearch_bt(ops, F, X, R, fnorm, y, Jy, w, gnew)
            ops.axpby(1.0, w, -1.0, X, y)
            snorm, xnorm = ops.norm2(y), ops.norm2(w)
            ops.copy(w, X)
            ops.copy(gnew, R)
            fnorm = gnorm
            res.its += 1
            res.fnorms.append(fnorm)
            if not math.isfinite(fnorm):
                res.reason = "DIVERGED_FNORM_NAN"
            elif fnorm < opt.snes_atol:
                res.reason = "CONVERGED_FNORM_ABS"
            elif fnorm <= ttol:
                res.reason = "CONVERGED_FNORM_RELATIVE"
            elif snorm < opt.snes_stol * xnorm:
                res.reason = "CONVERGED_SNORM_RELATIVE"
        if opt.snes_converged_reason:
            out("    Nonlinear solve %s due to %s iterations %d" % ("converged" if res.reason.startswith("CONV")
                                                                    else "did not converge", res.reason, res.its))
        if not res.reason.startswith("CONV"):
            raise RuntimeError("TSSolve: nonlinear solve failed at step %d (%s)" % (k, res.reason))
        return res

    restart = True
    hnext = h
    while t < tmax - 1e-12 * max(1.0, abs(tmax)) and k < opt.ts_max_steps:
        if opt.ts_monitor:
            out("%d TS dt %s time %s" % (k, fmt_g(h), fmt_g(t)))
        if not restart:
            st["k"] = min(st["k"] + 1, order)
            advance(t, Yacc)
        accept, newton = True, []
        while True:
            if restart:                              # TSBDF_Restart
                st["k"], st["n"] = 1, 0
                advance(t, Yacc)
                tm[0] = t + h / 2.0
                ops.copy(wk[1], wk[0])
                newton.append(stage(wk[0]))
                st["k"] = min(2, order)
                st["n"] += 1
                ops.copy(wk[0], wk[2])
                tm[2] = tm[0]
            tm[0] = t + h
            ne = min(st["k"] - (0 if accept else 1) + 1, st["n"])                      # TSBDF_Extrapolate
            c = _lagrange_vals(tm[0], tm[1:1 + ne])
            ops.set(0.0, wk[0])
            for i in range(ne):
                ops.axpy(c[i], wk[1 + i], wk[0])
            newton.append(stage(wk[0]))
            kl = min(st["k"], st["n"] - 1)                                             # TSBDF_VecLTE
            a = _lagrange_ders(tm[0], tm[:kl + 1]) + [0.0]
            b = _lagrange_ders(tm[0], tm[:kl + 2])
            ops.copy(wk[0], lte)
            for i in range(kl + 2):
                ops.axpy((a[i] - b[i]) / a[0], wk[i], lte)
            enorm = math.sqrt(ops.wrms2(wk[0], lte, opt.ts_atol, opt.ts_rtol) / n)
            ok, hnext = adapt_basic(h, enorm, accept, order=kl + 1)
            if ok:
                break
            accept = False
            rejected += 1
            h = hnext
        ops.copy(wk[0], Yacc)
        t += h
        steps.append((t, h, newton))
        h = match_step(t, hnext, tmax)
        restart = False
        k += 1
    if opt.ts_monitor:
        out("%d TS dt %s time %s" % (k, fmt_g(h), fmt_g(t)))
    ops.copy(Yacc, Y)
    ops.sync()
    seconds = time.perf_counter() - t0
    if opt.call_back_report:                                                           # pattern.c:127-135
        out("CALL-BACK REPORT")
        out("  solver type: bdf")
        out("  IFunction:   1  | IJacobian:   1")
        out("  RHSFunction: 1  | RHSJacobian: %d" % (0 if opt.no_rhsjacobian else 1))
    if opt.log_view:
        out("TSSolve %.6f s (%d steps, %d rejected)" % (seconds, k, rejected))
    rep = PatternReport(m=m, steps=steps, Y=Y, seconds=seconds, lines=lines)
    rep.rejected = rejected
    return rep


# ---------------------------------------------------------------------------------------------------------
# -ts_type arkimex: [PETSc] TSARKIMEX3 = ARK3(2)4L[2]SA (Kennedy & Carpenter 2003) + TSAdaptBasic + MATCHSTEP
# ---------------------------------------------------------------------------------------------------------
def _fr(a, b):
    return a / b


_G = _fr(1767732205903, 4055673282236)
ARK3_AI = ((0.0, 0.0, 0.0, 0.0), (_G, _G, 0.0, 0.0),
           (_fr(2746238789719, 10658868560708), _fr(-640167445237, 6845629431997), _G, 0.0),
           (_fr(1471266399579, 7840856788654), _fr(-4482444167858, 7529755066697), _fr(11266239266428, 11593286722821), _G))
ARK3_AE = ((0.0, 0.0, 0.0, 0.0), (_fr(1767732205903, 2027836641118), 0.0, 0.0, 0.0),
           (_fr(5535828885825, 10492691773637), _fr(788022342437, 10882634858940), 0.0, 0.0),
           (_fr(6485989280629, 16251701735622), _fr(-4246266847089, 9704473918619), _fr(10755448449292, 10357097424841), 0.0))
ARK3_B = ARK3_AI[3]
ARK3_BH = (_fr(2756255671327, 12835298489170), _fr(-10771552573575, 22201958757719), _fr(9247589265047, 10645013368117),
           _fr(2193209047091, 5459859503100))


def adapt_basic(h, enorm, prev_accept, order=3, safety=0.9, reject_safety=0.5, clip=(0.1, 10.0)):
    """[PETSc] TSAdaptChoose_Basic -> (accept, next h): the extra factor 1/2 only from the second consecutive rejection."""
    accept = enorm <= 1.0
    s = safety * (reject_safety if (not accept and not prev_accept) else 1.0)
    hfac = s * enorm ** (-1.0 / order) if enorm > 0.0 else float("inf")
    return accept, h * min(max(hfac, clip[0]), clip[1])


def match_step(t, hnext, tmax, fac=(0.01, 2.0)):
    """TS_EXACTFINALTIME_MATCHSTEP (pattern.c:118) as TSAdaptChoose applies it; t = time after the accepted step."""
    if t >= tmax:
        return hnext
    hmax, tend, out = tmax - t, t + hnext, hnext
    if tend > tmax:
        out = hmax
    if tend < tmax and hnext * fac[1] > hmax:
        out = hmax / 2.0
    if tend < tmax and hnext * (1.0 + fac[0]) > hmax:
        out = hmax
    return out


def _arkimex(ops, opt: PatternOptions, levels, m, out, lines) -> PatternReport:
    A = StageOperator(ops, levels, opt, imex=True)
    n = levels[0].n
    Y = levels[0].Y
    Z, R, d, Ynew, Yemb, zero = (ops.empty(n) for _ in range(6))
    Ys = [ops.empty(n) for _ in range(4)]
    FI = [ops.empty(n) for _ in range(4)]
    FE = [ops.empty(n) for _ in range(4)]
    work = [ops.empty(n) for _ in range(opt.gmres_restart + 1)]
    ops.set(0.0, zero)
    initial_state(ops, opt, m, Y)
    t0 = time.perf_counter()
    # ([PETSc] TSSolve: TS_EXACTFINALTIME_MATCHSTEP clips the first step to the final time)
    t, k, h, steps, rejected, ksp_total = 0.0, 0, min(opt.ts_dt, opt.ts_max_time), [], 0, 0
    tmax = opt.ts_max_time
    while t < tmax - 1e-12 * max(1.0, abs(tmax)) and k < opt.ts_max_steps:
        if opt.ts_monitor:
            out("%d TS dt %s time %s" % (k, fmt_g(h), fmt_g(t)))
        prev_accept = True
        while True:
            for i in range(4):
                ops.copy(Y, Z)
                for j in range(i):
                    if ARK3_AE[i][j] != 0.0:
                        ops.axpy(h * ARK3_AE[i][j], FE[j], Z)
                    if ARK3_AI[i][j] != 0.0:
                        ops.axpy(h * ARK3_AI[i][j], FI[j], Z)
                if ARK3_AI[i][i] == 0.0:                           # explicit first stage: Y_1 = Z, YdotI = -F(Y_1, 0)
                    ops.copy(Z, Ys[i])
                    ops.pattern_ifunction(m, m, opt.L, opt.Du, opt.Dv, Ys[i], zero, FI[i])
                    ops.axpby(-1.0, FI[i], 0.0, None, FI[i])
                else:
                    # F(Y_i, shift (Y_i - Z)) = 0 with shift = 1/(h a_ii): linear; [PETSc] runs Newton on it (rtol 1e-8)
                    shift = 1.0 / (h * ARK3_AI[i][i])
                    if A.shift != shift or A.Ainv is None:
                        A.setup(shift)                             # the same shift for the three implicit stages
                    ops.copy(Ys[i - 1], Ys[i])                     # initial guess: the previous stage

                    def resid(W, f):
                        ops.axpby(shift, W, -shift, Z, d)
                        ops.pattern_ifunction(m, m, opt.L, opt.Du, opt.Dv, W, d, f)

                    resid(Ys[i], R)
                    r0 = rn = ops.norm2(R)
                    its = 0
                    while rn > opt.snes_rtol * r0 and rn > opt.snes_atol and its < opt.snes_max_it:
                        kr = gmres(ops, A.mult, R, d, A.precond, opt.ksp_rtol, restart=opt.gmres_restart,
                                   max_it=opt.ksp_max_it, work=work)
                        ksp_total += kr.its
                        ops.axpy(-1.0, d, Ys[i])
                        resid(Ys[i], R)
                        rn = ops.norm2(R)
                        its += 1
                    if not math.isfinite(rn) or rn > opt.snes_rtol * r0 and rn > opt.snes_atol:
                        raise RuntimeError("TSSolve: stage solve failed at step %d" % k)
                    ops.axpby(shift, Ys[i], -shift, Z, FI[i])      # YdotI = shift (Y_i - Z)
                ops.pattern_rhsfunction(m, m, opt.phi, opt.kappa, Ys[i], FE[i])
            ops.copy(Y, Ynew)
            ops.copy(Y, Yemb)
            for j in range(4):
                ops.axpy(h * ARK3_B[j], FI[j], Ynew)
                ops.axpy(h * ARK3_B[j], FE[j], Ynew)
                ops.axpy(h * ARK3_BH[j], FI[j], Yemb)
                ops.axpy(h * ARK3_BH[j], FE[j], Yemb)
            enorm = math.sqrt(ops.wrms2(Ynew, Yemb, opt.ts_atol, opt.ts_rtol) / n)     # TSErrorWeightedNorm2
            accept, hnext = adapt_basic(h, enorm, prev_accept)
            if accept:
                break
            prev_accept = False
            rejected += 1
            h = hnext
        ops.copy(Ynew, Y)
        t += h
        steps.append((t, h, enorm))
        h = match_step(t, hnext, tmax)
        k += 1
    if opt.ts_monitor:
        out("%d TS dt %s time %s" % (k, fmt_g(h), fmt_g(t)))
    ops.sync()
    seconds = time.perf_counter() - t0
    if opt.call_back_report:                                                                           # pattern.c:127-135
        out("CALL-BACK REPORT")
        out("  solver type: arkimex")
        out("  IFunction:   1  | IJacobian:   1")
        out("  RHSFunction: 1  | RHSJacobian: 0")              # IMEX: the reaction Jacobian is never needed
    if opt.log_view:
        out("TSSolve %.6f s (%d steps, %d rejected, %d GMRES iterations)" % (seconds, k, rejected, ksp_total))
    rep = PatternReport(m=m, steps=steps, Y=Y, seconds=seconds, lines=lines)
    rep.rejected = rejected
    return rep
