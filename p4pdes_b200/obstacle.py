"""Host-side mirror of c/ch12/obstacle.c on top of the C ABI: the obstacle problem -lap u = 0, u >= psi on
(-2,2)^2 by [PETSc] SNESVINEWTONRSLS -- a reduced-space active-set Newton method -- over the Poisson kernels of fish.c
(the reference reuses Poisson2DFunctionLocal / Poisson2DJacobianLocal, obstacle.c:108-112).

    obstacle_main("-da_refine 2 -snes_monitor_short -ksp_rtol 1.0e-12 -snes_rtol 1.0e-10 -pc_type none", ctx)

takes obstacle.c's command line (c/ch12/makefile:17-27) and prints its lines (obstacle.c:154-175).  Every vector lives in
HBM; per Newton step the device does: residual (p4b_poisson_function), inactive-set mask and VI norm
(p4b_vi_inactive_mask, p4b_vec_pointwise_mult), CG on the Jacobian restricted to the inactive set -- applied matrix-free
as mask .* A (mask .* x) with p4b_stencil_apply -- and the projected backtracking line search (p4b_vec_pointwise_max).
The algorithm is restated in oracle/obstacle_oracle.py, which reproduces c/ch12/output/obstacle.test1 completely.

psi, u_exact (the Dirichlet data) are the reference's host functions (obstacle.c:16-47): sampled here on the host, as
the shim samples a user's g_bdry.  What is NOT provided: preconditioners on the reduced matrix other than none / jacobi
([PETSc]'s default ILU(0) and the ASM+LU of the goldens are sequential; -pc_type mg on a reduced system needs
[PETSc]'s DMCoarsen-with-active-set machinery) and -snes_type vinewtonssls."""
from __future__ import annotations

import math
import shlex
import time
from dataclasses import dataclass, field

import numpy as np

from . import lib as L

AFREE, A_, B_ = 0.697965148223374, 0.680259411891719, 0.471519893402112      # obstacle.c:41-43


def psi_host(x, y):
    """obstacle.c:16-27."""
    r = np.sqrt(x * x + y * y)
    r0 = 0.9
    psi0 = math.sqrt(1.0 - r0 * r0)
    dpsi0 = -r0 / psi0
    return np.where(r <= r0, np.sqrt(np.maximum(1.0 - r * r, 0.0)), psi0 + dpsi0 * (r - r0))


def u_exact_host(x, y):
    """obstacle.c:40-47."""
    r = np.sqrt(x * x + y * y)
    return np.where(r <= AFREE, psi_host(x, y), -A_ * np.log(np.maximum(r, 1e-300)) + B_)


@dataclass
class ObstacleOptions:
    grid_x: int = 3
    grid_y: int = 3
    refine: int = 0
    grid_sequence: int = 0
    snes_rtol: float = 1.0e-8
    snes_stol: float = 1.0e-8
    snes_atol: float = 1.0e-50
    snes_max_it: int = 50
    ksp_rtol: float = 1.0e-5
    ksp_max_it: int = 10000
    pc_type: str = ""
    snes_monitor: bool = False
    snes_converged_reason: bool = False
    ksp_converged_reason: bool = False


@dataclass
class ObstacleReport:
    m: int
    its: int
    reason: str
    fnorm: list
    ksp_its: list
    err1: float
    errinf: float
    area_err: float
    seconds: float
    u: object = None
    lines: list = field(default_factory=list)


def parse_options(argv) -> ObstacleOptions:
    if isinstance(argv, str):
        argv = shlex.split(argv)
    o = ObstacleOptions()
    flags = {"-snes_monitor_short": "snes_monitor", "-snes_monitor": "snes_monitor",
             "-snes_converged_reason": "snes_converged_reason", "-ksp_converged_reason": "ksp_converged_reason"}
    valued = {"-da_grid_x": ("grid_x", int), "-da_grid_y": ("grid_y", int), "-da_refine": ("refine", int),
              "-snes_grid_sequence": ("grid_sequence", int), "-snes_rtol": ("snes_rtol", float),
              "-snes_max_it": ("snes_max_it", int), "-ksp_rtol": ("ksp_rtol", float), "-ksp_max_it": ("ksp_max_it", int),
              "-pc_type": ("pc_type", str)}
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
        elif a == "-snes_type" and argv[i + 1] == "vinewtonrsls":
            i += 2
        elif a == "-ksp_type" and argv[i + 1] == "cg":
            i += 2
        else:
            raise L.P4BError("unknown or unsupported option %s" % a)
    if o.pc_type not in ("none", "jacobi"):
        raise L.P4BError("obstacle on the device: -pc_type none or jacobi on the reduced system (PETSc's default ILU(0), "
                         "ASM + LU and multigrid on a reduced matrix are not provided)")
    if o.grid_x != o.grid_y:
        raise L.P4BError("the device path takes square grids (obstacle.c's default)")
    return o


def _short(x):
    if x > 1e-9:
        return "%g" % x
    if x > 1e-11:
        return "%5.3e" % x
    return "< 1.e-11"


class _Level:
    """One grid of the sequence: geometry, bounds, Dirichlet data and work vectors on the device."""

    def __init__(self, ctx, m):
        self.m, self.n = m, m * m
        self.grid = L.make_grid(2, (m, m), (4.0, 4.0, 1.0))
        x = -2.0 + 4.0 * np.arange(m) / (m - 1)
        X, Y = np.meshgrid(x, x)
        self.psi = ctx.from_host(psi_host(X, Y).ravel())
        self.g = ctx.from_host(u_exact_host(X, Y).ravel())           # g_fcn = u_exact (obstacle.c:49-52)
        self.zero = ctx.zeros(self.n)                                # f_rhs = 0 (:55-57)
        names = ("u", "F", "mask", "y", "w", "Fw", "maskw", "r", "z", "p", "Ap", "t")
        for nm in names:
            setattr(self, nm, ctx.empty(self.n))


def _vi_norm(ctx, Lv, u, F, mask, t):
    ctx.vi_inactive_mask(u, Lv.psi, F, mask)
    ctx.pointwise_mult(mask, F, t)
    return ctx.norm2(t)


def _reduced_cg(ctx, Lv, opt, rhs_masked):
    """CG on mask .* A (mask .* x) = mask .* F from x = 0 ([PETSc] KSPCG, preconditioned norm; PC none or Jacobi =
    a constant scaling here: the stencil's diagonal is constant).  The iterates stay in the inactive subspace."""
    d = 1.0 if opt.pc_type == "none" else 1.0 / 4.0      # diag = 2 (scx + scy) = 4 for hx = hy
    x, r, z, p, Ap, t = Lv.y, Lv.r, Lv.z, Lv.p, Lv.Ap, Lv.t
    ctx.set(0.0, x)
    ctx.copy(rhs_masked, r)
    ctx.axpby(d, r, 0.0, None, z)
    beta = ctx.dot(z, r)
    dp = ctx.norm2(z)
    ttol = max(opt.ksp_rtol * dp, 1.0e-50)
    its = 0
    ctx.copy(z, p)
    while dp > ttol and its < opt.ksp_max_it:
        ctx.pointwise_mult(Lv.mask, p, t)                # (p is masked already; kept for clarity of the operator)
        ctx.stencil_apply(Lv.grid, t, Ap)
        ctx.pointwise_mult(Lv.mask, Ap, Ap)
        a = beta / ctx.dot(p, Ap)
        ctx.axpy(a, p, x)
        ctx.axpy(-a, Ap, r)
        ctx.axpby(d, r, 0.0, None, z)
        bnew = ctx.dot(z, r)
        dp = ctx.norm2(z)
        its += 1
        ctx.aypx(bnew / beta, z, p)
        beta = bnew
    return its, dp <= ttol


def _rsls(ctx, Lv, opt, out, indent):
    """[PETSc] SNESSolve_VINEWTONRSLS on one grid (oracle/obstacle_oracle.py:rsls); Lv.u holds the initial iterate."""
    pad = "  " * indent
    ctx.pointwise_max(Lv.u, Lv.psi, Lv.u)                                   # SNESVIProjectOntoBounds
    ctx.poisson_function(Lv.grid, Lv.u, Lv.zero, Lv.g, Lv.F)
    fnorm = _vi_norm(ctx, Lv, Lv.u, Lv.F, Lv.mask, Lv.t)
    norms, ksp = [fnorm], []
    f0 = fnorm
    if opt.snes_monitor:
        out("%s  0 SNES Function norm %s" % (pad, _short(fnorm)))
    its, reason = 0, "DIVERGED_MAX_IT"
    if fnorm < opt.snes_atol:
        reason = "CONVERGED_FNORM_ABS"
    while reason == "DIVERGED_MAX_IT" and its < opt.snes_max_it:
        ctx.pointwise_mult(Lv.mask, Lv.F, Lv.t)                              # F on the inactive set
        ctx.copy(Lv.t, Lv.Fw)                                               # (kept: _reduced_cg uses t as scratch)
        k, conv = _reduced_cg(ctx, Lv, opt, Lv.Fw)
        ksp.append(k)
        if opt.ksp_converged_reason:
            out("%s    Linear solve %s due to %s iterations %d" % (pad, "converged" if conv else "did not converge",
                                                                   "CONVERGED_RTOL" if conv else "DIVERGED_ITS", k))
        # initial slope (F, J y) over the inactive set
        ctx.stencil_apply(Lv.grid, Lv.y, Lv.Ap)
        ctx.pointwise_mult(Lv.mask, Lv.Ap, Lv.Ap)
        fy = ctx.dot(Lv.Fw, Lv.Ap)
        slope = -fy if fy > 0 else -fnorm * fnorm
        lam, ok, lamprev, gprev = 1.0, False, None, None
        for _ in range(40):                                                 # SNESLineSearchApply_BT on the projected path
            ctx.axpby(1.0, Lv.u, -lam, Lv.y, Lv.w)
            ctx.pointwise_max(Lv.w, Lv.psi, Lv.w)
            ctx.poisson_function(Lv.grid, Lv.w, Lv.zero, Lv.g, Lv.r)        # F(w) -> r (free after the linear solve)
            gn = _vi_norm(ctx, Lv, Lv.w, Lv.r, Lv.maskw, Lv.t)
            if 0.5 * gn * gn <= 0.5 * fnorm * fnorm + 1.0e-4 * lam * slope:
                ok = True
                break
            if lamprev is None:
                lamnew = -slope / (gn * gn - fnorm * fnorm - 2.0 * slope)
            else:
                t1 = 0.5 * (gn * gn - fnorm * fnorm) - lam * slope
                t2 = 0.5 * (gprev * gprev - fnorm * fnorm) - lamprev * slope
                a = (t1 / (lam * lam) - t2 / (lamprev * lamprev)) / (lam - lamprev)
                b = (-lamprev * t1 / (lam * lam) + lam * t2 / (lamprev * lamprev)) / (lam - lamprev)
                dd = max(b * b - 3.0 * a * slope, 0.0)
                lamnew = -slope / (2.0 * b) if a == 0.0 else (-b + math.sqrt(dd)) / (3.0 * a)
            lamnew = min(max(lamnew, 0.1 * lam), 0.5 * lam)
            lamprev, gprev, lam = lam, gn, lamnew
        if not ok:
            reason = "DIVERGED_LINE_SEARCH"
            break
        ctx.axpby(1.0, Lv.w, -1.0, Lv.u, Lv.t)
        snorm, xnorm = ctx.norm2(Lv.t), ctx.norm2(Lv.w)
        ctx.copy(Lv.w, Lv.u)
        ctx.copy(Lv.r, Lv.F)
        ctx.copy(Lv.maskw, Lv.mask)
        fnorm = gn
        its += 1
        norms.append(fnorm)
        if opt.snes_monitor:
            out("%s  %d SNES Function norm %s" % (pad, its, _short(fnorm)))
        if fnorm < opt.snes_atol:
            reason = "CONVERGED_FNORM_ABS"
        elif fnorm <= opt.snes_rtol * f0:
            reason = "CONVERGED_FNORM_RELATIVE"
        elif snorm < opt.snes_stol * xnorm:
            reason = "CONVERGED_SNORM_RELATIVE"
    if opt.snes_converged_reason:
        out("%s  Nonlinear solve %s due to %s iterations %d" % (pad, "converged" if reason.startswith("CONV") else
                                                                "did not converge", reason, its))
    return its, reason, norms, ksp


def obstacle_main(argv, ctx, echo=False, keep_solution=False) -> ObstacleReport:
    opt = parse_options(argv)
    lines = []

    def out(s):
        lines.append(s)
        if echo:
            print(s)

    m = opt.grid_x
    for _ in range(opt.refine):
        m = 2 * m - 1
    t0 = time.perf_counter()
    Lv = _Level(ctx, m)
    ctx.set(0.0, Lv.u)                                                       # VecSet(u_initial, 0.0), obstacle.c:121
    its = reason = norms = ksp = None
    for stage in range(opt.grid_sequence + 1):
        if stage > 0:                                                        # [PETSc] -snes_grid_sequence: refine, interpolate
            fine = _Level(ctx, 2 * Lv.m - 1)
            ctx.set(0.0, fine.u)
            ctx.prolong_add(fine.grid, Lv.u, fine.u)
            Lv = fine
        its, reason, norms, ksp = _rsls(ctx, Lv, opt, out, opt.grid_sequence - stage)
    seconds = time.perf_counter() - t0
    out("done on %d x %d grid ... %s, SNES iters = %d, last KSP iters = %d" % (Lv.m, Lv.m, reason, its, ksp[-1] if ksp else 0))
    # obstacle.c:160-175: active area (inactive mask of the converged state), errors against u_exact
    ctx.vi_inactive_mask(Lv.u, Lv.psi, Lv.F, Lv.mask)
    nact = Lv.n - int(round(ctx.dot(Lv.mask, Lv.mask)))
    dx = 4.0 / (Lv.m - 1)
    exactarea = math.pi * AFREE * AFREE
    area_err = abs(dx * dx * nact - exactarea) / exactarea
    usol = ctx.from_host(ctx.to_host(Lv.u)) if keep_solution else None
    ctx.axpby(1.0, Lv.u, -1.0, Lv.g, Lv.t)
    errinf = ctx.norminf(Lv.t)
    err1 = float(np.sum(np.abs(ctx.to_host(Lv.t)))) / (Lv.m * Lv.m)           # NORM_1 / (mx my), obstacle.c:170-171 (host sum)
    out("errors: av |u-uexact| = %.3e, |u-uexact|_inf = %.3e, active area error = %.3f%%" % (err1, errinf, 100.0 * area_err))
    return ObstacleReport(Lv.m, its, reason, norms, ksp, err1, errinf, area_err, seconds, usol, lines)
