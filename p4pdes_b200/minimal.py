"""Host-side mirror of c/ch7/minimal.c on top of the C ABI (include/p4b200.h): Newton-Krylov-multigrid on the device.

`minimal_main(argv)` takes the reference's own command line (c/ch7/minimal.c:69-103 for the `-ms_*` options, and the
PETSc options its makefile / c/ch8/cluster.sh:70 use: `-da_grid_x/_y -da_refine -snes_grid_sequence -snes_fd_color
-pc_type mg -ksp_type ...`), runs the sequence of minimal.c:main -- DMDA, InitialState, SNESSolve, error norm -- and
prints the reference's lines.  What PETSc does inside SNESSolve is restated here over device kernels only:

  residual            p4b_minimal_function      c/ch7/minimal.c:210-282
  Jacobian            p4b_minimal_jacobian_fd   [PETSc] -snes_fd_color (9-colour MatFDColoring), on EVERY multigrid level
                                                at the injected iterate ([PETSc] DM-provided PCMG operators)
  Newton + line search                          [PETSc] SNESSolve_NEWTONLS + SNESLineSearchApply_BT (cubic)
  Krylov              GMRES(30) | CG            [PETSc] KSPGMRES (left preconditioning) / KSPCG; dots and AXPYs are
                                                p4b_vec_* kernels
  preconditioner      V cycle on the assembled level Jacobians: p4b_stencil9_lin (Chebyshev + Jacobi), p4b_restrict /
                      p4b_prolong_add (DMDA Q1, R = P^T), p4b_dense_matvec on the base grid ([PETSc] PCMG / PCLU)
  grid sequencing     p4b_prolong_add of the iterate onto the refined grid ([PETSc] -snes_grid_sequence)

Every vector lives in HBM; the host sees scalars only.  There is no CPU path: `ops` must be a device Context.
(tests/ substitutes a NumPy stand-in for `ops` to exercise this file's control flow without a GPU -- test
infrastructure, never used by the product.)  PETSc's default smoother PC (SOR) and default PC (ILU) are sequential
and are not provided: `-pc_type mg|none`, Chebyshev/Jacobi smoothing.
"""
from __future__ import annotations

import math
import shlex
import time
from dataclasses import dataclass, field

import numpy as np

PROBLEMS = {"tent": 0, "catenoid": 1}


# ---------------------------------------------------------------------------------------------------------
# options (minimal.c:69-103 + the PETSc options of c/ch7/makefile, c/ch8/cluster.sh:70)
# ---------------------------------------------------------------------------------------------------------
@dataclass
class MinimalOptions:
    problem: str = "catenoid"
    q: float = -0.5
    catenoid_c: float = 1.1
    tent_H: float = 1.0
    exact_init: bool = False
    grid_x: int = 3
    grid_y: int = 3
    refine: int = 0
    grid_sequence: int = 0
    fd_color: bool = False
    mf_operator: bool = False
    poisson_jacobian: bool = False      # set by parse_options when neither of the two is given (minimal.c's default)
    mf_pmat: str = "fd"                 # -p4b_mf_pmat fd|poisson: what -snes_mf_operator preconditions with (newton())
    ksp_type: str = "gmres"
    ksp_rtol: float = 1.0e-5
    ksp_max_it: int = 10000
    gmres_restart: int = 30
    pc_type: str = "mg"
    mg_levels: int = 0
    smooth_its: int = 2
    snes_rtol: float = 1.0e-8
    snes_stol: float = 1.0e-8
    snes_atol: float = 1.0e-50
    snes_max_it: int = 50
    snes_monitor: bool = False
    snes_monitor_short: bool = False
    snes_converged_reason: bool = False
    ksp_converged_reason: bool = False
    log_view: bool = False


def parse_options(argv) -> MinimalOptions:
    if isinstance(argv, str):
        argv = shlex.split(argv)
    o = MinimalOptions()
    flags = {"-ms_exact_init": "exact_init", "-snes_fd_color": "fd_color", "-snes_mf_operator": "mf_operator",
             "-snes_monitor": "snes_monitor",
             "-snes_monitor_short": "snes_monitor_short", "-snes_converged_reason": "snes_converged_reason",
             "-ksp_converged_reason": "ksp_converged_reason", "-log_view": "log_view"}
    valued = {"-ms_problem": ("problem", str), "-ms_q": ("q", float), "-ms_catenoid_c": ("catenoid_c", float),
              "-ms_tent_H": ("tent_H", float), "-da_grid_x": ("grid_x", int), "-da_grid_y": ("grid_y", int),
              "-da_refine": ("refine", int), "-snes_grid_sequence": ("grid_sequence", int), "-ksp_type": ("ksp_type", str),
              "-ksp_rtol": ("ksp_rtol", float), "-ksp_max_it": ("ksp_max_it", int),
              "-ksp_gmres_restart": ("gmres_restart", int), "-pc_type": ("pc_type", str),
              "-pc_mg_levels": ("mg_levels", int), "-mg_levels_ksp_max_it": ("smooth_its", int),
              "-snes_rtol": ("snes_rtol", float), "-snes_stol": ("snes_stol", float), "-snes_atol": ("snes_atol", float),
              "-snes_max_it": ("snes_max_it", int), "-p4b_mf_pmat": ("mf_pmat", str)}
    accepted = {"-mg_levels_ksp_type": ("chebyshev",), "-mg_levels_pc_type": ("jacobi",), "-snes_type": ("newtonls",)}
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
        elif a in accepted:
            if argv[i + 1] not in accepted[a]:
                raise ValueError("%s %s: the device path provides %s only (PETSc's sor/ilu are sequential)"
                                 % (a, argv[i + 1], "|".join(accepted[a])))
            i += 2
        else:
            raise ValueError("unknown or unsupported option %s" % a)
    if o.problem not in PROBLEMS:
        raise ValueError("unknown problem type")                                       # minimal.c:127
    if o.problem == "tent" and o.exact_init:
        raise ValueError("initialization with exact solution only possible for -mse_problem catenoid")   # :109
    if o.problem == "catenoid" and o.catenoid_c < 1.0:
        raise ValueError("catenoid exact solution only valid if c >= 1")               # :116
    if o.exact_init and o.q != -0.5:
        raise ValueError("initialization with catenoid exact solution only possible if q=-0.5")          # :120
    if o.ksp_type not in ("gmres", "cg"):
        raise ValueError("-ksp_type %s: the device path provides gmres and cg" % o.ksp_type)
    if o.pc_type not in ("mg", "none"):
        raise ValueError("-pc_type %s: the device path provides mg and none (ilu/sor/icc are sequential)" % o.pc_type)
    # neither -snes_fd_color nor -snes_mf_operator: [PETSc] uses the Jacobian minimal.c registers, which is Poisson's
    # ("ONLY APPROXIMATE", minimal.c:142-145): Newton with the constant 5-point matrix on every level (Level.assemble)
    o.poisson_jacobian = not (o.fd_color or o.mf_operator)
    # -snes_mf_operator: [PETSc] builds the preconditioner from the REGISTERED Jacobian, i.e. Poisson's ("-p4b_mf_pmat
    # poisson": nothing is differenced but the operator's action); the default here ("fd") keeps round 1's stronger
    # choice, the FD-coloured Jacobian of the residual itself
    if o.mf_pmat not in ("fd", "poisson"):
        raise ValueError("-p4b_mf_pmat %s: fd or poisson" % o.mf_pmat)
    if o.mf_pmat == "poisson" and not o.mf_operator:
        raise ValueError("-p4b_mf_pmat poisson is an option of -snes_mf_operator")
    return o


# ---------------------------------------------------------------------------------------------------------
# levels and operators
# ---------------------------------------------------------------------------------------------------------
class Level:
    """One grid of the hierarchy: boundary data g, the assembled Jacobian (stencil9 planes), smoother data, work."""

    def __init__(self, ops, mx, my, opt: MinimalOptions):
        self.ops, self.mx, self.my, self.n = ops, mx, my, mx * my
        self.g = ops.empty(self.n)
        ops.minimal_sample(mx, my, PROBLEMS[opt.problem], opt.tent_H, opt.catenoid_c, self.g)
        self.vals = ops.empty(9 * self.n)
        self.u = ops.empty(self.n)              # the iterate this level's Jacobian is evaluated at
        self.F = ops.empty(self.n)
        self.x, self.b, self.t = ops.empty(self.n), ops.empty(self.n), ops.empty(self.n)
        self.emin = self.emax = 0.0
        self.scale, self.omega = 0.0, []
        self.grid = ops.grid2d(mx, my)
        # which matrix the level carries: the FD-coloured Jacobian of the residual, or the one minimal.c registers --
        # Poisson2DJacobianLocal (minimal.c:142-145), the same at every iterate, so filled once
        self.poisson = opt.poisson_jacobian or (opt.mf_operator and opt.mf_pmat == "poisson")
        self.poisson_ready = False

    def assemble(self, q, F_known=False):
        """J = dF/du at self.u by coloured finite differences (or the registered Poisson matrix, see __init__)."""
        ops = self.ops
        if self.poisson:
            if not self.poisson_ready:
                ops.poisson_stencil9(self.mx, self.my, 1.0, 1.0, 1.0, 1.0, self.vals)     # unit square, cx = cy = 1
                self.poisson_ready = True
            return
        if not F_known:
            ops.minimal_function(self.mx, self.my, q, self.u, self.g, self.F)
        ops.minimal_jacobian_fd(self.mx, self.my, q, self.u, self.g, self.F, self.vals)

    def set_smoother(self, its):
        lam = self.ops.stencil9_gershgorin(self.mx, self.my, self.vals, self.t)
        self.emin, self.emax = 0.1 * lam, 1.1 * lam
        # [PETSc] KSPSolve_Chebyshev, first kind
        self.scale = 2.0 / (self.emax + self.emin)
        alpha = 1.0 - self.scale * self.emin
        mu = 1.0 / alpha
        omegaprod = 2.0 / alpha
        cm1, ck = 1.0, mu
        self.omega = []
        for _ in range(1, its):
            cp1 = 2.0 * mu * ck - cm1
            self.omega.append(omegaprod * ck / cp1)
            cm1, ck = ck, cp1

    def mult(self, x, y):
        self.ops.stencil9_apply(self.mx, self.my, self.vals, x, y)


class AssembledMG:
    """[PETSc] PCMG, multiplicative V cycle, on assembled level Jacobians; levels[0] is the finest."""

    def __init__(self, ops, levels, opt: MinimalOptions):
        self.ops, self.levels, self.its = ops, levels, opt.smooth_its
        self.Ainv = None

    def setup(self, q):
        """Level Jacobians at the injected iterate (levels[0].u and .F are current), smoothers, dense base-grid inverse."""
        ops = self.ops
        if self.levels[0].poisson and self.Ainv is not None:
            return                                       # the registered Poisson matrix does not depend on the iterate
        for l, L in enumerate(self.levels):
            if l > 0:
                ops.inject2d(L.mx, L.my, self.levels[l - 1].u, L.u)
            L.assemble(q, F_known=(l == 0))
            if l < len(self.levels) - 1:
                L.set_smoother(self.its)
        C = self.levels[-1]
        if C.n > 4225:
            raise ValueError("base grid of the multigrid hierarchy has %d nodes (> 65 x 65): use a coarser -da_grid_x/_y "
                             "or more levels" % C.n)
        dense = stencil9_to_dense(ops.to_host(C.vals), C.mx, C.my)
        self.Ainv = ops.from_host(np.linalg.inv(dense).ravel())          # [PETSc] PCLU on the coarsest level

    def _smooth(self, L, zero_guess):
        """Chebyshev(its) + Jacobi on L: A x = b.  x is updated in place; t is work."""
        ops, its = self.ops, self.its
        if its <= 0:
            if zero_guess:
                ops.set(0.0, L.x)
            return
        s = L.scale
        pm1, pk = L.x, L.t
        if zero_guess:
            ops.set(0.0, pm1)
        # p1 = p0 + s B (b - A p0)
        ops.stencil9_lin(L.mx, L.my, L.vals, pm1, L.b, None, 0.0, 1.0, s, True, pk)
        for i in range(1, its):
            w = L.omega[i - 1]
            # p+ = (1-w) p- + w p + w s B (b - A p)   (written over p-)
            ops.stencil9_lin(L.mx, L.my, L.vals, pk, L.b, pm1, 1.0 - w, w, w * s, True, pm1)
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
        ops.stencil9_lin(L.mx, L.my, L.vals, L.x, L.b, None, 0.0, 0.0, 1.0, False, L.t)      # t = b - A x
        ops.restrict(L.grid, L.t, C.b)
        self._cycle(l + 1, True)
        ops.prolong_add(L.grid, C.x, L.x)
        self._smooth(L, False)

    def apply(self, r, z):
        L = self.levels[0]
        self.ops.copy(r, L.b)
        self._cycle(0, True)
        self.ops.copy(L.x, z)


def stencil9_to_dense(vals, mx, my):
    """Host copy of a (small) stencil9 matrix as a dense array (base-grid LU, tests)."""
    N = mx * my
    A = np.zeros((N, N))
    v = np.asarray(vals).reshape(9, my, mx)
    for dj in (-1, 0, 1):
        for di in (-1, 0, 1):
            s = 3 * (dj + 1) + (di + 1)
            for j in range(max(0, -dj), min(my, my - dj)):
                for i in range(max(0, -di), min(mx, mx - di)):
                    A[j * mx + i, (j + dj) * mx + (i + di)] = v[s, j, i]
    return A


def stencil9_to_csr(vals, mx, my):
    """(rowptr, colind, values) int32/float64 host arrays of the stencil9 matrix with the 9-point pattern clipped at the
    grid edge: what p4b_sell_create takes to build the column-indexed SELL-32 copy ([PETSc] MatConvert to AIJ/SELL)."""
    N = mx * my
    v = np.asarray(vals).reshape(9, my, mx)
    jj, ii = np.meshgrid(np.arange(my), np.arange(mx), indexing="ij")
    rows, cols, data = [], [], []
    for dj in (-1, 0, 1):
        for di in (-1, 0, 1):
            ok = (ii + di >= 0) & (ii + di < mx) & (jj + dj >= 0) & (jj + dj < my)
            rows.append((jj * mx + ii)[ok])
            cols.append(((jj + dj) * mx + ii + di)[ok])
            data.append(v[3 * (dj + 1) + (di + 1)][ok])
    rows, cols, data = np.concatenate(rows), np.concatenate(cols), np.concatenate(data)
    order = np.lexsort((cols, rows))
    rows, cols, data = rows[order], cols[order], data[order]
    rowptr = np.zeros(N + 1, dtype=np.int32)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    return rowptr, cols.astype(np.int32), data.astype(np.float64)


# ---------------------------------------------------------------------------------------------------------
# Krylov solvers (device vectors, host scalars)
# ---------------------------------------------------------------------------------------------------------
@dataclass
class KSPResult:
    its: int
    reason: str
    history: list = field(default_factory=list)


def gmres(ops, mult, b, x, precond, rtol=1.0e-5, abstol=1.0e-50, restart=30, max_it=10000, work=None) -> KSPResult:
    """[PETSc] KSPGMRES: left-preconditioned, restarted, x0 = 0, convergence on the preconditioned residual norm."""
    n = b.numel()
    V = work if work is not None else [ops.empty(n) for _ in range(restart + 1)]
    w, t = ops.empty(n), ops.empty(n)
    ops.set(0.0, x)
    precond(b, V[0])
    beta = ops.norm2(V[0])
    hist = [beta]
    ttol = max(rtol * beta, abstol)
    its = 0
    if not math.isfinite(beta):
        return KSPResult(0, "DIVERGED_NANORINF", hist)
    while beta > ttol and its < max_it:
        H = np.zeros((restart + 1, restart))
        gvec = np.zeros(restart + 1)
        gvec[0] = beta
        cs, sn = np.zeros(restart), np.zeros(restart)
        ops.axpby(1.0 / beta, V[0], 0.0, None, V[0])
        k = 0
        while k < restart and its < max_it:
            mult(V[k], t)
            precond(t, w)
            for i in range(k + 1):                       # modified Gram-Schmidt
                H[i, k] = ops.dot(w, V[i])
                ops.axpy(-H[i, k], V[i], w)
            H[k + 1, k] = ops.norm2(w)
            if H[k + 1, k] != 0.0:
                ops.axpby(1.0 / H[k + 1, k], w, 0.0, None, V[k + 1])
            for i in range(k):
                tmp = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]
                H[i, k] = tmp
            d = math.hypot(H[k, k], H[k + 1, k])
            cs[k], sn[k] = H[k, k] / d, H[k + 1, k] / d
            H[k, k], H[k + 1, k] = d, 0.0
            gvec[k + 1] = -sn[k] * gvec[k]
            gvec[k] = cs[k] * gvec[k]
            beta = abs(gvec[k + 1])
            its += 1
            k += 1
            hist.append(beta)
            if beta <= ttol:
                break
        y = np.linalg.solve(np.triu(H[:k, :k]), gvec[:k])
        for i in range(k):
            ops.axpy(float(y[i]), V[i], x)
        if beta <= ttol:
            break
        mult(x, t)                                       # restart: r = M^-1 (b - A x)
        ops.axpby(1.0, b, -1.0, t, t)
        precond(t, V[0])
        beta = ops.norm2(V[0])
    return KSPResult(its, "CONVERGED_RTOL" if beta <= ttol else "DIVERGED_ITS", hist)


def cg(ops, mult, b, x, precond, rtol=1.0e-5, abstol=1.0e-50, max_it=10000) -> KSPResult:
    """[PETSc] KSPCG, preconditioned norm."""
    n = b.numel()
    r, z, p, w = ops.empty(n), ops.empty(n), ops.empty(n), ops.empty(n)
    ops.set(0.0, x)
    ops.copy(b, r)
    precond(r, z)
    beta = ops.dot(z, r)
    dp = ops.norm2(z)
    hist = [dp]
    ttol = max(rtol * dp, abstol)
    its, beta_old = 0, None
    while dp > ttol and its < max_it:
        if beta_old is None:
            ops.copy(z, p)
        else:
            ops.aypx(beta / beta_old, z, p)              # p = z + (beta/beta_old) p
        mult(p, w)
        a = beta / ops.dot(p, w)
        ops.axpy(a, p, x)
        ops.axpy(-a, w, r)
        precond(r, z)
        beta_old = beta
        beta = ops.dot(z, r)
        dp = ops.norm2(z)
        its += 1
        hist.append(dp)
    return KSPResult(its, "CONVERGED_RTOL" if dp <= ttol else "DIVERGED_ITS", hist)


# ---------------------------------------------------------------------------------------------------------
# Newton with backtracking line search
# ---------------------------------------------------------------------------------------------------------
def linesearch_bt(ops, F, x, f, fnorm, y, Jy, w, g, alpha=1.0e-4, steptol=1.0e-12, max_it=40):
    """[PETSc] SNESLineSearchApply_BT, cubic.  On return w = x - lambda y and g = F(w); returns (gnorm, lambda).
    (minlambda uses max|y_i| where PETSc uses max|y_i| / max(|x_i|, 1) <= it: the failure test is marginally laxer.)"""
    ynorm_inf = ops.norminf(y)
    if ynorm_inf == 0.0:
        ops.copy(x, w)
        ops.copy(f, g)
        return fnorm, 0.0
    minlambda = steptol / ynorm_inf
    initslope = ops.dot(f, Jy)
    if initslope > 0.0:
        initslope = -initslope
    if initslope == 0.0:
        initslope = -1.0

    def trial(lam):
        ops.axpby(1.0, x, -lam, y, w)
        F(w, g)
        return ops.norm2(g)

    lam = 1.0
    gnorm = trial(lam)
    if 0.5 * gnorm * gnorm <= 0.5 * fnorm * fnorm + lam * alpha * initslope:
        return gnorm, lam
    lamprev, gnormprev = lam, gnorm
    lamtemp = -initslope / (gnorm * gnorm - fnorm * fnorm - 2.0 * initslope)
    lam = 0.5 * lam if lamtemp > 0.5 * lam else (0.1 * lam if lamtemp <= 0.1 * lam else lamtemp)
    gnorm = trial(lam)
    if 0.5 * gnorm * gnorm < 0.5 * fnorm * fnorm + lam * alpha * initslope:
        return gnorm, lam
    for _ in range(max_it):
        if lam <= minlambda:
            raise RuntimeError("SNES line search failed (DIVERGED_LINE_SEARCH): lambda below minlambda")
        t1 = 0.5 * (gnorm * gnorm - fnorm * fnorm) - lam * initslope
        t2 = 0.5 * (gnormprev * gnormprev - fnorm * fnorm) - lamprev * initslope
        a = (t1 / (lam * lam) - t2 / (lamprev * lamprev)) / (lam - lamprev)
        b = (-lamprev * t1 / (lam * lam) + lam * t2 / (lamprev * lamprev)) / (lam - lamprev)
        d = max(b * b - 3.0 * a * initslope, 0.0)
        lamtemp = -initslope / (2.0 * b) if a == 0.0 else (-b + math.sqrt(d)) / (3.0 * a)
        lamprev, gnormprev = lam, gnorm
        lam = 0.5 * lam if lamtemp > 0.5 * lam else (0.1 * lam if lamtemp <= 0.1 * lam else lamtemp)
        gnorm = trial(lam)
        if 0.5 * gnorm * gnorm < 0.5 * fnorm * fnorm + lam * alpha * initslope:
            return gnorm, lam
    raise RuntimeError("SNES line search failed (DIVERGED_LINE_SEARCH)")


@dataclass
class SNESResult:
    its: int = 0
    reason: str = ""
    fnorms: list = field(default_factory=list)
    ksp_its: list = field(default_factory=list)
    lambdas: list = field(default_factory=list)


def newton(ops, levels, opt: MinimalOptions, out, indent=0) -> SNESResult:
    """[PETSc] SNESSolve_NEWTONLS on levels[0] (iterate in levels[0].u, updated in place)."""
    L = levels[0]
    q = opt.q
    n = L.n
    pad = "  " * indent
    F = lambda u, f: ops.minimal_function(L.mx, L.my, q, u, L.g, f)
    y, Jy, w, gnew = ops.empty(n), ops.empty(n), ops.empty(n), ops.empty(n)
    mfw = ops.empty(n) if opt.mf_operator else None
    mg = AssembledMG(ops, levels, opt) if opt.pc_type == "mg" and len(levels) > 1 else None
    work = [ops.empty(n) for _ in range(opt.gmres_restart + 1)] if opt.ksp_type == "gmres" else None
    F(L.u, L.F)
    fnorm = ops.norm2(L.F)
    res = SNESResult(fnorms=[fnorm])

    def monitor(it, v):
        if opt.snes_monitor_short:
            out("%s%3d SNES Function norm %s" % (pad, it, _g6(v)))
        elif opt.snes_monitor:
            out("%s%3d SNES Function norm %.12e" % (pad, it, v))

    monitor(0, fnorm)
    if fnorm < opt.snes_atol:
        res.reason = "CONVERGED_FNORM_ABS"
    ttol = opt.snes_rtol * fnorm
    it = 0
    while not res.reason:
        if it >= opt.snes_max_it:
            res.reason = "DIVERGED_MAX_IT"
            break
        if mg is not None:
            mg.setup(q)
            precond = mg.apply
        elif opt.mf_operator and opt.pc_type == "none":
            precond = lambda r, z: ops.copy(r, z)     # matrix-free and unpreconditioned: no matrix at all
        else:
            L.assemble(q, F_known=True)
            if opt.pc_type == "mg":                   # a single level: the "multigrid" is the direct base-grid solve
                Ainv = ops.from_host(np.linalg.inv(stencil9_to_dense(ops.to_host(L.vals), L.mx, L.my)).ravel())
                precond = lambda r, z: ops.dense_matvec(n, Ainv, r, z)
            else:
                precond = lambda r, z: ops.copy(r, z)
        mult = L.mult
        if opt.mf_operator:
            # [PETSc] MatMFFD, "wp": J v = (F(u + h v) - F(u)) / h, h = sqrt(eps) sqrt(1 + ||u||) / ||v||  (pinned by
            # c/ch7/output/minimal.test3, tests/test_minimal_oracle.py)
            unorm = ops.norm2(L.u)

            def mult(v, outv, unorm=unorm):
                vn = ops.norm2(v)
                if vn == 0.0:
                    ops.set(0.0, outv)
                    return
                h = 1.4901161193847656e-08 * math.sqrt(1.0 + unorm) / vn
                ops.axpby(1.0, L.u, h, v, mfw)
                F(mfw, outv)
                ops.axpby(1.0 / h, outv, -1.0 / h, L.F, outv)
        if opt.ksp_type == "gmres":
            k = gmres(ops, mult, L.F, y, precond, opt.ksp_rtol, restart=opt.gmres_restart, max_it=opt.ksp_max_it,
                      work=work)
        else:
            k = cg(ops, mult, L.F, y, precond, opt.ksp_rtol, max_it=opt.ksp_max_it)
        res.ksp_its.append(k.its)
        if opt.ksp_converged_reason:
            out("%s    Linear solve %s due to %s iterations %d" % (pad, "converged" if k.reason.startswith("CONV")
                                                                     else "did not converge", k.reason, k.its))
        mult(y, Jy)
        try:
            gnorm, lam = linesearch_bt(ops, F, L.u, L.F, fnorm, y, Jy, w, gnew)
        except RuntimeError:                          # [PETSc] stops the solve with a reason, it is not an error
            res.reason = "DIVERGED_LINE_SEARCH"
            break
        res.lambdas.append(lam)
        ops.axpby(1.0, w, -1.0, L.u, y)               # step actually taken (y is free now)
        snorm = ops.norm2(y)
        xnorm = ops.norm2(w)
        ops.copy(w, L.u)
        ops.copy(gnew, L.F)
        fnorm = gnorm
        it += 1
        res.its = it
        res.fnorms.append(fnorm)
        monitor(it, fnorm)
        if not math.isfinite(fnorm):
            res.reason = "DIVERGED_FNORM_NAN"
        elif fnorm < opt.snes_atol:
            res.reason = "CONVERGED_FNORM_ABS"
        elif fnorm <= ttol:
            res.reason = "CONVERGED_FNORM_RELATIVE"
        elif snorm < opt.snes_stol * xnorm:
            res.reason = "CONVERGED_SNORM_RELATIVE"
    if opt.snes_converged_reason:
        out("%s  Nonlinear solve %s due to %s iterations %d" % (pad, "converged" if res.reason.startswith("CONV")
                                                                 else "did not converge", res.reason, res.its))
    return res


def _g6(v):
    """How -snes_monitor_short prints norms ([PETSc] SNESMonitorDefaultShort): %g above 1e-9, %5.3e down to 1e-11, then
    '< 1.e-11' (c/ch7/output/minimal.test1:6 "1.772e-10")."""
    if v > 1.0e-9:
        return "%g" % v
    if v > 1.0e-11:
        return "%5.3e" % v
    return "< 1.e-11"


# ---------------------------------------------------------------------------------------------------------
# minimal.c:main
# ---------------------------------------------------------------------------------------------------------
@dataclass
class MinimalReport:
    mx: int
    my: int
    stages: list
    errinf: float | None
    u: object
    seconds: float
    lines: list


SNES_REASONS = {2: "CONVERGED_FNORM_ABS", 3: "CONVERGED_FNORM_RELATIVE", 4: "CONVERGED_SNORM_RELATIVE",
                -5: "DIVERGED_MAX_IT", -6: "DIVERGED_LINE_SEARCH", -4: "DIVERGED_FNORM_NAN"}


def _minimal_native(opt: MinimalOptions, ctx, out, keep_solution) -> MinimalReport:
    """The same run through ONE C-ABI call, p4b_minimal_solve (csrc/nk_device.cu + nk_solver.hpp): the host logic in C++
    inside the library, no interpreter between the kernels."""
    import ctypes as C

    from . import lib as L
    o = L.MinimalOpts()
    L.check(ctx.lib.p4b_minimal_default_opts(C.byref(o)))
    o.problem, o.q, o.catenoid_c, o.tent_H = PROBLEMS[opt.problem], opt.q, opt.catenoid_c, opt.tent_H
    o.exact_init, o.grid_x, o.grid_y, o.refine = int(opt.exact_init), opt.grid_x, opt.grid_y, opt.refine
    o.grid_sequence, o.ksp_type, o.ksp_rtol = opt.grid_sequence, {"gmres": 0, "cg": 1}[opt.ksp_type], opt.ksp_rtol
    o.ksp_max_it, o.gmres_restart, o.pc_type = opt.ksp_max_it, opt.gmres_restart, {"none": 0, "mg": 1}[opt.pc_type]
    o.mg_levels, o.smooth_its = opt.mg_levels, opt.smooth_its
    o.snes_rtol, o.snes_stol, o.snes_atol, o.snes_max_it = opt.snes_rtol, opt.snes_stol, opt.snes_atol, opt.snes_max_it
    o.snes_monitor = 2 if opt.snes_monitor_short else (1 if opt.snes_monitor else 0)
    o.snes_converged_reason, o.ksp_converged_reason = int(opt.snes_converged_reason), int(opt.ksp_converged_reason)
    o.mf_operator = int(opt.mf_operator)
    o.jacobian = int(opt.poisson_jacobian or (opt.mf_operator and opt.mf_pmat == "poisson"))
    mx, my = opt.grid_x, opt.grid_y
    for _ in range(opt.refine + opt.grid_sequence):
        mx, my = 2 * mx - 1, 2 * my - 1
    u = ctx.empty(mx * my) if keep_solution else None
    res = L.MinimalResult()
    cb = L.LINE_FN(lambda line, _ctx: out(line.decode()))
    t0 = time.perf_counter()
    L.check(ctx.lib.p4b_minimal_solve(ctx.h, C.byref(o), cb, None, u.data_ptr() if u is not None else None, mx * my,
                                      C.byref(res)))
    seconds = time.perf_counter() - t0
    stages = []
    for s in range(res.nstages):
        st = res.stage[s]
        stages.append(SNESResult(its=st.its, reason=SNES_REASONS.get(st.reason, str(st.reason)),
                                 fnorms=[st.fnorm[k] for k in range(st.its + 1)],
                                 ksp_its=[st.ksp_its[k] for k in range(st.its)], lambdas=[st.lam[k] for k in range(st.its)]))
    if opt.log_view:
        out("SNESSolve (all grid-sequence stages) %.6f s" % seconds)
    return MinimalReport(mx=res.mx, my=res.my, stages=stages, errinf=res.errinf if res.errinf >= 0 else None, u=u,
                         seconds=seconds, lines=None)


def minimal_main(argv, ops, echo=False, keep_solution=True, native=False) -> MinimalReport:
    """native=True: the whole run is one call of p4b_minimal_solve (host logic in C++ inside the library) instead of this
    file's loops over the individual C-ABI calls; same options, same lines, same report."""
    opt = parse_options(argv)
    lines = []

    def out(s):
        lines.append(s)
        if echo:
            print(s)

    if native:
        rep = _minimal_native(opt, ops, out, keep_solution)
        rep.lines = lines
        return rep

    mx, my = opt.grid_x, opt.grid_y
    for _ in range(opt.refine):
        mx, my = 2 * mx - 1, 2 * my - 1
    # the DM hierarchy PCMG sees: this grid and its coarsenings down to the -da_grid base (or -pc_mg_levels of them);
    # grid sequencing then refines the whole hierarchy ([PETSc] SNESSolve with -snes_grid_sequence)
    def hierarchy(mx_, my_, nmax):
        shapes = [(mx_, my_)]
        while (len(shapes) < nmax if nmax else True):
            cx, cy = shapes[-1]
            if cx <= 3 or cy <= 3 or (cx - 1) % 2 or (cy - 1) % 2:
                break
            if not nmax and (cx, cy) == (opt.grid_x, opt.grid_y):
                break
            shapes.append(((cx - 1) // 2 + 1, (cy - 1) // 2 + 1))
        return shapes

    t0 = time.perf_counter()
    stages = []
    u_prev, prev_grid = None, None
    for stage in range(opt.grid_sequence + 1):
        if stage > 0:
            mx, my = 2 * mx - 1, 2 * my - 1
        shapes = hierarchy(mx, my, opt.mg_levels) if opt.pc_type == "mg" else [(mx, my)]
        levels = [Level(ops, sx, sy, opt) for (sx, sy) in shapes]
        L = levels[0]
        if stage == 0:
            if opt.exact_init:
                ops.copy(L.g, L.u)                                     # FormExactFromG (minimal.c:191-208)
            else:
                ops.initial_state2d(L.grid, L.g, L.u)                  # InitialState(ZEROS, gonboundary) (:157)
        else:
            ops.set(0.0, L.u)
            ops.prolong_add(L.grid, u_prev, L.u)                       # [PETSc] DMRefine + MatInterpolate
        res = newton(ops, levels, opt, out, indent=opt.grid_sequence - stage)
        stages.append(res)
        u_prev, prev_grid = L.u, L.grid
    ops.sync()
    seconds = time.perf_counter() - t0
    L = levels[0]
    errinf = None
    msg = "done on %d x %d grid and problem %s" % (L.mx, L.my, opt.problem)            # minimal.c:166-167
    if opt.problem == "catenoid" and opt.q == -0.5:
        e = ops.empty(L.n)
        ops.axpby(1.0, L.u, -1.0, L.g, e)
        errinf = ops.norminf(e)
        out(msg + ":  error |u-uexact|_inf = %.5e" % errinf)                            # :177-178
    else:
        out(msg + " ...")                                                              # :180
    if opt.log_view:
        out("SNESSolve (all grid-sequence stages) %.6f s" % seconds)
    return MinimalReport(mx=L.mx, my=L.my, stages=stages, errinf=errinf, u=L.u if keep_solution else None,
                         seconds=seconds, lines=lines)
