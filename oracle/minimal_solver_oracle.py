"""CPU oracle for the SOLVER around c/ch7/minimal.c  (TEST INFRASTRUCTURE ONLY -- never imported by the product).

minimal.c hands PETSc a residual callback (FormFunctionLocal, restated in minimal_pattern_oracle.py) and lets
`SNESSolve` do the rest (c/ch7/minimal.c:138-161).  This file restates, in NumPy/SciPy, the PETSc pieces the
reference's own runs use (c/ch7/makefile:15-25, c/ch8/cluster.sh:70), following SURVEY.md Appendix A10:

  fd_jacobian      [PETSc] SNESComputeJacobianDefaultColor / MatFDColoringApply with the default differencing "wp"
                   (Walker-Pernice): 9 colours of the DMDA BOX stencil, ONE step for every column,
                   h = sqrt(DBL_EPSILON) * sqrt(1 + ||u||_2).  Pinned by minimal.test1: with this h every printed digit
                   of its six residual norms is reproduced; the per-entry "ds" rule (h_m = eps*u_m, eps*umin at
                   u_m = 0) takes a visibly different Newton path from the zero interior (first FD Jacobian with ~1 %
                   rounding noise), and h = eps*(1 + ||u||) misses the last printed digit of three of the norms
  gmres            [PETSc] KSPGMRES(30), left preconditioning, preconditioned-residual norm, x0 = 0
  linesearch_bt    [PETSc] SNESLineSearchApply_BT, cubic backtracking, alpha = 1e-4, steptol 1e-12
  newton           [PETSc] SNESSolve_NEWTONLS + SNESConvergedDefault (rtol 1e-8, stol 1e-8, Jacobian every iteration)
  AssembledMG      [PETSc] PCMG on DM-provided operators: level Jacobians = the same FD Jacobian at the INJECTED
                   iterate, R = P^T (DMDA Q1), Chebyshev(2)/Jacobi smoothing, dense LU on the coarsest grid
  minimal          minimal.c:main incl. -snes_grid_sequence (DMRefine + Q1 interpolation of the iterate)

Pinned on the reference's goldens (tests/test_minimal_oracle.py): c/ch7/output/minimal.test1 (every Newton norm to
the printed precision, iteration count, error), minimal.test2 (CG+ILU iteration counts 5, 6), minimal.test4 (grid-sequenced Newton
iteration counts 3, 5, 5).  The Chebyshev/Jacobi MG variant has no golden (PETSc's default smoother PC is the
sequential SOR): parity unpinned there, as for fish.  Eigenvalue target: PETSc estimates lambda_max
with GMRES on a random right-hand side; the oracle and the device path both use the Gershgorin bound
max_i sum_j |a_ij| / |a_ii| of D^-1 A with PETSc's (0.1, 1.1) transform.
"""
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from . import fish_oracle as fo
from . import minimal_pattern_oracle as mpo

EPS_FD = 1.4901161193847656e-08      # PETSC_SQRT_MACHINE_EPSILON


def fd_step(u):
    """[PETSc] MatFDColoringApply, htype "wp": the differencing step, the same for every column."""
    return EPS_FD * float(np.sqrt(1.0 + np.sqrt(np.sum(np.asarray(u, dtype=np.float64) ** 2))))


def fd_jacobian(F, u, F0=None):
    """Coloured finite-difference Jacobian of F at u ((my, mx) array) as CSR; 9-point BOX pattern clipped at the edge.
    One evaluation of F per colour (i mod 3) + 3 (j mod 3)."""
    my, mx = u.shape
    N = mx * my
    if F0 is None:
        F0 = F(u)
    h = fd_step(u)
    vscale = 1.0 / h
    jj, ii = np.meshgrid(np.arange(my), np.arange(mx), indexing="ij")
    rows, cols, vals = [], [], []
    n = (jj * mx + ii)
    for cj in range(3):
        for ci in range(3):
            mask = (ii % 3 == ci) & (jj % 3 == cj)
            Fp = F(u + np.where(mask, h, 0.0))
            dF = Fp - F0
            # row (i, j) meets the column (i+di, j+dj) of this colour
            di = ci - ii % 3
            di = np.where(di > 1, di - 3, np.where(di < -1, di + 3, di))
            dj = cj - jj % 3
            dj = np.where(dj > 1, dj - 3, np.where(dj < -1, dj + 3, dj))
            i2, j2 = ii + di, jj + dj
            ok = (i2 >= 0) & (i2 < mx) & (j2 >= 0) & (j2 < my)
            m = (j2 * mx + i2)[ok]
            rows.append(n[ok])
            cols.append(m)
            vals.append(dF[ok] * vscale)
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))
    A.sort_indices()
    return A


def gmres(A, b, M, rtol=1e-5, abstol=1e-50, restart=30, max_it=10000):
    """Left-preconditioned restarted GMRES on M^-1 A x = M^-1 b from x = 0.  Returns (x, its, history of ||M^-1 r||)."""
    n = b.size
    x = np.zeros(n)
    r = M(b)
    beta = float(np.linalg.norm(r))
    hist = [beta]
    ttol = max(rtol * beta, abstol)
    its = 0
    while beta > ttol and its < max_it:
        V = np.zeros((restart + 1, n))
        H = np.zeros((restart + 1, restart))
        V[0] = r / beta
        gvec = np.zeros(restart + 1)
        gvec[0] = beta
        cs, sn = np.zeros(restart), np.zeros(restart)
        k = 0
        while k < restart and its < max_it:
            w = M(A @ V[k])
            for i in range(k + 1):            # modified Gram-Schmidt
                H[i, k] = float(w @ V[i])
                w = w - H[i, k] * V[i]
            H[k + 1, k] = float(np.linalg.norm(w))
            if H[k + 1, k] != 0.0:
                V[k + 1] = w / H[k + 1, k]
            for i in range(k):                # apply the previous Givens rotations
                t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]
                H[i, k] = t
            d = np.hypot(H[k, k], H[k + 1, k])
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
        x = x + V[:k].T @ y
        if beta <= ttol:
            break
        r = M(b - A @ x)
        beta = float(np.linalg.norm(r))
    return x, its, hist


def linesearch_bt(F, x, f, fnorm, y, Jy, alpha=1.0e-4, steptol=1.0e-12, maxstep=1.0e8, max_it=40):
    """SNESLineSearchApply_BT (cubic).  Search along -y from x; returns (x_new, f_new, fnorm_new, lambda)."""
    ynorm = float(np.linalg.norm(y))
    if ynorm == 0.0:
        return x, f, fnorm, 0.0
    if ynorm > maxstep:
        y = y * (maxstep / ynorm)
        Jy = Jy * (maxstep / ynorm)
        ynorm = maxstep
    minlambda = steptol / float(np.max(np.abs(y) / np.maximum(np.abs(x), 1.0)))
    initslope = float(f @ Jy)
    if initslope > 0.0:
        initslope = -initslope
    if initslope == 0.0:
        initslope = -1.0
    lam = 1.0
    w = x - lam * y
    g = F(w)
    gnorm = float(np.linalg.norm(g))
    if 0.5 * gnorm * gnorm <= 0.5 * fnorm * fnorm + lam * alpha * initslope:
        return w, g, gnorm, lam
    # quadratic fit
    lamprev, gnormprev = lam, gnorm
    lamtemp = -initslope / (gnorm * gnorm - fnorm * fnorm - 2.0 * initslope)
    lam = 0.5 * lam if lamtemp > 0.5 * lam else (0.1 * lam if lamtemp <= 0.1 * lam else lamtemp)
    w = x - lam * y
    g = F(w)
    gnorm = float(np.linalg.norm(g))
    if 0.5 * gnorm * gnorm < 0.5 * fnorm * fnorm + lam * alpha * initslope:
        return w, g, gnorm, lam
    for _ in range(max_it):
        if lam <= minlambda:
            raise RuntimeError("line search failed: lambda below minlambda")
        t1 = 0.5 * (gnorm * gnorm - fnorm * fnorm) - lam * initslope
        t2 = 0.5 * (gnormprev * gnormprev - fnorm * fnorm) - lamprev * initslope
        a = (t1 / (lam * lam) - t2 / (lamprev * lamprev)) / (lam - lamprev)
        b = (-lamprev * t1 / (lam * lam) + lam * t2 / (lamprev * lamprev)) / (lam - lamprev)
        d = b * b - 3.0 * a * initslope
        if d < 0.0:
            d = 0.0
        lamtemp = -initslope / (2.0 * b) if a == 0.0 else (-b + np.sqrt(d)) / (3.0 * a)
        lamprev, gnormprev = lam, gnorm
        lam = 0.5 * lam if lamtemp > 0.5 * lam else (0.1 * lam if lamtemp <= 0.1 * lam else lamtemp)
        w = x - lam * y
        g = F(w)
        gnorm = float(np.linalg.norm(g))
        if 0.5 * gnorm * gnorm < 0.5 * fnorm * fnorm + lam * alpha * initslope:
            return w, g, gnorm, lam
    raise RuntimeError("line search failed")


def gershgorin_jacobi(A):
    """max_i sum_j |a_ij| / |a_ii|: an upper bound of lambda_max(D^-1 A)."""
    A = sp.csr_matrix(A)
    return float(np.max(np.asarray(abs(A).sum(axis=1)).ravel() / np.abs(A.diagonal())))


class AssembledMG:
    """PCMG (multiplicative V cycle, R = P^T, same smoother pre and post) on assembled level matrices, finest first.
    Smoother: Chebyshev(smooth_its) + Jacobi with targets (0.1, 1.1) * gershgorin; coarsest: dense LU."""

    def __init__(self, mats, shapes, smooth_its=2):
        self.A = [sp.csr_matrix(a) for a in mats]
        self.P = []
        for (my, mx) in shapes[1:]:
            self.P.append(sp.csr_matrix(sp.kron(fo.interp1d(my), fo.interp1d(mx), format="csr")))   # coarse -> next finer
        self.eig = [(0.1 * gershgorin_jacobi(a), 1.1 * gershgorin_jacobi(a)) for a in self.A[:-1]]
        self.dinv = [1.0 / a.diagonal() for a in self.A]
        self.coarse = np.linalg.inv(self.A[-1].toarray())
        self.its = smooth_its

    def _smooth(self, l, b, x):
        class _J:
            def __init__(s, d): s.d = d
            def apply(s, r): return s.d * r
        return fo.chebyshev_smooth(self.A[l], _J(self.dinv[l]), b, x, self.eig[l][0], self.eig[l][1], self.its)

    def _cycle(self, l, b, x):
        if l == len(self.A) - 1:
            return self.coarse @ b
        x = self._smooth(l, b, x)
        r = b - self.A[l] @ x
        P = self.P[l]
        xc = self._cycle(l + 1, P.T @ r, np.zeros(P.shape[1]))
        x = x + P @ xc
        return self._smooth(l, b, x)

    def apply(self, r):
        return self._cycle(0, r, np.zeros_like(r))


@dataclass
class NewtonResult:
    u: np.ndarray
    its: int
    reason: str
    fnorms: list = field(default_factory=list)
    ksp_its: list = field(default_factory=list)
    lambdas: list = field(default_factory=list)


class MFFD:
    """[PETSc] MatMFFD with its default "wp" step (-snes_mf_operator): J v = (F(u + h v) - F(u)) / h,
    h = sqrt(eps) sqrt(1 + ||u||_2) / ||v||_2.  Pinned by c/ch7/output/minimal.test3: the areas MSEMonitor prints during the
    first grid-sequence stage (3 x 3: one unknown, the linear solve is exact whatever the preconditioner) are reproduced
    digit for digit with this operator and not with the FD-coloured matrix (they differ in the 8th digit)."""

    def __init__(self, F, u, f):
        self.F, self.u, self.f = F, u, f
        self.shape = (u.size, u.size)
        self.unorm = float(np.linalg.norm(u))

    def __matmul__(self, v):
        vn = float(np.linalg.norm(v))
        if vn == 0.0:
            return np.zeros_like(v)
        h = EPS_FD * np.sqrt(1.0 + self.unorm) / vn
        return (self.F((self.u.ravel() + h * v).reshape(self.u.shape)).ravel() - self.f.ravel()) / h


def newton(F, u0, make_pc, jac=None, ksp="gmres", snes_rtol=1.0e-8, snes_stol=1.0e-8, snes_atol=1.0e-50, max_it=50,
           ksp_rtol=1.0e-5, mf_operator=False, monitor=None):
    """SNESSolve_NEWTONLS with bt line search.  F maps (my, mx) -> (my, mx); make_pc(J, u) returns r -> M^-1 r.
    mf_operator: the Krylov operator is MFFD (above); J (jac or FD-coloured) only builds the preconditioner.
    monitor(it, u): called before the first and after every iteration ([PETSc] SNESMonitorSet)."""
    shape = u0.shape
    u = u0.copy()
    f = F(u)
    fnorm = float(np.linalg.norm(f))
    res = NewtonResult(u=u, its=0, reason="", fnorms=[fnorm])
    if monitor:
        monitor(0, u)
    if fnorm < snes_atol:
        res.reason = "CONVERGED_FNORM_ABS"
        return res
    ttol = snes_rtol * fnorm
    Ff = lambda v: F(v.reshape(shape)).ravel()
    for it in range(max_it):
        J = jac(u) if jac is not None else fd_jacobian(F, u, f)
        M = make_pc(J, u)
        A = MFFD(F, u, f) if mf_operator else J
        solver = gmres if ksp == "gmres" else fo.cg
        y, kits, _ = solver(A, f.ravel(), M, rtol=ksp_rtol)
        res.ksp_its.append(kits)
        try:
            xnew, fnew, fnormnew, lam = linesearch_bt(Ff, u.ravel(), f.ravel(), fnorm, y, A @ y)
        except RuntimeError:                   # [PETSc] SNES_DIVERGED_LINE_SEARCH: the solve stops at the current iterate
            res.reason = "DIVERGED_LINE_SEARCH"
            return res
        res.lambdas.append(lam)
        snorm = float(np.linalg.norm(xnew - u.ravel()))
        xnorm = float(np.linalg.norm(xnew))
        u = xnew.reshape(shape)
        f, fnorm = fnew.reshape(shape), fnormnew
        res.fnorms.append(fnorm)
        res.its = it + 1
        res.u = u
        if monitor:
            monitor(it + 1, u)
        if fnorm < snes_atol:
            res.reason = "CONVERGED_FNORM_ABS"
            return res
        if fnorm <= ttol:
            res.reason = "CONVERGED_FNORM_RELATIVE"
            return res
        if snorm < snes_stol * xnorm:
            res.reason = "CONVERGED_SNORM_RELATIVE"
            return res
    res.reason = "DIVERGED_MAX_IT"
    return res


def injection(u):
    return u[::2, ::2].copy()


def interpolate(uc):
    """DMDA Q1 interpolation of a coarse (my, mx) array to the refined grid (2my-1, 2mx-1)."""
    my, mx = uc.shape
    P = sp.kron(fo.interp1d(my), fo.interp1d(mx), format="csr")
    return (P @ uc.ravel()).reshape(2 * my - 1, 2 * mx - 1)


@dataclass
class MinimalResult:
    u: np.ndarray
    stages: list            # one NewtonResult per grid-sequence stage, coarsest first
    errinf: float | None
    mx: int
    my: int


def minimal(mx=3, my=3, refine=0, grid_sequence=0, problem="catenoid", q=-0.5, catenoid_c=1.1, tent_H=1.0,
            pc="ilu", ksp="gmres", mg_levels=0, smooth_its=2, snes_rtol=1.0e-8, ksp_rtol=1.0e-5, mf_operator=False,
            monitor=None, poisson_jacobian=False, max_it=50):
    """minimal.c:main with -da_grid_x mx -da_grid_y my -da_refine refine -snes_grid_sequence grid_sequence -snes_fd_color.
    pc: "ilu" (PETSc's default on one rank), "none", "mg" (Chebyshev/Jacobi PCMG, levels down to the base grid unless
    mg_levels).  poisson_jacobian: neither -snes_fd_color nor -snes_mf_operator -- Newton's matrix is the one minimal.c
    registers (minimal.c:142-145), Poisson2DJacobianLocal on the unit square (poissonfunctions.c:147-190), and "mg" is
    PCMG rediscretising THAT operator on every level (Chebyshev targets from the Gershgorin bound, as for the FD matrices)."""
    base = (my, mx)                                    # the -da_grid DMDA: the coarsest grid PCMG coarsens down to
    for _ in range(refine):
        mx, my = 2 * mx - 1, 2 * my - 1

    def problem_on(my_, mx_):
        g = mpo.minimal_g(mx_, my_, problem, tent_H, catenoid_c)
        return g, (lambda u: mpo.minimal_function(u, g, q))

    g, F = problem_on(my, mx)
    u = np.zeros((my, mx))
    bd = np.ones((my, mx), dtype=bool)
    bd[1:-1, 1:-1] = False
    u[bd] = g[bd]                                      # InitialState(ZEROS, gonboundary) (poissonfunctions.c:260-346)
    stages = []
    for stage in range(grid_sequence + 1):
        if stage > 0:
            u = interpolate(u)
            my, mx = u.shape
            g, F = problem_on(my, mx)
        shape = (my, mx)

        def make_pc(J, ucur, shape=shape):
            if pc == "ilu":
                return fo.ILU0PC(J).apply
            if pc == "none":
                return lambda r: r
            if pc == "mg":
                # levels: every grid from this one down to the base grid of the run ([PETSc] grid sequencing keeps the
                # whole DM hierarchy) or mg_levels of them; level Jacobians at the injected iterate
                shapes, us = [shape], [ucur]
                while (len(shapes) < mg_levels if mg_levels else shapes[-1] != base) and shapes[-1][0] > 3 \
                        and shapes[-1][1] > 3 and (shapes[-1][0] - 1) % 2 == 0 and (shapes[-1][1] - 1) % 2 == 0:
                    us.append(injection(us[-1]))
                    shapes.append(us[-1].shape)
                if len(shapes) == 1:
                    inv = np.linalg.inv(J.toarray())
                    return lambda r: inv @ r
                mats = [J]
                for s, uu in zip(shapes[1:], us[1:]):
                    if poisson_jacobian:                 # [PETSc] PCMG calls the registered callback on the coarsened DMs
                        mats.append(fo.jacobian(fo.Grid(2, (s[1], s[0], 1))))
                        continue
                    _, Fl = problem_on(s[0], s[1])
                    mats.append(fd_jacobian(Fl, uu))
                # AssembledMG wants interpolation shapes coarse -> finer for each finer level: shapes[1:], finest first
                mg = AssembledMG(mats, [None] + shapes[1:], smooth_its)
                return mg.apply
            raise ValueError(pc)

        jac = (lambda w, shape=shape: fo.jacobian(fo.Grid(2, (shape[1], shape[0], 1)))) if poisson_jacobian else None
        r = newton(F, u, make_pc, jac=jac, ksp=ksp, snes_rtol=snes_rtol, ksp_rtol=ksp_rtol, mf_operator=mf_operator,
                   max_it=max_it,
                   monitor=(lambda it, w, st=stage: monitor(st, it, w)) if monitor else None)
        stages.append(r)
        u = r.u
    errinf = None
    if problem == "catenoid" and q == -0.5:
        errinf = float(np.max(np.abs(u - g)))          # minimal.c:169-179 (g is the exact solution everywhere)
    return MinimalResult(u=u, stages=stages, errinf=errinf, mx=mx, my=my)


# ---------------------------------------------------------------------------------------------------------
# MSEMonitor (c/ch7/minimal.c:284-360, -ms_monitor): surface area and diffusivity bounds by tensor Gauss-Legendre
# quadrature of the Q1 interpolant (c/interlude/quadrature.h:14-25)
# ---------------------------------------------------------------------------------------------------------
GAUSS_LEGENDRE = {1: ((0.0,), (2.0,)),
                  2: ((-0.577350269189626, 0.577350269189626), (1.0, 1.0)),
                  3: ((-0.774596669241483, 0.0, 0.774596669241483), (0.555555555555556, 0.888888888888889, 0.555555555555556))}


def mse_monitor(u, q=-0.5, quaddegree=3):
    """(area, Dmin, Dmax) of the iterate u ((my, mx) array on the unit square)."""
    my, mx = u.shape
    hx, hy = 1.0 / (mx - 1), 1.0 / (my - 1)
    xi, w = GAUSS_LEGENDRE[quaddegree]
    a11, a10, a01, a00 = u[1:, 1:], u[1:, :-1], u[:-1, 1:], u[:-1, :-1]      # au[j][i], au[j][i-1], au[j-1][i], au[j-1][i-1]
    area, Dmin, Dmax = 0.0, np.inf, 0.0
    for r, wr in zip(xi, w):
        dx = hx * 0.5 * (r + 1.0)                       # x - (x_i - hx)
        for s_, ws in zip(xi, w):
            dy = hy * 0.5 * (s_ + 1.0)                  # y - (y_j - hy)
            ux = ((a11 - a10) * dy + (a01 - a00) * (hy - dy)) / (hx * hy)
            uy = ((a11 - a01) * dx + (a10 - a00) * (hx - dx)) / (hx * hy)
            W = ux * ux + uy * uy
            D = np.power(1.0 + W, q)
            Dmin, Dmax = min(Dmin, float(D.min())), max(Dmax, float(D.max()))
            area += wr * ws * float(np.sum(np.sqrt(1.0 + W)))
    return area * hx * hy / 4.0, Dmin, Dmax
