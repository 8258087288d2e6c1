"""CPU restatement of c/ch12/obstacle.c solved by [PETSc] SNESVINEWTONRSLS.  TEST INFRASTRUCTURE ONLY.

From the reference itself: psi, u_exact, the bounds (obstacle.c:16-47,236-255), the residual and Jacobian of
c/ch6/poissonfunctions.c (Poisson2DFunctionLocal :37-66, f = 0, g = u_exact on the boundary; the fish oracle's) on
(-2,2)^2, the active-set count and area (obstacle.c:187-219), the error report (:160-175).
From [PETSc] (src/snes/impls/vi/rs/virs.c, vi.c; un-vendored): the reduced-space active-set Newton method --
  project the initial iterate onto the bounds; F = F(u);
  inactive set I = { i : not (u_i <= psi_i + 1e-8 and F_i > 0) }   (upper bound = +infinity);  ||F||_VI = ||F_I||_2
  each iteration: solve J_II y_I = F_I (y = 0 on the active set), then a backtracking line search on ||F||_VI along
  the PROJECTED path  w(lambda) = max(u - lambda y, psi)  (SNESLineSearchBT with the VI projection and norm).
Pinned (tests/test_obstacle_oracle.py) on c/ch12/output/obstacle.test1 COMPLETELY (CG + ILU(0): the three monitored norms,
both KSP counts 11, the error line, the active-area error) and on obstacle.test2 COMPLETELY (4 ranks, GMRES + [PETSc]
PCASM with LU subdomain solves: four norms incl. the inexact-solve digits 4.84465e-06 and 1.511e-11, 3 iterations, last
KSP count 4) -- which pins PCASM as restricted additive Schwarz with overlap 1 (measured in the reduced matrix's own
graph) on the DMDA's 2 x 2 process grid (RASMPC below; 17 other readings -- unrestricted, overlap 0 or 2, 1 x 4 strips --
miss the golden's third norm), and on obstacle.test3 (-snes_grid_sequence 3 -pc_type mg: Newton counts 1, 2, 2, 3 of the four
grids, last KSP count 4, error line) with multigrid on the reduced systems (ReducedMG below: the inactive mask coarsened
by injection, reduced Q1 interpolation, Chebyshev(2)/SOR with the GMRES eigenvalue estimate; exact solves give 1, 1, 1, 2 and
a Jacobi smoother 1, 1, 2, 3 -- the golden prints counts only, so Galerkin and rediscretised reduced coarse operators are not
told apart), and on obstacle.test4 (-snes_grid_sequence 2 -snes_type vinewtonssls: Newton counts 4, 6, 5, last KSP count 6, error
line) with the semismooth method restated in ssls() below ([PETSc] src/snes/impls/vi/ss/vissls.c: Newton + backtracking
on the Fischer-Burmeister function of (u - psi, F(u)); without the initial projection onto the bounds the first count is 5)."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import fish_oracle as fo
from oracle import minimal_solver_oracle as mso

AFREE, A_, B_ = 0.697965148223374, 0.680259411891719, 0.471519893402112


def psi(x, y):
    r = np.sqrt(x * x + y * y)
    r0 = 0.9
    psi0 = np.sqrt(1.0 - r0 * r0)
    dpsi0 = -r0 / psi0
    return np.where(r <= r0, np.sqrt(np.maximum(1.0 - r * r, 0.0)), psi0 + dpsi0 * (r - r0))


def u_exact(x, y):
    r = np.sqrt(x * x + y * y)
    with np.errstate(divide="ignore"):
        return np.where(r <= AFREE, psi(x, y), -A_ * np.log(np.maximum(r, 1e-300)) + B_)


def grid_xy(m):
    x = -2.0 + 4.0 * np.arange(m) / (m - 1)
    return np.meshgrid(x, x)          # X[j, i] = x_i, Y[j, i] = y_j


def residual(u, g):
    """Poisson2DFunctionLocal with f = 0 on (-2,2)^2, hx = hy: scaled boundary rows, g substituted for boundary neighbours."""
    m = u.shape[0]
    F = 4.0 * (u - g)                  # boundary rows: scdiag (u - g), scx = scy = 1 for hx = hy
    v = u.copy()
    v[0, :], v[-1, :], v[:, 0], v[:, -1] = g[0, :], g[-1, :], g[:, 0], g[:, -1]
    F[1:-1, 1:-1] = 4.0 * u[1:-1, 1:-1] - v[1:-1, :-2] - v[1:-1, 2:] - v[:-2, 1:-1] - v[2:, 1:-1]
    return F


def jacobian(m):
    g = fo.Grid(2, (m, m, 1), (4.0, 4.0, 1.0))
    return fo.jacobian(g)


@dataclass
class ObstacleResult:
    m: int
    its: int
    fnorm: list
    ksp_its: list
    u: np.ndarray = field(repr=False, default=None)
    err1: float = 0.0
    errinf: float = 0.0
    area_err: float = 0.0
    reason: str = ""


def dmda_owner(m, px, py):
    """Rank that owns each node of an m x m DMDA on a px x py process grid ([PETSc]: the first m mod p ranks of a
    direction get one node more), natural ordering n = j m + i."""
    def split(n, p):
        base, rem = divmod(n, p)
        out, s = [], 0
        for r in range(p):
            c = base + (1 if r < rem else 0)
            out.append((s, s + c))
            s += c
        return out
    own = np.zeros((m, m), dtype=int)
    for ry, (y0, y1) in enumerate(split(m, py)):
        for rx, (x0, x1) in enumerate(split(m, px)):
            own[y0:y1, x0:x1] = ry * px + rx
    return own.ravel()


class RASMPC:
    """[PETSc] PCASM, defaults: one block per rank, overlap 1 (the block's rows plus everything they couple to in the
    matrix's graph), type RESTRICT (each block solves on its overlapped set but only its OWN rows of the result are
    added), -sub_pc_type lu (exact subdomain solves)."""

    def __init__(self, A, owner, overlap=1, restricted=True):
        A = sp.csr_matrix(A)
        self.n, self.blocks = A.shape[0], []
        for r in np.unique(owner):
            own = np.flatnonzero(owner == r)
            ext = set(own.tolist())
            for _ in range(overlap):
                ext |= {int(c) for i in ext for c in A.indices[A.indptr[i]:A.indptr[i + 1]]}
            ext = np.array(sorted(ext))
            keep = np.isin(ext, own) if restricted else np.ones(ext.size, dtype=bool)
            self.blocks.append((ext, spla.splu(sp.csc_matrix(A[ext][:, ext])), keep))

    def apply(self, r):
        z = np.zeros(self.n)
        for ext, lu, keep in self.blocks:
            z[ext[keep]] += lu.solve(r[ext])[keep]
        return z


class ReducedMG:
    """[PETSc] PCMG on the reduced system J_II of an RSLS step (DMSetVI / DMCoarsen on the inactive set): a coarse node is
    inactive iff the fine node it coincides with is; interpolation = the rows / columns of the DMDA Q1 interpolation that
    belong to the two inactive sets; coarse operators P^T A P; Chebyshev(2) + SOR smoothing with the targets (0.1, 1.1) x
    the GMRES estimate (fish_oracle.gmres_lambda_max); LU on the coarsest level."""

    def __init__(self, Jr, idx, m, nlevels, smoother="sor"):
        self.A, self.P = [sp.csr_matrix(Jr)], []
        mask = np.zeros(m * m, dtype=bool)
        mask[idx] = True
        for _ in range(1, nlevels):
            mc = (m - 1) // 2 + 1
            cmask = mask.reshape(m, m)[::2, ::2].ravel()
            P = sp.kron(fo.interp1d(mc), fo.interp1d(mc), format="csr")[np.flatnonzero(mask)][:, np.flatnonzero(cmask)]
            self.P.append(sp.csr_matrix(P))
            self.A.append(sp.csr_matrix(P.T @ self.A[-1] @ P))
            m, mask = mc, cmask
        self.coarse = spla.splu(sp.csc_matrix(self.A[-1]))
        self.pc, self.eig = [], []
        for A in self.A[:-1]:
            pc = fo.BlockSSORPC(A, [(0, A.shape[0])]) if smoother == "sor" else fo.JacobiPC(A)
            lam = fo.gmres_lambda_max(A, pc)
            self.pc.append(pc)
            self.eig.append((0.1 * lam, 1.1 * lam))

    def _cycle(self, l, b, x):
        if l == len(self.A) - 1:
            return self.coarse.solve(b)
        x = fo.chebyshev_smooth(self.A[l], self.pc[l], b, x, self.eig[l][0], self.eig[l][1], 2)
        xc = self._cycle(l + 1, self.P[l].T @ (b - self.A[l] @ x), np.zeros(self.P[l].shape[1]))
        return fo.chebyshev_smooth(self.A[l], self.pc[l], b, x + self.P[l] @ xc, self.eig[l][0], self.eig[l][1], 2)

    def apply(self, r):
        return self._cycle(0, r, np.zeros_like(r))


def vi_norm(u, F, lo):
    inact = ~((u <= lo + 1.0e-8) & (F > 0.0))
    return float(np.sqrt(np.sum(F[inact] ** 2))), inact


def rsls(m, u0=None, snes_rtol=1.0e-8, ksp_rtol=1.0e-5, pc="ilu", max_it=50, snes_stol=1.0e-8, snes_atol=1.0e-50,
         ranks=(2, 2), asm_overlap=1, asm_restricted=True, mg_levels=1, mg_smoother="sor"):
    """pc: "ilu" | "none" (KSPCG), "exact", "asm" = KSPGMRES(30) + RASMPC on the ranks[0] x ranks[1] process grid, or "mg" =
    KSPCG + ReducedMG with mg_levels levels (one level: the exact solve)."""
    owner = dmda_owner(m, *ranks) if pc == "asm" else None
    X, Y = grid_xy(m)
    lo, g = psi(X, Y), u_exact(X, Y)
    J = jacobian(m)
    u = np.maximum(np.zeros((m, m)) if u0 is None else u0, lo)            # SNESVIProjectOntoBounds
    F = residual(u, g)
    fnorm, inact = vi_norm(u, F, lo)
    norms, ksp_its = [fnorm], []
    f0 = fnorm
    its, reason = 0, "DIVERGED_MAX_IT"
    if fnorm < snes_atol:
        reason = "CONVERGED_FNORM_ABS"
    while reason == "DIVERGED_MAX_IT" and its < max_it:
        idx = np.flatnonzero(inact.ravel())
        Jr = sp.csr_matrix(J[idx][:, idx])
        rhs = F.ravel()[idx]
        if pc == "asm":
            y_i, k, _ = mso.gmres(Jr, rhs, RASMPC(Jr, owner[idx], asm_overlap, asm_restricted).apply, rtol=ksp_rtol)
        elif pc == "exact" or (pc == "mg" and (mg_levels < 2 or idx.size < 2)):
            y_i, k = spla.spsolve(sp.csc_matrix(Jr), rhs) if idx.size > 1 else rhs / Jr.toarray().ravel(), 1
        elif pc == "mg":
            y_i, k, _ = fo.cg(Jr, rhs, ReducedMG(Jr, idx, m, mg_levels, mg_smoother).apply, rtol=ksp_rtol)
        else:
            M = fo.ILU0PC(Jr).apply if pc == "ilu" else (lambda r: r.copy())
            y_i, k, _ = fo.cg(Jr, rhs, M, rtol=ksp_rtol)
        ksp_its.append(k)
        y = np.zeros(m * m)
        y[idx] = y_i
        y = y.reshape(m, m)
        # [PETSc] SNESLineSearchApply_BT with the VI projection: full step first, then quadratic / cubic backtracking
        lam, ok = 1.0, False
        gprev, lamprev = None, None
        fy = float(np.sum(F[inact] * (J @ y.ravel()).reshape(m, m)[inact]))      # initial slope (F, J y) on the inactive set
        slope = -fy if fy > 0 else -fnorm * fnorm
        for _ in range(40):
            w = np.maximum(u - lam * y, lo)
            Fw = residual(w, g)
            gn, inact_w = vi_norm(w, Fw, lo)
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
                d = b * b - 3.0 * a * slope
                d = max(d, 0.0)
                lamnew = -slope / (2.0 * b) if a == 0.0 else (-b + np.sqrt(d)) / (3.0 * a)
            lamnew = min(max(lamnew, 0.1 * lam), 0.5 * lam)
            lamprev, gprev, lam = lam, gn, lamnew
        if not ok:
            reason = "DIVERGED_LINE_SEARCH"
            break
        snorm = float(np.linalg.norm(w - u))
        xnorm = float(np.linalg.norm(w))
        u, F, fnorm, inact = w, Fw, gn, inact_w
        its += 1
        norms.append(fnorm)
        if fnorm < snes_atol:
            reason = "CONVERGED_FNORM_ABS"
        elif fnorm <= snes_rtol * f0:
            reason = "CONVERGED_FNORM_RELATIVE"
        elif snorm < snes_stol * xnorm:
            reason = "CONVERGED_SNORM_RELATIVE"
    # obstacle.c:160-175, 187-219
    act = int(np.sum((u <= lo + 1.0e-8) & (F > 0.0)))
    dx = 4.0 / (m - 1)
    exactarea = np.pi * AFREE * AFREE
    e = u - g
    return ObstacleResult(m, its, norms, ksp_its, u, float(np.sum(np.abs(e))) / (m * m), float(np.max(np.abs(e))),
                          abs(dx * dx * act - exactarea) / exactarea, reason)


def rsls_grid_sequence(nseq, base=3, **kw):
    """-snes_grid_sequence nseq from the base x base grid ([PETSc]: DMRefine + Q1 interpolation of the iterate, which rsls
    then projects onto the bounds); with pc="mg" stage s preconditions with s + 1 levels.  Returns the stages' results."""
    out, u, m = [], None, base
    for stage in range(nseq + 1):
        if u is not None:
            u = mso.interpolate(u)
            m = u.shape[0]
        r = rsls(m, u0=u, mg_levels=stage + 1, **kw)
        out.append(r)
        u = r.u
    return out


def fischer(a, b):
    """[PETSc] Fischer(a, b) = sqrt(a^2 + b^2) - (a + b), in the cancellation-free form PETSc evaluates."""
    n = np.sqrt(a * a + b * b)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(a + b <= 0.0, n - (a + b), -2.0 * a * b / np.where(n + a + b != 0.0, n + a + b, 1.0))


def ssls(m, u0=None, snes_rtol=1.0e-8, ksp_rtol=1.0e-5, pc="ilu", max_it=50, snes_stol=1.0e-8, snes_atol=1.0e-50, project=True):
    """[PETSc] SNESVINEWTONSSLS with a lower bound only: phi_i = Fischer(u_i - psi_i, F_i(u)); each iteration solves
    (Da + Db J) y = phi with Da = (u - psi)/n - 1, Db = F/n - 1, n = |(u - psi, F)| (an element of the B-subdifferential;
    (1/sqrt2 - 1) for both where u = psi and F = 0), KSPCG as obstacle.c:119 sets it + ILU(0), then SNESLineSearchBT on
    ||phi||_2 along u - lambda y (no projection of the trial points: phi itself carries the constraint)."""
    X, Y = grid_xy(m)
    lo, g = psi(X, Y), u_exact(X, Y)
    J = jacobian(m)
    u = np.zeros((m, m)) if u0 is None else u0.copy()
    if project:
        u = np.maximum(u, lo)                                   # SNESVIProjectOntoBounds

    def phi_of(w):
        Fw = residual(w, g)
        return fischer(w - lo, Fw), Fw

    phi, F = phi_of(u)
    pn = float(np.linalg.norm(phi))
    norms, ksp_its, p0 = [pn], [], pn
    its, reason = 0, "DIVERGED_MAX_IT"
    if pn < snes_atol:
        reason = "CONVERGED_FNORM_ABS"
    while reason == "DIVERGED_MAX_IT" and its < max_it:
        a, b = (u - lo).ravel(), F.ravel()
        n = np.sqrt(a * a + b * b)
        safe = np.where(n > 0.0, n, 1.0)
        da = np.where(n > 0.0, a / safe - 1.0, 1.0 / np.sqrt(2.0) - 1.0)
        db = np.where(n > 0.0, b / safe - 1.0, 1.0 / np.sqrt(2.0) - 1.0)
        Js = sp.csr_matrix(sp.diags(da) + sp.diags(db) @ J)
        if pc == "exact":
            y, k = spla.spsolve(sp.csc_matrix(Js), phi.ravel()), 1
        else:
            y, k, _ = fo.cg(Js, phi.ravel(), fo.ILU0PC(Js).apply if pc == "ilu" else (lambda r: r.copy()), rtol=ksp_rtol)
        ksp_its.append(k)
        try:
            xnew, _, _, _ = mso.linesearch_bt(lambda w: phi_of(w.reshape(m, m))[0].ravel(), u.ravel(), phi.ravel(), pn, y, Js @ y)
        except RuntimeError:
            reason = "DIVERGED_LINE_SEARCH"
            break
        snorm, xnorm = float(np.linalg.norm(xnew - u.ravel())), float(np.linalg.norm(xnew))
        u = xnew.reshape(m, m)
        phi, F = phi_of(u)
        pn = float(np.linalg.norm(phi))
        its += 1
        norms.append(pn)
        if pn < snes_atol:
            reason = "CONVERGED_FNORM_ABS"
        elif pn <= snes_rtol * p0:
            reason = "CONVERGED_FNORM_RELATIVE"
        elif snorm < snes_stol * xnorm:
            reason = "CONVERGED_SNORM_RELATIVE"
    act = int(np.sum((u <= lo + 1.0e-8) & (F > 0.0)))
    dx = 4.0 / (m - 1)
    exactarea = np.pi * AFREE * AFREE
    e = u - g
    return ObstacleResult(m, its, norms, ksp_its, u, float(np.sum(np.abs(e))) / (m * m), float(np.max(np.abs(e))),
                          abs(dx * dx * act - exactarea) / exactarea, reason)


def ssls_grid_sequence(nseq, base=3, **kw):
    out, u = [], None
    m = base
    for _ in range(nseq + 1):
        if u is not None:
            u = mso.interpolate(u)
            m = u.shape[0]
        out.append(ssls(m, u0=u, **kw))
        u = out[-1].u
    return out
