"""CPU oracle for the implicit time stepping around c/ch5/pattern.c  (TEST INFRASTRUCTURE ONLY).

pattern.c hands PETSc the split system F(t,Y,Ydot) = G(t,Y) (IFunction / RHSFunction and their Jacobians, restated in
minimal_pattern_oracle.py) and calls TSSolve (c/ch5/pattern.c:99-125).  This file restates the PETSc pieces of the
implicit runs the reference uses (c/ch5/makefile:52-53 `-ts_type beuler -pc_type mg`):

  backward Euler   [PETSc] TSTHETA, theta = 1, fixed step (TSAdapt none), TS_EXACTFINALTIME_MATCHSTEP: each step solves
                   R(Y) = F(t+dt, Y, (Y - Y_n)/dt) - G(t+dt, Y) = 0 from the guess Y = Y_n
  Jacobian         [PETSc] TSComputeIJacobian: shift dF/dYdot + dF/dY - dG/dY, shift = 1/dt; without FormRHSJacobianLocal
                   (-ptn_no_rhsjacobian) the dG/dY block is missing and Newton runs with the approximate Jacobian
  Newton / GMRES   minimal_solver_oracle.newton / gmres (SNESNEWTONLS + bt, KSPGMRES(30) left-preconditioned)
  PCMG             periodic DMDA Q1 interpolation (ratio 2), R = P^T, level operators rediscretised on every level at
                   the injected iterate, Chebyshev(2)/Jacobi, dense LU on the base grid

  ARKIMEX          [PETSc] TSARKIMEX3 = ARK3(2)4L[2]SA (Kennedy & Carpenter 2003): 4 stages, F implicit (ESDIRK, L-stable),
                   G explicit, 2nd-order embedded method; TSAdaptBasic (safety 0.9, clip [0.1, 10], exponent -1/3, an extra
                   factor 1/2 only from the SECOND consecutive rejection on), TSErrorWeightedNorm2 with atol = rtol = 1e-4,
                   TS_EXACTFINALTIME_MATCHSTEP (last step stretched by <= 1 %, or the last two steps made equal)

Pinned on c/ch5/output/pattern.test2 (backward Euler: all lines verbatim; the golden's KSP count (3) also comes out of the
Chebyshev/Jacobi multigrid) and on c/ch5/output/pattern.test1 and pattern.test4 (ARKIMEX: every "TS dt ... time ..."
line of the adaptive runs, 13 and 11 lines, to all printed digits, including the rejected step of test1).
Crank-Nicolson: pattern.test3 verbatim.  BDF: only the restart step of pattern.test5 (see pattern_bdf_first_step).
"""
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from . import fish_oracle as fo
from . import minimal_pattern_oracle as mpo
from . import minimal_solver_oracle as mso


def interp1d_periodic(M):
    """(2M x M): fine 2I <- coarse I; fine 2I+1 <- (coarse I + coarse (I+1) mod M) / 2."""
    rows, cols, vals = [], [], []
    for I in range(M):
        rows.append(2 * I); cols.append(I); vals.append(1.0)
        rows += [2 * I + 1, 2 * I + 1]; cols += [I, (I + 1) % M]; vals += [0.5, 0.5]
    return sp.csr_matrix((vals, (rows, cols)), shape=(2 * M, M))


def interpolation(Mx, My):
    """P: coarse (My, Mx, 2) -> fine (2My, 2Mx, 2), interleaved components."""
    return sp.csr_matrix(sp.kron(sp.kron(interp1d_periodic(My), interp1d_periodic(Mx)), sp.eye(2), format="csr"))


def rhs_jacobian(Y, phi=0.024, kappa=0.06):
    """FormRHSJacobianLocal, pattern.c:202-236: 2 x 2 pointwise blocks, interleaved ordering."""
    u, v = Y[..., 0].ravel(), Y[..., 1].ravel()
    n = u.size
    uv, v2 = u * v, v * v
    k = np.arange(n)
    rows = np.concatenate([2 * k, 2 * k, 2 * k + 1, 2 * k + 1])
    cols = np.concatenate([2 * k, 2 * k + 1, 2 * k, 2 * k + 1])
    vals = np.concatenate([-v2 - phi, -2.0 * uv, v2, 2.0 * uv - (phi + kappa)])
    return sp.csr_matrix((vals, (rows, cols)), shape=(2 * n, 2 * n))


def stage_jacobian(Y, shift, rhsjac=True, L=2.5, Du=8.0e-5, Dv=4.0e-5, phi=0.024, kappa=0.06):
    my, mx, _ = Y.shape
    J = mpo.pattern_ijacobian(mx, my, shift, L, Du, Dv)
    if rhsjac:
        J = J - rhs_jacobian(Y, phi, kappa)
    return sp.csr_matrix(J)


class PatternMG:
    """V cycle on rediscretised level operators (finest first), Chebyshev(its)/Jacobi, Gershgorin targets."""

    def __init__(self, Y, shift, base, rhsjac, its=2, rscale=1.0, **par):
        self.A, self.P = [], []
        self.rscale = rscale        # 1: [PETSc] R = P^T; 0.25: averaging restriction, consistent with the pointwise
                                    # (finite-difference) scaling of pattern.c's equations
        Yl = Y
        while True:
            self.A.append(stage_jacobian(Yl, shift, rhsjac, **par))
            my, mx, _ = Yl.shape
            if mx <= base or mx % 2 or my % 2:
                break
            self.P.append(interpolation(mx // 2, my // 2))
            Yl = Yl[::2, ::2, :].copy()                       # DMCreateInjection
        self.eig = [(0.1 * mso.gershgorin_jacobi(a), 1.1 * mso.gershgorin_jacobi(a)) for a in self.A[:-1]]
        self.dinv = [1.0 / a.diagonal() for a in self.A]
        self.coarse = np.linalg.inv(self.A[-1].toarray())
        self.its = its

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
        xc = self._cycle(l + 1, self.rscale * (self.P[l].T @ r), np.zeros(self.P[l].shape[1]))
        return self._smooth(l, b, x + self.P[l] @ xc)

    def apply(self, r):
        return self._cycle(0, r, np.zeros_like(r))


@dataclass
class PatternResult:
    Y: np.ndarray
    mx: int
    steps: list = field(default_factory=list)        # (t_after, dt, NewtonResult)
    lines: list = field(default_factory=list)


def fmt_g(v):
    """PETSc's %g: an integral value prints with a trailing '.' ("5.", "200.")."""
    s = "%g" % v
    return s + "." if s.lstrip("-").isdigit() else s


def pattern_beuler(grid=3, refine=0, dt=5.0, tmax=200.0, pc="mg", rhsjac=True, snes_rtol=1.0e-8, ksp_rtol=1.0e-5,
                   smooth_its=2, max_steps=10000, rscale=1.0, theta=1.0, L=2.5, Du=8.0e-5, Dv=4.0e-5, phi=0.024,
                   kappa=0.06):
    """[PETSc] TSTHETA with fixed steps: theta = 1 backward Euler (TSBEULER), theta = 1/2 with the endpoint form =
    Crank-Nicolson (TSCN): solve  F(t+dt, Y, (Y - Y_n)/(theta dt)) - G(Y) + (1-theta)/theta [F(t, Y_n, 0) - G(Y_n)] = 0."""
    mx = grid * 2 ** refine
    par = dict(L=L, Du=Du, Dv=Dv, phi=phi, kappa=kappa)
    Y = mpo.pattern_initial_state(mx, mx, L)
    res = PatternResult(Y=Y, mx=mx)
    res.lines.append("running on %d x %d grid with square cells of side h = %.6f ..." % (mx, mx, L / mx))
    t, k = 0.0, 0
    while t < tmax - 1e-14 * max(1.0, abs(tmax)) and k < max_steps:
        step = min(dt, tmax - t)
        res.lines.append("%d TS dt %s time %s" % (k, fmt_g(step), fmt_g(t)))
        Y0 = Y.copy()
        shift = 1.0 / (theta * step)
        aff = 0.0
        if theta != 1.0:
            aff = ((1.0 - theta) / theta) * (mpo.pattern_ifunction(Y0, 0.0 * Y0, L, Du, Dv) - mpo.pattern_rhsfunction(Y0, phi, kappa))
        R = lambda W: (mpo.pattern_ifunction(W, (W - Y0) * shift, L, Du, Dv) - mpo.pattern_rhsfunction(W, phi, kappa)) + aff
        jac = lambda W: stage_jacobian(W, shift, rhsjac, **par)

        def make_pc(J, W):
            if pc == "none":
                return lambda r: r
            if pc == "ilu":
                return fo.ILU0PC(J).apply
            return PatternMG(W, shift, grid, rhsjac, smooth_its, rscale, **par).apply

        nr = mso.newton(R, Y, make_pc, jac=jac, snes_rtol=snes_rtol, ksp_rtol=ksp_rtol)
        Y = nr.u
        t += step
        k += 1
        res.steps.append((t, step, nr))
    res.lines.append("%d TS dt %s time %s" % (k, fmt_g(res.steps[-1][1] if res.steps else dt), fmt_g(t)))
    res.Y = Y
    return res


# ---------------------------------------------------------------------------------------------------------
# ARKIMEX 3 (pattern.c's default TS type: c/ch5/pattern.c:115)
# ---------------------------------------------------------------------------------------------------------
from fractions import Fraction as _Fr

_G = _Fr(1767732205903, 4055673282236)
ARK3_AI = np.array([[0, 0, 0, 0], [_G, _G, 0, 0],
                    [_Fr(2746238789719, 10658868560708), _Fr(-640167445237, 6845629431997), _G, 0],
                    [_Fr(1471266399579, 7840856788654), _Fr(-4482444167858, 7529755066697),
                     _Fr(11266239266428, 11593286722821), _G]], dtype=float)
ARK3_AE = np.array([[0, 0, 0, 0], [_Fr(1767732205903, 2027836641118), 0, 0, 0],
                    [_Fr(5535828885825, 10492691773637), _Fr(788022342437, 10882634858940), 0, 0],
                    [_Fr(6485989280629, 16251701735622), _Fr(-4246266847089, 9704473918619),
                     _Fr(10755448449292, 10357097424841), 0]], dtype=float)
ARK3_B = ARK3_AI[3].copy()
ARK3_BH = np.array([_Fr(2756255671327, 12835298489170), _Fr(-10771552573575, 22201958757719),
                    _Fr(9247589265047, 10645013368117), _Fr(2193209047091, 5459859503100)], dtype=float)
ARK3_C = np.array([0.0, float(_Fr(1767732205903, 2027836641118)), 0.6, 1.0])


def adapt_basic(h, enorm, prev_accept, order=3, safety=0.9, reject_safety=0.5, clip=(0.1, 10.0)):
    """[PETSc] TSAdaptChoose_Basic: (accept, next h)."""
    accept = enorm <= 1.0
    s = safety * (reject_safety if (not accept and not prev_accept) else 1.0)
    hfac = s * enorm ** (-1.0 / order) if enorm > 0.0 else np.inf
    return accept, h * min(max(hfac, clip[0]), clip[1])


def match_step(t, hnext, tmax, fac=(0.01, 2.0)):
    """TS_EXACTFINALTIME_MATCHSTEP in TSAdaptChoose: t = time after the accepted step."""
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


def pattern_arkimex(grid=3, refine=0, dt=5.0, tmax=200.0, atol=1.0e-4, rtol=1.0e-4, solve=None, max_steps=10000, L=2.5,
                    Du=8.0e-5, Dv=4.0e-5, phi=0.024, kappa=0.06):
    """pattern.c with its default TS type.  `solve(shift, rhs)` solves (shift I - C L9) y = rhs (default: sparse direct, which
    is what PETSc's Newton iteration on the linear stage equation converges to)."""
    import scipy.sparse.linalg as spla
    m = grid * 2 ** refine
    Y = mpo.pattern_initial_state(m, m, L)
    n = Y.size
    Lap = mpo.pattern_ijacobian(m, m, 0.0, L, Du, Dv)          # F(Y, Ydot) = Ydot + Lap Y   (Lap = -C L9)
    I = sp.identity(n, format="csr")
    if solve is None:
        solve = lambda shift, rhs: spla.spsolve((shift * I + Lap).tocsc(), rhs)
    res = PatternResult(Y=Y, mx=m)
    res.lines.append("running on %d x %d grid with square cells of side h = %.6f ..." % (m, m, L / m))
    t, k, h = 0.0, 0, min(dt, tmax)                 # [PETSc] TSSolve: MATCHSTEP clips the first step to the final time
    rejected = 0
    while t < tmax - 1e-12 * max(1.0, abs(tmax)) and k < max_steps:
        res.lines.append("%d TS dt %s time %s" % (k, fmt_g(h), fmt_g(t)))
        prev_accept = True
        while True:
            y0 = Y.ravel()
            FI, FE = [], []
            for i in range(4):
                Z = y0.copy()
                for j in range(i):
                    Z = Z + h * (ARK3_AE[i, j] * FE[j] + ARK3_AI[i, j] * FI[j])
                if ARK3_AI[i, i] == 0.0:
                    Yi = Z
                    fi = -(Lap @ Yi)                               # explicit first stage: YdotI = -F(Y, 0)
                else:
                    shift = 1.0 / (h * ARK3_AI[i, i])
                    Yi = solve(shift, shift * Z)                   # F(Yi, shift (Yi - Z)) = 0
                    fi = shift * (Yi - Z)
                FI.append(fi)
                FE.append(mpo.pattern_rhsfunction(Yi.reshape(m, m, 2), phi, kappa).ravel())
            ynew = y0 + h * sum(ARK3_B[j] * (FI[j] + FE[j]) for j in range(4))
            yemb = y0 + h * sum(ARK3_BH[j] * (FI[j] + FE[j]) for j in range(4))
            tol = atol + rtol * np.maximum(np.abs(ynew), np.abs(yemb))
            enorm = float(np.sqrt(np.sum(((ynew - yemb) / tol) ** 2) / n))
            accept, hnext = adapt_basic(h, enorm, prev_accept)
            if accept:
                break
            prev_accept = False
            rejected += 1
            h = hnext
        Y = ynew.reshape(m, m, 2)
        t += h
        res.steps.append((t, h, enorm))
        h = match_step(t, hnext, tmax)
        k += 1
    res.lines.append("%d TS dt %s time %s" % (k, fmt_g(h), fmt_g(t)))
    res.Y = Y
    res.rejected = rejected
    return res


# ---------------------------------------------------------------------------------------------------------
# BDF2, first step only (pattern.test5).  [PETSc] TSBDF restarts with a backward-Euler HALF step (t0 -> t0 + dt/2), then
# takes the BDF2 step over the nodes {t0 + dt, t0, t0 + dt/2} (variable-step weights = derivatives of the Lagrange basis
# at the new time); the local truncation error is the difference between the 2-node and 3-node derivative formulas,
# alpha = (a - b)/a_0, applied to the history (TSBDF_VecLTE, order k + 1 = 2), fed to TSAdaptBasic.  Later steps (order-2
# LTE over four nodes, history rotation) have no golden and are not restated; the device path does not offer bdf.
# ---------------------------------------------------------------------------------------------------------
def lagrange_basis_ders(t, T):
    """d/dt of the Lagrange basis polynomials over the nodes T, at t."""
    n = len(T)
    d = np.zeros(n)
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


def pattern_bdf_first_step(grid=3, refine=4, dt=1.0, atol=1.0e-4, rtol=1.0e-4, L=2.5, Du=8.0e-5, Dv=4.0e-5, phi=0.024,
                           kappa=0.06):
    """Returns (Newton iterations of the two solves, proposed next step, state after the step)."""
    m = grid * 2 ** refine
    par = dict(L=L, Du=Du, Dv=Dv, phi=phi, kappa=kappa)
    Y0 = mpo.pattern_initial_state(m, m, L)

    def stage(times, hist, guess):
        a = lagrange_basis_ders(times[0], times)
        aff = sum(a[i] * hist[i - 1] for i in range(1, len(times)))
        R = lambda W: mpo.pattern_ifunction(W, a[0] * W + aff, L, Du, Dv) - mpo.pattern_rhsfunction(W, phi, kappa)
        return mso.newton(R, guess, lambda J, W: fo.ILU0PC(J).apply, jac=lambda W: stage_jacobian(W, a[0], True, **par))

    half = stage([0.5 * dt, 0.0], [Y0], Y0)
    full = stage([dt, 0.0, 0.5 * dt], [Y0, half.u], half.u)
    T = [dt, 0.0, 0.5 * dt]
    a = np.append(lagrange_basis_ders(dt, T[:2]), 0.0)
    b = lagrange_basis_ders(dt, T)
    alpha = (a - b) / a[0]
    lte = alpha[0] * full.u + alpha[1] * Y0 + alpha[2] * half.u
    tol = atol + rtol * np.maximum(np.abs(full.u), np.abs(full.u + lte))
    enorm = float(np.sqrt(np.mean((lte / tol) ** 2)))
    _, hnext = adapt_basic(dt, enorm, True, order=2)
    return (half.its, full.its), hnext, full.u


# ---------------------------------------------------------------------------------------------------------
# BDF2, every step.  [PETSc] TSStep_BDF with the default order 2 (restated from the algorithm of src/ts/impls/bdf/bdf.c as
# described above; pattern.test5 pins the restart step -- Newton counts 3 and 2 and the proposed next step 1.10972 -- and
# nothing later: the steps after the first are checked by their order of accuracy only, "parity unpinned" there).
#   history  time[0] = new time, time[1..] = accepted states, newest first; after a restart time[2] = the half-step state
#   stage    Ydot = sum_i dL_i(time[0]) work[i] over the nodes time[0..k]  (Lagrange-basis derivatives; shift = dL_0)
#   guess    Lagrange extrapolation over min(k + 1, n) history states (k one lower after a rejection)
#   LTE      order kl = min(k, n - 1): alpha = (a - b)/a_0, a = derivative weights over kl + 1 nodes, b over kl + 2;
#            error estimate sum_i alpha_i work[i], weighted norm as TSErrorWeightedNorm, TSAdaptBasic with order kl + 1
# ---------------------------------------------------------------------------------------------------------
def lagrange_basis_vals(t, T):
    n = len(T)
    v = np.ones(n)
    for k in range(n):
        for j in range(n):
            if j != k:
                v[k] *= (t - T[j]) / (T[k] - T[j])
    return v


def pattern_bdf(grid=3, refine=0, dt=5.0, tmax=200.0, atol=1.0e-4, rtol=1.0e-4, snes_rtol=1.0e-8, max_steps=10000, L=2.5,
                Du=8.0e-5, Dv=4.0e-5, phi=0.024, kappa=0.06, order=2, adapt=True):
    m = grid * 2 ** refine
    par = dict(L=L, Du=Du, Dv=Dv, phi=phi, kappa=kappa)
    Y = mpo.pattern_initial_state(m, m, L)
    res = PatternResult(Y=Y, mx=m)
    res.lines.append("running on %d x %d grid with square cells of side h = %.6f ..." % (m, m, L / m))
    time, work = [0.0] * 8, [None] * 8
    state = dict(k=0, n=0)
    newton_counts = []

    def advance(t, X):
        for i in range(7, 1, -1):
            time[i], work[i] = time[i - 1], work[i - 1]
        state["n"] = min(state["n"] + 1, 7)
        time[1], work[1] = t, X.copy()

    def solve(guess):
        nn = max(state["k"], 1) + 1
        a = lagrange_basis_ders(time[0], time[:nn])
        aff = sum(a[i] * work[i] for i in range(1, nn))
        R = lambda W: mpo.pattern_ifunction(W, a[0] * W + aff, L, Du, Dv) - mpo.pattern_rhsfunction(W, phi, kappa)
        r = mso.newton(R, guess, lambda J, W: fo.ILU0PC(J).apply, jac=lambda W: stage_jacobian(W, a[0], True, **par),
                       snes_rtol=snes_rtol)
        newton_counts.append(r.its)
        res.lines.append("    Nonlinear solve converged due to %s iterations %d" % (r.reason, r.its))
        return r.u

    t, k, h = 0.0, 0, min(dt, tmax)
    restart, rejected = True, 0
    while t < tmax - 1e-12 * max(1.0, abs(tmax)) and k < max_steps:
        res.lines.append("%d TS dt %s time %s" % (k, fmt_g(h), fmt_g(t)))
        if not restart:
            state["k"] = min(state["k"] + 1, order)
            advance(t, Y)
        accept = True
        while True:
            if restart:
                state["k"], state["n"] = 1, 0
                advance(t, Y)
                time[0] = t + h / 2.0
                work[0] = solve(work[1])
                state["k"] = min(2, order)
                state["n"] += 1
                work[2], time[2] = work[0].copy(), time[0]
            time[0] = t + h
            ne = min(state["k"] - (0 if accept else 1) + 1, state["n"])
            c = lagrange_basis_vals(time[0], time[1:1 + ne])
            work[0] = solve(sum(c[i] * work[1 + i] for i in range(ne)))
            kl = min(state["k"], state["n"] - 1)
            a = np.append(lagrange_basis_ders(time[0], time[:kl + 1]), 0.0)
            b = lagrange_basis_ders(time[0], time[:kl + 2])
            alpha = (a - b) / a[0]
            lte = sum(alpha[i] * work[i] for i in range(kl + 2))
            tol = atol + rtol * np.maximum(np.abs(work[0]), np.abs(work[0] + lte))
            enorm = float(np.sqrt(np.mean((lte / tol) ** 2)))
            if adapt:
                ok, hnext = adapt_basic(h, enorm, accept, order=kl + 1)
            else:
                ok, hnext = True, h
            if ok:
                break
            accept = False
            rejected += 1
            h = hnext
        Y = work[0]
        t += h
        res.steps.append((t, h, enorm))
        h = match_step(t, hnext, tmax)
        restart = False
        k += 1
    res.lines.append("%d TS dt %s time %s" % (k, fmt_g(h), fmt_g(t)))
    res.Y, res.rejected, res.newton_counts = Y, rejected, newton_counts
    return res
