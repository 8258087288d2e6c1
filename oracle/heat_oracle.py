"""CPU oracle for c/ch5/heat.c  (TEST INFRASTRUCTURE ONLY -- never imported by the product).

heat.c solves u_t = D0 laplacian(u) + f on the unit square, non-homogeneous Neumann in x, periodic in y, by the method of
lines on a DMDA (5 x 4 nodes before -da_refine; x is vertex-centred with mx - 1 cells, y periodic with my cells) and lets
[PETSc] TS integrate X_t = G(t, X).  Restated here, in NumPy:

  rhs            FormRHSFunctionLocal  (c/ch5/heat.c:141-163): centred differences, the Neumann data enters through the
                 mirrored ghost value ul = u[i+1] + 2 hx gamma(y) at i = 0 and ur = u[i-1] at i = mx-1
  jacobian       FormRHSJacobianLocal  (c/ch5/heat.c:166-208), as a CSR matrix (rows i = 0, mx-1 carry 2 D/hx^2)
  energy         EnergyMonitor         (c/ch5/heat.c:105-138): trapezoid in x, rectangle in y; nu = D0 dt / (hx hy)
  rk3bs          [PETSc] TSRK, default scheme "3bs" (Bogacki-Shampine 3(2)) + TSAdaptBasic + MATCHSTEP
  theta          [PETSc] TSTHETA: backward Euler (theta = 1) / Crank-Nicolson endpoint rule, fixed steps; the linear stage
                 systems are solved exactly here (sparse LU)

Pinned on the reference's goldens (tests/test_heat_oracle.py): c/ch5/output/heat.test2 -- the adaptive RK3bs step
sequence 0.001, 0.00226419, 0.00336791, 0.00336791 and the proposed 0.00889336, every printed digit; this is what fixes the
controller's exponent to 1/3 (the order of the scheme; 1/2 gives 0.00359127) -- and c/ch5/output/heat.test1 (100 fixed
backward-Euler steps; only the times are printed).  The solution itself appears in no golden: the energy identity of
heat.c's help text (total heat is conserved for this f and gamma) and agreement of the integrators with each other are the
size-independent checks."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .pattern_solver_oracle import adapt_basic, match_step


def spacings(mx, my):
    """heat.c:99-103: hx = 1/(mx-1), hy = 1/my (periodic direction)."""
    return 1.0 / (mx - 1), 1.0 / my


def f_source(x, y):
    return 3.0 * np.exp(-25.0 * (x - 0.6) * (x - 0.6)) * np.sin(2.0 * np.pi * y)          # heat.c:16-19


def gamma_neumann(y):
    return np.sin(6.0 * np.pi * y)                                                          # heat.c:21-23


def rhs(u, D0=1.0):
    """G(u) for u of shape (my, mx)."""
    my, mx = u.shape
    hx, hy = spacings(mx, my)
    x = hx * np.arange(mx)[None, :]
    y = hy * np.arange(my)[:, None]
    ul = np.empty_like(u)
    ur = np.empty_like(u)
    ul[:, 1:] = u[:, :-1]
    ul[:, 0] = u[:, 1] + 2.0 * hx * gamma_neumann(y[:, 0])
    ur[:, :-1] = u[:, 1:]
    ur[:, -1] = u[:, -2]
    uxx = (ul - 2.0 * u + ur) / (hx * hx)
    uyy = (np.roll(u, 1, axis=0) - 2.0 * u + np.roll(u, -1, axis=0)) / (hy * hy)
    return D0 * (uxx + uyy) + f_source(x, y)


def jacobian(mx, my, D0=1.0):
    """dG/du, (mx my) x (mx my), natural ordering n = j mx + i."""
    hx, hy = spacings(mx, my)
    hx2, hy2 = hx * hx, hy * hy
    rows, cols, vals = [], [], []
    for j in range(my):
        for i in range(mx):
            n = j * mx + i
            rows.append(n); cols.append(n); vals.append(-2.0 * D0 * (1.0 / hx2 + 1.0 / hy2))
            for jj in ((j - 1) % my, (j + 1) % my):
                rows.append(n); cols.append(jj * mx + i); vals.append(D0 / hy2)
            if i == 0:
                rows.append(n); cols.append(n + 1); vals.append(2.0 * D0 / hx2)
            elif i == mx - 1:
                rows.append(n); cols.append(n - 1); vals.append(2.0 * D0 / hx2)
            else:
                rows += [n, n]; cols += [n - 1, n + 1]; vals += [D0 / hx2, D0 / hx2]
    return sp.csr_matrix((vals, (rows, cols)), shape=(mx * my, mx * my))        # (duplicates at my = 2 add up, as INSERT would not: my >= 3)


def energy(u):
    """EnergyMonitor's integral of u."""
    my, mx = u.shape
    hx, hy = spacings(mx, my)
    w = np.ones(mx)
    w[0] = w[-1] = 0.5
    return float((u * w[None, :]).sum() * hx * hy)


def fmt_g(v):
    """[PETSc] %g with the trailing '.' PetscVSNPrintf gives integer-valued reals."""
    s = "%g" % v
    return s if any(c in s for c in ".en") else s + "."


def wrms(a, b, atol, rtol):
    tol = atol + rtol * np.maximum(np.abs(a), np.abs(b))
    return float(np.sqrt(np.mean(((a - b) / tol) ** 2)))


RK3BS_A = ((0.0, 0.0, 0.0, 0.0), (0.5, 0.0, 0.0, 0.0), (0.0, 0.75, 0.0, 0.0), (2.0 / 9.0, 1.0 / 3.0, 4.0 / 9.0, 0.0))
RK3BS_BEMBED = (7.0 / 24.0, 0.25, 1.0 / 3.0, 0.125)


def rk3bs(G, u0, dt, tmax, atol=1.0e-4, rtol=1.0e-4, max_steps=5000, order=3, monitor=None):
    """[PETSc] TSSolve with TSRK 3bs.  Returns (u, lines, steps) with lines the -ts_monitor output."""
    u, t, h, k = u0.copy(), 0.0, min(dt, tmax), 0
    lines, steps = [], []
    hnext = h
    while t < tmax - 1e-12 * max(1.0, abs(tmax)) and k < max_steps:
        if monitor:
            monitor(k, t, h, u)
        lines.append("%d TS dt %s time %s" % (k, fmt_g(h), fmt_g(t)))
        prev = True
        while True:
            K = []
            for i in range(4):
                Z = u.copy()
                for j in range(i):
                    if RK3BS_A[i][j]:
                        Z = Z + h * RK3BS_A[i][j] * K[j]
                K.append(G(Z))
            unew = u + h * sum(RK3BS_A[3][j] * K[j] for j in range(3))
            uemb = u + h * sum(RK3BS_BEMBED[j] * K[j] for j in range(4))
            accept, hnext = adapt_basic(h, wrms(unew, uemb, atol, rtol), prev, order=order)
            if accept:
                break
            prev, h = False, hnext
        u, t = unew, t + h
        steps.append(h)
        h = match_step(t, hnext, tmax)
        k += 1
    if monitor:
        monitor(k, t, h, u)
    lines.append("%d TS dt %s time %s" % (k, fmt_g(h), fmt_g(t)))
    return u, lines, steps


def theta(u0, dt, tmax, D0=1.0, theta=1.0, monitor=None):
    """[PETSc] TSTHETA with fixed steps on the (affine) heat system: (I/(theta dt) - J) w = ..., solved exactly."""
    my, mx = u0.shape
    J = jacobian(mx, my, D0)
    n = mx * my
    c = rhs(np.zeros((my, mx)), D0).ravel()              # G(u) = J u + c
    u, t, k = u0.ravel().copy(), 0.0, 0
    lines = []
    last = dt
    lu = {}
    while t < tmax - 1e-14 * max(1.0, abs(tmax)):
        h = min(dt, tmax - t)
        last = h
        if monitor:
            monitor(k, t, h, u.reshape(my, mx))
        lines.append("%d TS dt %s time %s" % (k, fmt_g(h), fmt_g(t)))
        key = round(h / dt, 9)
        if key not in lu:
            lu[key] = spla.splu(sp.csc_matrix(sp.identity(n) / (theta * h) - J))
        # (w - u)/(theta h) = G(w) + (1-theta)/theta [G(u) - (u - u)/..]: endpoint rule on an affine system
        b = u / (theta * h) + c + (1.0 - theta) / theta * (J @ u + c)
        u = lu[key].solve(b)
        t += h
        k += 1
    if monitor:
        monitor(k, t, last, u.reshape(my, mx))
    lines.append("%d TS dt %s time %s" % (k, fmt_g(last), fmt_g(t)))
    return u.reshape(my, mx), lines


def heat(grid=(5, 4), refine=0, ts_type="bdf", dt=0.001, tmax=0.1, D0=1.0, monitor_energy=False):
    """heat.c:main for -ts_type beuler | cn | rk (heat.c's default, bdf, is not restated here)."""
    mx = 1 + (grid[0] - 1) * 2 ** refine          # [PETSc] -da_refine: non-periodic M <- 1 + 2^n (M-1), periodic M <- 2^n M
    my = grid[1] * 2 ** refine
    hx, hy = spacings(mx, my)
    lines = ["solving on %d x %d grid for t0=%s to tf=%s ..." % (mx, my, fmt_g(0.0), fmt_g(tmax))]      # heat.c:86-88
    body = []
    mon = None
    if monitor_energy:
        mon = lambda k, t, h, u: body.append(("  energy = %9.2e     nu = %8.4f" % (energy(u), D0 * h / (hx * hy)), len(body)))
    u0 = np.zeros((my, mx))                       # heat.c:91
    if ts_type == "rk":
        u, tl, _ = rk3bs(lambda w: rhs(w, D0), u0, dt, tmax, monitor=mon)
    elif ts_type in ("beuler", "cn"):
        u, tl = theta(u0, dt, tmax, D0, 1.0 if ts_type == "beuler" else 0.5, monitor=mon)
    else:
        raise ValueError(ts_type)
    if monitor_energy:                            # the user monitor was set first: its line precedes -ts_monitor's
        for (e, _), l in zip(body, tl):
            lines += [e, l]
    else:
        lines += tl
    return u, lines
