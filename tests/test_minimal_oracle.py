"""The minimal.c solver oracle (oracle/minimal_solver_oracle.py) against the reference's goldens
(c/ch7/output/minimal.test{1,2,4}; commands in c/ch7/makefile:15-25).

The differencing rule is what the goldens pin.  PETSc's MatFDColoring defaults to the "wp" step
h = sqrt(eps) sqrt(1 + ||u||_2), the same for every column; with it (and the reference's default GMRES + ILU(0))
EVERY printed digit of minimal.test1's residual history is reproduced, minimal.test2's CG counts are exactly 5 and 6,
minimal.test4's Newton counts exactly 3, 5, 5.  (The per-entry "ds" rule perturbs the zero interior of the initial
iterate by 1.5e-14 and takes a visibly different, libm-dependent Newton path: 6 iterations, second norm 1.00.)"""
import numpy as np
import pytest

from oracle import minimal_solver_oracle as mo


def g6(v):
    return "%g" % float("%.6g" % v)


def test_golden_minimal_test1():
    # -snes_fd_color -ms_problem catenoid -ms_catenoid_c 2.0 -da_refine 1   (default KSP GMRES(30) + ILU(0))
    r = mo.minimal(refine=1, problem="catenoid", catenoid_c=2.0, pc="ilu")
    s = r.stages[0]
    assert (r.mx, r.my) == (5, 5)
    assert [g6(f) for f in s.fnorms[:5]] == ["1.08276", "0.69656", "0.170569", "0.00995652", "2.20675e-05"]   # minimal.test1:1-5
    assert "%5.3e" % s.fnorms[5] == "1.772e-10"              # :6 (SNESMonitorDefaultShort prints %5.3e below 1e-9)
    assert s.reason == "CONVERGED_FNORM_RELATIVE" and s.its == 5   # :7
    assert "%.5e" % r.errinf == "1.10603e-04"                # minimal.test1:8
    assert s.fnorms[-1] <= 1e-8 * s.fnorms[0]
    assert s.lambdas[0] < 1.0 and s.lambdas[-1] == 1.0       # the first step is damped, the last ones are full


def test_golden_minimal_test2():
    # -snes_fd_color -ms_q 0.0 -ksp_type cg -da_refine 2 -ms_problem tent   (CG + ILU(0); q = 0: Laplace)
    r = mo.minimal(refine=2, problem="tent", q=0.0, ksp="cg", pc="ilu")
    s = r.stages[0]
    assert (r.mx, r.my) == (9, 9)
    assert s.its == 2 and s.ksp_its == [5, 6]                # minimal.test2:3,5: "iterations 5" then "iterations 6"
    # the FD Jacobian of this (linear) problem is symmetric to the tolerance the golden checks (-mat_is_symmetric 1e-7)
    g = mo.mpo.minimal_g(9, 9, "tent", 1.0, 1.1)
    J = mo.fd_jacobian(lambda u: mo.mpo.minimal_function(u, g, 0.0), r.u)
    assert abs(J - J.T).max() <= 1.0e-7 * abs(J).max()


def test_golden_minimal_test4():
    # -snes_fd_color -snes_grid_sequence 2 -ms_problem tent: Newton iterations 3, 5, 5 on 3x3, 5x5, 9x9
    r = mo.minimal(grid_sequence=2, problem="tent", pc="ilu")
    assert [s.its for s in r.stages] == [3, 5, 5]            # minimal.test4:1-3
    assert all(s.reason == "CONVERGED_FNORM_RELATIVE" for s in r.stages)
    assert (r.mx, r.my) == (9, 9)


def test_fd_jacobian_matches_analytic_derivative_away_from_zero():
    rng = np.random.default_rng(3)
    mx, my = 9, 7
    g = mo.mpo.minimal_g(mx, my, "catenoid", 1.0, 1.1)
    u = g + 0.1 * rng.standard_normal((my, mx))
    F = lambda w: mo.mpo.minimal_function(w, g, -0.5)
    J = mo.fd_jacobian(F, u).toarray()
    # central differences with a much larger step as an independent check
    h = 1e-6
    for n in rng.choice(mx * my, 12, replace=False):
        e = np.zeros(mx * my)
        e[n] = h
        col = (F(u + e.reshape(my, mx)) - F(u - e.reshape(my, mx))).ravel() / (2 * h)
        np.testing.assert_allclose(J[:, n], col, rtol=0, atol=2e-6 * max(1.0, np.abs(col).max()))
    # boundary rows are identity rows, interior rows do not couple to boundary columns (minimal.c:227,230-256)
    bd = np.ones((my, mx), bool)
    bd[1:-1, 1:-1] = False
    b = bd.ravel()
    assert np.allclose(J[b][:, b], np.eye(b.sum()), atol=1e-7)
    assert np.all(J[~b][:, b] == 0.0)


def test_multigrid_preconditioned_newton_is_mesh_independent():
    its = []
    for seq in (2, 3, 4):
        r = mo.minimal(grid_sequence=seq, problem="tent", pc="mg")
        its.append(max(r.stages[-1].ksp_its))
        assert r.stages[-1].reason == "CONVERGED_FNORM_RELATIVE"
    assert max(its) - min(its) <= 2 and max(its) <= 10


def test_golden_minimal_test3_path_independent_lines():
    """c/ch7/output/minimal.test3 (-snes_mf_operator -pc_type mg -snes_grid_sequence 2 -ms_monitor -ms_quaddegree 2, 2 ranks):
    MSEMonitor (minimal.c:284-360) prints area and diffusivity bounds of every Newton iterate.  The Newton path of that run
    (matrix-free operator, Poisson preconditioner) is not restated, but the lines of the INITIAL iterate, of each
    CONVERGED stage and of each INTERPOLATED stage start do not depend on it: they pin the quadrature monitor, the
    discretisation, and the DMDA Q1 interpolation -snes_grid_sequence uses."""
    from oracle import fish_oracle as fo
    fmt = lambda t: "area = %.8f; %.4f <= D <= %.4f" % t
    g = mo.mpo.minimal_g(3, 3, "catenoid", 1.0, 1.1)
    u = np.zeros((3, 3))
    u[[0, -1], :] = g[[0, -1], :]
    u[:, [0, -1]] = g[:, [0, -1]]
    got = [fmt(mo.mse_monitor(u, -0.5, 2))]
    for stage in range(3):
        if stage:
            u = mo.interpolate(u)
            got.append(fmt(mo.mse_monitor(u, -0.5, 2)))
        gg = mo.mpo.minimal_g(u.shape[1], u.shape[0], "catenoid", 1.0, 1.1)
        u = mo.newton(lambda w, gg=gg: mo.mpo.minimal_function(w, gg, -0.5), u, lambda J, uu: fo.ILU0PC(J).apply,
                      snes_rtol=1e-12).u
        got.append(fmt(mo.mse_monitor(u, -0.5, 2)))
    assert got == ["area = 2.14201032; 0.2985 <= D <= 0.8826",        # minimal.test3:1
                   "area = 1.32217567; 0.6158 <= D <= 0.9510",        # :5-6   (3 x 3 converged)
                   "area = 1.32217583; 0.6035 <= D <= 0.9518",        # :8     (interpolated to 5 x 5)
                   "area = 1.33230217; 0.5755 <= D <= 0.9872",        # :11
                   "area = 1.33230220; 0.5684 <= D <= 0.9872",        # :13    (interpolated to 9 x 9)
                   "area = 1.33475385; 0.5156 <= D <= 0.9968"]        # :15-16
    assert "%.5e" % float(np.max(np.abs(u - gg))) == "6.79501e-04"    # :18


def test_golden_minimal_test3_matrix_free_operator():
    """minimal.test3 ran -snes_mf_operator: the Krylov operator is [PETSc] MatMFFD, J v = (F(u + h v) - F(u)) / h with the
    default "wp" step h = sqrt(eps) sqrt(1 + ||u||) / ||v||.  On the first grid (3 x 3: ONE unknown, so the linear solve is
    exact whatever preconditioner the golden used) every printed digit of the monitor lines is reproduced with that
    operator -- and not with the FD-coloured matrix, which differs in the 8th digit; the later stages inherit the golden's
    inexact multigrid solves (2 ranks, Chebyshev/SOR) in the last digits of two lines."""
    fmt = lambda t: "area = %.8f; %.4f <= D <= %.4f" % t
    got = {True: [], False: []}
    for mf in (True, False):
        mo.minimal(grid_sequence=2, pc="ilu", mf_operator=mf, monitor=lambda st, it, u, mf=mf: got[mf].append(fmt(mo.mse_monitor(u, -0.5, 2))))
    ref = "/root/reference/c/ch7/output/minimal.test3"
    golden = ["area = 2.14201032; 0.2985 <= D <= 0.8826", "area = 1.60235989; 0.4166 <= D <= 0.9873",
              "area = 1.39125969; 0.5789 <= D <= 0.9362", "area = 1.32285324; 0.6106 <= D <= 0.9561",
              "area = 1.32217567; 0.6158 <= D <= 0.9510", "area = 1.32217567; 0.6158 <= D <= 0.9510",
              "area = 1.32217583; 0.6035 <= D <= 0.9518", "area = 1.33231935; 0.5757 <= D <= 0.9872",
              "area = 1.33230219; 0.5755 <= D <= 0.9872", "area = 1.33230217; 0.5755 <= D <= 0.9872",
              "area = 1.33230220; 0.5684 <= D <= 0.9872", "area = 1.33475595; 0.5153 <= D <= 0.9968",
              "area = 1.33475385; 0.5156 <= D <= 0.9968", "area = 1.33475385; 0.5156 <= D <= 0.9968"]
    import os
    if os.path.exists(ref):
        assert [l.strip() for l in open(ref) if "area" in l] == golden
    assert got[True][:7] == golden[:7]                       # minimal.test3:1-6,8: the whole first stage + the interpolated iterate
    assert got[False][1:4] != golden[1:4]                    # the assembled FD-coloured operator does not give these digits
    area = lambda l: float(l.split()[2].rstrip(";"))
    assert len(got[True]) == 14 and max(abs(area(a) - area(b)) for a, b in zip(got[True], golden)) <= 6e-8
    assert [l.split(";")[1] for l in got[True]] == [l.split(";")[1] for l in golden]


def test_golden_minimal_test3_with_the_registered_poisson_preconditioner():
    """minimal.test3's actual route: -snes_mf_operator makes [PETSc] precondition with the matrix the REGISTERED Jacobian
    callback fills, Poisson2DJacobianLocal (minimal.c:142-145, help text :9-10), under -pc_type mg rediscretised per level
    (= fish.c's PCMG).  Restated (poisson_jacobian=True): the golden's Newton counts 5, 3, 3 (:7, :12, :17), its error line
    (:18) and every monitor line up to three that carry the golden's own inexact 2-rank Chebyshev/SOR solves (<= 6e-8 in the area)."""
    fmt = lambda t: "area = %.8f; %.4f <= D <= %.4f" % t
    golden = ["area = 2.14201032; 0.2985 <= D <= 0.8826", "area = 1.60235989; 0.4166 <= D <= 0.9873",
              "area = 1.39125969; 0.5789 <= D <= 0.9362", "area = 1.32285324; 0.6106 <= D <= 0.9561",
              "area = 1.32217567; 0.6158 <= D <= 0.9510", "area = 1.32217567; 0.6158 <= D <= 0.9510",
              "area = 1.32217583; 0.6035 <= D <= 0.9518", "area = 1.33231935; 0.5757 <= D <= 0.9872",
              "area = 1.33230219; 0.5755 <= D <= 0.9872", "area = 1.33230217; 0.5755 <= D <= 0.9872",
              "area = 1.33230220; 0.5684 <= D <= 0.9872", "area = 1.33475595; 0.5153 <= D <= 0.9968",
              "area = 1.33475385; 0.5156 <= D <= 0.9968", "area = 1.33475385; 0.5156 <= D <= 0.9968"]
    got = []
    o = mo.minimal(grid_sequence=2, pc="mg", mf_operator=True, poisson_jacobian=True,
                   monitor=lambda st, it, u: got.append(fmt(mo.mse_monitor(u, -0.5, 2))))
    assert [s.its for s in o.stages] == [5, 3, 3]
    assert all(s.reason == "CONVERGED_FNORM_RELATIVE" for s in o.stages)
    assert "%.5e" % o.errinf == "6.79501e-04"
    same = [a == b for a, b in zip(got, golden)]
    assert len(got) == 14 and same.count(False) <= 3
    area = lambda l: float(l.split()[2].rstrip(";"))
    assert max(abs(area(a) - area(b)) for a, b in zip(got, golden)) <= 6e-8
    assert [l.split(";")[1] for l in got] == [l.split(";")[1] for l in golden]
