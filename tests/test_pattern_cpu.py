"""pattern.c implicit runs without a GPU: the solver oracle against the reference's golden (c/ch5/output/pattern.test2,
command c/ch5/makefile:52-53) and the control flow of the device driver p4pdes_b200/pattern.py on the NumPy stand-in."""
import os

import numpy as np
import pytest

from oracle import pattern_solver_oracle as po
from p4pdes_b200 import pattern as pp
from tests.fake_ops import FakeOps

GOLDEN_TEST2 = """running on 12 x 12 grid with square cells of side h = 0.208333 ...
0 TS dt 1. time 0.
      Linear solve converged due to CONVERGED_RTOL iterations 3
    Nonlinear solve converged due to CONVERGED_FNORM_RELATIVE iterations 1
1 TS dt 1. time 1.""".split("\n")
TEST2 = ("-da_refine 2 -ts_monitor -ts_dt 1 -ts_max_time 1 -ts_type beuler -pc_type mg -snes_converged_reason "
         "-ksp_converged_reason -snes_rtol 1.0e-1 -ptn_no_rhsjacobian")


def test_golden_file_is_what_we_pin():
    ref = "/root/reference/c/ch5/output/pattern.test2"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present")
    assert open(ref).read().rstrip("\n").split("\n") == GOLDEN_TEST2


def test_oracle_reproduces_pattern_test2():
    r = po.pattern_beuler(grid=3, refine=2, dt=1.0, tmax=1.0, rhsjac=False, snes_rtol=1e-1)
    assert r.lines == [GOLDEN_TEST2[0], GOLDEN_TEST2[1], GOLDEN_TEST2[4]]
    (t, dt, nr), = r.steps
    assert (t, dt, nr.its, nr.ksp_its, nr.reason) == (1.0, 1.0, 1, [3], "CONVERGED_FNORM_RELATIVE")


def test_driver_prints_pattern_test2_verbatim():
    r = pp.pattern_main(TEST2, FakeOps())
    assert r.lines == GOLDEN_TEST2


@pytest.mark.parametrize("argv,okw", [
    ("-da_grid_x 4 -da_grid_y 4 -da_refine 3 -ts_type beuler -ts_dt 5 -ts_max_time 12 -pc_type mg",
     dict(grid=4, refine=3, dt=5.0, tmax=12.0)),
    ("-da_refine 3 -ts_type beuler -ts_dt 2 -ts_max_time 4 -pc_type mg -ptn_no_rhsjacobian -snes_rtol 1e-6",
     dict(grid=3, refine=3, dt=2.0, tmax=4.0, rhsjac=False, snes_rtol=1e-6)),
    ("-da_refine 2 -ts_type beuler -ts_dt 5 -ts_max_time 5 -pc_type none", dict(grid=3, refine=2, dt=5.0, tmax=5.0, pc="none")),
    ("-da_grid_x 4 -da_grid_y 4 -da_refine 4 -ts_type beuler -ts_dt 5 -ts_max_time 5 -pc_type mg -p4b_mg_rscale 0.25",
     dict(grid=4, refine=4, dt=5.0, tmax=5.0, rscale=0.25)),
])
def test_driver_matches_oracle(argv, okw):
    r = pp.pattern_main(argv, FakeOps())
    o = po.pattern_beuler(**okw)
    assert [(t, dt) for t, dt, _ in r.steps] == [(t, dt) for t, dt, _ in o.steps]        # incl. the matched final step
    assert [s[2].its for s in r.steps] == [s[2].its for s in o.steps]
    assert [s[2].ksp_its for s in r.steps] == [s[2].ksp_its for s in o.steps]
    assert np.max(np.abs(r.Y.a.reshape(o.Y.shape) - o.Y)) <= 1e-12


@pytest.mark.parametrize("argv,msg", [
    ("-da_refine 2 -ts_type rk", "arkimex"),
    ("-ts_type beuler -pc_type ilu", "sequential"),
    ("-ts_type beuler -ptn_no_ijacobian", "not provided"),
    ("-ts_type beuler -da_grid_x 4 -da_grid_y 6", "requires mx == my"),   # pattern.c:89
])
def test_error_paths(argv, msg):
    with pytest.raises(ValueError, match=msg):
        pp.pattern_main(argv, FakeOps())


def test_periodic_interpolation_is_partition_of_unity_and_restriction_its_transpose():
    P = po.interpolation(6, 4)
    assert P.shape == (2 * 12 * 8, 2 * 6 * 4)
    np.testing.assert_allclose(P @ np.ones(P.shape[1]), 1.0)
    np.testing.assert_allclose(P.T @ np.ones(P.shape[0]), 4.0)            # every coarse node collects weight 4 in 2-D


def test_averaging_restriction_gives_mesh_independent_krylov_counts():
    """pattern.c's equations carry no cell-volume factor, so [PETSc]'s R = P^T over-weights the coarse correction 4x and
    the GMRES count grows with resolution; the averaging restriction (-p4b_mg_rscale 0.25) keeps it flat."""
    petsc = [max(po.pattern_beuler(grid=4, refine=r, dt=5.0, tmax=5.0).steps[0][2].ksp_its) for r in (5, 6)]
    avg = [max(po.pattern_beuler(grid=4, refine=r, dt=5.0, tmax=5.0, rscale=0.25).steps[0][2].ksp_its) for r in (5, 6)]
    assert avg[1] <= avg[0] <= 8 and petsc[1] >= 2 * avg[1]


# ---- ARKIMEX (pattern.c's default TS type): goldens c/ch5/output/pattern.test1, pattern.test4 -------------------------
GOLDEN_TEST1 = """running on 16 x 16 grid with square cells of side h = 0.156250 ...
0 TS dt 5. time 0.
1 TS dt 10.0587 time 5.
2 TS dt 9.52508 time 15.0587
3 TS dt 11.0962 time 24.5838
4 TS dt 12.5472 time 35.68
5 TS dt 15.0847 time 48.2272
6 TS dt 19.2147 time 63.3119
7 TS dt 27.8378 time 82.5266
8 TS dt 36.8439 time 110.364
9 TS dt 26.3958 time 147.208
10 TS dt 16.1864 time 167.627
11 TS dt 16.1864 time 183.814
12 TS dt 35.668 time 200.""".split("\n")
GOLDEN_TEST4 = """running on 24 x 24 grid with square cells of side h = 0.104167 ...
0 TS dt 5. time 0.
1 TS dt 7.32225 time 5.
2 TS dt 9.34935 time 12.3223
3 TS dt 12.4736 time 21.6716
4 TS dt 17.0057 time 34.1452
5 TS dt 25.5789 time 51.1509
6 TS dt 28.718 time 76.7298
7 TS dt 26.7656 time 105.448
8 TS dt 33.8933 time 132.213
9 TS dt 33.8933 time 166.107
10 TS dt 83.5115 time 200.
CALL-BACK REPORT
  solver type: arkimex
  IFunction:   1  | IJacobian:   1
  RHSFunction: 1  | RHSJacobian: 0""".split("\n")
TEST1 = "-da_grid_x 4 -da_grid_y 4 -da_refine 2 -ts_monitor"                  # c/ch5/makefile:50
TEST4 = "-da_refine 3 -ptn_call_back_report -ts_monitor"                      # c/ch5/makefile:59


def test_arkimex_golden_files_are_what_we_pin():
    for name, want in (("pattern.test1", GOLDEN_TEST1), ("pattern.test4", GOLDEN_TEST4)):
        ref = "/root/reference/c/ch5/output/" + name
        if not os.path.exists(ref):
            pytest.skip("reference tree not present")
        assert open(ref).read().rstrip("\n").split("\n") == want


def test_ark3_tableau_satisfies_its_order_conditions():
    AI, AE, b, bh, c = po.ARK3_AI, po.ARK3_AE, po.ARK3_B, po.ARK3_BH, po.ARK3_C
    np.testing.assert_allclose(AI.sum(1), c, atol=1e-15)
    np.testing.assert_allclose(AE.sum(1), c, atol=1e-15)
    for A in (AI, AE):                                     # third order, including the coupling conditions
        assert abs(b.sum() - 1) < 1e-15 and abs(b @ c - 0.5) < 1e-15 and abs(b @ c ** 2 - 1 / 3) < 1e-15
        assert abs(b @ A @ c - 1 / 6) < 1e-15
    assert abs(bh.sum() - 1) < 1e-15 and abs(bh @ c - 0.5) < 1e-15 and abs(bh @ c ** 2 - 1 / 3) > 1e-3    # embedded: order 2
    assert np.allclose(np.array(pp.ARK3_AI), AI) and np.allclose(np.array(pp.ARK3_AE), AE)
    assert np.allclose(pp.ARK3_BH, bh)


def test_oracle_reproduces_the_adaptive_arkimex_goldens_verbatim():
    r = po.pattern_arkimex(grid=4, refine=2)
    assert r.lines == GOLDEN_TEST1 and r.rejected == 1          # the step proposed at t = 147.208 is rejected once
    r = po.pattern_arkimex(grid=3, refine=3)
    assert r.lines == GOLDEN_TEST4[:12] and r.rejected == 0


def test_driver_prints_the_arkimex_goldens_verbatim():
    assert pp.pattern_main(TEST1, FakeOps()).lines == GOLDEN_TEST1
    assert pp.pattern_main(TEST4, FakeOps()).lines == GOLDEN_TEST4


def test_step_size_controller_rules():
    # accepted step: h * 0.9 * enorm^(-1/3), clipped to [0.1, 10]
    assert pp.adapt_basic(2.0, 0.001, True) == (True, 2.0 * 0.9 * 0.001 ** (-1 / 3))
    assert pp.adapt_basic(2.0, 1e-9, True) == (True, 20.0)
    # first rejection keeps the safety factor, the second consecutive one halves it
    assert pp.adapt_basic(2.0, 8.0, True) == (False, 2.0 * 0.9 * 0.5)
    assert pp.adapt_basic(2.0, 8.0, False) == (False, 2.0 * 0.45 * 0.5)
    # MATCHSTEP: overshoot -> land exactly; within 1 % -> stretch; less than two steps left -> two equal steps
    assert pp.match_step(190.0, 20.0, 200.0) == 10.0
    assert pp.match_step(190.0, 9.95, 200.0) == 10.0
    assert pp.match_step(180.0, 15.0, 200.0) == 10.0
    assert pp.match_step(100.0, 15.0, 200.0) == 15.0


GOLDEN_TEST3 = """running on 12 x 12 grid with square cells of side h = 0.208333 ...
0 TS dt 5. time 0.
    Nonlinear solve converged due to CONVERGED_FNORM_RELATIVE iterations 3
1 TS dt 5. time 5.
    Nonlinear solve converged due to CONVERGED_FNORM_RELATIVE iterations 3
2 TS dt 5. time 10.""".split("\n")
# c/ch5/makefile:56 runs `-ts_type cn -snes_fd_color` on 2 ranks (default PC); the device path uses the analytic
# Jacobian and multigrid: the printed lines (step sequence, 3 Newton iterations per step) are the same
TEST3 = "-da_refine 2 -ts_monitor -ts_type cn -ts_max_time 10 -snes_converged_reason -pc_type mg"


def test_crank_nicolson_golden_pattern_test3():
    ref = "/root/reference/c/ch5/output/pattern.test3"
    if os.path.exists(ref):
        assert open(ref).read().rstrip("\n").split("\n") == GOLDEN_TEST3
    for pc in ("ilu", "mg"):
        r = po.pattern_beuler(grid=3, refine=2, dt=5.0, tmax=10.0, theta=0.5, pc=pc)
        assert [s[2].its for s in r.steps] == [3, 3]
        assert r.lines == [GOLDEN_TEST3[0], GOLDEN_TEST3[1], GOLDEN_TEST3[3], GOLDEN_TEST3[5]]
    assert pp.pattern_main(TEST3, FakeOps()).lines == GOLDEN_TEST3


def test_crank_nicolson_is_second_order_and_backward_euler_first_order():
    # temporal convergence against a fine-step reference on a small grid
    ref = po.pattern_beuler(grid=3, refine=2, dt=0.125, tmax=8.0, theta=0.5, pc="ilu").Y
    err = {}
    for theta in (1.0, 0.5):
        err[theta] = [np.max(np.abs(po.pattern_beuler(grid=3, refine=2, dt=dt, tmax=8.0, theta=theta, pc="ilu").Y - ref))
                      for dt in (2.0, 1.0)]
    assert 1.7 < err[1.0][0] / err[1.0][1] < 2.3          # O(dt)
    assert 3.3 < err[0.5][0] / err[0.5][1] < 4.8          # O(dt^2)


def test_bdf_restart_step_golden_pattern_test5():
    """c/ch5/output/pattern.test5 (-da_refine 4 -ts_type bdf -ts_max_time 1): the two nonlinear solves of the restart
    (3 and 2 Newton iterations) and the step the controller proposes next (1.10972)."""
    (n1, n2), hnext, _ = po.pattern_bdf_first_step(grid=3, refine=4, dt=1.0)
    assert (n1, n2) == (3, 2)                               # pattern.test5:3-4
    assert po.fmt_g(float("%.6g" % hnext)) == "1.10972"     # pattern.test5:5  "1 TS dt 1.10972 time 1."


GOLDEN_TEST5 = """running on 48 x 48 grid with square cells of side h = 0.052083 ...
0 TS dt 1. time 0.
    Nonlinear solve converged due to CONVERGED_FNORM_RELATIVE iterations 3
    Nonlinear solve converged due to CONVERGED_FNORM_RELATIVE iterations 2
1 TS dt 1.10972 time 1.
CALL-BACK REPORT
  solver type: bdf
  IFunction:   1  | IJacobian:   1
  RHSFunction: 1  | RHSJacobian: 1""".split("\n")
TEST5 = "-da_refine 4 -ptn_call_back_report -ts_type bdf -ts_max_time 1 -snes_converged_reason -ts_monitor"   # c/ch5/makefile:62


def test_bdf_oracle_every_step():
    """The full BDF2 stepping of the oracle: pattern.test5's solver lines verbatim (the restart step is all the golden
    holds), and -- since nothing pins the later steps -- their order of accuracy with fixed steps: error ratios ~4."""
    ref = "/root/reference/c/ch5/output/pattern.test5"
    if os.path.exists(ref):
        assert open(ref).read().rstrip("\n").split("\n") == GOLDEN_TEST5
    r = po.pattern_bdf(grid=3, refine=4, dt=5.0, tmax=1.0)
    assert r.lines == GOLDEN_TEST5[:5] and r.newton_counts == [3, 2]
    fine = po.pattern_bdf(grid=3, refine=2, dt=0.125, tmax=8.0, adapt=False, snes_rtol=1e-12).Y
    err = [np.abs(po.pattern_bdf(grid=3, refine=2, dt=dt, tmax=8.0, adapt=False, snes_rtol=1e-12).Y - fine).max()
           for dt in (2.0, 1.0, 0.5)]
    assert 3.5 < err[0] / err[1] < 4.5 and 3.5 < err[1] / err[2] < 4.7
    # adaptive: rejections happen, the final time is matched, and the answer agrees with ARKIMEX to the controller's tolerance
    a = po.pattern_bdf(grid=4, refine=2, dt=5.0, tmax=200.0)
    b = po.pattern_arkimex(grid=4, refine=2, dt=5.0, tmax=200.0)
    assert a.rejected > 0 and a.steps[-1][0] == 200.0 and np.abs(a.Y - b.Y).max() < 5e-2


@pytest.mark.parametrize("pc", ["none", "mg"])
def test_driver_prints_pattern_test5_verbatim(pc):
    assert pp.pattern_main(TEST5 + " -pc_type " + pc, FakeOps()).lines == GOLDEN_TEST5


def test_driver_bdf_matches_the_oracle_on_an_adaptive_run():
    r = pp.pattern_main("-da_grid_x 4 -da_grid_y 4 -da_refine 2 -ts_type bdf -ts_monitor -pc_type none -ksp_rtol 1e-10", FakeOps())
    o = po.pattern_bdf(grid=4, refine=2, dt=5.0, tmax=200.0)
    assert [l for l in r.lines if " TS dt " in l] == [l for l in o.lines if " TS dt " in l]
    assert r.rejected == o.rejected == 5 and np.abs(r.Y.a.reshape(o.Y.shape) - o.Y).max() <= 1e-9
