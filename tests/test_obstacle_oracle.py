"""oracle/obstacle_oracle.py ([PETSc] SNESVINEWTONRSLS restated) against the reference's goldens c/ch12/output/obstacle.test1-4
(c/ch12/makefile:16-27), and the product's host logic (p4pdes_b200/obstacle.py) over the NumPy stand-in of the device
operations against the same lines."""
import numpy as np

from oracle import obstacle_oracle as oo
from p4pdes_b200.obstacle import obstacle_main
from tests.fake_ops import FakeOps

TEST1 = ["  0 SNES Function norm 3.86571", "  1 SNES Function norm 1.3323", "  2 SNES Function norm < 1.e-11",
         "done on 9 x 9 grid ... CONVERGED_FNORM_RELATIVE, SNES iters = 2, last KSP iters = 11",
         "errors: av |u-uexact| = 3.076e-03, |u-uexact|_inf = 1.334e-02, active area error = 47.016%"]
TEST3_ERRORS = "errors: av |u-uexact| = 2.707e-03, |u-uexact|_inf = 1.428e-02, active area error = 18.430%"


def short(x):
    return "%g" % x if x > 1e-9 else ("%5.3e" % x if x > 1e-11 else "< 1.e-11")


def test_oracle_reproduces_obstacle_test1_completely():
    # -da_refine 2 -snes_monitor_short -ksp_rtol 1.0e-12 -snes_rtol 1.0e-10 -ksp_converged_reason, default PC = ILU(0), CG
    r = oo.rsls(9, snes_rtol=1e-10, ksp_rtol=1e-12, pc="ilu")
    assert ["  %d SNES Function norm %s" % (i, short(v)) for i, v in enumerate(r.fnorm)] == TEST1[:3]
    assert r.ksp_its == [11, 11] and r.its == 2 and r.reason == "CONVERGED_FNORM_RELATIVE"          # both "iterations 11" lines
    assert ("errors: av |u-uexact| = %.3e, |u-uexact|_inf = %.3e, active area error = %.3f%%"
            % (r.err1, r.errinf, 100 * r.area_err)) == TEST1[4]


def test_oracle_reproduces_obstacle_test2_completely():
    """c/ch12/makefile:20 on 4 ranks: -da_refine 2 -snes_monitor_short -ksp_type gmres -pc_type asm -sub_pc_type lu.  The
    third and fourth norm carry the inexactness of the GMRES(rtol 1e-5) + additive-Schwarz solves: they pin PCASM's defaults
    (restricted, overlap 1, one block per rank of the 2 x 2 DMDA) -- every other reading misses them."""
    golden = ["  0 SNES Function norm 3.86571", "  1 SNES Function norm 1.3323", "  2 SNES Function norm 4.84465e-06",
              "  3 SNES Function norm 1.511e-11",
              "done on 9 x 9 grid ... CONVERGED_FNORM_RELATIVE, SNES iters = 3, last KSP iters = 4",
              "errors: av |u-uexact| = 3.076e-03, |u-uexact|_inf = 1.334e-02, active area error = 47.016%"]
    r = oo.rsls(9, pc="asm", ranks=(2, 2))
    got = ["  %d SNES Function norm %s" % (i, short(v)) for i, v in enumerate(r.fnorm)]
    got.append("done on 9 x 9 grid ... %s, SNES iters = %d, last KSP iters = %d" % (r.reason, r.its, r.ksp_its[-1]))
    got.append("errors: av |u-uexact| = %.3e, |u-uexact|_inf = %.3e, active area error = %.3f%%" % (r.err1, r.errinf, 100 * r.area_err))
    assert got == golden
    for kw in (dict(ranks=(1, 4)), dict(ranks=(4, 1)), dict(asm_overlap=0), dict(asm_overlap=2), dict(asm_restricted=False)):
        v = oo.rsls(9, pc="asm", **kw)
        assert short(v.fnorm[2]) != "4.84465e-06" and v.its == 3          # same Newton path, other linear-solve errors


def test_oracle_reproduces_obstacle_test3():
    """c/ch12/makefile:23: -snes_grid_sequence 3 -snes_converged_reason -pc_type mg.  Everything the golden prints: the
    Newton counts of the four grids, the last KSP count, the error line.  The counts need the INEXACT multigrid solves (rtol
    1e-5): exact solves converge one iteration earlier on three of the four grids."""
    st = oo.rsls_grid_sequence(3, pc="mg")
    assert [r.m for r in st] == [3, 5, 9, 17] and all(r.reason == "CONVERGED_FNORM_RELATIVE" for r in st[1:])
    assert [r.its for r in st] == [1, 2, 2, 3] and st[-1].ksp_its[-1] == 4
    r = st[-1]
    assert ("errors: av |u-uexact| = %.3e, |u-uexact|_inf = %.3e, active area error = %.3f%%"
            % (r.err1, r.errinf, 100 * r.area_err)) == TEST3_ERRORS
    assert [r.its for r in oo.rsls_grid_sequence(3, pc="exact")] == [1, 1, 1, 2]
    assert [r.its for r in oo.rsls_grid_sequence(3, pc="mg", mg_smoother="jacobi")] == [1, 1, 2, 3]


def test_oracle_reproduces_obstacle_test4_semismooth():
    """c/ch12/makefile:28: -snes_grid_sequence 2 -snes_converged_reason -snes_type vinewtonssls (KSPCG + ILU(0) as obstacle.c
    leaves them).  Everything the golden prints: Newton counts 4, 6, 5, last KSP count 6, the error line (the same discrete
    solution as the reduced-space method's)."""
    st = oo.ssls_grid_sequence(2)
    assert [r.m for r in st] == [3, 5, 9] and all(r.reason == "CONVERGED_FNORM_RELATIVE" for r in st)
    assert [r.its for r in st] == [4, 6, 5] and st[-1].ksp_its[-1] == 6
    r = st[-1]
    assert ("errors: av |u-uexact| = %.3e, |u-uexact|_inf = %.3e, active area error = %.3f%%"
            % (r.err1, r.errinf, 100 * r.area_err)) == TEST1[4]
    assert [r.its for r in oo.ssls_grid_sequence(2, project=False)] == [5, 6, 5]       # the initial projection is needed
    # Fischer-Burmeister: zero exactly on the complementarity set, both evaluation branches agree
    a, b = np.array([0.0, 2.0, 0.0, 1e-9, 3.0]), np.array([1.5, 0.0, 0.0, 1e-9, -4.0])
    np.testing.assert_allclose(oo.fischer(a, b), np.sqrt(a * a + b * b) - (a + b), atol=1e-15)
    assert np.all(oo.fischer(a[:3], b[:3]) == 0.0)


def test_oracle_error_lines_of_the_other_goldens():
    # obstacle.test2 (GMRES + ASM/LU on 4 ranks) and test4 (vinewtonssls) end on the same discrete solution as test1
    r = oo.rsls(9, pc="exact")
    assert ("%.3e %.3e %.3f" % (r.err1, r.errinf, 100 * r.area_err)) == "3.076e-03 1.334e-02 47.016"
    # obstacle.test3: grid sequence to 17 x 17; the converged state does not depend on the path
    r = oo.rsls(17, pc="exact", snes_rtol=1e-10)
    assert ("errors: av |u-uexact| = %.3e, |u-uexact|_inf = %.3e, active area error = %.3f%%"
            % (r.err1, r.errinf, 100 * r.area_err)) == TEST3_ERRORS


def test_host_logic_prints_the_goldens_over_the_stand_in():
    rep = obstacle_main("-da_refine 2 -snes_monitor_short -ksp_rtol 1.0e-12 -snes_rtol 1.0e-10 -pc_type none", FakeOps())
    assert rep.lines[:3] == TEST1[:3] and rep.lines[-1] == TEST1[4]
    assert rep.lines[3].startswith("done on 9 x 9 grid ... CONVERGED_FNORM_RELATIVE, SNES iters = 2, last KSP iters = ")
    rep = obstacle_main("-snes_grid_sequence 3 -snes_converged_reason -pc_type jacobi", FakeOps())
    assert rep.lines[-1] == TEST3_ERRORS and rep.lines[-2].startswith("done on 17 x 17 grid ... CONVERGED_FNORM_RELATIVE")
    assert [l.count("Nonlinear solve converged") for l in rep.lines[:4]] == [1, 1, 1, 1]
    assert [len(l) - len(l.lstrip()) for l in rep.lines[:4]] == [8, 6, 4, 2]               # PETSc's grid-sequence indentation
    # same states as the oracle with the same (unpreconditioned CG) solves
    want = oo.rsls(33, pc="none", ksp_rtol=1e-10)
    got = obstacle_main("-da_refine 4 -pc_type none -ksp_rtol 1e-10", FakeOps(), keep_solution=True)
    assert got.its == want.its and got.ksp_its == want.ksp_its
    np.testing.assert_allclose(got.fnorm, want.fnorm, rtol=1e-9, atol=1e-14)
    assert np.max(np.abs(got.u.a.reshape(33, 33) - want.u)) < 1e-12
