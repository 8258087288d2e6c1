"""oracle/bratu_oracle.py against the reference's golden c/ch7/solns/output/bratu2D.test1 (c/ch7/solns/makefile:12):
the residual callback (||F(u0)|| = 9.04754 on the 9 x 9 grid with Liouville's boundary data), the discretisation
(converged error 3.169e-04) and -- with cycle="petsc", the restatement of [PETSc] SNESFASCycle_Full found by matching this
golden -- every printed digit of its residual norms and its count of NGS calls.  The textbook F cycle the device runs has
the same components and convergence class but not the same iterates."""
import numpy as np
import pytest

from oracle import bratu_oracle as bo

GOLDEN = {"norms": [9.04754, 0.000449564, 1.87245e-06, 8.93257e-09], "its": 3, "errinf": "3.169e-04", "m": 9}


@pytest.mark.parametrize("order", ["lexicographic", "redblack"])
def test_golden_bratu2d_test1(order):
    r = bo.fas_solve(refine=2, lam=1.0, exact=True, rtol=1.0e-8, order=order)
    assert r.m == GOLDEN["m"]
    assert "%g" % float("%.6g" % r.fnorm[0]) == "9.04754"                 # pure callback KAT
    assert "%.3e" % r.errinf == GOLDEN["errinf"]                           # cycle-independent
    assert r.its == GOLDEN["its"]                                          # F cycle per outer iteration, as the golden's 3
    for mine, theirs in zip(r.fnorm[1:], GOLDEN["norms"][1:]):            # same size (PETSc's cycle differs in detail)
        assert 0.2 < mine / theirs < 5.0


def test_golden_bratu2d_test1_verbatim_with_petscs_full_cycle():
    """c/ch7/solns/output/bratu2D.test1:1-6.  The cycle is pinned by the numbers: each ingredient switched off misses them."""
    r = bo.fas_solve(refine=2, cycle="petsc")
    assert r.its == 3 and r.ngs_calls == 58                                # :5 "iterations 3", :6 "NGS calls = 58"
    assert ["%g" % float("%.6g" % x) for x in r.fnorm] == ["9.04754", "0.000449564", "1.87245e-06", "8.93257e-09"]      # :1-4
    assert "%.3e" % r.errinf == GOLDEN["errinf"]                           # :7
    for variant, second in (("no_downsweep", "3.75084e-06"), ("no_presmooth", "0.000107976"), ("no_final_v", "0.00392143")):
        v = bo.fas_solve(refine=2, cycle="petsc", variant=variant)
        assert "%g" % float("%.6g" % v.fnorm[2]) == second != "1.87245e-06"
    # the reference's sweep sets a boundary node when the loop reaches it (bratu2D.c:250-256): with the edges set
    # beforehand (the red-black ordering does that) the first cycle already differs
    rb = bo.fas_solve(refine=2, cycle="petsc", order="redblack")
    assert "%g" % float("%.6g" % rb.fnorm[1]) != "0.000449564" and rb.its == 3
    # the device's textbook F cycle = this cycle without the smoothing on the way down: same first iterate
    tb = bo.fas_solve(refine=2)
    assert "%g" % float("%.6g" % tb.fnorm[1]) == "0.000449564" and tb.ngs_calls == 54


def test_transfer_operators_are_the_dmda_q1_pair():
    from oracle import fish_oracle as fo
    g = fo.refined_grid(2, 2)                                              # 5 x 5 coarse of 9 x 9
    P = fo.interpolation(fo.refined_grid(2, 1))
    rng = np.random.default_rng(0)
    xc, r = rng.standard_normal((5, 5)), rng.standard_normal((9, 9))
    np.testing.assert_allclose(bo.prolong(xc).ravel(), P @ xc.ravel(), atol=1e-14)
    np.testing.assert_allclose(bo.restrict(r).ravel(), P.T @ r.ravel(), atol=1e-14)


def test_error_decays_like_h2_and_cycles_are_h_independent():
    errs, its = [], []
    for refine in (3, 4, 5):
        r = bo.fas_solve(refine=refine, order="redblack", rtol=1.0e-10)
        errs.append(r.errinf)
        its.append(r.its)
    assert 3.5 < errs[0] / errs[1] < 4.5 and 3.5 < errs[1] / errs[2] < 4.5
    assert max(its) <= 5
