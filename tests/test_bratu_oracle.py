"""oracle/bratu_oracle.py against the reference's golden c/ch7/solns/output/bratu2D.test1 (c/ch7/solns/makefile:12):
what the golden pins -- the residual callback (||F(u0)|| = 9.04754 on the 9 x 9 grid with Liouville's boundary data) and
the discretisation (converged error 3.169e-04) -- and what it cannot pin without PETSc's SNESFAS source (the cycle:
same number of outer iterations, norms of the same size; parity of the cycle UNPINNED, see the oracle's header)."""
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
