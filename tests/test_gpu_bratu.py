"""c/ch7/solns/bratu2D.c on the device: kernels and the FAS + NGS solve through the C ABI against
oracle/bratu_oracle.py (red-black ordering on both sides), the golden's pinned numbers, and the reference's headline
command line (bratu2D.c:15) at a size a test can afford."""
import numpy as np
import pytest
import torch

from oracle import bratu_oracle as bo
from p4pdes_b200 import lib as L
from p4pdes_b200.bratu import bratu_main
from p4pdes_b200.fish import Context

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]
FAS = ("-snes_type fas -snes_fas_type full -fas_levels_snes_type ngs -fas_levels_snes_ngs_sweeps 2 -fas_levels_snes_max_it 1 "
       "-fas_coarse_snes_type ngs -fas_coarse_snes_ngs_sweeps 2 -fas_coarse_snes_max_it 4")


@pytest.fixture(scope="module")
def ctx():
    return Context()


@pytest.mark.parametrize("m,exact,lam", [(9, 1, 1.0), (33, 1, 1.0), (65, 0, 3.0), (129, 0, 6.0)])
def test_residual_and_ngs_kernels(ctx, m, exact, lam):
    rng = np.random.default_rng(m)
    u = 0.3 * rng.standard_normal((m, m))
    b = 0.1 * rng.standard_normal((m, m))
    g = bo.boundary_values(m, bool(exact))
    du, db, dF = (torch.from_numpy(a.ravel().copy()).cuda() for a in (u, b, u))
    L.check(ctx.lib.p4b_bratu_function(ctx.h, m, m, lam, exact, du.data_ptr(), db.data_ptr(), dF.data_ptr()))
    want = bo.residual(u, lam, g, b)
    assert np.max(np.abs(dF.cpu().numpy().reshape(m, m) - want)) <= 1e-13 * max(1.0, np.abs(want).max())
    L.check(ctx.lib.p4b_bratu_function(ctx.h, m, m, lam, exact, du.data_ptr(), None, dF.data_ptr()))
    assert np.max(np.abs(dF.cpu().numpy().reshape(m, m) - bo.residual(u, lam, g))) <= 1e-13 * max(1.0, np.abs(want).max())
    L.check(ctx.lib.p4b_bratu_ngs(ctx.h, m, m, lam, exact, 2, db.data_ptr(), du.data_ptr()))
    want_u = bo.ngs(u, b, lam, g, 2, order="redblack")
    assert np.max(np.abs(du.cpu().numpy().reshape(m, m) - want_u)) <= 1e-12
    gd = ctx.empty(m * m)
    L.check(ctx.lib.p4b_bratu_exact(ctx.h, m, m, exact, gd.data_ptr()))
    assert np.max(np.abs(gd.cpu().numpy().reshape(m, m) - g)) <= 1e-14


def test_golden_bratu2d_test1_on_device(ctx):
    """c/ch7/solns/makefile:12.  Pinned: the first norm and the final error; the cycle is the oracle's (red-black)."""
    rep = bratu_main("-lb_exact -snes_converged_reason -lb_showcounts " + FAS + " -da_refine 2 -snes_monitor_short", ctx)
    assert rep.lines[0] == "  0 SNES Function norm 9.04754 "
    assert rep.lines[-1] == "done on 9 x 9 grid:   error |u-uexact|_inf = 3.169e-04"
    assert "Nonlinear solve converged due to CONVERGED_FNORM_RELATIVE iterations 3" in rep.lines
    want = bo.fas_solve(refine=2, order="redblack")
    np.testing.assert_allclose(rep.fnorm, want.fnorm, rtol=1e-7, atol=1e-14 * want.fnorm[0])
    assert (rep.residual_calls, rep.ngs_calls) == (want.residual_calls + 0, want.ngs_calls)


@pytest.mark.parametrize("argv,kw", [
    ("-lb_exact -snes_rtol 1.0e-10 " + FAS + " -da_refine 5", dict(refine=5, rtol=1e-10)),
    ("-lb_lambda 5.0 -snes_type fas -snes_fas_type multiplicative -fas_levels_snes_type ngs -fas_coarse_snes_type ngs "
     "-fas_coarse_snes_max_it 4 -fas_levels_snes_ngs_sweeps 2 -fas_coarse_snes_ngs_sweeps 2 -da_refine 4 -snes_fas_levels 3",
     dict(refine=4, lam=5.0, exact=False, levels=3, full_cycle=False)),                    # V cycles, 17^2 coarse grid, no exact solution
])
def test_solve_equals_the_oracle(ctx, argv, kw):
    rep = bratu_main(argv, ctx, keep_solution=True)
    want = bo.fas_solve(order="redblack", **kw)
    assert rep.its == want.its and rep.reason > 0
    np.testing.assert_allclose(rep.fnorm, want.fnorm, rtol=1e-6, atol=1e-14 * want.fnorm[0])
    m = want.m
    assert np.max(np.abs(rep.u.cpu().numpy().reshape(m, m) - want.u)) <= 1e-10
    if want.errinf is not None:
        assert abs(rep.errinf - want.errinf) <= 1e-12


def test_headline_command_line_scaled_down(ctx):
    """bratu2D.c:15 is -da_grid_x 6 -da_grid_y 6 ... -da_refine 12 (20481^2, 31.5 s on 20 cores); here -da_refine 8
    (1281^2): converges in one full cycle to discretisation accuracy, like the reference's run."""
    rep = bratu_main("-da_grid_x 6 -da_grid_y 6 -lb_exact -snes_rtol 1.0e-10 -snes_converged_reason -lb_showcounts " + FAS +
                     " -da_refine 8", ctx)
    assert (rep.mx, rep.my) == (1281, 1281) and rep.reason > 0 and rep.its <= 3
    assert rep.errinf < 1e-6
    print("bratu2D 1281^2: %d cycles, %.1f ms, error %.3e, %d residual / %d NGS calls" %
          (rep.its, rep.solve_ms, rep.errinf, rep.residual_calls, rep.ngs_calls))
