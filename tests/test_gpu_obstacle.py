"""c/ch12/obstacle.c on the device: the reduced-space active-set Newton method over the Poisson kernels,
through the C ABI, against oracle/obstacle_oracle.py (which reproduces c/ch12/output/obstacle.test1 completely) and the
goldens' own lines."""
import numpy as np
import pytest
import torch

from oracle import obstacle_oracle as oo
from p4pdes_b200.fish import Context
from p4pdes_b200.obstacle import obstacle_main

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


@pytest.fixture(scope="module")
def ctx():
    return Context()


def test_vi_kernels(ctx):
    rng = np.random.default_rng(3)
    n = 100003
    u, lo, F = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    u[::7] = lo[::7] + 5e-9                      # inside the 1e-8 activity tolerance
    du, dlo, dF, out = (ctx.from_host(a) for a in (u, lo, F, np.zeros(n)))
    ctx.vi_inactive_mask(du, dlo, dF, out)
    assert np.array_equal(ctx.to_host(out), np.where((u <= lo + 1e-8) & (F > 0), 0.0, 1.0))
    ctx.pointwise_mult(du, dF, out)
    assert np.array_equal(ctx.to_host(out), u * F)
    ctx.pointwise_max(du, dlo, out)
    assert np.array_equal(ctx.to_host(out), np.maximum(u, lo))


def test_golden_obstacle_test1_on_device(ctx):
    """c/ch12/makefile:17 with -pc_type none (the golden's ILU(0) is sequential): the monitored norms and the error line
    are the golden's; only the KSP count belongs to the preconditioner."""
    rep = obstacle_main("-da_refine 2 -snes_monitor_short -ksp_rtol 1.0e-12 -snes_rtol 1.0e-10 -pc_type none", ctx)
    assert rep.lines[:3] == ["  0 SNES Function norm 3.86571", "  1 SNES Function norm 1.3323", "  2 SNES Function norm < 1.e-11"]
    assert rep.lines[3].startswith("done on 9 x 9 grid ... CONVERGED_FNORM_RELATIVE, SNES iters = 2, last KSP iters = ")
    assert rep.lines[4] == "errors: av |u-uexact| = 3.076e-03, |u-uexact|_inf = 1.334e-02, active area error = 47.016%"


def test_golden_obstacle_test3_error_line_on_device(ctx):
    rep = obstacle_main("-snes_grid_sequence 3 -snes_converged_reason -pc_type jacobi", ctx)
    assert rep.lines[-1] == "errors: av |u-uexact| = 2.707e-03, |u-uexact|_inf = 1.428e-02, active area error = 18.430%"
    assert rep.lines[-2].startswith("done on 17 x 17 grid ... CONVERGED_FNORM_RELATIVE")


@pytest.mark.parametrize("refine,pc", [(4, "none"), (5, "jacobi"), (6, "none")])
def test_device_equals_the_oracle(ctx, refine, pc):
    m = 2 ** (refine + 1) + 1
    want = oo.rsls(m, pc="none", ksp_rtol=1e-10)          # Jacobi = a constant scaling here: same iterates
    got = obstacle_main("-da_refine %d -pc_type %s -ksp_rtol 1e-10" % (refine, pc), ctx, keep_solution=True)
    assert got.its == want.its and got.reason == want.reason
    assert all(abs(a - b) <= 1 for a, b in zip(got.ksp_its, want.ksp_its))
    np.testing.assert_allclose(got.fnorm, want.fnorm, rtol=1e-7, atol=1e-13)
    assert np.max(np.abs(got.u.cpu().numpy().reshape(m, m) - want.u)) < 1e-10
    assert abs(got.area_err - want.area_err) < 1e-12 and abs(got.errinf - want.errinf) < 1e-10


def test_grid_sequence_to_larger_grids(ctx):
    """From u = 0 the active set moves about one cell per Newton step, so fine grids are reached the reference's way, by
    -snes_grid_sequence (c/ch12/makefile:23): a few steps per grid, errors and the free boundary improve with h."""
    errs = []
    for seq in (5, 6, 7):                                 # 65^2, 129^2, 257^2 from the 3 x 3 DMDA
        rep = obstacle_main("-snes_grid_sequence %d -snes_converged_reason -pc_type jacobi" % seq, ctx)
        assert rep.reason.startswith("CONVERGED") and rep.its <= 6 and rep.m == 2 ** (seq + 1) + 1
        errs.append((rep.errinf, rep.area_err))
    assert errs[2][0] < errs[0][0] and errs[2][1] < errs[0][1]
