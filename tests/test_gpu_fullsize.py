"""BASELINE-size runs (257^3, 513^3) checked through size-independent properties -- the oracle cannot be run
at these sizes in test time: h-independent iteration counts, O(h^2) error decay, true-residual reduction,
symmetry of A and of the preconditioner, linearity, reproducibility."""
import numpy as np
import pytest
import torch

from p4pdes_b200 import lib as L
from p4pdes_b200.fish import Context, Multigrid, mg_options

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return Context()


def solve(ctx, refine, levels, rtol=1e-10):
    g = L.refined_grid(3, refine)
    mg = Multigrid(ctx, g, mg_options(levels=levels))
    n = mg.nlocal
    b, x, u0, ue = ctx.empty(n), ctx.empty(n), ctx.empty(n), ctx.empty(n)
    mg.fish_setup("manuexp", True, b=b, u0=u0, uexact=ue)
    res = mg.cg_solve(b, x, rtol=rtol)
    return g, mg, b, x, u0, ue, res


def test_iteration_counts_are_h_independent_and_errors_decay_like_h2(ctx):
    errs, its = {}, {}
    for refine in (5, 6, 7, 8):          # 65^3 .. 513^3, coarse grid 9^3 throughout (-pc_mg_levels refine-1)
        g, mg, b, x, u0, ue, res = solve(ctx, refine, refine - 1)
        assert res.reason == L.CONVERGED_RTOL
        ctx.axpy(-1.0, x, u0)
        ctx.axpy(-1.0, ue, u0)
        errs[refine] = ctx.norminf(u0)
        its[refine] = res.its
        # true residual of the solve: ||b - A x|| / ||b||
        r = ctx.empty(mg.nlocal)
        ctx.stencil_residual(g, b, x, r)
        assert ctx.norm2(r) / ctx.norm2(b) < 1e-9
        mg.close()
        del b, x, u0, ue, r
        torch.cuda.empty_cache()
    # 12-13 iterations at rtol 1e-10, independent of h
    assert all(11 <= v <= 14 for v in its.values()), its
    assert max(its.values()) - min(its.values()) <= 2
    for r in (6, 7, 8):                  # second-order discretisation: error quarters per refinement
        assert 3.5 < errs[r - 1] / errs[r] < 4.5, errs


def test_operator_and_preconditioner_are_symmetric_at_257(ctx):
    g = L.refined_grid(3, 7)
    mg = Multigrid(ctx, g, mg_options(levels=6))
    n = mg.nlocal
    gen = torch.Generator(device="cuda").manual_seed(1)
    u = torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)
    v = torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)
    Au, Av, Mu, Mv = (ctx.empty(n) for _ in range(4))
    ctx.stencil_apply(g, u, Au)
    ctx.stencil_apply(g, v, Av)
    assert abs(ctx.dot(v, Au) - ctx.dot(u, Av)) <= 1e-12 * abs(ctx.dot(v, Au))          # A = A^T (fish.test8 "symmetric")
    mg.apply(u, Mu)
    mg.apply(v, Mv)
    assert abs(ctx.dot(v, Mu) - ctx.dot(u, Mv)) <= 1e-10 * abs(ctx.dot(v, Mu))          # M^-1 symmetric: CG is applicable
    assert ctx.dot(u, Mu) > 0 and ctx.dot(v, Mv) > 0                                    # and positive
    # linearity of the cycle
    w = 2.0 * u - 3.0 * v
    Mw = ctx.empty(n)
    mg.apply(w, Mw)
    ref = 2.0 * Mu - 3.0 * Mv
    assert float((Mw - ref).norm() / ref.norm()) < 1e-12
    mg.close()


def test_solve_is_bitwise_reproducible_at_257(ctx):
    g, mg, b, x, u0, ue, res = solve(ctx, 7, 6)
    x1 = x.clone()
    res2 = mg.cg_solve(b, x, rtol=1e-10)
    assert res.history == res2.history and torch.equal(x, x1)      # fixed-order reductions, no fp atomics
    mg.close()


def _direct_parity(ctx, refine, levels):
    """The same solve by the C oracle (oracle/fish_cpu.c, pinned by tests/test_fish_cpu_oracle.py) and on the device:
    the north-star bar -- equal KSP iteration count, preconditioned residual history within 1e-10 relative, solution
    within 1e-12 relative (fp64)."""
    from oracle import fish_cpu as fc
    want = fc.solve(dim=3, refine=refine, levels=levels, rtol=1e-10, want_arrays=True)
    g, mg, b, x, u0, ue, res = solve(ctx, refine, levels)
    assert res.reason == L.CONVERGED_RTOL and res.its == want["its"], (res.its, want["its"])
    np.testing.assert_allclose(res.history, want["history"], rtol=1e-10, atol=4e-16 * want["history"][0])
    bw = torch.from_numpy(want["b"]).cuda()
    assert float(torch.linalg.vector_norm(b - bw) / torch.linalg.vector_norm(bw)) < 1e-13      # F(u0): the right-hand side
    del bw
    ctx.axpy(-1.0, x, u0)                       # u = u0 - y (SNESSolve_KSPONLY)
    uw = torch.from_numpy(want["u"]).cuda()
    rel = float(torch.linalg.vector_norm(u0 - uw) / torch.linalg.vector_norm(uw))
    assert rel < 1e-12, rel
    mg.close()
    return res.its, rel


def test_direct_oracle_parity_at_257(ctx):
    """BASELINE config 2: fish 3-D 257^3, -da_refine 7 -pc_mg_levels 6 (c/ch8/cluster.sh:63 with Chebyshev/Jacobi)."""
    its, rel = _direct_parity(ctx, 7, 6)
    print("257^3: %d its, solution rel. diff vs C oracle %.2e" % (its, rel))


@pytest.mark.slow
def test_direct_oracle_parity_at_513(ctx):
    """BASELINE config 3's grid on one GPU: 513^3, -da_refine 8 -pc_mg_levels 7 (the bench.py workload)."""
    its, rel = _direct_parity(ctx, 8, 7)
    print("513^3: %d its, solution rel. diff vs C oracle %.2e" % (its, rel))
