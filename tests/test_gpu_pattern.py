"""GPU parity of the pattern.c implicit time-stepping path (BASELINE config 5) through the C ABI."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import minimal_pattern_oracle as mpo
from oracle import minimal_solver_oracle as mso
from oracle import pattern_solver_oracle as po
from p4pdes_b200 import pattern as pp
from p4pdes_b200.fish import Context
from tests.test_pattern_cpu import GOLDEN_TEST2, TEST2

pytestmark = pytest.mark.gpu
PAR = (2.5, 8.0e-5, 4.0e-5, 0.024, 0.06)


@pytest.fixture(scope="module")
def ctx():
    return Context()


def dev(ctx, a):
    return ctx.from_host(np.asarray(a, dtype=np.float64))


@pytest.mark.parametrize("m", [6, 12, 64, 130])
@pytest.mark.parametrize("rhsjac", [True, False])
def test_stage_jacobian_kernels(ctx, m, rhsjac):
    rng = np.random.default_rng(m)
    Y = mpo.pattern_initial_state(m, m) + 0.05 * rng.standard_normal((m, m, 2))
    shift = 0.37
    J = po.stage_jacobian(Y, shift, rhsjac)
    n = 2 * m * m
    X, b, pm1 = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    dY = dev(ctx, Y) if rhsjac else None
    dX, db, dp, out = dev(ctx, X), dev(ctx, b), dev(ctx, pm1), ctx.empty(n)
    ctx.pattern_jac_apply(m, *PAR, shift, dY, dX, out)
    np.testing.assert_allclose(ctx.to_host(out), J @ X, rtol=1e-13, atol=1e-13)
    ctx.pattern_jac_lin(m, *PAR, shift, dY, dX, db, dp, 0.3, 0.7, 0.45, True, out)
    want = 0.3 * pm1 + 0.7 * X + 0.45 * (b - J @ X) / J.diagonal()
    np.testing.assert_allclose(ctx.to_host(out), want, rtol=1e-13, atol=1e-13)
    ctx.pattern_jac_lin(m, *PAR, shift, dY, dX, db, None, 0.0, 0.0, 1.0, False, out)
    np.testing.assert_allclose(ctx.to_host(out), b - J @ X, rtol=1e-13, atol=1e-13)
    ctx.pattern_jac_lin(m, *PAR, shift, dY, dX, db, dp, 0.3, 0.7, 0.45, True, dp)         # in place over pm1
    np.testing.assert_allclose(ctx.to_host(dp), want, rtol=1e-13, atol=1e-13)
    lam = ctx.pattern_jac_gershgorin(m, *PAR, shift, dY, out)
    assert abs(lam - mso.gershgorin_jacobi(J)) <= 1e-12 * lam
    if m <= 12:
        dense = pp.dense_stage_jacobian(m, shift, Y if rhsjac else None, *PAR)
        np.testing.assert_allclose(dense, J.toarray(), rtol=1e-14, atol=1e-16)


@pytest.mark.parametrize("M", [3, 4, 16, 65])
def test_periodic_transfer_kernels(ctx, M):
    rng = np.random.default_rng(M)
    P = po.interpolation(M, M)
    rf, xc, xf = rng.standard_normal(P.shape[0]), rng.standard_normal(P.shape[1]), rng.standard_normal(P.shape[0])
    bc = ctx.empty(P.shape[1])
    ctx.pattern_restrict(M, M, dev(ctx, rf), bc)
    np.testing.assert_allclose(ctx.to_host(bc), P.T @ rf, rtol=1e-14, atol=1e-14)
    dxf = dev(ctx, xf)
    ctx.pattern_prolong_add(M, M, dev(ctx, xc), dxf)
    np.testing.assert_allclose(ctx.to_host(dxf), xf + P @ xc, rtol=1e-14, atol=1e-14)
    ctx.pattern_inject(M, M, dev(ctx, rf), bc)
    np.testing.assert_array_equal(ctx.to_host(bc).reshape(M, M, 2), rf.reshape(2 * M, 2 * M, 2)[::2, ::2, :])


def test_golden_pattern_test2_verbatim_on_device(ctx):
    assert pp.pattern_main(TEST2, ctx).lines == GOLDEN_TEST2                 # c/ch5/output/pattern.test2


@pytest.mark.parametrize("argv,okw", [
    ("-da_grid_x 4 -da_grid_y 4 -da_refine 4 -ts_type beuler -ts_dt 5 -ts_max_time 12 -pc_type mg",
     dict(grid=4, refine=4, dt=5.0, tmax=12.0)),
    ("-da_refine 4 -ts_type beuler -ts_dt 2 -ts_max_time 4 -pc_type mg -ptn_no_rhsjacobian -snes_rtol 1e-6",
     dict(grid=3, refine=4, dt=2.0, tmax=4.0, rhsjac=False, snes_rtol=1e-6)),
    ("-da_grid_x 4 -da_grid_y 4 -da_refine 5 -ts_type beuler -ts_dt 5 -ts_max_time 10 -pc_type mg -p4b_mg_rscale 0.25",
     dict(grid=4, refine=5, dt=5.0, tmax=10.0, rscale=0.25)),
])
def test_device_time_stepping_matches_oracle(ctx, argv, okw):
    r = pp.pattern_main(argv, ctx)
    o = po.pattern_beuler(**okw)
    assert [(t, dt) for t, dt, _ in r.steps] == [(t, dt) for t, dt, _ in o.steps]
    assert [s[2].its for s in r.steps] == [s[2].its for s in o.steps]
    for a, b in zip(r.steps, o.steps):
        assert all(abs(x - y) <= 1 for x, y in zip(a[2].ksp_its, b[2].ksp_its))
        np.testing.assert_allclose(a[2].fnorms[0], b[2].fnorms[0], rtol=1e-9)
    Y = ctx.to_host(r.Y).reshape(o.Y.shape)
    assert np.max(np.abs(Y - o.Y)) <= 1e-9                                   # Newton tolerance 1e-8 on ||R||


def test_config5_at_full_size(ctx):
    """-da_grid_x 4 -da_grid_y 4 -da_refine 9 (2048 x 2048 x 2 = 8.4 M unknowns, 10 levels),
    backward Euler + Newton-GMRES-MG.  Size-independent properties: every stage solve converges with bounded Krylov
    counts; mass-like invariants stay in range (0 <= v, u <= 1); the pattern has started to grow from the seeded patch."""
    # -p4b_mg_rscale 0.25: averaging restriction.  With PETSc's R = P^T the pointwise-scaled equations of pattern.c get a
    # 4x over-weighted coarse correction and GMRES needs hundreds of iterations at this resolution (pattern.py).
    r = pp.pattern_main("-da_grid_x 4 -da_grid_y 4 -da_refine 9 -ts_type beuler -ts_dt 5 -ts_max_time 10 -pc_type mg "
                        "-p4b_mg_rscale 0.25", ctx)
    assert r.m == 2048 and len(r.steps) == 2
    assert all(s[2].reason.startswith("CONVERGED") for s in r.steps)
    assert max(max(s[2].ksp_its) for s in r.steps) <= 12
    Y = ctx.to_host(r.Y).reshape(2048, 2048, 2)
    assert Y[..., 0].max() <= 1.0 + 1e-9 and Y[..., 1].min() >= -1e-9 and Y[..., 1].max() > 0.1
    print("pattern 2048^2 x 2: %.3f s for 2 steps, Newton its %s, KSP its %s"
          % (r.seconds, [s[2].its for s in r.steps], [s[2].ksp_its for s in r.steps]))
