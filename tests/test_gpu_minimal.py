"""GPU parity of the minimal.c Newton-Krylov-multigrid path (BASELINE config 4) through the C ABI:
kernels against the oracle on identical inputs, whole solves against the oracle and the reference's goldens."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import minimal_pattern_oracle as mpo
from oracle import minimal_solver_oracle as mo
from p4pdes_b200 import lib as L
from p4pdes_b200 import minimal as pm
from p4pdes_b200.fish import Context

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return Context()


def dev(ctx, a):
    return ctx.from_host(np.asarray(a, dtype=np.float64))


@pytest.mark.parametrize("mx,my,problem,q", [(9, 7, "catenoid", -0.5), (33, 17, "tent", -0.5), (65, 65, "catenoid", -0.25),
                                             (130, 37, "tent", 0.0)])
def test_fd_jacobian_and_stencil9_kernels(ctx, mx, my, problem, q):
    rng = np.random.default_rng(5)
    g = mpo.minimal_g(mx, my, problem, 1.0, 1.1)
    # the differencing step is h = sqrt(eps) sqrt(1 + ||u||_2) for every column ("wp"); the rounding noise of F divided
    # by it is what separates two implementations of the same finite-difference formula
    u = 1.0 + 0.3 * rng.random((my, mx))
    F = lambda w: mpo.minimal_function(w, g, q)
    du, dg, dF = dev(ctx, u), dev(ctx, g), ctx.empty(mx * my)
    ctx.minimal_function(mx, my, q, du, dg, dF)
    np.testing.assert_allclose(ctx.to_host(dF).reshape(my, mx), F(u), rtol=1e-12, atol=1e-13)
    vals = ctx.empty(9 * mx * my)
    ctx.minimal_jacobian_fd(mx, my, q, du, dg, dF, vals)
    rp, ci, d = pm.stencil9_to_csr(ctx.to_host(vals), mx, my)
    J = sp.csr_matrix((d, ci, rp), shape=(mx * my, mx * my))
    Jo = mo.fd_jacobian(F, u)
    # same differencing, same colours: the two differ by the rounding of F (1e-16 relative) divided by h ~ 1e-7
    assert abs(J - Jo).max() <= 2e-6 * abs(Jo).max()
    # structure: identity boundary rows, no coupling of interior rows to boundary columns
    bd = np.ones((my, mx), bool)
    bd[1:-1, 1:-1] = False
    b = bd.ravel()
    Jd = J.toarray()
    assert np.allclose(Jd[b][:, b], np.eye(b.sum()), atol=1e-7) and np.all(Jd[~b][:, b] == 0.0)
    # y = A x and the fused smoother step on the device matrix
    x, bb, pm1 = rng.standard_normal(mx * my), rng.standard_normal(mx * my), rng.standard_normal(mx * my)
    dx_, db, dp, dy = dev(ctx, x), dev(ctx, bb), dev(ctx, pm1), ctx.empty(mx * my)
    ctx.stencil9_apply(mx, my, vals, dx_, dy)
    np.testing.assert_allclose(ctx.to_host(dy), J @ x, rtol=1e-13, atol=1e-13)
    ctx.stencil9_lin(mx, my, vals, dx_, db, dp, 0.3, 0.7, 0.45, True, dy)
    want = 0.3 * pm1 + 0.7 * x + 0.45 * (bb - J @ x) / J.diagonal()
    np.testing.assert_allclose(ctx.to_host(dy), want, rtol=1e-13, atol=1e-13)
    ctx.stencil9_lin(mx, my, vals, dx_, db, None, 0.0, 0.0, 1.0, False, dy)
    np.testing.assert_allclose(ctx.to_host(dy), bb - J @ x, rtol=1e-13, atol=1e-13)
    ctx.stencil9_lin(mx, my, vals, dx_, db, dp, 0.3, 0.7, 0.45, True, dp)          # in place over pm1
    np.testing.assert_allclose(ctx.to_host(dp), want, rtol=1e-13, atol=1e-13)
    lam = ctx.stencil9_gershgorin(mx, my, vals, dy)
    assert abs(lam - mo.gershgorin_jacobi(J)) <= 1e-12 * lam
    # the column-indexed SELL-32 copy of the same matrix ([PETSc] AIJ/SELL Mat) multiplies identically
    A = C.c_void_p()
    L.check(ctx.lib.p4b_sell_create(ctx.h, mx * my, rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p),
                                    d.ctypes.data_as(C.c_void_p), C.byref(A)))
    ys = ctx.empty(mx * my)
    L.check(ctx.lib.p4b_sell_spmv(A, dx_.data_ptr(), ys.data_ptr()))
    ctx.stencil9_apply(mx, my, vals, dx_, dy)
    np.testing.assert_allclose(ctx.to_host(ys), ctx.to_host(dy), rtol=1e-14, atol=1e-14)
    ctx.lib.p4b_sell_destroy(A)


def test_small_helpers(ctx):
    rng = np.random.default_rng(6)
    cmx, cmy = 9, 5
    uf = rng.standard_normal((2 * cmy - 1, 2 * cmx - 1))
    duc = ctx.empty(cmx * cmy)
    ctx.inject2d(cmx, cmy, dev(ctx, uf), duc)
    np.testing.assert_array_equal(ctx.to_host(duc).reshape(cmy, cmx), uf[::2, ::2])
    n = 77
    Ainv, b = rng.standard_normal((n, n)), rng.standard_normal(n)
    dx = ctx.empty(n)
    ctx.dense_matvec(n, dev(ctx, Ainv), dev(ctx, b), dx)
    np.testing.assert_allclose(ctx.to_host(dx), Ainv @ b, rtol=1e-13, atol=1e-13)
    x, y = rng.standard_normal(1000), rng.standard_normal(1000)
    out = ctx.empty(1000)
    ctx.axpby(0.5, dev(ctx, x), -2.0, dev(ctx, y), out)
    np.testing.assert_allclose(ctx.to_host(out), 0.5 * x - 2.0 * y, rtol=1e-15)
    ctx.axpby(3.0, dev(ctx, x), 0.0, None, out)
    np.testing.assert_allclose(ctx.to_host(out), 3.0 * x, rtol=1e-15)
    ctx.copy(dev(ctx, y), out)
    np.testing.assert_array_equal(ctx.to_host(out), y)


def test_golden_minimal_test1_lines_on_device(ctx):
    r = pm.minimal_main("-snes_fd_color -snes_converged_reason -snes_monitor_short -ms_problem catenoid "
                        "-ms_catenoid_c 2.0 -da_refine 1", ctx)
    assert r.lines[0] == "  0 SNES Function norm 1.08276"                                       # minimal.test1:1
    assert r.lines[-1] == "done on 5 x 5 grid and problem catenoid:  error |u-uexact|_inf = 1.10603e-04"   # :8
    assert abs(r.stages[0].its - 5) <= 1 and r.stages[0].reason == "CONVERGED_FNORM_RELATIVE"


@pytest.mark.parametrize("argv,okw", [
    ("-snes_fd_color -snes_grid_sequence 3 -ms_problem tent -pc_type mg", dict(grid_sequence=3, problem="tent", pc="mg")),
    ("-snes_fd_color -da_refine 4 -pc_type mg -ksp_type cg -ms_q 0.0 -ms_problem tent",
     dict(refine=4, problem="tent", q=0.0, pc="mg", ksp="cg")),
    ("-snes_fd_color -da_grid_x 5 -da_grid_y 9 -snes_grid_sequence 3 -pc_type mg -ms_catenoid_c 1.5",
     dict(mx=5, my=9, grid_sequence=3, pc="mg", catenoid_c=1.5)),
    ("-snes_fd_color -da_refine 5 -pc_type mg -pc_mg_levels 4", dict(refine=5, pc="mg", mg_levels=4)),
])
def test_device_solve_matches_oracle(ctx, argv, okw):
    r = pm.minimal_main(argv, ctx)
    o = mo.minimal(**okw)
    assert (r.mx, r.my) == (o.mx, o.my)
    # device and NumPy residuals differ in rounding (rsqrt vs power), the linear solves stop at rtol 1e-5: counts within +-1
    for a, b in zip(r.stages, o.stages):
        assert a.reason == b.reason == "CONVERGED_FNORM_RELATIVE"
        # (a long globalisation phase -- a cold start on a fine grid takes ~10 damped steps -- amplifies the difference)
        assert abs(a.its - b.its) <= max(1, b.its // 6)
        assert abs(max(a.ksp_its) - max(b.ksp_its)) <= 1
        assert a.fnorms[-1] <= 1e-8 * a.fnorms[0]
    u = ctx.to_host(r.u).reshape(o.u.shape)
    # both converge to the discrete solution to the Newton tolerance (||F|| <= 1e-8 ||F0||)
    assert np.max(np.abs(u - o.u)) <= 1e-7 * max(1.0, np.max(np.abs(o.u)))
    if o.errinf is not None:
        assert abs(r.errinf - o.errinf) <= 1e-7


def test_cluster_configuration_at_full_size(ctx):
    """c/ch8/cluster.sh:70 / BASELINE config 4: 33 x 33 base grid, -snes_grid_sequence 6 -> 2049 x 2049, Newton-GMRES-MG.
    Size-independent properties: every stage converges, Krylov iterations per Newton step stay bounded (multigrid),
    the error against the exact catenoid decays like h^2 from stage to stage."""
    r = pm.minimal_main("-da_grid_x 33 -da_grid_y 33 -snes_grid_sequence 6 -snes_fd_color -pc_type mg", ctx)
    assert (r.mx, r.my) == (2049, 2049)
    assert all(s.reason.startswith("CONVERGED") for s in r.stages)
    assert max(max(s.ksp_its) for s in r.stages[1:]) <= 12
    assert r.stages[-1].its <= 4                                  # grid sequencing: a good initial iterate
    e = {}
    for seq in (2, 3):
        e[seq] = pm.minimal_main("-da_grid_x 33 -da_grid_y 33 -snes_grid_sequence %d -snes_fd_color -pc_type mg" % seq,
                                 ctx).errinf
    assert 3.0 < e[2] / e[3] < 5.0                                # O(h^2)
    assert r.errinf < e[3] / 30.0
    print("minimal 2049^2: %.3f s, Newton its %s, max KSP its %s, error %.3e"
          % (r.seconds, [s.its for s in r.stages], [max(s.ksp_its) for s in r.stages], r.errinf))
