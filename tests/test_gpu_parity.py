"""GPU parity: every kernel and the whole CG+GMG solve, through the C ABI, against the CPU oracle.

Tolerances are the north star's: iteration count +-1 (we assert equality where the oracle is exact),
residual norms 1e-10 relative, solution 1e-12 relative (fp64).  Sizes are those the oracle finishes
in seconds; BASELINE-size runs are checked through size-independent properties (test_gpu_fullsize.py).
"""
import numpy as np
import pytest
import torch

from oracle import fish_oracle as fo
from p4pdes_b200 import lib as L
from p4pdes_b200.fish import Context, Multigrid, fish_main, mg_options

pytestmark = pytest.mark.gpu

GRIDS = [
    (1, (17,), (1, 1, 1), (1.0, 1.0, 1.0)),
    (1, (129,), (2.0, 1, 1), (3.0, 1.0, 1.0)),
    (2, (9, 9), (1, 1, 1), (1.0, 1.0, 1.0)),
    (2, (33, 17), (1.0, 2.0, 1), (1.0, 2.5, 1.0)),
    (2, (129, 129), (1, 1, 1), (1.0, 1.0, 1.0)),
    (3, (9, 9, 9), (1, 1, 1), (0.01, 2.0, 100.0)),
    (3, (17, 9, 33), (1.0, 0.5, 2.0), (1.0, 1.0, 1.0)),
    (3, (33, 33, 33), (1, 1, 1), (1.0, 1.0, 1.0)),
    (3, (65, 65, 65), (1, 1, 1), (1.0, 1.0, 1.0)),
    (2, (2049, 33), (1, 1, 1), (1.0, 1.0, 1.0)),          # wide 2-D rows: the plane-marching kernel's 2-D path
]


@pytest.fixture(scope="module")
def ctx():
    return Context()


@pytest.fixture(params=["generic", "march"])
def path(request):
    """Run a test once on the one-thread-per-node kernels and once with the TMA-staged plane-marching
    kernel forced on for every 3-D level it can take (normally only planes >= 128^2 use it)."""
    L.tune("march_enabled", 1)
    L.tune("march_min_plane", 1 if request.param == "march" else 1 << 30)
    yield request.param
    L.tune("march_min_plane", 16384)


def ogrid(dim, m, Ls):
    mm = tuple(m) + (1,) * (3 - len(m))
    return fo.Grid(dim, mm, tuple(float(x) for x in Ls))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).ravel()).cuda()


def relerr(a, b):
    a = a.cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("dim,m,Ls,c", GRIDS)
def test_stencil_apply_and_residual(ctx, path, dim, m, Ls, c):
    og = ogrid(dim, m, Ls)
    g = L.make_grid(dim, m, Ls, c)
    A = fo.jacobian(og, c)
    rng = np.random.default_rng(1)
    u = rng.standard_normal(og.n)
    b = rng.standard_normal(og.n)
    du, db, dy = dev(u), dev(b), ctx.empty(og.n)
    ctx.stencil_apply(g, du, dy)
    assert relerr(dy, A @ u) < 1e-14
    ctx.stencil_residual(g, db, du, dy)
    assert relerr(dy, b - A @ u) < 1e-14


# grids whose coarse rows are >= 64 wide take the marching restriction / 4-node prolongation kernels; ragged
# row strips (11 = 2*4 + 3 coarse rows), fewer coarse planes than one chunk, and a chunked z range
TRANSFER_GRIDS = GRIDS + [
    (3, (129, 21, 13), (1, 1, 1), (1.0, 1.0, 1.0)),
    (3, (131, 9, 37), (1.0, 0.5, 2.0), (1.0, 2.0, 0.5)),
    (3, (257, 17, 9), (1, 1, 1), (1.0, 1.0, 1.0)),
]


@pytest.mark.parametrize("dim,m,Ls,c", TRANSFER_GRIDS)
def test_transfer_matches_q1_interpolation(ctx, dim, m, Ls, c):
    og = ogrid(dim, m, Ls)
    if any((og.m[d] - 1) % 2 or og.m[d] <= 3 for d in range(dim)):
        pytest.skip("not coarsenable")
    oc = og.coarsen()
    g = L.make_grid(dim, m, Ls, c)
    P = fo.interpolation(oc)
    rng = np.random.default_rng(2)
    r = rng.standard_normal(og.n)
    xc = rng.standard_normal(oc.n)
    xf = rng.standard_normal(og.n)
    dbc = ctx.empty(oc.n)
    ctx.restrict(g, dev(r), dbc)
    assert relerr(dbc, P.T @ r) < 1e-14
    dxf = dev(xf)
    ctx.prolong_add(g, dev(xc), dxf)
    assert relerr(dxf, xf + P @ xc) < 1e-14
    # the same through vectors that start on an odd element (8-byte, not 16-byte, aligned)
    r1 = torch.zeros(og.n + 1, dtype=torch.float64, device="cuda")
    r1[1:] = dev(r)
    b1 = torch.zeros(oc.n + 1, dtype=torch.float64, device="cuda")
    ctx.restrict(g, r1[1:], b1[1:])
    assert relerr(b1[1:], P.T @ r) < 1e-14
    assert torch.equal(b1[1:], dbc)                       # and bit-identical to the aligned call
    # fused residual + restriction
    A = fo.jacobian(og, c)
    ctx.residual_restrict(g, dev(r), dev(xf), dbc)
    assert relerr(dbc, P.T @ (r - A @ xf)) < 1e-13


@pytest.mark.parametrize("dim,m,Ls,c", GRIDS[2:8])
@pytest.mark.parametrize("its,zero", [(1, True), (2, True), (2, False), (3, False), (4, True)])
def test_chebyshev_jacobi_smoother(ctx, path, dim, m, Ls, c, its, zero):
    og = ogrid(dim, m, Ls)
    g = L.make_grid(dim, m, Ls, c)
    A = fo.jacobian(og, c)
    lam = fo.lambda_max_jacobi(og, c)
    emin, emax = 0.1 * lam, 1.1 * lam
    rng = np.random.default_rng(3)
    b = rng.standard_normal(og.n)
    x0 = np.zeros(og.n) if zero else rng.standard_normal(og.n)
    want = fo.chebyshev_smooth(A, fo.JacobiPC(A), b, x0, emin, emax, its)
    dx = dev(x0)
    ctx.cheb_jacobi(g, emin, emax, its, zero, dev(b), dx, ctx.empty(og.n))
    assert relerr(dx, want) < 1e-13


@pytest.mark.parametrize("dim,refine,problem,c", [
    (1, 3, "manupoly", (1.0, 1.0, 1.0)), (1, 5, "manuexp", (1.0, 1.0, 1.0)),
    (2, 3, "manuexp", (1.0, 1.0, 1.0)), (2, 4, "manupoly", (1.0, 3.0, 1.0)),
    (3, 2, "manupoly", (0.01, 2.0, 100.0)), (3, 4, "manuexp", (1.0, 1.0, 1.0)), (3, 3, "zero", (1.0, 1.0, 1.0)),
])
@pytest.mark.parametrize("gonb", [True, False])
def test_form_function_and_initial_state(ctx, dim, refine, problem, c, gonb):
    og = fo.refined_grid(dim, refine)
    g = L.refined_grid(dim, refine, c=c)
    n = og.n
    f, gb, u, F = ctx.empty(n), ctx.empty(n), ctx.empty(n), ctx.empty(n)
    ctx.fish_sample(g, problem, f, gb)
    ctx.initial_state(g, gb, gonb, u)
    u0 = fo.initial_state(og, problem, gonb)
    assert relerr(u, u0) < 1e-15 if np.linalg.norm(u0) > 0 else float(u.abs().max()) == 0.0
    ctx.poisson_function(g, u, f, gb, F)
    assert relerr(F, fo.form_function(og, u0, problem, c)) < 1e-13
    # and at a random state (exercises the g-for-boundary-neighbour substitution)
    ur = np.random.default_rng(4).standard_normal(n)
    ctx.poisson_function(g, dev(ur), f, gb, F)
    assert relerr(F, fo.form_function(og, ur, problem, c)) < 1e-13


def test_fish_test1_residual_norm_kat(ctx, goldens):
    # pure-callback KAT: "0 SNES Function norm 0.925546" (c/ch6/output/fish.test1:1)
    rep = fish_main("-fsh_dim 1 -fsh_problem manupoly -da_refine 3 -pc_type mg -ksp_rtol 1.0e-12 "
                    "-snes_monitor_short -ksp_converged_reason", ctx)
    g = goldens["fish.test1"]
    assert rep.lines[0] == "  0 SNES Function norm %s" % g["snes_fnorm0"]
    assert rep.lines[2] == "  1 SNES Function norm < 1.e-11"
    assert rep.lines[3] == "problem manupoly on %s grid:" % g["gridstr"]
    assert rep.lines[4] == "  error |u-uexact|_inf = %s, |u-uexact|_h = %s" % (g["errinf"], g["err2h"])


@pytest.mark.parametrize("name,opts", [
    ("fish.test2", "-fsh_dim 1 -fsh_problem manupoly -da_refine 1"),
    ("fish.test6", "-fsh_dim 3 -da_refine 2 -fsh_problem manupoly -fsh_cx 0.01 -fsh_cy 2 -fsh_cz 100"),
    ("fish.test7", "-fsh_dim 3 -fsh_problem manupoly -da_refine 2"),
])
def test_reference_golden_error_norms(ctx, goldens, name, opts):
    # converged error norms are solver independent: the device path must print the reference's digits
    g = goldens[name]
    rep = fish_main(opts + " -pc_type mg -ksp_rtol 1e-12", ctx)
    assert rep.lines[-2] == "problem %s on %s grid:" % (g["problem"], g["gridstr"])
    assert rep.lines[-1] == "  error |u-uexact|_inf = %s, |u-uexact|_h = %s" % (g["errinf"], g["err2h"])


MG_CASES = [
    (2, 3, dict()),
    (2, 5, dict(cycle="w")),
    (2, 6, dict(levels=4)),
    (3, 3, dict()),
    (3, 4, dict(cycle="w", smoother_ksp="richardson", smoother_its=1)),
    (3, 5, dict(levels=4, smoother_its=3)),
    (3, 4, dict(eig=(0.2, 2.2))),
    (1, 6, dict()),
]


@pytest.mark.parametrize("dim,refine,kw", MG_CASES)
@pytest.mark.parametrize("fuse", [True, False])
def test_pcmg_apply(ctx, path, dim, refine, kw, fuse):
    og = fo.refined_grid(dim, refine)
    g = L.refined_grid(dim, refine)
    M = fo.PCMG(og, opts=fo.MGOptions(**kw))
    mg = Multigrid(ctx, g, mg_options(levels=kw.get("levels", 0), cycle=kw.get("cycle", "v"),
                                      smoother=kw.get("smoother_ksp", "chebyshev"),
                                      smooth_its=kw.get("smoother_its", 2), eig=kw.get("eig"), fuse=fuse))
    assert mg.nlevels == M.nlev
    for l in range(M.nlev):
        mm, eig = mg.level_info(l)
        assert mm == M.grids[l].m
        if l > 0:
            assert abs(eig[0] - M.eig[l][0]) < 1e-14 and abs(eig[1] - M.eig[l][1]) < 1e-14
    r = np.random.default_rng(5).standard_normal(og.n)
    z = ctx.empty(og.n)
    mg.apply(dev(r), z)
    assert relerr(z, M.apply(r)) < 1e-12
    mg.close()


SOLVE_CASES = [
    ("-fsh_dim 2 -da_refine 3", dict(dim=2, refine=3)),
    ("-fsh_dim 2 -da_refine 4", dict(dim=2, refine=4)),
    ("-fsh_dim 2 -da_refine 6 -ksp_rtol 1e-10", dict(dim=2, refine=6, rtol=1e-10)),            # BASELINE config C1
    ("-fsh_dim 2 -da_refine 6", dict(dim=2, refine=6)),
    ("-fsh_dim 3 -da_refine 3", dict(dim=3, refine=3)),
    ("-fsh_dim 3 -da_refine 4 -ksp_rtol 1e-10", dict(dim=3, refine=4, rtol=1e-10)),
    ("-fsh_dim 3 -da_refine 5 -ksp_rtol 1e-10", dict(dim=3, refine=5, rtol=1e-10)),
    ("-fsh_dim 3 -da_refine 5 -ksp_rtol 1e-10 -pc_mg_levels 4", dict(dim=3, refine=5, rtol=1e-10, mg=dict(levels=4))),
    ("-fsh_dim 3 -da_refine 3 -fsh_problem manupoly -fsh_cx 0.5 -fsh_cy 2 -fsh_cz 3 -ksp_rtol 1e-8",
     dict(dim=3, refine=3, problem="manupoly", c=(0.5, 2.0, 3.0), rtol=1e-8)),
    ("-fsh_dim 2 -da_refine 4 -pc_mg_cycle_type w -mg_levels_ksp_type richardson -mg_levels_ksp_max_it 1",
     dict(dim=2, refine=4, mg=dict(cycle="w", smoother_ksp="richardson", smoother_its=1))),
    ("-fsh_dim 3 -da_refine 4 -mg_levels_ksp_chebyshev_eigenvalues 0.2,2.2 -ksp_rtol 1e-10",
     dict(dim=3, refine=4, rtol=1e-10, mg=dict(eig=(0.2, 2.2)))),
    ("-fsh_dim 2 -da_refine 3 -fsh_initial_gonboundary false", dict(dim=2, refine=3, gonboundary=False)),
    ("-fsh_dim 1 -da_refine 6 -ksp_rtol 1e-10", dict(dim=1, refine=6, rtol=1e-10)),
]


@pytest.mark.parametrize("opts,okw", SOLVE_CASES)
@pytest.mark.parametrize("fuse", [True, False])
def test_fish_solve_matches_oracle(ctx, path, opts, okw, fuse):
    okw = dict(okw)
    mgkw = okw.pop("mg", {})
    want = fo.fish(mg=fo.MGOptions(**mgkw), **okw)
    rep = fish_main(opts + " -pc_type mg -ksp_converged_reason -ksp_monitor" + ("" if fuse else " -p4b_no_fuse"),
                    ctx, keep_solution=True)
    # same KSP iteration count (north star: +-1; the oracle is deterministic so we demand equality)
    assert rep.ksp.its == want.its
    assert rep.ksp.reason == L.CONVERGED_RTOL
    # residual history ||M^-1 r_i|| within 1e-10 relative
    hist = np.array(rep.ksp.history)
    np.testing.assert_allclose(hist, np.array(want.history), rtol=1e-10)
    # solution within 1e-12 relative
    assert relerr(rep.u, want.u) < 1e-12
    assert abs(rep.fnorm0 - want.fnorm0) <= 1e-13 * want.fnorm0
    assert "%.3e" % rep.errinf == "%.3e" % want.errinf
    assert "%.3e" % rep.err2h == "%.3e" % want.err2h
    assert ("    Linear solve converged due to CONVERGED_RTOL iterations %d" % want.its) in rep.lines


def test_cuda_graph_replay_of_coarse_levels():
    # on a capturable (non-default) stream the sub-cycle below the finest level is replayed as a CUDA graph;
    # results must be the same as the kernel-by-kernel launch sequence (and as the oracle)
    want = fo.fish(dim=3, refine=5, rtol=1e-10, mg=fo.MGOptions(levels=4))
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        c2 = Context()
        g = L.refined_grid(3, 5)
        outs = []
        for use_graph in (True, False):
            mg = Multigrid(c2, g, mg_options(levels=4, use_graph=use_graph))
            b, x = c2.empty(mg.nlocal), c2.empty(mg.nlocal)
            mg.fish_setup("manuexp", True, b=b)
            n0 = c2.lib.p4b_launch_count()
            for _ in range(3):                       # replays, not just the capture pass
                res = mg.cg_solve(b, x, rtol=1e-10)
            outs.append((res.its, res.history, x.clone(), c2.lib.p4b_launch_count() - n0))
            mg.close()
        side.synchronize()
    assert outs[0][0] == outs[1][0] == want.its
    np.testing.assert_allclose(outs[0][1], want.history, rtol=1e-10)
    assert torch.equal(outs[0][2], outs[1][2])       # same kernels, same order: bitwise equal
    assert outs[0][3] == outs[1][3]                  # the launch counter accounts for replayed kernels
    assert relerr(outs[0][2], want.y) < 1e-12


def test_cg_with_jacobi_and_no_pc(ctx):
    og = fo.refined_grid(2, 4)
    g = L.refined_grid(2, 4)
    A = fo.jacobian(og)
    b = np.random.default_rng(6).standard_normal(og.n)
    mg = Multigrid(ctx, g)
    for pc, M in (("none", lambda r: r.copy()), ("jacobi", fo.JacobiPC(A).apply)):
        xw, its, hist = fo.cg(A, b, M, rtol=1e-8)
        x = ctx.empty(og.n)
        res = mg.cg_solve(dev(b), x, rtol=1e-8, pc=pc)
        assert abs(res.its - its) <= 1
        assert relerr(x, xw) < 1e-7
        np.testing.assert_allclose(res.history[:10], hist[:10], rtol=1e-9)
    mg.close()


def test_host_buffer_entry_points(ctx):
    # the e2e path bench.py times: host buffers in, host buffers out
    want = fo.fish(dim=3, refine=4, rtol=1e-10)
    og = want.grid
    g = L.refined_grid(3, 4)
    mg = Multigrid(ctx, g)
    bh = torch.from_numpy(want.b.copy()).pin_memory()
    xh = torch.empty(og.n, dtype=torch.float64).pin_memory()
    res = mg.cg_solve_host(bh, xh, rtol=1e-10)
    assert res.its == want.its
    assert relerr(xh, want.y) < 1e-12
    # SNESKSPONLY on host buffers
    x, y, z = og.coords()
    fh = torch.from_numpy((fo.f_rhs(3, "manuexp", x, y, z, (1, 1, 1)) * np.ones(og.shape)).ravel().copy()).pin_memory()
    gh = torch.from_numpy((fo.u_exact(3, "manuexp", x, y, z) * np.ones(og.shape)).ravel().copy()).pin_memory()
    uh = torch.from_numpy(fo.initial_state(og, "manuexp").ravel().copy()).pin_memory()
    res = mg.fish_solve_host(fh, gh, uh, rtol=1e-10)
    assert res.its == want.its
    assert relerr(uh, want.u) < 1e-12
    mg.close()


def test_vector_kernels(ctx):
    rng = np.random.default_rng(7)
    for n in (1, 31, 1000, 1 << 20, (1 << 22) + 3):
        x, y = rng.standard_normal(n), rng.standard_normal(n)
        dx, dy = dev(x), dev(y)
        assert abs(ctx.dot(dx, dy) - x @ y) <= 1e-12 * np.sqrt(n) * max(1.0, abs(x @ y))
        assert abs(ctx.norm2(dx) - np.linalg.norm(x)) <= 1e-13 * np.linalg.norm(x)
        assert ctx.norminf(dx) == np.abs(x).max()
        ctx.axpy(0.75, dx, dy)
        assert relerr(dy, y + 0.75 * x) < 1e-15
        ctx.aypx(-0.5, dx, dy)
        assert relerr(dy, x - 0.5 * (y + 0.75 * x)) < 1e-15
    # reductions are fixed-order: bitwise reproducible
    dx = dev(rng.standard_normal(1 << 22))
    assert ctx.dot(dx, dx) == ctx.dot(dx, dx)


def test_errors_are_loud(ctx):
    with pytest.raises(L.P4BError, match="levels"):
        Multigrid(ctx, L.refined_grid(2, 2), mg_options(levels=9))
    with pytest.raises(L.P4BError):
        fish_main("-fsh_dim 4 -pc_type mg", ctx)
    # a coarsest grid too large for the dense inverse: refused when the multigrid cycle asks for it (a one-level
    # hierarchy of any size is what -pc_type none / jacobi run on)
    mg = Multigrid(ctx, L.refined_grid(3, 5), mg_options(levels=2))
    n = mg.nlocal
    b, x = ctx.empty(n), ctx.empty(n)
    b.fill_(1.0)
    with pytest.raises(L.P4BError, match="coarsest"):
        mg.cg_solve(b, x, rtol=1e-5)
    res = mg.cg_solve(b, x, rtol=1e-5, pc="jacobi")
    assert res.reason == L.CONVERGED_RTOL
    mg.close()
