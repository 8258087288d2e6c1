"""minimal.c on the device through p4b_minimal_solve (host logic in C++ inside the library): the device instantiation of the
C++ host logic that tests/test_native_nk_cpu.py checks on the CPU against the Python oracle.
First run on a B200 in round 2 (profiles/r02_pending.md) and promoted to the `gpu` marker."""
import numpy as np
import pytest
import torch

from p4pdes_b200 import minimal as pm
from p4pdes_b200.fish import Context

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


@pytest.fixture(scope="module")
def ctx():
    return Context()


@pytest.mark.parametrize("argv", [
    "-snes_fd_color -snes_converged_reason -snes_monitor_short -ksp_converged_reason -snes_grid_sequence 3 -ms_problem tent "
    "-pc_type mg",
    "-snes_fd_color -da_refine 4 -pc_type mg -ksp_type cg -ms_q 0.0 -ms_problem tent -snes_converged_reason",
    "-snes_fd_color -da_grid_x 17 -da_grid_y 17 -snes_grid_sequence 2 -pc_type mg -ms_problem tent",
])
def test_native_solve_equals_the_python_host(ctx, argv):
    """p4b_minimal_solve (host logic in C++ inside the library, csrc/nk_solver.hpp) against the Python host running the
    same algorithm through the individual C-ABI calls: same lines, same counts, same solution.  (The two differ only in
    the base-grid inverse: banded LU in C++, LAPACK in Python.)"""
    a = pm.minimal_main(argv, ctx)
    b = pm.minimal_main(argv, ctx, native=True)
    assert (a.mx, a.my) == (b.mx, b.my)
    assert [s.its for s in a.stages] == [s.its for s in b.stages]
    assert [s.ksp_its for s in a.stages] == [s.ksp_its for s in b.stages]
    assert [s.reason for s in a.stages] == [s.reason for s in b.stages]
    for s, t in zip(a.stages, b.stages):
        np.testing.assert_allclose(s.fnorms, t.fnorms, rtol=1e-2, atol=1e-10 * s.fnorms[0])
    assert [l.split(" norm ")[0] for l in a.lines] == [l.split(" norm ")[0] for l in b.lines]
    ua, ub = ctx.to_host(a.u), ctx.to_host(b.u)
    assert np.max(np.abs(ua - ub)) <= 1e-9 * max(1.0, np.max(np.abs(ua)))


def test_native_solve_cluster_configuration(ctx):
    r = pm.minimal_main("-da_grid_x 33 -da_grid_y 33 -snes_grid_sequence 6 -snes_fd_color -pc_type mg", ctx, native=True)
    assert (r.mx, r.my) == (2049, 2049) and all(s.reason.startswith("CONVERGED") for s in r.stages)
    assert max(max(s.ksp_its) for s in r.stages[1:]) <= 12 and r.errinf < 1e-7
    print("minimal 2049^2 native: %.3f s, Newton its %s" % (r.seconds, [s.its for s in r.stages]))


def test_callback_contract_on_device(ctx):
    """p4b_snes2d_solve: the residual is a HOST callback (here the NumPy restatement of minimal.c's FormFunctionLocal plays
    the user's callback), the algebra runs on the device.  Must agree with the run whose residual is the device kernel."""
    import ctypes as C
    from oracle import minimal_pattern_oracle as mpo
    from p4pdes_b200 import lib as L
    calls = [0]
    gcache = {}

    def residual(_user, mx, my, u_ptr, F_ptr):
        calls[0] += 1
        if (mx, my) not in gcache:
            gcache[(mx, my)] = mpo.minimal_g(mx, my, "tent", 1.0, 1.1)
        u = np.ctypeslib.as_array(u_ptr, shape=(my, mx))
        np.ctypeslib.as_array(F_ptr, shape=(my, mx))[:] = mpo.minimal_function(u, gcache[(mx, my)], -0.5)
        return 0

    o = L.MinimalOpts()
    L.check(ctx.lib.p4b_minimal_default_opts(C.byref(o)))
    o.grid_x = o.grid_y = 3
    o.grid_sequence = 3
    g0 = mpo.minimal_g(3, 3, "tent", 1.0, 1.1)
    u0 = np.zeros((3, 3))
    u0[[0, -1], :] = g0[[0, -1], :]
    u0[:, [0, -1]] = g0[:, [0, -1]]
    out = np.zeros(17 * 17)
    res = L.MinimalResult()
    cb, line = L.RESIDUAL2D_FN(residual), L.LINE_FN(lambda s, c: None)
    ref = pm.minimal_main("-snes_fd_color -snes_grid_sequence 3 -ms_problem tent -pc_type mg", ctx)
    # route 0: recognition off, every evaluation is the host callback (nine per level Jacobian)
    L.check(ctx.lib.p4b_tune(b"recognise_residual", 0))
    try:
        L.check(ctx.lib.p4b_snes2d_solve(ctx.h, C.byref(o), cb, None, u0.ctypes.data_as(C.c_void_p), line, None,
                                         out.ctypes.data_as(C.c_void_p), out.size, C.byref(res)))
    finally:
        L.check(ctx.lib.p4b_tune(b"recognise_residual", 1))
    assert ctx.lib.p4b_snes2d_last_route() == 0
    assert (res.mx, res.my, res.nstages) == (17, 17, 4)
    assert [res.stage[s].its for s in range(4)] == [s.its for s in ref.stages]
    assert calls[0] > 9 * sum(s.its for s in ref.stages)
    assert np.max(np.abs(out - ctx.to_host(ref.u))) <= 1e-8
    # route 1 (default): the callback is recognised as the library's kernel after 2 probes per grid + 1 (grids 3, 5, 9, 17)
    calls[0] = 0
    out2 = np.zeros(17 * 17)
    L.check(ctx.lib.p4b_snes2d_solve(ctx.h, C.byref(o), cb, None, u0.ctypes.data_as(C.c_void_p), line, None,
                                     out2.ctypes.data_as(C.c_void_p), out2.size, C.byref(res)))
    assert ctx.lib.p4b_snes2d_last_route() == 1 and calls[0] == 2 * 4 + 1 + 4      # probes + one re-verification per stage
    assert [res.stage[s].its for s in range(4)] == [s.its for s in ref.stages]
    assert np.max(np.abs(out2 - ctx.to_host(ref.u))) <= 1e-8
