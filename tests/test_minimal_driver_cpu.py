"""Control flow of the device driver p4pdes_b200/minimal.py (option parsing, Newton + bt line search, GMRES / CG,
multigrid cycle on assembled level Jacobians, grid sequencing), exercised WITHOUT a GPU by handing it the NumPy
stand-in tests/fake_ops.py instead of the device Context, and compared with the independent oracle.  The GPU tests
(tests/test_gpu_minimal.py) run the very same driver on the device."""
import numpy as np
import pytest

from oracle import minimal_solver_oracle as mo
from p4pdes_b200 import minimal as pm
from tests.fake_ops import FakeOps


def test_reference_command_lines_parse():
    o = pm.parse_options("-da_grid_x 33 -da_grid_y 33 -snes_grid_sequence 6 -snes_fd_color -snes_converged_reason "
                         "-snes_monitor -ksp_converged_reason -pc_type mg -log_view")            # c/ch8/cluster.sh:70
    assert (o.grid_x, o.grid_y, o.grid_sequence, o.pc_type, o.fd_color) == (33, 33, 6, "mg", True)
    o = pm.parse_options("-snes_fd_color -ms_q 0.0 -ksp_type cg -ksp_converged_reason -da_refine 2 -ms_problem tent")
    assert (o.q, o.ksp_type, o.refine, o.problem) == (0.0, "cg", 2, "tent")


@pytest.mark.parametrize("argv,msg", [
    ("-snes_fd_color -ms_problem sphere", "unknown problem type"),                             # minimal.c:127
    ("-snes_fd_color -ms_problem tent -ms_exact_init", "only possible for -mse_problem catenoid"),   # :109
    ("-snes_fd_color -ms_catenoid_c 0.5", "only valid if c >= 1"),                             # :116
    ("-snes_fd_color -ms_exact_init -ms_q -0.25", "only possible if q=-0.5"),                  # :120
    ("-snes_fd_color -pc_type ilu", "sequential"),
    ("-snes_fd_color -mg_levels_pc_type sor", "sequential"),
    ("-snes_fd_color -p4b_mf_pmat poisson", "option of -snes_mf_operator"),
])
def test_error_paths(argv, msg):
    with pytest.raises(ValueError, match=msg):
        pm.parse_options(argv)


def test_driver_reproduces_golden_lines_of_minimal_test1():
    ops = FakeOps()
    r = pm.minimal_main("-snes_fd_color -snes_converged_reason -snes_monitor_short -ms_problem catenoid "
                        "-ms_catenoid_c 2.0 -da_refine 1", ops)
    assert r.lines[0] == "  0 SNES Function norm 1.08276"                                   # minimal.test1:1
    assert r.lines[-2] == "  Nonlinear solve converged due to CONVERGED_FNORM_RELATIVE iterations 5"   # :7
    assert r.lines[-1] == "done on 5 x 5 grid and problem catenoid:  error |u-uexact|_inf = 1.10603e-04"   # :8
    # :2-6 -- the golden ran GMRES + ILU(0) to rtol 1e-5; the inexactness of THAT linear solve is visible in the 4th
    # digit of the norms (exact Newton steps give 0.69662, 0.170628, ...), so another preconditioner agrees to ~5e-3
    golden = [1.08276, 0.69656, 0.170569, 0.00995652, 2.20675e-05, 1.772e-10]
    got = [float(l.split()[-1]) for l in r.lines[:6]]
    assert len(r.lines) == 8
    np.testing.assert_allclose(got[:5], golden[:5], rtol=1e-2)
    assert 5.0e-11 < got[5] < 1.0e-9                          # (dominated by the last linear solve's own tolerance)


@pytest.mark.parametrize("argv,okw", [
    ("-snes_fd_color -snes_grid_sequence 3 -ms_problem tent -pc_type mg", dict(grid_sequence=3, problem="tent", pc="mg")),
    ("-snes_fd_color -da_refine 3 -pc_type mg -ksp_type cg -ms_q 0.0 -ms_problem tent",
     dict(refine=3, problem="tent", q=0.0, pc="mg", ksp="cg")),
    ("-snes_fd_color -da_grid_x 5 -da_grid_y 9 -snes_grid_sequence 2 -pc_type mg -ms_catenoid_c 1.5",
     dict(mx=5, my=9, grid_sequence=2, pc="mg", catenoid_c=1.5)),
    ("-snes_fd_color -da_refine 2 -pc_type none -ms_problem tent", dict(refine=2, problem="tent", pc="none")),
    ("-snes_fd_color -da_refine 4 -pc_type mg -pc_mg_levels 3", dict(refine=4, pc="mg", mg_levels=3)),
    ("-snes_mf_operator -snes_grid_sequence 2 -pc_type mg", dict(grid_sequence=2, pc="mg", mf_operator=True)),
    ("-snes_mf_operator -da_refine 2 -pc_type none -ms_problem tent", dict(refine=2, problem="tent", pc="none", mf_operator=True)),
])
def test_driver_matches_oracle(argv, okw):
    ops = FakeOps()
    r = pm.minimal_main(argv, ops)
    o = mo.minimal(**okw)
    assert (r.mx, r.my) == (o.mx, o.my)
    assert [s.its for s in r.stages] == [s.its for s in o.stages]
    assert [s.ksp_its for s in r.stages] == [s.ksp_its for s in o.stages]
    for a, b in zip(r.stages, o.stages):
        # (late norms inherit the 1e-5 linear-solve tolerance: rounding-level differences show at ~1e-5 relative)
        np.testing.assert_allclose(a.fnorms, b.fnorms, rtol=1e-3, atol=1e-10 * b.fnorms[0])
    u = r.u.a.reshape(o.u.shape)
    assert np.max(np.abs(u - o.u)) <= 1e-11 * max(1.0, np.max(np.abs(o.u)))
    if o.errinf is not None:
        assert abs(r.errinf - o.errinf) <= 1e-11


@pytest.mark.parametrize("argv,okw", [
    ("-da_refine 2 -pc_type none -ms_problem tent -ms_q 0.0", dict(refine=2, problem="tent", q=0.0, pc="none")),
    ("-da_refine 3 -pc_type none", dict(refine=3, pc="none")),
    ("-da_refine 2 -pc_type none -ksp_type cg -ms_problem tent", dict(refine=2, problem="tent", pc="none", ksp="cg")),
    ("-snes_grid_sequence 2 -pc_type none -ms_catenoid_c 1.5", dict(grid_sequence=2, pc="none", catenoid_c=1.5)),
    ("-da_refine 3 -pc_type mg -ms_problem tent -ms_q 0.0", dict(refine=3, problem="tent", q=0.0, pc="mg")),
    ("-da_refine 3 -pc_type mg -pc_mg_levels 2 -snes_max_it 8", dict(refine=3, pc="mg", mg_levels=2, max_it=8)),
    # -snes_mf_operator with what [PETSc] preconditions it with: the registered Poisson matrix (minimal.test3's route)
    ("-snes_mf_operator -p4b_mf_pmat poisson -snes_grid_sequence 2 -pc_type mg",
     dict(grid_sequence=2, pc="mg", mf_operator=True)),
    ("-snes_mf_operator -p4b_mf_pmat poisson -da_refine 3 -pc_type mg -ms_problem tent",
     dict(refine=3, problem="tent", pc="mg", mf_operator=True)),
])
def test_default_route_uses_the_registered_poisson_jacobian(argv, okw):
    """minimal.c:142-145: without -snes_fd_color / -snes_mf_operator Newton's matrix is Poisson2DJacobianLocal ("ONLY
    APPROXIMATE": exact for -ms_q 0, a slowly converging fixed-point iteration otherwise).  Driver == oracle, step by step."""
    r = pm.minimal_main(argv, FakeOps())
    o = mo.minimal(poisson_jacobian=True, **okw)
    assert [s.its for s in r.stages] == [s.its for s in o.stages]
    assert [s.ksp_its for s in r.stages] == [s.ksp_its for s in o.stages]
    assert [s.reason for s in r.stages] == [s.reason for s in o.stages]
    for a, b in zip(r.stages, o.stages):
        np.testing.assert_allclose(a.fnorms, b.fnorms, rtol=1e-3, atol=1e-10 * b.fnorms[0])
    assert np.max(np.abs(r.u.a.reshape(o.u.shape) - o.u)) <= 1e-10
    if okw.get("q") == 0.0 and okw["pc"] == "none":
        # Laplace: the Poisson matrix IS the Jacobian up to the boundary rows' scaling (4 against minimal.c's 1).  (With
        # multigrid the coarse corrections leave O(ksp_rtol) on the boundary rows, which that scaling then removes only by a
        # factor 3/4 per step, in PETSc as here: the oracle shows the same 13 iterations.)
        assert r.stages[0].its <= 3


def test_stencil9_csr_round_trip():
    rng = np.random.default_rng(0)
    mx, my = 7, 5
    vals = rng.standard_normal(9 * mx * my)
    rp, ci, d = pm.stencil9_to_csr(vals, mx, my)
    import scipy.sparse as sp
    A = sp.csr_matrix((d, ci, rp), shape=(mx * my, mx * my)).toarray()
    np.testing.assert_array_equal(A, pm.stencil9_to_dense(vals, mx, my))
    assert rp[-1] == (3 * mx - 2) * (3 * my - 2)
