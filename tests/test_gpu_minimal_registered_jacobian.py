"""minimal.c WITHOUT -snes_fd_color on the device: the matrix is the one the driver registers, Poisson2DJacobianLocal
(c/ch7/minimal.c:142-145, "ONLY APPROXIMATE"; c/ch6/poissonfunctions.c:152-193) -- Newton's matrix by default, the
preconditioner's under -snes_mf_operator (-p4b_mf_pmat poisson, [PETSc]'s choice; minimal.test3 ran that way).
The kernel (p4b_poisson_stencil9) against the oracle's matrix; the Python host, the C++ host inside the library and the
unchanged minimal.c under the shim against the oracle (poisson_jacobian=True) and against each other.
CPU counterparts: tests/test_minimal_driver_cpu.py, test_native_nk_cpu.py, test_shim_minimal_cpu.py."""
import json
import os
import re
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import fish_oracle as fo
from oracle import minimal_solver_oracle as mo
from p4pdes_b200 import minimal as pm
from p4pdes_b200.fish import Context

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "p4pdes_b200", "bin", "minimal")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "minimal_goldens.json")))

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


@pytest.fixture(scope="module")
def ctx():
    return Context()


@pytest.mark.parametrize("mx,my,Lx,Ly,cx,cy", [(3, 3, 1.0, 1.0, 1.0, 1.0), (9, 5, 1.0, 1.0, 1.0, 1.0), (33, 65, 2.0, 0.5, 1.0, 3.0),
                                               (130, 37, 1.0, 1.0, 1.0, 1.0)])
def test_poisson_stencil9_kernel_is_the_matrix_the_callback_inserts(ctx, mx, my, Lx, Ly, cx, cy):
    vals = ctx.empty(9 * mx * my)
    ctx.poisson_stencil9(mx, my, Lx, Ly, cx, cy, vals)
    rp, ci, d = pm.stencil9_to_csr(ctx.to_host(vals), mx, my)
    J = sp.csr_matrix((d, ci, rp), shape=(mx * my, mx * my))
    Jo = fo.jacobian(fo.Grid(2, (mx, my, 1), (Lx, Ly, 1.0)), (cx, cy, 1.0))
    assert abs(J - Jo).max() <= 1e-14 * abs(Jo).max()
    J.eliminate_zeros()
    assert J.nnz == Jo.nnz


CASES = [
    ("-da_refine 2 -pc_type none -ms_problem tent -ms_q 0.0", dict(refine=2, problem="tent", q=0.0, pc="none")),
    ("-da_refine 3 -pc_type mg -snes_max_it 7", dict(refine=3, pc="mg", max_it=7)),
    ("-snes_grid_sequence 3 -pc_type mg -ms_problem tent -ms_tent_H 0.3 -snes_max_it 12",
     dict(grid_sequence=3, pc="mg", problem="tent", tent_H=0.3, max_it=12)),
    ("-snes_mf_operator -p4b_mf_pmat poisson -snes_grid_sequence 2 -pc_type mg", dict(grid_sequence=2, pc="mg", mf_operator=True)),
    ("-snes_mf_operator -p4b_mf_pmat poisson -da_refine 3 -pc_type mg -ms_problem tent",
     dict(refine=3, problem="tent", pc="mg", mf_operator=True)),
]
_ORACLE = {}


def oracle_run(argv, okw):
    if argv not in _ORACLE:                       # (the two hosts are compared with the same oracle run)
        _ORACLE[argv] = mo.minimal(poisson_jacobian=True, **okw)
    return _ORACLE[argv]


@pytest.mark.parametrize("native", [False, True])
@pytest.mark.parametrize("argv,okw", CASES)
def test_device_solve_matches_oracle(ctx, argv, okw, native):
    r = pm.minimal_main(argv, ctx, native=native)
    o = oracle_run(argv, okw)
    assert (r.mx, r.my) == (o.mx, o.my)
    for a, b in zip(r.stages, o.stages):
        # (a stagnating stage ends on the step-size test or on -snes_max_it, whichever rounding lets come first)
        assert a.reason == b.reason or {a.reason, b.reason} <= {"CONVERGED_SNORM_RELATIVE", "DIVERGED_MAX_IT"}
        assert abs(a.its - b.its) <= 1
        assert abs(max(a.ksp_its, default=0) - max(b.ksp_its, default=0)) <= 1
        k = min(len(a.fnorms), len(b.fnorms), 6)
        np.testing.assert_allclose(a.fnorms[:k], b.fnorms[:k], rtol=1e-3, atol=1e-9 * b.fnorms[0])
    if all(s.reason.startswith("CONVERGED_FNORM") for s in o.stages):
        u = ctx.to_host(r.u).reshape(o.u.shape)
        assert np.max(np.abs(u - o.u)) <= 1e-7 * max(1.0, np.max(np.abs(o.u)))


@pytest.mark.parametrize("native", [False, True])
def test_a_failed_line_search_is_a_reason_not_an_error(ctx, native):
    """The catenoid from the zero interior on 33 x 33: the Poisson step is no descent direction for ||F||^2 and the cubic
    backtracking gives up -- [PETSc] reports DIVERGED_LINE_SEARCH and returns the iterate (the oracle shows the same)."""
    r = pm.minimal_main("-da_refine 4 -pc_type mg -snes_converged_reason", ctx, native=native)
    assert r.stages[0].reason == "DIVERGED_LINE_SEARCH" and r.stages[0].its == 0
    assert r.lines[0] == "  Nonlinear solve did not converge due to DIVERGED_LINE_SEARCH iterations 0"


def test_golden_minimal_test3_as_petsc_ran_it(ctx):
    """c/ch7/output/minimal.test3: Newton counts 5, 3, 3 and the error line (path-independent digits)."""
    r = pm.minimal_main("-snes_mf_operator -p4b_mf_pmat poisson -snes_converged_reason -pc_type mg -snes_grid_sequence 2", ctx,
                        native=True)
    assert [s.its for s in r.stages] == [5, 3, 3]
    assert r.lines[-1] == "done on 9 x 9 grid and problem catenoid:  error |u-uexact|_inf = 6.79501e-04"     # minimal.test3:18


def test_larger_grid_mf_operator_poisson_preconditioner(ctx):
    """33 x 33 -> 513 x 513 by grid sequencing: the Poisson multigrid keeps the Krylov counts bounded (the diffusivity of the
    catenoid stays in [0.5, 1], minimal.test3's D range); nothing is differenced but the operator's action."""
    r = pm.minimal_main("-da_grid_x 33 -da_grid_y 33 -snes_grid_sequence 4 -snes_mf_operator -p4b_mf_pmat poisson -pc_type mg",
                        ctx, native=True)
    assert (r.mx, r.my) == (513, 513) and all(s.reason.startswith("CONVERGED") for s in r.stages)
    assert max(max(s.ksp_its) for s in r.stages) <= 40 and r.errinf < 2e-6


@pytest.mark.skipif(not os.path.exists(EXE), reason="p4pdes_b200/bin/minimal was not built")
def test_unchanged_minimal_c_without_fd_color():
    def run(argv, ok=True):
        p = subprocess.run([EXE] + argv.split(), capture_output=True, text=True, timeout=600)
        assert (p.returncode == 0) == ok, p.stderr
        return p.stdout.splitlines(), p

    # Laplace's equation (-ms_q 0): the registered matrix is the Jacobian up to the boundary rows' scaling
    lines, _ = run("-da_refine 2 -pc_type none -ms_problem tent -ms_q 0.0 -snes_converged_reason -snes_monitor_short")
    assert lines[0] == "  0 SNES Function norm 1.65831" and lines[1].startswith("  1 SNES Function norm 6.26")
    assert lines[-2] == "  Nonlinear solve converged due to CONVERGED_FNORM_RELATIVE iterations 2"
    # minimal.test3 with its own command line + PETSc's preconditioner matrix
    g = GOLD["minimal.test3"]
    lines, _ = run(g["options"] + " -mg_levels_pc_type jacobi -p4b_mf_pmat poisson")
    assert len(lines) == len(g["lines"])
    for a, b in zip(lines, g["lines"]):
        if "area" in b:
            fa, fb = [float(x) for x in re.findall(r"[0-9.]+", a)], [float(x) for x in re.findall(r"[0-9.]+", b)]
            assert a[:a.index("area")] == b[:b.index("area")] and fa[1:] == fb[1:] and abs(fa[0] - fb[0]) <= 1e-7
        else:
            assert a == b
