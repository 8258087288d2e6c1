"""The C++ Newton-Krylov-multigrid host logic (p4pdes_b200/csrc/nk_solver.hpp -- what p4b_minimal_solve runs on the
device) exercised WITHOUT a GPU: oracle/Makefile instantiates the same template with plain C++ loops
(oracle/native/host_ops.hpp, test infrastructure) and this test compares the runs with the independent Python oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import minimal_solver_oracle as mo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "nk_host_test")


@pytest.fixture(scope="module")
def exe():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "nk_host_test"])
    return EXE


def run(exe, *args):
    p = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    lines = p.stdout.rstrip("\n").split("\n")
    return lines[:-1], json.loads(lines[-1])


@pytest.mark.parametrize("argv,okw", [
    (("-snes_grid_sequence", 3, "-ms_problem", "tent", "-pc_type", "mg"), dict(grid_sequence=3, problem="tent", pc="mg")),
    (("-da_refine", 3, "-pc_type", "mg", "-ksp_type", "cg", "-ms_q", 0.0, "-ms_problem", "tent"),
     dict(refine=3, problem="tent", q=0.0, pc="mg", ksp="cg")),
    (("-da_grid_x", 5, "-da_grid_y", 9, "-snes_grid_sequence", 2, "-pc_type", "mg", "-ms_problem", "tent"),
     dict(mx=5, my=9, grid_sequence=2, pc="mg", problem="tent")),
    (("-da_refine", 2, "-pc_type", "none", "-ms_problem", "tent"), dict(refine=2, problem="tent", pc="none")),
    (("-da_refine", 4, "-pc_type", "mg", "-pc_mg_levels", 3, "-ms_problem", "tent"),
     dict(refine=4, pc="mg", mg_levels=3, problem="tent")),
    (("-snes_mf_operator", "-snes_grid_sequence", 2, "-pc_type", "mg", "-ms_problem", "tent"),
     dict(grid_sequence=2, pc="mg", problem="tent", mf_operator=True)),
    (("-snes_mf_operator", "-da_refine", 2, "-pc_type", "none"), dict(refine=2, pc="none", mf_operator=True)),
])
def test_native_solver_matches_python_oracle(exe, argv, okw):
    _, d = run(exe, "-snes_fd_color", *argv)
    o = mo.minimal(**okw)
    assert (d["mx"], d["my"]) == (o.mx, o.my)
    assert [s["its"] for s in d["stages"]] == [s.its for s in o.stages]
    assert [s["ksp_its"] for s in d["stages"]] == [s.ksp_its for s in o.stages]
    assert all(s["reason"] == t.reason for s, t in zip(d["stages"], o.stages))
    for k, (s, t) in enumerate(zip(d["stages"], o.stages)):
        # (late norms inherit the 1e-5 linear-solve tolerance; the differencing step h ~ 1e-7 leaves ~1e-8 of rounding)
        np.testing.assert_allclose(s["fnorm"], t.fnorms, rtol=1e-4, atol=1e-10 * t.fnorms[0])
        np.testing.assert_allclose(s["lambda"], t.lambdas, rtol=1e-6)
    assert abs(d["sum"] - float(o.u.sum())) <= 1e-9 * abs(float(o.u.sum()))
    if o.errinf is not None:
        assert abs(d["errinf"] - o.errinf) <= 1e-10
    assert d["allocs"] == d["frees"]                      # every vector the solver took from Ops went back


@pytest.mark.parametrize("argv,okw", [
    (("-da_refine", 2, "-pc_type", "none", "-ms_problem", "tent", "-ms_q", 0.0), dict(refine=2, problem="tent", q=0.0, pc="none")),
    (("-da_refine", 3, "-pc_type", "mg", "-snes_max_it", 7), dict(refine=3, pc="mg", max_it=7)),
    (("-snes_grid_sequence", 2, "-pc_type", "mg", "-ms_problem", "tent", "-ms_tent_H", 0.3, "-snes_max_it", 12),
     dict(grid_sequence=2, pc="mg", problem="tent", tent_H=0.3, max_it=12)),
    (("-snes_mf_operator", "-p4b_mf_pmat", "poisson", "-snes_grid_sequence", 2, "-pc_type", "mg"),
     dict(grid_sequence=2, pc="mg", mf_operator=True)),
])
def test_native_solver_on_the_registered_poisson_jacobian(exe, argv, okw):
    """MinimalOpts.jacobian = 1 (no -snes_fd_color): the matrix minimal.c registers, minimal.c:142-145 -- Newton's matrix,
    or the preconditioner's under -snes_mf_operator; the oracle with poisson_jacobian=True, step by step."""
    _, d = run(exe, *argv)
    o = mo.minimal(poisson_jacobian=True, **okw)
    assert [s["its"] for s in d["stages"]] == [s.its for s in o.stages]
    assert [s["ksp_its"] for s in d["stages"]] == [s.ksp_its for s in o.stages]
    assert all(s["reason"] == t.reason for s, t in zip(d["stages"], o.stages))
    for s, t in zip(d["stages"], o.stages):
        np.testing.assert_allclose(s["fnorm"], t.fnorms, rtol=1e-4, atol=1e-10 * t.fnorms[0])
    assert abs(d["sum"] - float(o.u.sum())) <= 1e-9 * abs(float(o.u.sum()))
    assert d["allocs"] == d["frees"]


def test_native_solver_on_the_catenoid_cold_start(exe):
    """The catenoid runs start at u = 0 in the interior (the case the per-entry "ds" differencing could not handle:
    tests/test_minimal_oracle.py): same Newton counts on every stage, same solution and error."""
    _, d = run(exe, "-snes_fd_color", "-da_grid_x", 5, "-da_grid_y", 9, "-snes_grid_sequence", 2, "-pc_type", "mg",
               "-ms_catenoid_c", 1.5)
    o = mo.minimal(mx=5, my=9, grid_sequence=2, pc="mg", catenoid_c=1.5)
    assert all(s["reason"] == "CONVERGED_FNORM_RELATIVE" for s in d["stages"])
    assert [s["its"] for s in d["stages"]] == [s.its for s in o.stages]
    assert abs(d["errinf"] - o.errinf) <= 1e-8 and abs(d["sum"] - float(o.u.sum())) <= 1e-7 * abs(float(o.u.sum()))


def test_native_solver_prints_the_reference_lines(exe):
    lines, d = run(exe, "-snes_fd_color", "-ms_problem", "catenoid", "-ms_catenoid_c", 2.0, "-da_refine", 1, "-monitor")
    assert lines[0] == "  0 SNES Function norm 1.08276"                                            # minimal.test1:1
    assert lines[-1] == "done on 5 x 5 grid and problem catenoid:  error |u-uexact|_inf = 1.10603e-04"   # :8
    assert lines[-2] == "  Nonlinear solve converged due to CONVERGED_FNORM_RELATIVE iterations 5"        # :7
    # :2-6: the golden's own linear solves (ILU(0), rtol 1e-5) show in the 4th digit -- see test_minimal_driver_cpu.py
    golden = [1.08276, 0.69656, 0.170569, 0.00995652, 2.20675e-05]
    got = [float(l.split()[-1]) for l in lines if "SNES Function norm" in l]
    assert len(got) == 6
    np.testing.assert_allclose(got[:5], golden, rtol=1e-2)
    assert all(l.startswith("    Linear solve converged due to CONVERGED_RTOL iterations ") for l in lines if "Linear" in l)


def test_banded_inverse_agrees_with_lapack(exe):
    # the base-grid solve of the native path (banded LU + n solves) against numpy on the Python driver's dense copy:
    # same Newton / Krylov counts with a 17 x 17 base grid (289 unknowns, bandwidth 18)
    _, d = run(exe, "-snes_fd_color", "-da_grid_x", 17, "-da_grid_y", 17, "-snes_grid_sequence", 1, "-pc_type", "mg",
               "-ms_problem", "tent")
    o = mo.minimal(mx=17, my=17, grid_sequence=1, pc="mg", problem="tent")
    assert [s["ksp_its"] for s in d["stages"]] == [s.ksp_its for s in o.stages]
    assert abs(d["sum"] - float(o.u.sum())) <= 1e-9 * abs(float(o.u.sum()))


# ---- pattern.c: the native time-stepping host (csrc/ts_solver.hpp) against the reference's goldens --------------------
from tests.test_pattern_cpu import (GOLDEN_TEST1, GOLDEN_TEST2, GOLDEN_TEST3, GOLDEN_TEST4, GOLDEN_TEST5, TEST1,  # noqa: E402
                                    TEST2, TEST3, TEST4, TEST5)


@pytest.mark.parametrize("argv,golden", [(TEST1, GOLDEN_TEST1), (TEST2, GOLDEN_TEST2), (TEST3, GOLDEN_TEST3),
                                         (TEST4, GOLDEN_TEST4), (TEST5 + " -pc_type none", GOLDEN_TEST5),
                                         (TEST5 + " -pc_type mg", GOLDEN_TEST5)])
def test_native_time_stepper_prints_the_pattern_goldens_verbatim(exe, argv, golden):
    """c/ch5/output/pattern.test1-5: adaptive ARKIMEX3 (incl. the rejected step), backward Euler, Crank-Nicolson, BDF2."""
    lines, d = run(exe, "-pattern", *argv.split())
    assert lines == golden
    assert d["allocs"] == d["frees"]


def test_native_time_stepper_matches_python_oracle(exe):
    from oracle import pattern_solver_oracle as po
    _, d = run(exe, "-pattern", "-da_grid_x", 4, "-da_grid_y", 4, "-da_refine", 3, "-ts_type", "beuler", "-ts_dt", 5,
               "-ts_max_time", 12, "-pc_type", "mg")
    o = po.pattern_beuler(grid=4, refine=3, dt=5.0, tmax=12.0)
    assert d["nsteps"] == len(o.steps) and d["step_newton"] == [s[2].its for s in o.steps]
    assert d["ksp_its_total"] == sum(sum(s[2].ksp_its) for s in o.steps)
    want = float(np.sum(o.Y[..., 0] + 3.0 * o.Y[..., 1]))
    assert abs(d["sum"] - want) <= 1e-10 * abs(want)
    _, d = run(exe, "-pattern", "-da_grid_x", 4, "-da_grid_y", 4, "-da_refine", 2, "-ts_max_time", 60)
    o = po.pattern_arkimex(grid=4, refine=2, tmax=60.0)
    assert d["nsteps"] == len(o.steps) and d["rejected"] == o.rejected
    want = float(np.sum(o.Y[..., 0] + 3.0 * o.Y[..., 1]))
    assert abs(d["sum"] - want) <= 1e-7 * abs(want)           # stage solves: Newton rtol 1e-8 vs the oracle's direct solves


# ---- the callback contract: the REFERENCE's own compiled FormFunctionLocal drives the native solver ---------------------
REFLIB = os.path.join(ROOT, "oracle", "_ref", "libfishref.so")


@pytest.mark.skipif(not os.path.exists(REFLIB), reason="oracle/_ref/libfishref.so not built (no /root/reference)")
def test_reference_callback_drives_the_native_solver(exe):
    """p4b_residual2d_fn (include/p4b200.h) is the FormFunctionLocal contract.  Here the callback is c/ch7/minimal.c's own
    FormFunctionLocal, compiled unchanged (oracle/refstub), and the solver is the product's host logic on plain-C++
    vector operations: golden error of minimal.test1, golden Newton counts of minimal.test4 (with multigrid instead of the
    sequential ILU), and agreement with the run whose residual is the restated kernel formula."""
    _, d = run(exe, "-callback", REFLIB, "-snes_fd_color", "-ms_problem", "catenoid", "-ms_catenoid_c", 2.0, "-da_refine", 1)
    assert "%.5e" % d["errinf"] == "1.10603e-04"                                     # minimal.test1:8
    assert abs(d["stages"][0]["its"] - 5) <= 1 and d["callbacks"] > 9 * d["stages"][0]["its"]
    _, d = run(exe, "-callback", REFLIB, "-snes_fd_color", "-ms_problem", "tent", "-snes_grid_sequence", 2, "-pc_type", "mg")
    assert [s["its"] for s in d["stages"]] == [3, 5, 5]                              # minimal.test4:1-3
    _, own = run(exe, "-snes_fd_color", "-ms_problem", "tent", "-snes_grid_sequence", 2, "-pc_type", "mg")
    assert [s["ksp_its"] for s in d["stages"]] == [s["ksp_its"] for s in own["stages"]]
    assert abs(d["sum"] - own["sum"]) <= 1e-9 * abs(own["sum"])


def test_native_bdf_matches_the_oracle_on_an_adaptive_run(exe):
    """Nothing pins BDF beyond the restart step (pattern.test5): the C++ host and the NumPy oracle, two statements of the
    same algorithm, print the same 25 adaptive steps (5 rejections) to every digit."""
    from oracle import pattern_solver_oracle as po
    for extra in (("-pc_type", "none", "-ksp_rtol", "1e-10"), ("-pc_type", "mg")):
        lines, d = run(exe, "-pattern", "-da_grid_x", 4, "-da_grid_y", 4, "-da_refine", 2, "-ts_type", "bdf", "-ts_monitor", *extra)
        o = po.pattern_bdf(grid=4, refine=2, dt=5.0, tmax=200.0)
        assert [l for l in lines if " TS dt " in l] == [l for l in o.lines if " TS dt " in l]
        assert (d["nsteps"], d["rejected"]) == (len(o.steps), o.rejected) and o.rejected == 5
        want = float((o.Y[..., 0] + 3.0 * o.Y[..., 1]).sum())
        assert abs(d["sum"] - want) <= 1e-8 * abs(want)


def test_classical_gram_schmidt_with_batched_dots_changes_nothing_visible(exe):
    """p4b_tune("gmres_cgs", 1) / -gmres_cgs: [PETSc]'s default orthogonalisation (classical Gram-Schmidt, the dots of a step
    as one VecMDot) instead of modified Gram-Schmidt.  Same iterates up to rounding: the goldens stay verbatim, the
    Krylov totals equal."""
    for argv, golden in ((TEST1, GOLDEN_TEST1), (TEST4, GOLDEN_TEST4), (TEST5 + " -pc_type mg", GOLDEN_TEST5)):
        lines, d = run(exe, "-pattern", *argv.split(), "-gmres_cgs")
        ref, d0 = run(exe, "-pattern", *argv.split())
        assert lines == golden == ref and d["ksp_its_total"] == d0["ksp_its_total"]
    a, da = run(exe, "-snes_fd_color", "-snes_grid_sequence", 3, "-pc_type", "mg", "-monitor")
    b, db = run(exe, "-snes_fd_color", "-snes_grid_sequence", 3, "-pc_type", "mg", "-monitor", "-gmres_cgs")
    assert a == b and [s["ksp_its"] for s in da["stages"]] == [s["ksp_its"] for s in db["stages"]]
