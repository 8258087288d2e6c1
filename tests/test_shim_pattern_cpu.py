"""The reference's UNCHANGED c/ch5/pattern.c through the PETSc-shaped shim (p4pdes_b200/shim/petscshim.c:TSSolve), on the CPU.

oracle/Makefile compiles pattern.c from where it lies under /root/reference against include/petsc.h and links it with
the shim source and the host stand-in for the library calls TSSolve makes (oracle/native/p4b_standin.cpp: the same
time-stepping template over plain loops; test infrastructure).  What is checked is the shim's host logic: the periodic
two-component DMDA (ghosted a[j][i] views, coordinates), identification of the model's numbers from the registered
callbacks, verification of those callbacks (functions at a generic state, every Jacobian row), which callbacks are
invoked for which -ts_type, option mapping, and that callbacks which are NOT the model are refused -- against the
reference's goldens (tests/golden/pattern_goldens.json) and a variants driver written for this purpose
(tests/shim_cases/ts_variants.c).  The device run of the same binary is tests/test_gpu_r2_shim_pattern.py."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "pattern_goldens.json")))
MG = " -pc_type mg -mg_levels_pc_type jacobi"


@pytest.fixture(scope="module")
def exe():
    path = os.path.join(ROOT, "oracle", "_ref", "pattern_shim_host")
    if os.path.exists("/root/reference/c/ch5/pattern.c"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref/pattern_shim_host"])
    if not os.path.exists(path):
        pytest.skip("needs the reference tree to compile pattern.c")
    return path


@pytest.fixture(scope="module")
def variants(exe, tmp_path_factory):
    """tests/shim_cases/ts_variants.c against the same shim + stand-in objects."""
    d = tmp_path_factory.mktemp("tsv")
    obj, out = str(d / "tsv.o"), str(d / "tsv")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), "-c",
                           os.path.join(ROOT, "tests", "shim_cases", "ts_variants.c"), "-o", obj])
    o = os.path.join(ROOT, "oracle", "_ref", "obj")
    subprocess.check_call(["g++", obj, os.path.join(o, "petscshim.o"), os.path.join(o, "p4b_standin.o"), "-o", out, "-lm"])
    return out


def run(exe, argv, check=True):
    p = subprocess.run([exe] + argv.split(), capture_output=True, text=True, timeout=300)
    if check:
        assert p.returncode == 0, p.stderr
    return p.stdout.splitlines(), p


@pytest.mark.parametrize("name,extra", [
    ("pattern.test1", " -pc_type none"),       # adaptive ARKIMEX3, 12 steps
    ("pattern.test1", MG),
    ("pattern.test2", " -mg_levels_pc_type jacobi"),       # backward Euler, -pc_type mg is in the golden's own options
    ("pattern.test3", " -pc_type none"),       # Crank-Nicolson, -snes_fd_color
    ("pattern.test4", MG),                     # incl. the rejected step and pattern.c's own CALL-BACK REPORT
    ("pattern.test4", " -pc_type none"),
    ("pattern.test5", " -pc_type none"),       # BDF2: the restart step; all four callbacks are called (and checked)
    ("pattern.test5", MG),
])
def test_goldens_verbatim(exe, name, extra):
    g = GOLD[name]
    lines, _ = run(exe, g["options"] + extra)
    assert lines == g["lines"]


def test_call_back_report_reflects_what_petsc_would_call(exe):
    """pattern.c:127-135 prints flags its own callbacks set.  IMEX never calls the RHS Jacobian (pattern.test4: 'RHSJacobian:
    0'); the fully implicit types call all four (pattern.test5: all 1); -snes_fd_color calls no Jacobian callback."""
    tail = lambda argv: run(exe, argv)[0][-2:]
    assert tail("-da_refine 2 -ptn_call_back_report -ts_max_time 5" + MG) == [
        "  IFunction:   1  | IJacobian:   1", "  RHSFunction: 1  | RHSJacobian: 0"]
    assert tail("-da_refine 2 -ptn_call_back_report -ts_type beuler -ts_max_time 5" + MG) == [
        "  IFunction:   1  | IJacobian:   1", "  RHSFunction: 1  | RHSJacobian: 1"]
    assert tail("-da_refine 2 -ptn_call_back_report -ts_type cn -ts_max_time 5 -ptn_no_rhsjacobian" + MG) == [
        "  IFunction:   1  | IJacobian:   1", "  RHSFunction: 1  | RHSJacobian: 0"]
    assert tail("-da_refine 2 -ptn_call_back_report -ts_type cn -ts_max_time 5 -snes_fd_color" + MG) == [
        "  IFunction:   1  | IJacobian:   0", "  RHSFunction: 1  | RHSJacobian: 0"]


def test_model_parameters_are_identified_from_the_callbacks(exe):
    """-ptn_* options are pattern.c's, not the shim's: changed values reach the device path through the probes.  A stiffer
    diffusion changes the adaptive step sequence; the run with the defaults given explicitly is the golden."""
    g = GOLD["pattern.test1"]
    same, _ = run(exe, g["options"] + " -ptn_Du 8.0e-5 -ptn_Dv 4.0e-5 -ptn_phi 0.024 -ptn_kappa 0.06 -ptn_L 2.5 -pc_type none")
    assert same == g["lines"]
    other, _ = run(exe, g["options"] + " -ptn_Du 3.0e-4 -ptn_kappa 0.05 -pc_type none")
    assert other[0] == g["lines"][0] and other[1] == g["lines"][1] and other[2:] != g["lines"][2:]
    assert other[-1].endswith("time 200.")


@pytest.mark.parametrize("argv,code,msg", [
    ("-da_refine 2", 56, "ILU"),
    ("-da_refine 2 -pc_type mg", 56, "SOR"),
    ("-da_refine 2 -pc_type none -ts_type rk", 56, "-ts_type rk is not provided"),
    ("-da_refine 2 -pc_type none -ptn_no_ijacobian", 56, "no IJacobian callback registered"),
    ("-da_grid_x 4 -da_refine 2 -pc_type none", 1, "pattern.c requires mx == my"),
    ("-da_grid_x 64 -da_grid_y 64" + MG, 61, "coarser -da_grid"),
])
def test_error_paths(exe, argv, code, msg):
    _, p = run(exe, argv, check=False)
    assert p.returncode == code and msg in p.stderr and "PETSC ERROR" in p.stderr


def test_other_parameter_values_run_and_other_models_take_the_general_route(variants):
    ok, _ = run(variants, "-variant 0 -da_refine 2 -ts_monitor" + MG)
    assert ok[-1].startswith("done: |Y|_2 = ") and ok[-2].endswith("time 20.")
    # callbacks that are not the model: with multigrid asked for, a refusal that names the way out ...
    for v, what in ((1, "RHSFunction is not G"), (2, "IFunction is not F")):
        _, p = run(variants, "-variant %d -da_refine 2" % v + MG, check=False)
        assert p.returncode == 56 and what in p.stderr and "max deviation" in p.stderr and "pass -pc_type none" in p.stderr
    # a forcing that vanishes at t = 0 is seen by the probes at later times (one probe at t = 0 is not enough)
    _, p = run(variants, "-variant 5 -da_refine 2" + MG, check=False)
    assert p.returncode == 56 and "but not at t = " in p.stderr and "pass -pc_type none" in p.stderr
    v5, _ = run(variants, "-variant 5 -da_refine 2 -pc_type none -ts_type beuler -ts_dt 2 -ts_max_time 6 -snes_rtol 1e-10")
    v0, _ = run(variants, "-variant 0 -da_refine 2 -pc_type none -ts_type beuler -ts_dt 2 -ts_max_time 6 -snes_rtol 1e-10")
    assert v5[-1].startswith("done: |Y|_2 = ") and v5[-1] != v0[-1]
    # ... a Jacobian callback that contradicts its own (model) functions is refused outright
    _, p = run(variants, "-variant 3 -da_refine 2" + MG, check=False)
    assert p.returncode == 56 and "IJacobian does not insert" in p.stderr
    # a wrong RHS Jacobian is only seen where PETSc would call it: the IMEX default never does
    ok4, _ = run(variants, "-variant 4 -da_refine 2 -ts_monitor" + MG)
    assert ok4 == ok
    _, p = run(variants, "-variant 4 -da_refine 2 -ts_type beuler" + MG, check=False)
    assert p.returncode == 56 and "RHSJacobian does not insert" in p.stderr


def test_general_matrix_free_route(exe, variants):
    """p4b_ts2d_solve under the shim: F and G are the user's host callbacks, the stage operator the differenced residual.
    (1) The model itself through this route (recognition switched off) gives what the kernel route gives -- the unchanged
    pattern.c prints pattern.test4's adaptive steps verbatim (its own report then says the Jacobian callback was never
    called); (2) a system the library has no kernels for (extra cubic reaction term) agrees with an independent NumPy
    backward-Euler solve of the same equations."""
    g = GOLD["pattern.test4"]
    lines, _ = run(exe, g["options"] + " -pc_type none -p4b_recognise_residual 0 -log_view")
    assert lines[:len(g["lines"]) - 2] == g["lines"][:-2] and lines[len(g["lines"]) - 2] == "  IFunction:   1  | IJacobian:   0"
    assert "TS: callbacks evaluated on the host, matrix-free stage operator (not the library's model)" in lines
    argv = "-da_refine 2 -pc_type none -ts_type beuler -ts_dt 2 -ts_max_time 6 -snes_rtol 1e-10"
    a, _ = run(variants, "-variant 0 " + argv)
    b, _ = run(variants, "-variant 0 -p4b_recognise_residual 0 " + argv)
    assert a == b
    c, _ = run(variants, "-variant 1 " + argv)
    # the same three steps in NumPy: Newton on F(W, (W - Y)/dt) - G(W) with the analytic Jacobian
    import numpy as np
    import scipy.sparse as sp
    from oracle import fish_oracle as fo
    from oracle import minimal_pattern_oracle as mpo
    from oracle import minimal_solver_oracle as mso
    from oracle import pattern_solver_oracle as po
    m, par = 16, dict(L=2.0, Du=6.0e-5, Dv=3.5e-5, phi=0.03, kappa=0.055)
    sx = np.sin(2.0 * np.pi * np.arange(m) / m)
    Y = np.zeros((m, m, 2))
    Y[..., 1] = 0.25 * (sx[None, :] ** 2) * (sx[:, None] ** 2)
    Y[..., 0] = 1.0 - 2.0 * Y[..., 1]
    dt = 2.0

    def G(W):
        out = mpo.pattern_rhsfunction(W, par["phi"], par["kappa"])
        out[..., 0] += 1.0e-3 * W[..., 0] ** 3
        return out

    for _ in range(3):
        Y0 = Y.copy()
        R = lambda W: mpo.pattern_ifunction(W, (W - Y0) / dt, par["L"], par["Du"], par["Dv"]) - G(W)

        def jac(W):
            d = np.zeros_like(W)
            d[..., 0] = 3.0e-3 * W[..., 0] ** 2
            return (po.stage_jacobian(W, 1.0 / dt, True, **par) - sp.diags(d.ravel())).tocsr()

        Y = mso.newton(R, Y0, lambda J, W: fo.ILU0PC(J).apply, jac=jac, snes_rtol=1e-12).u
    got = float(c[-1].split()[-1])
    assert abs(got - np.linalg.norm(Y)) <= 1e-8 * np.linalg.norm(Y) and c != a
