"""The reference's UNCHANGED c/ch7/minimal.c through the PETSc-shaped shim (p4pdes_b200/shim/petscshim.c), on the CPU.

oracle/Makefile compiles minimal.c + poissonfunctions.c from where they lie under /root/reference against
include/petsc.h and links them with the shim source and -- in place of libp4b200.so -- the host stand-in
oracle/native/p4b_standin.cpp (the SAME solver template over plain loops; test infrastructure).  What is checked is the
shim's host logic: option handling, the DMDALocalInfo / a[j][i] views it builds around FormFunctionLocal on every level
and stage, SNESMonitorSet monitors seeing the stage's DM and iterate, the DM / solution replacement under
-snes_grid_sequence, error behaviour -- against the reference's goldens (tests/golden/minimal_goldens.json).
The device instantiation of the same binary (p4pdes_b200/bin/minimal) is tests/test_gpu_r2_shim_minimal.py."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "minimal_goldens.json")))
EXTRA = " -pc_type mg -mg_levels_pc_type jacobi"      # the device path has no ILU / SOR: name the preconditioner


@pytest.fixture(scope="module")
def exe():
    path = os.path.join(ROOT, "oracle", "_ref", "minimal_shim_host")
    if os.path.exists("/root/reference/c/ch7/minimal.c"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref/minimal_shim_host"])
    if not os.path.exists(path):
        pytest.skip("needs the reference tree to compile minimal.c")
    return path


def run(exe, argv, check=True):
    p = subprocess.run([exe] + argv.split(), capture_output=True, text=True, timeout=120)
    if check:
        assert p.returncode == 0, p.stderr
    return p.stdout.splitlines(), p


def test_golden_test4_verbatim(exe):
    """-snes_fd_color -snes_converged_reason -snes_grid_sequence 2 -ms_problem tent: three stages, tab levels 2, 1, 0."""
    g = GOLD["minimal.test4"]
    for extra in (" -pc_type none", EXTRA):
        lines, _ = run(exe, g["options"] + extra)
        assert lines == g["lines"]


def test_golden_test1(exe):
    g = GOLD["minimal.test1"]
    lines, _ = run(exe, g["options"] + " -pc_type none")
    assert len(lines) == len(g["lines"]) and lines[0] == g["lines"][0] and lines[-2:] == g["lines"][-2:]
    # the norms in between carry the golden's own ILU(0)/rtol-1e-5 linear solves in the 4th digit (test_minimal_driver_cpu.py)
    got = [float(l.split()[-1]) for l in lines[1:5]]
    want = [float(l.split()[-1]) for l in g["lines"][1:5]]
    np.testing.assert_allclose(got, want, rtol=1e-2)
    assert re.fullmatch(r"  5 SNES Function norm \d\.\d{3}e-10", lines[5])          # the %5.3e branch of the short monitor


def test_golden_test2_structure(exe):
    g = GOLD["minimal.test2"]
    lines, _ = run(exe, g["options"] + " -pc_type none")
    want = [l for l in g["lines"] if not l.startswith("Matrix is symmetric")]      # -mat_is_symmetric is not provided
    assert len(lines) == len(want) == 3 and lines[-1] == want[-1]
    assert all(re.fullmatch(r"    Linear solve converged due to CONVERGED_RTOL iterations \d+", l) for l in lines[:2])


def test_golden_test3_monitor_lines(exe):
    """-ms_monitor (SNESMonitorSet(MSEMonitor), minimal.c:146-148,286-345) under -snes_grid_sequence 2 -pc_type mg: the
    golden ran -snes_mf_operator on 2 ranks; with -snes_fd_color the Newton path is the same to ~1e-7 in the printed
    areas, and every other character -- tab levels, iteration counts, the lines of the initial, interpolated and
    converged iterates, the final error -- is identical."""
    g = GOLD["minimal.test3"]
    opts = g["options"].replace("-snes_mf_operator", "-snes_fd_color") + " -mg_levels_pc_type jacobi"
    lines, _ = run(exe, opts)
    assert len(lines) == len(g["lines"])
    nexact = 0
    for a, b in zip(lines, g["lines"]):
        if "area" in b:
            fa, fb = [float(x) for x in re.findall(r"[0-9.]+", a)], [float(x) for x in re.findall(r"[0-9.]+", b)]
            assert a[:a.index("area")] == b[:b.index("area")]                      # PetscViewerASCIIAddTab(tab level)
            np.testing.assert_allclose(fa, fb, rtol=0, atol=2e-7)
            nexact += a == b
        else:
            assert a == b
    assert nexact >= 10        # all but the mid-stage iterates, which depend on how the linear systems were solved


def test_golden_test3_with_its_own_command_line(exe):
    """c/ch7/makefile:22: -snes_mf_operator -snes_converged_reason -pc_type mg -snes_grid_sequence 2 -ms_monitor -ms_quaddegree 2
    (plus the device smoother's name).  The Krylov operator is the differenced residual ([PETSc] MatMFFD "wp"), whose
    rounding noise is 1e-8 of J v: the printed areas (8 decimals) agree with the golden to a few units of the last digit
    -- 9e-8 on the one line where the inexact multigrid solves (the golden's and ours) show -- and every other character is equal."""
    g = GOLD["minimal.test3"]
    for extra in (" -mg_levels_pc_type jacobi", " -mg_levels_pc_type jacobi -p4b_recognise_residual 0"):
        lines, _ = run(exe, g["options"] + extra)
        assert len(lines) == len(g["lines"])
        worst = 0.0
        for a, b in zip(lines, g["lines"]):
            if "area" in b:
                fa, fb = [float(x) for x in re.findall(r"[0-9.]+", a)], [float(x) for x in re.findall(r"[0-9.]+", b)]
                assert a[:a.index("area")] == b[:b.index("area")] and fa[1:] == fb[1:]
                worst = max(worst, abs(fa[0] - fb[0]))
            else:
                assert a == b
        assert worst <= 1e-7


def test_solution_and_dm_after_grid_sequencing(exe):
    """minimal.c:161-177 fetches the refined DM and the solution from the SNES: the reported grid and error prove both."""
    lines, _ = run(exe, "-snes_fd_color -snes_grid_sequence 3 -da_grid_x 5 -da_grid_y 4" + EXTRA)
    m = re.fullmatch(r"done on (\d+) x (\d+) grid and problem catenoid:  error \|u-uexact\|_inf = (\S+)", lines[-1])
    assert m and (int(m.group(1)), int(m.group(2))) == (33, 25) and float(m.group(3)) < 2e-4
    # -da_refine and -snes_grid_sequence reach the same grid and the same discrete solution
    l2, _ = run(exe, "-snes_fd_color -da_refine 3 -da_grid_x 5 -da_grid_y 4 -snes_rtol 1e-12" + EXTRA)
    assert abs(float(l2[-1].split()[-1]) - float(m.group(3))) <= 1e-9


@pytest.mark.parametrize("argv,code,msg", [
    ("-snes_fd_color -pc_type none -p4b_mf_pmat ilu", 56, "-p4b_mf_pmat: fd or poisson"),
    ("-snes_fd_color", 56, "ILU"),
    ("-snes_fd_color -pc_type mg", 56, "SOR"),
    ("-snes_fd_color -pc_type jacobi", 56, "-pc_type mg and -pc_type none"),
    ("-snes_fd_color -pc_type none -ksp_type bcgs", 56, "gmres and cg"),
    ("-snes_mf -pc_type none", 56, "not provided"),
    ("-snes_fd_color -pc_type none -ms_problem tent -ms_exact_init", 2, "only possible for -mse_problem catenoid"),
    ("-snes_fd_color -pc_type none -ms_catenoid_c 0.5", 3, "c >= 1"),
    ("-snes_fd_color -pc_type mg -mg_levels_pc_type jacobi -da_grid_x 129 -da_grid_y 129", 61, "65 x 65"),
])
def test_error_paths(exe, argv, code, msg):
    _, p = run(exe, argv, check=False)
    assert p.returncode == code and msg in p.stderr and "PETSC ERROR" in p.stderr


# ---- recognition of the registered residual (include/p4b200.h "Recognition"; csrc/nk_solver.hpp probe_minimal_model) ----
def _route(p):
    m = re.search(r"standin: route (\d), q (\S+), residual callbacks (\d+)", p.stderr)
    return int(m.group(1)), float(m.group(2).rstrip(",")), int(m.group(3))


def run_report(exe, argv):
    p = subprocess.run([exe] + argv.split(), capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, P4B_STANDIN_REPORT="1"))
    assert p.returncode == 0, p.stderr
    return p.stdout.splitlines(), _route(p)


def test_minimal_c_is_recognised_and_both_routes_print_the_same(exe):
    """The unchanged minimal.c's FormFunctionLocal IS the residual the library has as a kernel: after 2 probes per grid + 1
    the solve calls it no more.  With recognition switched off every evaluation is the host callback (hundreds); the
    two routes print the same lines."""
    g = GOLD["minimal.test4"]
    a, (route, q, ncb) = run_report(exe, g["options"] + EXTRA)
    # grids 3x3, 5x5, 9x9: two probes each + one to identify q, then one re-verification at each converged iterate
    assert a == g["lines"] and route == 1 and q == -0.5 and ncb == 2 * 3 + 1 + 3
    b, (route0, _, ncb0) = run_report(exe, g["options"] + EXTRA + " -p4b_recognise_residual 0")
    assert b == g["lines"] and route0 == 0 and ncb0 > 200
    # another exponent and boundary problem, with the monitor: identified exactly as the option was parsed
    argv = "-snes_fd_color -snes_converged_reason -da_refine 3 -ms_q -0.3 -ms_catenoid_c 1.3 -ms_monitor" + EXTRA
    c, (route, q, _) = run_report(exe, argv)
    d, (route0, _, _) = run_report(exe, argv + " -p4b_recognise_residual 0")
    assert route == 1 and q == float("-0.3") and route0 == 0 and c == d


@pytest.fixture(scope="module")
def snes_variants(exe, tmp_path_factory):
    d = tmp_path_factory.mktemp("sv")
    obj, out = str(d / "sv.o"), str(d / "sv")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), "-c",
                           os.path.join(ROOT, "tests", "shim_cases", "snes_variants.c"), "-o", obj])
    o = os.path.join(ROOT, "oracle", "_ref", "obj")
    subprocess.check_call(["g++", obj, os.path.join(o, "petscshim.o"), os.path.join(o, "p4b_standin.o"), "-o", out, "-lm"])
    return out


def test_a_residual_that_is_not_the_model_stays_a_host_callback(snes_variants):
    """tests/shim_cases/snes_variants.c: the model written differently, with its own Dirichlet data and exponent, is
    recognised; the model plus a reaction term is not, is solved through host callbacks, and its answer is the one an
    independent NumPy Newton solve of the same equations gives."""
    from oracle import fish_oracle as fo
    from oracle import minimal_pattern_oracle as mpo
    from oracle import minimal_solver_oracle as mo
    argv = "-snes_fd_color -da_refine 2 -snes_rtol 1e-12" + EXTRA
    a, (route, q, ncb) = run_report(snes_variants, "-variant 0 " + argv)
    b, (route0, _, _) = run_report(snes_variants, "-variant 0 -p4b_recognise_residual 0 " + argv)
    assert route == 1 and q == float("-0.35") and ncb == 2 * 3 + 1 + 1 and route0 == 0 and a == b
    c, (route1, _, ncb1) = run_report(snes_variants, "-variant 1 " + argv)
    assert route1 == 0 and ncb1 > 100 and c != a
    m = 17
    x = np.linspace(0.0, 1.0, m)
    X, Y = np.meshgrid(x, x)
    g = 0.4 * np.sin(3.0 * X + 1.0) * np.cos(2.0 * Y) + 0.2 * X * Y
    h = 1.0 / (m - 1)
    inner = np.zeros((m, m), bool)
    inner[1:-1, 1:-1] = True
    for variant, lines in ((0, a), (1, c)):
        F = lambda u, v=variant: mpo.minimal_function(u, g, -0.35) + (5.0 * h * h * u ** 3 * inner if v else 0.0)
        r = mo.newton(F, np.full((m, m), 0.1), lambda J, uu: fo.ILU0PC(J).apply, snes_rtol=1e-12)
        got = re.fullmatch(r"done on 17 x 17 grid: sum (\S+) max (\S+)", lines[-1])
        assert abs(float(got.group(1)) - r.u.sum()) <= 1e-8 * abs(r.u.sum()) and abs(float(got.group(2)) - r.u.max()) <= 1e-9


def test_a_term_the_probes_cannot_see_is_caught_at_the_converged_iterate(snes_variants):
    """ADVICE r1 (probe_minimal_model): a callback that equals the model at the probe states but not where the solve ends
    up.  Variant 2 adds a term that acts only where u < -0.05; the probes' interior values lie in [0, 0.5].  The callback
    is recognised, the device solve converges -- to the model's answer, not the caller's -- the re-verification at the
    converged iterate sees the difference, says so on stderr, and the solve is repeated through the host callback."""
    argv = "-snes_fd_color -da_refine 2 -snes_rtol 1e-12" + EXTRA
    p = subprocess.run([snes_variants] + ("-variant 2 " + argv).split(), capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, P4B_STANDIN_REPORT="1"))
    assert p.returncode == 0, p.stderr
    assert "matched the library's kernel at the probes but not at a converged iterate" in p.stderr
    assert _route(p)[0] == 0                                   # the answer came from the host-callback route
    forced, _ = run_report(snes_variants, "-variant 2 -p4b_recognise_residual 0 " + argv)
    model, _ = run_report(snes_variants, "-variant 0 " + argv)
    assert p.stdout.splitlines()[-1] == forced[-1] and forced[-1] != model[-1]


# ---- the Jacobian callback minimal.c REGISTERS (minimal.c:142-145: Poisson's, "ONLY APPROXIMATE") -------------------------
@pytest.mark.parametrize("argv", [
    "-da_refine 2 -pc_type none -ms_problem tent -ms_q 0.0 -snes_converged_reason -snes_monitor_short",
    "-da_refine 3 -pc_type mg -snes_converged_reason -snes_max_it 6 -snes_monitor",
    "-da_grid_x 5 -da_grid_y 9 -snes_grid_sequence 2 -pc_type mg -ms_problem tent -ms_tent_H 0.3 -snes_converged_reason "
    "-snes_max_it 12 -snes_monitor -ksp_converged_reason",
    "-snes_mf_operator -p4b_mf_pmat poisson -snes_grid_sequence 2 -pc_type mg -snes_converged_reason -snes_monitor "
    "-ksp_converged_reason",
])
def test_without_fd_color_the_registered_poisson_jacobian_is_newtons_matrix(exe, argv):
    """What [PETSc] does with the unchanged minimal.c when -snes_fd_color is absent: the matrix comes from the registered
    callback.  The shim runs that callback (start and final grid), checks that its rows are the library's Poisson matrix and
    lets the kernel stand in for it; the printed lines are those of the Python host over the NumPy stand-in, character for
    character (that driver == the oracle: tests/test_minimal_driver_cpu.py)."""
    from p4pdes_b200 import minimal as pm
    from tests.fake_ops import FakeOps
    lines, _ = run(exe, argv + (" -mg_levels_pc_type jacobi" if "-pc_type mg" in argv else ""))
    want = pm.minimal_main(argv, FakeOps()).lines
    assert len(lines) == len(want)
    for a, b in zip(lines, want):
        if "SNES Function norm" in a and a != b:        # late norms carry the rounding of the differenced operator
            assert a.split()[:4] == b.split()[:4] and abs(float(a.split()[-1]) - float(b.split()[-1])) <= 1e-4 * float(b.split()[-1]) + 1e-11
        else:
            assert a == b


def test_golden_test3_on_petscs_own_route(exe):
    """minimal.test3 once more, now as [PETSc] ran it: -snes_mf_operator preconditioned by the REGISTERED Poisson matrix
    (-p4b_mf_pmat poisson; the default keeps the FD-coloured Jacobian).  Newton counts 5 / 3 / 3, tab levels, error line
    equal; areas to <= 1e-7 (the golden's Chebyshev/SOR solves on 2 ranks against Chebyshev/Jacobi here)."""
    g = GOLD["minimal.test3"]
    lines, _ = run(exe, g["options"] + " -mg_levels_pc_type jacobi -p4b_mf_pmat poisson")
    assert len(lines) == len(g["lines"])
    for a, b in zip(lines, g["lines"]):
        if "area" in b:
            fa, fb = [float(x) for x in re.findall(r"[0-9.]+", a)], [float(x) for x in re.findall(r"[0-9.]+", b)]
            assert a[:a.index("area")] == b[:b.index("area")] and fa[1:] == fb[1:] and abs(fa[0] - fb[0]) <= 1e-7
        else:
            assert a == b


def test_a_registered_jacobian_that_is_not_the_librarys_is_refused(snes_variants):
    # no callback registered and no -snes_fd_color: PETSc would difference densely; here an error that names the option
    _, p = run(snes_variants, "-variant 0 -pc_type none", check=False)
    assert p.returncode == 73 and "no Jacobian callback registered" in p.stderr and "-snes_fd_color" in p.stderr
    # the Laplacian rows of the unit square: accepted, and a (slowly converging) Newton iteration runs on them
    ok, p = run(snes_variants, "-variant 0 -jac 1 -da_refine 1 -pc_type none -snes_max_it 4 -snes_converged_reason")
    assert ok[0].strip() == "Nonlinear solve did not converge due to DIVERGED_MAX_IT iterations 4" and ok[-1].startswith("done on 9 x 9")
    fd, _ = run(snes_variants, "-variant 0 -jac 1 -da_refine 1 -pc_type none -snes_fd_color -snes_converged_reason")
    assert "CONVERGED_FNORM_RELATIVE" in fd[0]          # with -snes_fd_color the callback is not used at all
    for k, why in ((2, "deviation"), (3, "column to a boundary node was not dropped")):
        _, p = run(snes_variants, "-variant 0 -jac %d -da_refine 1 -pc_type none" % k, check=False)
        assert p.returncode == 56 and "is not Poisson2DJacobianLocal" in p.stderr and why in p.stderr \
            and "pass -snes_fd_color" in p.stderr
