"""Both routes of the shim's SNESSolve / TSSolve on the device, with drivers that are NOT the reference's
(tests/shim_cases/*.c, compiled here against include/petsc.h and the in-tree libraries): the model written differently
must be recognised and run device-resident; a system the library has no kernels for must run through host callbacks.
The expected numbers were produced on the CPU through the host stand-in and confirmed there by independent NumPy solves
(tests/test_shim_minimal_cpu.py, tests/test_shim_pattern_cpu.py).  First run on a B200 in round 2 (profiles/r02_pending.md) and promoted to the `gpu` marker."""
import os
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "p4pdes_b200", "lib")
MG = " -pc_type mg -mg_levels_pc_type jacobi"

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device"),
              pytest.mark.skipif(not os.path.exists(os.path.join(LIB, "libpetsc_p4b200.so")), reason="shim library not built")]


@pytest.fixture(scope="module")
def drivers(tmp_path_factory):
    d = tmp_path_factory.mktemp("variants")
    out = {}
    for name in ("snes_variants", "ts_variants", "ksponly_varcoef"):
        exe = str(d / name)
        subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-O2", "-I", os.path.join(ROOT, "include"),
                               os.path.join(ROOT, "tests", "shim_cases", name + ".c"), "-o", exe, "-L", LIB, "-lpetsc_p4b200",
                               "-lp4b200", "-Wl,-rpath," + LIB, "-lm"])
        out[name] = exe
    return out


def run(exe, argv):
    p = subprocess.run([exe] + argv.split(), capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    return p.stdout.splitlines()


def test_snes_routes(drivers):
    argv = "-snes_fd_color -da_refine 2 -snes_rtol 1e-12 -log_view" + MG
    a = run(drivers["snes_variants"], "-variant 0 " + argv)
    b = run(drivers["snes_variants"], "-variant 0 -p4b_recognise_residual 0 " + argv)
    c = run(drivers["snes_variants"], "-variant 1 " + argv)
    assert "SNES newtonls: residual recognised as the library's kernel: evaluated on the device" in a
    assert "SNES newtonls: residual evaluated by the host callback" in b and "SNES newtonls: residual evaluated by the host callback" in c
    val = lambda lines: [float(x) for x in lines[0].split("sum ")[1].replace("max", "").split()]
    for got, want in ((val(a), (2.328160025196e+01, 3.999862342713e-01)), (val(b), (2.328160025196e+01, 3.999862342713e-01)),
                      (val(c), (2.314385272737e+01, 3.999862342713e-01))):
        assert abs(got[0] - want[0]) <= 1e-8 * want[0] and abs(got[1] - want[1]) <= 1e-9


def test_ts_routes(drivers):
    argv = "-da_refine 2 -pc_type none -ts_type beuler -ts_dt 2 -ts_max_time 6 -snes_rtol 1e-10"
    a = run(drivers["ts_variants"], "-variant 0 " + argv)
    b = run(drivers["ts_variants"], "-variant 0 -p4b_recognise_residual 0 " + argv)
    c = run(drivers["ts_variants"], "-variant 1 " + argv)
    norm = lambda lines: float(lines[-1].split()[-1])
    assert abs(norm(a) - 1.4146570989e+01) <= 1e-8 * 14.0 and abs(norm(b) - norm(a)) <= 1e-8 * 14.0
    assert abs(norm(c) - 1.4210174635e+01) <= 1e-8 * 14.0
    p = subprocess.run([drivers["ts_variants"]] + ("-variant 1 -da_refine 2" + MG).split(), capture_output=True, text=True)
    assert p.returncode == 56 and "pass -pc_type none" in p.stderr


def test_assembled_mat_type(drivers):
    """-mat_type sellcuda: the Jacobian callback's values as inserted -> CSR -> SELL-32 on the device, KSPCG over p4b_sell_spmv.
    (1) the unchanged fish.c gives the structured type's iteration counts and report; (2) a variable-coefficient problem the
    structured type refuses is solved, with the sums an independent NumPy solve gives (tests/test_shim_fish_cpu.py)."""
    fishexe = os.path.join(ROOT, "p4pdes_b200", "bin", "fish")
    if os.path.exists(fishexe):
        common = "-fsh_dim 3 -fsh_problem manupoly -da_refine 3 -ksp_rtol 1.0e-12 -pc_type none -ksp_converged_reason -snes_monitor_short"
        assert run(fishexe, common) == run(fishexe, common + " -mat_type sellcuda")
    p = subprocess.run([drivers["ksponly_varcoef"]] + "-da_refine 3 -pc_type none".split(), capture_output=True, text=True)
    assert p.returncode == 56 and "-mat_type sellcuda" in p.stderr
    lines = run(drivers["ksponly_varcoef"], "-da_refine 3 -pc_type none -ksp_rtol 1e-13 -mat_type sellcuda")
    got = lines[-1].split()
    assert abs(float(got[7]) - 8.938150776637e+02) <= 1e-8 * 893.8 and abs(float(got[9]) - 7.822558518301e-01) <= 1e-9
