"""The reference's UNCHANGED c/ch5/heat.c through the PETSc-shaped shim (p4pdes_b200/shim/petscshim.c), on the CPU.

heat.c is not pattern.c's DMDA (one component, DM_BOUNDARY_NONE in x, periodic in y, RHSFunction + RHSJacobian only):
TSSolve (ts_solve_any_dmda) probes the registered RHSFunction against the library's heat kernel -- heat.c's IS that
function, so its run is p4b_heat_solve (G and the stage operator are kernels) -- and otherwise, or with
-p4b_recognise_residual 0, takes the callback route (p4b_ts_solve_callbacks: the user's G on ghosted a[j][i] views on the
host, the integrators and the vector algebra behind the C ABI).  Here the library is the host stand-in
oracle/native/p4b_standin.cpp (the same templates over plain loops); on a GPU box: tests/test_gpu_heat.py.  Checked: both goldens verbatim, TSMonitorSet / TSGetDM / TSGetTimeStep through heat.c's own
EnergyMonitor, the solution (through the binary viewer) against the oracle, every -ts_type, and the error paths."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import heat_oracle as ho
from p4pdes_b200 import petscbin

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "heat_goldens.json")))


@pytest.fixture(scope="module")
def exe():
    path = os.path.join(ROOT, "oracle", "_ref", "heat_shim_host")
    if os.path.exists("/root/reference/c/ch5/heat.c"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref/heat_shim_host"])
    if not os.path.exists(path):
        pytest.skip("needs the reference tree to compile heat.c")
    return path


def run(exe, argv, check=True):
    p = subprocess.run([exe] + argv.split(), capture_output=True, text=True, timeout=300)
    if check:
        assert p.returncode == 0, p.stderr
    return p.stdout.splitlines(), p


ROUTES = [("", "equals the library's heat-equation kernel"), (" -p4b_recognise_residual 0", "callbacks evaluated on the host")]


@pytest.mark.parametrize("route,says", ROUTES)
def test_goldens_verbatim(exe, route, says):
    lines, p = run(exe, GOLD["heat.test2"]["options"] + route)               # explicit: no preconditioner to name
    assert lines == GOLD["heat.test2"]["lines"]
    assert says in p.stderr                                                  # the route is always said
    lines, _ = run(exe, GOLD["heat.test1"]["options"] + " -pc_type none" + route)    # (PETSc's default ILU is not on the device)
    assert lines == GOLD["heat.test1"]["lines"]


def states(exe, argv, tmp_path):
    t, u = str(tmp_path / "t.dat"), str(tmp_path / "u.dat")
    run(exe, argv + " -ts_monitor binary:%s -ts_monitor_solution binary:%s" % (t, u))
    return np.array(petscbin.read_file(t)), petscbin.read_file(u)


@pytest.mark.parametrize("route", [r for r, _ in ROUTES])
@pytest.mark.parametrize("ts_type,tol", [("rk", 1e-13), ("beuler", 2e-7), ("cn", 2e-7)])
def test_solution_equals_the_oracle(exe, tmp_path, ts_type, tol, route):
    """The states the binary viewer records (c/ch5/MOVIES.md:44) against the NumPy integration of the same system: the
    explicit scheme is the same arithmetic; the implicit ones inherit the stage solves' tolerances (Newton 1e-8 on a
    matrix-free GMRES at 1e-5 against sparse LU)."""
    extra = "" if ts_type == "rk" else " -pc_type none"
    T, U = states(exe, "-da_refine 2 -ts_type %s -ts_max_time 0.01%s%s" % (ts_type, extra, route), tmp_path)
    want_t = []
    mon = lambda k, t, h, w: want_t.append(t)
    if ts_type == "rk":
        u, _, _ = ho.rk3bs(ho.rhs, np.zeros((16, 17)), 0.001, 0.01, monitor=mon)
    else:
        u, _ = ho.theta(np.zeros((16, 17)), 0.001, 0.01, theta=1.0 if ts_type == "beuler" else 0.5, monitor=mon)
    np.testing.assert_allclose(T, want_t, rtol=1e-12, atol=1e-15)
    assert U[-1].shape == (17 * 16,) and np.max(np.abs(u)) > 1e-3
    assert np.max(np.abs(U[-1].reshape(16, 17) - u)) <= tol * np.max(np.abs(u))


def test_energy_monitor_through_tsmonitorset(exe):
    """-ht_monitor (heat.c:51-52,71-73): heat.c's own monitor, registered with TSMonitorSet, asks TSGetDM and TSGetTimeStep
    inside the solve; its line precedes -ts_monitor's (set first).  Same text as the oracle's up to the energy's rounding."""
    lines, _ = run(exe, "-da_refine 2 -ts_type beuler -pc_type none -ts_max_time 0.02 -ts_monitor -ht_monitor")
    _, want = ho.heat(refine=2, ts_type="beuler", tmax=0.02, monitor_energy=True)
    assert len(lines) == len(want) == 1 + 2 * 21
    for a, b in zip(lines, want):
        if "energy" in a:
            assert a.split("nu =")[1] == b.split("nu =")[1] == "   0.2560" and abs(float(a.split()[2])) < 1e-15
        else:
            assert a == b
    # adaptive steps: nu follows the step the integrator is about to take
    lines, _ = run(exe, "-da_refine 1 -ts_type rk -ts_max_time 0.01 -ts_monitor -ht_monitor")
    _, want = ho.heat(refine=1, ts_type="rk", tmax=0.01, monitor_energy=True)
    assert [l.split("nu =")[1] for l in lines if "nu =" in l] == [l.split("nu =")[1] for l in want if "nu =" in l]
    assert [l for l in lines if " TS dt " in l] == GOLD["heat.test2"]["lines"][1:]


def test_default_type_bdf_and_other_options(exe, tmp_path):
    """heat.c:74 sets TSBDF (order 2, adaptive): no golden; it must reach the final time and agree with the other
    integrators to its accuracy.  -ht_D0 is heat.c's own option and reaches the callbacks."""
    T, U = states(exe, "-da_refine 1 -pc_type none -ts_max_time 0.01", tmp_path)
    assert abs(T[-1] - 0.01) < 1e-15 and len(T) >= 4
    ref, _, _ = ho.rk3bs(ho.rhs, np.zeros((8, 9)), 1e-4, 0.01, atol=1e-10, rtol=1e-10)
    assert np.max(np.abs(U[-1].reshape(8, 9) - ref)) <= 2e-2 * np.max(np.abs(ref))
    T2, U2 = states(exe, "-da_refine 1 -ts_type rk -ts_max_time 0.01 -ht_D0 0.25", tmp_path)
    want, _, _ = ho.rk3bs(lambda w: ho.rhs(w, 0.25), np.zeros((8, 9)), 0.001, 0.01)
    assert np.max(np.abs(U2[-1].reshape(8, 9) - want)) <= 1e-13 * np.max(np.abs(want))


@pytest.mark.parametrize("argv,code,msg", [
    ("-da_refine 1 -ts_type beuler", 56, "ILU"),
    ("-da_refine 1 -ts_type beuler -pc_type mg", 56, "-pc_type none"),
    ("-da_refine 1 -ts_type ssp", 56, "-ts_type ssp is not provided"),
    ("-da_refine 1 -ts_type rk -ts_rk_type 5dp", 56, "3bs"),
])
def test_error_paths(exe, argv, code, msg):
    _, p = run(exe, argv, check=False)
    assert p.returncode == code and msg in p.stderr and "PETSC ERROR" in p.stderr


@pytest.fixture(scope="module")
def variants(exe, tmp_path_factory):
    """tests/shim_cases/heat_variants.c against the same shim + stand-in objects."""
    d = tmp_path_factory.mktemp("hv")
    obj, out = str(d / "hv.o"), str(d / "hv")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), "-c",
                           os.path.join(ROOT, "tests", "shim_cases", "heat_variants.c"), "-o", obj])
    o = os.path.join(ROOT, "oracle", "_ref", "obj")
    subprocess.check_call(["g++", obj, os.path.join(o, "petscshim.o"), os.path.join(o, "p4b_standin.o"), "-o", out, "-lm"])
    return out


def test_recognition_of_the_heat_kernel_and_its_limits(variants):
    """The model written differently (and with another diffusivity) is recognised: 4 host evaluations of G in the whole run
    (two to identify D0, one at a generic state, one at the final state).  The model plus a cubic term is not: host
    callbacks throughout, and a different answer.  A term the probes cannot see is caught at the final state: loud error."""
    tail = lambda l: (float(l.split()[7]), float(l.split()[9]), int(l.split("(")[1].split()[0]))
    a, p = run(variants, "-variant 0 -da_refine 2 -pc_type none")
    assert "equals the library's heat-equation kernel (D0 = 1;" in p.stderr and tail(a[-1])[2] == 4
    b, p = run(variants, "-variant 0 -da_refine 2 -pc_type none -p4b_recognise_residual 0")
    assert "callbacks evaluated on the host" in p.stderr and tail(b[-1])[2] > 100
    assert abs(tail(a[-1])[1] - tail(b[-1])[1]) <= 1e-7 * tail(b[-1])[1]            # same answer on both routes
    c, p = run(variants, "-variant 0 -D0 0.3 -da_refine 2 -ts_type rk")
    assert "(D0 = 0.3;" in p.stderr and tail(c[-1])[2] == 4
    d, p = run(variants, "-variant 1 -da_refine 2 -pc_type none")
    assert "callbacks evaluated on the host" in p.stderr and tail(d[-1])[2] > 100
    assert abs(tail(d[-1])[1] - tail(a[-1])[1]) > 1e-6 * tail(a[-1])[1]
    _, p = run(variants, "-variant 2 -da_refine 2 -pc_type none", check=False)
    assert p.returncode == 56 and "not at the final state" in p.stderr and "-p4b_recognise_residual 0" in p.stderr
    e, p = run(variants, "-variant 2 -da_refine 2 -pc_type none -p4b_recognise_residual 0")
    assert tail(e[-1])[0] < -1e-3                                                     # the extra sink removes heat
