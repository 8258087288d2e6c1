"""The reference's UNCHANGED c/ch5/heat.c on the device: p4pdes_b200/bin/heat = heat.c compiled against include/petsc.h,
linked with the shim and libp4b200.so (p4pdes_b200/build.py:DRIVERS; the prebuilt binary travels to the GPU box).
heat.c's FormRHSFunctionLocal is recognised as the library's heat kernel (probed, re-verified at the final state), so its run
is device-resident (p4b_heat_solve: G and the matrix-free stage operator are kernels, the integrators [PETSc] TSRK 3bs,
TSTHETA, TSBDF and all vector algebra on the device); with -p4b_recognise_residual 0 G stays heat.c's host callback
(p4b_ts_solve_callbacks).  Both routes, and the kernels against the oracle.  The same binary over the host stand-in:
tests/test_shim_heat_cpu.py."""
import json
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import heat_oracle as ho
from p4pdes_b200 import petscbin

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "p4pdes_b200", "bin", "heat")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "heat_goldens.json")))

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device"),
              pytest.mark.skipif(not os.path.exists(EXE), reason="p4pdes_b200/bin/heat was not built")]


def run(argv):
    p = subprocess.run([EXE] + argv.split(), capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    return p.stdout.splitlines(), p


ROUTES = [("", "equals the library's heat-equation kernel"), (" -p4b_recognise_residual 0", "callbacks evaluated on the host")]


@pytest.mark.parametrize("mx,my,D0", [(9, 8, 1.0), (17, 16, 0.3), (130, 37, 2.0)])
def test_heat_kernels_against_the_oracle(mx, my, D0):
    """p4b_heat_rhs = FormRHSFunctionLocal (heat.c:141-163), p4b_heat_jac_apply = shift I - the rows of
    FormRHSJacobianLocal (:166-208)."""
    from p4pdes_b200.fish import Context
    ctx = Context()
    rng = np.random.default_rng(11)
    u, x = rng.standard_normal((my, mx)), rng.standard_normal(mx * my)
    du, dG = ctx.from_host(u.ravel()), ctx.empty(mx * my)
    ctx.heat_rhs(mx, my, D0, du, dG)
    want = ho.rhs(u, D0).ravel()
    np.testing.assert_allclose(ctx.to_host(dG), want, rtol=1e-13, atol=1e-12 * np.max(np.abs(want)))
    ctx.heat_jac_apply(mx, my, D0, 7.5, ctx.from_host(x), dG)
    want = 7.5 * x - ho.jacobian(mx, my, D0) @ x
    np.testing.assert_allclose(ctx.to_host(dG), want, rtol=1e-13, atol=1e-12 * np.max(np.abs(want)))


@pytest.mark.parametrize("route,says", ROUTES)
def test_goldens_verbatim_on_device(route, says):
    lines, p = run(GOLD["heat.test2"]["options"] + route)
    assert lines == GOLD["heat.test2"]["lines"]                 # adaptive RK3bs: every digit of the step sequence
    assert says in p.stderr
    lines, _ = run(GOLD["heat.test1"]["options"] + " -pc_type none" + route)
    assert lines == GOLD["heat.test1"]["lines"]


@pytest.mark.parametrize("route", [r for r, _ in ROUTES])
@pytest.mark.parametrize("ts_type,tol", [("rk", 1e-12), ("beuler", 2e-7), ("cn", 2e-7)])
def test_solution_equals_the_oracle_on_device(tmp_path, ts_type, tol, route):
    t, u = str(tmp_path / "t.dat"), str(tmp_path / "u.dat")
    extra = ("" if ts_type == "rk" else " -pc_type none") + route
    run("-da_refine 2 -ts_type %s -ts_max_time 0.01%s -ts_monitor binary:%s -ts_monitor_solution binary:%s" % (ts_type, extra, t, u))
    T, U = np.array(petscbin.read_file(t)), petscbin.read_file(u)
    want_t = []
    mon = lambda k, tt, h, w: want_t.append(tt)
    if ts_type == "rk":
        ref, _, _ = ho.rk3bs(ho.rhs, np.zeros((16, 17)), 0.001, 0.01, monitor=mon)
    else:
        ref, _ = ho.theta(np.zeros((16, 17)), 0.001, 0.01, theta=1.0 if ts_type == "beuler" else 0.5, monitor=mon)
    np.testing.assert_allclose(T, want_t, rtol=1e-10, atol=1e-14)
    assert np.max(np.abs(U[-1].reshape(16, 17) - ref)) <= tol * np.max(np.abs(ref))


def test_energy_monitor_and_default_bdf_on_device(tmp_path):
    lines, _ = run("-da_refine 2 -ts_type beuler -pc_type none -ts_max_time 0.02 -ts_monitor -ht_monitor")
    assert len(lines) == 1 + 2 * 21
    e = [l for l in lines if "energy" in l]
    assert all(l.split("nu =")[1] == "   0.2560" and abs(float(l.split()[2])) < 1e-15 for l in e)
    t, u = str(tmp_path / "t.dat"), str(tmp_path / "u.dat")
    run("-da_refine 1 -pc_type none -ts_max_time 0.01 -ts_monitor binary:%s -ts_monitor_solution binary:%s" % (t, u))
    T, U = np.array(petscbin.read_file(t)), petscbin.read_file(u)
    ref, _, _ = ho.rk3bs(ho.rhs, np.zeros((8, 9)), 1e-4, 0.01, atol=1e-10, rtol=1e-10)
    assert abs(T[-1] - 0.01) < 1e-15 and np.max(np.abs(U[-1].reshape(8, 9) - ref)) <= 2e-2 * np.max(np.abs(ref))


def test_a_fine_grid_runs_device_resident():
    """513 x 512 nodes, explicit RK3bs: the step is stability-limited (h^2), every stage is one kernel + a few axpys; the
    host sees the state twice (start, end).  Energy stays at rounding level (heat.c's conservation claim)."""
    lines, p = run("-da_refine 7 -ts_type rk -ts_max_time 2e-5 -ts_monitor")
    assert lines[0] == "solving on 513 x 512 grid for t0=0. to tf=2e-05 ..." and lines[-1].endswith("time 2e-05")
    assert "time stepping on the device" in p.stderr and len(lines) >= 4
