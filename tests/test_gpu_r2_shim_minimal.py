"""The reference's UNCHANGED c/ch7/minimal.c on the device: p4pdes_b200/bin/minimal = minimal.c + poissonfunctions.c
compiled against include/petsc.h, linked with the shim and libp4b200.so (p4pdes_b200/build.py:DRIVERS; the prebuilt
binary travels to the GPU box).  First run on a B200 in round 2 (profiles/r02_pending.md) and promoted to the `gpu` marker.  The same binary over the host stand-in is checked on the CPU (tests/test_shim_minimal_cpu.py)."""
import json
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "p4pdes_b200", "bin", "minimal")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "minimal_goldens.json")))
EXTRA = " -pc_type mg -mg_levels_pc_type jacobi"

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device"),
              pytest.mark.skipif(not os.path.exists(EXE), reason="p4pdes_b200/bin/minimal was not built")]


def run(argv):
    p = subprocess.run([EXE] + argv.split(), capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr
    return p.stdout.splitlines()


def test_golden_test4_verbatim_on_device():
    g = GOLD["minimal.test4"]
    for extra in (" -pc_type none", EXTRA):
        assert run(g["options"] + extra) == g["lines"]


def test_golden_test1_on_device():
    g = GOLD["minimal.test1"]
    lines = run(g["options"] + " -pc_type none")
    assert len(lines) == len(g["lines"]) and lines[0] == g["lines"][0] and lines[-2:] == g["lines"][-2:]
    got = [float(l.split()[-1]) for l in lines[1:5]]
    want = [float(l.split()[-1]) for l in g["lines"][1:5]]
    np.testing.assert_allclose(got, want, rtol=1e-2)


def test_golden_test3_monitor_on_device():
    g = GOLD["minimal.test3"]
    lines = run(g["options"].replace("-snes_mf_operator", "-snes_fd_color") + " -mg_levels_pc_type jacobi")
    assert len(lines) == len(g["lines"])
    for a, b in zip(lines, g["lines"]):
        if "area" in b:
            fa, fb = [float(x) for x in re.findall(r"[0-9.]+", a)], [float(x) for x in re.findall(r"[0-9.]+", b)]
            assert a[:a.index("area")] == b[:b.index("area")]
            np.testing.assert_allclose(fa, fb, rtol=0, atol=2e-7)
        else:
            assert a == b


def test_cluster_configuration_through_the_unchanged_driver():
    """c/ch8/cluster.sh:70 (BASELINE config 4): ./minimal -da_grid_x 33 -da_grid_y 33 -snes_grid_sequence 6 -snes_fd_color
    -pc_type mg.  minimal.c's FormFunctionLocal is recognised as the library's kernel (2 probes per grid + 1), so the solve
    is device-resident; with recognition off the same run evaluates the host callback nine times per level Jacobian."""
    lines = run("-da_grid_x 33 -da_grid_y 33 -snes_grid_sequence 6 -snes_fd_color -snes_converged_reason -log_view" + EXTRA)
    assert all("Nonlinear solve converged due to CONVERGED_" in l for l in lines[:7])   # the last stage stops on SNORM
    m = re.fullmatch(r"done on 2049 x 2049 grid and problem catenoid:  error \|u-uexact\|_inf = (\S+)", lines[7])
    assert m and float(m.group(1)) < 1e-7
    assert "SNES newtonls: residual recognised as the library's kernel: evaluated on the device" in lines
    t_dev = float(re.search(r"SNESSolve (\S+)", "\n".join(lines)).group(1))
    print("unchanged minimal.c, 2049^2, device residual: SNESSolve %.3f s" % t_dev)


def test_both_routes_agree_on_device():
    argv = "-snes_fd_color -snes_converged_reason -snes_grid_sequence 3 -da_grid_x 9 -da_grid_y 9 -log_view" + EXTRA
    a = run(argv)
    b = run(argv + " -p4b_recognise_residual 0")
    assert a[:5] == b[:5]
    assert "SNES newtonls: residual evaluated by the host callback" in b
