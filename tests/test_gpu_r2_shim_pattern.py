"""The reference's UNCHANGED c/ch5/pattern.c on the device: p4pdes_b200/bin/pattern = pattern.c compiled against
include/petsc.h, linked with the shim and libp4b200.so (p4pdes_b200/build.py:DRIVERS; the prebuilt binary travels to the
GPU box).  First run on a B200 in round 2 (profiles/r02_pending.md) and promoted to the `gpu` marker.  The same binary
over the host stand-in is checked on the CPU (tests/test_shim_pattern_cpu.py)."""
import json
import os
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "p4pdes_b200", "bin", "pattern")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "pattern_goldens.json")))
MG = " -pc_type mg -mg_levels_pc_type jacobi"

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device"),
              pytest.mark.skipif(not os.path.exists(EXE), reason="p4pdes_b200/bin/pattern was not built")]


def run(argv):
    p = subprocess.run([EXE] + argv.split(), capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr
    return p.stdout.splitlines()


@pytest.mark.parametrize("name,extra", [("pattern.test1", " -pc_type none"), ("pattern.test1", MG),
                                        ("pattern.test2", " -mg_levels_pc_type jacobi"), ("pattern.test3", " -pc_type none"),
                                        ("pattern.test4", MG), ("pattern.test5", MG)])
def test_goldens_verbatim_on_device(name, extra):
    g = GOLD[name]
    assert run(g["options"] + extra) == g["lines"]


def test_baseline_configuration_through_the_unchanged_driver():
    """2048 x 2048, two implicit steps (identification and verification of the callbacks run once, on the
    host, at this size: ~6 sweeps over 8.4 M unknowns)."""
    lines = run("-da_grid_x 8 -da_grid_y 8 -da_refine 8 -ts_type beuler -ts_dt 5 -ts_max_time 10 -ts_monitor "
                "-snes_converged_reason -p4b_mg_rscale 0.25" + MG)
    assert lines[0] == "running on 2048 x 2048 grid with square cells of side h = 0.001221 ..."
    assert lines[-1] == "2 TS dt 5. time 10." and sum("CONVERGED_FNORM_RELATIVE" in l for l in lines) == 2
