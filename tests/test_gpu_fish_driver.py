"""The reference's UNCHANGED c/ch6/fish.c + poissonfunctions.c, compiled against the PETSc-shaped shim
(include/petsc.h, libpetsc_p4b200.so) and run on the GPU with the reference's own command lines
(c/ch6/makefile:11-33, c/ch8/cluster.sh:63), output compared as text the way c/testit.sh does."""
import os
import subprocess

import pytest

from oracle import fish_oracle as fo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FISH = os.path.join(ROOT, "p4pdes_b200", "bin", "fish")
JAC = " -mg_levels_pc_type jacobi"     # the device smoother (north star); PETSc's default SOR is sequential


def fish(opts, expect_rc=0):
    if not os.path.exists(FISH):
        from p4pdes_b200 import build as b
        b.build_drivers()
    assert os.path.exists(FISH), "fish driver was not built (needs /root/reference at build time)"
    p = subprocess.run([FISH] + opts.split(), capture_output=True, text=True, timeout=600)
    assert p.returncode == expect_rc, (p.returncode, p.stdout, p.stderr)
    return p.stdout, p.stderr


def test_golden_test1_lines(goldens):
    # c/ch6/makefile:12 with the Jacobi smoother: every line except the iteration count is solver independent
    g = goldens["fish.test1"]
    out, _ = fish(g["options"] + JAC)
    lines = out.splitlines()
    assert lines[0] == "  0 SNES Function norm %s" % g["snes_fnorm0"]
    assert lines[1].startswith("    Linear solve converged due to CONVERGED_RTOL iterations ")
    assert lines[2] == "  1 SNES Function norm < 1.e-11"
    assert lines[3] == "problem %s on %s grid:" % (g["problem"], g["gridstr"])
    assert lines[4] == "  error |u-uexact|_inf = %s, |u-uexact|_h = %s" % (g["errinf"], g["err2h"])
    assert len(lines) == 5


@pytest.mark.parametrize("name,opts", [
    ("fish.test6", "-fsh_dim 3 -da_refine 2 -fsh_problem manupoly -ksp_converged_reason -fsh_cx 0.01 -fsh_cy 2 -fsh_cz 100"),
    ("fish.test7", "-fsh_dim 3 -fsh_problem manupoly -ksp_converged_reason -da_refine 2"),
])
def test_golden_error_lines(goldens, name, opts):
    g = goldens[name]
    out, _ = fish(opts + " -pc_type mg -ksp_rtol 1.0e-12" + JAC)
    lines = out.splitlines()
    assert lines[-2] == "problem %s on %s grid:" % (g["problem"], g["gridstr"])
    assert lines[-1] == "  error |u-uexact|_inf = %s, |u-uexact|_h = %s" % (g["errinf"], g["err2h"])


@pytest.mark.parametrize("opts,okw", [
    ("-fsh_dim 2 -da_refine 6 -pc_type mg -ksp_rtol 1e-10 -ksp_converged_reason", dict(dim=2, refine=6, rtol=1e-10)),
    ("-fsh_dim 3 -da_refine 4 -pc_type mg -ksp_rtol 1e-10 -ksp_converged_reason", dict(dim=3, refine=4, rtol=1e-10)),
    # c/ch8/cluster.sh:63 at a size the oracle can check
    ("-fsh_dim 3 -da_refine 5 -pc_mg_levels 4 -pc_type mg -snes_type ksponly -ksp_converged_reason",
     dict(dim=3, refine=5, mg=dict(levels=4))),
    ("-fsh_dim 2 -da_refine 4 -pc_type mg -pc_mg_cycle_type w -mg_levels_ksp_type richardson -mg_levels_ksp_max_it 1 "
     "-ksp_converged_reason", dict(dim=2, refine=4, mg=dict(cycle="w", smoother_ksp="richardson", smoother_its=1))),
    ("-fsh_dim 2 -fsh_initial_gonboundary false -da_refine 3 -pc_type mg -ksp_converged_reason",
     dict(dim=2, refine=3, gonboundary=False)),
])
def test_output_matches_oracle(opts, okw):
    okw = dict(okw)
    want = fo.fish(mg=fo.MGOptions(**okw.pop("mg", {})), **okw)
    out, _ = fish(opts + JAC + " -ksp_monitor")
    lines = out.splitlines()
    gs = {1: "%d point 1D", 2: "%d x %d point 2D", 3: "%d x %d x %d point 3D"}[want.grid.dim] % want.grid.m[:want.grid.dim]
    assert "    Linear solve converged due to CONVERGED_RTOL iterations %d" % want.its in lines
    assert lines[-2] == "problem manuexp on %s grid:" % gs
    assert lines[-1] == "  error |u-uexact|_inf = %.3e, |u-uexact|_h = %.3e" % (want.errinf, want.err2h)
    hist = [float(l.split()[-1]) for l in lines if "KSP Residual norm" in l]
    assert len(hist) == len(want.history)
    for a, b in zip(hist, want.history):
        assert abs(a - b) <= 1e-10 * b


def test_reference_error_paths():
    # fish.c:188-193,214-215 checks run unchanged; the shim refuses what only sequential PETSc code provides
    _, err = fish("-fsh_cx -1 -fsh_problem manupoly -pc_type mg" + JAC, expect_rc=2)
    assert "positivity required" in err
    _, err = fish("-fsh_cx 2 -pc_type mg" + JAC, expect_rc=3)
    assert "cx=cy=cz=1 required" in err
    _, err = fish("-fsh_dim 4 -pc_type mg" + JAC, expect_rc=1)
    assert "invalid dim" in err
    _, err = fish("-fsh_dim 2 -da_refine 2", expect_rc=56)
    assert "ILU" in err
    _, err = fish("-fsh_dim 2 -da_refine 2 -pc_type mg", expect_rc=56)
    assert "SOR" in err


def test_log_view_reports_solve_time():
    out, _ = fish("-fsh_dim 3 -da_refine 5 -pc_type mg -pc_mg_levels 4 -ksp_rtol 1e-10 -log_view" + JAC)
    assert "KSPSolve" in out and "SNESSolve" in out
