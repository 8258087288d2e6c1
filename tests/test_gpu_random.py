"""-fsh_initial_type random (c/ch6/poissonfunctions.c:267-271) and -ptn_noisy_init (c/ch5/pattern.c:159-165,
c/ch8/cluster.sh:56) on the device: the VecSetRandom stream of tests/test_random_stream.py under the drivers."""
import os
import subprocess

import numpy as np
import pytest
import torch

from p4pdes_b200 import pattern as pp
from p4pdes_b200.fish import Context, fish_main

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    return Context()


def test_fish_random_initial_iterate(ctx):
    base = "-fsh_dim 3 -da_refine 4 -pc_type mg -ksp_rtol 1e-12 -ksp_converged_reason"
    a = fish_main(base, ctx, keep_solution=True)
    b = fish_main(base + " -fsh_initial_type random", ctx, keep_solution=True)
    assert b.fnorm0 > 1.5 * a.fnorm0                     # the random interior makes the first residual larger
    assert float((a.u - b.u).abs().max()) < 1e-9         # ... and CG converges to the same discrete solution
    assert "%.3e" % a.errinf == "%.3e" % b.errinf
    # boundary values of g are on the boundary, the stream's first values just inside (natural ordering)
    u0 = ctx.rander48(33 ** 3).cpu().numpy().reshape(33, 33, 33)
    assert 0.0 <= u0.min() and u0.max() < 1.0


def test_pattern_noisy_init_hosts_agree(ctx):
    argv = "-da_refine 4 -ts_monitor -ts_max_time 30 -pc_type mg -ptn_noisy_init 0.15"
    a = pp.pattern_main(argv, ctx)
    b = pp.pattern_main(argv, ctx, native=True)
    q = pp.pattern_main(argv.replace(" -ptn_noisy_init 0.15", ""), ctx)
    assert [l for l in a.lines if "TS dt" in l] == [l for l in b.lines if "TS dt" in l]
    assert a.lines != q.lines                             # the noise changes the adaptive step sequence
    assert float((a.Y - b.Y).abs().max()) < 1e-9
    # the initial state itself: level * stream under the patch, u = n_u + 1 - 2 v (pattern.c:159-175)
    m = 3 * 2 ** 4
    Y = ctx.empty(2 * m * m)
    ctx.pattern_initial_state_noisy(m, m, 2.5, 0.15, Y)
    Y0 = ctx.empty(2 * m * m)
    ctx.pattern_initial_state(m, m, 2.5, Y0)
    r = ctx.rander48(2 * m * m).cpu().numpy().reshape(-1, 2)
    y, y0 = Y.cpu().numpy().reshape(-1, 2), Y0.cpu().numpy().reshape(-1, 2)
    np.testing.assert_allclose(y[:, 1], y0[:, 1] + 0.15 * r[:, 1], rtol=0, atol=1e-15)
    np.testing.assert_allclose(y[:, 0], 0.15 * r[:, 0] + 1.0 - 2.0 * y[:, 1], rtol=0, atol=1e-15)


def test_unchanged_pattern_c_with_the_cluster_script_noise():
    """c/ch8/cluster.sh:56 on a small grid: the unchanged driver under the shim, device residuals, equal to the hosts."""
    exe = os.path.join(ROOT, "p4pdes_b200", "bin", "pattern")
    if not os.path.exists(exe):
        pytest.skip("unchanged driver not built")
    argv = "-da_refine 4 -ts_monitor -ts_max_time 30 -pc_type mg -mg_levels_pc_type jacobi -ptn_noisy_init 0.15"
    p = subprocess.run([exe] + argv.split(), capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    want = pp.pattern_main(argv.replace(" -mg_levels_pc_type jacobi", ""), Context())
    got = [l for l in p.stdout.splitlines() if "TS dt" in l]
    assert got and got == [l for l in want.lines if "TS dt" in l]


def test_ts_binary_monitors_on_device(ctx, tmp_path):
    """c/ch5/MOVIES.md:44 on the device: times and states of every step as PETSc binary records (native host and the
    unchanged pattern.c under the shim write the same files)."""
    from p4pdes_b200 import petscbin
    t, u = str(tmp_path / "t.dat"), str(tmp_path / "u.dat")
    argv = "-da_refine 3 -ts_max_time 30 -pc_type mg"
    ref = pp.pattern_main(argv + " -ts_monitor", ctx, native=True)
    rep = pp.pattern_main(argv + " -ts_monitor binary:%s -ts_monitor_solution binary:%s" % (t, u), ctx, native=True)
    times, states = petscbin.read_file(t), petscbin.read_file(u)
    nlines = len([l for l in ref.lines if " TS dt " in l])
    assert len(times) == len(states) == nlines and not any(" TS dt " in l for l in rep.lines)
    assert abs(times[-1] - 30.0) < 1e-12 and times[0] == 0.0
    assert float(np.max(np.abs(states[-1] - rep.Y.cpu().numpy()))) == 0.0
    exe = os.path.join(ROOT, "p4pdes_b200", "bin", "pattern")
    if os.path.exists(exe):
        t2, u2 = str(tmp_path / "t2.dat"), str(tmp_path / "u2.dat")
        p = subprocess.run([exe] + (argv + " -mg_levels_pc_type jacobi -ts_monitor binary:%s -ts_monitor_solution binary:%s"
                                    % (t2, u2)).split(), capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr
        s2 = petscbin.read_file(u2)
        assert len(s2) == len(states) and float(np.max(np.abs(s2[-1] - states[-1]))) < 1e-9
        np.testing.assert_allclose(petscbin.read_file(t2), times, rtol=1e-12)
