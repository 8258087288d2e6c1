"""The VecSetRandom stream ([PETSc] PetscRandom of the default type rander48; c/ch5/pattern.c:163 -ptn_noisy_init,
c/ch6/poissonfunctions.c:268-270 -fsh_initial_type random).  PETSc's rander48 is the 48-bit linear congruential
generator of drand48 seeded with 0x12345678; PETSc cannot be installed here and no golden of the reference uses a random
vector, so equality with PETSc itself is unpinned -- what is pinned is the recurrence, on glibc's srand48/drand48."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from p4pdes_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stream(n, seed=0x12345678):
    lib = L.load()
    state = C.c_ulonglong(lib.p4b_rander48_seed(seed))
    out = np.empty(n)
    L.check(lib.p4b_rander48_fill(C.byref(state), n, out.ctypes.data))
    return out, state.value


def test_stream_is_the_drand48_recurrence():
    libc = C.CDLL("libc.so.6")
    libc.drand48.restype = C.c_double
    for seed in (0x12345678, 0x12345678 + 76543, 1):
        libc.srand48(C.c_long(seed))
        want = np.array([libc.drand48() for _ in range(1000)])
        got, _ = stream(1000, seed)
        assert np.array_equal(got, want)
    a, st = stream(10)
    b, _ = stream(25)
    state = C.c_ulonglong(st)
    rest = np.empty(15)
    L.check(L.load().p4b_rander48_fill(C.byref(state), 15, rest.ctypes.data))
    assert np.array_equal(np.concatenate([a, rest]), b)          # the state carries over between calls
    assert 0.0 <= b.min() and b.max() < 1.0


def run_host(exe, argv):
    p = subprocess.run([exe] + argv.split(), capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    return p.stdout.splitlines()


def test_unchanged_pattern_c_with_noisy_init_runs_on_the_host_stand_in():
    """c/ch8/cluster.sh:56 passes -ptn_noisy_init 0.15; through the shim VecSetRandom now fills the Vec and pattern.c's own
    InitialState adds the patch.  The noise changes the trajectory (different step sequence from the noiseless run)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "pattern_shim_host")
    if not os.path.exists(exe):
        pytest.skip("reference tree absent: the unchanged driver was not built")
    base = "-da_refine 3 -ts_monitor -ts_max_time 20 -pc_type mg -mg_levels_pc_type jacobi"
    quiet = run_host(exe, base)
    noisy = run_host(exe, base + " -ptn_noisy_init 0.15")
    assert any("TS dt" in l for l in noisy) and noisy != quiet
    assert noisy == run_host(exe, base + " -ptn_noisy_init 0.15")      # a fixed stream: runs are repeatable


def test_unchanged_fish_c_with_random_initial_iterate():
    """-fsh_initial_type random: CG from a random iterate reaches the same discrete solution (same error norms)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "fish_shim_host")
    if not os.path.exists(exe):
        pytest.skip("reference tree absent: the unchanged driver was not built")
    base = "-fsh_dim 2 -da_refine 4 -pc_type mg -mg_levels_pc_type jacobi -ksp_rtol 1e-12"
    zeros = run_host(exe, base)
    rand = run_host(exe, base + " -fsh_initial_type random")
    assert zeros[-1] == rand[-1] and "error |u-uexact|_inf" in rand[-1]
