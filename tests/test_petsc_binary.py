"""PETSc binary viewers on the TS monitors (c/ch5/MOVIES.md:44: `-ts_monitor binary:t.dat -ts_monitor_solution binary:u.dat`;
read by c/ch5/plotTS.py:44-46 through PetscBinaryIO): the unchanged pattern.c under the shim writes them (CPU: host
stand-in), p4pdes_b200/petscbin.py reads them back the way PetscBinaryIO.readBinaryFile does."""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import minimal_pattern_oracle as mpo
from p4pdes_b200 import petscbin

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_record_layout_is_big_endian_classid_first(tmp_path):
    f = tmp_path / "x.dat"
    with open(f, "wb") as fh:
        petscbin.write_real(fh, 2.5)
        petscbin.write_vec(fh, [1.0, -2.0, 3.5])
    raw = f.read_bytes()
    assert raw[:4] == struct.pack(">i", 1211213) and raw[4:12] == struct.pack(">d", 2.5)
    assert raw[12:20] == struct.pack(">ii", 1211214, 3) and raw[20:] == struct.pack(">3d", 1.0, -2.0, 3.5)
    back = petscbin.read_file(f)
    assert back[0] == 2.5 and np.array_equal(back[1], [1.0, -2.0, 3.5])


def test_unchanged_pattern_c_writes_the_movie_files(tmp_path):
    exe = os.path.join(ROOT, "oracle", "_ref", "pattern_shim_host")
    if not os.path.exists(exe):
        pytest.skip("reference tree absent: the unchanged driver was not built")
    t, u = str(tmp_path / "t.dat"), str(tmp_path / "u.dat")
    base = "-da_refine 3 -ts_max_time 30 -pc_type mg -mg_levels_pc_type jacobi"
    p = subprocess.run([exe] + (base + " -ts_monitor").split(), capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    want_t = [float(l.split()[-1].rstrip(".")) for l in p.stdout.splitlines() if " TS dt " in l]
    p = subprocess.run([exe] + (base + " -ts_monitor binary:%s -ts_monitor_solution binary:%s" % (t, u)).split(),
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert not any(" TS dt " in l for l in p.stdout.splitlines())          # a binary viewer prints nothing
    times, states = petscbin.read_file(t), petscbin.read_file(u)
    assert len(times) == len(states) == len(want_t) and len(times) >= 3
    np.testing.assert_allclose(times, want_t, rtol=1e-5)                  # the monitor lines print 6 digits
    m = 3 * 2 ** 3
    assert all(s.shape == (2 * m * m,) for s in states)
    np.testing.assert_allclose(states[0], mpo.pattern_initial_state(m, m, 2.5).ravel(), atol=1e-15)      # step 0 = InitialState
    assert np.abs(states[-1] - states[0]).max() > 1e-3                    # ... and the trajectory moves
    # plotTS.py's reshaping: U = array(readBinaryFile(ufile)).T, one column per frame
    U = np.array(states).transpose()
    assert U.shape == (2 * m * m, len(times))
