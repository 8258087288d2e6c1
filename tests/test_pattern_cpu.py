"""pattern.c implicit runs without a GPU: the solver oracle against the reference's golden (c/ch5/output/pattern.test2,
command c/ch5/makefile:52-53) and the control flow of the device driver p4pdes_b200/pattern.py on the NumPy stand-in."""
import os

import numpy as np
import pytest

from oracle import pattern_solver_oracle as po
from p4pdes_b200 import pattern as pp
from tests.fake_ops import FakeOps

GOLDEN_TEST2 = """running on 12 x 12 grid with square cells of side h = 0.208333 ...
0 TS dt 1. time 0.
      Linear solve converged due to CONVERGED_RTOL iterations 3
    Nonlinear solve converged due to CONVERGED_FNORM_RELATIVE iterations 1
1 TS dt 1. time 1.""".split("\n")
TEST2 = ("-da_refine 2 -ts_monitor -ts_dt 1 -ts_max_time 1 -ts_type beuler -pc_type mg -snes_converged_reason "
         "-ksp_converged_reason -snes_rtol 1.0e-1 -ptn_no_rhsjacobian")


def test_golden_file_is_what_we_pin():
    ref = "/root/reference/c/ch5/output/pattern.test2"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present")
    assert open(ref).read().rstrip("\n").split("\n") == GOLDEN_TEST2


def test_oracle_reproduces_pattern_test2():
    r = po.pattern_beuler(grid=3, refine=2, dt=1.0, tmax=1.0, rhsjac=False, snes_rtol=1e-1)
    assert r.lines == [GOLDEN_TEST2[0], GOLDEN_TEST2[1], GOLDEN_TEST2[4]]
    (t, dt, nr), = r.steps
    assert (t, dt, nr.its, nr.ksp_its, nr.reason) == (1.0, 1.0, 1, [3], "CONVERGED_FNORM_RELATIVE")


def test_driver_prints_pattern_test2_verbatim():
    r = pp.pattern_main(TEST2, FakeOps())
    assert r.lines == GOLDEN_TEST2


@pytest.mark.parametrize("argv,okw", [
    ("-da_grid_x 4 -da_grid_y 4 -da_refine 3 -ts_type beuler -ts_dt 5 -ts_max_time 12 -pc_type mg",
     dict(grid=4, refine=3, dt=5.0, tmax=12.0)),
    ("-da_refine 3 -ts_type beuler -ts_dt 2 -ts_max_time 4 -pc_type mg -ptn_no_rhsjacobian -snes_rtol 1e-6",
     dict(grid=3, refine=3, dt=2.0, tmax=4.0, rhsjac=False, snes_rtol=1e-6)),
    ("-da_refine 2 -ts_type beuler -ts_dt 5 -ts_max_time 5 -pc_type none", dict(grid=3, refine=2, dt=5.0, tmax=5.0, pc="none")),
    ("-da_grid_x 4 -da_grid_y 4 -da_refine 4 -ts_type beuler -ts_dt 5 -ts_max_time 5 -pc_type mg -p4b_mg_rscale 0.25",
     dict(grid=4, refine=4, dt=5.0, tmax=5.0, rscale=0.25)),
])
def test_driver_matches_oracle(argv, okw):
    r = pp.pattern_main(argv, FakeOps())
    o = po.pattern_beuler(**okw)
    assert [(t, dt) for t, dt, _ in r.steps] == [(t, dt) for t, dt, _ in o.steps]        # incl. the matched final step
    assert [s[2].its for s in r.steps] == [s[2].its for s in o.steps]
    assert [s[2].ksp_its for s in r.steps] == [s[2].ksp_its for s in o.steps]
    assert np.max(np.abs(r.Y.a.reshape(o.Y.shape) - o.Y)) <= 1e-12


@pytest.mark.parametrize("argv,msg", [
    ("-da_refine 2", "beuler"),                                           # the reference's default is arkimex: not built
    ("-ts_type beuler -pc_type ilu", "sequential"),
    ("-ts_type beuler -ptn_noisy_init 0.2", "not provided"),
    ("-ts_type beuler -da_grid_x 4 -da_grid_y 6", "requires mx == my"),   # pattern.c:89
])
def test_error_paths(argv, msg):
    with pytest.raises(ValueError, match=msg):
        pp.pattern_main(argv, FakeOps())


def test_periodic_interpolation_is_partition_of_unity_and_restriction_its_transpose():
    P = po.interpolation(6, 4)
    assert P.shape == (2 * 12 * 8, 2 * 6 * 4)
    np.testing.assert_allclose(P @ np.ones(P.shape[1]), 1.0)
    np.testing.assert_allclose(P.T @ np.ones(P.shape[0]), 4.0)            # every coarse node collects weight 4 in 2-D


def test_averaging_restriction_gives_mesh_independent_krylov_counts():
    """pattern.c's equations carry no cell-volume factor, so [PETSc]'s R = P^T over-weights the coarse correction 4x and
    the GMRES count grows with resolution; the averaging restriction (-p4b_mg_rscale 0.25) keeps it flat."""
    petsc = [max(po.pattern_beuler(grid=4, refine=r, dt=5.0, tmax=5.0).steps[0][2].ksp_its) for r in (5, 6)]
    avg = [max(po.pattern_beuler(grid=4, refine=r, dt=5.0, tmax=5.0, rscale=0.25).steps[0][2].ksp_its) for r in (5, 6)]
    assert avg[1] <= avg[0] <= 8 and petsc[1] >= 2 * avg[1]
