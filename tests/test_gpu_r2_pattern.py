"""pattern.c on the device: First run on a B200 in round 2 (profiles/r02_pending.md) and promoted to the `gpu` marker.
Everything they exercise is also CPU-checked (tests/test_pattern_cpu.py,
tests/test_native_nk_cpu.py: the same driver / the same C++ host logic on NumPy / plain-C++ operations reproduce the
goldens verbatim); here is the device instantiation: p4b_vec_wrms2, the ARKIMEX and Crank-Nicolson runs of
the Python host, and p4b_pattern_solve."""
import numpy as np
import pytest
import torch

from p4pdes_b200 import pattern as pp
from p4pdes_b200.fish import Context
from tests.test_pattern_cpu import (GOLDEN_TEST1, GOLDEN_TEST2, GOLDEN_TEST3, GOLDEN_TEST4, GOLDEN_TEST5, TEST1, TEST2,
                                    TEST3, TEST4, TEST5)

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


@pytest.fixture(scope="module")
def ctx():
    return Context()


def dev(ctx, a):
    return ctx.from_host(np.asarray(a, dtype=np.float64))


def test_weighted_error_norm_kernel(ctx):
    rng = np.random.default_rng(9)
    for n in (1, 1000, 2 * 2048 * 2048 + 3):
        x, y = rng.standard_normal(n), rng.standard_normal(n)
        want = np.sum(((x - y) / (1e-4 + 1e-3 * np.maximum(np.abs(x), np.abs(y)))) ** 2)
        got = ctx.wrms2(dev(ctx, x), dev(ctx, y), 1e-4, 1e-3)
        assert abs(got - want) <= 1e-12 * want


def _ts_numbers(lines):
    return [(float(w[3]), float(w[5])) for w in (l.split() for l in lines) if len(w) == 6 and w[1] == "TS"]


@pytest.mark.parametrize("argv,golden", [(TEST1, GOLDEN_TEST1), (TEST4, GOLDEN_TEST4)])
def test_arkimex_goldens_on_device(ctx, argv, golden):
    """pattern.c's default TS type with adaptive steps: c/ch5/output/pattern.test1 / test4.  The step sizes are compared as
    numbers (6 printed digits; a value that sits on a rounding boundary may print differently after 1e-10-level
    differences in the stage solves); everything else verbatim."""
    r = pp.pattern_main(argv, ctx)
    assert len(r.lines) == len(golden)
    np.testing.assert_allclose(_ts_numbers(r.lines), _ts_numbers(golden), rtol=3e-6)
    assert [l for l in r.lines if " TS " not in l] == [l for l in golden if " TS " not in l]
    if r.lines != golden:
        print("printed digits differ:", [(a, b) for a, b in zip(r.lines, golden) if a != b])


def test_golden_pattern_test3_crank_nicolson_on_device(ctx):
    assert pp.pattern_main(TEST3, ctx).lines == GOLDEN_TEST3                 # c/ch5/output/pattern.test3


@pytest.mark.parametrize("argv,golden", [(TEST1, GOLDEN_TEST1), (TEST2, GOLDEN_TEST2), (TEST3, GOLDEN_TEST3),
                                         (TEST4, GOLDEN_TEST4), (TEST5 + " -pc_type mg", GOLDEN_TEST5)])
def test_native_time_stepper_on_device(ctx, argv, golden):
    """p4b_pattern_solve (host logic in C++ inside the library, csrc/ts_solver.hpp; CPU-checked against the same goldens
    in tests/test_native_nk_cpu.py) on the device: step sizes as numbers, everything else verbatim."""
    r = pp.pattern_main(argv, ctx, native=True)
    assert len(r.lines) == len(golden)
    np.testing.assert_allclose(_ts_numbers(r.lines), _ts_numbers(golden), rtol=3e-6)
    assert [l for l in r.lines if " TS " not in l] == [l for l in golden if " TS " not in l]


def test_native_time_stepper_equals_the_python_host(ctx):
    argv = "-da_grid_x 4 -da_grid_y 4 -da_refine 5 -ts_type beuler -ts_dt 5 -ts_max_time 10 -pc_type mg -p4b_mg_rscale 0.25"
    a = pp.pattern_main(argv, ctx)
    b = pp.pattern_main(argv, ctx, native=True)
    assert [(t, dt) for t, dt, _ in a.steps] == [(t, dt) for t, dt, _ in b.steps]
    assert [s[2].its for s in a.steps] == [s[2] for s in b.steps]
    assert float((a.Y - b.Y).abs().max()) <= 1e-10


def test_c_example_hosts_print_the_goldens():
    """examples/pattern_native.c / minimal_native.c: a C program, one C-ABI call per run."""
    import subprocess
    from p4pdes_b200 import build as p4build
    exes = p4build.build_examples()
    for argv, golden in ((TEST1, GOLDEN_TEST1), (TEST2, GOLDEN_TEST2), (TEST3, GOLDEN_TEST3), (TEST4, GOLDEN_TEST4)):
        out = subprocess.run([exes["pattern_native"], *argv.split()], capture_output=True, text=True, timeout=300, check=True)
        lines = out.stdout.rstrip("\n").split("\n")
        assert len(lines) == len(golden)
        np.testing.assert_allclose(_ts_numbers(lines), _ts_numbers(golden), rtol=3e-6)
        assert [l for l in lines if " TS " not in l] == [l for l in golden if " TS " not in l]
    out = subprocess.run([exes["minimal_native"], "-snes_fd_color", "-snes_converged_reason", "-snes_monitor_short",
                          "-ms_problem", "catenoid", "-ms_catenoid_c", "2.0", "-da_refine", "1"], capture_output=True,
                         text=True, timeout=300, check=True)
    lines = out.stdout.rstrip("\n").split("\n")
    assert lines[0] == "  0 SNES Function norm 1.08276"                                                  # minimal.test1:1
    assert lines[-1] == "done on 5 x 5 grid and problem catenoid:  error |u-uexact|_inf = 1.10603e-04"  # :8


def test_python_host_bdf_on_device(ctx):
    """-ts_type bdf through the Python host (p4pdes_b200/pattern.py:_bdf): pattern.test5 and the native host's steps."""
    r = pp.pattern_main(TEST5 + " -pc_type mg", ctx)
    assert len(r.lines) == len(GOLDEN_TEST5)
    np.testing.assert_allclose(_ts_numbers(r.lines), _ts_numbers(GOLDEN_TEST5), rtol=3e-6)
    assert [l for l in r.lines if " TS " not in l] == [l for l in GOLDEN_TEST5 if " TS " not in l]
    argv = "-da_grid_x 4 -da_grid_y 4 -da_refine 3 -ts_type bdf -ts_monitor -ts_max_time 60 -pc_type mg"
    a, b = pp.pattern_main(argv, ctx), pp.pattern_main(argv, ctx, native=True)
    np.testing.assert_allclose(_ts_numbers(a.lines), _ts_numbers(b.lines), rtol=1e-6)
