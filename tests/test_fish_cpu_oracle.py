"""Pin the C oracle (oracle/fish_cpu.c, OpenMP) -- the checker of the BASELINE-size GPU parity tests and the CPU arm of
bench.py -- against (1) the reference's goldens it can run (c/ch6/output/fish.test1, fish.test3: Chebyshev(2)+SOR PCMG)
and (2) the NumPy oracle (oracle/fish_oracle.py, itself pinned on fish.test1-8) in 1-/2-/3-D with the BASELINE smoother
(Chebyshev(2)/Jacobi): equal KSP counts, preconditioned-residual history to 1e-12, solution to 1e-13."""
import numpy as np
import pytest

from oracle import fish_cpu as fc
from oracle import fish_oracle as fo


def fmt(x):
    return "%.3e" % x


def test_c_oracle_reproduces_golden_fish_test1(goldens):
    g = goldens["fish.test1"]
    r = fc.solve(dim=1, refine=3, problem="manupoly", rtol=1e-12, smoother_pc="sor", threads=1)
    assert "%g" % float("%.6g" % r["fnorm0"]) == g["snes_fnorm0"]
    assert r["its"] == g["ksp_its"] and r["fnorm1"] < 1e-11
    assert fmt(r["errinf"]) == g["errinf"] and fmt(r["err2h"]) == g["err2h"]


def test_c_oracle_reproduces_golden_fish_test3(goldens):
    g = goldens["fish.test3"]
    r = fc.solve(dim=2, refine=1, problem="manuexp", gonboundary=False, smoother_pc="sor", threads=1)
    assert r["its"] == g["ksp_its"]
    assert fmt(r["errinf"]) == g["errinf"] and fmt(r["err2h"]) == g["err2h"]


@pytest.mark.parametrize("dim,refine,problem,levels,rtol", [
    (1, 5, "manupoly", 0, 1e-12),
    (2, 6, "manuexp", 0, 1e-10),          # BASELINE config 1 (129^2)
    (2, 5, "manupoly", 3, 1e-8),
    (3, 4, "manuexp", 0, 1e-10),
    (3, 5, "manuexp", 4, 1e-10),          # 65^3, coarse grid 9^3 as in the 257^3 / 513^3 configurations
])
def test_c_oracle_equals_numpy_oracle(dim, refine, problem, levels, rtol):
    want = fo.fish(dim, refine, problem, rtol=rtol, mg=fo.MGOptions(levels=levels or None))
    for threads in (1, 4):                # the OpenMP reductions must not change the counts
        got = fc.solve(dim=dim, refine=refine, problem=problem, levels=levels, rtol=rtol, threads=threads,
                       want_arrays=True)
        assert got["its"] == want.its
        # relative 1e-12, plus rounding noise at eps x the first norm (entries 10 orders below it carry fewer digits)
        np.testing.assert_allclose(got["history"], want.history, rtol=1e-12 if threads == 1 else 1e-10,
                                   atol=4e-16 * want.history[0])
        u = got["u"]
        assert np.linalg.norm(u - want.u.ravel()) <= 1e-13 * np.linalg.norm(want.u)
        assert abs(got["errinf"] - want.errinf) <= 1e-12 and abs(got["fnorm0"] - want.fnorm0) <= 1e-12 * want.fnorm0


def test_stream_triad_figure_is_reported():
    gbs = fc.stream_triad_gbs(n=1 << 22, reps=2)
    assert gbs > 0.5
