"""CPU-side checks of the C ABI: the library builds, loads, exports every declared symbol, and the
host-only entry points behave.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import pytest

from p4pdes_b200 import build as p4build
from p4pdes_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    p4build.build()
    return L.load()


def test_header_and_bindings_agree(lib):
    hdr = open(os.path.join(ROOT, "include", "p4b200.h")).read()
    declared = set(re.findall(r"\b(p4b_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(L.exported_symbols())
    for name in declared:
        assert hasattr(lib, name), "libp4b200.so does not export %s" % name


def test_version(lib):
    assert lib.p4b_version() == 100


def test_slab_range_matches_dmda_split(lib):
    # DMDA ownership: first (m % P) ranks get one more (17 -> 9+8, 9 -> 5+4, 5 -> 3+2, 3 -> 2+1)
    for m, P, want in ((17, 2, [(0, 9), (9, 8)]), (9, 2, [(0, 5), (5, 4)]), (5, 2, [(0, 3), (3, 2)]),
                       (3, 2, [(0, 2), (2, 1)]), (513, 8, None)):
        got = []
        for r in range(P):
            s, c = C.c_int(), C.c_int()
            assert lib.p4b_slab_range(m, P, r, C.byref(s), C.byref(c)) == 0
            got.append((s.value, c.value))
        if want:
            assert got == want
        assert sum(c for _, c in got) == m
        assert all(got[r][0] + got[r][1] == got[r + 1][0] for r in range(P - 1))
    s, c = C.c_int(), C.c_int()
    assert lib.p4b_slab_range(10, 2, 5, C.byref(s), C.byref(c)) != 0
    assert b"slab" in lib.p4b_last_error()


def test_lambda_max_matches_survey_table(lib):
    # 1.7071 (5), 1.9239 (9), 1.9808 (17), 1.9952 (33)
    for m, want in ((5, 1.7071), (9, 1.9239), (17, 1.9808), (33, 1.9952)):
        for dim in (1, 2, 3):
            g = L.make_grid(dim, (m,) * dim)
            lam = C.c_double()
            assert lib.p4b_lambda_max_jacobi(C.byref(g), C.byref(lam)) == 0
            assert abs(lam.value - want) < 5e-5


def test_default_options(lib):
    o = L.MGOpts()
    assert lib.p4b_mg_default_opts(C.byref(o)) == 0
    assert (o.cycle, o.smoother, o.smooth_its, o.est_lo, o.est_hi) == (L.CYCLE_V, L.SMOOTH_CHEBYSHEV, 2, 0.1, 1.1)


def test_invalid_grid_is_reported(lib):
    g = L.make_grid(3, (9, 9, 9), c=(1.0, -1.0, 1.0))
    lam = C.c_double()
    assert lib.p4b_lambda_max_jacobi(C.byref(g), C.byref(lam)) == 2          # fish.c:188-190 error code 2
    assert b"positivity" in lib.p4b_last_error()


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.p4b_ctx_create(0, None, C.byref(h))
    assert rc != 0 and b"no CPU fallback" in lib.p4b_last_error()
    from p4pdes_b200.fish import Context
    with pytest.raises(L.P4BError):
        Context()


def test_option_parser_mirrors_fish_checks():
    from p4pdes_b200.fish import parse_options
    o = parse_options("-fsh_dim 3 -da_refine 7 -pc_mg_levels 6 -pc_type mg -snes_type ksponly -ksp_converged_reason")
    assert (o.dim, o.da_refine, o.pc_mg_levels, o.ksp_converged_reason) == (3, 7, 6, True)
    with pytest.raises(L.P4BError, match="MANUEXP"):
        parse_options("-fsh_cx 2 -pc_type mg")
    with pytest.raises(L.P4BError, match="positivity"):
        parse_options("-fsh_problem manupoly -fsh_cy -1 -pc_type mg")
    with pytest.raises(L.P4BError, match="sequential"):
        parse_options("-pc_type mg -mg_levels_pc_type sor")
    with pytest.raises(L.P4BError, match="ILU"):
        parse_options("-fsh_dim 2")


def test_minimal_solve_structs_match_the_header(lib, tmp_path):
    """ctypes mirrors of p4b_minimal_opts / p4b_minimal_result against the C header: defaults written by the library
    arrive in the right fields, and the sizes agree with what a C compiler makes of include/p4b200.h."""
    import subprocess
    o = L.MinimalOpts()
    assert lib.p4b_minimal_default_opts(C.byref(o)) == 0
    assert (o.problem, o.q, o.catenoid_c, o.tent_H, o.exact_init) == (1, -0.5, 1.1, 1.0, 0)
    assert (o.grid_x, o.grid_y, o.refine, o.grid_sequence) == (3, 3, 0, 0)
    assert (o.ksp_type, o.ksp_rtol, o.ksp_max_it, o.gmres_restart) == (0, 1.0e-5, 10000, 30)
    assert (o.pc_type, o.mg_levels, o.smooth_its) == (1, 0, 2)
    assert (o.snes_rtol, o.snes_stol, o.snes_atol, o.snes_max_it) == (1.0e-8, 1.0e-8, 1.0e-50, 50)
    assert (o.snes_monitor, o.snes_converged_reason, o.ksp_converged_reason) == (0, 0, 0)
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "p4b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(p4b_minimal_opts),sizeof(p4b_minimal_stage),sizeof(p4b_minimal_result),'
                   'offsetof(p4b_minimal_result,errinf),offsetof(p4b_minimal_stage,fnorm));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    assert got == [C.sizeof(L.MinimalOpts), C.sizeof(L.MinimalStage), C.sizeof(L.MinimalResult),
                   L.MinimalResult.errinf.offset, L.MinimalStage.fnorm.offset]


def test_pattern_solve_structs_match_the_header(lib, tmp_path):
    import subprocess
    o = L.PatternOpts()
    assert lib.p4b_pattern_default_opts(C.byref(o)) == 0
    assert (o.L, o.Du, o.Dv, o.phi, o.kappa) == (2.5, 8.0e-5, 4.0e-5, 0.024, 0.06)                  # pattern.c:47-52
    assert (o.no_rhsjacobian, o.call_back_report, o.grid_x, o.grid_y, o.refine) == (0, 0, 3, 3, 0)
    assert (o.ts_type, o.ts_dt, o.ts_max_time, o.ts_max_steps) == (0, 5.0, 200.0, 5000)            # pattern.c:115-117
    assert (o.ts_rtol, o.ts_atol, o.ts_monitor, o.pc_type, o.smooth_its, o.mg_rscale) == (1e-4, 1e-4, 0, 1, 2, 1.0)
    assert (o.snes_rtol, o.snes_stol, o.snes_atol, o.snes_max_it) == (1e-8, 1e-8, 1e-50, 50)
    assert (o.ksp_rtol, o.ksp_max_it, o.gmres_restart, o.snes_converged_reason, o.ksp_converged_reason) == (1e-5, 10000, 30, 0, 0)
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "p4b200.h"\nint main(void){printf("%zu %zu %zu %zu\\n",'
                   'sizeof(p4b_pattern_opts),sizeof(p4b_pattern_result),offsetof(p4b_pattern_result,step_newton),'
                   'offsetof(p4b_pattern_result,error));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    assert got == [C.sizeof(L.PatternOpts), C.sizeof(L.PatternResult), L.PatternResult.step_newton.offset,
                   L.PatternResult.error.offset]


def test_c_example_hosts_build_as_c99_and_refuse_to_run_without_a_device():
    """examples/*.c are what INTEGRATION.md shows a C maintainer: they must compile as pedantic C99 against the header and
    link against the library; without a CUDA device they fail loudly (no CPU fallback behind the C ABI)."""
    import subprocess
    import torch
    exes = p4build.build_examples()
    assert set(exes) == {"minimal_native", "pattern_native"}
    if torch.cuda.is_available():
        pytest.skip("a device is present: the run itself is covered by tests/test_gpu_r2_*.py")
    for exe in exes.values():
        p = subprocess.run([exe], capture_output=True, text=True, timeout=60)
        assert p.returncode == 1 and "no CPU fallback" in p.stderr
        assert subprocess.run([exe, "-bogus"], capture_output=True).returncode == 2
