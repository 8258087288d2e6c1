"""The reference's UNCHANGED c/ch6/fish.c through the PETSc-shaped shim, on the CPU: the shim's work around the solve --
callbacks on the host, recognition of the Jacobian callback's values on every level ("stencilcuda"), Vec bookkeeping,
the reference's own report -- with the library replaced by the host stand-in (oracle/native/p4b_standin.cpp: the
recognised operator + Jacobi-CG, NO multigrid, so iteration counts are not the device path's and are not asserted).
The device run of the same driver is tests/test_gpu_fish_driver.py (validated on a B200)."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "fish_goldens.json")))
JAC = " -mg_levels_pc_type jacobi"


@pytest.fixture(scope="module")
def exe():
    path = os.path.join(ROOT, "oracle", "_ref", "fish_shim_host")
    if os.path.exists("/root/reference/c/ch6/fish.c"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref/fish_shim_host"])
    if not os.path.exists(path):
        pytest.skip("needs the reference tree to compile fish.c")
    return path


def fish(exe, opts, expect_rc=0):
    p = subprocess.run([exe] + opts.split(), capture_output=True, text=True, timeout=300)
    assert p.returncode == expect_rc, (p.returncode, p.stdout, p.stderr)
    return p.stdout.splitlines(), p.stderr


def test_golden_test1_lines(exe):
    g = GOLD["fish.test1"]
    lines, _ = fish(exe, g["options"] + JAC)
    assert lines[0] == "  0 SNES Function norm %s" % g["snes_fnorm0"]
    assert lines[1].startswith("    Linear solve converged due to CONVERGED_RTOL iterations ")
    assert lines[2] == "  1 SNES Function norm < 1.e-11"
    assert lines[3:] == ["problem %s on %s grid:" % (g["problem"], g["gridstr"]),
                         "  error |u-uexact|_inf = %s, |u-uexact|_h = %s" % (g["errinf"], g["err2h"])]


@pytest.mark.parametrize("name,opts", [      # (the goldens' own -snes_fd_color / -pc_mg_galerkin are not on the device path)
    ("fish.test6", "-fsh_dim 3 -da_refine 2 -fsh_problem manupoly -ksp_converged_reason -fsh_cx 0.01 -fsh_cy 2 -fsh_cz 100"),
    ("fish.test7", "-fsh_dim 3 -fsh_problem manupoly -ksp_converged_reason -da_refine 2"),
    ("fish.test2", "-fsh_dim 1 -fsh_problem manupoly -da_refine 1"),
])
def test_golden_error_lines(exe, name, opts):
    g = GOLD[name]
    lines, _ = fish(exe, opts + " -pc_type mg -ksp_rtol 1.0e-12" + JAC)
    assert lines[-2:] == ["problem %s on %s grid:" % (g["problem"], g["gridstr"]),
                          "  error |u-uexact|_inf = %s, |u-uexact|_h = %s" % (g["errinf"], g["err2h"])]


def test_reference_error_paths(exe):
    _, err = fish(exe, "-fsh_cx -1 -fsh_problem manupoly -pc_type mg" + JAC, expect_rc=2)
    assert "positivity required" in err
    _, err = fish(exe, "-fsh_dim 4 -pc_type mg" + JAC, expect_rc=1)
    assert "invalid dim" in err
    _, err = fish(exe, "-fsh_dim 2 -da_refine 2", expect_rc=56)
    assert "ILU" in err
    _, err = fish(exe, "-fsh_dim 2 -da_refine 2 -pc_type mg", expect_rc=56)
    assert "SOR" in err
    _, err = fish(exe, "-fsh_dim 2 -da_refine 2 -pc_type mg -snes_grid_sequence 1" + JAC, expect_rc=56)
    assert "newtonls only" in err


@pytest.mark.parametrize("name,verbatim", [("fish.test2", True), ("fish.test5", False), ("fish.test8", False)])
def test_goldens_with_fd_color_and_symmetry_report(exe, name, verbatim):
    """fish.test2,5,8 run `-snes_fd_color -mat_is_symmetric tol` (c/ch6/makefile:14,23,32).  The shim honours -snes_fd_color by
    checking F(u+v) - F(u) = A v (user's residual on the host, device MatMult of the recognised Jacobian) and reports the
    symmetry it measures on the device operator.  test2 (1-D, 5 points) is reproduced verbatim; test5/8's error norms carry
    the algebraic error of the golden's own CG + ILU(0) solve at rtol 1e-5 (the oracle reproduces them with ILU,
    tests/test_oracle_goldens.py), so only their report lines are compared here."""
    g = GOLD[name]
    out, _ = fish(exe, g["options"] + " -pc_type none")
    ref = "/root/reference/c/ch6/output/%s" % name
    tol = g["options"].split("-mat_is_symmetric ")[1].split()[0]
    assert out[:2] == ["Matrix is symmetric (tolerance %g)" % float(tol)] * 2 and g["symmetric_lines"] == 2
    assert out[2] == "problem %s on %s grid:" % (g["problem"], g["gridstr"]) and len(out) == 4
    if verbatim:
        assert out[3] == "  error |u-uexact|_inf = %s, |u-uexact|_h = %s" % (g["errinf"], g["err2h"])
        if os.path.exists(ref):
            assert out == open(ref).read().splitlines()
    else:
        tight, _ = fish(exe, g["options"] + " -pc_type none -ksp_rtol 1e-12")
        inf_t, inf_g = float(tight[3].split()[3].rstrip(",")), float(g["errinf"])
        assert abs(inf_t - inf_g) <= 0.15 * inf_g          # discretisation error; the golden adds its solver's 1e-5


# ---- the assembled Mat type: -mat_type sellcuda (values kept as inserted -> CSR -> SELL-32 SpMV; KSPCG as a host loop) ----
def test_assembled_type_gives_the_structured_type_s_answer(exe):
    """The unchanged fish.c with -mat_type sellcuda: same iteration counts, same residual history (to 1e-10 relative),
    same report as the matrix-free structured type (here both over the host stand-in: two independent CG loops)."""
    for opts in ("-fsh_dim 3 -fsh_problem manupoly -da_refine 2 -ksp_rtol 1.0e-12 -fsh_cx 0.01 -fsh_cy 2 -fsh_cz 100 -pc_type none",
                 "-fsh_dim 2 -da_refine 4 -pc_type jacobi", "-fsh_dim 1 -da_refine 5 -pc_type none -ksp_rtol 1e-10"):
        common = opts + " -ksp_converged_reason -ksp_monitor -snes_monitor_short"
        a, _ = fish(exe, common)
        b, _ = fish(exe, common + " -mat_type sellcuda")
        assert len(a) == len(b) and any("KSP Residual norm" in l for l in a)
        for la, lb in zip(a, b):
            if "KSP Residual norm" in la:            # 13 printed digits: the two loops round differently in the last one
                assert la.split()[0] == lb.split()[0] and abs(float(la.split()[-1]) - float(lb.split()[-1])) <= 1e-10 * float(la.split()[-1])
            else:
                assert la == lb
    _, err = fish(exe, "-fsh_dim 2 -da_refine 3 -pc_type mg -mg_levels_pc_type jacobi -mat_type sellcuda", expect_rc=56)
    assert "-mat_type stencilcuda" in err
    _, err = fish(exe, "-fsh_dim 2 -da_refine 3 -pc_type none -mat_type dense", expect_rc=86)
    assert "registered: stencilcuda, sellcuda" in err


def test_variable_coefficients_need_and_get_the_assembled_type(exe, tmp_path):
    """tests/shim_cases/ksponly_varcoef.c: -div(a grad u) + c u = f with a varying: the structured type refuses the Jacobian
    callback's values and names the way out; the assembled type solves, and the answer is NumPy's for the same discretisation."""
    import numpy as np
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    obj, out = str(tmp_path / "vc.o"), str(tmp_path / "vc")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), "-c",
                           os.path.join(ROOT, "tests", "shim_cases", "ksponly_varcoef.c"), "-o", obj])
    o = os.path.join(ROOT, "oracle", "_ref", "obj")
    subprocess.check_call(["g++", obj, os.path.join(o, "petscshim.o"), os.path.join(o, "p4b_standin.o"), "-o", out, "-lm"])
    _, err = fish(out, "-da_refine 3 -pc_type none", expect_rc=56)
    assert "diagonal is not constant" in err and "-mat_type sellcuda" in err
    lines, _ = fish(out, "-da_refine 3 -pc_type none -ksp_rtol 1e-13 -mat_type sellcuda -ksp_converged_reason")
    m = 33
    h = 1.0 / (m - 1)
    coef = lambda x, y: 1.0 + 0.5 * np.sin(3.0 * x) * np.cos(2.0 * y)
    g = lambda x, y: np.sin(x) + y * y
    idx = lambda i, j: j * m + i
    A = sp.lil_matrix((m * m, m * m))
    b = np.zeros(m * m)
    for j in range(m):
        for i in range(m):
            x, y, r = i * h, j * h, idx(i, j)
            if i in (0, m - 1) or j in (0, m - 1):
                A[r, r], b[r] = 1.0, g(x, y)
                continue
            nb = [(i + 1, j, coef(x + 0.5 * h, y)), (i - 1, j, coef(x - 0.5 * h, y)), (i, j + 1, coef(x, y + 0.5 * h)),
                  (i, j - 1, coef(x, y - 0.5 * h))]
            A[r, r] = sum(a for _, _, a in nb) + h * h * 2.0
            b[r] = h * h * (1.0 + x * y)
            for ii, jj, a in nb:
                if ii in (0, m - 1) or jj in (0, m - 1):
                    b[r] += a * g(ii * h, jj * h)
                else:
                    A[r, idx(ii, jj)] = -a
    u = spla.spsolve(A.tocsc(), b)
    got = lines[-1].split()
    assert lines[-1].startswith("done on 33 x 33 grid: sum ")
    assert abs(float(got[7]) - u.sum()) <= 1e-9 * abs(u.sum()) and abs(float(got[9]) - u[idx(m // 2, m // 2)]) <= 1e-10
    # the -snes_fd_color consistency check and the symmetry report work on the assembled matrix as well
    chk, _ = fish(out, "-da_refine 2 -pc_type jacobi -mat_type sellcuda -mat_is_symmetric 1e-10 -snes_fd_color")
    assert chk[:2] == ["Matrix is symmetric (tolerance 1e-10)"] * 2 and chk[2].startswith("done on 17 x 17 grid")
    # Jacobi with a varying diagonal: D^-1 applied as a diagonal SELL matrix; fewer iterations, same answer
    jl, _ = fish(out, "-da_refine 3 -pc_type jacobi -ksp_rtol 1e-13 -mat_type sellcuda -ksp_converged_reason")
    its = lambda ls: int(ls[0].split()[-1])
    assert its(jl) < its(lines) and abs(float(jl[-1].split()[7]) - u.sum()) <= 1e-9 * abs(u.sum())
