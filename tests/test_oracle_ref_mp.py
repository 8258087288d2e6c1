"""minimal.c / pattern.c callbacks: the NumPy restatement against the reference's own compiled code
(oracle/_ref/libfishref.so) and against the goldens' pure-callback known answers."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import minimal_pattern_oracle as mp

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libfishref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libfishref.so not built (no /root/reference)")
P = lambda a: a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def ref():
    lib = C.CDLL(LIB)
    lib.ref_pattern_ijacobian.restype = C.c_long
    lib.ref_minimal_function.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    lib.ref_minimal_g.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p]
    lib.ref_pattern_ifunction.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ref_pattern_rhsfunction.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    lib.ref_pattern_ijacobian.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_long,
                                          C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


@pytest.mark.parametrize("mx,my", [(5, 5), (9, 17), (33, 33)])
@pytest.mark.parametrize("problem,q", [("catenoid", -0.5), ("tent", -0.5), ("tent", 0.0), ("catenoid", -0.25)])
def test_minimal_function_matches_reference(ref, mx, my, problem, q):
    M = (C.c_int * 3)(mx, my, 1)
    pid = 0 if problem == "tent" else 1
    g = np.zeros((my, mx))
    ref.ref_minimal_g(M, pid, 1.0, 1.1, P(g))
    np.testing.assert_allclose(mp.minimal_g(mx, my, problem, 1.0, 1.1), g, rtol=1e-14, atol=1e-16)   # libm vs numpy: ulps
    u = np.random.default_rng(0).standard_normal((my, mx)) * 0.3
    FF = np.zeros((my, mx))
    assert ref.ref_minimal_function(M, pid, q, 1.0, 1.1, P(u), P(FF)) == 0
    np.testing.assert_allclose(mp.minimal_function(u, g, q), FF, rtol=1e-13, atol=1e-14)


def test_minimal_test1_initial_function_norm():
    # c/ch7/output/minimal.test1:1  "  0 SNES Function norm 1.08276": c/ch7/makefile:16 runs
    # -ms_problem catenoid -ms_catenoid_c 2.0 -da_refine 1 (3x3 -> 5x5 grid), a pure-callback known answer
    g = mp.minimal_g(5, 5, "catenoid", 1.0, 2.0)
    u0 = np.zeros((5, 5))
    bd = np.ones((5, 5), bool)
    bd[1:-1, 1:-1] = False
    u0[bd] = g[bd]                                  # InitialState(zeros, g on the boundary), minimal.c:157
    F = mp.minimal_function(u0, g, -0.5)
    assert "%g" % float("%.6g" % np.linalg.norm(F)) == "1.08276"


@pytest.mark.parametrize("mx,my", [(12, 12), (16, 24)])
def test_pattern_callbacks_match_reference(ref, mx, my):
    rng = np.random.default_rng(1)
    Y = mp.pattern_initial_state(mx, my) + 0.01 * rng.standard_normal((my, mx, 2))
    Yd = rng.standard_normal((my, mx, 2))
    F = np.zeros((my, mx, 2))
    assert ref.ref_pattern_ifunction(mx, my, 2.5, 8.0e-5, 4.0e-5, P(Y), P(Yd), P(F)) == 0
    np.testing.assert_allclose(mp.pattern_ifunction(Y, Yd), F, rtol=1e-13, atol=1e-15)
    G = np.zeros((my, mx, 2))
    assert ref.ref_pattern_rhsfunction(mx, my, 0.024, 0.06, P(Y), P(G)) == 0
    np.testing.assert_allclose(mp.pattern_rhsfunction(Y), G, rtol=1e-14, atol=1e-16)
    cap = 18 * 2 * mx * my
    row, col, val = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap)
    nnz = ref.ref_pattern_ijacobian(mx, my, 2.5, 8.0e-5, 4.0e-5, 0.37, cap, P(row), P(col), P(val))
    assert nnz == 18 * mx * my
    import scipy.sparse as sp
    Jref = sp.csr_matrix((val[:nnz], (row[:nnz], col[:nnz])), shape=(2 * mx * my, 2 * mx * my))
    J = mp.pattern_ijacobian(mx, my, 0.37)
    assert abs(J - Jref).max() <= 1e-15 * abs(Jref).max()
    # the IFunction is affine in (Y, Ydot) with exactly this Jacobian: F(Y, s*Y) = J(s) Y
    np.testing.assert_allclose(mp.pattern_ifunction(Y, 0.37 * Y).ravel(), J @ Y.ravel(), rtol=1e-12, atol=1e-14)
