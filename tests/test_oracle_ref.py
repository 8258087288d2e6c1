"""Validate the oracle's restatement of the discretisation against the REFERENCE'S OWN CODE: oracle/_ref/libfishref.so
is c/ch6/poissonfunctions.c + c/ch6/fish.c compiled unchanged from /root/reference (oracle/refstub/Makefile) with a
host-only stub of the few PETSc calls the callbacks make.  Skipped when the library is absent (it is built by
__graft_entry__.build() wherever /root/reference exists and travels to the GPU box as a built file)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import fish_oracle as fo

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libfishref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libfishref.so not built (no /root/reference)")
PROB = {"manupoly": 0, "manuexp": 1, "zero": 2}

CASES = [
    (1, (17,), (1.0, 1.0, 1.0), (1.0, 1.0, 1.0)),
    (1, (9,), (2.0, 1.0, 1.0), (3.0, 1.0, 1.0)),
    (2, (9, 9), (1.0, 1.0, 1.0), (1.0, 1.0, 1.0)),
    (2, (17, 9), (1.0, 2.0, 1.0), (1.0, 2.5, 1.0)),
    (3, (9, 9, 9), (1.0, 1.0, 1.0), (0.01, 2.0, 100.0)),
    (3, (9, 5, 17), (1.0, 0.5, 2.0), (1.0, 1.0, 1.0)),
]


@pytest.fixture(scope="module")
def ref():
    lib = C.CDLL(LIB)
    lib.ref_jacobian.restype = C.c_long
    return lib


def arrs(dim, m, Ls, c):
    mm = tuple(m) + (1,) * (3 - len(m))
    return (C.c_int * 3)(*mm), (C.c_double * 3)(*Ls), (C.c_double * 3)(*c), fo.Grid(dim, mm, tuple(Ls))


@pytest.mark.parametrize("dim,m,Ls,c", CASES)
@pytest.mark.parametrize("problem", ["manupoly", "manuexp", "zero"])
def test_residual_matches_reference_callback(ref, dim, m, Ls, c, problem):
    if problem == "manuexp":
        c = (1.0, 1.0, 1.0)
    M, L, cc, og = arrs(dim, m, Ls, c)
    u = np.random.default_rng(0).standard_normal(og.n)
    F = np.zeros(og.n)
    assert ref.ref_function(dim, M, L, cc, PROB[problem], u.ctypes.data_as(C.c_void_p), F.ctypes.data_as(C.c_void_p)) == 0
    mine = fo.form_function(og, u, problem, c).ravel()
    np.testing.assert_allclose(mine, F, rtol=1e-13, atol=1e-13 * np.abs(F).max())


@pytest.mark.parametrize("dim,m,Ls,c", CASES)
def test_jacobian_matches_reference_callback(ref, dim, m, Ls, c):
    M, L, cc, og = arrs(dim, m, Ls, c)
    cap = 8 * og.n
    row, col, val = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap)
    nnz = ref.ref_jacobian(dim, M, L, cc, cap, row.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p),
                           val.ctypes.data_as(C.c_void_p))
    assert nnz > 0
    import scipy.sparse as sp
    Aref = sp.csr_matrix((val[:nnz], (row[:nnz], col[:nnz])), shape=(og.n, og.n))
    A = fo.jacobian(og, c)
    assert (abs(A - Aref)).max() <= 1e-15 * abs(Aref).max()
    assert A.nnz == Aref.nnz                       # same sparsity: columns to boundary nodes dropped


@pytest.mark.parametrize("dim,m,Ls,c", CASES)
@pytest.mark.parametrize("problem", ["manupoly", "manuexp"])
def test_initial_state_and_exact_solution(ref, dim, m, Ls, c, problem):
    M, L, cc, og = arrs(dim, m, Ls, c)
    for gonb in (0, 1):
        u = np.full(og.n, 7.0)
        assert ref.ref_initial_state(dim, M, L, PROB[problem], gonb, u.ctypes.data_as(C.c_void_p)) == 0
        np.testing.assert_allclose(fo.initial_state(og, problem, bool(gonb)).ravel(), u, rtol=1e-15, atol=0)
    ue = np.zeros(og.n)
    assert ref.ref_uexact(dim, M, L, PROB[problem], ue.ctypes.data_as(C.c_void_p)) == 0
    x, y, z = og.coords()
    np.testing.assert_allclose((fo.u_exact(dim, problem, x, y, z) * np.ones(og.shape)).ravel(), ue, rtol=1e-15, atol=0)


def test_fish_test1_function_norm_from_reference_code(ref, goldens):
    # the golden's first line, produced by the reference's own callback on the reference's own initial state
    M, L, cc, og = arrs(1, (17,), (1.0, 1.0, 1.0), (1.0, 1.0, 1.0))
    u = np.zeros(17)
    ref.ref_initial_state(1, M, L, 0, 1, u.ctypes.data_as(C.c_void_p))
    F = np.zeros(17)
    ref.ref_function(1, M, L, cc, 0, u.ctypes.data_as(C.c_void_p), F.ctypes.data_as(C.c_void_p))
    assert "%g" % float("%.6g" % np.linalg.norm(F)) == goldens["fish.test1"]["snes_fnorm0"]
