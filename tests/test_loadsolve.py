"""c/ch2/loadsolve.c's path: PETSc binary Mat / Vec files -> Krylov solve (p4pdes_b200/loadsolve.py, petscbin.py), on the
CPU through the NumPy stand-in for the device context.  Device: tests/test_gpu_loadsolve.py."""
import json
import os
import struct

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from p4pdes_b200 import loadsolve as ls
from p4pdes_b200 import petscbin
from tests.fake_ops import FakeOps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "loadsolve_goldens.json")))["loadsolve.test1"]


def test_mat_record_layout_and_round_trip(tmp_path):
    csr, b, _ = ls.tri_system(4)
    f = tmp_path / "A.dat"
    with open(f, "wb") as fh:
        petscbin.write_mat(fh, *csr)
    raw = f.read_bytes()
    # [PETSc] MatView_SeqAIJ_Binary: class id, M, N, nz; row lengths; column indices; values -- all big-endian
    assert raw[:16] == struct.pack(">iiii", 1211216, 4, 4, 10)
    assert raw[16:32] == struct.pack(">4i", 2, 3, 3, 2)
    assert raw[32:72] == struct.pack(">10i", 0, 1, 0, 1, 2, 1, 2, 3, 2, 3)
    assert raw[72:] == struct.pack(">10d", 3, -1, -1, 3, -1, -1, 3, -1, -1, 3)
    ((m, n), (rp, ci, v)), = petscbin.read_file(f)
    assert (m, n) == (4, 4) and np.array_equal(rp, csr[0]) and np.array_equal(ci, csr[1]) and np.array_equal(v, csr[2])
    f.write_bytes(raw[:-8])
    with pytest.raises(ValueError, match="truncated"):
        petscbin.read_file(f)


def golden_run(ops, tmp_path, extra):
    csr, b, _ = ls.tri_system(4)
    A, bb = str(tmp_path / "A.dat"), str(tmp_path / "b.dat")
    ls.write_system(A, bb, csr, b)
    opts = GOLD["options"].replace("-fA A.dat", "-fA " + A).replace("-fb b.dat", "-fb " + bb)
    return ls.loadsolve_main(opts + extra, ops), A, bb


def check_golden(rep, A, bb):
    want = [l.replace("A.dat", A).replace("b.dat", bb) for l in GOLD["lines"]]
    assert len(rep.lines) == len(want)
    for a, b in zip(rep.lines, want):
        if "type:" in b:
            assert a.strip().startswith("type:")         # the golden names PETSc's seqaij / seq, this path its own types
        else:
            assert a == b


@pytest.mark.parametrize("extra", [" -pc_type none -ksp_rtol 1e-12", " -pc_type jacobi -ksp_rtol 1e-12",
                                   " -ksp_type cg -pc_type jacobi -ksp_rtol 1e-12"])
def test_golden_loadsolve_test1(tmp_path, extra):
    """c/ch2/output/loadsolve.test1: the -verbose lines, the ASCII views of A and b (PETSc's formats) and the solution
    exp(cos i) to the printed digits.  (The golden's GMRES + ILU(0) solves a tridiagonal system exactly; here the
    tolerance is tightened instead.)"""
    rep, A, bb = golden_run(FakeOps(), tmp_path, extra)
    check_golden(rep, A, bb)
    assert rep.reason == "CONVERGED_RTOL" and rep.its <= 4


@pytest.mark.parametrize("ksp,pc", [("gmres", "none"), ("gmres", "jacobi"), ("cg", "jacobi"), ("cg", "none")])
def test_larger_systems_against_scipy(tmp_path, ksp, pc):
    m = 2000
    csr, b, xexact = ls.tri_system(m)
    A, bb = str(tmp_path / "A.dat"), str(tmp_path / "b.dat")
    ls.write_system(A, bb, csr, b)
    rep = ls.loadsolve_main("-fA %s -fb %s -ksp_type %s -pc_type %s -ksp_rtol 1e-10 -ksp_converged_reason" % (A, bb, ksp, pc),
                            FakeOps())
    assert rep.reason == "CONVERGED_RTOL" and rep.lines[-1].startswith("Linear solve converged due to CONVERGED_RTOL iterations")
    M = sp.csr_matrix((csr[2], csr[1], csr[0]), shape=(m, m))
    np.testing.assert_allclose(rep.x, spla.spsolve(M.tocsc(), b), rtol=1e-8)
    np.testing.assert_allclose(rep.x, xexact, rtol=1e-8)
    assert rep.its <= 40                                   # diagonally dominant: cond <= 5


def test_options_and_errors(tmp_path):
    csr, b, _ = ls.tri_system(5)
    A, bb = str(tmp_path / "A.dat"), str(tmp_path / "b.dat")
    ls.write_system(A, bb, csr, b)
    rep = ls.loadsolve_main("-fA %s -pc_type none -verbose" % A, FakeOps())             # loadsolve.c:84-90: no -fb
    assert rep.lines[-1] == "right-hand-side vector b not provided ... using zero vector of length 5"
    assert rep.its == 0 and not rep.x.any()
    for argv, msg in (("-pc_type none", "no input matrix provided"), ("-fA %s" % A, "ILU"),
                      ("-fA %s -pc_type none -ksp_type bcgs" % A, "gmres and cg"), ("-fA %s -pc_type none" % bb, "one Mat record")):
        with pytest.raises(ValueError, match=msg):
            ls.loadsolve_main(argv, FakeOps())
    ls.write_system(A, bb, csr, np.ones(4))
    with pytest.raises(ValueError, match="do not match"):
        ls.loadsolve_main("-fA %s -fb %s -pc_type none" % (A, bb), FakeOps())
    with open(A, "wb") as fh:
        petscbin.write_mat(fh, [0, 1, 2], [0, 1], [1.0, 1.0], ncols=3)
    with pytest.raises(ValueError, match="square"):
        ls.loadsolve_main("-fA %s -pc_type none" % A, FakeOps())
