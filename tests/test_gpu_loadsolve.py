"""c/ch2/loadsolve.c's path on the device (p4pdes_b200/loadsolve.py): PETSc binary Mat / Vec -> SELL-32 on the GPU ->
GMRES / CG over the SpMV and Vec kernels.  CPU counterpart (same host code over the NumPy stand-in): tests/test_loadsolve.py."""
import numpy as np
import pytest
import torch

from p4pdes_b200 import loadsolve as ls
from p4pdes_b200.fish import Context
from tests.test_loadsolve import check_golden, golden_run

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


@pytest.fixture(scope="module")
def ctx():
    return Context()


@pytest.mark.parametrize("extra", [" -pc_type none -ksp_rtol 1e-12", " -ksp_type cg -pc_type jacobi -ksp_rtol 1e-12"])
def test_golden_loadsolve_test1_on_device(ctx, tmp_path, extra):
    rep, A, bb = golden_run(ctx, tmp_path, extra)
    check_golden(rep, A, bb)                      # c/ch2/output/loadsolve.test1
    assert rep.reason == "CONVERGED_RTOL"


@pytest.mark.parametrize("ksp,pc", [("gmres", "jacobi"), ("cg", "none")])
def test_large_tridiagonal_system_on_device(ctx, tmp_path, ksp, pc):
    """loadsolve.c:17-19's large example (there m = 10^7; 10^6 here keeps the files at 36 MB)."""
    m = 1000000
    csr, b, xexact = ls.tri_system(m)
    A, bb = str(tmp_path / "A.dat"), str(tmp_path / "b.dat")
    ls.write_system(A, bb, csr, b)
    rep = ls.loadsolve_main("-fA %s -fb %s -ksp_type %s -pc_type %s -ksp_rtol 1e-10" % (A, bb, ksp, pc), ctx)
    assert rep.reason == "CONVERGED_RTOL" and rep.its <= 40 and rep.n == m
    # the residual is reduced by 1e-10 in the 2-norm (cond(A) <= 5): the error is small in norm, single entries to ~1e-7
    assert np.linalg.norm(rep.x - xexact) <= 1e-8 * np.linalg.norm(xexact)
    np.testing.assert_allclose(rep.x, xexact, rtol=1e-6)
