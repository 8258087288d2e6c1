"""GPU parity of the minimal.c / pattern.c callback kernels and the SELL SpMV against the CPU oracle
(which is itself checked against the reference's compiled code in tests/test_oracle_ref_mp.py)."""
import numpy as np
import pytest
import torch

from oracle import fish_oracle as fo
from oracle import minimal_pattern_oracle as mp
from p4pdes_b200 import callbacks as cb
from p4pdes_b200.fish import Context

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return Context()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).ravel()).cuda()


def rel(a, b):
    a, b = a.cpu().numpy().ravel(), np.asarray(b).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("mx,my", [(5, 5), (9, 17), (33, 33), (129, 65), (513, 513)])
@pytest.mark.parametrize("problem,q", [("catenoid", -0.5), ("tent", -0.5), ("tent", 0.0), ("catenoid", -0.3)])
def test_minimal_form_function(ctx, mx, my, problem, q):
    g = mp.minimal_g(mx, my, problem, 1.0, 1.1)
    dg = cb.minimal_g(ctx, mx, my, problem, 1.0, 1.1)
    assert rel(dg, g) < 1e-14
    u = np.random.default_rng(0).standard_normal((my, mx)) * 0.3
    FF = cb.minimal_form_function(ctx, mx, my, dev(u), dg, q)
    assert rel(FF, mp.minimal_function(u, g, q)) < 1e-13


def test_minimal_test1_function_norm_on_device(ctx):
    # c/ch7/output/minimal.test1:1  "  0 SNES Function norm 1.08276"
    g = cb.minimal_g(ctx, 5, 5, "catenoid", 1.0, 2.0)
    u0 = g.clone().reshape(5, 5)
    u0[1:-1, 1:-1] = 0.0
    FF = cb.minimal_form_function(ctx, 5, 5, u0.reshape(-1).contiguous(), g, -0.5)
    assert "%g" % float("%.6g" % ctx.norm2(FF)) == "1.08276"


@pytest.mark.parametrize("mx,my", [(12, 12), (16, 24), (48, 48), (512, 512)])
def test_pattern_callbacks(ctx, mx, my):
    rng = np.random.default_rng(1)
    Y0 = mp.pattern_initial_state(mx, my)
    assert rel(cb.pattern_initial_state(ctx, mx, my), Y0) < 1e-14
    Y = Y0 + 0.01 * rng.standard_normal(Y0.shape)
    Yd = rng.standard_normal(Y0.shape)
    dY, dYd = dev(Y), dev(Yd)
    assert rel(cb.pattern_rhs_function(ctx, mx, my, dY), mp.pattern_rhsfunction(Y)) < 1e-14
    assert rel(cb.pattern_ifunction(ctx, mx, my, dY, dYd), mp.pattern_ifunction(Y, Yd)) < 1e-13
    J = mp.pattern_ijacobian(mx, my, 0.37)
    assert rel(cb.pattern_ijacobian_mult(ctx, mx, my, 0.37, dY), J @ Y.ravel()) < 1e-13


def test_sell_spmv_matches_assembled_jacobians(ctx):
    rng = np.random.default_rng(2)
    mats = [fo.jacobian(fo.refined_grid(3, 4)),                 # fish 7-point, 33^3 (rows of length 1..7)
            fo.jacobian(fo.refined_grid(2, 6), (1.0, 3.0, 1.0)),
            mp.pattern_ijacobian(48, 48, 0.2)]                  # pattern 9-point periodic, 2 dof
    import scipy.sparse as sp
    mats.append(sp.random(1000, 1000, density=0.01, random_state=3, format="csr") + sp.eye(1000, format="csr"))
    mats.append(sp.csr_matrix((37, 37)))                        # empty rows
    for A in mats:
        A = sp.csr_matrix(A)
        A.sort_indices()
        S = cb.SellMatrix(ctx, A.indptr, A.indices, A.data)
        assert S.nnz == A.nnz and S.padded_nnz >= A.nnz
        x = rng.standard_normal(A.shape[1])
        y = S.mult(dev(x))
        want = A @ x
        assert float(np.linalg.norm(y.cpu().numpy() - want)) <= 1e-13 * max(np.linalg.norm(want), 1.0)
        S.close()
