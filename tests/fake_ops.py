"""NumPy stand-in for the device Context, so that the control flow of p4pdes_b200/minimal.py (Newton, line search,
GMRES/CG, the multigrid cycle, grid sequencing) can be exercised on a machine without a GPU.

TEST INFRASTRUCTURE ONLY: the product never imports this file; every operation here is a restatement from oracle/ of
the C-ABI call of the same name (include/p4b200.h).  GPU tests run the same driver on the real Context."""
import numpy as np
import scipy.sparse as sp

from oracle import fish_oracle as fo
from oracle import minimal_pattern_oracle as mpo
from oracle import minimal_solver_oracle as mso
from oracle import pattern_solver_oracle as pso


class Vec:
    def __init__(self, a):
        self.a = np.asarray(a, dtype=np.float64).ravel().copy()

    def numel(self):
        return self.a.size


class FakeOps:
    def __init__(self):
        self.calls = {}

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    # memory
    def empty(self, n):
        return Vec(np.full(int(n), np.nan))

    def zeros(self, n):
        return Vec(np.zeros(int(n)))

    def to_host(self, t):
        return t.a.copy()

    def from_host(self, a):
        return Vec(a)

    def sync(self):
        pass

    # vectors
    def dot(self, x, y):
        return float(x.a @ y.a)

    def norm2(self, x):
        return float(np.linalg.norm(x.a))

    def norminf(self, x):
        return float(np.max(np.abs(x.a)))

    def wrms2(self, x, y, atol, rtol):
        return float(np.sum(((x.a - y.a) / (atol + rtol * np.maximum(np.abs(x.a), np.abs(y.a)))) ** 2))

    def axpy(self, a, x, y):
        y.a += a * x.a

    def aypx(self, a, x, y):
        y.a[:] = x.a + a * y.a

    def axpby(self, a, x, b, y, out):
        v = np.zeros(out.a.size)
        if x is not None:
            v = a * x.a
        if y is not None:
            v = v + b * y.a
        out.a[:] = v

    def copy(self, x, y):
        y.a[:] = x.a

    def set(self, a, y):
        y.a[:] = a

    # grids
    def grid2d(self, mx, my):
        return (mx, my)

    def initial_state2d(self, grid, g, u):
        mx, my = grid
        gg = g.a.reshape(my, mx)
        uu = np.zeros((my, mx))
        uu[0, :], uu[-1, :], uu[:, 0], uu[:, -1] = gg[0, :], gg[-1, :], gg[:, 0], gg[:, -1]
        u.a[:] = uu.ravel()

    def _P(self, grid):
        mx, my = grid
        return sp.kron(fo.interp1d((my - 1) // 2 + 1), fo.interp1d((mx - 1) // 2 + 1), format="csr")

    def restrict(self, grid, rf, bc):
        self._count("restrict")
        bc.a[:] = self._P(grid).T @ rf.a

    def prolong_add(self, grid, xc, xf):
        self._count("prolong_add")
        if hasattr(grid, "mx"):
            grid = (grid.mx, grid.my)
        xf.a += self._P(grid) @ xc.a

    # fish.c kernels on an L.Grid (obstacle.py drives them): restated from the fish oracle
    @staticmethod
    def _fo_grid(g):
        return fo.Grid(g.dim, (g.mx, g.my, g.mz), (g.Lx, g.Ly, g.Lz))

    def stencil_apply(self, g, u, y):
        y.a[:] = fo.jacobian(self._fo_grid(g), (g.cx, g.cy, g.cz)) @ u.a

    def poisson_function(self, g, u, f, gb, F):
        # Poisson2DFunctionLocal with node-sampled f and g (c/ch6/poissonfunctions.c:37-66); square 2-D grids, hx = hy
        m = g.mx
        uu, gg, ff = u.a.reshape(m, m), gb.a.reshape(m, m), f.a.reshape(m, m)
        h = g.Lx / (m - 1)
        out = 4.0 * (uu - gg)
        v = uu.copy()
        v[0, :], v[-1, :], v[:, 0], v[:, -1] = gg[0, :], gg[-1, :], gg[:, 0], gg[:, -1]
        out[1:-1, 1:-1] = (4.0 * uu[1:-1, 1:-1] - v[1:-1, :-2] - v[1:-1, 2:] - v[:-2, 1:-1] - v[2:, 1:-1]
                           - h * h * ff[1:-1, 1:-1])
        F.a[:] = out.ravel()

    def vi_inactive_mask(self, u, lower, F, mask):
        mask.a[:] = np.where((u.a <= lower.a + 1.0e-8) & (F.a > 0.0), 0.0, 1.0)

    def pointwise_mult(self, x, y, out):
        out.a[:] = x.a * y.a

    def pointwise_max(self, x, y, out):
        out.a[:] = np.maximum(x.a, y.a)

    def inject2d(self, cmx, cmy, uf, uc):
        uc.a[:] = uf.a.reshape(2 * cmy - 1, 2 * cmx - 1)[::2, ::2].ravel()

    # minimal.c callbacks and assembled Jacobians
    def minimal_sample(self, mx, my, problem, tent_H, c, g):
        g.a[:] = mpo.minimal_g(mx, my, "tent" if problem == 0 else "catenoid", tent_H, c).ravel()

    def minimal_function(self, mx, my, q, u, g, FF):
        self._count("minimal_function")
        FF.a[:] = mpo.minimal_function(u.a.reshape(my, mx), g.a.reshape(my, mx), q).ravel()

    def minimal_jacobian_fd(self, mx, my, q, u, g, F0, vals):
        self._count("minimal_jacobian_fd")
        gg = g.a.reshape(my, mx)
        A = mso.fd_jacobian(lambda w: mpo.minimal_function(w, gg, q), u.a.reshape(my, mx), F0.a.reshape(my, mx)).tocoo()
        v = np.zeros((9, my * mx))
        dj = A.col // mx - A.row // mx
        di = A.col % mx - A.row % mx
        v[3 * (dj + 1) + (di + 1), A.row] = A.data
        vals.a[:] = v.ravel()

    def sell_matrix(self, rowptr, colind, vals):
        A = sp.csr_matrix((vals, colind, rowptr), shape=(len(rowptr) - 1, len(rowptr) - 1))

        class _M:
            def mult(self, x, y):
                y.a[:] = A @ x.a
                return y
        return _M()

    def poisson_stencil9(self, mx, my, Lx, Ly, cx, cy, vals):
        self._count("poisson_stencil9")
        A = fo.jacobian(fo.Grid(2, (mx, my, 1), (Lx, Ly, 1.0)), (cx, cy, 1.0)).tocoo()
        v = np.zeros((9, my * mx))
        dj = A.col // mx - A.row // mx
        di = A.col % mx - A.row % mx
        v[3 * (dj + 1) + (di + 1), A.row] = A.data
        vals.a[:] = v.ravel()

    def _csr(self, mx, my, vals):
        from p4pdes_b200.minimal import stencil9_to_csr
        rp, ci, d = stencil9_to_csr(vals.a, mx, my)
        return sp.csr_matrix((d, ci, rp), shape=(mx * my, mx * my))

    def stencil9_apply(self, mx, my, vals, x, y):
        self._count("stencil9_apply")
        y.a[:] = self._csr(mx, my, vals) @ x.a

    def stencil9_lin(self, mx, my, vals, u, b, pm1, ca, cb, cg, jacobi, out):
        self._count("stencil9_lin")
        A = self._csr(mx, my, vals)
        r = (b.a if b is not None else 0.0) - A @ u.a
        if jacobi:
            r = r / A.diagonal()
        o = cb * u.a + cg * r
        if pm1 is not None:
            o = o + ca * pm1.a
        out.a[:] = o

    def stencil9_gershgorin(self, mx, my, vals, work):
        return mso.gershgorin_jacobi(self._csr(mx, my, vals))

    def dense_matvec(self, n, Ainv, b, x):
        x.a[:] = Ainv.a.reshape(n, n) @ b.a

    # pattern.c implicit stage equation
    def pattern_initial_state(self, mx, my, Lside, Y):
        Y.a[:] = mpo.pattern_initial_state(mx, my, Lside).ravel()

    def pattern_initial_state_noisy(self, mx, my, Lside, level, Y):
        # pattern.c:159-175 on the VecSetRandom stream (host code of the library: loads without a GPU)
        import ctypes as C
        from p4pdes_b200 import lib as L
        lib = L.load()
        state = C.c_ulonglong(lib.p4b_rander48_seed(0x12345678))
        r = np.empty(2 * mx * my)
        L.check(lib.p4b_rander48_fill(C.byref(state), r.size, r.ctypes.data))
        y = mpo.pattern_initial_state(mx, my, Lside).reshape(-1, 2)
        r = level * r.reshape(-1, 2)
        v = y[:, 1] + r[:, 1]
        Y.a[:] = np.stack([r[:, 0] + 1.0 - 2.0 * v, v], axis=1).ravel()

    def pattern_ifunction(self, mx, my, Lside, Du, Dv, Y, Ydot, F):
        F.a[:] = mpo.pattern_ifunction(Y.a.reshape(my, mx, 2), Ydot.a.reshape(my, mx, 2), Lside, Du, Dv).ravel()

    def pattern_rhsfunction(self, mx, my, phi, kappa, Y, G):
        G.a[:] = mpo.pattern_rhsfunction(Y.a.reshape(my, mx, 2), phi, kappa).ravel()

    def _pJ(self, m, Lside, Du, Dv, phi, kappa, shift, Y):
        J = mpo.pattern_ijacobian(m, m, shift, Lside, Du, Dv)
        if Y is not None:
            J = J - pso.rhs_jacobian(Y.a.reshape(m, m, 2), phi, kappa)
        return sp.csr_matrix(J)

    def pattern_jac_apply(self, m, Lside, Du, Dv, phi, kappa, shift, Y, X, out):
        self._count("pattern_jac_apply")
        out.a[:] = self._pJ(m, Lside, Du, Dv, phi, kappa, shift, Y) @ X.a

    def pattern_jac_lin(self, m, Lside, Du, Dv, phi, kappa, shift, Y, X, b, pm1, ca, cb, cg, jacobi, out):
        self._count("pattern_jac_lin")
        J = self._pJ(m, Lside, Du, Dv, phi, kappa, shift, Y)
        r = (b.a if b is not None else 0.0) - J @ X.a
        if jacobi:
            r = r / J.diagonal()
        o = cb * X.a + cg * r
        if pm1 is not None:
            o = o + ca * pm1.a
        out.a[:] = o

    def pattern_jac_gershgorin(self, m, Lside, Du, Dv, phi, kappa, shift, Y, work):
        return mso.gershgorin_jacobi(self._pJ(m, Lside, Du, Dv, phi, kappa, shift, Y))

    def pattern_restrict(self, Mx, My, rf, bc):
        bc.a[:] = pso.interpolation(Mx, My).T @ rf.a

    def pattern_prolong_add(self, Mx, My, xc, xf):
        xf.a += pso.interpolation(Mx, My) @ xc.a

    def pattern_inject(self, Mx, My, yf, yc):
        yc.a[:] = yf.a.reshape(2 * My, 2 * Mx, 2)[::2, ::2, :].ravel()
