"""Host-side mirrors of the minimal.c / pattern.c callbacks and of the assembled-matrix SpMV (C ABI wrappers).

Same names and argument meaning as the reference callbacks (c/ch7/minimal.c:210-282, c/ch5/pattern.c:146-318);
arrays are torch float64 CUDA tensors in DMDA natural ordering (i fastest; pattern's (u,v) interleaved).
"""
import ctypes as C

import numpy as np

from . import lib as L

MINIMAL_PROBLEMS = {"tent": 0, "catenoid": 1}


def minimal_g(ctx, mx, my, problem="catenoid", tent_H=1.0, catenoid_c=1.1):
    g = ctx.empty(mx * my)
    L.check(ctx.lib.p4b_minimal_sample(ctx.h, mx, my, MINIMAL_PROBLEMS[problem], tent_H, catenoid_c, g.data_ptr()))
    return g


def minimal_form_function(ctx, mx, my, u, g, q=-0.5, out=None):
    """FormFunctionLocal of minimal.c."""
    FF = out if out is not None else ctx.empty(mx * my)
    L.check(ctx.lib.p4b_minimal_function(ctx.h, mx, my, q, u.data_ptr(), g.data_ptr(), FF.data_ptr()))
    return FF


def pattern_initial_state(ctx, mx, my, Lside=2.5):
    Y = ctx.empty(2 * mx * my)
    L.check(ctx.lib.p4b_pattern_initial_state(ctx.h, mx, my, Lside, Y.data_ptr()))
    return Y


def pattern_rhs_function(ctx, mx, my, Y, phi=0.024, kappa=0.06, out=None):
    G = out if out is not None else ctx.empty(2 * mx * my)
    L.check(ctx.lib.p4b_pattern_rhsfunction(ctx.h, mx, my, phi, kappa, Y.data_ptr(), G.data_ptr()))
    return G


def pattern_ifunction(ctx, mx, my, Y, Ydot, Lside=2.5, Du=8.0e-5, Dv=4.0e-5, out=None):
    F = out if out is not None else ctx.empty(2 * mx * my)
    L.check(ctx.lib.p4b_pattern_ifunction(ctx.h, mx, my, Lside, Du, Dv, Y.data_ptr(), Ydot.data_ptr(), F.data_ptr()))
    return F


def pattern_ijacobian_mult(ctx, mx, my, shift, X, Lside=2.5, Du=8.0e-5, Dv=4.0e-5, out=None):
    JX = out if out is not None else ctx.empty(2 * mx * my)
    L.check(ctx.lib.p4b_pattern_ijacobian_mult(ctx.h, mx, my, Lside, Du, Dv, shift, X.data_ptr(), JX.data_ptr()))
    return JX


class SellMatrix:
    """An assembled matrix on the device (SELL-32), built from host CSR arrays (what MatSetValuesStencil + assembly
    produce in the reference: poissonfunctions.c:140-258, pattern.c:220-306)."""

    def __init__(self, ctx, rowptr, colind, vals):
        self.ctx = ctx
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        colind = np.ascontiguousarray(colind, dtype=np.int32)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        self.nrows = rowptr.size - 1
        self.h = C.c_void_p()
        L.check(ctx.lib.p4b_sell_create(ctx.h, self.nrows, rowptr.ctypes.data_as(C.c_void_p),
                                        colind.ctypes.data_as(C.c_void_p), vals.ctypes.data_as(C.c_void_p),
                                        C.byref(self.h)))
        n, nnz, pad = C.c_int(), C.c_longlong(), C.c_longlong()
        L.check(ctx.lib.p4b_sell_info(self.h, C.byref(n), C.byref(nnz), C.byref(pad)))
        self.nnz, self.padded_nnz = nnz.value, pad.value

    def mult(self, x, y=None):
        y = y if y is not None else self.ctx.empty(self.nrows)
        L.check(self.ctx.lib.p4b_sell_spmv(self.h, x.data_ptr(), y.data_ptr()))
        return y

    def close(self):
        if self.h:
            self.ctx.lib.p4b_sell_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
