"""CPU oracle for the callbacks of c/ch7/minimal.c and c/ch5/pattern.c  (TEST INFRASTRUCTURE ONLY).

NumPy restatements, each citing the reference lines it follows.  They are checked against the reference's
OWN compiled code (oracle/_ref/libfishref.so, built unchanged from c/ch7/minimal.c and c/ch5/pattern.c by
oracle/refstub/Makefile; tests/test_oracle_ref_mp.py) and against the pure-callback known answers of the goldens
(c/ch7/output/minimal.test1:1 "0 SNES Function norm 1.08276").  The Newton / TS drivers around these callbacks live
in PETSc; their restatements are minimal_solver_oracle.py and pattern_solver_oracle.py (pinned on
c/ch7/output/minimal.test1,2,4 and c/ch5/output/pattern.test2).
"""
import numpy as np


# ---------------------------------------------------------------- minimal.c ----
def g_bdry_tent(x, y, tent_H):
    """minimal.c:27-34."""
    return np.where(x < 1.0e-8, 2.0 * tent_H * np.where(y < 0.5, y, 1.0 - y), 0.0) + 0.0 * (x + y)


def g_bdry_catenoid(x, y, c):
    """minimal.c:36-42."""
    return c * np.cosh(x / c) * np.sin(np.arccos((y / c) / np.cosh(x / c)))


def minimal_g(mx, my, problem="catenoid", tent_H=1.0, catenoid_c=1.1):
    """g sampled at every node of the unit-square grid (FormExactFromG, minimal.c:191-208)."""
    x = (np.arange(mx) * (1.0 / (mx - 1))).reshape(1, mx)
    y = (np.arange(my) * (1.0 / (my - 1))).reshape(my, 1)
    if problem == "tent":
        return g_bdry_tent(x, y, tent_H) * np.ones((my, mx))
    return g_bdry_catenoid(x, y, catenoid_c) * np.ones((my, mx))


def minimal_function(u, g, q=-0.5):
    """FormFunctionLocal, minimal.c:210-282.  u, g: (my, mx) arrays (g = boundary function at the nodes).

    Boundary rows FF = u - g (unscaled, :227); interior rows use g instead of u for every neighbour that is a
    boundary node, including the four diagonal ones (:230-256)."""
    my, mx = u.shape
    hx, hy = 1.0 / (mx - 1), 1.0 / (my - 1)
    bd = np.zeros((my, mx), dtype=bool)
    bd[0, :] = bd[-1, :] = True
    bd[:, 0] = bd[:, -1] = True
    w = np.where(bd, g, u)

    def sh(dj, di):
        out = np.zeros_like(w)
        js = slice(max(dj, 0), my + min(dj, 0))
        jd = slice(max(-dj, 0), my + min(-dj, 0))
        is_ = slice(max(di, 0), mx + min(di, 0))
        id_ = slice(max(-di, 0), mx + min(-di, 0))
        out[jd, id_] = w[js, is_]
        return out

    ue, uw, un, us = sh(0, 1), sh(0, -1), sh(1, 0), sh(-1, 0)
    une, unw, use, usw = sh(1, 1), sh(1, -1), sh(-1, 1), sh(-1, -1)
    DD = lambda s: np.power(1.0 + s, q)          # minimal.c:46-48
    dux = (ue - u) / hx
    duy = (un + une - us - use) / (4.0 * hy)
    De = DD(dux * dux + duy * duy)
    dux = (u - uw) / hx
    duy = (unw + un - usw - us) / (4.0 * hy)
    Dw = DD(dux * dux + duy * duy)
    dux = (ue + une - uw - unw) / (4.0 * hx)
    duy = (un - u) / hy
    Dn = DD(dux * dux + duy * duy)
    dux = (ue + use - uw - usw) / (4.0 * hx)
    duy = (u - us) / hy
    Ds = DD(dux * dux + duy * duy)
    FF = -(hy / hx) * (De * (ue - u) - Dw * (u - uw)) - (hx / hy) * (Dn * (un - u) - Ds * (u - us))
    return np.where(bd, u - g, FF)


# ---------------------------------------------------------------- pattern.c ----
def pattern_initial_state(mx, my, L=2.5):
    """InitialState without noise, pattern.c:146-179.  Returns (my, mx, 2) with [...,0]=u, [...,1]=v.
    Coordinates of a periodic DMDA with DMDASetUniformCoordinates(0,L,0,L): x_i = i L/mx (pattern.c:91-93)."""
    x = (np.arange(mx) * (L / mx)).reshape(1, mx)
    y = (np.arange(my) * (L / my)).reshape(my, 1)
    ledge = (L - 0.5) / 2.0
    redge = L - ledge
    inside = (x >= ledge) & (x <= redge) & (y >= ledge) & (y <= redge)
    sx, sy = np.sin(4.0 * np.pi * x), np.sin(4.0 * np.pi * y)
    v = np.where(inside, 0.5 * sx * sx * sy * sy, 0.0)
    u = 1.0 - 2.0 * v
    return np.stack([u, v], axis=-1)


def pattern_rhsfunction(Y, phi=0.024, kappa=0.06):
    """FormRHSFunctionLocal, pattern.c:185-199."""
    u, v = Y[..., 0], Y[..., 1]
    uv2 = u * v * v
    return np.stack([-uv2 + phi * (1.0 - u), uv2 - (phi + kappa) * v], axis=-1)


def _lap9(a):
    """[1 4 1; 4 -20 4; 1 4 1] with periodic wrap (pattern.c:252-257)."""
    r = lambda dj, di: np.roll(np.roll(a, -dj, axis=0), -di, axis=1)
    return (r(1, -1) + 4.0 * r(1, 0) + r(1, 1) + 4.0 * r(0, -1) - 20.0 * a + 4.0 * r(0, 1)
            + r(-1, -1) + 4.0 * r(-1, 0) + r(-1, 1))


def pattern_ifunction(Y, Ydot, L=2.5, Du=8.0e-5, Dv=4.0e-5):
    """FormIFunctionLocal, pattern.c:242-267: F = Ydot - C L9(Y), C = D/(6 h^2), h = L/mx."""
    my, mx, _ = Y.shape
    h = L / mx
    Cu, Cv = Du / (6.0 * h * h), Dv / (6.0 * h * h)
    return np.stack([Ydot[..., 0] - Cu * _lap9(Y[..., 0]), Ydot[..., 1] - Cv * _lap9(Y[..., 1])], axis=-1)


def pattern_ijacobian(mx, my, shift, L=2.5, Du=8.0e-5, Dv=4.0e-5):
    """FormIJacobianLocal, pattern.c:274-318, as a SciPy CSR matrix on the interleaved (u,v) ordering."""
    import scipy.sparse as sp
    h = L / mx
    C = (Du / (6.0 * h * h), Dv / (6.0 * h * h))
    n = mx * my
    idx = np.arange(n).reshape(my, mx)
    rows, cols, vals = [], [], []
    for c in (0, 1):
        for dj, di, wgt in ((0, 0, None), (0, -1, 4.0), (0, 1, 4.0), (-1, 0, 4.0), (1, 0, 4.0),
                            (-1, -1, 1.0), (1, -1, 1.0), (-1, 1, 1.0), (1, 1, 1.0)):
            nb = np.roll(np.roll(idx, -dj, axis=0), -di, axis=1)
            rows.append(2 * idx.ravel() + c)
            cols.append(2 * nb.ravel() + c)
            vals.append(np.full(n, shift + 20.0 * C[c] if wgt is None else -wgt * C[c]))
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(2 * n, 2 * n))
