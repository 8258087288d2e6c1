"""CPU oracle for the fish.c CG + geometric-multigrid path  (TEST INFRASTRUCTURE ONLY).

This file is a NumPy/SciPy restatement of the algorithm the reference runs when
``c/ch6/fish.c`` is executed with ``-pc_type mg``.  It is the parity checker for
the CUDA path; nothing in the product (``p4pdes_b200/``) may import it.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg use it.

What is restated, and from where (paths relative to /root/reference):

* discretisation:  ``c/ch6/poissonfunctions.c:4-115`` (residuals), ``:117-258``
  (Jacobians), ``:260-346`` (InitialState); problem functions ``c/ch6/fish.c:15-82``;
  error report ``c/ch6/fish.c:248-280``.
* the solver lives in PETSc, which is NOT vendored in the reference and has NO pinned
  version (``c/ch6/makefile:1-2`` includes ``${PETSC_DIR}/lib/petsc/conf/variables``).
  Its published algorithms are restated here: DMDA Q1
  interpolation, PCMG multiplicative V/W cycle with R = P^T and rediscretised coarse
  operators, KSPCHEBYSHEV (first kind, PETSc counting), PCJACOBI, PCSOR (local
  symmetric sweep; only so the goldens can be checked), KSPRICHARDSON, exact coarse
  solve, KSPCG with the preconditioned norm, SNESKSPONLY.

Pinning (tests/test_oracle_goldens.py): the restatement reproduces the reference's
golden outputs ``c/ch6/output/fish.test1`` (complete), ``fish.test3`` (complete),
``fish.test4`` (complete, 2-rank block SSOR + W cycle), ``fish.test5,6,8`` (CG + the default
ILU(0) PC: iteration count of test6 and all error norms) and the error norms of
``fish.test2,7``, and ``fish.test7`` complete: its 11 iterations (-pc_mg_galerkin, 2 ranks) need P^T A P coarse
operators, the two-rank block SSOR AND [PETSc]'s GMRES-estimated Chebyshev targets (lambda_hat = 1.22, not 1, for the
block smoother: MGOptions.estimate = "gmres"); with the analytic targets the count is 23.  The discretisation part is
additionally checked against the reference's OWN compiled code (oracle/_ref/libfishref.so = c/ch6/poissonfunctions.c +
c/ch6/fish.c built unchanged by oracle/refstub/Makefile; tests/test_oracle_ref.py).  No reference golden uses Chebyshev+Jacobi or a grid larger
than 17^2 / 9^3, so at BASELINE sizes parity is pinned only transitively
(oracle validated on goldens -> oracle run at size).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# --------------------------------------------------------------------------------------
# problem functions: c/ch6/fish.c:15-82
# --------------------------------------------------------------------------------------

PROBLEMS = ("manupoly", "manuexp", "zero")


def u_exact(dim, problem, x, y, z):
    """fish.c:15-41 (g_bdry_ptr table fish.c:115-118)."""
    if problem == "zero":
        return np.zeros(np.broadcast(x, y, z).shape)
    if problem == "manupoly":
        a = x * x * (1.0 - x * x)
        if dim == 1:
            return a + 0.0 * y + 0.0 * z
        a = a * y * y * (y * y - 1.0)
        if dim == 2:
            return a + 0.0 * z
        return a * z * z * (z * z - 1.0)
    if problem == "manuexp":
        if dim == 1:
            return -np.exp(x) + 0.0 * y + 0.0 * z
        if dim == 2:
            return -x * np.exp(y) + 0.0 * z
        return -x * np.exp(y + z)
    raise ValueError(problem)


def f_rhs(dim, problem, x, y, z, c):
    """fish.c:45-82 (f_rhs_ptr table fish.c:120-123).  c = (cx,cy,cz)."""
    cx, cy, cz = c
    if problem == "zero":
        return np.zeros(np.broadcast(x, y, z).shape)
    if problem == "manupoly":
        if dim == 1:
            return cx * 12.0 * x * x - 2.0 + 0.0 * y + 0.0 * z
        aa = x * x * (1.0 - x * x)
        bb = y * y * (y * y - 1.0)
        ddaa = 2.0 * (1.0 - 6.0 * x * x)
        ddbb = 2.0 * (6.0 * y * y - 1.0)
        if dim == 2:
            return -(cx * ddaa * bb + cy * aa * ddbb) + 0.0 * z
        cc = z * z * (z * z - 1.0)
        ddcc = 2.0 * (6.0 * z * z - 1.0)
        return -(cx * ddaa * bb * cc + cy * aa * ddbb * cc + cz * aa * bb * ddcc)
    if problem == "manuexp":
        if dim == 1:
            return np.exp(x) + 0.0 * y + 0.0 * z
        if dim == 2:
            return x * np.exp(y) + 0.0 * z
        return 2.0 * x * np.exp(y + z)
    raise ValueError(problem)


# --------------------------------------------------------------------------------------
# grid + discretisation
# --------------------------------------------------------------------------------------

@dataclass
class Grid:
    """A DMDA-like vertex-centred grid; arrays are indexed [k, j, i] (i fastest)."""
    dim: int
    m: tuple            # (mx, my, mz); unused dims are 1
    L: tuple = (1.0, 1.0, 1.0)

    @property
    def shape(self):
        mx, my, mz = self.m
        return (mz, my, mx)

    @property
    def n(self):
        return self.m[0] * self.m[1] * self.m[2]

    def h(self):
        """poissonfunctions.c:74-76: h = (max-min)/(m-1) per used dim."""
        return tuple(self.L[d] / (self.m[d] - 1) if d < self.dim else 1.0 for d in range(3))

    def coords(self):
        """x = xmin + i*h  (poissonfunctions.c:83-87), as broadcastable [k,j,i] arrays."""
        hx, hy, hz = self.h()
        mx, my, mz = self.m
        x = (np.arange(mx) * hx).reshape(1, 1, mx)
        y = (np.arange(my) * hy).reshape(1, my, 1) if self.dim >= 2 else np.zeros((1, 1, 1))
        z = (np.arange(mz) * hz).reshape(mz, 1, 1) if self.dim >= 3 else np.zeros((1, 1, 1))
        return x, y, z

    def bdry_mask(self):
        """True where the node is on the Dirichlet boundary (poissonfunctions.c:12,45,88-90)."""
        mx, my, mz = self.m
        b = np.zeros(self.shape, dtype=bool)
        b[:, :, 0] = True
        b[:, :, mx - 1] = True
        if self.dim >= 2:
            b[:, 0, :] = True
            b[:, my - 1, :] = True
        if self.dim >= 3:
            b[0, :, :] = True
            b[mz - 1, :, :] = True
        return b

    def coarsen(self):
        """DMDA coarsening, ratio 2, non-periodic: m -> (m-1)/2+1."""
        mc = []
        for d in range(3):
            if d < self.dim:
                if (self.m[d] - 1) % 2 != 0 or self.m[d] < 3:
                    raise ValueError("grid %r cannot be coarsened" % (self.m,))
                mc.append((self.m[d] - 1) // 2 + 1)
            else:
                mc.append(1)
        return Grid(self.dim, tuple(mc), self.L)


def refined_grid(dim, refine, base=3, L=(1.0, 1.0, 1.0)):
    """-da_refine n on the 3^d base DMDA of fish.c:199-212: m <- 1 + 2^n (m-1)."""
    m = 1 + (2 ** refine) * (base - 1)
    return Grid(dim, tuple(m if d < dim else 1 for d in range(3)), tuple(L))


def stencil_coeffs(g: Grid, c=(1.0, 1.0, 1.0)):
    """(sc[3], scdiag, vol) exactly as each dimension's callback computes them.

    1-D poissonfunctions.c:13-21,130-137: diag cx*2/h, off -cx/h, rhs scale h.
    2-D :37-40:  scx = cx*hy/hx, scy = cy*hx/hy, darea = hx*hy.
    3-D :77-81:  sc_d = c_d*dvol/(h_d*h_d), dvol = hx*hy*hz.
    """
    hx, hy, hz = g.h()
    cx, cy, cz = c
    if g.dim == 1:
        return (cx / hx, 0.0, 0.0), cx * 2.0 / hx, hx
    if g.dim == 2:
        scx = cx * hy / hx
        scy = cy * hx / hy
        return (scx, scy, 0.0), 2.0 * (scx + scy), hx * hy
    dvol = hx * hy * hz
    scx = cx * dvol / (hx * hx)
    scy = cy * dvol / (hy * hy)
    scz = cz * dvol / (hz * hz)
    return (scx, scy, scz), 2.0 * (scx + scy + scz), dvol


def _shift(a, axis, s, fill):
    """a shifted so that out[idx] = a[idx + s] along axis; out-of-range -> fill."""
    out = np.full_like(a, fill)
    n = a.shape[axis]
    src = [slice(None)] * a.ndim
    dst = [slice(None)] * a.ndim
    if s > 0:
        src[axis] = slice(s, n)
        dst[axis] = slice(0, n - s)
    else:
        src[axis] = slice(0, n + s)
        dst[axis] = slice(-s, n)
    out[tuple(dst)] = a[tuple(src)]
    return out


def form_function(g: Grid, u, problem, c=(1.0, 1.0, 1.0)):
    """F(u): Poisson{1,2,3}DFunctionLocal, poissonfunctions.c:4-115.

    Interior rows use g (not u) for neighbours that lie on the boundary; boundary rows
    are scdiag*(u-g).  The 1-D form keeps the reference's own operation order.
    """
    u = np.asarray(u, dtype=np.float64).reshape(g.shape)
    x, y, z = g.coords()
    gb = u_exact(g.dim, problem, x, y, z) * np.ones(g.shape)
    f = f_rhs(g.dim, problem, x, y, z, c) * np.ones(g.shape)
    bm = g.bdry_mask()
    ug = np.where(bm, gb, u)           # neighbour values: g on the boundary, u inside
    sc, scdiag, vol = stencil_coeffs(g, c)
    F = np.empty(g.shape)
    if g.dim == 1:
        hx = g.h()[0]
        uw = _shift(ug, 2, -1, 0.0)
        ue = _shift(ug, 2, +1, 0.0)
        Fi = c[0] * (2.0 * u - uw - ue) / hx - hx * f
        Fb = (u - gb) * (c[0] * (2.0 / hx))
        return np.where(bm, Fb, Fi)
    uw = _shift(ug, 2, -1, 0.0)
    ue = _shift(ug, 2, +1, 0.0)
    us = _shift(ug, 1, -1, 0.0)
    un = _shift(ug, 1, +1, 0.0)
    if g.dim == 2:
        Fi = scdiag * u - sc[0] * (uw + ue) - sc[1] * (us + un) - vol * f
    else:
        ud = _shift(ug, 0, -1, 0.0)
        uu = _shift(ug, 0, +1, 0.0)
        Fi = scdiag * u - sc[0] * (uw + ue) - sc[1] * (us + un) - sc[2] * (uu + ud) - vol * f
    Fb = (u - gb) * scdiag
    F = np.where(bm, Fb, Fi)
    return F


def jacobian(g: Grid, c=(1.0, 1.0, 1.0)):
    """Assembled Jacobian of Poisson{1,2,3}DJacobianLocal, poissonfunctions.c:117-258.

    Constant diagonal scdiag on every row; off-diagonals -sc_d only between two
    interior nodes (columns to boundary nodes are dropped, :133-138,:172-181,:221-245).
    """
    sc, scdiag, _ = stencil_coeffs(g, c)
    mx, my, mz = g.m
    n = g.n
    idx = np.arange(n).reshape(g.shape)
    interior = ~g.bdry_mask()
    rows = [idx.ravel()]
    cols = [idx.ravel()]
    vals = [np.full(n, scdiag)]
    for d, axis in ((0, 2), (1, 1), (2, 0)):
        if d >= g.dim:
            continue
        for s in (-1, +1):
            nb_int = _shift(interior, axis, s, False)
            sel = interior & nb_int
            nb_idx = _shift(idx, axis, s, -1)
            rows.append(idx[sel])
            cols.append(nb_idx[sel])
            vals.append(np.full(int(sel.sum()), -sc[d]))
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    return A


def initial_state(g: Grid, problem, gonboundary=True):
    """InitialState with ZEROS, poissonfunctions.c:260-346."""
    u = np.zeros(g.shape)
    if gonboundary:
        x, y, z = g.coords()
        gb = u_exact(g.dim, problem, x, y, z) * np.ones(g.shape)
        bm = g.bdry_mask()
        u[bm] = gb[bm]
    return u


# --------------------------------------------------------------------------------------
# DMDA Q1 interpolation and level hierarchy
# --------------------------------------------------------------------------------------

def interp1d(mc):
    """1-D vertex-centred ratio-2 interpolation, (2mc-1) x mc: weights 1 / (1/2,1/2)."""
    mf = 2 * (mc - 1) + 1
    rows, cols, vals = [], [], []
    for K in range(mc):
        rows.append(2 * K); cols.append(K); vals.append(1.0)
    for K in range(mc - 1):
        rows += [2 * K + 1, 2 * K + 1]; cols += [K, K + 1]; vals += [0.5, 0.5]
    return sp.csr_matrix((vals, (rows, cols)), shape=(mf, mc))


def interpolation(gc: Grid):
    """P: coarse -> fine (tensor product; boundary nodes included, no special-casing)."""
    Px = interp1d(gc.m[0])
    P = Px
    if gc.dim >= 2:
        P = sp.kron(interp1d(gc.m[1]), P, format="csr")
    if gc.dim >= 3:
        P = sp.kron(interp1d(gc.m[2]), P, format="csr")
    return sp.csr_matrix(P)


# --------------------------------------------------------------------------------------
# smoothers
# --------------------------------------------------------------------------------------

def split_rows(m_last, nranks):
    """DMDA ownership of the slowest dimension over nranks: first (m % P) ranks get one more."""
    base, rem = divmod(m_last, nranks)
    sizes = [base + (1 if r < rem else 0) for r in range(nranks)]
    starts = np.concatenate([[0], np.cumsum(sizes)])
    return [(int(starts[r]), int(starts[r + 1])) for r in range(nranks)]


class JacobiPC:
    """PCJACOBI: z = r / diag(A)."""
    def __init__(self, A, **kw):
        self.dinv = 1.0 / A.diagonal()

    def apply(self, r):
        return self.dinv * r


class BlockSSORPC:
    """PCSOR default: omega=1 local symmetric sweep per rank block.

    Oracle-only (the device path does not implement SOR: it is sequential).  The blocks
    are contiguous row ranges (natural ordering split along the slowest dimension).
    """
    def __init__(self, A, blocks=None, **kw):
        n = A.shape[0]
        self.blocks = blocks or [(0, n)]
        self.parts = []
        for (s, e) in self.blocks:
            Ab = sp.csr_matrix(A[s:e, s:e])
            D = sp.diags(Ab.diagonal())
            Lo = sp.tril(Ab, k=-1, format="csr")
            Up = sp.triu(Ab, k=1, format="csr")
            self.parts.append((sp.csr_matrix(D + Lo), Ab.diagonal(), sp.csr_matrix(D + Up)))

    def apply(self, r):
        z = np.empty_like(r)
        for (s, e), (DL, d, DU) in zip(self.blocks, self.parts):
            y = spla.spsolve_triangular(DL, r[s:e], lower=True)
            z[s:e] = spla.spsolve_triangular(DU, d * y, lower=False)
        return z


class ILU0PC:
    """PCILU, zero fill, natural ordering: PETSc's default PC on one rank.

    Oracle-only (sequential triangular solves).  Exists so that the goldens which use the
    default PC (fish.test5, fish.test6, fish.test8) pin the KSPCG restatement a second time.
    """
    def __init__(self, A, **kw):
        A = sp.csr_matrix(A).copy()
        A.sort_indices()
        n = A.shape[0]
        ip, ix, v = A.indptr, A.indices, A.data.copy()
        diag = np.zeros(n, dtype=np.int64)
        for i in range(n):
            pos = {int(ix[p]): p for p in range(ip[i], ip[i + 1])}
            for p in range(ip[i], ip[i + 1]):
                k = int(ix[p])
                if k >= i:
                    break
                v[p] /= v[diag[k]]
                for q in range(diag[k] + 1, ip[k + 1]):
                    j = int(ix[q])
                    if j in pos:
                        v[pos[j]] -= v[p] * v[q]
            diag[i] = pos[i]
        LU = sp.csr_matrix((v, ix, ip), shape=A.shape)
        self.Lo = sp.csr_matrix(sp.tril(LU, -1) + sp.eye(n))
        self.Up = sp.csr_matrix(sp.triu(LU, 0))

    def apply(self, r):
        y = spla.spsolve_triangular(self.Lo, r, lower=True)
        return spla.spsolve_triangular(self.Up, y, lower=False)


def chebyshev_smooth(A, pc, b, x, emin, emax, max_it):
    """KSPCHEBYSHEV, first kind, PETSc counting: max_it = number of PC applications."""
    scale = 2.0 / (emax + emin)
    alpha = 1.0 - scale * emin
    mu = 1.0 / alpha
    omegaprod = 2.0 / alpha
    c_m1, c_k = 1.0, mu
    p_m1 = x
    r = b - A @ p_m1
    p_k = p_m1 + scale * pc.apply(r)
    for _ in range(1, max_it):
        r = b - A @ p_k
        c_p1 = 2.0 * mu * c_k - c_m1
        omega = omegaprod * c_k / c_p1
        p_p1 = (1.0 - omega) * p_m1 + omega * p_k + omega * scale * pc.apply(r)
        p_m1, p_k = p_k, p_p1
        c_m1, c_k = c_k, c_p1
    return p_k


def richardson_smooth(A, pc, b, x, max_it, scale=1.0):
    """KSPRICHARDSON: x <- x + scale * B (b - A x)."""
    for _ in range(max_it):
        x = x + scale * pc.apply(b - A @ x)
    return x


def lambda_max_jacobi(g: Grid, c=(1.0, 1.0, 1.0)):
    """Analytic lambda_max(D^-1 A) = 1 + sum_d sc_d cos(pi/(m_d-1)) / sum_d sc_d  (SURVEY section 7).

    Stands in for PETSc's GMRES(10)-on-random-rhs estimate, which is not reproducible.
    Grids whose interior is a single node per dimension give cos(pi/2)=0 -> lambda = 1.
    """
    sc, _, _ = stencil_coeffs(g, c)
    num = sum(sc[d] * math.cos(math.pi / (g.m[d] - 1)) for d in range(g.dim))
    den = sum(sc[d] for d in range(g.dim))
    return 1.0 + num / den


# --------------------------------------------------------------------------------------
# PCMG
# --------------------------------------------------------------------------------------

@dataclass
class MGOptions:
    levels: int | None = None            # -pc_mg_levels (default: down to the 3^d grid)
    cycle: str = "v"                     # -pc_mg_cycle_type v|w
    smoother_ksp: str = "chebyshev"      # -mg_levels_ksp_type chebyshev|richardson
    smoother_pc: str = "jacobi"          # -mg_levels_pc_type jacobi|sor
    smoother_its: int = 2                # -mg_levels_ksp_max_it
    eig: tuple | None = None             # -mg_levels_ksp_chebyshev_eigenvalues emin,emax (explicit)
    esteig: tuple = (0.1, 1.1)           # targets (0.1, 1.1)*lambda_hat when eig is None
    nranks: int = 1                      # rank blocks for SOR (goldens only)
    galerkin: bool = False               # -pc_mg_galerkin (P^T A P coarse operators)
    seed: int = 0                        # of the noise vector of estimate="gmres"
    estimate: str = "analytic"           # lambda_hat for the Chebyshev targets: "analytic" (lambda_max_jacobi; 1 for SOR) or
                                         # "gmres" = [PETSc] KSPChebyshev's own estimate, gmres_lambda_max below


def gmres_lambda_max(A, pc, its=10, seed=0):
    """[PETSc] KSPChebyshev's default eigenvalue estimate: `its` GMRES iterations on B A with a noisy
    right-hand side, lambda_hat = the largest Ritz value (eigenvalues of the Arnoldi Hessenberg matrix).  PETSc draws the
    noise from its own PetscRandom; the estimate is insensitive to it at the digits that matter (tests/test_oracle_goldens.py
    runs four different streams), so a seeded NumPy generator stands in."""
    rng = np.random.default_rng(seed)
    r = pc.apply(rng.uniform(-1.0, 1.0, A.shape[0]))
    V = [r / np.linalg.norm(r)]
    H = np.zeros((its + 1, its))
    k = its
    for j in range(its):
        w = pc.apply(A @ V[j])
        for i in range(j + 1):                   # modified Gram-Schmidt
            H[i, j] = w @ V[i]
            w = w - H[i, j] * V[i]
        H[j + 1, j] = np.linalg.norm(w)
        if H[j + 1, j] < 1.0e-14:
            k = j + 1
            break
        V.append(w / H[j + 1, j])
    return float(np.max(np.linalg.eigvals(H[:k, :k]).real))


class PCMG:
    """Multiplicative multigrid preconditioner as PETSc's PCMG applies it."""

    def __init__(self, gfine: Grid, c=(1.0, 1.0, 1.0), opts: MGOptions | None = None):
        self.opts = o = opts or MGOptions()
        grids = [gfine]
        while True:
            if o.levels is not None and len(grids) >= o.levels:
                break
            gg = grids[-1]
            if any(gg.m[d] <= 3 or (gg.m[d] - 1) % 2 for d in range(gg.dim)):
                break
            grids.append(gg.coarsen())
        grids.reverse()                  # grids[0] = coarsest ... grids[-1] = finest
        self.grids = grids
        self.nlev = len(grids)
        self.P = [None] + [interpolation(grids[l - 1]) for l in range(1, self.nlev)]
        if o.galerkin:
            self.A = [None] * self.nlev
            self.A[-1] = jacobian(grids[-1], c)
            for l in range(self.nlev - 1, 0, -1):
                self.A[l - 1] = sp.csr_matrix(self.P[l].T @ self.A[l] @ self.P[l])
        else:
            self.A = [jacobian(gg, c) for gg in grids]      # rediscretisation (fish.c:7)
        self.coarse = spla.splu(sp.csc_matrix(self.A[0]))
        self.pc, self.eig = [None], [None]
        for l in range(1, self.nlev):
            gg = grids[l]
            if o.smoother_pc == "jacobi":
                pc = JacobiPC(self.A[l])
                lam = lambda_max_jacobi(gg, c)
            elif o.smoother_pc == "sor":
                rows = split_rows(gg.m[gg.dim - 1], o.nranks)
                plane = gg.n // gg.m[gg.dim - 1]
                pc = BlockSSORPC(self.A[l], [(s * plane, e * plane) for s, e in rows])
                lam = 1.0                # boundary rows give lambda_max(B A) = 1 exactly
            else:
                raise ValueError(o.smoother_pc)
            if o.estimate == "gmres":
                lam = gmres_lambda_max(self.A[l], pc, seed=getattr(o, "seed", 0))
            elif o.estimate != "analytic":
                raise ValueError(o.estimate)
            self.pc.append(pc)
            self.eig.append(o.eig if o.eig is not None else (o.esteig[0] * lam, o.esteig[1] * lam))

    def smooth(self, l, b, x):
        o = self.opts
        if o.smoother_ksp == "chebyshev":
            emin, emax = self.eig[l]
            return chebyshev_smooth(self.A[l], self.pc[l], b, x, emin, emax, o.smoother_its)
        if o.smoother_ksp == "richardson":
            return richardson_smooth(self.A[l], self.pc[l], b, x, o.smoother_its)
        raise ValueError(o.smoother_ksp)

    def mcycle(self, l, b, x):
        if l == 0:
            return self.coarse.solve(b)
        x = self.smooth(l, b, x)
        r = b - self.A[l] @ x
        bc = self.P[l].T @ r
        xc = np.zeros_like(bc)
        cycles = 1 if (l == 1 or self.opts.cycle == "v") else 2
        for _ in range(cycles):
            xc = self.mcycle(l - 1, bc, xc)
        x = x + self.P[l] @ xc
        return self.smooth(l, b, x)

    def apply(self, r):
        return self.mcycle(self.nlev - 1, r, np.zeros_like(r))


# --------------------------------------------------------------------------------------
# KSPCG (preconditioned norm) and the fish driver
# --------------------------------------------------------------------------------------

def cg(A, b, M, rtol=1e-5, abstol=1e-50, max_it=10000):
    """KSPSolve_CG, left preconditioning, KSP_NORM_PRECONDITIONED.

    Returns (x, its, history) with history[i] = ||M^-1 r_i||_2 (what -ksp_monitor prints).
    """
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    beta = float(z @ r)
    dp = float(np.sqrt(z @ z))
    hist = [dp]
    ttol = max(rtol * dp, abstol)
    its = 0
    p = None
    beta_old = None
    while dp > ttol and its < max_it:
        if p is None:
            p = z.copy()
        else:
            p = z + (beta / beta_old) * p
        w = A @ p
        a = beta / float(p @ w)
        x = x + a * p
        r = r - a * w
        z = M(r)
        beta_old = beta
        beta = float(z @ r)
        dp = float(np.sqrt(z @ z))
        its += 1
        hist.append(dp)
    return x, its, hist


@dataclass
class FishResult:
    grid: Grid
    its: int
    fnorm0: float
    fnorm1: float
    errinf: float
    err2h: float
    history: list
    u: np.ndarray = field(repr=False, default=None)
    b: np.ndarray = field(repr=False, default=None)
    y: np.ndarray = field(repr=False, default=None)


def error_norms(g: Grid, u, problem):
    """fish.c:248-276: |u-uexact|_inf and |u-uexact|_2 / sqrt(prod (m_d-1))."""
    x, y, z = g.coords()
    ue = u_exact(g.dim, problem, x, y, z) * np.ones(g.shape)
    e = (u.reshape(g.shape) - ue).ravel()
    normconst = math.sqrt(float(np.prod([g.m[d] - 1 for d in range(g.dim)], dtype=np.float64)))
    return float(np.abs(e).max()), float(np.sqrt(e @ e)) / normconst


def fish(dim=2, refine=0, problem="manuexp", c=(1.0, 1.0, 1.0), L=(1.0, 1.0, 1.0),
         gonboundary=True, rtol=1e-5, pc="mg", mg: MGOptions | None = None, max_it=10000):
    """./fish -fsh_dim dim -da_refine refine ... with SNESKSPONLY + KSPCG (fish.c:128-286)."""
    g = refined_grid(dim, refine, L=L)
    u0 = initial_state(g, problem, gonboundary)
    F0 = form_function(g, u0, problem, c).ravel()
    A = jacobian(g, c)
    if pc == "mg":
        M = PCMG(g, c, mg).apply
    elif pc == "none":
        M = lambda r: r.copy()
    elif pc == "ilu":
        M = ILU0PC(A).apply
    elif pc == "exact":
        lu = spla.splu(sp.csc_matrix(A))
        M = lu.solve
    else:
        raise ValueError(pc)
    y, its, hist = cg(A, F0, M, rtol=rtol, max_it=max_it)
    u = u0.ravel() - y
    F1 = form_function(g, u, problem, c).ravel()
    errinf, err2h = error_norms(g, u, problem)
    return FishResult(g, its, float(np.linalg.norm(F0)), float(np.linalg.norm(F1)), errinf, err2h, hist,
                      u=u.reshape(g.shape), b=F0, y=y)
