"""Slab-decomposed NumPy restatement of the CG + V-cycle, run under torch.distributed (gloo) on CPU.

TEST INFRASTRUCTURE.  It follows the *device* algorithm's distribution exactly -- the level/slab plan comes
from the product library (p4b_plan_levels through ctypes: who owns which planes, which levels are
replicated), ghost planes are exchanged before every stencil / transfer that needs them, dot products are
all-reduced -- while the arithmetic is the oracle's.  Comparing its result with the single-rank oracle
checks the host-side sharding logic without a GPU.
"""
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fish_oracle as fo  # noqa: E402
from p4pdes_b200 import lib as L  # noqa: E402


class Lev:
    pass


def build(dim, refine, levels, rank, world, rep_points=None):
    g = L.refined_grid(dim, refine)
    if rep_points is not None:
        L.tune("rep_points", rep_points)
    opts = L.MGOpts()
    L.load().p4b_mg_default_opts(opts)
    opts.levels = levels
    plan = L.plan_levels(g, opts, world)
    levs = []
    for pl in plan:
        lv = Lev()
        og = fo.Grid(dim, pl["m"])
        sc, diag, vol = fo.stencil_coeffs(og)
        # slots (nx, ny, nz): 2-D grids are (mx, 1, my)
        if dim == 3:
            lv.n = pl["m"]; lv.act = (1, 1, 1); lv.c = sc
        else:
            lv.n = (pl["m"][0], 1, pl["m"][1]); lv.act = (1, 0, 1); lv.c = (sc[0], 0.0, sc[1])
        lv.diag = diag
        lv.rep = pl["replicated"]
        lv.zs_all, lv.zm_all = pl["zs"], pl["zm"]
        lv.own = (pl["zs"][rank], pl["zm"][rank])
        lv.zs, lv.zm = (0, lv.n[2]) if (lv.rep or world == 1) else lv.own
        lam = fo.lambda_max_jacobi(og)
        lv.emin, lv.emax = 0.1 * lam, 1.1 * lam
        shape = (lv.zm + 2, lv.n[1], lv.n[0])
        lv.x, lv.b, lv.t = np.zeros(shape), np.zeros(shape), np.zeros(shape)
        levs.append(lv)
    return g, levs


def halo(lv, v, rank, world):
    if world == 1 or lv.rep:
        return
    reqs = []
    lo, hi = lv.zs > 0, lv.zs + lv.zm < lv.n[2]
    rlo = torch.empty(v[0].shape, dtype=torch.float64)
    rhi = torch.empty(v[0].shape, dtype=torch.float64)
    if lo:
        reqs.append(dist.isend(torch.from_numpy(v[1].copy()), rank - 1))
        reqs.append(dist.irecv(rlo, rank - 1))
    if hi:
        reqs.append(dist.isend(torch.from_numpy(v[lv.zm].copy()), rank + 1))
        reqs.append(dist.irecv(rhi, rank + 1))
    for r in reqs:
        r.wait()
    if lo:
        v[0] = rlo.numpy()
    if hi:
        v[lv.zm + 1] = rhi.numpy()


def masks(lv):
    nx, ny, nz = lv.n
    k = np.arange(lv.zs - 1, lv.zs + lv.zm + 1).reshape(-1, 1, 1)
    j = np.arange(ny).reshape(1, -1, 1)
    i = np.arange(nx).reshape(1, 1, -1)
    bd = (i == 0) | (i == nx - 1)
    if lv.act[1]:
        bd = bd | (j == 0) | (j == ny - 1)
    bd = bd | (k == 0) | (k == nz - 1) | (k < 0) | (k > nz - 1)
    return np.broadcast_to(bd, (lv.zm + 2, ny, nx))


def apply_A(lv, u):
    """A u on owned planes (ghosts of u must be current); returns an array with ghost planes zeroed."""
    bd = masks(lv)
    um = np.where(bd, 0.0, u)          # masking the data == dropping columns to boundary nodes
    out = np.zeros_like(u)
    core = slice(1, lv.zm + 1)
    s = lv.diag * u[core]
    acc = np.zeros_like(s)
    acc[:, :, 1:] += lv.c[0] * um[core][:, :, :-1]
    acc[:, :, :-1] += lv.c[0] * um[core][:, :, 1:]
    if lv.act[1]:
        acc[:, 1:, :] += lv.c[1] * um[core][:, :-1, :]
        acc[:, :-1, :] += lv.c[1] * um[core][:, 1:, :]
    acc += lv.c[2] * (um[0:lv.zm] + um[2:lv.zm + 2])
    out[core] = np.where(bd[core], s, s - acc)
    return out


def cheb(lv, zero_guess, its, rank, world):
    scale = 2.0 / (lv.emax + lv.emin)
    alpha = 1.0 - scale * lv.emin
    mu, omegaprod = 1.0 / alpha, 2.0 / alpha
    cm1, ck = 1.0, mu
    pm1 = np.zeros_like(lv.x) if zero_guess else lv.x
    halo(lv, pm1, rank, world)
    pk = pm1 + scale * (lv.b - apply_A(lv, pm1)) / lv.diag
    for _ in range(1, its):
        halo(lv, pk, rank, world)
        r = lv.b - apply_A(lv, pk)
        cp1 = 2.0 * mu * ck - cm1
        om = omegaprod * ck / cp1
        pm1, pk = pk, (1.0 - om) * pm1 + om * pk + om * scale * r / lv.diag
        cm1, ck = ck, cp1
    lv.x = pk


def w1(d):
    return 0.5 if d else 1.0


def restrict(F, C, rf, own):
    """Planes own=(czs,czm) of b_c = P^T r from fine slab array rf (ghosts current)."""
    czs, czm = own
    nx, ny, _ = C.n
    out = np.zeros((czm, ny, nx))
    for Kl in range(czm):
        K = czs + Kl
        fk = 2 * K
        for dk in (-1, 0, 1):
            kf = fk + dk
            if kf < 0 or kf >= F.n[2]:
                continue
            pl = rf[kf - F.zs + 1]
            for dj in ((-1, 0, 1) if F.act[1] else (0,)):
                for di in (-1, 0, 1):
                    w = w1(dk) * w1(dj) * w1(di)
                    js = np.arange(ny) * (2 if F.act[1] else 1) + dj
                    is_ = np.arange(nx) * 2 + di
                    jm = (js >= 0) & (js < F.n[1])
                    im = (is_ >= 0) & (is_ < F.n[0])
                    sub = pl[np.ix_(js[jm], is_[im])]
                    out[Kl][np.ix_(np.where(jm)[0], np.where(im)[0])] += w * sub
    return out


def prolong_add(F, C, xc, xf):
    """xf[owned] += P xc ; xc is the coarse slab array (ghosts current) or the replicated full array."""
    for kl in range(F.zm):
        k = F.zs + kl
        K0, ok = k >> 1, k & 1
        acc = np.zeros((F.n[1], F.n[0]))
        for dk in range(ok + 1):
            pl = xc[K0 + dk - C.zs + 1]
            if F.act[1]:
                py = np.zeros((F.n[1], C.n[0]))
                py[0::2] = pl
                py[1::2] = 0.5 * (pl[:-1] + pl[1:])
            else:
                py = pl
            px = np.zeros((F.n[1], F.n[0]))
            px[:, 0::2] = py
            px[:, 1::2] = 0.5 * (py[:, :-1] + py[:, 1:])
            acc += px
        xf[kl + 1] += (0.5 if ok else 1.0) * acc


def gather_rep(lv, v, world):
    if world == 1:
        return
    full = torch.from_numpy(v[1:lv.zm + 1].copy())
    parts = []
    for r in range(world):
        buf = torch.zeros((lv.zm_all[r],) + tuple(full.shape[1:]), dtype=torch.float64)
        if r == dist.get_rank():
            buf.copy_(full[lv.zs_all[r]:lv.zs_all[r] + lv.zm_all[r]])
        if lv.zm_all[r] > 0:
            dist.broadcast(buf, r)
        parts.append(buf)
    v[1:lv.zm + 1] = torch.cat(parts).numpy()


def coarse_solve(lv, dim):
    og = fo.Grid(dim, lv.n if dim == 3 else (lv.n[0], lv.n[2], 1))
    A = fo.jacobian(og)
    import scipy.sparse.linalg as spla
    lv.x[1:lv.zm + 1] = spla.spsolve(A.tocsc(), lv.b[1:lv.zm + 1].ravel()).reshape(lv.x[1:lv.zm + 1].shape)


def cycle(levs, l, zero_guess, dim, rank, world):
    lv = levs[l]
    if l == 0:
        coarse_solve(lv, dim)
        return
    C = levs[l - 1]
    cheb(lv, zero_guess, 2, rank, world)
    halo(lv, lv.x, rank, world)
    lv.t = lv.b - apply_A(lv, lv.x)
    lv.t[0] = 0.0
    lv.t[-1] = 0.0
    halo(lv, lv.t, rank, world)
    boundary = C.rep and not lv.rep and world > 1
    own = C.own if boundary else (C.zs, C.zm)
    part = restrict(lv, C, lv.t, own)
    C.b[:] = 0.0
    C.b[1 + own[0] - C.zs: 1 + own[0] - C.zs + own[1]] = part
    if boundary:
        gather_rep(C, C.b, world)
    cycle(levs, l - 1, True, dim, rank, world)
    halo(C, C.x, rank, world)
    prolong_add(lv, C, C.x, lv.x)
    cheb(lv, False, 2, rank, world)


def allsum(v, world):
    if world == 1:
        return float(v)
    t = torch.tensor([v], dtype=torch.float64)
    dist.all_reduce(t)
    return float(t.item())


def solve(dim, refine, levels, rtol, rank, world, rep_points=None):
    g, levs = build(dim, refine, levels, rank, world, rep_points)
    T = levs[-1]
    og = fo.refined_grid(dim, refine)
    u0 = fo.initial_state(og, "manuexp")
    b_full = fo.form_function(og, u0, "manuexp").reshape(T.n[2], T.n[1], T.n[0])
    core = slice(1, T.zm + 1)
    b = b_full[T.zs:T.zs + T.zm]
    x = np.zeros_like(b)
    r = b.copy()

    def M(rr):
        T.b[:] = 0.0
        T.b[core] = rr
        cycle(levs, len(levs) - 1, True, dim, rank, world)
        return T.x[core].copy()

    z = M(r)
    beta = allsum(np.vdot(z, r), world)
    dp = math.sqrt(allsum(np.vdot(z, z), world))
    hist = [dp]
    ttol = rtol * dp
    its = 0
    p = np.zeros((T.zm + 2,) + b.shape[1:])
    beta_old = 1.0
    while dp > ttol and its < 100:
        p[core] = z if its == 0 else z + (beta / beta_old) * p[core]
        halo(T, p, rank, world)
        w = apply_A(T, p)[core]
        a = beta / allsum(np.vdot(p[core], w), world)
        x += a * p[core]
        r -= a * w
        z = M(r)
        beta_old = beta
        beta = allsum(np.vdot(z, r), world)
        dp = math.sqrt(allsum(np.vdot(z, z), world))
        its += 1
        hist.append(dp)
    return x, its, hist, (T.zs, T.zm), levs


def worker(rank, world, port, dim, refine, levels, rtol, q, rep_points=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, its, hist, slab, levs = solve(dim, refine, levels, rtol, rank, world, rep_points)
    q.put((rank, x, its, hist, slab, [lv.rep for lv in levs]))
    dist.barrier()
    dist.destroy_process_group()
