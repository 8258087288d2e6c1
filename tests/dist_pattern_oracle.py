"""CPU (gloo) emulation of the y-slab algorithm of pattern.c's multi-GPU path (csrc/nk_device.cu SlabPatternOps): the
product's slab plan (p4b_pattern_slab_plan, host code of the library) + a ring exchange of one ghost row per side + the
"no wrap in y" index rules, restated in NumPy, must give the rows of the periodic single-rank operators of
oracle/pattern_solver_oracle.py.  Test infrastructure."""
import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from oracle import minimal_pattern_oracle as mpo
from oracle import pattern_solver_oracle as pso
from p4pdes_b200 import lib as L


def plan(m, grid_x, nranks, rank):
    lib = L.load()
    arr = [(C.c_int * 32)() for _ in range(4)]
    nl = lib.p4b_pattern_slab_plan(m, grid_x, 1, nranks, rank, *arr)
    assert nl > 0, L.last_error() if hasattr(L, "last_error") else nl
    return [dict(m=arr[0][i], dist=bool(arr[1][i]), ys=arr[2][i], ym=arr[3][i]) for i in range(nl)]


def ring_halo(owned, rank, world):
    """owned: (ym, mx, 2) -> (ym + 2, mx, 2) with the neighbours' rows (periodic ring)."""
    prev, nxt = (rank - 1) % world, (rank + 1) % world
    lo, hi = torch.empty(owned.shape[1:], dtype=torch.float64), torch.empty(owned.shape[1:], dtype=torch.float64)
    first, last = torch.from_numpy(owned[0].copy()), torch.from_numpy(owned[-1].copy())
    reqs = [dist.isend(first, prev, tag=1), dist.isend(last, nxt, tag=2), dist.irecv(hi, nxt, tag=1), dist.irecv(lo, prev, tag=2)]
    for r in reqs:
        r.wait()
    return np.concatenate([lo.numpy()[None], owned, hi.numpy()[None]])


def stencil_rows(g, Ydot, Lside, Du, Dv, m):
    """F = Ydot - C L9(Y) on the owned rows of a ghosted slab g (ym + 2 rows): wrap in x, none in y."""
    h = Lside / m
    Cc = np.array([Du, Dv]) / (6.0 * h * h)
    c = g[1:-1]
    w, e = np.roll(g, 1, axis=1), np.roll(g, -1, axis=1)
    lap = (w[2:] + 4.0 * g[2:] + e[2:] + 4.0 * w[1:-1] - 20.0 * c + 4.0 * e[1:-1] + w[:-2] + 4.0 * g[:-2] + e[:-2])
    return Ydot - Cc * lap


def worker(rank, world, port, m, grid_x, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)                       # the same global fields on every rank
    pl = plan(m, grid_x, world, rank)
    out = {"rank": rank, "plan": pl}
    # (1) stencil on the finest (distributed) level
    Y, D = rng.standard_normal((m, m, 2)), rng.standard_normal((m, m, 2))
    lv = pl[0]
    own = slice(lv["ys"], lv["ys"] + lv["ym"])
    F = stencil_rows(ring_halo(Y[own], rank, world), D[own], 2.5, 8e-5, 4e-5, m)
    out["stencil"] = float(np.max(np.abs(F - mpo.pattern_ifunction(Y, D)[own])))
    # (2) transfers down the hierarchy: restriction b_c = P^T r, prolongation x_f += P x_c, rows of the periodic operators
    errs = []
    for lf, lc in zip(pl[:-1], pl[1:]):
        mf, mc = lf["m"], lc["m"]
        if not lf["dist"]:
            break
        P = pso.interpolation(mc, mc)
        r = rng.standard_normal((mf, mf, 2))
        xc = rng.standard_normal((mc, mc, 2))
        fown = slice(lf["ys"], lf["ys"] + lf["ym"])
        ymc, ysc = lf["ym"] // 2, lf["ys"] // 2
        # restriction of this rank's coarse rows from its fine rows + the ghost row below
        g = ring_halo(r[fown], rank, world)              # g[k] = fine row ys - 1 + k
        bc = np.zeros((ymc, mc, 2))
        for J in range(ymc):
            for dj, wj in ((-1, 0.5), (0, 1.0), (1, 0.5)):
                row = g[2 * J + dj + 1]
                rx = row[0::2] + 0.5 * (np.roll(row, 1, axis=0)[0::2] + np.roll(row, -1, axis=0)[0::2])
                bc[J] += wj * rx
        want = (P.T @ r.ravel()).reshape(mc, mc, 2)
        errs.append(float(np.max(np.abs(bc - want[ysc:ysc + ymc]))))
        if not lc["dist"]:                               # first replicated level: all-gather, every rank gets the whole b_c
            parts = [torch.empty(bc.shape, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(bc))
            errs.append(float(np.max(np.abs(np.concatenate([p.numpy() for p in parts]) - want))))
            csl = np.concatenate([xc[ysc:ysc + ymc], xc[(ysc + ymc) % mc][None]])      # rows + the wrapped row above
        else:
            assert lc["ys"] == ysc and lc["ym"] == ymc
            csl = ring_halo(xc[ysc:ysc + ymc], rank, world)[1:]
        # prolongation onto this rank's fine rows
        xf = np.zeros((lf["ym"], mf, 2))
        for j in range(lf["ym"]):
            J0 = j >> 1
            rows = csl[J0] if j % 2 == 0 else 0.5 * (csl[J0] + csl[J0 + 1])
            xf[j, 0::2] = rows
            xf[j, 1::2] = 0.5 * (rows + np.roll(rows, -1, axis=0))
        wantf = (P @ xc.ravel()).reshape(mf, mf, 2)
        errs.append(float(np.max(np.abs(xf - wantf[fown]))))
    out["transfer"] = max(errs) if errs else None
    q.put(out)
    dist.barrier()
    dist.destroy_process_group()
