"""World-size-2 (and 3) gloo runs on CPU of the slab-decomposed algorithm: the plan the product library
makes (p4b_plan_levels) + ghost exchange + replicated coarse levels must reproduce the single-rank oracle."""
import numpy as np
import pytest
import torch.multiprocessing as mp

from oracle import fish_oracle as fo
from p4pdes_b200 import lib as L
from tests import dist_oracle


def run(world, dim, refine, levels, rtol, port, rep_points=None):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=dist_oracle.worker, args=(r, world, port, dim, refine, levels, rtol, q, rep_points))
             for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return outs


@pytest.mark.parametrize("world,dim,refine,levels,port,rep", [
    (2, 3, 4, 0, 29611, None),   # 33^3: only the finest level is distributed, 17^3 and below replicated
    (2, 2, 7, 0, 29612, 100),    # 2-D 257^2 in y-slabs, replication threshold lowered: 5 distributed levels
    (3, 3, 4, 3, 29613, None),   # uneven slabs (33 = 11+11+11), -pc_mg_levels 3
    (2, 3, 4, 0, 29614, 1),      # 33^3 .. 9^3 distributed (every rank keeps >= 2 planes), 5^3 and 3^3 replicated
    (4, 3, 5, 4, 29615, 1),      # 65^3 on 4 ranks, levels 65/33/17/9: 9^3 cannot be split 4 ways -> replicated
])
def test_slab_algorithm_matches_single_rank_oracle(world, dim, refine, levels, port, rep):
    rtol = 1e-10
    want = fo.fish(dim=dim, refine=refine, rtol=rtol, mg=fo.MGOptions(levels=levels or None))
    outs = run(world, dim, refine, levels, rtol, port, rep)
    y = np.concatenate([o[1].ravel() for o in outs])
    slabs = [o[4] for o in outs]
    n_last = want.grid.m[dim - 1]
    assert sum(s[1] for s in slabs) == n_last and slabs[0][0] == 0
    for o in outs:
        assert o[2] == want.its
        np.testing.assert_allclose(o[3], want.history, rtol=1e-10)
    assert np.linalg.norm(y - want.y) / np.linalg.norm(want.y) < 1e-12
    assert outs[0][5][-1] is False          # the finest level is always distributed
    if rep is not None and rep <= 100:
        assert outs[0][5].count(False) >= 3  # several distributed levels were exercised


def test_plan_ownership_rules():
    # "coarse plane K belongs to the owner of fine plane 2K"; every plane owned exactly once; the DMDA split on top
    for dim, refine, P in ((3, 8, 8), (3, 7, 4), (3, 5, 2), (2, 9, 8), (3, 6, 3)):
        plan = L.plan_levels(L.refined_grid(dim, refine), None, P)
        top = plan[-1]
        nz = top["m"][dim - 1]
        assert top["zs"] == [r * (nz // P) + min(r, nz % P) for r in range(P)]
        for lf, lc in zip(plan[1:][::-1], plan[:-1][::-1]):
            ncz = lc["m"][dim - 1]
            owner = [-1] * ncz
            for r in range(P):
                for K in range(lc["zs"][r], lc["zs"][r] + lc["zm"][r]):
                    assert owner[K] == -1
                    owner[K] = r
                    assert lf["zs"][r] <= 2 * K < lf["zs"][r] + lf["zm"][r]
            assert -1 not in owner
        assert not top["replicated"]
        for l in plan:
            if not l["replicated"]:
                assert min(l["zm"]) >= 2
        # replication is monotone: once a level is replicated every coarser one is
        reps = [l["replicated"] for l in plan]
        assert reps == sorted(reps, reverse=True)
    # one rank: nothing is replicated
    assert not any(l["replicated"] for l in L.plan_levels(L.refined_grid(3, 4), None, 1))
