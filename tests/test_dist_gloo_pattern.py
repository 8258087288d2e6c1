"""World-size-2 and -4 gloo runs on CPU of the y-slab rules behind pattern.c's multi-GPU path (BASELINE config 5): the
slab plan of the product library, the ring exchange and the no-wrap-in-y index rules reproduce the periodic operators."""
import pytest
import torch.multiprocessing as mp

from tests import dist_pattern_oracle as dpo


def run(world, port, m, grid_x):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=dpo.worker, args=(r, world, port, m, grid_x, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=300) for _ in range(world)], key=lambda o: o["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return outs


@pytest.mark.parametrize("world,port,m,grid_x", [(2, 29621, 32, 4), (4, 29622, 64, 4), (2, 29623, 48, 3)])
def test_slab_rules_reproduce_the_periodic_operators(world, port, m, grid_x):
    outs = run(world, port, m, grid_x)
    for o in outs:
        assert o["stencil"] < 1e-12 and o["transfer"] is not None and o["transfer"] < 1e-12
    pl = outs[0]["plan"]
    assert pl[0]["dist"] and not pl[-1]["dist"]                   # finest distributed, base grid replicated
    rows = [o["plan"][0] for o in outs]
    assert [r["ys"] for r in rows] == [k * m // world for k in range(world)] and all(r["ym"] == m // world for r in rows)


def test_plan_rules():
    # BASELINE config 5: 2048^2 on 8 ranks, base grid 4 (-da_grid_x 4 -da_refine 9)
    pl = dpo.plan(2048, 4, 8, 3)
    assert [l["m"] for l in pl] == [2048 >> k for k in range(10)]
    assert [l["dist"] for l in pl] == [True] * 8 + [False] * 2    # 2048 .. 16 on slabs (256 .. 2 rows each), 8 and 4 replicated
    assert pl[0]["ys"] == 3 * 256 and pl[7]["ym"] == 2 and pl[8]["ym"] == 8
    # rows that do not split evenly are refused
    import ctypes as C
    from p4pdes_b200 import lib as L
    arr = [(C.c_int * 32)() for _ in range(4)]
    assert L.load().p4b_pattern_slab_plan(48, 3, 1, 5, 0, *arr) < 0
