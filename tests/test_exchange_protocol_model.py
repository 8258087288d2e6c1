"""Exhaustive model check of the fused ghost-exchange protocol (p4pdes_b200/csrc/comm.h HaloPort).

The GPU tests show that the protocol gives the right bits on the runs that were made; this test explores EVERY
interleaving of a small abstract model -- 3 ranks in a line (so the middle rank has both neighbours), each running the
same kernel sequence in stream order -- and checks the property the design relies on: every ghost-plane read sees
exactly the value the sequential program expects (the neighbour's latest push of that vector before the reading
kernel), i.e. data has arrived (no stale read) and has not been overwritten by a later push (no early overwrite).

Model of one kernel on one rank (what the CUDA code does, comm.h):
  * per slab side that has a neighbour, a group of boundary CTAs: read the rank's exchange count e, wait until BOTH
    neighbours' flags are >= e (port_wait), read the ghost plane of the consumed vector, store the boundary plane of
    the produced vector into the neighbour's ghost plane (port_store), arrive (port_signal);
  * as soon as all boundary groups of a pushing kernel have arrived -- the interior may still be running --
    flags of both neighbours := e + 1, count := e + 1;
  * interior CTAs run concurrently and touch no ghost plane; the next kernel of the rank starts when everything of
    this one is finished (stream order).
Kernel sequences are the ones mg.cu issues (V cycle + CG updates), including the x, t, x pattern of the post-smoother
and producers that read no ghosts.  The negative tests show the checker sees real bugs: dropping the wait, or a kernel
that reads the ghosts of the vector it pushes, is reported.
"""
import sys

import pytest

# a kernel = (name, vector whose ghosts it reads or None, vector it pushes or None)
V_CYCLE_AND_CG = [
    ("r_update", None, "b"),        # r -= a w            pushes b (= r)
    ("cheb_zero", "b", "x"),        # zero-guess smoother  reads b ghosts, pushes x
    ("residual", "x", "t"),
    ("restrict", "t", "cb"),        # coarse right-hand side (pushed on the coarse level)
    ("c_cheb_zero", "cb", "cx"),
    ("c_cheb_next", "cx", "cx2"),   # stand-in for the rest of the coarse sweep; result consumed by the prolongation
    ("prolong", "cx2", "x"),
    ("cheb_first", "x", "t"),
    ("cheb_next", "t", "x"),        # x, t, x: the same vector two pushes apart
    ("xp_update", None, "p"),       # producer that reads no ghosts
    ("apply_dot", "p", None),       # consumer that pushes nothing
    ("r_update", None, "b"),
    ("cheb_zero", "b", "x"),
]


def explore(seq, nranks=3, wait=True, max_states=4_000_000):
    """DFS over all interleavings.  Returns None if no ghost read can go wrong, else a description of a bad read."""
    kernels = list(seq)
    nk = len(kernels)
    # expected[n][v] = index of the kernel whose push of v a read in kernel n must see
    expected = []
    latest = {}
    for n, (name, rd, push) in enumerate(kernels):
        expected.append(dict(latest))
        if push is not None:
            latest[push] = n
    sides = {r: [s for s in (0, 1) if (r > 0 if s == 0 else r < nranks - 1)] for r in range(nranks)}
    vectors = sorted({k[2] for k in kernels if k[2]})
    vidx = {v: i for i, v in enumerate(vectors)}

    # state: per rank (kernel index, phase of the lower boundary group, phase of the upper one, interior done,
    #                  signalled, exchange count latched by each group),
    #        exchange count per rank, flags per rank (from lower, from upper neighbour),
    #        ghost version per (rank, side, vector) = index of the kernel whose push it holds
    # group phases: 0 not started, 1 latched e, 2 waited, 3 ghost read, 4 pushed and arrived, 5 = no neighbour there
    def fresh(r):
        return (0 if 0 in sides[r] else 5, 0 if 1 in sides[r] else 5, 0, 0, -1, -1)

    ranks0 = tuple((0,) + fresh(r) for r in range(nranks))
    init = (ranks0, tuple(0 for _ in range(nranks)), tuple((0, 0) for _ in range(nranks)),
            tuple(tuple(tuple(-1 for _ in vectors) for _ in (0, 1)) for _ in range(nranks)))
    seen = set()
    stack = [init]
    while stack:
        st = stack.pop()
        if st in seen:
            continue
        seen.add(st)
        if len(seen) > max_states:
            raise RuntimeError("state space larger than expected")
        ranks, epochs, flags, ghosts = st
        for r in range(nranks):
            k, p0, p1, interior, signalled, e0, e1 = ranks[r]
            if k >= nk:
                continue
            name, rd, push = kernels[k]

            def with_rank(newr, ep=epochs, fl=flags, gh=ghosts):
                rs = list(ranks)
                rs[r] = newr
                return tuple(rs), ep, fl, gh

            if not interior:                                  # interior CTAs: no ghost plane involved
                stack.append(with_rank((k, p0, p1, 1, signalled, e0, e1)))
            for s in sides[r]:                                # boundary groups
                ph, el = (p0, e0) if s == 0 else (p1, e1)

                def upd(newph, newe=None):
                    ne = el if newe is None else newe
                    return ((k, newph, p1, interior, signalled, ne, e1) if s == 0
                            else (k, p0, newph, interior, signalled, e0, ne))

                if ph == 0:                                   # latch the exchange count
                    stack.append(with_rank(upd(1, epochs[r])))
                elif ph == 1:                                 # port_wait: BOTH neighbours' flags >= e
                    if not wait or all(flags[r][t] >= el for t in sides[r]):
                        stack.append(with_rank(upd(2)))
                elif ph == 2:                                 # ghost read of the consumed vector on this side
                    if rd is not None and rd in vidx:
                        want, got = expected[k].get(rd, -1), ghosts[r][s][vidx[rd]]
                        if want >= 0 and got != want:
                            return ("rank %d, kernel %d (%s), side %d: ghost of %s holds the push of kernel %d, expected %d"
                                    % (r, k, name, s, rd, got, want))
                    stack.append(with_rank(upd(3)))
                elif ph == 3:                                 # port_store into the neighbour's ghost plane, arrive
                    gh = ghosts
                    if push is not None:
                        q = r - 1 if s == 0 else r + 1
                        g = [[list(y) for y in x] for x in ghosts]
                        g[q][1 - s][vidx[push]] = k           # my lower boundary plane is the neighbour's UPPER ghost
                        gh = tuple(tuple(tuple(y) for y in x) for x in g)
                    stack.append(with_rank(upd(4), gh=gh))
            done = p0 in (4, 5) and p1 in (4, 5)
            if done and not signalled:                        # port_signal: as soon as the boundary groups are done
                ep, fl = epochs, flags
                if push is not None:
                    e = epochs[r] + 1
                    ep = tuple(e if q == r else epochs[q] for q in range(nranks))
                    f = [list(x) for x in flags]
                    if r > 0:
                        f[r - 1][1] = e                       # the lower neighbour's "from upper" flag
                    if r < nranks - 1:
                        f[r + 1][0] = e
                    fl = tuple(tuple(x) for x in f)
                stack.append(with_rank((k, p0, p1, interior, 1, e0, e1), ep, fl))
            if done and signalled and interior:               # stream order: the next kernel starts now
                stack.append(with_rank((k + 1,) + fresh(r)))
    return None


def test_every_interleaving_of_the_v_cycle_sequence_reads_the_expected_ghosts():
    assert explore(V_CYCLE_AND_CG, nranks=3) is None
    assert explore(V_CYCLE_AND_CG, nranks=2) is None


def test_repeated_cycles_are_safe():
    assert explore(V_CYCLE_AND_CG[1:11] * 2, nranks=3) is None


def test_checker_sees_a_missing_wait():
    bad = explore(V_CYCLE_AND_CG, wait=False)
    assert bad is not None and "expected" in bad


def test_the_one_pattern_the_protocol_cannot_serve_is_detected_and_never_issued():
    """A kernel that reads the ghosts of the very vector it pushes would overwrite a neighbour's ghost plane before the
    neighbour's copy of the same kernel has read it.  The stencil kernels never do this (operand and result are distinct
    buffers; in-place updates are element-wise and read no ghosts).  Two consecutive pushes of one vector by kernels that
    do not read it are fine."""
    assert explore([("a", None, "x"), ("b", "x", "x"), ("c", "x", None)]) is not None
    assert explore([("a", None, "x"), ("b", None, "x"), ("c", "x", None)]) is None
    assert all(rd is None or rd != push for _, rd, push in V_CYCLE_AND_CG)
