"""Micro-benchmark of the callback / SpMV kernels at the BASELINE config sizes (C4: 2049^2, C5: 2048^2 x 2 dof) and at
sizes that exceed the 126 MB L2, with CUDA events on the launching stream.  Prints one JSON line per kernel.

    python profiles/microbench.py > profiles/r01_microbench.jsonl
    python profiles/microbench.py --only-heat      (the kernels added late in round 2)
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from p4pdes_b200 import callbacks as cb  # noqa: E402
from p4pdes_b200 import lib as L  # noqa: E402
from p4pdes_b200.fish import Context  # noqa: E402

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6553.9


def timeit(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, size, bytes_alg, ms, note=""):
    gbs = bytes_alg / ms / 1e6
    print(json.dumps({"kernel": name, "size": size, "alg_bytes": bytes_alg, "ms": ms, "GBs": gbs,
                      "frac_of_hbm_peak": gbs / PEAK, "working_set_MB": bytes_alg / 1e6, "note": note}), flush=True)


def heat_and_poisson(ctx):
    """Round 2, late additions: heat.c's kernels (c/ch5/heat.c:141-208) and the Poisson matrix fill (minimal.c:142-145)."""
    for mx, my in ((2049, 2048), (8193, 8192)):
        n = mx * my
        u = torch.randn(n, dtype=torch.float64, device="cuda")
        G = ctx.empty(n)
        note = "L2 resident" if 16 * n < 100e6 else "exceeds L2"
        ms = timeit(lambda: ctx.heat_rhs(mx, my, 1.0, u, G))
        report("heat_rhs", "%d x %d" % (mx, my), 16.0 * n, ms, note)
        ms = timeit(lambda: ctx.heat_jac_apply(mx, my, 1.0, 1000.0, u, G))
        report("heat_jac_apply", "%d x %d" % (mx, my), 16.0 * n, ms, note)
        del u, G
    for m in (2049, 4097):
        n = m * m
        vals = ctx.empty(9 * n)
        ms = timeit(lambda: ctx.poisson_stencil9(m, m, 1.0, 1.0, 1.0, 1.0, vals))
        report("poisson_stencil9 (fill)", "%d^2" % m, 72.0 * n, ms, "write-only, %s" % ("L2 resident" if 72 * n < 100e6 else "exceeds L2"))
        del vals


def main():
    ctx = Context()
    if "--only-heat" in sys.argv:
        heat_and_poisson(ctx)
        return
    for m in (2049, 8193):
        n = m * m
        g = cb.minimal_g(ctx, m, m, "catenoid", 1.0, 1.1)
        u = g.clone() * 0.9
        FF = ctx.empty(n)
        ms = timeit(lambda: cb.minimal_form_function(ctx, m, m, u, g, -0.5, out=FF))
        report("minimal_function", "%d^2" % m, 16.0 * n, ms, "L2 resident" if 16 * n < 100e6 else "exceeds L2")
        del g, u, FF
    for m in (2048, 6144):
        n = m * m
        Y = cb.pattern_initial_state(ctx, m, m)
        Yd = Y.clone()
        F = ctx.empty(2 * n)
        ms = timeit(lambda: cb.pattern_ifunction(ctx, m, m, Y, Yd, out=F))
        report("pattern_ifunction", "%d^2 x2" % m, 48.0 * n, ms, "L2 resident" if 48 * n < 100e6 else "exceeds L2")
        ms = timeit(lambda: cb.pattern_rhs_function(ctx, m, m, Y, out=F))
        report("pattern_rhsfunction", "%d^2 x2" % m, 32.0 * n, ms)
        ms = timeit(lambda: cb.pattern_ijacobian_mult(ctx, m, m, 0.2, Y, out=F))
        report("pattern_ijacobian_mult", "%d^2 x2" % m, 32.0 * n, ms)
        del Y, Yd, F
    # SELL SpMV on the assembled 5-point fish Jacobian (2049^2) and 7-point (257^3): built with scipy on the host
    from oracle import fish_oracle as fo
    for og, label in ((fo.Grid(2, (2049, 2049, 1)), "fish 2-D 2049^2 5-pt"), (fo.Grid(3, (257, 257, 257)), "fish 3-D 257^3 7-pt")):
        A = fo.jacobian(og)
        A.sort_indices()
        S = cb.SellMatrix(ctx, A.indptr, A.indices, A.data)
        x = torch.randn(og.n, dtype=torch.float64, device="cuda")
        y = ctx.empty(og.n)
        ms = timeit(lambda: S.mult(x, y))
        report("sell_spmv", label, 12.0 * S.padded_nnz + 16.0 * og.n, ms, "nnz %d padded %d" % (S.nnz, S.padded_nnz))
        # the matrix-free kernel for the same operator
        g = L.make_grid(og.dim, og.m[:og.dim])
        ms2 = timeit(lambda: ctx.stencil_apply(g, x, y))
        report("stencil_apply (matrix-free)", label, 16.0 * og.n, ms2, "same operator, %.1fx faster than SpMV" % (ms / ms2))
        S.close()
        del x, y
    # ---- assembled 9-point Jacobian of minimal.c (stencil9 layout) and its column-indexed SELL-32 copy ----------------
    import ctypes as C
    from p4pdes_b200.minimal import stencil9_to_csr
    for m in (2049, 4097):
        n = m * m
        g = cb.minimal_g(ctx, m, m, "catenoid", 1.0, 1.1)
        u = g.clone() * 0.9
        F0, vals = ctx.empty(n), ctx.empty(9 * n)
        ctx.minimal_function(m, m, -0.5, u, g, F0)
        ms = timeit(lambda: ctx.minimal_jacobian_fd(m, m, -0.5, u, g, F0, vals), reps=5, warm=2)
        # per colour: perturb (16 N), residual (16 N + g), extract (reads u, F0, Fp: 24 N, writes N values);
        # nine colours fill the 9 N coefficients once (72 N) after a 72 N memset
        report("minimal_jacobian_fd (9 colours x 3 kernels)", "%d^2" % m, (9 * (16 + 24 + 24) + 72 + 72) * float(n), ms,
               "%.2f ms per assembled Jacobian" % ms)
        x, b, y = torch.randn(n, dtype=torch.float64, device="cuda"), torch.randn(n, dtype=torch.float64, device="cuda"), ctx.empty(n)
        ms1 = timeit(lambda: ctx.stencil9_apply(m, m, vals, x, y))
        report("stencil9_apply (y = A x)", "%d^2" % m, 88.0 * n, ms1)
        ms2 = timeit(lambda: ctx.stencil9_lin(m, m, vals, x, b, y, 0.3, 0.7, 0.4, True, y))
        report("stencil9_lin (Chebyshev+Jacobi step)", "%d^2" % m, 104.0 * n, ms2)
        if m == 2049:
            rp, ci, d = stencil9_to_csr(ctx.to_host(vals), m, m)
            S = cb.SellMatrix(ctx, rp, ci, d)
            ms3 = timeit(lambda: S.mult(x, y))
            report("sell_spmv (same Jacobian, column-indexed copy)", "%d^2" % m, 12.0 * S.padded_nnz + 16.0 * n, ms3,
                   "stencil9 layout is %.2fx faster" % (ms3 / ms1))
            S.close()
        del g, u, F0, vals, x, b, y
    # ---- pattern.c stage Jacobian, matrix-free, and the periodic transfer ------------------------------------------------
    PAR = (2.5, 8.0e-5, 4.0e-5, 0.024, 0.06)
    for m in (2048, 6144):
        n = m * m
        Y = cb.pattern_initial_state(ctx, m, m)
        X, B, out = torch.randn(2 * n, dtype=torch.float64, device="cuda"), torch.randn(2 * n, dtype=torch.float64, device="cuda"), ctx.empty(2 * n)
        ms = timeit(lambda: ctx.pattern_jac_apply(m, *PAR, 0.2, Y, X, out))
        report("pattern_jac_apply (J X, matrix-free)", "%d^2 x2" % m, 48.0 * n, ms)
        ms = timeit(lambda: ctx.pattern_jac_lin(m, *PAR, 0.2, Y, X, B, out, 0.3, 0.7, 0.4, True, out))
        report("pattern_jac_lin (Chebyshev+Jacobi step)", "%d^2 x2" % m, 80.0 * n, ms)
        xc = ctx.empty(n // 2)
        ms = timeit(lambda: ctx.pattern_restrict(m // 2, m // 2, X, xc))
        report("pattern_restrict", "%d^2 x2" % m, 16.0 * n + 4.0 * n, ms)
        ms = timeit(lambda: ctx.pattern_prolong_add(m // 2, m // 2, xc, out))
        report("pattern_prolong_add", "%d^2 x2" % m, 32.0 * n + 4.0 * n, ms)
        del Y, X, B, out, xc


if __name__ == "__main__":
    main()
