"""ctypes wrapper of oracle/fish_cpu.c (the OpenMP C restatement).  TEST INFRASTRUCTURE / CPU BASELINE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libfishcpu.so")


class Opts(C.Structure):
    _fields_ = [("dim", C.c_int), ("m", C.c_int * 3), ("L", C.c_double * 3), ("c", C.c_double * 3),
                ("problem", C.c_int), ("gonboundary", C.c_int), ("levels", C.c_int), ("cycle", C.c_int),
                ("smoother_ksp", C.c_int), ("smoother_pc", C.c_int), ("smooth_its", C.c_int),
                ("emin", C.c_double), ("emax", C.c_double), ("rtol", C.c_double), ("max_it", C.c_int),
                ("threads", C.c_int)]


class Result(C.Structure):
    _fields_ = [("its", C.c_int), ("nlevels", C.c_int), ("threads", C.c_int), ("fnorm0", C.c_double),
                ("fnorm1", C.c_double), ("errinf", C.c_double), ("err2h", C.c_double), ("seconds", C.c_double),
                ("nhist", C.c_int), ("hist", C.c_double * 256)]


_lib = None


def load():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "fish_cpu.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", HERE, "libfishcpu.so"])
        _lib = C.CDLL(LIB)
        _lib.fishcpu_solve.restype = C.c_int
        _lib.fishcpu_solve.argtypes = [C.POINTER(Opts), C.POINTER(Result), C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.fishcpu_stream_triad_gbs.restype = C.c_double
        _lib.fishcpu_stream_triad_gbs.argtypes = [C.c_long, C.c_int, C.c_int]
    return _lib


PROBLEMS = {"manupoly": 0, "manuexp": 1, "zero": 2}


def solve(dim=3, refine=3, problem="manuexp", c=(1.0, 1.0, 1.0), L=(1.0, 1.0, 1.0), gonboundary=True, levels=0,
          cycle="v", smoother_ksp="chebyshev", smoother_pc="jacobi", smooth_its=2, eig=None, rtol=1e-5,
          max_it=10000, threads=0, want_arrays=False):
    lib = load()
    m = 1 + (2 ** refine) * 2
    o = Opts()
    o.dim = dim
    for d in range(3):
        o.m[d] = m if d < dim else 1
        o.L[d] = L[d]
        o.c[d] = c[d]
    o.problem = PROBLEMS[problem]
    o.gonboundary = int(gonboundary)
    o.levels = levels
    o.cycle = 1 if cycle == "v" else 2
    o.smoother_ksp = 0 if smoother_ksp == "chebyshev" else 1
    o.smoother_pc = 0 if smoother_pc == "jacobi" else 1
    o.smooth_its = smooth_its
    o.emin, o.emax = (eig if eig else (0.0, 0.0))
    o.rtol = rtol
    o.max_it = max_it
    o.threads = threads or 0
    r = Result()
    n = m ** dim
    u = b = y = None
    if want_arrays:
        u, b, y = np.empty(n), np.empty(n), np.empty(n)
    rc = lib.fishcpu_solve(C.byref(o), C.byref(r), u.ctypes.data if u is not None else None,
                           b.ctypes.data if b is not None else None, y.ctypes.data if y is not None else None)
    if rc:
        raise RuntimeError("fishcpu_solve failed with code %d" % rc)
    out = {"its": r.its, "nlevels": r.nlevels, "threads": r.threads, "fnorm0": r.fnorm0, "fnorm1": r.fnorm1,
           "errinf": r.errinf, "err2h": r.err2h, "seconds": r.seconds, "history": [r.hist[i] for i in range(r.nhist)],
           "n": n, "m": m}
    if want_arrays:
        out.update(u=u, b=b, y=y)
    return out


def timed_solve(refine=7, levels=6, rtol=1e-10, threads=None, repeats=1):
    """One timed 3-D manuexp solve with the BASELINE options (Chebyshev(2)/Jacobi V-cycle)."""
    best = None
    for _ in range(repeats):
        r = solve(dim=3, refine=refine, levels=levels, rtol=rtol, threads=threads or 0)
        if best is None or r["seconds"] < best["seconds"]:
            best = r
    return best


def stream_triad_gbs(n=1 << 26, reps=5, threads=0):
    return load().fishcpu_stream_triad_gbs(n, reps, threads)
