"""c/ch2/loadsolve.c on the device: load A (and b) from PETSc binary files, solve A x = b with a Krylov method.

    ./tri -ksp_view_mat binary:A.dat -ksp_view_rhs binary:b.dat ; ./loadsolve -fA A.dat -fb b.dat      (loadsolve.c:4-8)

becomes  loadsolve_main("-fA A.dat -fb b.dat -pc_type jacobi", ctx):  [PETSc] MatLoad / VecLoad = p4pdes_b200/petscbin.py,
the matrix goes to the device as SELL-32 (p4b_sell_create, [PETSc] MATSELL), KSPSolve = the Krylov hosts of minimal.py
(KSPGMRES(30), PETSc's default, or KSPCG) over the SpMV and Vec kernels, -pc_type none | jacobi.  PETSc's default
preconditioner (ILU(0)) is sequential and not provided: it must be replaced by name, as for every driver here.
tri_system restates c/ch2/tri.c:31-52 (the system the reference's own test feeds loadsolve, c/ch2/makefile:34-36), so that
the golden c/ch2/output/loadsolve.test1 can be reproduced without PETSc writing the files.
The -ksp_view_* options print PETSc's ASCII views (small systems; the layout of the golden)."""
from __future__ import annotations

import math
import shlex
from dataclasses import dataclass, field

import numpy as np

from . import petscbin
from .minimal import cg, gmres


def tri_system(m=4):
    """c/ch2/tri.c:31-52: A = tridiag(-1, 3, -1), xexact_i = exp(cos i), b = A xexact.  CSR arrays + the two vectors."""
    rowptr, colind, vals = [0], [], []
    for i in range(m):
        for j, v in ((i - 1, -1.0), (i, 3.0), (i + 1, -1.0)):
            if 0 <= j < m:
                colind.append(j)
                vals.append(v)
        rowptr.append(len(colind))
    xexact = np.exp(np.cos(np.arange(m, dtype=np.float64)))
    b = 3.0 * xexact
    b[1:] -= xexact[:-1]
    b[:-1] -= xexact[1:]
    return (np.array(rowptr, np.int32), np.array(colind, np.int32), np.array(vals)), b, xexact


def write_system(pathA, pathb, csr, b):
    with open(pathA, "wb") as fh:
        petscbin.write_mat(fh, *csr)
    if pathb:
        with open(pathb, "wb") as fh:
            petscbin.write_vec(fh, b)


def fmt_g(v):
    """%g as [PETSc] prints it: integer-valued reals get a trailing '.'."""
    s = "%g" % v
    return s if any(c in s for c in ".en") else s + "."


@dataclass
class LoadSolveReport:
    lines: list = field(default_factory=list)
    its: int = 0
    reason: str = ""
    x: object = None
    n: int = 0


def loadsolve_main(argv, ops, echo=False) -> LoadSolveReport:
    """ops: p4pdes_b200.fish.Context (or the NumPy stand-in of the tests)."""
    if isinstance(argv, str):
        argv = shlex.split(argv)
    o = {"-fA": "", "-fb": "", "-ksp_type": "gmres", "-pc_type": "", "-ksp_rtol": "1e-5", "-ksp_max_it": "10000",
         "-ksp_gmres_restart": "30"}
    flags = {"-verbose": False, "-ksp_converged_reason": False, "-ksp_view_mat": False, "-ksp_view_rhs": False,
             "-ksp_view_solution": False}
    i = 0
    while i < len(argv):
        if argv[i] in flags:
            flags[argv[i]] = True
            i += 1
        elif argv[i] in o:
            o[argv[i]] = argv[i + 1]
            i += 2
        else:
            raise ValueError("unknown or unsupported option %s" % argv[i])
    rep = LoadSolveReport()

    def out(s):
        rep.lines.append(s)
        if echo:
            print(s)

    if not o["-fA"]:
        raise ValueError("no input matrix provided ... ending  (usage: loadsolve -fA A.dat)")          # loadsolve.c:47-50
    if o["-ksp_type"] not in ("gmres", "cg"):
        raise ValueError("-ksp_type %s: gmres and cg are provided" % o["-ksp_type"])
    if o["-pc_type"] not in ("none", "jacobi"):
        raise ValueError("PETSc's default PC (ILU(0) on one rank) is sequential and not provided on the device: pass "
                         "-pc_type none or -pc_type jacobi")
    if flags["-verbose"]:
        out("reading matrix from %s ..." % o["-fA"])
    recs = petscbin.read_file(o["-fA"])
    if len(recs) != 1 or not isinstance(recs[0], tuple):
        raise ValueError("%s does not hold one Mat record" % o["-fA"])
    (m, n), (rowptr, colind, vals) = recs[0]
    if flags["-verbose"]:
        out("matrix has size m x n = %d x %d ..." % (m, n))
    if m != n:
        raise ValueError("only works for square matrices")                                             # :66-68
    if o["-fb"]:
        if flags["-verbose"]:
            out("reading vector from %s ..." % o["-fb"])
        recs = petscbin.read_file(o["-fb"])
        if len(recs) != 1 or not isinstance(recs[0], np.ndarray):
            raise ValueError("%s does not hold one Vec record" % o["-fb"])
        b = recs[0]
        if b.size != m:
            raise ValueError("size of matrix and vector do not match")                                 # :80-82
    else:
        if flags["-verbose"]:
            out("right-hand-side vector b not provided ... using zero vector of length %d" % m)
        b = np.zeros(m)
    A = ops.sell_matrix(rowptr, colind, vals)
    db, dx = ops.from_host(b), ops.empty(m)
    if o["-pc_type"] == "jacobi":
        diag = np.zeros(m)
        rows = np.repeat(np.arange(m), np.diff(rowptr))
        on = colind == rows
        diag[rows[on]] = vals[on]
        if np.any(diag == 0.0):
            raise ValueError("-pc_type jacobi: zero (or missing) diagonal entry")
        dinv = ops.from_host(1.0 / diag)
        precond = lambda r, z: ops.pointwise_mult(dinv, r, z)
    else:
        precond = lambda r, z: ops.copy(r, z)
    mult = lambda v, w: A.mult(v, w)
    rtol, max_it = float(o["-ksp_rtol"]), int(o["-ksp_max_it"])
    if float(np.max(np.abs(b))) == 0.0:                    # [PETSc]: a zero right-hand side converges at once, x = 0
        ops.set(0.0, dx)
        rep.its, rep.reason = 0, "CONVERGED_ATOL"
    else:
        if o["-ksp_type"] == "gmres":
            k = gmres(ops, mult, db, dx, precond, rtol, restart=int(o["-ksp_gmres_restart"]), max_it=max_it)
        else:
            k = cg(ops, mult, db, dx, precond, rtol, max_it=max_it)
        rep.its, rep.reason = k.its, k.reason
    x = np.asarray(ops.to_host(dx), dtype=np.float64)
    # [PETSc] KSPSolve's viewers, in its order: matrix, right-hand side, (solve,) reason, solution
    if flags["-ksp_view_mat"]:
        out("Mat Object: 1 MPI process")
        out("  type: sellcuda")
        for r in range(m):
            out("row %d:" % r + "".join(" (%d, %s) " % (colind[q], fmt_g(vals[q])) for q in range(rowptr[r], rowptr[r + 1])))
    if flags["-ksp_view_rhs"]:
        out("Vec Object: 1 MPI process")
        out("  type: cuda")
        for v in b:
            out(fmt_g(v))
    if flags["-ksp_converged_reason"]:
        out("Linear solve %s due to %s iterations %d" % ("converged" if rep.reason.startswith("CONV") else "did not converge",
                                                         rep.reason, rep.its))
    if flags["-ksp_view_solution"]:
        out("Vec Object: 1 MPI process")
        out("  type: cuda")
        for v in x:
            out(fmt_g(v))
    if not math.isfinite(float(np.sum(x))):
        rep.reason = rep.reason or "DIVERGED_NANORINF"
    rep.x, rep.n = x, m
    if hasattr(A, "close"):
        A.close()
    return rep
