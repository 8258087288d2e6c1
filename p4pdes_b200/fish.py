"""Host-side mirror of c/ch6/fish.c on top of the C ABI (include/p4b200.h).

`fish_main(argv)` takes the reference's own command line (`-fsh_dim 3 -da_refine 4 -pc_type mg ...`,
c/ch6/fish.c:154-185 and the PETSc options listed in SURVEY.md Appendix B), runs the same sequence
as fish.c:main (DMDA -> InitialState -> SNESSolve[KSPONLY: F(u0), KSPCG + PCMG] -> error norms) on
the GPU, and prints the same lines the reference prints, so tests can diff text like c/testit.sh does.

PyTorch is used for device tensors, streams and torch.distributed only.
"""
from __future__ import annotations

import ctypes as C
import math
import shlex
from dataclasses import dataclass, field

import torch

from . import lib as L


class Context:
    """p4b_ctx: one CUDA device + stream (+ NCCL communicator when torch.distributed is initialised)."""

    def __init__(self, device: int | None = None, stream: "torch.cuda.Stream | None" = None, distributed=False):
        self.lib = L.load()
        if not torch.cuda.is_available():
            raise L.P4BError("no CUDA device: p4pdes_b200 has no CPU path")
        self.device = torch.cuda.current_device() if device is None else int(device)
        torch.cuda.set_device(self.device)
        self.stream = stream or torch.cuda.current_stream(self.device)
        self.h = C.c_void_p()
        L.check(self.lib.p4b_ctx_create(self.device, C.c_void_p(self.stream.cuda_stream), C.byref(self.h)))
        self.rank, self.nranks = 0, 1
        if distributed:
            self.init_distributed()

    def init_distributed(self):
        """Ship the NCCL unique id from rank 0 over torch.distributed and build the library's communicator."""
        import torch.distributed as dist
        if not dist.is_initialized():
            raise L.P4BError("torch.distributed is not initialised")
        self.rank, self.nranks = dist.get_rank(), dist.get_world_size()
        if self.nranks == 1:
            return
        buf = C.create_string_buffer(128)
        if self.rank == 0:
            L.check(self.lib.p4b_comm_unique_id(buf))
        obj = [bytes(buf.raw) if self.rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        idbuf = C.create_string_buffer(obj[0], 128)
        L.check(self.lib.p4b_comm_init(self.h, idbuf, self.rank, self.nranks))

    def sync(self):
        L.check(self.lib.p4b_ctx_sync(self.h))

    def empty(self, n):
        return torch.empty(int(n), dtype=torch.float64, device="cuda:%d" % self.device)

    def zeros(self, n):
        return torch.zeros(int(n), dtype=torch.float64, device="cuda:%d" % self.device)

    def close(self):
        if self.h:
            self.lib.p4b_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- building blocks (single slab) ----
    def stencil_apply(self, g, u, y):
        L.check(self.lib.p4b_stencil_apply(self.h, C.byref(g), u.data_ptr(), y.data_ptr()))

    def stencil_residual(self, g, b, u, r):
        L.check(self.lib.p4b_stencil_residual(self.h, C.byref(g), b.data_ptr(), u.data_ptr(), r.data_ptr()))

    def cheb_jacobi(self, g, emin, emax, its, zero_guess, b, x, work):
        L.check(self.lib.p4b_cheb_jacobi(self.h, C.byref(g), emin, emax, its, int(zero_guess), b.data_ptr(),
                                         x.data_ptr(), work.data_ptr()))

    def restrict(self, gf, rf, bc):
        L.check(self.lib.p4b_restrict(self.h, C.byref(gf), rf.data_ptr(), bc.data_ptr()))

    def prolong_add(self, gf, xc, xf):
        L.check(self.lib.p4b_prolong_add(self.h, C.byref(gf), xc.data_ptr(), xf.data_ptr()))

    def residual_restrict(self, gf, b, x, bc):
        L.check(self.lib.p4b_residual_restrict(self.h, C.byref(gf), b.data_ptr(), x.data_ptr(), bc.data_ptr()))

    def dot(self, x, y):
        r = C.c_double()
        L.check(self.lib.p4b_vec_dot(self.h, x.numel(), x.data_ptr(), y.data_ptr(), C.byref(r)))
        return r.value

    def norm2(self, x):
        r = C.c_double()
        L.check(self.lib.p4b_vec_norm2(self.h, x.numel(), x.data_ptr(), C.byref(r)))
        return r.value

    def wrms2(self, x, y, atol, rtol):
        """sum_i ((x_i - y_i) / (atol + rtol max(|x_i|, |y_i|)))^2  ([PETSc] TSErrorWeightedNorm2)."""
        r = C.c_double()
        L.check(self.lib.p4b_vec_wrms2(self.h, x.numel(), x.data_ptr(), y.data_ptr(), atol, rtol, C.byref(r)))
        return r.value

    def norminf(self, x):
        r = C.c_double()
        L.check(self.lib.p4b_vec_norminf(self.h, x.numel(), x.data_ptr(), C.byref(r)))
        return r.value

    def axpy(self, a, x, y):
        L.check(self.lib.p4b_vec_axpy(self.h, x.numel(), a, x.data_ptr(), y.data_ptr()))

    def aypx(self, a, x, y):
        L.check(self.lib.p4b_vec_aypx(self.h, x.numel(), a, x.data_ptr(), y.data_ptr()))

    def axpby(self, a, x, b, y, out):
        """out = a x + b y; x or y may be None (treated as zero) and may alias out."""
        ptr = lambda t: t.data_ptr() if t is not None else None
        L.check(self.lib.p4b_vec_axpby(self.h, out.numel(), a, ptr(x), b, ptr(y), out.data_ptr()))

    def copy(self, x, y):
        L.check(self.lib.p4b_vec_copy(self.h, x.numel(), x.data_ptr(), y.data_ptr()))

    def set(self, a, y):
        L.check(self.lib.p4b_vec_set(self.h, y.numel(), a, y.data_ptr()))

    def vi_inactive_mask(self, u, lower, F, mask):
        L.check(self.lib.p4b_vi_inactive_mask(self.h, u.numel(), u.data_ptr(), lower.data_ptr(), F.data_ptr(), mask.data_ptr()))

    def pointwise_mult(self, x, y, out):
        L.check(self.lib.p4b_vec_pointwise_mult(self.h, x.numel(), x.data_ptr(), y.data_ptr(), out.data_ptr()))

    def pointwise_max(self, x, y, out):
        L.check(self.lib.p4b_vec_pointwise_max(self.h, x.numel(), x.data_ptr(), y.data_ptr(), out.data_ptr()))

    def to_host(self, t):
        self.sync()
        return t.detach().cpu().numpy()

    def from_host(self, a):
        import numpy as np
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).ravel()).to("cuda:%d" % self.device)

    # ---- 2-D assembled-Jacobian path (minimal.c; include/p4b200.h "assembled Jacobians") ----
    def grid2d(self, mx, my):
        return L.make_grid(2, (mx, my), (1.0, 1.0, 1.0), (1.0, 1.0, 1.0))

    def initial_state2d(self, grid, g, u):
        self.initial_state(grid, g, True, u)

    def minimal_sample(self, mx, my, problem, tent_H, catenoid_c, g):
        L.check(self.lib.p4b_minimal_sample(self.h, mx, my, problem, tent_H, catenoid_c, g.data_ptr()))

    def minimal_function(self, mx, my, q, u, g, FF):
        L.check(self.lib.p4b_minimal_function(self.h, mx, my, q, u.data_ptr(), g.data_ptr(), FF.data_ptr()))

    def minimal_jacobian_fd(self, mx, my, q, u, g, F0, vals):
        L.check(self.lib.p4b_minimal_jacobian_fd(self.h, mx, my, q, u.data_ptr(), g.data_ptr(), F0.data_ptr(),
                                                 vals.data_ptr()))

    def heat_rhs(self, mx, my, D0, u, G):
        L.check(self.lib.p4b_heat_rhs(self.h, mx, my, D0, u.data_ptr(), G.data_ptr()))

    def heat_jac_apply(self, mx, my, D0, shift, X, JX):
        L.check(self.lib.p4b_heat_jac_apply(self.h, mx, my, D0, shift, X.data_ptr(), JX.data_ptr()))

    def sell_matrix(self, rowptr, colind, vals):
        """An assembled matrix on this device ([PETSc] MATSELL), from host CSR arrays."""
        from .callbacks import SellMatrix
        return SellMatrix(self, rowptr, colind, vals)

    def poisson_stencil9(self, mx, my, Lx, Ly, cx, cy, vals):
        L.check(self.lib.p4b_poisson_stencil9(self.h, mx, my, Lx, Ly, cx, cy, vals.data_ptr()))

    def stencil9_apply(self, mx, my, vals, x, y):
        L.check(self.lib.p4b_stencil9_apply(self.h, mx, my, vals.data_ptr(), x.data_ptr(), y.data_ptr()))

    def stencil9_lin(self, mx, my, vals, u, b, pm1, ca, cb, cg, jacobi, out):
        ptr = lambda t: t.data_ptr() if t is not None else None
        L.check(self.lib.p4b_stencil9_lin(self.h, mx, my, vals.data_ptr(), u.data_ptr(), ptr(b), ptr(pm1), ca, cb, cg,
                                          int(bool(jacobi)), out.data_ptr()))

    def stencil9_gershgorin(self, mx, my, vals, work):
        r = C.c_double()
        L.check(self.lib.p4b_stencil9_gershgorin(self.h, mx, my, vals.data_ptr(), work.data_ptr(), C.byref(r)))
        return r.value

    def inject2d(self, cmx, cmy, ufine, ucoarse):
        L.check(self.lib.p4b_inject2d(self.h, cmx, cmy, ufine.data_ptr(), ucoarse.data_ptr()))

    def dense_matvec(self, n, Ainv, b, x):
        L.check(self.lib.p4b_dense_matvec(self.h, n, Ainv.data_ptr(), b.data_ptr(), x.data_ptr()))

    # ---- pattern.c implicit stage equation (include/p4b200.h "pattern.c implicit stage equation") ----
    def pattern_initial_state(self, mx, my, Lside, Y):
        L.check(self.lib.p4b_pattern_initial_state(self.h, mx, my, Lside, Y.data_ptr()))

    def rander48(self, n, seed=0x12345678):
        """n values of the VecSetRandom stream ([PETSc] rander48 restated, include/p4b200.h) as a device tensor."""
        state = C.c_ulonglong(self.lib.p4b_rander48_seed(seed))
        host = torch.empty(n, dtype=torch.float64)
        L.check(self.lib.p4b_rander48_fill(C.byref(state), n, host.data_ptr()))
        return host.to(self.device)

    def pattern_initial_state_noisy(self, mx, my, Lside, level, Y):
        noise = self.rander48(2 * mx * my)
        L.check(self.lib.p4b_pattern_initial_state_noisy(self.h, mx, my, Lside, noise.data_ptr(), level, Y.data_ptr()))

    def pattern_ifunction(self, mx, my, Lside, Du, Dv, Y, Ydot, F):
        L.check(self.lib.p4b_pattern_ifunction(self.h, mx, my, Lside, Du, Dv, Y.data_ptr(), Ydot.data_ptr(), F.data_ptr()))

    def pattern_rhsfunction(self, mx, my, phi, kappa, Y, G):
        L.check(self.lib.p4b_pattern_rhsfunction(self.h, mx, my, phi, kappa, Y.data_ptr(), G.data_ptr()))

    def pattern_jac_apply(self, m, Lside, Du, Dv, phi, kappa, shift, Y, X, out):
        L.check(self.lib.p4b_pattern_jac_apply(self.h, m, m, Lside, Du, Dv, phi, kappa, shift,
                                               Y.data_ptr() if Y is not None else None, X.data_ptr(), out.data_ptr()))

    def pattern_jac_lin(self, m, Lside, Du, Dv, phi, kappa, shift, Y, X, b, pm1, ca, cb, cg, jacobi, out):
        ptr = lambda t: t.data_ptr() if t is not None else None
        L.check(self.lib.p4b_pattern_jac_lin(self.h, m, m, Lside, Du, Dv, phi, kappa, shift, ptr(Y), X.data_ptr(), ptr(b),
                                             ptr(pm1), ca, cb, cg, int(bool(jacobi)), out.data_ptr()))

    def pattern_jac_gershgorin(self, m, Lside, Du, Dv, phi, kappa, shift, Y, work):
        r = C.c_double()
        L.check(self.lib.p4b_pattern_jac_gershgorin(self.h, m, m, Lside, Du, Dv, phi, kappa, shift,
                                                    Y.data_ptr() if Y is not None else None, work.data_ptr(), C.byref(r)))
        return r.value

    def pattern_restrict(self, Mx, My, rf, bc):
        L.check(self.lib.p4b_pattern_restrict(self.h, Mx, My, rf.data_ptr(), bc.data_ptr()))

    def pattern_prolong_add(self, Mx, My, xc, xf):
        L.check(self.lib.p4b_pattern_prolong_add(self.h, Mx, My, xc.data_ptr(), xf.data_ptr()))

    def pattern_inject(self, Mx, My, yf, yc):
        L.check(self.lib.p4b_pattern_inject(self.h, Mx, My, yf.data_ptr(), yc.data_ptr()))

    def fish_sample(self, g, problem, f=None, gb=None):
        L.check(self.lib.p4b_fish_sample(self.h, C.byref(g), L.PROBLEMS[problem],
                                         f.data_ptr() if f is not None else None,
                                         gb.data_ptr() if gb is not None else None))

    def initial_state(self, g, gb, gonboundary, u):
        L.check(self.lib.p4b_initial_state(self.h, C.byref(g), gb.data_ptr() if gb is not None else None,
                                           int(gonboundary), u.data_ptr()))

    def poisson_function(self, g, u, f, gb, F):
        L.check(self.lib.p4b_poisson_function(self.h, C.byref(g), u.data_ptr(), f.data_ptr(), gb.data_ptr(),
                                              F.data_ptr()))


def mg_options(levels=0, cycle="v", smoother="chebyshev", smooth_its=2, eig=None, esteig=(0.1, 1.1), fuse=True,
               use_graph=True) -> L.MGOpts:
    o = L.MGOpts()
    L.load().p4b_mg_default_opts(C.byref(o))
    o.levels = int(levels or 0)
    o.cycle = {"v": L.CYCLE_V, "w": L.CYCLE_W}[cycle]
    o.smoother = {"chebyshev": L.SMOOTH_CHEBYSHEV, "richardson": L.SMOOTH_RICHARDSON}[smoother]
    o.smooth_its = int(smooth_its)
    if eig is not None:
        o.emin, o.emax = float(eig[0]), float(eig[1])
    o.est_lo, o.est_hi = float(esteig[0]), float(esteig[1])
    o.fuse = int(bool(fuse))
    o.use_graph = int(bool(use_graph))
    return o


class Multigrid:
    """p4b_mg: the PCMG hierarchy and KSPCG workspace for one grid on one context."""

    def __init__(self, ctx: Context, grid: L.Grid, opts: L.MGOpts | None = None):
        self.ctx, self.grid = ctx, grid
        self.lib = ctx.lib
        self.opts = opts or mg_options()
        self.h = C.c_void_p()
        L.check(self.lib.p4b_mg_create(ctx.h, C.byref(grid), C.byref(self.opts), C.byref(self.h)))
        s, c, n = C.c_int(), C.c_int(), C.c_size_t()
        L.check(self.lib.p4b_mg_local_range(self.h, C.byref(s), C.byref(c), C.byref(n)))
        self.zs, self.zm, self.nlocal = s.value, c.value, n.value

    @property
    def nlevels(self):
        n = C.c_int()
        L.check(self.lib.p4b_mg_nlevels(self.h, C.byref(n)))
        return n.value

    def level_info(self, l):
        m = (C.c_int * 3)()
        e = (C.c_double * 2)()
        L.check(self.lib.p4b_mg_level_info(self.h, l, m, e))
        return tuple(m), tuple(e)

    def apply(self, r, z):
        L.check(self.lib.p4b_mg_apply(self.h, r.data_ptr(), z.data_ptr()))

    def cg_solve(self, b, x, rtol=1e-5, abstol=1e-50, max_it=10000, pc="mg") -> L.KSPResult:
        res = L.KSPResult()
        pcid = {"none": L.PC_NONE, "jacobi": L.PC_JACOBI, "mg": L.PC_MG}[pc]
        L.check(self.lib.p4b_cg_solve(self.h, pcid, b.data_ptr(), x.data_ptr(), rtol, abstol, max_it, C.byref(res)))
        return res

    def cg_solve_host(self, b_host, x_host, rtol=1e-5, abstol=1e-50, max_it=10000, pc="mg") -> L.KSPResult:
        """b_host/x_host: CPU float64 tensors (pinned for full PCIe speed) of the local slab."""
        res = L.KSPResult()
        pcid = {"none": L.PC_NONE, "jacobi": L.PC_JACOBI, "mg": L.PC_MG}[pc]
        L.check(self.lib.p4b_cg_solve_host(self.h, pcid, b_host.data_ptr(), x_host.data_ptr(), rtol, abstol, max_it,
                                           C.byref(res)))
        return res

    def fish_solve_host(self, f_host, gb_host, u_host, rtol=1e-5, abstol=1e-50, max_it=10000) -> L.KSPResult:
        res = L.KSPResult()
        L.check(self.lib.p4b_fish_solve_host(self.h, f_host.data_ptr(), gb_host.data_ptr(), u_host.data_ptr(), rtol,
                                             abstol, max_it, C.byref(res)))
        return res

    def fish_setup(self, problem, gonboundary=True, b=None, u0=None, uexact=None):
        """b = F(u0), u0, uexact of fish.c's built-in problems on this rank's slab (device tensors)."""
        ptr = lambda t: t.data_ptr() if t is not None else None
        L.check(self.lib.p4b_mg_fish_setup(self.h, L.PROBLEMS[problem], int(gonboundary), ptr(b), ptr(u0), ptr(uexact)))

    def profile(self, on=True):
        """1/True: finest-level kernels; 2: trace mode (every level, every exchange, CUDA graph bypassed)."""
        L.check(self.lib.p4b_profile_enable(self.h, int(on)))

    def profile_reset(self):
        L.check(self.lib.p4b_profile_reset(self.h))

    def profile_stats(self):
        out = {}
        for i, name in enumerate(L.KERNEL_CLASSES):
            s = L.KernelStat()
            L.check(self.lib.p4b_profile_get(self.h, i, C.byref(s)))
            if s.launches:
                out[name] = {"launches": s.launches, "ms": s.ms, "bytes": s.bytes}
        return out

    def profile_trace(self):
        """{level: {class: {launches, ms, bytes}}} collected in trace mode (level 0 = coarsest)."""
        out = {}
        for l in range(self.nlevels):
            for i, name in enumerate(L.KERNEL_CLASSES):
                s = L.KernelStat()
                L.check(self.lib.p4b_profile_get_level(self.h, l, i, C.byref(s)))
                if s.launches:
                    out.setdefault(l, {})[name] = {"launches": s.launches, "ms": s.ms, "bytes": s.bytes}
        return out

    def close(self):
        if self.h:
            self.lib.p4b_mg_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------------
# the fish driver
# ----------------------------------------------------------------------------------------------------

@dataclass
class FishOptions:
    """Options of fish.c:140-185 plus the PETSc options the reference's scripts pass."""
    dim: int = 2
    problem: str = "manuexp"
    initial_type: str = "zeros"
    gonboundary: bool = True
    cx: float = 1.0
    cy: float = 1.0
    cz: float = 1.0
    Lx: float = 1.0
    Ly: float = 1.0
    Lz: float = 1.0
    da_refine: int = 0
    da_grid: tuple = (3, 3, 3)
    ksp_type: str = "cg"
    ksp_rtol: float = 1e-5
    ksp_atol: float = 1e-50
    ksp_max_it: int = 10000
    pc_type: str = "mg"
    pc_mg_levels: int = 0
    pc_mg_cycle_type: str = "v"
    mg_levels_ksp_type: str = "chebyshev"
    mg_levels_ksp_max_it: int = 2
    mg_levels_pc_type: str = "jacobi"
    mg_eigenvalues: tuple | None = None
    mg_esteig: tuple = (0.1, 1.1)
    ksp_converged_reason: bool = False
    ksp_monitor: bool = False
    snes_monitor_short: bool = False
    fuse: bool = True


def parse_options(argv) -> FishOptions:
    """Parse a reference-style option string/list.  Unknown options raise (PETSc would warn at exit)."""
    if isinstance(argv, str):
        argv = shlex.split(argv)
    o = FishOptions()
    have_pc = False
    i = 0

    def val():
        nonlocal i
        i += 1
        if i >= len(argv):
            raise L.P4BError("option %s needs a value" % argv[i - 1])
        return argv[i]

    def boolval():
        nonlocal i
        if i + 1 < len(argv) and not argv[i + 1].startswith("-"):
            i += 1
            return argv[i].lower() in ("1", "true", "yes", "on")
        return True

    while i < len(argv):
        a = argv[i]
        if a == "-fsh_dim": o.dim = int(val())
        elif a == "-fsh_problem": o.problem = val()
        elif a == "-fsh_initial_type": o.initial_type = val()
        elif a == "-fsh_initial_gonboundary": o.gonboundary = boolval()
        elif a in ("-fsh_cx", "-fsh_cy", "-fsh_cz", "-fsh_Lx", "-fsh_Ly", "-fsh_Lz"):
            setattr(o, a[5:], float(val()))
        elif a == "-da_refine": o.da_refine = int(val())
        elif a in ("-da_grid_x", "-da_grid_y", "-da_grid_z"):
            gxyz = list(o.da_grid)
            gxyz["xyz".index(a[-1])] = int(val())
            o.da_grid = tuple(gxyz)
        elif a == "-ksp_type": o.ksp_type = val()
        elif a == "-ksp_rtol": o.ksp_rtol = float(val())
        elif a == "-ksp_atol": o.ksp_atol = float(val())
        elif a == "-ksp_max_it": o.ksp_max_it = int(val())
        elif a == "-pc_type": o.pc_type = val(); have_pc = True
        elif a == "-pc_mg_levels": o.pc_mg_levels = int(val())
        elif a == "-pc_mg_cycle_type": o.pc_mg_cycle_type = val()
        elif a == "-mg_levels_ksp_type": o.mg_levels_ksp_type = val()
        elif a == "-mg_levels_ksp_max_it": o.mg_levels_ksp_max_it = int(val())
        elif a == "-mg_levels_pc_type": o.mg_levels_pc_type = val()
        elif a == "-mg_levels_ksp_chebyshev_eigenvalues":
            o.mg_eigenvalues = tuple(float(t) for t in val().split(","))
        elif a == "-mg_levels_ksp_chebyshev_esteig":
            t = [float(s) for s in val().split(",")]
            o.mg_esteig = (t[1], t[3]) if len(t) == 4 else (t[0], t[1])
        elif a == "-snes_type":
            if val() != "ksponly":
                raise L.P4BError("fish is linear: only -snes_type ksponly is provided (fish.c:231)")
        elif a == "-ksp_converged_reason": o.ksp_converged_reason = True
        elif a == "-ksp_monitor": o.ksp_monitor = True
        elif a in ("-snes_monitor_short", "-snes_monitor"): o.snes_monitor_short = True
        elif a == "-p4b_no_fuse": o.fuse = False
        else:
            raise L.P4BError("unknown option %s" % a)
        i += 1
    if not have_pc:
        raise L.P4BError("the default PETSc PC (ILU) is sequential and not provided on the device: pass -pc_type mg|jacobi|none")
    if o.dim not in (1, 2, 3):
        raise L.P4BError("invalid dim for DMDA creation")                        # fish.c:215
    if o.cx <= 0 or o.cy <= 0 or o.cz <= 0:
        raise L.P4BError("positivity required for coefficients cx,cy,cz")        # fish.c:189
    if o.problem == "manuexp" and (o.cx != 1.0 or o.cy != 1.0 or o.cz != 1.0):
        raise L.P4BError("cx=cy=cz=1 required for problem MANUEXP")              # fish.c:192
    if o.problem not in L.PROBLEMS:
        raise L.P4BError("unknown -fsh_problem %s" % o.problem)
    if o.initial_type not in ("zeros", "random"):
        raise L.P4BError("unknown -fsh_initial_type %s (zeros, random)" % o.initial_type)
    if o.ksp_type != "cg":
        raise L.P4BError("only -ksp_type cg is provided on the device")
    if o.pc_type not in ("mg", "jacobi", "none"):
        raise L.P4BError("-pc_type %s is not provided on the device (mg, jacobi, none)" % o.pc_type)
    if o.pc_type == "mg" and o.mg_levels_pc_type != "jacobi":
        raise L.P4BError("-mg_levels_pc_type %s is sequential; the device smoother is jacobi" % o.mg_levels_pc_type)
    return o


@dataclass
class FishReport:
    options: FishOptions
    grid: L.Grid
    ksp: L.KSPResult
    fnorm0: float
    fnorm1: float
    errinf: float
    err2h: float
    lines: list = field(default_factory=list)
    u: "torch.Tensor | None" = None

    @property
    def text(self):
        return "\n".join(self.lines) + "\n"


def _snes_short(x):
    """PETSc's -snes_monitor_short number format ([PETSc] SNESMonitorDefaultShort): %g above 1e-9, %5.3e down to
    1e-11, '< 1.e-11' below that."""
    if x > 1e-9:
        return "%g" % x
    if x > 1e-11:
        return "%5.3e" % x
    return "< 1.e-11"


def grid_string(g: L.Grid):
    """fish.c:260-272."""
    if g.dim == 1:
        return "%d point 1D" % g.mx
    if g.dim == 2:
        return "%d x %d point 2D" % (g.mx, g.my)
    return "%d x %d x %d point 3D" % (g.mx, g.my, g.mz)


def fish_main(argv, ctx: Context | None = None, keep_solution=False, echo=False) -> FishReport:
    """Run fish.c:main on the device (single rank).  Returns the report; report.text is what fish prints."""
    o = parse_options(argv)
    ctx = ctx or Context()
    m = [1 + (2 ** o.da_refine) * (o.da_grid[d] - 1) if d < o.dim else 1 for d in range(3)]
    g = L.make_grid(o.dim, m, (o.Lx, o.Ly, o.Lz), (o.cx, o.cy, o.cz))
    n = g.n
    f, gb, u, F = ctx.empty(n), ctx.empty(n), ctx.empty(n), ctx.empty(n)
    ctx.fish_sample(g, o.problem, f, gb)                       # f_rhs, g_bdry tables (fish.c:115-123)
    if o.initial_type == "random":                             # poissonfunctions.c:267-271: VecSetRandom, then g on the boundary
        u.copy_(ctx.rander48(n))
        ctx.initial_state(g, gb, int(o.gonboundary) | 2, u)
    else:
        ctx.initial_state(g, gb, o.gonboundary, u)             # InitialState (fish.c:238)
    lines = []

    def out(s):
        lines.append(s)
        if echo:
            print(s)

    # SNESSolve_KSPONLY (fish.c:239): F0 = F(u0); J y = F0; u = u0 - y
    ctx.poisson_function(g, u, f, gb, F)
    fnorm0 = ctx.norm2(F)
    if o.snes_monitor_short:
        out("  0 SNES Function norm %s" % _snes_short(fnorm0))
    # [PETSc] PCSetUp_MG on a DMDA: refine+1 levels unless -pc_mg_levels, the grid before -da_refine being the coarsest
    opts = mg_options(levels=o.pc_mg_levels or (o.da_refine + 1 if o.pc_type == "mg" else 1), cycle=o.pc_mg_cycle_type, smoother=o.mg_levels_ksp_type,
                      smooth_its=o.mg_levels_ksp_max_it, eig=o.mg_eigenvalues, esteig=o.mg_esteig, fuse=o.fuse)
    mg = Multigrid(ctx, g, opts)
    y = ctx.empty(n)
    res = mg.cg_solve(F, y, rtol=o.ksp_rtol, abstol=o.ksp_atol, max_it=o.ksp_max_it, pc=o.pc_type)
    if o.ksp_monitor:
        for i, r in enumerate(res.history):
            out("    %d KSP Residual norm %.12e" % (i, r))
    if o.ksp_converged_reason:
        if res.reason > 0:
            out("    Linear solve converged due to %s iterations %d" % (L.REASONS[res.reason], res.its))
        else:
            out("    Linear solve did not converge due to %s iterations %d" % (L.REASONS.get(res.reason, "?"), res.its))
    ctx.axpy(-1.0, y, u)
    ctx.poisson_function(g, u, f, gb, F)
    fnorm1 = ctx.norm2(F)
    if o.snes_monitor_short:
        out("  1 SNES Function norm %s" % _snes_short(fnorm1))
    # error report (fish.c:248-280): gb holds u_exact at every node
    usol = u.clone() if keep_solution else None
    ctx.axpy(-1.0, gb, u)
    errinf = ctx.norminf(u)
    err2 = ctx.norm2(u)
    normconst = math.sqrt(float(math.prod((g.m[d] - 1) for d in range(g.dim))))
    err2h = err2 / normconst
    out("problem %s on %s grid:" % (o.problem, grid_string(g)))
    out("  error |u-uexact|_inf = %.3e, |u-uexact|_h = %.3e" % (errinf, err2h))
    mg.close()
    return FishReport(o, g, res, fnorm0, fnorm1, errinf, err2h, lines, usol)
