"""ctypes bindings of include/p4b200.h (the C ABI of libp4b200.so).

PyTorch supplies device memory (tensors), streams and torch.distributed; every numerical
operation goes through the C ABI into the hand-written sm_100a kernels.  There is no
fallback: if the library is missing, or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# P4B_LIB: an alternative build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("P4B_LIB") or os.path.join(HERE, "lib", "libp4b200.so")

P4B_MAX_HIST = 256
CYCLE_V, CYCLE_W = 1, 2
SMOOTH_CHEBYSHEV, SMOOTH_RICHARDSON = 0, 1
PC_NONE, PC_JACOBI, PC_MG = 0, 1, 2
PROBLEMS = {"manupoly": 0, "manuexp": 1, "zero": 2}
CONVERGED_RTOL, CONVERGED_ATOL, DIVERGED_ITS, DIVERGED_NAN = 2, 3, -3, -9
REASONS = {2: "CONVERGED_RTOL", 3: "CONVERGED_ATOL", -3: "DIVERGED_ITS", -4: "DIVERGED_DTOL", -8: "DIVERGED_INDEFINITE_PC",
           -9: "DIVERGED_NANORINF", -10: "DIVERGED_INDEFINITE_MAT"}
KERNEL_CLASSES = ["apply_dot", "residual", "cheb_zero", "cheb_first", "cheb_next", "restrict", "prolong_add",
                  "axpy2", "dot2", "aypx", "resid_restrict", "xp_update", "r_update",
                  "halo", "gather", "allreduce", "coarse_solve", "subcycle"]
# the classes after r_update are exchanges / latency items, not roofline kernels (no algorithmic bytes)
N_ROOFLINE_CLASSES = 13


class Grid(C.Structure):
    _fields_ = [("dim", C.c_int), ("mx", C.c_int), ("my", C.c_int), ("mz", C.c_int),
                ("Lx", C.c_double), ("Ly", C.c_double), ("Lz", C.c_double),
                ("cx", C.c_double), ("cy", C.c_double), ("cz", C.c_double)]

    @property
    def n(self):
        return self.mx * self.my * self.mz

    @property
    def m(self):
        return (self.mx, self.my, self.mz)


class MGOpts(C.Structure):
    _fields_ = [("levels", C.c_int), ("cycle", C.c_int), ("smoother", C.c_int), ("smooth_its", C.c_int),
                ("emin", C.c_double), ("emax", C.c_double), ("est_lo", C.c_double), ("est_hi", C.c_double),
                ("fuse", C.c_int), ("use_graph", C.c_int)]


class KSPResult(C.Structure):
    _fields_ = [("its", C.c_int), ("reason", C.c_int), ("rnorm0", C.c_double), ("rnorm", C.c_double),
                ("nhist", C.c_int), ("hist", C.c_double * P4B_MAX_HIST), ("solve_ms", C.c_double)]

    @property
    def history(self):
        return [self.hist[i] for i in range(self.nhist)]


class MinimalOpts(C.Structure):
    """p4b_minimal_opts (include/p4b200.h)."""
    _fields_ = [("problem", C.c_int), ("q", C.c_double), ("catenoid_c", C.c_double), ("tent_H", C.c_double),
                ("exact_init", C.c_int), ("grid_x", C.c_int), ("grid_y", C.c_int), ("refine", C.c_int),
                ("grid_sequence", C.c_int), ("ksp_type", C.c_int), ("ksp_rtol", C.c_double), ("ksp_max_it", C.c_int),
                ("gmres_restart", C.c_int), ("pc_type", C.c_int), ("mg_levels", C.c_int), ("smooth_its", C.c_int),
                ("snes_rtol", C.c_double), ("snes_stol", C.c_double), ("snes_atol", C.c_double), ("snes_max_it", C.c_int),
                ("snes_monitor", C.c_int), ("snes_converged_reason", C.c_int), ("ksp_converged_reason", C.c_int),
                ("mf_operator", C.c_int), ("jacobian", C.c_int)]


class MinimalStage(C.Structure):
    _fields_ = [("mx", C.c_int), ("my", C.c_int), ("its", C.c_int), ("reason", C.c_int), ("nksp", C.c_int),
                ("ksp_its", C.c_int * 64), ("lam", C.c_double * 64), ("fnorm", C.c_double * 65)]


class MinimalResult(C.Structure):
    _fields_ = [("mx", C.c_int), ("my", C.c_int), ("nstages", C.c_int), ("stage", MinimalStage * 16),
                ("errinf", C.c_double), ("error", C.c_int), ("errmsg", C.c_char * 256)]


class PatternOpts(C.Structure):
    """p4b_pattern_opts (include/p4b200.h)."""
    _fields_ = [("L", C.c_double), ("Du", C.c_double), ("Dv", C.c_double), ("phi", C.c_double), ("kappa", C.c_double),
                ("no_rhsjacobian", C.c_int), ("call_back_report", C.c_int), ("grid_x", C.c_int), ("grid_y", C.c_int),
                ("refine", C.c_int), ("ts_type", C.c_int), ("ts_dt", C.c_double), ("ts_max_time", C.c_double),
                ("ts_max_steps", C.c_int), ("ts_rtol", C.c_double), ("ts_atol", C.c_double), ("ts_monitor", C.c_int),
                ("pc_type", C.c_int), ("smooth_its", C.c_int), ("mg_rscale", C.c_double), ("snes_rtol", C.c_double),
                ("snes_stol", C.c_double), ("snes_atol", C.c_double), ("snes_max_it", C.c_int), ("ksp_rtol", C.c_double),
                ("ksp_max_it", C.c_int), ("gmres_restart", C.c_int), ("snes_converged_reason", C.c_int),
                ("ksp_converged_reason", C.c_int)]


class PatternResult(C.Structure):
    _fields_ = [("m", C.c_int), ("nsteps", C.c_int), ("rejected", C.c_int), ("ksp_its_total", C.c_longlong),
                ("newton_its_total", C.c_longlong), ("t_final", C.c_double), ("dt_last", C.c_double),
                ("step_t", C.c_double * 512), ("step_dt", C.c_double * 512), ("step_newton", C.c_int * 512),
                ("error", C.c_int)]


LINE_FN = C.CFUNCTYPE(None, C.c_char_p, C.c_void_p)
RESIDUAL2D_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double))
IFUNCTION2D_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double),
                             C.POINTER(C.c_double))
RHSFUNCTION2D_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double))
TS_STEP_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_double), C.c_size_t)
MONITOR2D_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_double))


class BratuOpts(C.Structure):
    """p4b_bratu_opts (include/p4b200.h)."""
    _fields_ = [("lam", C.c_double), ("exact", C.c_int), ("grid_x", C.c_int), ("grid_y", C.c_int), ("refine", C.c_int),
                ("levels", C.c_int), ("snes_rtol", C.c_double), ("snes_max_it", C.c_int), ("smooth_sweeps", C.c_int),
                ("smooth_its", C.c_int), ("coarse_sweeps", C.c_int), ("coarse_its", C.c_int), ("full_cycle", C.c_int),
                ("monitor", C.c_int), ("converged_reason", C.c_int)]


class BratuResult(C.Structure):
    _fields_ = [("mx", C.c_int), ("my", C.c_int), ("its", C.c_int), ("reason", C.c_int), ("nnorm", C.c_int),
                ("fnorm", C.c_double * 64), ("errinf", C.c_double), ("residual_calls", C.c_longlong),
                ("ngs_calls", C.c_longlong), ("solve_ms", C.c_double)]


class KernelStat(C.Structure):
    _fields_ = [("launches", C.c_longlong), ("ms", C.c_double), ("bytes", C.c_double)]


class P4BError(RuntimeError):
    pass


_P = C.c_void_p
_D = C.c_void_p      # device pointer
_SIGS = {
    "p4b_version": (C.c_int, []),
    "p4b_last_error": (C.c_char_p, []),
    "p4b_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "p4b_launch_count": (C.c_longlong, []),
    "p4b_tune": (C.c_int, [C.c_char_p, C.c_long]),
    "p4b_kernel_name": (C.c_char_p, [C.c_int]),
    "p4b_ctx_create": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "p4b_ctx_destroy": (C.c_int, [_P]),
    "p4b_ctx_sync": (C.c_int, [_P]),
    "p4b_comm_unique_id": (C.c_int, [_P]),
    "p4b_comm_init": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "p4b_comm_stats": (C.c_int, [_P, C.POINTER(C.c_ulonglong * 5), C.c_int]),
    "p4b_slab_range": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "p4b_malloc": (C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    "p4b_free": (C.c_int, [_P, _P]),
    "p4b_memcpy_h2d": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "p4b_memcpy_d2h": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "p4b_stencil_apply": (C.c_int, [_P, C.POINTER(Grid), _D, _D]),
    "p4b_stencil_residual": (C.c_int, [_P, C.POINTER(Grid), _D, _D, _D]),
    "p4b_cheb_jacobi": (C.c_int, [_P, C.POINTER(Grid), C.c_double, C.c_double, C.c_int, C.c_int, _D, _D, _D]),
    "p4b_restrict": (C.c_int, [_P, C.POINTER(Grid), _D, _D]),
    "p4b_prolong_add": (C.c_int, [_P, C.POINTER(Grid), _D, _D]),
    "p4b_residual_restrict": (C.c_int, [_P, C.POINTER(Grid), _D, _D, _D]),
    "p4b_lambda_max_jacobi": (C.c_int, [C.POINTER(Grid), C.POINTER(C.c_double)]),
    "p4b_vec_dot": (C.c_int, [_P, C.c_size_t, _D, _D, C.POINTER(C.c_double)]),
    "p4b_vec_wrms2": (C.c_int, [_P, C.c_size_t, _D, _D, C.c_double, C.c_double, C.POINTER(C.c_double)]),
    "p4b_vec_norm2": (C.c_int, [_P, C.c_size_t, _D, C.POINTER(C.c_double)]),
    "p4b_vec_norminf": (C.c_int, [_P, C.c_size_t, _D, C.POINTER(C.c_double)]),
    "p4b_vec_axpy": (C.c_int, [_P, C.c_size_t, C.c_double, _D, _D]),
    "p4b_vec_aypx": (C.c_int, [_P, C.c_size_t, C.c_double, _D, _D]),
    "p4b_vec_set": (C.c_int, [_P, C.c_size_t, C.c_double, _D]),
    "p4b_fish_sample": (C.c_int, [_P, C.POINTER(Grid), C.c_int, _D, _D]),
    "p4b_initial_state": (C.c_int, [_P, C.POINTER(Grid), _D, C.c_int, _D]),
    "p4b_poisson_function": (C.c_int, [_P, C.POINTER(Grid), _D, _D, _D, _D]),
    "p4b_mg_default_opts": (C.c_int, [C.POINTER(MGOpts)]),
    "p4b_mg_create": (C.c_int, [_P, C.POINTER(Grid), C.POINTER(MGOpts), C.POINTER(_P)]),
    "p4b_mg_create_stencil": (C.c_int, [_P, C.POINTER(Grid), C.POINTER(MGOpts), C.POINTER(C.c_double), C.c_int,
                                        C.POINTER(_P)]),
    "p4b_mg_destroy": (C.c_int, [_P]),
    "p4b_plan_levels": (C.c_int, [C.POINTER(Grid), C.POINTER(MGOpts), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "p4b_mg_nlevels": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "p4b_mg_level_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "p4b_mg_local_range": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "p4b_mg_matmult": (C.c_int, [_P, _D, _D]),
    "p4b_mg_apply": (C.c_int, [_P, _D, _D]),
    "p4b_cg_solve": (C.c_int, [_P, C.c_int, _D, _D, C.c_double, C.c_double, C.c_int, C.POINTER(KSPResult)]),
    "p4b_cg_solve_host": (C.c_int, [_P, C.c_int, _P, _P, C.c_double, C.c_double, C.c_int, C.POINTER(KSPResult)]),
    "p4b_fish_solve_host": (C.c_int, [_P, _P, _P, _P, C.c_double, C.c_double, C.c_int, C.POINTER(KSPResult)]),
    "p4b_mg_fish_setup": (C.c_int, [_P, C.c_int, C.c_int, _D, _D, _D]),
    "p4b_minimal_sample": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _D]),
    "p4b_minimal_function": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, _D, _D, _D]),
    "p4b_pattern_initial_state": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, _D]),
    "p4b_pattern_initial_state_noisy": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, _D, C.c_double, _D]),
    "p4b_vi_inactive_mask": (C.c_int, [_P, C.c_size_t, _D, _D, _D, _D]),
    "p4b_vec_pointwise_mult": (C.c_int, [_P, C.c_size_t, _D, _D, _D]),
    "p4b_vec_pointwise_max": (C.c_int, [_P, C.c_size_t, _D, _D, _D]),
    "p4b_ctx_create_own_stream": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "p4b_bratu_default_opts": (C.c_int, [C.POINTER(BratuOpts)]),
    "p4b_bratu_function": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, C.c_int, _D, _D, _D]),
    "p4b_bratu_ngs": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _D, _D]),
    "p4b_bratu_exact": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _D]),
    "p4b_bratu_solve": (C.c_int, [_P, C.POINTER(BratuOpts), LINE_FN, C.c_void_p, _D, C.c_size_t, C.POINTER(BratuResult)]),
    "p4b_set_ts_step_monitor": (C.c_int, [TS_STEP_FN, C.c_void_p]),
    "p4b_pattern_slab_plan": (C.c_int, [C.c_int] * 5 + [C.POINTER(C.c_int)] * 4),
    "p4b_rander48_seed": (C.c_ulonglong, [C.c_ulong]),
    "p4b_rander48_fill": (C.c_int, [C.POINTER(C.c_ulonglong), C.c_size_t, C.c_void_p]),
    "p4b_pattern_rhsfunction": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, C.c_double, _D, _D]),
    "p4b_pattern_ifunction": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _D, _D, _D]),
    "p4b_pattern_ijacobian_mult": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                             _D, _D]),
    "p4b_minimal_jacobian_fd": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, _D, _D, _D, _D]),
    "p4b_poisson_stencil9": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _D]),
    "p4b_stencil9_apply": (C.c_int, [_P, C.c_int, C.c_int, _D, _D, _D]),
    "p4b_stencil9_lin": (C.c_int, [_P, C.c_int, C.c_int, _D, _D, _D, _D, C.c_double, C.c_double, C.c_double, C.c_int,
                                   _D]),
    "p4b_dense_matvec": (C.c_int, [_P, C.c_int, _D, _D, _D]),
    "p4b_stencil9_gershgorin": (C.c_int, [_P, C.c_int, C.c_int, _D, _D, C.POINTER(C.c_double)]),
    "p4b_inject2d": (C.c_int, [_P, C.c_int, C.c_int, _D, _D]),
    "p4b_vec_axpby": (C.c_int, [_P, C.c_size_t, C.c_double, _D, C.c_double, _D, _D]),
    "p4b_vec_copy": (C.c_int, [_P, C.c_size_t, _D, _D]),
    "p4b_pattern_jac_apply": (C.c_int, [_P, C.c_int, C.c_int] + [C.c_double] * 6 + [_D, _D, _D]),
    "p4b_pattern_jac_lin": (C.c_int, [_P, C.c_int, C.c_int] + [C.c_double] * 6 + [_D, _D, _D, _D, C.c_double, C.c_double,
                                                                                  C.c_double, C.c_int, _D]),
    "p4b_pattern_jac_gershgorin": (C.c_int, [_P, C.c_int, C.c_int] + [C.c_double] * 6 + [_D, _D, C.POINTER(C.c_double)]),
    "p4b_pattern_restrict": (C.c_int, [_P, C.c_int, C.c_int, _D, _D]),
    "p4b_pattern_prolong_add": (C.c_int, [_P, C.c_int, C.c_int, _D, _D]),
    "p4b_pattern_inject": (C.c_int, [_P, C.c_int, C.c_int, _D, _D]),
    "p4b_minimal_default_opts": (C.c_int, [C.POINTER(MinimalOpts)]),
    "p4b_minimal_solve": (C.c_int, [_P, C.POINTER(MinimalOpts), LINE_FN, _P, _D, C.c_size_t, C.POINTER(MinimalResult)]),
    "p4b_snes2d_solve": (C.c_int, [_P, C.POINTER(MinimalOpts), RESIDUAL2D_FN, _P, _P, LINE_FN, _P, _P, C.c_size_t,
                                  C.POINTER(MinimalResult)]),
    "p4b_snes2d_solve_monitored": (C.c_int, [_P, C.POINTER(MinimalOpts), RESIDUAL2D_FN, MONITOR2D_FN, _P, _P, LINE_FN, _P, _P,
                                            C.c_size_t, C.POINTER(MinimalResult)]),
    "p4b_snes2d_last_route": (C.c_int, []),
    "p4b_vec_mdot": (C.c_int, [_P, C.c_size_t, C.c_int, _P, _D, _P]),
    "p4b_pattern_default_opts": (C.c_int, [C.POINTER(PatternOpts)]),
    "p4b_pattern_solve": (C.c_int, [_P, C.POINTER(PatternOpts), LINE_FN, _P, _D, C.c_size_t, C.POINTER(PatternResult)]),
    "p4b_pattern_solve_from": (C.c_int, [_P, C.POINTER(PatternOpts), _D, LINE_FN, _P, _D, C.c_size_t,
                                        C.POINTER(PatternResult)]),
    "p4b_ts2d_solve": (C.c_int, [_P, C.POINTER(PatternOpts), IFUNCTION2D_FN, RHSFUNCTION2D_FN, _P, _P, C.c_size_t, LINE_FN,
                                _P, C.POINTER(PatternResult)]),
    "p4b_ts_solve_callbacks": (C.c_int, [_P, C.POINTER(PatternOpts), IFUNCTION2D_FN, RHSFUNCTION2D_FN, _P, _P, C.c_size_t,
                                        LINE_FN, _P, C.POINTER(PatternResult)]),
    "p4b_ts_time_step": (C.c_double, []),
    "p4b_heat_rhs": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, _D, _D]),
    "p4b_heat_jac_apply": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, C.c_double, _D, _D]),
    "p4b_heat_solve": (C.c_int, [_P, C.POINTER(PatternOpts), C.c_int, C.c_int, C.c_double, _P, LINE_FN, _P,
                                C.POINTER(PatternResult)]),
    "p4b_sell_create": (C.c_int, [_P, C.c_int, _P, _P, _P, C.POINTER(_P)]),
    "p4b_sell_spmv": (C.c_int, [_P, _D, _D]),
    "p4b_sell_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "p4b_sell_destroy": (C.c_int, [_P]),
    "p4b_profile_enable": (C.c_int, [_P, C.c_int]),
    "p4b_profile_reset": (C.c_int, [_P]),
    "p4b_profile_get": (C.c_int, [_P, C.c_int, C.POINTER(KernelStat)]),
    "p4b_profile_get_level": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(KernelStat)]),
}

_lib = None


def exported_symbols():
    """Every symbol include/p4b200.h declares (the tests check the .so exports each one)."""
    return sorted(_SIGS)


def load(path: str | None = None):
    """dlopen libp4b200.so and attach the signatures.  Raises if it is not built: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise P4BError("libp4b200.so is not built (%s); run `python -m p4pdes_b200.build`. "
                       "There is no CPU fallback." % path)
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise P4BError("p4b200 error %d: %s" % (rc, load().p4b_last_error().decode()))


def plan_levels(grid: Grid, opts: "MGOpts | None", nranks: int):
    """Host-only view of the hierarchy and slab ownership (p4b_plan_levels).  Returns a list of dicts,
    coarsest level first: {"m": (mx,my,mz), "zs": [...per rank], "zm": [...], "replicated": bool}."""
    lib = load()
    nl = C.c_int()
    m3 = (C.c_int * (3 * 16))()
    zs = (C.c_int * (16 * nranks))()
    zm = (C.c_int * (16 * nranks))()
    rep = (C.c_int * 16)()
    check(lib.p4b_plan_levels(C.byref(grid), C.byref(opts) if opts is not None else None, nranks, C.byref(nl), m3,
                              zs, zm, rep))
    return [{"m": (m3[3 * l], m3[3 * l + 1], m3[3 * l + 2]),
             "zs": [zs[l * nranks + r] for r in range(nranks)],
             "zm": [zm[l * nranks + r] for r in range(nranks)],
             "replicated": bool(rep[l])} for l in range(nl.value)]


def tune(key: str, value: int):
    check(load().p4b_tune(key.encode(), int(value)))


def make_grid(dim, m, L=(1.0, 1.0, 1.0), c=(1.0, 1.0, 1.0)) -> Grid:
    m = tuple(m) + (1,) * (3 - len(m))
    return Grid(dim, m[0], m[1] if dim >= 2 else 1, m[2] if dim >= 3 else 1, L[0], L[1], L[2], c[0], c[1], c[2])


def refined_grid(dim, refine, base=3, L=(1.0, 1.0, 1.0), c=(1.0, 1.0, 1.0)) -> Grid:
    """-da_refine n on the 3^d DMDA of fish.c:199-212: m <- 1 + 2^n (m - 1)."""
    m = 1 + (2 ** refine) * (base - 1)
    return make_grid(dim, (m,) * dim, L, c)
