"""ctypes binding of include/somar_b200.h (the C ABI of libsomar_b200.so).

The library is the product; this module only loads it and declares the prototypes.  It fails
loudly when the shared object is missing -- there is no CPU fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsomar_b200.so")

SB_MAX_HISTORY = 64

RELAX_NONE, RELAX_JACOBI, RELAX_JACOBIRB, RELAX_GS, RELAX_GSRB, RELAX_VERTLINE = 0, 2, 3, 4, 5, 6
STATUS_NAMES = {-1: "UNDEFINED", 0: "DIVERGED", 1: "CONVERGED", 2: "SINGULAR", 3: "MAXITERS", 4: "HANG"}
MAP_CARTESIAN, MAP_STRETCHED, MAP_CALLBACK = 0, 1, 2
CELL, FACE_X, FACE_Y, FACE_Z = -1, 0, 1, 2
MODE_MG, MODE_LEPTIC, MODE_LEPTIC_MG = 1, 2, 3

MAP_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int, C.c_void_p)

I3 = C.c_int * 3
D3 = C.c_double * 3


class LevelDesc(C.Structure):
    _fields_ = [
        ("dim", C.c_int),
        ("domain_lo", I3),
        ("domain_hi", I3),
        ("periodic", I3),
        ("dXi", D3),
        ("num_boxes", C.c_int),
        ("box_lo", C.POINTER(C.c_int)),
        ("box_hi", C.POINTER(C.c_int)),
        ("box_rank", C.POINTER(C.c_int)),
        ("map_kind", C.c_int),
        ("map_xmin", D3),
        ("map_xmax", D3),
        ("map_ampl", D3),
        ("map_fn", MAP_FN),
        ("map_user", C.c_void_p),
        ("bc_alpha", (C.c_double * 2) * 3),
        ("bc_beta", (C.c_double * 2) * 3),
        ("alpha", C.c_double),
        ("beta", C.c_double),
        ("relax_method", C.c_int),
        ("num_crse_boxes", C.c_int),
        ("crse_box_lo", C.POINTER(C.c_int)),
        ("crse_box_hi", C.POINTER(C.c_int)),
        ("crse_box_rank", C.POINTER(C.c_int)),
        ("crse_domain_lo", I3),
        ("crse_domain_hi", I3),
    ]


class BottomOptions(C.Structure):
    _fields_ = [
        ("absTol", C.c_double), ("relTol", C.c_double), ("small", C.c_double), ("hang", C.c_double),
        ("convergenceMetric", C.c_double),
        ("maxIters", C.c_int), ("maxRestarts", C.c_int), ("normType", C.c_int), ("verbosity", C.c_int),
        ("numSmoothPrecond", C.c_int),
    ]


class MGOptions(C.Structure):
    _fields_ = [
        ("absTol", C.c_double), ("relTol", C.c_double), ("convergenceMetric", C.c_double), ("hang", C.c_double),
        ("numSmoothDown", C.c_int), ("numSmoothUp", C.c_int), ("numSmoothBottom", C.c_int), ("numSmoothPrecond", C.c_int),
        ("prolongOrder", C.c_int), ("prolongOrderFMG", C.c_int), ("numSmoothUpFMG", C.c_int),
        ("maxDepth", C.c_int), ("numCycles", C.c_int), ("maxIters", C.c_int), ("normType", C.c_int), ("verbosity", C.c_int),
        ("bottom", BottomOptions),
    ]


class SolverStatus(C.Structure):
    _fields_ = [
        ("status", C.c_int), ("num_iters", C.c_int),
        ("init_res_norm", C.c_double), ("final_res_norm", C.c_double),
        ("num_norms", C.c_int), ("res_norms", C.c_double * SB_MAX_HISTORY),
        ("solve_mode", C.c_int), ("max_depth", C.c_int), ("device_ms", C.c_double),
    ]

    @property
    def norms(self):
        return [self.res_norms[i] for i in range(self.num_norms)]


P = C.c_void_p
PP = C.POINTER(C.c_void_p)
DP = C.POINTER(C.c_double)
IP = C.POINTER(C.c_int)
F3 = C.c_void_p * 3

# name -> (restype, argtypes); every symbol include/somar_b200.h declares.
PROTOTYPES = {
    "sb_last_error": (C.c_char_p, []),
    "sb_version": (C.c_int, []),
    "sb_context_create": (C.c_int, [PP, C.c_int, C.c_int, C.c_int]),
    "sb_context_destroy": (C.c_int, [P]),
    "sb_context_sync": (C.c_int, [P]),
    "sb_context_launch_count": (C.c_longlong, [P]),
    "sb_context_timer_start": (C.c_int, [P]),
    "sb_context_timer_stop": (C.c_int, [P, DP]),
    "sb_context_profile": (C.c_int, [P, C.c_int]),
    "sb_context_profile_get": (C.c_int, [P, C.c_char_p, DP, C.POINTER(C.c_longlong)]),
    "sb_comm_get_unique_id": (C.c_int, [C.c_void_p]),
    "sb_comm_init": (C.c_int, [P, C.c_void_p]),
    "sb_plan_tile": (C.c_int, [C.POINTER(LevelDesc), C.c_int, C.c_int, IP, IP, IP, IP, IP]),
    "sb_plan_schedule": (C.c_int, [C.POINTER(LevelDesc), C.c_int, IP, C.c_int, IP]),
    "sb_plan_cf_stencils": (C.c_int, [IP, IP, IP, IP, C.c_int, IP, IP, C.c_int, C.c_int, C.c_int, C.c_int, IP, IP, DP, DP, DP]),
    "sb_op_create": (C.c_int, [P, C.POINTER(LevelDesc), PP]),
    "sb_op_set_metric": (C.c_int, [P, C.c_int, C.c_int, DP, IP, IP]),
    "sb_op_finalize": (C.c_int, [P]),
    "sb_op_destroy": (C.c_int, [P]),
    "sb_op_has_null_space": (C.c_int, [P, IP]),
    "sb_op_halo_mode": (C.c_int, [P, IP]),
    "sb_plan_line_tile_order": (C.c_int, [C.c_int, C.c_int, C.c_int, IP, C.c_int, IP]),
    "sb_op_new_mg_operator": (C.c_int, [P, IP, PP]),
    "sb_op_get_info": (C.c_int, [P, IP, IP, DP, IP]),
    "sb_op_get_coefficient": (C.c_int, [P, C.c_int, DP, C.c_longlong]),
    "sb_field_create": (C.c_int, [P, C.c_int, PP]),
    "sb_field_destroy": (C.c_int, [P]),
    "sb_field_upload": (C.c_int, [P, DP, IP, IP]),
    "sb_field_download": (C.c_int, [P, DP, IP, IP]),
    "sb_op_apply_bcs": (C.c_int, [P, P, C.c_int]),
    "sb_op_apply_op": (C.c_int, [P, P, P, C.c_int]),
    "sb_op_residual": (C.c_int, [P, P, P, P, C.c_int]),
    "sb_op_relax": (C.c_int, [P, P, P, C.c_int]),
    "sb_op_precond": (C.c_int, [P, P, P, C.c_int]),
    "sb_op_remove_kernel": (C.c_int, [P, P]),
    "sb_op_norm": (C.c_int, [P, P, C.c_int, C.c_double, DP]),
    "sb_op_dot": (C.c_int, [P, P, P, DP]),
    "sb_op_incr": (C.c_int, [P, P, P, C.c_double]),
    "sb_op_axby": (C.c_int, [P, P, P, P, C.c_double, C.c_double]),
    "sb_op_scale": (C.c_int, [P, P, C.c_double]),
    "sb_op_set_to_zero": (C.c_int, [P, P]),
    "sb_op_assign_local": (C.c_int, [P, P, P]),
    "sb_op_mg_restrict": (C.c_int, [P, P, P, P]),
    "sb_op_mg_prolong": (C.c_int, [P, P, P, P, C.c_int]),
    "sb_op_level_divergence": (C.c_int, [P, P, F3]),
    "sb_op_level_gradient": (C.c_int, [P, F3, P, C.c_int]),
    "sb_op_flux_incr": (C.c_int, [P, F3, F3, C.c_double]),
    "sb_op_send_to_advecting_velocity": (C.c_int, [P, F3, C.c_int]),
    "sb_op_send_to_cartesian_velocity": (C.c_int, [P, F3, C.c_int]),
    "sb_op_apply_bcs_amr": (C.c_int, [P, P, P, C.c_int, C.c_int]),
    "sb_op_amr_operator": (C.c_int, [P, P, P, P, P, C.c_int, P]),
    "sb_op_amr_operator_nf": (C.c_int, [P, P, P, P, C.c_int]),
    "sb_op_amr_operator_nc": (C.c_int, [P, P, P, P, C.c_int, P]),
    "sb_op_amr_residual": (C.c_int, [P, P, P, P, P, P, C.c_int, P]),
    "sb_op_amr_norm_level": (C.c_int, [P, P, P, C.c_int, DP]),
    "sb_op_get_flux": (C.c_int, [P, F3, P]),
    "sb_op_reflux": (C.c_int, [P, P, P, P, P]),
    "sb_op_reflux_flux": (C.c_int, [P, P, F3, F3, P]),
    "sb_op_comp_divergence": (C.c_int, [P, P, F3, PP, P]),
    "sb_op_comp_gradient": (C.c_int, [P, F3, P, P, C.c_int, C.c_int]),
    "sb_op_average_down": (C.c_int, [P, P, P]),
    "sb_op_get_patch": (C.c_int, [P, IP, IP, IP, IP]),
    "sb_amr_solver_create": (C.c_int, [PP, C.c_int, C.c_int, C.c_int, C.POINTER(MGOptions), PP]),
    "sb_amr_solver_destroy": (C.c_int, [P]),
    "sb_amr_solver_solve": (C.c_int, [P, PP, PP, C.c_int, C.c_int, C.c_double, C.POINTER(SolverStatus)]),
    "sb_mg_default_options": (None, [C.POINTER(MGOptions)]),
    "sb_mg_quick_and_dirty_options": (None, [C.POINTER(MGOptions)]),
    "sb_mgsolver_create": (C.c_int, [P, C.POINTER(MGOptions), IP, C.c_int, PP]),
    "sb_hybrid_solver_create": (C.c_int, [P, C.POINTER(MGOptions), PP]),
    "sb_solver_destroy": (C.c_int, [P]),
    "sb_solver_get_schedule": (C.c_int, [P, IP, C.c_int, IP]),
    "sb_solver_set_options": (C.c_int, [P, C.POINTER(MGOptions)]),
    "sb_solver_solve": (C.c_int, [P, P, P, C.c_int, C.c_int, C.c_double, C.POINTER(SolverStatus)]),
    "sb_solver_vcycle": (C.c_int, [P, P, P]),
    "sb_solver_precond_vcycle": (C.c_int, [P, P, P]),
    "sb_context_stream_wait": (C.c_int, [P, C.c_int, C.c_int]),
    "sb_context_stream_sync": (C.c_int, [P, C.c_int]),
    "sb_field_upload_async": (C.c_int, [P, DP, IP, IP]),
    "sb_field_download_async": (C.c_int, [P, DP, IP, IP]),
    "sb_project_correct": (C.c_int, [P, F3, P, C.c_double, C.c_int, P, DP, DP, C.POINTER(SolverStatus)]),
    "sb_project_predict": (C.c_int, [P, F3, P, C.c_double, C.c_int, DP, IP, C.POINTER(SolverStatus)]),
    "sb_project_host": (C.c_int, [P, DP * 3, DP, DP, C.c_double, DP, DP, C.POINTER(SolverStatus)]),
}

_lib = None


class SomarB200Error(RuntimeError):
    pass


def load():
    """dlopen libsomar_b200.so and attach prototypes.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SomarB200Error(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C somar_b200/csrc).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise SomarB200Error(load().sb_last_error().decode())
