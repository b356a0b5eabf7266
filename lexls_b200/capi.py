"""ctypes binding of the C-ABI (include/lexls_b200.h).  The CUDA library is the only implementation: importing this
module fails loudly when liblexls_b200.so has not been built, and creating a batch fails when no GPU is usable."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("LEXLS_B200_LIB") or os.path.join(_HERE, "liblexls_b200.so")  # the variable: scripts/build_variants.py

# every symbol include/lexls_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "lexls_b200_abi_version", "lexls_b200_last_error", "lexls_b200_default_params", "lexls_b200_device_count",
    "lexls_b200_create_batch", "lexls_b200_destroy", "lexls_b200_reset", "lexls_b200_set_params", "lexls_b200_set_stream",
    "lexls_b200_set_data", "lexls_b200_set_data_dev", "lexls_b200_set_simple_bounds", "lexls_b200_set_simple_bounds_dev",
    "lexls_b200_set_fixed_variables", "lexls_b200_set_active_guess", "lexls_b200_set_active_guess_dev", "lexls_b200_set_x0",
    "lexls_b200_set_x0_dev", "lexls_b200_set_v0", "lexls_b200_solve_batch", "lexls_b200_solve_batch_sync",
    "lexls_b200_synchronize", "lexls_b200_get_x", "lexls_b200_get_v", "lexls_b200_get_constraint_violation",
    "lexls_b200_get_active_set", "lexls_b200_get_active_order", "lexls_b200_get_status", "lexls_b200_get_counters",
    "lexls_b200_get_ranks", "lexls_b200_get_column_permutations", "lexls_b200_device_results", "lexls_b200_kernel_info",
    "lexls_b200_last_solve_ms", "lexls_b200_measure_fp64_peak", "lexls_b200_request_outputs", "lexls_b200_get_lexqr",
    "lexls_b200_get_lambda", "lexls_b200_get_working_set_log", "lexls_b200_validate", "lexls_b200_create_batch_multi",
    "lexls_b200_shard_count", "lexls_b200_get_shard", "lexls_b200_get_xstar", "lexls_b200_set_active_guess_ordered",
    "lexls_b200_check_guard_bands", "lexls_b200_set_regularization_factor",
    "lexls_b200_get_x_least_norm", "lexls_b200_set_async_copies", "lexls_b200_wait_copies",
]
OUT_LEXQR, OUT_LAMBDA, OUT_XSTAR, OUT_LEAST_NORM = 1, 2, 4, 8

OK, EINVAL, ECUDA, EUNSUPPORTED, EBOUNDS, ECAPACITY = range(6)
LSE, LSI = 0, 1
NCOUNTERS = 8


class Params(C.Structure):
    """POD mirror of ParametersLexLSI (reference include/lexls/typedefs.h:127-295)."""

    _fields_ = [
        ("max_number_of_factorizations", C.c_uint32),
        ("tol_linear_dependence", C.c_double),
        ("tol_wrong_sign_lambda", C.c_double),
        ("tol_correct_sign_lambda", C.c_double),
        ("tol_feasibility", C.c_double),
        ("cycling_handling_enabled", C.c_int32),
        ("cycling_max_counter", C.c_uint32),
        ("cycling_relax_step", C.c_double),
        ("regularization_type", C.c_int32),
        ("modify_x_guess_enabled", C.c_int32),
        ("modify_type_active_enabled", C.c_int32),
        ("modify_type_inactive_enabled", C.c_int32),
        ("set_min_init_ctr_violation", C.c_int32),
        ("use_phase1_v0", C.c_int32),
        ("log_working_set_enabled", C.c_int32),
        ("deactivate_first_wrong_sign", C.c_int32),
        ("variable_regularization_factor", C.c_double),
        ("max_number_of_cg_iterations", C.c_uint32),
    ]

    def __init__(self, **kw):
        super().__init__()
        lib().lexls_b200_default_params(C.byref(self))
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


class KernelInfo(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("grid", C.c_uint32), ("block", C.c_uint32), ("smem_bytes", C.c_uint32),
                ("problems_per_sm", C.c_uint32), ("launches_per_solve", C.c_uint32), ("warps_per_problem", C.c_uint32)]


class LexLSError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("lexls_b200 error %d: %s" % (code, message))
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError("lexls_b200: %s is missing — build it with `python -m lexls_b200.build` (nvcc, sm_100a). "
                              "There is no CPU fallback." % SO_PATH)
        L = C.CDLL(SO_PATH)
        L.lexls_b200_last_error.restype = C.c_char_p
        for name in SYMBOLS:
            getattr(L, name)  # raises AttributeError on a missing export
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise LexLSError(rc, lib().lexls_b200_last_error().decode())
