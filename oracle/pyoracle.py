"""ctypes front-end of the CPU oracle (oracle/oracle_capi.cpp).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblexls_oracle.so")
_lib = None


class OracleParams(C.Structure):
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
        ("modify_x_guess_enabled", C.c_int32),
        ("modify_type_active_enabled", C.c_int32),
        ("modify_type_inactive_enabled", C.c_int32),
        ("set_min_init_ctr_violation", C.c_int32),
        ("use_phase1_v0", C.c_int32),
        ("log_working_set_enabled", C.c_int32),
        ("deactivate_first_wrong_sign", C.c_int32),
    ]

    def __init__(self, **kw):
        super().__init__()
        lib().oracle_default_params(C.byref(self))
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


def _cpu_id():
    """Model + flags of the host CPU: the library is built with -march=native, so a copy built on another machine
    (the build container vs the GPU box) must be rebuilt, not loaded."""
    try:
        model = flags = ""
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name") and not model:
                    model = ln.split(":", 1)[1].strip()
                elif ln.startswith("flags") and not flags:
                    flags = ln.split(":", 1)[1].strip()
                if model and flags:
                    break
        import hashlib
        return model + " " + hashlib.sha1(flags.encode()).hexdigest()
    except Exception:
        return "unknown"


def build(force=False):
    """Compile oracle/_build/liblexls_oracle.so (g++ -O3 -march=native; no dependencies)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "lexls_oracle.hpp", "Makefile")]
    stamp = os.path.join(_HERE, "_build", "cpu.txt")
    cpu = _cpu_id()
    same_cpu = os.path.exists(stamp) and open(stamp).read() == cpu
    if (not force) and same_cpu and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs):
        return _SO
    subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    with open(stamp, "w") as f:
        f.write(cpu)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_last_error.restype = C.c_char_p
    return _lib


def set_regularization(reg_type=0, factors=None, variable_factor=0.0, cg_iterations=10):
    """Regularization for the following *_solve_batch calls (lexlse.h:274-411): type 1 TIKHONOV, 2 TIKHONOV_CG, 3 R, 4 R_NO_Z,
    5 RT_NO_Z, 6 RT_NO_Z_CG, 7 TIKHONOV_1, 8 TIKHONOV_2, 9 TEST; factors[k] per objective (setRegularizationFactor);
    cg_iterations = max_number_of_CG_iterations (typedefs.h:91); reg_type=0 switches it off again."""
    lib().oracle_set_cg_iterations(C.c_uint32(int(cg_iterations)))
    f = np.ascontiguousarray(factors if factors is not None else [], dtype=np.float64)
    lib().oracle_set_regularization(C.c_int(int(reg_type)), C.c_double(float(variable_factor)),
                                    f.ctypes.data_as(C.POINTER(C.c_double)) if f.size else None, C.c_uint32(f.size))


def set_least_norm(on):
    """LexLSE batches: solveLeastNorm (lexlse.h:1131-1205) instead of solve() after the factorization."""
    lib().oracle_set_least_norm(C.c_int(1 if on else 0))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ptr_array(arrs, ctype):
    n = len(arrs)
    out = (C.POINTER(ctype) * n)()
    for i, a in enumerate(arrs):
        out[i] = a.ctypes.data_as(C.POINTER(ctype)) if a is not None else None
    return out


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def lse_solve_batch(nVar, ObjDim, data, fixed_index=None, fixed_value=None, fixed_type=None,
                    tol_linear_dependence=1e-12, threads=1, want_factors=False):
    """data[k]: array [B, nVar+1, m_k] (i.e. each problem column-major m_k x (nVar+1))."""
    L = lib()
    nObj = len(ObjDim)
    dims = np.asarray(ObjDim, dtype=np.uint32)
    data = [_f64(d) for d in data]
    B = data[0].shape[0] if nObj else 0
    nCtr = int(dims.sum())
    nFixed = 0
    fi = fv = ft = None
    if fixed_index is not None:
        fi = np.ascontiguousarray(fixed_index, dtype=np.uint32).reshape(B, -1)
        nFixed = fi.shape[1]
        fv = _f64(fixed_value).reshape(B, nFixed)
        if fixed_type is not None:
            ft = np.ascontiguousarray(fixed_type, dtype=np.uint8).reshape(B, nFixed)
    x = np.zeros((B, nVar))
    v = np.zeros((B, nCtr))
    ranks = np.zeros((B, nObj), dtype=np.uint32)
    total_rank = np.zeros(B, dtype=np.uint32)
    margins = np.zeros((B, 2))
    lexqr = np.zeros((B, nVar + 1, nCtr)) if want_factors else None
    hh = np.zeros((B, nCtr)) if want_factors else None
    colperm = np.zeros((B, nVar), dtype=np.uint32) if want_factors else None
    u32 = C.POINTER(C.c_uint32)
    rc = L.oracle_lse_solve_batch(
        C.c_uint32(nVar), C.c_uint32(nObj), dims.ctypes.data_as(u32), C.c_uint64(B), _ptr_array(data, C.c_double),
        C.c_uint32(nFixed), fi.ctypes.data_as(u32) if fi is not None else None, _dp(fv),
        ft.ctypes.data_as(C.POINTER(C.c_uint8)) if ft is not None else None, C.c_double(tol_linear_dependence),
        _dp(x), _dp(v), ranks.ctypes.data_as(u32), _dp(lexqr), _dp(hh),
        colperm.ctypes.data_as(u32) if colperm is not None else None, total_rank.ctypes.data_as(u32), _dp(margins),
        C.c_int(threads))
    if rc != 0:
        raise RuntimeError(L.oracle_last_error().decode())
    out = dict(x=x, v=v, ranks=ranks, total_rank=total_rank, margins=margins)
    if want_factors:
        out.update(lexqr=lexqr, hh=hh, colperm=colperm)
    return out


def _prep_lsi(nVar, ObjDim, ObjType, data, var_index, guess, v0):
    nObj = len(ObjDim)
    dims = np.asarray(ObjDim, dtype=np.uint32)
    types = np.asarray(ObjType if ObjType is not None else [0] * nObj, dtype=np.int32)
    data = [_f64(d) for d in data]
    vi = np.ascontiguousarray(var_index, dtype=np.uint32) if var_index is not None else None
    g = None
    if guess is not None:
        g = [np.ascontiguousarray(a, dtype=np.uint8) if a is not None else None for a in guess]
    vv = None
    if v0 is not None:
        vv = [_f64(a) if a is not None else None for a in v0]
    return nObj, dims, types, data, vi, g, vv


def lsi_solve_batch(nVar, ObjDim, ObjType, data, var_index=None, guess=None, x0=None, v0=None, params=None, threads=1,
                    guess_ordered=None, want_xstar=False):
    """data[k]: [B, nVar+2, m_k] (general) or [B, 2, m_0] (simple bounds); var_index [B, m_0] 0-based.
    guess_ordered = (obj, ctr, type) arrays [B, count]: api_activate calls in CALL order (type 0 = padding), used instead of
    guess.  want_xstar: also return xstar [B, nVar], lexdata and lexqr_star [B, nVar+1, M] (get_xStar / get_data / get_lexqr
    called after solve(), as tests/test_numerical_error.cpp does)."""
    L = lib()
    if guess_ordered is not None or want_xstar:
        return _lsi_solve_batch_ex(L, nVar, ObjDim, ObjType, data, var_index, guess, x0, v0, params, threads, guess_ordered,
                                   want_xstar)
    nObj, dims, types, data, vi, g, vv = _prep_lsi(nVar, ObjDim, ObjType, data, var_index, guess, v0)
    B = data[0].shape[0]
    mtot = int(dims.sum())
    x0a = _f64(x0) if x0 is not None else None
    x = np.zeros((B, nVar))
    v = [np.zeros((B, int(m))) for m in dims]
    viol = [np.zeros((B, int(m))) for m in dims]
    act = [np.zeros((B, int(m)), dtype=np.uint8) for m in dims]
    status = np.zeros(B, dtype=np.int32)
    counters = np.zeros((B, 8), dtype=np.uint32)
    margins = np.zeros((B, 4))
    order = np.zeros((B, mtot), dtype=np.uint32)
    u32 = C.POINTER(C.c_uint32)
    p = params if params is not None else OracleParams()
    rc = L.oracle_lsi_solve_batch(
        C.c_uint32(nVar), C.c_uint32(nObj), dims.ctypes.data_as(u32), types.ctypes.data_as(C.POINTER(C.c_int32)),
        C.c_uint64(B), _ptr_array(data, C.c_double), vi.ctypes.data_as(u32) if vi is not None else None,
        _ptr_array(g, C.c_uint8) if g is not None else None, _dp(x0a),
        _ptr_array(vv, C.c_double) if vv is not None else None, C.byref(p), _dp(x), _ptr_array(v, C.c_double),
        _ptr_array(viol, C.c_double), _ptr_array(act, C.c_uint8), status.ctypes.data_as(C.POINTER(C.c_int32)),
        counters.ctypes.data_as(u32), _dp(margins), order.ctypes.data_as(u32), C.c_int(threads))
    if rc != 0:
        raise RuntimeError(L.oracle_last_error().decode())
    return dict(x=x, v=v, viol=viol, active=act, status=status, counters=counters, margins=margins, active_order=order)


def _lsi_solve_batch_ex(L, nVar, ObjDim, ObjType, data, var_index, guess, x0, v0, params, threads, guess_ordered, want_xstar):
    nObj, dims, types, data, vi, g, vv = _prep_lsi(nVar, ObjDim, ObjType, data, var_index, guess, v0)
    B = data[0].shape[0]
    mtot = int(dims.sum())
    M = int(sum(int(m) for m, t in zip(dims, types) if t != 1))
    x0a = _f64(x0) if x0 is not None else None
    x = np.zeros((B, nVar))
    v = [np.zeros((B, int(m))) for m in dims]
    viol = [np.zeros((B, int(m))) for m in dims]
    act = [np.zeros((B, int(m)), dtype=np.uint8) for m in dims]
    status = np.zeros(B, dtype=np.int32)
    counters = np.zeros((B, 8), dtype=np.uint32)
    margins = np.zeros((B, 4))
    order = np.zeros((B, mtot), dtype=np.uint32)
    u32, u8 = C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
    go = gc = gt = None
    gcount = 0
    if guess_ordered is not None:
        go = np.ascontiguousarray(guess_ordered[0], dtype=np.uint32).reshape(B, -1)
        gc = np.ascontiguousarray(guess_ordered[1], dtype=np.uint32).reshape(B, -1)
        gt = np.ascontiguousarray(guess_ordered[2], dtype=np.uint8).reshape(B, -1)
        gcount = go.shape[1]
    xstar = np.zeros((B, nVar)) if want_xstar else None
    lexdata = np.zeros((B, nVar + 1, M)) if want_xstar else None
    lexqr_star = np.zeros((B, nVar + 1, M)) if want_xstar else None
    p = params if params is not None else OracleParams()
    rc = L.oracle_lsi_solve_batch_ex(
        C.c_uint32(nVar), C.c_uint32(nObj), dims.ctypes.data_as(u32), types.ctypes.data_as(C.POINTER(C.c_int32)),
        C.c_uint64(B), _ptr_array(data, C.c_double), vi.ctypes.data_as(u32) if vi is not None else None,
        _ptr_array(g, C.c_uint8) if g is not None else None, _dp(x0a),
        _ptr_array(vv, C.c_double) if vv is not None else None, C.byref(p), _dp(x), _ptr_array(v, C.c_double),
        _ptr_array(viol, C.c_double), _ptr_array(act, C.c_uint8), status.ctypes.data_as(C.POINTER(C.c_int32)),
        counters.ctypes.data_as(u32), _dp(margins), order.ctypes.data_as(u32), C.c_int(threads), C.c_uint32(gcount),
        go.ctypes.data_as(u32) if go is not None else None, gc.ctypes.data_as(u32) if gc is not None else None,
        gt.ctypes.data_as(u8) if gt is not None else None, _dp(xstar), _dp(lexdata), _dp(lexqr_star))
    if rc != 0:
        raise RuntimeError(L.oracle_last_error().decode())
    out = dict(x=x, v=v, viol=viol, active=act, status=status, counters=counters, margins=margins, active_order=order)
    if want_xstar:
        out.update(xstar=xstar, lexdata=lexdata, lexqr_star=lexqr_star)
    return out


def lsi_solve_detail(nVar, ObjDim, ObjType, data, var_index=None, guess=None, x0=None, params=None, log_capacity=1024):
    """One problem (arrays without the batch axis) with multipliers and the working-set log."""
    L = lib()
    nObj, dims, types, data, vi, g, _ = _prep_lsi(nVar, ObjDim, ObjType, data, var_index, guess, None)
    x0a = _f64(x0) if x0 is not None else None
    x = np.zeros(nVar)
    status = C.c_int32(0)
    counters = np.zeros(8, dtype=np.uint32)
    lam = [np.zeros((nObj, int(m))) for m in dims]  # column-major m_k x nObj
    log = np.zeros((log_capacity, 6))
    u32 = C.POINTER(C.c_uint32)
    p = params if params is not None else OracleParams()
    n = L.oracle_lsi_solve_detail(
        C.c_uint32(nVar), C.c_uint32(nObj), dims.ctypes.data_as(u32), types.ctypes.data_as(C.POINTER(C.c_int32)),
        _ptr_array(data, C.c_double), vi.ctypes.data_as(u32) if vi is not None else None,
        _ptr_array(g, C.c_uint8) if g is not None else None, _dp(x0a), C.byref(p), _dp(x), C.byref(status),
        counters.ctypes.data_as(u32), _ptr_array(lam, C.c_double), _dp(log), C.c_int(log_capacity))
    if n < 0:
        raise RuntimeError(L.oracle_last_error().decode())
    return dict(x=x, status=int(status.value), counters=counters, lambdas=[l.T.copy() for l in lam], log=log[:min(n, log_capacity)])
