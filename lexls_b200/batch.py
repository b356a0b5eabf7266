"""Batched LexLSE / LexLSI front-end over the C-ABI.

Mirrors the reference's call sequence (interfaces/matlab-octave/lexlse.cpp:146-207, lexlsi.cpp:527-625):
construct with the shape, setParameters, setData per level, optional warm start, solve, read results — for a whole
batch of same-shape problems at once.  Arrays are numpy (host); `*_dev` methods take raw device pointers (ints), e.g.
``tensor.data_ptr()`` of a CUDA tensor, so that torch is only plumbing for device memory.

Level data layout: ``data[p]`` is the column-major ``m_k x cols`` matrix of problem ``p`` — as a numpy array of shape
``[B, cols, m_k]`` (C-contiguous), i.e. ``data[p].T`` is the usual ``[m_k, cols]`` matrix ``[A | b]`` or ``[A | lb | ub]``.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import Params, check, lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class _Batch:
    kind = None

    def __init__(self, nVar, ObjDim, ObjType=None, batch=1, device=0, devices=None, strict=True):
        """devices: a list of CUDA ordinals shards the batch over several GPUs of the box (contiguous ranges, one stream and
        host thread per device inside the library: lexls_b200_create_batch_multi); the arrays of every call stay [B, ...].
        strict: setData raises on lb > ub / duplicate VarIndex where the reference throws (one synchronisation per call);
        strict=False defers those checks to solve() / validate() — one synchronisation per batch."""
        self.nVar = int(nVar)
        self.ObjDim = [int(m) for m in ObjDim]
        self.nObj = len(self.ObjDim)
        self.ObjType = [int(t) for t in ObjType] if ObjType is not None else None
        self.batch = int(batch)
        self.device = int(device)
        self.devices = [int(d) for d in devices] if devices is not None else None
        self.strict = bool(strict)
        dims = (C.c_uint32 * self.nObj)(*self.ObjDim)
        types = (C.c_int32 * self.nObj)(*self.ObjType) if self.ObjType is not None else None
        h = C.c_void_p()
        if self.devices is not None:
            devs = (C.c_int * len(self.devices))(*self.devices)
            check(lib().lexls_b200_create_batch_multi(self.kind, C.c_uint32(self.nVar), C.c_uint32(self.nObj), dims, types,
                                                      C.c_uint64(self.batch), devs, C.c_int(len(self.devices)), C.byref(h)))
        else:
            check(lib().lexls_b200_create_batch(self.kind, C.c_uint32(self.nVar), C.c_uint32(self.nObj), dims, types,
                                                C.c_uint64(self.batch), C.c_int(self.device), C.byref(h)))
        self._h = h

    def set_async_copies(self, enabled=True):
        """Host set* calls only enqueue their copies; the caller keeps the input arrays alive and unmodified until the next
        synchronising call (solve(), synchronize(), get_*)."""
        check(lib().lexls_b200_set_async_copies(self._h, C.c_int(1 if enabled else 0)))
        self._async = bool(enabled)
        self._keep = []

    def _hold(self, *arrays):
        """async copies: the converted input arrays must outlive the enqueued copy (released at the next synchronisation)."""
        if getattr(self, "_async", False):
            self._keep.extend(arrays)

    def _release(self):
        if getattr(self, "_async", False):
            self._keep = []

    def validate(self):
        """Waits for the set* calls issued so far and raises what their device-side checks found (lexlsi.h:430, :474, :485)."""
        check(lib().lexls_b200_validate(self._h))
        self._release()

    def wait_copies(self):
        """async copies: returns when every enqueued host->device copy has left the caller's arrays (the device-side repacks
        are deferred to the next solve / read); the point at which a pipelined caller hands the host link to the next handle."""
        check(lib().lexls_b200_wait_copies(self._h))
        self._release()

    def check_guard_bands(self):
        """LEXLS_B200_CANARY=1: number of device buffers whose trailing guard band was written into (0 = clean)."""
        v = C.c_uint32()
        check(lib().lexls_b200_check_guard_bands(self._h, C.byref(v)))
        return v.value

    def shards(self):
        """[(device, first, count)] of a multi-device batch (one entry for a single-device one)."""
        n = C.c_int()
        check(lib().lexls_b200_shard_count(self._h, C.byref(n)))
        out = []
        for i in range(n.value):
            first, count, dev = C.c_uint64(), C.c_uint64(), C.c_int()
            check(lib().lexls_b200_get_shard(self._h, C.c_int(i), None, C.byref(first), C.byref(count), C.byref(dev)))
            out.append((dev.value, first.value, count.value))
        return out

    def close(self):
        if getattr(self, "_h", None):
            lib().lexls_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- configuration ------------------------------------------------------------------------------------------
    def setParameters(self, params):
        check(lib().lexls_b200_set_params(self._h, C.byref(params)))

    def setRegularizationFactor(self, level, factor):
        """lexlsi.h:497 / lexlse.h:1478 — damping factor of one objective (with Params(regularization_type=...))."""
        check(lib().lexls_b200_set_regularization_factor(self._h, C.c_uint32(level), C.c_double(float(factor))))

    def set_stream(self, stream_ptr):
        check(lib().lexls_b200_set_stream(self._h, C.c_void_p(stream_ptr)))

    def _cols(self, level):
        raise NotImplementedError

    def setData(self, level, data):
        """data: [B, cols, m] float64 (each problem column-major m x cols)."""
        m, cols = self.ObjDim[level], self._cols(level)
        a = np.ascontiguousarray(data, dtype=np.float64)
        if a.shape != (self.batch, cols, m):
            raise ValueError("level %d: expected shape %s, got %s" % (level, (self.batch, cols, m), a.shape))
        check(lib().lexls_b200_set_data(self._h, C.c_uint32(level), _dp(a), C.c_uint32(max(m, 1)), C.c_uint64(m * cols)))
        self._hold(a)
        if self.strict:
            self.validate()

    def setData_dev(self, level, ptr, ld=None, problem_stride=None):
        m, cols = self.ObjDim[level], self._cols(level)
        ld = m if ld is None else ld
        problem_stride = ld * cols if problem_stride is None else problem_stride
        check(lib().lexls_b200_set_data_dev(self._h, C.c_uint32(level), C.cast(C.c_void_p(ptr), C.POINTER(C.c_double)),
                                            C.c_uint32(max(ld, 1)), C.c_uint64(problem_stride)))

    # ---- solve --------------------------------------------------------------------------------------------------
    def solve(self, sync=True):
        check(lib().lexls_b200_solve_batch_sync(self._h) if sync else lib().lexls_b200_solve_batch(self._h))

    def synchronize(self):
        check(lib().lexls_b200_synchronize(self._h))
        self._release()

    # ---- results ------------------------------------------------------------------------------------------------
    def get_x(self, out=None):
        """x of every problem, [B, nVar]; `out` may be a caller buffer (e.g. pinned host memory)."""
        x = np.empty((self.batch, self.nVar)) if out is None else out
        if x.shape != (self.batch, self.nVar) or x.dtype != np.float64 or not x.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 [B, nVar] array")
        check(lib().lexls_b200_get_x(self._h, _dp(x)))
        self._release()
        return x

    def get_v(self, level):
        v = np.empty((self.batch, self.ObjDim[level]))
        check(lib().lexls_b200_get_v(self._h, C.c_uint32(level), _dp(v)))
        return v

    def get_status(self):
        s = np.empty(self.batch, dtype=np.int32)
        check(lib().lexls_b200_get_status(self._h, s.ctypes.data_as(C.POINTER(C.c_int32))))
        return s

    def get_counters(self):
        c = np.empty((self.batch, capi.NCOUNTERS), dtype=np.uint32)
        check(lib().lexls_b200_get_counters(self._h, c.ctypes.data_as(C.POINTER(C.c_uint32))))
        return c

    def get_ranks(self):
        G = self.nObj - (1 if (self.ObjType and self.ObjType[0] == 1) else 0)
        r = np.empty((self.batch, G), dtype=np.uint32)
        check(lib().lexls_b200_get_ranks(self._h, r.ctypes.data_as(C.POINTER(C.c_uint32))))
        return r

    def get_x_least_norm(self):
        """solveLeastNorm (lexlse.h:1046-1262): the minimiser of least norm, [B, nVar]; needs request_outputs(least_norm=True)."""
        x = np.empty((self.batch, self.nVar))
        check(lib().lexls_b200_get_x_least_norm(self._h, _dp(x)))
        return x

    def request_outputs(self, lexqr=False, multipliers=False, xstar=False, least_norm=False):
        """Optional outputs written by the next solves: the factorized work matrix (get_lexqr) / LexLSI multipliers /
        LexLSI get_xStar with get_data and get_lexqr of that last equality-constrained problem."""
        check(lib().lexls_b200_request_outputs(self._h, (capi.OUT_LEXQR if lexqr else 0) | (capi.OUT_LAMBDA if multipliers else 0)
                                               | (capi.OUT_XSTAR if xstar else 0) | (capi.OUT_LEAST_NORM if least_norm else 0)))

    def get_lexqr(self):
        """(lexqr [B, nVar+1, M] — i.e. per problem column-major M x (nVar+1) —, hh [B, M])."""
        M = sum(d for k, d in enumerate(self.ObjDim) if not (self.ObjType and self.ObjType[k] == 1))
        q = np.empty((self.batch, self.nVar + 1, M))
        hh = np.empty((self.batch, M))
        check(lib().lexls_b200_get_lexqr(self._h, q.ctypes.data_as(C.POINTER(C.c_double)), hh.ctypes.data_as(C.POINTER(C.c_double))))
        return q, hh

    def device_results(self):
        x, s, c = C.POINTER(C.c_double)(), C.POINTER(C.c_int32)(), C.POINTER(C.c_uint32)()
        check(lib().lexls_b200_device_results(self._h, C.byref(x), C.byref(s), C.byref(c)))
        return (C.cast(x, C.c_void_p).value, C.cast(s, C.c_void_p).value, C.cast(c, C.c_void_p).value)

    def kernel_info(self):
        info = capi.KernelInfo()
        check(lib().lexls_b200_kernel_info(self._h, C.byref(info)))
        return dict(name=info.name.decode(), grid=info.grid, block=info.block, smem_bytes=info.smem_bytes,
                    problems_per_sm=info.problems_per_sm, launches_per_solve=info.launches_per_solve,
                    warps_per_problem=info.warps_per_problem)

    def last_solve_ms(self):
        ms = C.c_float()
        check(lib().lexls_b200_last_solve_ms(self._h, C.byref(ms)))
        return float(ms.value)


class LexLSEBatch(_Batch):
    """Batch of LexLSE problems (reference include/lexls/lexlse.h): setData / fixVariables / factorize+solve."""
    kind = capi.LSE

    def _cols(self, level):
        return self.nVar + 1

    def fixVariables(self, var_index, value, types=None):
        """lexlse.h:1398-1419 for the whole batch: var_index, value [B, nFixed]."""
        vi = np.ascontiguousarray(var_index, dtype=np.uint32).reshape(self.batch, -1)
        vv = np.ascontiguousarray(value, dtype=np.float64).reshape(self.batch, -1)
        tt = np.ascontiguousarray(types, dtype=np.uint8).reshape(self.batch, -1) if types is not None else None
        check(lib().lexls_b200_set_fixed_variables(
            self._h, C.c_uint32(vi.shape[1]), vi.ctypes.data_as(C.POINTER(C.c_uint32)), _dp(vv),
            tt.ctypes.data_as(C.POINTER(C.c_uint8)) if tt is not None else None))

    def get_column_permutations(self):
        p = np.empty((self.batch, self.nVar), dtype=np.uint32)
        check(lib().lexls_b200_get_column_permutations(self._h, p.ctypes.data_as(C.POINTER(C.c_uint32))))
        return p


class LexLSIBatch(_Batch):
    """Batch of LexLSI problems (reference include/lexls/lexlsi.h)."""
    kind = capi.LSI

    def _cols(self, level):
        if self.ObjType and self.ObjType[level] == 1:
            return 2
        return self.nVar + 2

    def setData(self, level, data, var_index=None):
        if self.ObjType and self.ObjType[level] == 1:
            if var_index is None:
                raise ValueError("a simple-bounds level needs var_index")
            m = self.ObjDim[0]
            a = np.ascontiguousarray(data, dtype=np.float64)
            if a.shape != (self.batch, 2, m):
                raise ValueError("simple bounds: expected shape %s, got %s" % ((self.batch, 2, m), a.shape))
            vi = np.ascontiguousarray(var_index, dtype=np.uint32)
            shared = vi.ndim == 1
            if not shared and vi.shape != (self.batch, m):
                raise ValueError("var_index must be [m] or [B, m]")
            check(lib().lexls_b200_set_simple_bounds(self._h, vi.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                     C.c_uint64(0 if shared else m), _dp(a), C.c_uint32(max(m, 1)),
                                                     C.c_uint64(2 * m)))
            self._hold(a, vi)
            if self.strict:
                self.validate()
        else:
            super().setData(level, data)

    def set_active_guess(self, level, types):
        if types is None:
            check(lib().lexls_b200_set_active_guess(self._h, C.c_uint32(level), None))
            return
        t = np.ascontiguousarray(types, dtype=np.uint8)
        if t.shape != (self.batch, self.ObjDim[level]):
            raise ValueError("guess must be [B, m]")
        check(lib().lexls_b200_set_active_guess(self._h, C.c_uint32(level), t.ctypes.data_as(C.POINTER(C.c_uint8))))
        self._hold(t)

    def set_active_guess_ordered(self, obj, ctr, types):
        """api_activate calls in CALL order (lexlsi.h:120-136): obj, ctr, types are [B, count]; type 0 pads ragged lists."""
        if obj is None:
            check(lib().lexls_b200_set_active_guess_ordered(self._h, C.c_uint32(0), None, None, None, C.c_uint64(0)))
            return
        o = np.ascontiguousarray(obj, dtype=np.uint32).reshape(self.batch, -1)
        c = np.ascontiguousarray(ctr, dtype=np.uint32).reshape(self.batch, -1)
        t = np.ascontiguousarray(types, dtype=np.uint8).reshape(self.batch, -1)
        if not (o.shape == c.shape == t.shape):
            raise ValueError("obj, ctr, types must have the same [B, count] shape")
        u32 = C.POINTER(C.c_uint32)
        check(lib().lexls_b200_set_active_guess_ordered(self._h, C.c_uint32(o.shape[1]), o.ctypes.data_as(u32), c.ctypes.data_as(u32),
                                                        t.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint64(o.shape[1])))

    def get_xStar(self):
        """(xstar [B, nVar], data [B, nVar+1, M], lexqr [B, nVar+1, M], hh [B, M]): get_xStar (lexlsi.h:519) and what get_data /
        get_lexqr return after it.  Needs request_outputs(xstar=True) before the solve."""
        M = sum(d for k, d in enumerate(self.ObjDim) if not (self.ObjType and self.ObjType[k] == 1))
        xs = np.empty((self.batch, self.nVar))
        data = np.empty((self.batch, self.nVar + 1, M))
        q = np.empty((self.batch, self.nVar + 1, M))
        hh = np.empty((self.batch, M))
        check(lib().lexls_b200_get_xstar(self._h, _dp(xs), _dp(data), _dp(q), _dp(hh)))
        return xs, data, q, hh

    def set_x0(self, x0):
        if x0 is None:
            check(lib().lexls_b200_set_x0(self._h, None))
            return
        a = np.ascontiguousarray(x0, dtype=np.float64)
        if a.shape != (self.batch, self.nVar):
            raise ValueError("x0 must be [B, nVar]")
        check(lib().lexls_b200_set_x0(self._h, _dp(a)))
        self._hold(a)

    def set_v0(self, level, v0):
        a = np.ascontiguousarray(v0, dtype=np.float64)
        if a.shape != (self.batch, self.ObjDim[level]):
            raise ValueError("v0 must be [B, m]")
        check(lib().lexls_b200_set_v0(self._h, C.c_uint32(level), _dp(a)))

    def reset(self):
        check(lib().lexls_b200_reset(self._h))

    def getConstraintViolation(self, level):
        v = np.empty((self.batch, self.ObjDim[level]))
        check(lib().lexls_b200_get_constraint_violation(self._h, C.c_uint32(level), _dp(v)))
        return v

    def getActiveCtr(self, level):
        t = np.empty((self.batch, self.ObjDim[level]), dtype=np.uint8)
        check(lib().lexls_b200_get_active_set(self._h, C.c_uint32(level), t.ctypes.data_as(C.POINTER(C.c_uint8))))
        return t

    def getLambda(self, level):
        """Multipliers of every objective for the constraints of `level` (getLambda, lexlsi.h:552-605): [B, nObj, m]
        (per problem the reference's m x nObj column-major matrix).  Needs request_outputs(multipliers=True)."""
        lam = np.empty((self.batch, self.nObj, self.ObjDim[level]))
        check(lib().lexls_b200_get_lambda(self._h, C.c_uint32(level), _dp(lam)))
        return lam

    def getWorkingSetLog(self, capacity=64):
        """(entries [B, capacity, 6] = {obj, ctr, type, alpha_or_lambda, rank, cycling_detected}, count [B]); needs
        Params(log_working_set_enabled=1) at solve time (getWorkingSetLog, lexlsi.h:739)."""
        e = np.zeros((self.batch, capacity, 6))
        n = np.zeros(self.batch, dtype=np.uint32)
        check(lib().lexls_b200_get_working_set_log(self._h, C.c_uint32(capacity), _dp(e), n.ctypes.data_as(C.POINTER(C.c_uint32))))
        return e, n

    def getActiveCtr_order(self, level):
        o = np.empty((self.batch, self.ObjDim[level]), dtype=np.uint32)
        check(lib().lexls_b200_get_active_order(self._h, C.c_uint32(level), o.ctypes.data_as(C.POINTER(C.c_uint32))))
        return o


def measure_fp64_peak(device=0):
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    check(lib().lexls_b200_measure_fp64_peak(C.c_int(device), C.byref(a), C.byref(b), C.byref(c)))
    return dict(dfma_tflops=a.value, dmma_tflops=b.value, sm_clock_mhz=c.value)
