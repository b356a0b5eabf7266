"""Shared helpers of the parity tests (GPU result vs CPU oracle on the same inputs)."""
import numpy as np

# north_star tolerance: relative 1e-9 on x and on the per-level residual norms ||v_k||.
# The kernels reproduce the oracle's summation order, so the observed difference is normally exactly 0;
# the tests assert the stated tolerance and report the observed maximum.
RTOL = 1e-9


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(np.linalg.norm(b), 1e-300)
    return float(np.linalg.norm(a - b) / den)


def assert_close_batch(got, ref, what, rtol=RTOL):
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    num = np.linalg.norm((got - ref).reshape(got.shape[0], -1), axis=1)
    den = np.maximum(np.linalg.norm(ref.reshape(ref.shape[0], -1), axis=1), 1e-12)
    worst = float((num / den).max()) if got.size else 0.0
    assert worst <= rtol, "%s: worst relative error %.3e > %.1e (problem %d)" % (what, worst, rtol, int((num / den).argmax()))
    return worst


def solve_lse_gpu(w_nVar, ObjDim, data, fixed=None, tol=None):
    import lexls_b200 as lx
    B = data[0].shape[0]
    with lx.LexLSEBatch(w_nVar, ObjDim, batch=B) as s:
        if tol is not None:
            s.setParameters(lx.Params(tol_linear_dependence=tol))
        if fixed is not None:
            s.fixVariables(*fixed)
        for k, d in enumerate(data):
            s.setData(k, d)
        s.solve()
        out = dict(x=s.get_x(), v=np.concatenate([s.get_v(k) for k in range(len(ObjDim))], axis=1), ranks=s.get_ranks(),
                   counters=s.get_counters(), colperm=s.get_column_permutations(), status=s.get_status(), info=s.kernel_info())
    return out


def solve_lsi_gpu(nVar, ObjDim, ObjType, data, var_index=None, guess=None, x0=None, v0=None, params=None):
    import lexls_b200 as lx
    B = data[0].shape[0]
    with lx.LexLSIBatch(nVar, ObjDim, ObjType, batch=B) as s:
        if params is not None:
            s.setParameters(params)
        for k, d in enumerate(data):
            if ObjType is not None and ObjType[k] == 1:
                s.setData(k, d, var_index=var_index)
            else:
                s.setData(k, d)
        if x0 is not None:
            s.set_x0(x0)
        if v0 is not None:
            for k, a in enumerate(v0):
                if a is not None:
                    s.set_v0(k, a)
        if guess is not None:
            for k, g in enumerate(guess):
                if g is not None:
                    s.set_active_guess(k, g)
        s.solve()
        P = len(ObjDim)
        out = dict(x=s.get_x(), v=[s.get_v(k) for k in range(P)], viol=[s.getConstraintViolation(k) for k in range(P)],
                   active=[s.getActiveCtr(k) for k in range(P)], order=[s.getActiveCtr_order(k) for k in range(P)],
                   status=s.get_status(), counters=s.get_counters(), ranks=s.get_ranks(), info=s.kernel_info())
    return out


def to_params(lx, **kw):
    return lx.Params(**kw)


def oracle_params(oracle, **kw):
    return oracle.OracleParams(**kw)


def compare_lsi(got, ref, P, rtol=RTOL):
    """Discrete outcomes must be identical; values within rtol."""
    assert np.array_equal(got["status"], ref["status"]), "termination status differs"
    for name, idx in (("factorizations", 0), ("activations", 1), ("deactivations", 2), ("iterations", 3), ("cycling", 4),
                      ("total rank", 5), ("active count", 6)):
        assert np.array_equal(got["counters"][:, idx], ref["counters"][:, idx]), name + " count differs"
    off = 0
    for k in range(P):
        assert np.array_equal(got["active"][k], ref["active"][k]), "final active set differs at level %d" % k
        m = got["active"][k].shape[1]
        assert np.array_equal(got["order"][k], ref["active_order"][:, off:off + m]), "working-set order differs at level %d" % k
        off += m
    worst = assert_close_batch(got["x"], ref["x"], "x", rtol)
    for k in range(P):
        if got["v"][k].shape[1] == 0:
            continue
        nv_g = np.linalg.norm(got["v"][k], axis=1)
        nv_r = np.linalg.norm(ref["v"][k], axis=1)
        assert np.all(np.abs(nv_g - nv_r) <= rtol * np.maximum(nv_r, 1e-9)), "||v_%d|| differs" % k
        nv_g = np.linalg.norm(got["viol"][k], axis=1)
        nv_r = np.linalg.norm(ref["viol"][k], axis=1)
        assert np.all(np.abs(nv_g - nv_r) <= rtol * np.maximum(nv_r, 1e-9)), "||violation_%d|| differs" % k
    return worst
