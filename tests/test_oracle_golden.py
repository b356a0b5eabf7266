"""CPU: the oracle against every fixture / known-answer test the reference holds for this path."""
import os

import numpy as np
import pytest


def _lsi_from_dat(oracle, h, guess=None, **par):
    data = [np.ascontiguousarray(M.T)[None] for M in h.objectives]
    return oracle.lsi_solve_batch(h.nVar, h.ObjDim, h.ObjType, data, var_index=h.var_index[None] if h.var_index is not None else None,
                                  guess=guess, params=oracle.OracleParams(**par) if par else None)


def test_test01_dat_solution(oracle, golden_dir):
    """tests/test_01.dat: n=88, levels [74 simple bounds, 33, 3, 2, 97]; x must equal the file's #Solution (15 digits)."""
    from lexls_b200 import datfile
    h = datfile.read_hierarchy(os.path.join(golden_dir, "test_01.dat"))
    assert (h.nVar, h.ObjDim, h.ObjType) == (88, [74, 33, 3, 2, 97], [1, 0, 0, 0, 0])
    cold = _lsi_from_dat(oracle, h)
    warm = _lsi_from_dat(oracle, h, guess=[g[None] for g in h.active_set_guess])
    for r in (cold, warm):
        assert r["status"][0] == 0
        assert r["counters"][0, 5] == 88  # total rank: x is unique
        rel = np.linalg.norm(r["x"][0] - h.solution) / np.linalg.norm(h.solution)
        assert rel < 1e-9, rel
    assert cold["counters"][0, 0] > 5 and warm["counters"][0, 0] == 1  # the file's guess is the optimal active set
    # (the two final working sets differ by one weakly active bound, so only x is compared)


def test_escande_multipliers(oracle):
    """interfaces/matlab-octave/tests/lexlsi/lambda_test.m:1-47: (x_1=1) > (2 x_2=1) > ... > (n x_n=1) > (sum x_i = 1):
    x_k = 1/k and the last column of the multipliers is [-w, -w/2, ..., -w/n, w], w = sum_{i=2..n} 1/i."""
    for n in (3, 5, 9):
        dims = [1] * (n + 1)
        data = []
        for k in range(n):
            row = np.zeros(n + 2)
            row[k] = k + 1
            row[n] = row[n + 1] = 1.0
            data.append(row.reshape(n + 2, 1))
        row = np.ones(n + 2)
        data.append(row.reshape(n + 2, 1))
        r = oracle.lsi_solve_detail(n, dims, [0] * (n + 1), data)
        assert r["status"] == 0
        assert np.allclose(r["x"], 1.0 / np.arange(1, n + 1), rtol=0, atol=1e-14)
        w = sum(1.0 / i for i in range(2, n + 1))
        expect = np.array([-w / k for k in range(1, n + 1)] + [w])
        got = np.array([r["lambdas"][k][0, n] for k in range(n + 1)])
        assert np.allclose(got, expect, rtol=0, atol=1e-13), (got, expect)
        # all earlier columns: objective j only sees its own residual (zero here)
        for k in range(n + 1):
            assert np.allclose(r["lambdas"][k][0, :n], 0.0, atol=1e-13)


def test_pinv_identity(oracle):
    """interfaces/matlab-octave/examples/example_lexlse.m:17-29: {A1 x = b1 ; I x = 0}  ==  pinv(A1) b1."""
    rng = np.random.default_rng(0)
    B, m, n = 50, 3, 5
    A = rng.random((B, m, n))
    b = rng.random((B, m))
    l1 = np.concatenate([A, b[:, :, None]], axis=2).transpose(0, 2, 1)
    l2 = np.concatenate([np.broadcast_to(np.eye(n), (B, n, n)), np.zeros((B, n, 1))], axis=2).transpose(0, 2, 1)
    r = oracle.lse_solve_batch(n, [m, n], [np.ascontiguousarray(l1), np.ascontiguousarray(l2)])
    for p in range(B):
        assert np.allclose(r["x"][p], np.linalg.pinv(A[p]) @ b[p], rtol=0, atol=1e-11)


def test_fixed_variables_equal_general_reposing(oracle):
    """interfaces/matlab-octave/tests/lexlse/test_lexlse.m:13-20 + compare_results.m: fixing variables == an extra first level
    of equalities e_i' x = value (tolerance 1e-10 on x and residuals)."""
    rng = np.random.default_rng(1)
    B, n, dims, nf = 40, 8, [4, 5], 3
    data = [np.ascontiguousarray(rng.standard_normal((B, n + 1, m))) for m in dims]
    idx = np.stack([rng.permutation(n)[:nf] for _ in range(B)]).astype(np.uint32)
    val = rng.standard_normal((B, nf))
    a = oracle.lse_solve_batch(n, dims, data, fixed_index=idx, fixed_value=val)
    first = np.zeros((B, n + 1, nf))
    for p in range(B):
        for j in range(nf):
            first[p, idx[p, j], j] = 1.0
            first[p, n, j] = val[p, j]
    b = oracle.lse_solve_batch(n, [nf] + dims, [first] + data)
    assert np.abs(a["x"] - b["x"]).max() < 1e-10
    assert np.abs(a["v"] - b["v"][:, nf:]).max() < 1e-10
    assert np.array_equal(a["total_rank"], b["total_rank"])


def test_lexlse_matches_nullspace_cascade(oracle):
    """Cross-check of factorize+solve+get_v against a cascade of dense least-squares problems in numpy
    (x_k = x_{k-1} + N_{k-1} argmin ||A_k N_{k-1} z - (b_k - A_k x_{k-1})||)."""
    rng = np.random.default_rng(2)
    B, n, dims = 30, 10, [3, 4, 2, 5]
    data = [np.ascontiguousarray(rng.standard_normal((B, n + 1, m))) for m in dims]
    r = oracle.lse_solve_batch(n, dims, data)
    for p in range(B):
        x = np.zeros(n)
        N = np.eye(n)
        off = 0
        for k, m in enumerate(dims):
            A = data[k][p][:-1, :].T
            b = data[k][p][-1, :]
            if N.shape[1] > 0:
                AN = A @ N
                z = np.linalg.lstsq(AN, b - A @ x, rcond=None)[0]
                x = x + N @ z
                u, s, vt = np.linalg.svd(AN)
                rk = int((s > 1e-10).sum())
                N = N @ vt[rk:].T
            res = A @ x - b
            assert abs(np.linalg.norm(r["v"][p, off:off + m]) - np.linalg.norm(res)) < 1e-9
            off += m
        # total rank 10 -> x unique
        assert np.allclose(r["x"][p], x, atol=1e-8)


def test_lsi_final_working_set_equals_lse(oracle):
    """tests/test_numerical_error.cpp:79-137 (procedure): re-solving the final active set as equalities with LexLSE reproduces
    LexLSI's x."""
    from lexls_b200 import problems
    w = problems.workload("tiny_lsi", 48)
    data, vi = problems.generate(w)
    r = oracle.lsi_solve_batch(w.nVar, w.ObjDim, w.ObjType, data, var_index=np.tile(vi, (48, 1)))
    assert (r["status"] == 0).all()
    for p in range(48):
        levels, dims = [], []
        fidx, fval = [], []
        off = 0
        for k, m in enumerate(w.ObjDim):
            order = r["active_order"][p, off:off + m]
            order = order[order != 0xFFFFFFFF]
            types = r["active"][k][p]
            if w.ObjType[k] == 1:
                for i in order:
                    fidx.append(vi[i])
                    fval.append(data[k][p][0, i] if types[i] == 1 else data[k][p][1, i])
            else:
                rows = []
                for i in order:
                    rhs = data[k][p][w.nVar, i] if types[i] == 1 else data[k][p][w.nVar + 1, i]
                    rows.append(np.concatenate([data[k][p][:w.nVar, i], [rhs]]))
                M = np.array(rows).reshape(len(rows), w.nVar + 1)
                levels.append(np.ascontiguousarray(M.T)[None])
                dims.append(len(rows))
            off += m
        e = oracle.lse_solve_batch(w.nVar, dims, levels, fixed_index=np.array([fidx], dtype=np.uint32) if fidx else None,
                                   fixed_value=np.array([fval]) if fidx else None)
        # LexLSI reaches x by steps x += alpha*dx, so the agreement is to rounding, not bitwise
        assert np.abs(e["x"][0] - r["x"][p]).max() < 1e-10


def test_kkt_invariants(oracle):
    """Solved LexLSI problems: bounds hold up to the reported violation, v is consistent, and active-set types are coherent."""
    from lexls_b200 import problems
    w = problems.workload("c3", 64)
    data, vi = problems.generate(w)
    r = oracle.lsi_solve_batch(w.nVar, w.ObjDim, w.ObjType, data, var_index=np.tile(vi, (64, 1)))
    assert (r["status"] == 0).all()
    for k, m in enumerate(w.ObjDim):
        x = r["x"]
        if w.ObjType[k] == 1:
            Ax = x[:, vi]
            lb, ub = data[k][:, 0, :], data[k][:, 1, :]
        else:
            Ax = np.einsum("bcm,bc->bm", data[k][:, :w.nVar, :], x)
            lb, ub = data[k][:, w.nVar, :], data[k][:, w.nVar + 1, :]
        viol = np.where(Ax < lb, Ax - lb, np.where(Ax > ub, Ax - ub, 0.0))
        assert np.abs(viol - r["viol"][k]).max() < 1e-9
        assert np.abs(r["v"][k] - r["viol"][k]).max() < 1e-8       # at the solution v is the minimal violation
        act = r["active"][k]
        assert np.all(np.abs(Ax - lb)[act == 1] <= np.abs(r["v"][k])[act == 1] + 1e-9)
        assert np.all(np.abs(Ax - ub)[act == 2] <= np.abs(r["v"][k])[act == 2] + 1e-9)
        assert np.all(np.abs(lb - ub)[act == 3] < 1e-15)


def test_oracle_is_deterministic_across_threads(oracle):
    from lexls_b200 import problems
    w = problems.workload("tiny_lsi", 32)
    data, vi = problems.generate(w)
    a = oracle.lsi_solve_batch(w.nVar, w.ObjDim, w.ObjType, data, var_index=np.tile(vi, (32, 1)), threads=1)
    b = oracle.lsi_solve_batch(w.nVar, w.ObjDim, w.ObjType, data, var_index=np.tile(vi, (32, 1)), threads=4)
    assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["counters"], b["counters"])


def test_oracle_rejects_bad_bounds(oracle):
    d = np.zeros((1, 4, 2))
    d[0, 2, :] = 1.0   # lb
    d[0, 3, :] = 0.0   # ub < lb
    with pytest.raises(RuntimeError):
        oracle.lsi_solve_batch(2, [2], [0], [d])


def _two_level(rng, B, n, m1, m2):
    A1, b1 = rng.standard_normal((B, m1, n)), rng.standard_normal((B, m1))
    A2, b2 = rng.standard_normal((B, m2, n)), rng.standard_normal((B, m2))
    lv = lambda A, b: np.ascontiguousarray(np.concatenate([A, b[:, :, None]], axis=2).transpose(0, 2, 1))
    return A1, b1, A2, b2, [lv(A1, b1), lv(A2, b2)]


def test_tikhonov_regularization_is_the_constrained_ridge_solution(oracle):
    """REGULARIZATION_TIKHONOV / TIKHONOV_2 / TIKHONOV_1 (lexlse.h:317-331, :378-396) on {A1 x = b1 (full row rank)} >
    {A2 x = b2, factor a}: x = x1 + Z z with z = argmin |A2 (x1 + Z z) - b2|^2 + a^2 |x1 + Z z|^2, Z = null(A1), x1 = pinv(A1) b1."""
    rng = np.random.default_rng(5)
    B, n, m1, m2, a = 20, 7, 3, 6, 0.35
    A1, b1, A2, b2, data = _two_level(rng, B, n, m1, m2)
    for reg_type in (1, 7, 8):
        oracle.set_regularization(reg_type, [0.0, a])
        try:
            r = oracle.lse_solve_batch(n, [m1, m2], data)
        finally:
            oracle.set_regularization(0)
        for p in range(B):
            x1 = np.linalg.pinv(A1[p]) @ b1[p]
            Z = np.linalg.svd(A1[p])[2][m1:].T
            M = np.vstack([A2[p] @ Z, a * Z])
            rhs = np.concatenate([b2[p] - A2[p] @ x1, -a * x1])
            z = np.linalg.lstsq(M, rhs, rcond=None)[0]
            assert np.abs(r["x"][p] - (x1 + Z @ z)).max() < 1e-11, reg_type


def test_conjugate_gradient_regularization_converges_to_the_direct_variants(oracle):
    """REGULARIZATION_TIKHONOV_CG solves the TIKHONOV problem and RT_NO_Z_CG the RT_NO_Z problem by CGLS
    (lexlse.h:2256-2279, :2331-2357, :2367-2550): with enough iterations they return the direct solutions; with the default
    cap of 10 they stop early (typedefs.h:121) and differ."""
    rng = np.random.default_rng(6)
    B, n, m1, m2 = 12, 9, 3, 4
    _, _, _, _, data = _two_level(rng, B, n, m1, m2)
    factors = [0.2, 0.5]

    def run(reg_type, it=10):
        oracle.set_regularization(reg_type, factors, cg_iterations=it)
        try:
            return oracle.lse_solve_batch(n, [m1, m2], data)["x"]
        finally:
            oracle.set_regularization(0)

    for cg, direct in ((2, 1), (6, 5)):
        xd = run(direct)
        assert np.abs(run(cg, 400) - xd).max() < 1e-9
        assert np.abs(run(cg, 1) - xd).max() > 1e-6  # one step is not the solution: the cap is honoured
