"""GPU parity, round 2: the BASELINE batches at FULL size, get_xStar / get_data, api_activate in call order, multi-device
handles, deferred input validation.  Everything goes through the C-ABI (lexls_b200.batch -> liblexls_b200.so)."""
import numpy as np
import pytest

from helpers import compare_lsi, solve_lse_gpu, solve_lsi_gpu

pytestmark = pytest.mark.gpu


def _lsi_both(oracle, w, data, vi, guess=None, x0=None, **par):
    import lexls_b200 as lx
    B = data[0].shape[0]
    vib = np.tile(vi, (B, 1)) if (vi is not None and vi.ndim == 1) else vi
    ref = oracle.lsi_solve_batch(w.nVar, w.ObjDim, w.ObjType, data, var_index=vib, guess=guess, x0=x0,
                                 params=oracle.OracleParams(**par) if par else None, threads=0)
    got = solve_lsi_gpu(w.nVar, w.ObjDim, w.ObjType, data, var_index=vi, guess=guess, x0=x0, params=lx.Params(**par) if par else None)
    return got, ref


# ---- full BASELINE sizes ------------------------------------------------------------------------
def test_c2_full_size_65536_bit_identical(oracle):
    from lexls_b200 import problems
    w = problems.workload("c2")
    assert w.batch == 65536
    data, _ = problems.generate(w)
    ref = oracle.lse_solve_batch(w.nVar, w.ObjDim, data, threads=0)
    got = solve_lse_gpu(w.nVar, w.ObjDim, data)
    assert np.array_equal(got["ranks"], ref["ranks"])
    assert np.array_equal(got["x"], ref["x"]) and np.array_equal(got["v"], ref["v"])
    assert (got["ranks"][:, :2] < 20).any(axis=1).sum() > 1000  # the rank-deficient tenth of the batch is really there


def test_c3_full_size_16384_cold_and_warm(oracle):
    from lexls_b200 import problems
    w = problems.workload("c3")
    assert w.batch == 16384
    data, vi = problems.generate(w)
    got, ref = _lsi_both(oracle, w, data, vi)
    compare_lsi(got, ref, len(w.ObjDim))
    assert np.array_equal(got["x"], ref["x"]) and (got["status"] == 0).all()
    d2 = problems.perturb(w, data)
    got2, ref2 = _lsi_both(oracle, w, d2, vi, guess=got["active"], x0=got["x"])
    compare_lsi(got2, ref2, len(w.ObjDim))
    assert np.array_equal(got2["x"], ref2["x"]) and (got2["status"] == 0).all()
    assert got2["counters"][:, 0].mean() < 1.5 < got["counters"][:, 0].mean()


def test_c4_64_problems_solved_at_full_factorization_count(oracle):
    """n = 200, 10 levels x 40 rows, cold start: 310-380 factorizations per problem until PROBLEM_SOLVED (the default cap of 200
    would stop every one of them), work matrix in the global scratch, Schur update on the FP64 tensor cores."""
    from lexls_b200 import problems
    w = problems.workload("c4", 64)
    data, vi = problems.generate(w)
    got, ref = _lsi_both(oracle, w, data, vi, max_number_of_factorizations=1000)
    assert got["info"]["name"].endswith("_lodglobal")
    compare_lsi(got, ref, len(w.ObjDim))
    assert np.array_equal(got["x"], ref["x"])
    assert (got["status"] == 0).all() and got["counters"][:, 0].min() > 200


@pytest.mark.parametrize("dmma", ["0", "1"])
def test_global_work_matrix_dmma_and_dfma_agree_with_the_oracle(oracle, monkeypatch, dmma):
    """The Schur update of the global-work-matrix kernels: mma.sync m8n8k4 f64 tiles (default) and the DFMA strips give the same
    bits as the oracle's ascending-k chain."""
    monkeypatch.setenv("LEXLS_B200_DMMA", dmma)
    monkeypatch.setenv("LEXLS_B200_FORCE_LODGLOBAL", "1")
    from lexls_b200 import problems
    rng = np.random.default_rng(5)
    for n, dims in ((37, [9, 14, 6, 11]), (70, [33, 5, 41]), (21, [3, 30])):
        B = 19
        data = [np.ascontiguousarray(rng.standard_normal((B, n + 1, m))) for m in dims]
        ref = oracle.lse_solve_batch(n, dims, data, threads=0)
        got = solve_lse_gpu(n, dims, data)
        assert got["info"]["name"].endswith("_lodglobal")
        assert np.array_equal(got["x"], ref["x"]) and np.array_equal(got["v"], ref["v"])
    w = problems.workload("c4", 4)
    data, vi = problems.generate(w)
    got, ref = _lsi_both(oracle, w, data, vi, max_number_of_factorizations=12)
    compare_lsi(got, ref, len(w.ObjDim))
    assert np.array_equal(got["x"], ref["x"])


# ---- get_xStar / get_data / get_lexqr after it (lexlsi.h:519, :607, :612) --------------------------------------------
@pytest.mark.parametrize("name,count,cap", [("tiny_lsi", 24, None), ("c3", 16, None), ("c3", 12, 3)])
def test_get_xstar_and_get_data_match_the_oracle(oracle, name, count, cap):
    import lexls_b200 as lx
    from lexls_b200 import problems
    w = problems.workload(name, count)
    data, vi = problems.generate(w)
    par = dict(max_number_of_factorizations=cap) if cap else {}
    vib = np.tile(vi, (w.batch, 1))
    ref = oracle.lsi_solve_batch(w.nVar, w.ObjDim, w.ObjType, data, var_index=vib, params=oracle.OracleParams(**par) if par else None,
                                 threads=0, want_xstar=True)
    with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch) as s:
        if par:
            s.setParameters(lx.Params(**par))
        s.request_outputs(xstar=True, lexqr=True, multipliers=True)
        for k, d in enumerate(data):
            s.setData(k, d, var_index=vi) if w.ObjType[k] == 1 else s.setData(k, d)
        s.solve()
        xs, dat, q, hh = s.get_xStar()
        x, status = s.get_x(), s.get_status()
    assert np.array_equal(status, ref["status"]) and np.array_equal(x, ref["x"])
    if cap:
        assert (status == 2).any()  # some problems stopped by the cap: xStar belongs to a working set that was never factorized
    assert np.array_equal(xs, ref["xstar"]), "get_xStar differs from the oracle"
    assert np.array_equal(dat, ref["lexdata"]), "get_data differs from the oracle"
    assert np.array_equal(q, ref["lexqr_star"]), "get_lexqr after get_xStar differs from the oracle"
    solved = status == 0
    assert np.allclose(xs[solved], x[solved], rtol=0, atol=1e-9)  # a solved problem ends on the solution of its last LexLSE


# ---- api_activate in CALL order (lexlsi.h:120-136 -> workingset.h:67-77) ----------------------------------------------
def test_active_guess_in_call_order_matches_the_oracle(oracle):
    import lexls_b200 as lx
    from lexls_b200 import problems
    w = problems.workload("c3", 48)
    data, vi = problems.generate(w)
    vib = np.tile(vi, (w.batch, 1))
    cold = oracle.lsi_solve_batch(w.nVar, w.ObjDim, w.ObjType, data, var_index=vib, threads=0)
    # the final active inequalities of the cold solve, shuffled per problem = an arbitrary call order of api_activate
    rng = np.random.default_rng(3)
    lists = []
    for p in range(w.batch):
        e = [(k, i, int(cold["active"][k][p, i])) for k in range(len(w.ObjDim)) for i in range(w.ObjDim[k])
             if cold["active"][k][p, i] in (1, 2)]
        rng.shuffle(e)
        lists.append(e)
    cnt = max(len(e) for e in lists) + 2
    obj = np.zeros((w.batch, cnt), dtype=np.uint32)
    ctr = np.zeros((w.batch, cnt), dtype=np.uint32)
    typ = np.zeros((w.batch, cnt), dtype=np.uint8)
    for p, e in enumerate(lists):
        for j, (k, i, t) in enumerate(e):
            obj[p, j], ctr[p, j], typ[p, j] = k, i, t
    d2 = problems.perturb(w, data)
    ref = oracle.lsi_solve_batch(w.nVar, w.ObjDim, w.ObjType, d2, var_index=vib, x0=cold["x"], threads=0, guess_ordered=(obj, ctr, typ))
    sorted_ref = oracle.lsi_solve_batch(w.nVar, w.ObjDim, w.ObjType, d2, var_index=vib, x0=cold["x"], threads=0, guess=cold["active"])
    with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch) as s:
        for k, d in enumerate(d2):
            s.setData(k, d, var_index=vi) if w.ObjType[k] == 1 else s.setData(k, d)
        s.set_x0(cold["x"])
        s.set_active_guess_ordered(obj, ctr, typ)
        s.solve()
        P = len(w.ObjDim)
        got = dict(x=s.get_x(), v=[s.get_v(k) for k in range(P)], viol=[s.getConstraintViolation(k) for k in range(P)],
                   active=[s.getActiveCtr(k) for k in range(P)], order=[s.getActiveCtr_order(k) for k in range(P)],
                   status=s.get_status(), counters=s.get_counters())
        # switching back to the per-row form replaces the list
        for k in range(P):
            s.set_active_guess(k, cold["active"][k])
        s.solve()
        x_sorted = s.get_x()
    compare_lsi(got, ref, len(w.ObjDim))
    assert np.array_equal(got["x"], ref["x"])
    assert np.array_equal(x_sorted, sorted_ref["x"])
    # the call order is visible in the working-set order (and may be in the last bits of x)
    assert not np.array_equal(ref["active_order"], sorted_ref["active_order"])


# ---- multi-device handles (lexls_b200_create_batch_multi) -------------------------------------------------------------
def _device_count():
    import ctypes as C
    from lexls_b200 import capi
    n = C.c_int()
    capi.check(capi.lib().lexls_b200_device_count(C.byref(n)))
    return n.value


@pytest.mark.parametrize("ndev", [1, 2, 4, 8])
def test_multi_device_batch_equals_single_device(oracle, ndev):
    import lexls_b200 as lx
    from lexls_b200 import problems
    if _device_count() < ndev:
        pytest.skip("needs %d CUDA devices" % ndev)
    devices = list(range(ndev))
    # LexLSE: ragged split (batch not a multiple of the device count), fixed variables, factorization output
    w = problems.workload("c2", 1003)
    data, _ = problems.generate(w)
    ref = oracle.lse_solve_batch(w.nVar, w.ObjDim, data, threads=0, want_factors=True)
    with lx.LexLSEBatch(w.nVar, w.ObjDim, batch=w.batch, devices=devices) as s:
        sh = s.shards()
        assert len(sh) == ndev and sum(c for _, _, c in sh) == w.batch and [d for d, _, _ in sh] == devices
        s.request_outputs(lexqr=True)
        for k, d in enumerate(data):
            s.setData(k, d)
        s.solve()
        x, ranks = s.get_x(), s.get_ranks()
        v = np.concatenate([s.get_v(k) for k in range(len(w.ObjDim))], axis=1)
        q, hh = s.get_lexqr()
        assert s.last_solve_ms() > 0
    assert np.array_equal(x, ref["x"]) and np.array_equal(v, ref["v"]) and np.array_equal(ranks, ref["ranks"])
    assert np.array_equal(q, ref["lexqr"]) and np.array_equal(hh, ref["hh"])
    # LexLSI: cold, then warm start through the same multi-device handle
    w = problems.workload("c3", 517)
    data, vi = problems.generate(w)
    vib = np.tile(vi, (w.batch, 1))
    ref = oracle.lsi_solve_batch(w.nVar, w.ObjDim, w.ObjType, data, var_index=vib, threads=0)
    P = len(w.ObjDim)
    with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch, devices=devices, strict=False) as s:
        def run():
            s.solve()
            return dict(x=s.get_x(), v=[s.get_v(k) for k in range(P)], viol=[s.getConstraintViolation(k) for k in range(P)],
                        active=[s.getActiveCtr(k) for k in range(P)], order=[s.getActiveCtr_order(k) for k in range(P)],
                        status=s.get_status(), counters=s.get_counters())
        for k, d in enumerate(data):
            s.setData(k, d, var_index=vi) if w.ObjType[k] == 1 else s.setData(k, d)
        got = run()
        compare_lsi(got, ref, P)
        assert np.array_equal(got["x"], ref["x"])
        d2 = problems.perturb(w, data)
        ref2 = oracle.lsi_solve_batch(w.nVar, w.ObjDim, w.ObjType, d2, var_index=vib, guess=got["active"], x0=got["x"], threads=0)
        for k, d in enumerate(d2):
            s.setData(k, d, var_index=vi) if w.ObjType[k] == 1 else s.setData(k, d)
        for k in range(P):
            s.set_active_guess(k, got["active"][k])
        s.set_x0(got["x"])
        got2 = run()
        compare_lsi(got2, ref2, P)
        assert np.array_equal(got2["x"], ref2["x"])


@pytest.mark.parametrize("ndev", [1, 2])
def test_sharded_batch_python(oracle, ndev):
    """lexls_b200.ShardedBatch (one single-device handle per GPU inside one Python process) against the oracle."""
    import lexls_b200 as lx
    from lexls_b200 import problems
    if _device_count() < ndev:
        pytest.skip("needs %d CUDA devices" % ndev)
    w = problems.workload("c2", 301)
    data, _ = problems.generate(w)
    ref = oracle.lse_solve_batch(w.nVar, w.ObjDim, data, threads=0)
    with lx.ShardedBatch(lx.LexLSEBatch, w.nVar, w.ObjDim, batch=w.batch, devices=list(range(ndev))) as s:
        for k, d in enumerate(data):
            s.setData(k, d)
        s.solve()
        x = s.gather("get_x")
        ranks = s.gather("get_ranks")
    assert np.array_equal(x, ref["x"]) and np.array_equal(ranks, ref["ranks"])


# ---- asynchronous copies: staging arena + deferred repacks ------------------------------------------------------------
@pytest.mark.gpu
def test_async_copies_defer_the_repacks_and_give_the_same_bits():
    """set_async_copies: every host input goes into its own arena slice and its repack is launched by the next solve / read.
    Several rounds (the arena grows and wraps during the first ones), results equal the default mode bit for bit; a repeated
    setData of one level between solves, a result read with inputs pending, invalid inputs, and the switch back."""
    import lexls_b200 as lx
    from lexls_b200 import problems
    w = problems.workload("c3", 300)
    data, vi = problems.generate(w)
    rounds = [data, problems.perturb(w, data), data, problems.perturb(w, problems.perturb(w, data))]
    nl = len(w.ObjDim)

    def feed(s, d):
        for k, a in enumerate(d):
            s.setData(k, a, var_index=vi) if w.ObjType[k] == 1 else s.setData(k, a)

    ref = []
    with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch) as s:
        for d in rounds:
            feed(s, d)
            s.solve()
            ref.append((s.get_x(), s.get_status(), [s.getActiveCtr(k) for k in range(nl)]))
    with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch, strict=False) as s:
        s.set_async_copies(True)
        for r, d in enumerate(rounds):
            feed(s, d)
            if r == 1:
                s.setData(1, rounds[0][1])  # the same level twice in one round: the later call wins
                s.setData(1, d[1])
            s.wait_copies()
            if r == 2:
                assert s.get_status().shape == (w.batch,)  # a read with inputs pending flushes them first
            s.solve()
            x, st = s.get_x(), s.get_status()
            assert np.array_equal(x, ref[r][0]) and np.array_equal(st, ref[r][1])
            for k in range(nl):
                assert np.array_equal(s.getActiveCtr(k), ref[r][2][k])
        # warm start through the arena (guess + x0 are host inputs too)
        feed(s, rounds[1])
        for k in range(nl):
            s.set_active_guess(k, ref[0][2][k])
        s.set_x0(ref[0][0])
        s.solve()
        xw = s.get_x()
        bad = rounds[0][1].copy()
        bad[7, w.nVar, 0], bad[7, w.nVar + 1, 0] = 2.0, 1.0  # lb > ub in one problem
        s.setData(1, bad)
        with pytest.raises(lx.LexLSError):
            s.solve()
        s.set_async_copies(False)
        s.reset()  # forget the guess and x0 of the warm start above
        feed(s, rounds[0])
        s.solve()
        assert np.array_equal(s.get_x(), ref[0][0])
    with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch) as s:
        feed(s, rounds[1])
        for k in range(nl):
            s.set_active_guess(k, ref[0][2][k])
        s.set_x0(ref[0][0])
        s.solve()
        assert np.array_equal(s.get_x(), xw)


# ---- deferred validation: one synchronisation per batch ---------------------------------------------------------------
def test_deferred_validation_reports_at_solve():
    import lexls_b200 as lx
    with lx.LexLSIBatch(4, [3, 2], [1, 0], batch=5, strict=False) as s:
        bad = np.zeros((5, 2, 3))
        bad[3, 0, 1] = 1.0  # lb > ub in one problem of the batch
        s.setData(0, bad, var_index=np.array([0, 1, 2], dtype=np.uint32))  # not reported here ...
        s.setData(1, np.zeros((5, 6, 2)))
        with pytest.raises(lx.LexLSError) as e:
            s.solve()  # ... but here: the device refuses the solve, the first synchronisation reports it
        assert e.value.code == 4 and "(simple)" in str(e.value)
        with pytest.raises(lx.LexLSError):
            s.solve()  # the offending data must be set again
        ok = np.zeros((5, 2, 3))
        ok[:, 1, :] = 1.0
        s.setData(0, ok, var_index=np.array([0, 1, 2], dtype=np.uint32))
        gen = np.zeros((5, 6, 2))
        gen[:, 4, :] = 2.0   # lb
        gen[:, 5, :] = 1.0   # ub < lb
        s.setData(1, gen)
        with pytest.raises(lx.LexLSError) as e:
            s.validate()
        assert e.value.code == 4 and "(general)" in str(e.value)
        gen[:, 5, :] = 3.0
        s.setData(1, gen)
        s.setData(0, ok, var_index=np.array([0, 1, 2], dtype=np.uint32))
        s.solve()
        assert (s.get_status() == 0).all()


def test_working_set_log_count_never_exceeds_capacity(oracle):
    import lexls_b200 as lx
    from lexls_b200 import problems
    w = problems.workload("c3", 8)
    data, vi = problems.generate(w)
    with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch) as s:
        s.setParameters(lx.Params(log_working_set_enabled=1))
        for k, d in enumerate(data):
            s.setData(k, d, var_index=vi) if w.ObjType[k] == 1 else s.setData(k, d)
        s.solve()
        full, n_full = s.getWorkingSetLog(capacity=256)
        cut, n_cut = s.getWorkingSetLog(capacity=3)
    assert n_full.max() > 3 and (n_cut == np.minimum(n_full, 3)).all()
    assert np.array_equal(cut[:, :3, :], full[:, :3, :])


# ---- deactivate_first_wrong_sign (lexlsi.h:1048-1105) -----------------------------------------------------------------
def test_deactivate_first_wrong_sign_matches_the_oracle(oracle):
    """Removal rule 'first': of the constraints with a wrong-sign multiplier, the one activated first (the reference's
    chronological WS list) — different trajectories from the default 'largest' rule, identical to the oracle's."""
    import lexls_b200 as lx
    from lexls_b200 import problems
    rng = np.random.default_rng(17)
    # general hierarchies only: 4 levels of inequalities / equalities, n = 14
    n, dims, B = 14, [6, 5, 7, 14], 96
    data = []
    for k, m in enumerate(dims):
        A = rng.standard_normal((B, m, n))
        ub = rng.random((B, m))
        lb = ub - rng.random((B, m)) if k != 1 else ub
        data.append(np.ascontiguousarray(np.concatenate([A, lb[:, :, None], ub[:, :, None]], axis=2).transpose(0, 2, 1)))

    class W:
        nVar, ObjDim, ObjType = n, dims, [0, 0, 0, 0]
    par = dict(deactivate_first_wrong_sign=1, max_number_of_factorizations=300, log_working_set_enabled=1)
    got, ref = _lsi_both(oracle, W, data, None, **par)
    compare_lsi(got, ref, len(dims))
    assert np.array_equal(got["x"], ref["x"])
    assert got["counters"][:, 2].sum() > 0  # constraints really were removed
    largest = oracle.lsi_solve_batch(n, dims, [0] * 4, data, params=oracle.OracleParams(max_number_of_factorizations=300), threads=0)
    assert not np.array_equal(largest["counters"][:, :4], ref["counters"][:, :4])  # the rule matters
    # warm start with a call list: the guess enters WS in call order
    w = problems.workload("c4", 6)
    d4, _ = problems.generate(w)
    got4, ref4 = _lsi_both(oracle, w, d4, None, deactivate_first_wrong_sign=1, max_number_of_factorizations=40)
    compare_lsi(got4, ref4, len(w.ObjDim))
    assert np.array_equal(got4["x"], ref4["x"])


def test_deactivate_first_wrong_sign_is_refused_with_simple_bounds():
    import lexls_b200 as lx
    with lx.LexLSIBatch(4, [3, 2], [1, 0], batch=2) as s:
        with pytest.raises(lx.LexLSError) as e:
            s.setParameters(lx.Params(deactivate_first_wrong_sign=1))
        assert e.value.code == 3 and "lexlse.h:596" in str(e.value)
    with lx.LexLSIBatch(4, [3, 2], [0, 0], batch=2) as s:
        s.setParameters(lx.Params(deactivate_first_wrong_sign=1))


def test_result_buffers_keep_their_guard_bands(monkeypatch):
    """LEXLS_B200_CANARY=1 (stand-in for compute-sanitizer memcheck, which is closed on the GPU pool): every device buffer of a
    handle is followed by a guard band; after solves of every kernel family no band may have been written into."""
    import lexls_b200 as lx
    from lexls_b200 import problems
    monkeypatch.setenv("LEXLS_B200_CANARY", "1")
    rng = np.random.default_rng(2)

    def lse(n, dims, B, **envs):
        for k, v in envs.items():
            monkeypatch.setenv("LEXLS_B200_" + k, str(v))
        data = [np.ascontiguousarray(rng.standard_normal((B, n + 1, m))) for m in dims]
        with lx.LexLSEBatch(n, dims, batch=B) as s:
            s.request_outputs(lexqr=True)
            s.fixVariables(np.tile(np.array([1, 0], dtype=np.uint32), (B, 1)), rng.standard_normal((B, 2)))
            for k, d in enumerate(data):
                s.setData(k, d)
            s.solve()
            s.get_x(), s.get_lexqr()
            bad = s.check_guard_bands()
            name = s.kernel_info()["name"]
        for k in envs:
            monkeypatch.delenv("LEXLS_B200_" + k)
        return name, bad

    def lsi(name, B, **envs):
        for k, v in envs.items():
            monkeypatch.setenv("LEXLS_B200_" + k, str(v))
        w = problems.workload(name, B)
        data, vi = problems.generate(w)
        with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=B) as s:
            s.setParameters(lx.Params(log_working_set_enabled=1, max_number_of_factorizations=8))
            s.request_outputs(lexqr=True, multipliers=True, xstar=True)
            for k, d in enumerate(data):
                s.setData(k, d, var_index=vi) if w.ObjType[k] == 1 else s.setData(k, d)
            s.solve()
            s.get_x(), s.get_xStar(), s.getWorkingSetLog(16)
            bad = s.check_guard_bands()
            kname = s.kernel_info()["name"]
        for k in envs:
            monkeypatch.delenv("LEXLS_B200_" + k)
        return kname, bad

    seen = {}
    for res in (lse(40, [20, 20, 20], 33), lse(12, [5, 4, 6], 7), lse(30, [9, 14, 6], 9, WARPS=2), lse(30, [9, 14, 6], 9, WARPS=4),
                lse(30, [9, 14, 6], 5, WARPS=8), lse(37, [9, 14, 6, 11], 5, FORCE_LODGLOBAL=1, LG_WARPS=8),
                lse(37, [9, 14, 6, 11], 5, FORCE_LODGLOBAL=1, LG_WARPS=16, DMMA=0),
                lsi("tiny_lsi", 9), lsi("c3", 11), lsi("c3", 5, WARPS=1), lsi("c3", 5, WARPS=4), lsi("c3", 5, WARPS=8),
                lsi("c3", 5, DATA="smem", WARPS=4), lsi("c3", 4, FORCE_LODGLOBAL=1, LG_WARPS=8), lsi("c3", 4, FORCE_LODGLOBAL=1, LG_WARPS=16),
                lsi("c4", 2)):
        seen[res[0]] = seen.get(res[0], 0) + res[1]
    assert len(seen) >= 12, seen
    assert all(v == 0 for v in seen.values()), seen


# ---- regularization (lexlse.h:274-411, :1700-1760, :2078-2251, :2592-2625) ------------------------------------------------
REG_TYPES = {1: "TIKHONOV", 2: "TIKHONOV_CG", 3: "R", 4: "R_NO_Z", 5: "RT_NO_Z", 6: "RT_NO_Z_CG", 7: "TIKHONOV_1", 8: "TIKHONOV_2", 9: "TEST"}


@pytest.mark.parametrize("reg_type", sorted(REG_TYPES))
@pytest.mark.parametrize("variant", ["w2", "w4", "lg8", "lg16"])
def test_regularized_lexlse_matches_the_oracle(oracle, monkeypatch, reg_type, variant):
    import lexls_b200 as lx
    env = {"w2": dict(WARPS=2), "w4": dict(WARPS=4), "lg8": dict(FORCE_LODGLOBAL=1, LG_WARPS=8), "lg16": dict(FORCE_LODGLOBAL=1, LG_WARPS=16)}[variant]
    for k, v in env.items():
        monkeypatch.setenv("LEXLS_B200_" + k, str(v))
    rng = np.random.default_rng(100 + reg_type)
    for n, dims, fixed in ((14, [4, 3, 5], False), (11, [3, 4, 8], True), (9, [2, 12], False)):
        B = 7
        data = [np.ascontiguousarray(rng.standard_normal((B, n + 1, m))) for m in dims]
        data[0][::3, 1, :] = data[0][::3, 0, :]  # a duplicated column: lower rank, ties
        factors = [0.3, 0.05, 0.7][:len(dims)]
        fi = np.tile(np.array([2, 0], dtype=np.uint32), (B, 1)) if fixed else None
        fv = rng.standard_normal((B, 2)) if fixed else None
        oracle.set_regularization(reg_type, factors)
        try:
            ref = oracle.lse_solve_batch(n, dims, data, fixed_index=fi, fixed_value=fv, threads=0, want_factors=True)
        finally:
            oracle.set_regularization(0)
        with lx.LexLSEBatch(n, dims, batch=B) as s:
            s.setParameters(lx.Params(regularization_type=reg_type))
            for k, f in enumerate(factors):
                s.setRegularizationFactor(k, f)
            s.request_outputs(lexqr=True)
            if fixed:
                s.fixVariables(fi, fv)
            for k, d in enumerate(data):
                s.setData(k, d)
            s.solve()
            x, ranks = s.get_x(), s.get_ranks()
            q, hh = s.get_lexqr()
            name = s.kernel_info()["name"]
        assert not name.endswith("_w1"), name
        assert np.array_equal(ranks, ref["ranks"])
        assert np.array_equal(q, ref["lexqr"]), "factorization (incl. the regularized right-hand sides) differs: %s %s" % (REG_TYPES[reg_type], name)
        assert np.array_equal(x, ref["x"])
    # unregularized reference differs (the factors really act), and a zero factor reproduces it exactly
    plain = oracle.lse_solve_batch(n, dims, data, threads=0)
    assert not np.array_equal(plain["x"], ref["x"])
    with lx.LexLSEBatch(n, dims, batch=B) as s:
        s.setParameters(lx.Params(regularization_type=reg_type))
        for k, d in enumerate(data):
            s.setData(k, d)
        s.solve()
        assert np.array_equal(s.get_x(), plain["x"])


def test_variable_regularization_factor_matches_the_oracle(oracle):
    import lexls_b200 as lx
    rng = np.random.default_rng(77)
    n, dims, B = 12, [4, 4, 6], 16
    data = [np.ascontiguousarray(rng.standard_normal((B, n + 1, m))) for m in dims]
    data[1][:, :, 3] = data[1][:, :, 2] * (1 + 1e-7)  # nearly dependent rows: a badly conditioned level
    oracle.set_regularization(8, [0.2, 0.4, 0.1], variable_factor=0.5)
    try:
        ref = oracle.lse_solve_batch(n, dims, data, threads=0)
    finally:
        oracle.set_regularization(0)
    with lx.LexLSEBatch(n, dims, batch=B) as s:
        s.setParameters(lx.Params(regularization_type=8, variable_regularization_factor=0.5))
        for k, f in enumerate([0.2, 0.4, 0.1]):
            s.setRegularizationFactor(k, f)
        for k, d in enumerate(data):
            s.setData(k, d)
        s.solve()
        assert np.array_equal(s.get_x(), ref["x"])


@pytest.mark.parametrize("reg_type", [1, 2, 4, 5, 6])
def test_regularized_lexlsi_matches_the_oracle(oracle, reg_type):
    """The active-set loop with a damped inner LexLSE (lexlsi.h:331-333, :497; objective.h:489-493: the simple-bounds objective
    is never regularized)."""
    import lexls_b200 as lx
    from lexls_b200 import problems
    w = problems.workload("c3", 64)
    data, vi = problems.generate(w)
    vib = np.tile(vi, (w.batch, 1))
    factors = [9.0, 0.0, 0.02, 0.0, 0.05]  # objective 0 is SIMPLE_BOUNDS: its factor must be ignored
    par = dict(regularization_type=reg_type, max_number_of_factorizations=60)
    oracle.set_regularization(reg_type, factors)
    try:
        ref = oracle.lsi_solve_batch(w.nVar, w.ObjDim, w.ObjType, data, var_index=vib, params=oracle.OracleParams(max_number_of_factorizations=60), threads=0)
    finally:
        oracle.set_regularization(0)
    P = len(w.ObjDim)
    with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch) as s:
        s.setParameters(lx.Params(**par))
        for k, f in enumerate(factors):
            s.setRegularizationFactor(k, f)
        for k, d in enumerate(data):
            s.setData(k, d, var_index=vi) if w.ObjType[k] == 1 else s.setData(k, d)
        s.solve()
        got = dict(x=s.get_x(), v=[s.get_v(k) for k in range(P)], viol=[s.getConstraintViolation(k) for k in range(P)],
                   active=[s.getActiveCtr(k) for k in range(P)], order=[s.getActiveCtr_order(k) for k in range(P)],
                   status=s.get_status(), counters=s.get_counters())
    compare_lsi(got, ref, P)
    assert np.array_equal(got["x"], ref["x"])
    plain = oracle.lsi_solve_batch(w.nVar, w.ObjDim, w.ObjType, data, var_index=vib, threads=0)
    assert not np.array_equal(plain["x"], ref["x"])


def test_unknown_regularization_types_are_refused():
    import lexls_b200 as lx
    with lx.LexLSEBatch(4, [2, 2], batch=1) as s:
        for t in (-1, 10, 12):
            with pytest.raises(lx.LexLSError):
                s.setParameters(lx.Params(regularization_type=t))


def test_conjugate_gradient_iteration_cap_matches_the_oracle(oracle):
    """max_number_of_CG_iterations (typedefs.h:91): 1, 3 and 40 steps of the two CGLS types give the oracle's bits, and the
    long run reaches the direct variant (TIKHONOV / RT_NO_Z) to 1e-9."""
    import lexls_b200 as lx
    rng = np.random.default_rng(77)
    n, dims, B = 13, [4, 3, 5], 9
    data = [np.ascontiguousarray(rng.standard_normal((B, n + 1, m))) for m in dims]
    factors = [0.3, 0.05, 0.7]

    def gpu(reg_type, it):
        with lx.LexLSEBatch(n, dims, batch=B) as s:
            s.setParameters(lx.Params(regularization_type=reg_type, max_number_of_cg_iterations=it))
            for k, f in enumerate(factors):
                s.setRegularizationFactor(k, f)
            for k, d in enumerate(data):
                s.setData(k, d)
            s.solve()
            return s.get_x()

    for cg, direct in ((2, 1), (6, 5)):
        for it in (1, 3, 40):
            oracle.set_regularization(cg, factors, cg_iterations=it)
            try:
                ref = oracle.lse_solve_batch(n, dims, data, threads=0)
            finally:
                oracle.set_regularization(0)
            assert np.array_equal(gpu(cg, it), ref["x"]), (cg, it)
        assert np.abs(gpu(cg, 400) - gpu(direct, 10)).max() < 1e-9


# ---- solveLeastNorm (lexlse.h:1046-1262) ----------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", ["w1", "w2", "lg8"])
def test_least_norm_solution_matches_the_oracle_and_pinv(oracle, monkeypatch, variant):
    import lexls_b200 as lx
    env = {"w1": {}, "w2": dict(WARPS=2), "lg8": dict(FORCE_LODGLOBAL=1, LG_WARPS=8)}[variant]
    for k, v in env.items():
        monkeypatch.setenv("LEXLS_B200_" + k, str(v))
    rng = np.random.default_rng(31)
    for n, dims, fixed in ((10, [3, 2, 2], False), (16, [4, 5, 3], True), (8, [3, 9], False)):
        B = 9
        data = [np.ascontiguousarray(rng.standard_normal((B, n + 1, m))) for m in dims]
        fi = np.tile(np.array([1, 4], dtype=np.uint32), (B, 1)) if fixed else None
        fv = rng.standard_normal((B, 2)) if fixed else None
        oracle.set_least_norm(True)
        try:
            ref = oracle.lse_solve_batch(n, dims, data, fixed_index=fi, fixed_value=fv, threads=0)
        finally:
            oracle.set_least_norm(False)
        basic = oracle.lse_solve_batch(n, dims, data, fixed_index=fi, fixed_value=fv, threads=0)
        with lx.LexLSEBatch(n, dims, batch=B) as s:
            s.request_outputs(least_norm=True)
            if fixed:
                s.fixVariables(fi, fv)
            for k, d in enumerate(data):
                s.setData(k, d)
            s.solve()
            x_ln, x = s.get_x_least_norm(), s.get_x()
        assert np.array_equal(x, basic["x"]) and np.array_equal(x_ln, ref["x"])
        assert (np.linalg.norm(x_ln, axis=1) <= np.linalg.norm(x, axis=1) + 1e-12).all()
        if not fixed and sum(dims) <= n:
            for p in range(B):  # consistent equalities: the least-norm solution is the pseudo-inverse solution
                A = np.concatenate([d[p].T[:, :n] for d in data])
                b = np.concatenate([d[p].T[:, n] for d in data])
                assert np.abs(x_ln[p] - np.linalg.pinv(A) @ b).max() < 1e-9


# ---- reflector application: one thread per column while more columns are open than thread pairs ------------------------------
@pytest.mark.parametrize("warps,n,dims", [(4, 80, [30, 25, 30]), (8, 150, [20, 30, 25]), (4, 70, [10, 70])])
def test_wide_shapes_on_the_multi_warp_kernels_match_the_oracle(oracle, monkeypatch, warps, n, dims):
    """Shapes with more open columns than half the threads of a group (qr_level_smem switches between one thread and a thread
    pair per column): bit-identical x, ranks and factorization, rank-deficient problems included."""
    monkeypatch.setenv("LEXLS_B200_WARPS", str(warps))
    monkeypatch.setenv("LEXLS_B200_NO_AUTO_LODGLOBAL", "1")
    rng = np.random.default_rng(warps * 1000 + n)
    B = 6
    data = [np.ascontiguousarray(rng.standard_normal((B, n + 1, m))) for m in dims]
    data[0][::2, 3, :] = data[0][::2, 1, :]  # duplicated column: rank deficiency and pivot ties in half of the problems
    ref = oracle.lse_solve_batch(n, dims, data, threads=0, want_factors=True)
    import lexls_b200 as lx
    with lx.LexLSEBatch(n, dims, batch=B) as s:
        s.request_outputs(lexqr=True)
        for k, d in enumerate(data):
            s.setData(k, d)
        s.solve()
        assert s.kernel_info()["name"] == "lse_w%d" % warps
        q, _ = s.get_lexqr()
        assert np.array_equal(s.get_x(), ref["x"]) and np.array_equal(s.get_ranks(), ref["ranks"])
        assert np.array_equal(q, ref["lexqr"])
