"""GPU parity: batched LexLSI (device-resident active-set loop) vs the CPU oracle on the same seeded inputs."""
import os

import numpy as np
import pytest

from helpers import compare_lsi, solve_lsi_gpu

pytestmark = pytest.mark.gpu


def _both(oracle, w, data, vi, guess=None, x0=None, v0=None, **par):
    import lexls_b200 as lx
    B = data[0].shape[0]
    vib = np.tile(vi, (B, 1)) if (vi is not None and vi.ndim == 1) else vi
    ref = oracle.lsi_solve_batch(w.nVar, w.ObjDim, w.ObjType, data, var_index=vib, guess=guess, x0=x0, v0=v0,
                                 params=oracle.OracleParams(**par) if par else None, threads=0)
    got = solve_lsi_gpu(w.nVar, w.ObjDim, w.ObjType, data, var_index=vi, guess=guess, x0=x0, v0=v0,
                        params=lx.Params(**par) if par else None)
    return got, ref


def test_tiny_lsi_cold(oracle):
    from lexls_b200 import problems
    w = problems.workload("tiny_lsi")
    data, vi = problems.generate(w)
    got, ref = _both(oracle, w, data, vi)
    compare_lsi(got, ref, len(w.ObjDim))
    assert np.array_equal(got["x"], ref["x"]), "bit-exact parity expected"
    assert (got["status"] == 0).all()


def test_c3_cold_and_warm(oracle):
    from lexls_b200 import problems
    w = problems.workload("c3", 1024)
    data, vi = problems.generate(w)
    got, ref = _both(oracle, w, data, vi)
    compare_lsi(got, ref, len(w.ObjDim))
    assert np.array_equal(got["x"], ref["x"])
    assert got["counters"][:, 0].max() > 3  # the active-set loop really iterated
    # warm start: perturbed problem, previous active set and x as the guess
    d2 = problems.perturb(w, data)
    got2, ref2 = _both(oracle, w, d2, vi, guess=got["active"], x0=got["x"])
    compare_lsi(got2, ref2, len(w.ObjDim))
    assert np.array_equal(got2["x"], ref2["x"])
    assert got2["counters"][:, 0].mean() < got["counters"][:, 0].mean()


@pytest.mark.parametrize("lg_warps", ["8", "16"])
def test_c4_shape_work_matrix_in_global_memory(oracle, monkeypatch, lg_warps):
    """BASELINE configs[3] shape (10 levels, n = 200, 40 rows each): the work matrix (401 x 201 doubles) exceeds one SM's
    shared memory, so the kernel keeps it in an L2-resident scratch buffer (8 warps per problem and two problems per SM by
    default, 16 warps when two level views do not fit).  Bounded by the factorization cap."""
    monkeypatch.setenv("LEXLS_B200_LG_WARPS", lg_warps)
    from lexls_b200 import problems
    w = problems.workload("c4", 6)
    data, vi = problems.generate(w)
    got, ref = _both(oracle, w, data, vi, max_number_of_factorizations=10)
    assert got["info"]["name"] == "lsi_w%s_lodglobal" % lg_warps
    compare_lsi(got, ref, len(w.ObjDim))
    assert np.array_equal(got["x"], ref["x"])
    assert (got["status"] == 2).all() and (got["counters"][:, 0] == 10).all()


def test_test01_dat_golden(oracle, golden_dir):
    """The reference's own fixture: x must match the file's #Solution (cold and with the file's active-set guess)."""
    from lexls_b200 import datfile

    class W:
        pass
    h = datfile.read_hierarchy(os.path.join(golden_dir, "test_01.dat"))
    w = W()
    w.nVar, w.ObjDim, w.ObjType = h.nVar, h.ObjDim, h.ObjType
    data = [np.ascontiguousarray(M.T)[None] for M in h.objectives]
    for guess in (None, [g[None] for g in h.active_set_guess]):
        got, ref = _both(oracle, w, data, h.var_index, guess=guess)
        compare_lsi(got, ref, len(w.ObjDim))
        rel = np.linalg.norm(got["x"][0] - h.solution) / np.linalg.norm(h.solution)
        assert rel < 1e-9, rel
        assert got["status"][0] == 0


@pytest.mark.parametrize("flags", [dict(modify_type_active_enabled=1, modify_type_inactive_enabled=1),
                                   dict(modify_x_guess_enabled=1, modify_type_inactive_enabled=1),
                                   dict(set_min_init_ctr_violation=0), dict(use_phase1_v0=1)])
def test_hot_start_flags(oracle, flags):
    from lexls_b200 import problems
    w = problems.workload("tiny_lsi", 96)
    data, vi = problems.generate(w)
    rng = np.random.default_rng(5)
    x0 = 0.3 * rng.standard_normal((96, w.nVar))
    guess = [rng.integers(0, 3, size=(96, m)).astype(np.uint8) for m in w.ObjDim]
    got, ref = _both(oracle, w, data, vi, guess=guess, x0=x0, **flags)
    compare_lsi(got, ref, len(w.ObjDim))


def test_x0_and_v0(oracle):
    from lexls_b200 import problems
    w = problems.workload("tiny_lsi", 64)
    data, vi = problems.generate(w)
    got, ref = _both(oracle, w, data, vi)
    # restart from the solution with its own residuals: must terminate after one factorization
    got2, ref2 = _both(oracle, w, data, vi, guess=got["active"], x0=got["x"], v0=got["v"])
    compare_lsi(got2, ref2, len(w.ObjDim))
    # partially specified v0 is disregarded (lexlsi.h:758-793)
    v0p = list(got["v"])
    v0p[1] = None
    got3, ref3 = _both(oracle, w, data, vi, x0=got["x"], v0=v0p)
    compare_lsi(got3, ref3, len(w.ObjDim))


def test_cycling_handling_and_iteration_cap(oracle):
    from lexls_b200 import problems
    w = problems.workload("tiny_lsi", 128)
    data, vi = problems.generate(w)
    got, ref = _both(oracle, w, data, vi, cycling_handling_enabled=1, cycling_max_counter=3, cycling_relax_step=1e-8)
    compare_lsi(got, ref, len(w.ObjDim))
    got, ref = _both(oracle, w, data, vi, max_number_of_factorizations=2)
    compare_lsi(got, ref, len(w.ObjDim))
    assert (got["status"] == 2).any()


def test_no_simple_bounds_and_degenerate_rows(oracle):
    """general levels only; equality rows with zero normal (not activated, lexlsi.h:377); duplicated inequality rows (ties)."""
    from lexls_b200 import problems
    rng = np.random.default_rng(9)
    B, n = 80, 9

    class W:
        nVar, ObjDim, ObjType = n, [4, 6, 9], [0, 0, 0]
    data = []
    for m in W.ObjDim:
        A = rng.standard_normal((B, m, n))
        ub = rng.random((B, m))
        lb = ub - rng.random((B, m))
        data.append(np.concatenate([A, lb[:, :, None], ub[:, :, None]], axis=2))
    data[0][:, 0, :n] = 0.0
    data[0][:, 0, n] = data[0][:, 0, n + 1] = 0.25        # 0*x = 0.25 : equality with zero normal
    data[1][:, 1, :] = data[1][:, 0, :]                    # duplicated inequality -> exact ratio ties
    data[2][:, :, n] = data[2][:, :, n + 1]                # last level: equalities
    data = [np.ascontiguousarray(d.transpose(0, 2, 1)) for d in data]
    got, ref = _both(oracle, W, data, None)
    compare_lsi(got, ref, 3)


@pytest.mark.parametrize("env", [{"LEXLS_B200_WARPS": "1"}, {"LEXLS_B200_WARPS": "1", "LEXLS_B200_DATA": "smem"},
                                 {"LEXLS_B200_WARPS": "4"}, {"LEXLS_B200_DATA": "smem"}, {"LEXLS_B200_FORCE_LODGLOBAL": "1"}])
def test_kernel_variants_match(oracle, env, monkeypatch):
    from lexls_b200 import problems
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    w = problems.workload("c3", 192)
    data, vi = problems.generate(w)
    got, ref = _both(oracle, w, data, vi)
    compare_lsi(got, ref, len(w.ObjDim))
    assert np.array_equal(got["x"], ref["x"])


def test_invalid_inputs_are_rejected():
    import lexls_b200 as lx
    with lx.LexLSIBatch(4, [3, 2], [1, 0], batch=2) as s:
        bad = np.zeros((2, 2, 3))
        bad[:, 0, :] = 1.0  # lb > ub
        with pytest.raises(lx.LexLSError) as e:
            s.setData(0, bad, var_index=np.array([0, 1, 2], dtype=np.uint32))
        assert e.value.code == 4
        ok = np.zeros((2, 2, 3))
        ok[:, 1, :] = 1.0
        with pytest.raises(lx.LexLSError):
            s.setData(0, ok, var_index=np.array([0, 1, 1], dtype=np.uint32))  # duplicate indices, lexlsi.h:485
        with pytest.raises(lx.LexLSError):
            s.solve()  # setData missing
        with pytest.raises(lx.LexLSError):
            s.setParameters(lx.Params(regularization_type=10))  # typedefs.h:32-44 ends at REGULARIZATION_TEST = 9


@pytest.mark.parametrize("name,count,cap", [("tiny_lsi", 12, None), ("c3", 6, None), ("c3", 4, 3)])
def test_get_lambda_matches_the_oracle(oracle, name, count, cap):
    """getLambda (lexlsi.h:552-605): per-level multiplier matrices, also for problems stopped by the factorization cap
    (the reference re-factorizes the final working set first)."""
    import lexls_b200 as lx
    from lexls_b200 import problems
    w = problems.workload(name, count)
    data, vi = problems.generate(w)
    par = dict(max_number_of_factorizations=cap) if cap else {}
    with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch) as s:
        if par:
            s.setParameters(lx.Params(**par))
        s.request_outputs(multipliers=True, lexqr=True)
        for k, d in enumerate(data):
            s.setData(k, d, var_index=vi) if w.ObjType[k] == 1 else s.setData(k, d)
        s.solve()
        lam = [s.getLambda(k) for k in range(len(w.ObjDim))]
        status = s.get_status()
        x = s.get_x()
    nz = 0
    for p in range(w.batch):
        one = [d[p] for d in data]
        r = oracle.lsi_solve_detail(w.nVar, w.ObjDim, w.ObjType, one, var_index=vi,
                                    params=oracle.OracleParams(**par) if par else None)
        assert r["status"] == status[p]
        assert np.array_equal(r["x"], x[p])
        for k in range(len(w.ObjDim)):
            # oracle: [m_k, nObj]; device: [nObj, m_k]
            assert np.array_equal(lam[k][p].T, r["lambdas"][k]), (p, k)
            nz += int(np.count_nonzero(lam[k][p]))
    assert nz > 0
    if cap:
        assert (status == 2).any()


def test_full_size_c3_idempotence_and_feasibility():
    """Size-independent properties at BASELINE's full C3 size (16,384 problems, no oracle): every problem ends
    PROBLEM_SOLVED; the top-priority simple bounds hold; re-solving from the solution (x0 = x, guess = final active set)
    takes exactly one factorization and reproduces x and the active set."""
    import lexls_b200 as lx
    from lexls_b200 import problems
    w = problems.workload("c3")
    data, vi = problems.generate(w)
    nl = len(w.ObjDim)
    with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch) as s:
        for k, d in enumerate(data):
            s.setData(k, d, var_index=vi) if w.ObjType[k] == 1 else s.setData(k, d)
        s.solve()
        x, st, cnt = s.get_x(), s.get_status(), s.get_counters()
        act = [s.getActiveCtr(k) for k in range(nl)]
        viol0 = s.getConstraintViolation(0)
        assert (st == 0).all()
        assert np.abs(viol0).max() <= 1e-9
        lb, ub = data[0][:, 0, :], data[0][:, 1, :]
        assert (x[:, vi] >= lb - 1e-9).all() and (x[:, vi] <= ub + 1e-9).all()
        assert cnt[:, 0].mean() > 3
        for k in range(nl):
            s.set_active_guess(k, act[k])
        s.set_x0(x)
        s.solve()
        x2, st2, cnt2 = s.get_x(), s.get_status(), s.get_counters()
        act2 = [s.getActiveCtr(k) for k in range(nl)]
    assert (st2 == 0).all()
    assert (cnt2[:, 0] == 1).all(), "a converged active set must verify in one factorization"
    for k in range(nl):
        assert np.array_equal(act[k], act2[k])
    num = np.linalg.norm(x2 - x, axis=1)
    den = np.maximum(np.linalg.norm(x, axis=1), 1e-12)
    assert (num / den).max() <= 1e-9


@pytest.mark.parametrize("name,count,par", [("tiny_lsi", 10, {}), ("c3", 6, {}),
                                             ("c3", 4, dict(cycling_handling_enabled=1, cycling_max_counter=3)),
                                             ("c3", 3, dict(max_number_of_factorizations=4))])
def test_working_set_log_matches_the_oracle(oracle, name, count, par):
    """getWorkingSetLog (lexlsi.h:739): the sequence of added / removed constraints with step lengths, multipliers, ranks."""
    import lexls_b200 as lx
    from lexls_b200 import problems
    w = problems.workload(name, count)
    data, vi = problems.generate(w)
    par = dict(par, log_working_set_enabled=1)
    with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch) as s:
        s.setParameters(lx.Params(**par))
        for k, d in enumerate(data):
            s.setData(k, d, var_index=vi) if w.ObjType[k] == 1 else s.setData(k, d)
        s.solve()
        entries, n = s.getWorkingSetLog(capacity=256)
        status = s.get_status()
    total = 0
    for p in range(w.batch):
        r = oracle.lsi_solve_detail(w.nVar, w.ObjDim, w.ObjType, [d[p] for d in data], var_index=vi,
                                    params=oracle.OracleParams(**par))
        assert r["status"] == status[p]
        assert n[p] == len(r["log"]), (p, n[p], len(r["log"]))
        assert np.array_equal(entries[p, :n[p]], r["log"]), p
        total += int(n[p])
    assert total > 0


def test_pipelined_batch_matches_plain_calls():
    """PipelinedBatch (chunks with overlapped copies) returns the same bits as one plain batch: LexLSE and warm-started LexLSI."""
    import lexls_b200 as lx
    from lexls_b200 import problems
    w = problems.workload("c2", 1000)
    data, _ = problems.generate(w)
    with lx.LexLSEBatch(w.nVar, w.ObjDim, batch=w.batch) as s:
        for k, d in enumerate(data):
            s.setData(k, d)
        s.solve()
        x_ref = s.get_x()
    x = np.empty_like(x_ref)
    with lx.PipelinedBatch(lx.LexLSEBatch, w.nVar, w.ObjDim, batch=w.batch, chunks=3) as pb:
        pb.solve_host(data, x)
        assert np.array_equal(pb.gather("get_ranks").shape, (w.batch, 3))
    assert np.array_equal(x, x_ref)

    w = problems.workload("c3", 500)
    data, vi = problems.generate(w)
    nl = len(w.ObjDim)
    with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch) as s:
        for k, d in enumerate(data):
            s.setData(k, d, var_index=vi) if w.ObjType[k] == 1 else s.setData(k, d)
        s.solve()
        guess, x0 = [s.getActiveCtr(k) for k in range(nl)], s.get_x()
        d2 = problems.perturb(w, data)
        for k, d in enumerate(d2):
            s.setData(k, d, var_index=vi) if w.ObjType[k] == 1 else s.setData(k, d)
        for k, g in enumerate(guess):
            s.set_active_guess(k, g)
        s.set_x0(x0)
        s.solve()
        x_ref, st_ref = s.get_x(), s.get_status()
    x = np.empty_like(x_ref)
    with lx.PipelinedBatch(lx.LexLSIBatch, w.nVar, w.ObjDim, w.ObjType, batch=w.batch, chunks=4) as pb:
        pb.solve_host(d2, x, var_index=vi, guess=guess, x0=x0)
        st = pb.gather("get_status")
    assert np.array_equal(x, x_ref) and np.array_equal(st, st_ref)
