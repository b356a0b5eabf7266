"""GPU parity: batched LexLSE (factorize + solve + get_v) vs the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from helpers import RTOL, assert_close_batch, solve_lse_gpu

pytestmark = pytest.mark.gpu


def _check(oracle, nVar, ObjDim, data, fixed=None, tol=None):
    ref = oracle.lse_solve_batch(nVar, ObjDim, data, fixed_index=fixed[0] if fixed else None,
                                 fixed_value=fixed[1] if fixed else None, fixed_type=None,
                                 tol_linear_dependence=1e-12 if tol is None else tol, threads=0, want_factors=True)
    got = solve_lse_gpu(nVar, ObjDim, data, fixed=fixed, tol=tol)
    assert np.array_equal(got["ranks"], ref["ranks"]), "per-level ranks differ"
    assert np.array_equal(got["counters"][:, 5], ref["total_rank"]), "total rank differs"
    tr = ref["total_rank"]
    mask = np.arange(nVar)[None, :] < tr[:, None]
    assert np.array_equal(np.where(mask, got["colperm"], 0), np.where(mask, ref["colperm"], 0)), "pivot sequence differs"
    wx = assert_close_batch(got["x"], ref["x"], "x")
    off = 0
    for m in ObjDim:
        if m:
            ng = np.linalg.norm(got["v"][:, off:off + m], axis=1)
            nr = np.linalg.norm(ref["v"][:, off:off + m], axis=1)
            assert np.all(np.abs(ng - nr) <= RTOL * np.maximum(nr, 1e-9)), "||v_k|| differs"
        off += m
    exact = bool(np.array_equal(got["x"], ref["x"]) and np.array_equal(got["v"], ref["v"]))
    return wx, exact, got


def test_tiny_lse_bit_exact(oracle):
    from lexls_b200 import problems
    w = problems.workload("tiny_lse")
    data, _ = problems.generate(w)
    wx, exact, got = _check(oracle, w.nVar, w.ObjDim, data)
    assert exact, "the kernel follows the oracle's summation order; results should be bit-identical (worst rel %.3e)" % wx


@pytest.mark.parametrize("name,batch", [("c2", 2048), ("c2_60", 512)])
def test_c2_subset(oracle, name, batch):
    from lexls_b200 import problems
    w = problems.workload(name, batch)
    data, _ = problems.generate(w)
    wx, exact, got = _check(oracle, w.nVar, w.ObjDim, data)
    assert exact
    assert (got["status"] == 0).all()
    # algorithmic flop counter agrees with the closed formula
    fl = np.array([problems.lqr_flops(w.nVar, w.ObjDim, list(r)) for r in got["ranks"]])
    assert np.all(np.abs(got["counters"][:, 7].astype(np.int64) * 1000 - fl) <= 1000)


def test_fixed_variables(oracle):
    rng = np.random.default_rng(3)
    B, n, dims = 96, 14, [6, 5, 7]
    data = [np.ascontiguousarray(rng.standard_normal((B, n + 1, m))) for m in dims]
    nf = 4
    idx = np.stack([rng.permutation(n)[:nf] for _ in range(B)]).astype(np.uint32)
    val = rng.standard_normal((B, nf))
    wx, exact, _ = _check(oracle, n, dims, data, fixed=(idx, val))
    assert exact


@pytest.mark.parametrize("dims,n", [([3, 0, 4], 9), ([1], 5), ([12, 12], 6), ([40], 40), ([5, 5, 5, 5, 5, 5, 5, 5], 33), ([70, 50], 90)])
def test_edge_shapes(oracle, dims, n):
    """empty level, single row, more rows than columns, square, many levels, wider than one warp."""
    rng = np.random.default_rng(11)
    B = 40
    data = [np.ascontiguousarray(rng.standard_normal((B, n + 1, m))) for m in dims]
    # duplicate rows / zero rows to provoke rank decisions and pivot ties
    if dims[0] >= 3:
        data[0][: B // 2, :, 1] = data[0][: B // 2, :, 0]
        data[0][: B // 4, :n, 2] = 0.0
    wx, exact, _ = _check(oracle, n, dims, data)
    assert exact


@pytest.mark.parametrize("warps", ["2", "4"])
def test_multi_warp_groups_match(oracle, warps, monkeypatch):
    """Groups of several warps per problem must give the same bits as the warp-per-problem kernel."""
    from lexls_b200 import problems
    monkeypatch.setenv("LEXLS_B200_WARPS", warps)
    w = problems.workload("c2", 256)
    data, _ = problems.generate(w)
    wx, exact, got = _check(oracle, w.nVar, w.ObjDim, data)
    assert got["info"]["warps_per_problem"] == int(warps)
    assert exact


@pytest.mark.parametrize("lg_warps", ["8", "16"])
def test_lod_in_global_memory_variant(oracle, monkeypatch, lg_warps):
    monkeypatch.setenv("LEXLS_B200_FORCE_LODGLOBAL", "1")
    monkeypatch.setenv("LEXLS_B200_LG_WARPS", lg_warps)
    from lexls_b200 import problems
    w = problems.workload("tiny_lse")
    data, _ = problems.generate(w)
    wx, exact, got = _check(oracle, w.nVar, w.ObjDim, data)
    assert got["info"]["name"] == "lse_w%s_lodglobal" % lg_warps
    assert exact


def test_solution_is_lexicographic_least_squares():
    """Size-independent property at BASELINE's full C2 size: level-1 residual is the least-squares residual (numpy)."""
    from lexls_b200 import problems
    import lexls_b200 as lx
    w = problems.workload("c2")
    data, _ = problems.generate(w)
    with lx.LexLSEBatch(w.nVar, w.ObjDim, batch=w.batch) as s:
        for k, d in enumerate(data):
            s.setData(k, d)
        s.solve()
        x = s.get_x()
        ranks = s.get_ranks()
    assert ranks.shape == (w.batch, 3)
    # full-rank problems: levels 1 and 2 are satisfied exactly (20 + 20 = 40 equations, 40 unknowns)
    full = (ranks[:, 0] == 20) & (ranks[:, 1] == 20)
    assert full.sum() > 0.85 * w.batch
    for k in range(2):
        A = data[k][full][:, :-1, :].transpose(0, 2, 1)
        b = data[k][full][:, -1, :]
        r = np.einsum("bij,bj->bi", A, x[full]) - b
        assert np.abs(r).max() < 1e-7
    # sample: compare against numpy's lstsq cascade on a few rank-deficient problems
    idx = np.nonzero(~full)[0][:8]
    for p in idx:
        A1 = data[0][p][:-1, :].T
        b1 = data[0][p][-1, :]
        r_gpu = np.linalg.norm(A1 @ x[p] - b1)
        r_opt = np.linalg.norm(A1 @ np.linalg.lstsq(A1, b1, rcond=None)[0] - b1)
        assert abs(r_gpu - r_opt) <= 1e-8 * max(1.0, r_opt)


def test_get_lexqr_matches_the_oracle_factorization(oracle):
    """get_lexqr (lexlse.h:1626): R, the Householder vectors, the eliminated blocks and the scalars — bit for bit."""
    import lexls_b200 as lx
    from lexls_b200 import problems
    for name, batch in (("tiny_lse", None), ("c2", 128)):
        w = problems.workload(name, batch)
        data, _ = problems.generate(w)
        ref = oracle.lse_solve_batch(w.nVar, w.ObjDim, data, threads=0, want_factors=True)
        with lx.LexLSEBatch(w.nVar, w.ObjDim, batch=w.batch) as s:
            s.request_outputs(lexqr=True)
            for k, d in enumerate(data):
                s.setData(k, d)
            s.solve()
            q, hh = s.get_lexqr()
        assert np.array_equal(q, ref["lexqr"]), "factorized work matrix differs"
        assert np.array_equal(hh, ref["hh"]), "Householder scalars differ"


@pytest.mark.parametrize("dims,n", [([40, 7], 47), ([33, 20, 5], 30), ([4, 4], 3), ([17, 3, 29], 15), ([21, 21], 47), ([40, 40], 16),
                                    ([2, 38, 1], 31)])
def test_register_window_shapes(oracle, dims, n):
    """Shapes that exercise every instance of the register-window panel (1..3 column slots, 20- and 40-row windows, windows
    that end inside a 4-row group, more rows than columns) with rank deficiency, zero columns and exact pivot ties."""
    rng = np.random.default_rng(5)
    B = 48
    data = [np.ascontiguousarray(rng.standard_normal((B, n + 1, m))) for m in dims]
    data[0][: B // 3, 1, :] = data[0][: B // 3, 0, :]           # duplicated column -> equal norms -> first-maximum rule
    data[0][B // 3: B // 2, 2, :] = 0.0                          # zero column
    if len(dims) > 1 and dims[1] >= 2:
        data[1][: B // 2, :, 1] = 2.0 * data[1][: B // 2, :, 0]  # dependent rows -> rank found inside the level
    data[-1][B - 4:, :, :] = 0.0                                 # an all-zero level (rank 0)
    wx, exact, got = _check(oracle, n, dims, data)
    assert exact
    assert got["info"]["warps_per_problem"] == 1


def test_register_window_with_fixed_variables(oracle):
    rng = np.random.default_rng(9)
    B, n, dims = 32, 20, [9, 14, 6]
    data = [np.ascontiguousarray(rng.standard_normal((B, n + 1, m))) for m in dims]
    idx = np.tile(np.array([3, 0, 17, 8], dtype=np.uint32), (B, 1))
    val = rng.standard_normal((B, 4))
    wx, exact, _ = _check(oracle, n, dims, data, fixed=(idx, val))
    assert exact


def test_two_problems_per_warp_kernel(oracle, monkeypatch):
    """The experimental half-warp-per-problem kernel (LEXLS_B200_PAIR=1): odd batch (a last warp with one problem), pairs whose
    two problems leave a level at different steps (rank deficiency, zero levels) — same bits as the oracle."""
    from lexls_b200 import problems
    monkeypatch.setenv("LEXLS_B200_PAIR", "1")
    w = problems.workload("c2", 257)
    data, _ = problems.generate(w)
    wx, exact, got = _check(oracle, w.nVar, w.ObjDim, data)
    if got["info"]["name"] != "lse_pair":
        pytest.skip("the experiment is not compiled into the product library (build with -DLEXLS_B200_WITH_PAIR)")
    assert got["info"]["name"] == "lse_pair" and exact
    rng = np.random.default_rng(21)
    for dims, n in (([17, 3, 20], 15), ([20, 20], 47), ([9, 12, 5], 30)):
        B = 37
        data = [np.ascontiguousarray(rng.standard_normal((B, n + 1, m))) for m in dims]
        data[0][::2, 1, :] = data[0][::2, 0, :]            # every other problem: duplicated column (ties, lower rank)
        data[0][1::3, 2, :] = 0.0                           # zero column in every third
        if dims[1] >= 2:
            data[1][::4, :, 1] = -data[1][::4, :, 0]        # dependent rows: the level ends early for one half only
        data[-1][B - 3:, :, :] = 0.0
        wx, exact, got = _check(oracle, n, dims, data)
        assert got["info"]["name"] == "lse_pair" and exact
