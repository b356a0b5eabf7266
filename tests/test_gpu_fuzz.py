"""GPU parity on randomly drawn SHAPES (seeded): every kernel family is crossed — the register-window panel (<= 48 columns,
<= 40 rows per level), the two-warp and the 4 / 8-warp shared-memory routines, levels of one row, empty levels, more rows
than columns, rank-deficient and duplicated rows, with and without a simple-bounds level.  Discrete outcomes (ranks, pivot
sequence, status, counters, active sets, working-set order) must be identical to the oracle's; x and ||v_k|| within 1e-9."""
import numpy as np
import pytest

from helpers import compare_lsi
from test_gpu_lse import _check
from test_gpu_lsi import _both

pytestmark = pytest.mark.gpu


def _lse_case(seed, wide=False):
    rng = np.random.default_rng(1000 + seed + (500 if wide else 0))
    n = int(rng.integers(100, 400)) if wide else int(rng.integers(1, 100))
    P = int(rng.integers(1, 5 if wide else 7))
    dims = [int(rng.integers(0, 90 if wide else 46)) if rng.random() < 0.8 else int(rng.integers(0, 3)) for _ in range(P)]
    if sum(dims) == 0:
        dims[0] = 2
    B = 8 if wide else 24
    data = []
    for m in dims:
        d = rng.standard_normal((B, n + 1, m))
        if m >= 2 and rng.random() < 0.5:
            d[: B // 2, :, m - 1] = d[: B // 2, :, 0]                      # duplicated row
        if m >= 3 and rng.random() < 0.3:
            d[: B // 3, :n, 1] = 0.0                                       # zero normal
        if m >= 4 and rng.random() < 0.3:
            d[:, :, 2] = 0.5 * d[:, :, 0] - 2.0 * d[:, :, 3]               # dependent row
        data.append(np.ascontiguousarray(d))
    nfix = int(rng.integers(0, min(n, 4))) if rng.random() < 0.4 else 0
    fixed = None
    if nfix:
        idx = np.stack([rng.permutation(n)[:nfix] for _ in range(B)]).astype(np.uint32)
        fixed = (idx, rng.standard_normal((B, nfix)))
    return n, dims, data, fixed


@pytest.mark.parametrize("seed", range(24))
def test_random_lse_shapes(oracle, seed):
    n, dims, data, fixed = _lse_case(seed)
    tol = [None, 1e-9, 1e-6, 1e-3][seed % 4]  # tol_linear_dependence decides the ranks of the dependent rows
    wx, exact, _ = _check(oracle, n, dims, data, fixed=fixed, tol=tol)
    assert exact, "n=%d dims=%s fixed=%s tol=%s: not bit-identical (worst rel err %.3e)" % (n, dims, fixed is not None, tol, wx)


def _lsi_case(seed, wide=False):
    rng = np.random.default_rng(5000 + seed + (500 if wide else 0))
    n = int(rng.integers(65, 260)) if wide else int(rng.integers(2, 80))
    P = int(rng.integers(1, 5 if wide else 6))
    B = 6 if wide else 16
    sb = rng.random() < 0.5

    class W:
        pass
    W.nVar, W.ObjDim, W.ObjType = n, [], []
    data, vi = [], None
    for k in range(P):
        if k == 0 and sb:
            m = int(rng.integers(1, n + 1))
            vi = np.sort(rng.permutation(n)[:m]).astype(np.uint32)
            ub = 0.3 + rng.random((B, m))
            lb = -0.3 - rng.random((B, m))
            if rng.random() < 0.5:
                lb[:, 0] = ub[:, 0]                                        # a fixed variable (equality bound)
            data.append(np.ascontiguousarray(np.stack([lb, ub], axis=1)))
            W.ObjDim.append(m)
            W.ObjType.append(1)
            continue
        m = int(rng.integers(1, 60 if wide else 31))
        A = rng.standard_normal((B, m, n))
        ub = rng.standard_normal((B, m)) + 0.5
        lb = ub - rng.random((B, m)) * 2.0
        eq = rng.random(m) < rng.random()
        lb[:, eq] = ub[:, eq]
        if m >= 2 and rng.random() < 0.4:
            A[:, m - 1, :] = A[:, 0, :]                                    # parallel constraints
        if m >= 3 and rng.random() < 0.2:
            A[:, 1, :] = 0.0                                               # 0 * x in [lb, ub]
        data.append(np.ascontiguousarray(np.concatenate([A, lb[:, :, None], ub[:, :, None]], axis=2).transpose(0, 2, 1)))
        W.ObjDim.append(m)
        W.ObjType.append(0)
    return W, data, vi


@pytest.mark.parametrize("seed", range(24))
def test_random_lsi_shapes(oracle, seed):
    W, data, vi = _lsi_case(seed)
    got, ref = _both(oracle, W, data, vi, max_number_of_factorizations=60)
    compare_lsi(got, ref, len(W.ObjDim))
    assert np.array_equal(got["x"], ref["x"]), "n=%d dims=%s types=%s: x not bit-identical" % (W.nVar, W.ObjDim, W.ObjType)
    # warm start from the result (previous active set and x as the guess)
    got2, ref2 = _both(oracle, W, data, vi, guess=got["active"], x0=got["x"], max_number_of_factorizations=60)
    compare_lsi(got2, ref2, len(W.ObjDim))
    assert np.array_equal(got2["x"], ref2["x"])


@pytest.mark.parametrize("seed", range(6))
def test_random_wide_lse_shapes(oracle, seed):
    """100-400 columns, up to 90 rows per level: the 4 / 8-warp shared-memory kernels and the global-work-matrix kernel."""
    n, dims, data, fixed = _lse_case(seed, wide=True)
    wx, exact, got = _check(oracle, n, dims, data, fixed=fixed)
    assert exact, "n=%d dims=%s (%s): not bit-identical (worst rel err %.3e)" % (n, dims, got["info"]["name"], wx)


@pytest.mark.parametrize("seed", range(6))
def test_random_wide_lsi_shapes(oracle, seed):
    W, data, vi = _lsi_case(seed, wide=True)
    got, ref = _both(oracle, W, data, vi, max_number_of_factorizations=25)
    compare_lsi(got, ref, len(W.ObjDim))
    assert np.array_equal(got["x"], ref["x"]), "n=%d dims=%s (%s): x not bit-identical" % (W.nVar, W.ObjDim, got["info"]["name"])


@pytest.mark.parametrize("lg_warps", ["8", "16"])
@pytest.mark.parametrize("seed", range(5))
def test_random_shapes_with_the_work_matrix_in_global_memory(oracle, monkeypatch, seed, lg_warps):
    """The same random shapes forced through the global-work-matrix kernels (level view in shared memory, odd first rows,
    deferred column swaps of the rows outside the level), 8 and 16 warps per problem."""
    monkeypatch.setenv("LEXLS_B200_FORCE_LODGLOBAL", "1")
    monkeypatch.setenv("LEXLS_B200_LG_WARPS", lg_warps)
    n, dims, data, fixed = _lse_case(seed)
    wx, exact, got = _check(oracle, n, dims, data, fixed=fixed)
    assert got["info"]["name"] == "lse_w%s_lodglobal" % lg_warps
    assert exact, "LexLSE n=%d dims=%s: not bit-identical (worst rel err %.3e)" % (n, dims, wx)
    W, data, vi = _lsi_case(seed + 1)
    got, ref = _both(oracle, W, data, vi, max_number_of_factorizations=40)
    assert got["info"]["name"] == "lsi_w%s_lodglobal" % lg_warps
    compare_lsi(got, ref, len(W.ObjDim))
    assert np.array_equal(got["x"], ref["x"]), "LexLSI n=%d dims=%s types=%s: x not bit-identical" % (W.nVar, W.ObjDim, W.ObjType)


@pytest.mark.parametrize("seed", range(16))
def test_random_parameters_and_guesses(oracle, seed):
    """Random ParametersLexLSI (tolerances, cycling guard, iteration cap, hot-start flags) with random active-set guesses, x0 and
    (sometimes) v0 on random shapes: every option of the path against the oracle."""
    rng = np.random.default_rng(9000 + seed)
    W, data, vi = _lsi_case(100 + seed)
    B = data[0].shape[0]
    par = dict(max_number_of_factorizations=int(rng.choice([1, 2, 5, 40])),
               tol_linear_dependence=float(rng.choice([1e-12, 1e-9, 1e-6])),
               tol_wrong_sign_lambda=float(rng.choice([1e-8, 1e-6, 1e-10])),
               tol_correct_sign_lambda=float(rng.choice([1e-12, 1e-9, 0.0])),
               tol_feasibility=float(rng.choice([1e-13, 1e-10, 1e-8])),
               cycling_handling_enabled=int(rng.integers(0, 2)), cycling_max_counter=int(rng.integers(1, 6)),
               cycling_relax_step=float(rng.choice([1e-8, 1e-6])),
               modify_x_guess_enabled=int(rng.integers(0, 2)), modify_type_active_enabled=int(rng.integers(0, 2)),
               modify_type_inactive_enabled=int(rng.integers(0, 2)), set_min_init_ctr_violation=int(rng.integers(0, 2)))
    guess = x0 = v0 = None
    mode = int(rng.integers(0, 4))
    if mode >= 1:
        guess = [rng.choice([0, 0, 1, 2], size=(B, m)).astype(np.uint8) for m in W.ObjDim]
    if mode >= 2:
        x0 = 0.5 * rng.standard_normal((B, W.nVar))
        if rng.random() < 0.5:
            par["use_phase1_v0"] = 1
    if mode == 3:
        v0 = [0.1 * rng.standard_normal((B, m)) for m in W.ObjDim]
    got, ref = _both(oracle, W, data, vi, guess=guess, x0=x0, v0=v0, **par)
    compare_lsi(got, ref, len(W.ObjDim))
    assert np.array_equal(got["x"], ref["x"]), "n=%d dims=%s types=%s par=%s" % (W.nVar, W.ObjDim, W.ObjType, par)


@pytest.mark.parametrize("lg_warps", ["8", "16"])
def test_global_work_matrix_with_more_row_pairs_than_threads(oracle, monkeypatch, lg_warps):
    """1,100 rows = 550 row pairs per column: more pairs outside a level than the group has threads (256 / 512), so the column
    swaps of the rows outside the level run both the pipelined path (one pair per thread) and the immediate loop."""
    monkeypatch.setenv("LEXLS_B200_FORCE_LODGLOBAL", "1")
    monkeypatch.setenv("LEXLS_B200_LG_WARPS", lg_warps)
    rng = np.random.default_rng(77)
    n, dims, B = 60, [20, 20, 20] + [80] * 13, 3   # 1,100 rows; the first three levels eliminate all 60 columns
    data = [np.ascontiguousarray(rng.standard_normal((B, n + 1, m))) for m in dims]
    data[1][:, :, 5] = data[1][:, :, 4]      # a dependent row: level 3 gets a column to eliminate as well
    wx, exact, got = _check(oracle, n, dims, data)
    assert got["info"]["name"] == "lse_w%s_lodglobal" % lg_warps
    assert exact, "not bit-identical (worst rel err %.3e)" % wx
