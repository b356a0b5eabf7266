"""Synthetic workloads of BASELINE.json's configs, generated on the host with a counter-based RNG
(numpy Philox keyed by (seed, first problem index)) so that every rank / test regenerates identical data.

Level arrays use the batch layout of :mod:`lexls_b200.batch`: ``[B, cols, m_k]`` (each problem column-major).

Recipes lifted from the reference:
  * dense random levels ``A ~ N(0,1)``, ``b ~ N(0,1)``                       (interfaces/matlab-octave/examples/example_lexlse.m:8-13)
  * rank-deficient hierarchies ``A_k = randn(m_k, sum_{i<k} m_i + r_k) [C; randn(r_k, n)]`` with the ``max|C|`` rescale
    (interfaces/matlab-octave/tests/implementation/utility/define_problem.m:45-54)
  * inequality bounds ``ub ~ U(0,1)``, ``lb = ub - U(0,1)``                   (interfaces/matlab-octave/examples/example_lexlsi.m:8-14)
"""
from dataclasses import dataclass
from typing import List, Optional

import numpy as np


@dataclass
class Workload:
    name: str
    kind: str                 # "lse" | "lsi"
    nVar: int
    ObjDim: List[int]
    ObjType: List[int]        # 0 general, 1 simple bounds
    level_kind: List[str]     # per level: "eq" | "ineq" | "bounds" | "posture"
    batch: int
    seed: int
    rank_deficient_fraction: float = 0.0

    @property
    def nCtr(self):
        return sum(self.ObjDim)


def workload(name: str, batch: Optional[int] = None) -> Workload:
    """BASELINE.json configs[1..3] by name: 'c2', 'c2_60', 'c3', 'c4' (+ small variants for tests)."""
    if name == "c2":      # batched LexLSE 3 levels n=40 m=60 (20/20/20), 65,536 problems
        w = Workload("c2_lexlse_n40_m20x3", "lse", 40, [20, 20, 20], [0, 0, 0], ["eq"] * 3, 65536, 12345, 0.10)
    elif name == "c2_60":  # the other reading of "m=60": 60 rows per level
        w = Workload("c2_lexlse_n40_m60x3", "lse", 40, [60, 60, 60], [0, 0, 0], ["eq"] * 3, 65536, 12345, 0.10)
    elif name == "c3":    # batched LexLSI humanoid-IK shape, 5 levels n=38, warm start, 16,384 problems
        w = Workload("c3_lexlsi_ik_n38_5lvl", "lsi", 38, [38, 12, 6, 6, 38], [1, 0, 0, 0, 0],
                     ["bounds", "eq", "ineq", "eq", "posture"], 16384, 2025)
    elif name == "c4":    # LexLSI 10 levels n=200 m=400, 1,024 problems
        w = Workload("c4_lexlsi_n200_m40x10", "lsi", 200, [40] * 10, [0] * 10, ["eq", "ineq"] * 5, 1024, 777)
    elif name == "tiny_lse":
        w = Workload("tiny_lse", "lse", 12, [5, 4, 6], [0, 0, 0], ["eq"] * 3, 64, 7, 0.25)
    elif name == "tiny_lsi":
        w = Workload("tiny_lsi", "lsi", 10, [10, 4, 5, 10], [1, 0, 0, 0], ["bounds", "eq", "ineq", "posture"], 64, 11)
    else:
        raise KeyError(name)
    if batch is not None:
        w.batch = int(batch)
    return w


def _rng(seed, start):
    return np.random.Generator(np.random.Philox(key=[int(seed), int(start)]))


def generate(w: Workload, start: int = 0, count: Optional[int] = None):
    """Returns (data, var_index): data[k] is [count, cols, m_k]; var_index is [m_0] (shared) or None."""
    B = w.batch - start if count is None else count
    g = _rng(w.seed, start)
    n = w.nVar
    data = []
    var_index = None
    if w.kind == "lse":
        nrd = int(round(B * w.rank_deficient_fraction))
        A_prev = None  # stacked, rescaled previous levels for the rank-deficient sub-batch [nrd, rows, n]
        for k, m in enumerate(w.ObjDim):
            A = g.standard_normal((B, m, n))
            b = g.standard_normal((B, m, 1))
            if nrd > 0:
                r = max(m - 2, 1)
                prev_rows = 0 if A_prev is None else A_prev.shape[1]
                mix = g.standard_normal((nrd, m, prev_rows + r))
                fresh = g.standard_normal((nrd, r, n))
                C = fresh if A_prev is None else np.concatenate([A_prev, fresh], axis=1)
                Ard = mix @ C
                A[:nrd] = Ard
                stack = Ard if A_prev is None else np.concatenate([A_prev, Ard], axis=1)
                s = np.abs(stack).max(axis=(1, 2), keepdims=True)
                A_prev = stack / np.maximum(s, 1.0)
            data.append(np.ascontiguousarray(np.concatenate([A, b], axis=2).transpose(0, 2, 1)))
        return data, None
    for k, (m, kind) in enumerate(zip(w.ObjDim, w.level_kind)):
        if kind == "bounds":
            ub = 0.5 + g.random((B, m))
            lb = -0.5 - g.random((B, m))
            data.append(np.ascontiguousarray(np.stack([lb, ub], axis=1)))
            var_index = np.arange(m, dtype=np.uint32)
            continue
        if kind == "posture":
            A = np.broadcast_to(np.eye(m, n), (B, m, n)).copy()
            b = 0.1 * g.standard_normal((B, m))
            lb = ub = b
        else:
            A = g.standard_normal((B, m, n))
            if kind == "eq":
                lb = ub = g.standard_normal((B, m))
            else:
                ub = g.random((B, m))
                lb = ub - g.random((B, m))
        M = np.concatenate([A, lb[:, :, None], ub[:, :, None]], axis=2)
        data.append(np.ascontiguousarray(M.transpose(0, 2, 1)))
    return data, var_index


def perturb(w: Workload, data, scale=1e-2, start: int = 0):
    """Warm-start companion problem: shift b / (lb, ub) by scale*N(0,1), keeping ub - lb."""
    g = _rng(w.seed + 1000003, start)
    out = []
    for k, (m, kind) in enumerate(zip(w.ObjDim, w.level_kind)):
        d = data[k].copy()
        B = d.shape[0]
        shift = scale * g.standard_normal((B, m))
        if w.kind == "lse":
            d[:, -1, :] += shift
        else:
            d[:, -2, :] += shift
            d[:, -1, :] += shift
        out.append(d)
    return out


def lqr_flops(nVar, dims, ranks, nFixed=0):
    """Algorithmic flop count of one factorize()+solve() (BASELINE.md / SURVEY.md 8d), mul+add = 2."""
    F = 0
    col = nFixed
    P = len(dims)
    for k, (m, r) in enumerate(zip(dims, ranks)):
        c = nVar - col
        if c <= 0:
            break
        F += 2 * m * c
        for j in range(r):
            mm = m - j
            F += 2 * mm + ((3 * mm + 4 + 4 * mm * (c - j)) if mm > 1 else 0) + 2 * (c - j - 1)
        Mk = sum(dims[k + 1:])
        if k < P - 1 and r > 0:
            F += Mk * r * r + 2 * Mk * r * (c - r + 1)
        col += r
    F += sum(ranks) ** 2
    return F
