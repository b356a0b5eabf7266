"""Small solves of EVERY kernel family for compute-sanitizer (one tool per gpurun call: memcheck, racecheck, initcheck, synccheck):

    compute-sanitizer --tool memcheck python scripts/sanitize_small.py

Families (the plan() switches of lexls_b200.cu are read from the environment at solve time):
  LexLSE  lse_w1 (register-window panel), lse_w2 / w4 / w8 (shared-memory routines), lse_w8_lodglobal / lse_w16_lodglobal
          (work matrix in global memory; Schur update as DFMA strips and as DMMA tiles), with fixed variables and get_lexqr
  LexLSI  lsi_w1 (panel), lsi_w2 / w4 / w8 with the objective data in L2 and staged in shared memory by TMA, lsi_w8 / w16
          _lodglobal; cold + warm start, getLambda, get_lexqr, get_xStar, the working-set log, an api_activate call list,
          cycling handling, a factorization cap
Every result is checked against nothing here (parity is tests/); the point is the sanitizer's report."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lexls_b200 as lx
from lexls_b200 import problems

KNOBS = ("LEXLS_B200_WARPS", "LEXLS_B200_DATA", "LEXLS_B200_FORCE_LODGLOBAL", "LEXLS_B200_LG_WARPS", "LEXLS_B200_DMMA")


def env(**kw):
    for k in KNOBS:
        os.environ.pop(k, None)
    for k, v in kw.items():
        os.environ["LEXLS_B200_" + k] = str(v)


def lse(n, dims, batch, fixed=False, seed=0):
    rng = np.random.default_rng(seed)
    data = [np.ascontiguousarray(rng.standard_normal((batch, n + 1, m))) for m in dims]
    data[0][::3, 1, :] = data[0][::3, 0, :]  # rank deficiency in every third problem
    with lx.LexLSEBatch(n, dims, batch=batch) as s:
        s.request_outputs(lexqr=True)
        if fixed:
            s.fixVariables(np.tile(np.array([2, 0], dtype=np.uint32), (batch, 1)), rng.standard_normal((batch, 2)))
        for k, d in enumerate(data):
            s.setData(k, d)
        s.solve()
        s.get_lexqr()
        for k in range(len(dims)):
            s.get_v(k)
        return s.kernel_info()["name"], float(np.abs(s.get_x()).sum())


def lsi(name, batch, params=None, ordered=False):
    w = problems.workload(name, batch)
    data, vi = problems.generate(w)
    P = len(w.ObjDim)
    with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch) as s:
        s.setParameters(lx.Params(log_working_set_enabled=1, **(params or {})))
        s.request_outputs(multipliers=True, lexqr=True, xstar=True)
        for k, d in enumerate(data):
            s.setData(k, d, var_index=vi) if w.ObjType[k] == 1 else s.setData(k, d)
        s.solve()
        act = [s.getActiveCtr(k) for k in range(P)]
        x = s.get_x()
        s.getWorkingSetLog(capacity=32)
        s.get_xStar()
        for k in range(P):
            s.getLambda(k)
            s.get_v(k)
            s.getConstraintViolation(k)
            s.getActiveCtr_order(k)
        # warm start from the result, once per row type and once as a call list
        d2 = problems.perturb(w, data)
        for k, d in enumerate(d2):
            s.setData(k, d, var_index=vi) if w.ObjType[k] == 1 else s.setData(k, d)
        s.set_x0(x)
        if ordered:
            cnt = sum(w.ObjDim)
            o = np.zeros((w.batch, cnt), dtype=np.uint32)
            c = np.zeros((w.batch, cnt), dtype=np.uint32)
            t = np.zeros((w.batch, cnt), dtype=np.uint8)
            for p in range(w.batch):
                j = 0
                for k in reversed(range(P)):
                    for i in range(w.ObjDim[k]):
                        if act[k][p, i] in (1, 2):
                            o[p, j], c[p, j], t[p, j] = k, i, act[k][p, i]
                            j += 1
            s.set_active_guess_ordered(o, c, t)
        else:
            for k in range(P):
                s.set_active_guess(k, act[k])
        s.solve()
        return s.kernel_info()["name"], float(np.abs(s.get_x()).sum())


if __name__ == "__main__":
    runs = []
    env()
    runs.append(lse(12, [5, 4, 6], 16))
    runs.append(lse(40, [20, 20, 20], 12, fixed=True))
    for wv in (2, 4, 8):
        env(WARPS=wv)
        runs.append(lse(30, [9, 14, 6, 11], 6, fixed=(wv == 4), seed=wv))
    for lg in (8, 16):
        for dm in (0, 1):
            env(FORCE_LODGLOBAL=1, LG_WARPS=lg, DMMA=dm)
            runs.append(lse(37, [9, 14, 6, 11], 5, fixed=(dm == 1), seed=lg + dm))
    env()
    runs.append(lsi("tiny_lsi", 12))
    runs.append(lsi("c3", 8, ordered=True))
    runs.append(lsi("c3", 6, params=dict(max_number_of_factorizations=3)))
    runs.append(lsi("tiny_lsi", 8, params=dict(cycling_handling_enabled=1, cycling_max_counter=3)))
    env(WARPS=1)
    runs.append(lsi("c3", 6))
    for wv in (4, 8):
        env(WARPS=wv)
        runs.append(lsi("c3", 4, ordered=(wv == 8)))
    env(DATA="smem")
    runs.append(lsi("c3", 6))
    env(DATA="smem", WARPS=4)
    runs.append(lsi("tiny_lsi", 6))
    for lg in (8, 16):
        for dm in (0, 1):
            env(FORCE_LODGLOBAL=1, LG_WARPS=lg, DMMA=dm)
            runs.append(lsi("c3", 4, params=dict(max_number_of_factorizations=6)))
    env()
    w = problems.workload("c4", 2)
    data, vi = problems.generate(w)
    with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch) as s:
        s.setParameters(lx.Params(max_number_of_factorizations=4))
        for k, d in enumerate(data):
            s.setData(k, d)
        s.solve()
        runs.append((s.kernel_info()["name"], float(np.abs(s.get_x()).sum())))
    for name, val in runs:
        print("%-22s %.6e" % (name, val))
    print("families:", sorted(set(n for n, _ in runs)))
