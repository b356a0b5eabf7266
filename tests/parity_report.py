"""Full-size parity report: the CUDA path vs the CPU oracle on the complete BASELINE batches (C2: 65,536 LexLSE problems;
C3: 16,384 LexLSI problems cold and warm-started; with the argument `c4` also C4: 1,024 LexLSI problems of 10 levels, n = 200,
cold start up to the 200-factorization cap — minutes of oracle time).  Counts bit-identical results and discrete mismatches, and lists the
near-tie margins the oracle measures (pivot gap, rank margin, ratio-test gap, multiplier gap) — the cases a different
summation order (e.g. Eigen's) could decide differently.   usage: python tests/parity_report.py > profiles/rN_parity_report.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lexls_b200 as lx
import oracle
from lexls_b200 import problems

oracle.build()
out = {}

def quant(a):
    a = np.asarray(a, dtype=np.float64)
    a = a[np.isfinite(a) & (a < 1e299)]
    if a.size == 0:
        return None
    return {"min": float(a.min()), "p0.1": float(np.quantile(a, 0.001)), "p1": float(np.quantile(a, 0.01)), "median": float(np.median(a))}

# ---- C2 -------------------------------------------------------------------------------------------------------------
w = problems.workload("c2")
data, _ = problems.generate(w)
t0 = time.time()
ref = oracle.lse_solve_batch(w.nVar, w.ObjDim, data, threads=0, want_factors=True)
t_or = time.time() - t0
with lx.LexLSEBatch(w.nVar, w.ObjDim, batch=w.batch) as s:
    s.request_outputs(lexqr=True)
    for k, d in enumerate(data):
        s.setData(k, d)
    s.solve()
    x, ranks, cp = s.get_x(), s.get_ranks(), s.get_column_permutations()
    v = np.concatenate([s.get_v(k) for k in range(len(w.ObjDim))], axis=1)
    q, hh = s.get_lexqr()
    info = s.kernel_info()
tr = ref["total_rank"]
mask = np.arange(w.nVar)[None, :] < tr[:, None]
out["c2"] = {
    "workload": w.name, "problems": int(w.batch), "kernel": info["name"], "oracle_seconds": t_or,
    "x_bit_identical": int((x == ref["x"]).all(axis=1).sum()),
    "v_bit_identical": int((v == ref["v"]).all(axis=1).sum()),
    "factorization_bit_identical": int(((q == ref["lexqr"]).all(axis=(1, 2)) & (hh == ref["hh"]).all(axis=1)).sum()),
    "ranks_identical": int((ranks == ref["ranks"]).all(axis=1).sum()),
    "pivot_sequence_identical": int((np.where(mask, cp, 0) == np.where(mask, ref["colperm"], 0)).all(axis=1).sum()),
    "max_rel_err_x": float((np.linalg.norm(x - ref["x"], axis=1) / np.maximum(np.linalg.norm(ref["x"], axis=1), 1e-300)).max()),
    "problems_with_a_rank_deficient_level": int(((ranks[:, 0] < w.ObjDim[0]) | (ranks[:, 1] < w.ObjDim[1])).sum()),
}

# ---- C3 cold + warm -------------------------------------------------------------------------------------------------
w = problems.workload("c3")
data, vi = problems.generate(w)
vib = np.tile(vi, (w.batch, 1))
nl = len(w.ObjDim)

def run_gpu(data, guess=None, x0=None, w=w, vi=vi):
    nl = len(w.ObjDim)
    with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch) as s:
        for k, d in enumerate(data):
            s.setData(k, d, var_index=vi) if w.ObjType[k] == 1 else s.setData(k, d)
        if guess is not None:
            for k, g in enumerate(guess):
                s.set_active_guess(k, g)
            s.set_x0(x0)
        s.solve()
        return dict(x=s.get_x(), status=s.get_status(), counters=s.get_counters(), active=[s.getActiveCtr(k) for k in range(nl)],
                    v=[s.get_v(k) for k in range(nl)], info=s.kernel_info())

def compare(tag, got, ref, w=w):
    nl = len(w.ObjDim)
    act_same = np.ones(w.batch, dtype=bool)
    v_same = np.ones(w.batch, dtype=bool)
    for k in range(nl):
        act_same &= (got["active"][k] == ref["active"][k]).all(axis=1)
        v_same &= (got["v"][k] == ref["v"][k]).all(axis=1)
    m = ref["margins"]
    out[tag] = {
        "problems": int(w.batch), "kernel": got["info"]["name"],
        "status_identical": int((got["status"] == ref["status"]).sum()),
        "counters_identical(fact,act,deact,iter)": int((got["counters"][:, :4] == ref["counters"][:, :4]).all(axis=1).sum()),
        "active_set_identical": int(act_same.sum()),
        "x_bit_identical": int((got["x"] == ref["x"]).all(axis=1).sum()),
        "v_bit_identical": int(v_same.sum()),
        "max_rel_err_x": float((np.linalg.norm(got["x"] - ref["x"], axis=1) / np.maximum(np.linalg.norm(ref["x"], axis=1), 1e-300)).max()),
        "mean_factorizations": float(got["counters"][:, 0].mean()),
        "status_counts": {str(int(k)): int((got["status"] == k).sum()) for k in np.unique(got["status"])},
        "near_tie_margins": {"pivot_gap_rel": quant(m[:, 0]), "rank_margin_log10": quant(m[:, 1]), "ratio_test_gap": quant(m[:, 2]),
                             "multiplier_gap": quant(m[:, 3])},
    }

ref = oracle.lsi_solve_batch(w.nVar, w.ObjDim, w.ObjType, data, var_index=vib, threads=0)
got = run_gpu(data)
compare("c3_cold", got, ref)
d2 = problems.perturb(w, data)
ref2 = oracle.lsi_solve_batch(w.nVar, w.ObjDim, w.ObjType, d2, var_index=vib, guess=got["active"], x0=got["x"], threads=0)
got2 = run_gpu(d2, guess=got["active"], x0=got["x"])
compare("c3_warm", got2, ref2)

# ---- C4 (optional: the oracle needs minutes) --------------------------------------------------------------------------
if "c4" in sys.argv[1:]:
    w4 = problems.workload("c4")
    d4, v4 = problems.generate(w4)
    t0 = time.time()
    ref4 = oracle.lsi_solve_batch(w4.nVar, w4.ObjDim, w4.ObjType, d4, var_index=None, threads=0)
    t_or = time.time() - t0
    got4 = run_gpu(d4, w=w4, vi=v4)
    compare("c4_cold", got4, ref4, w=w4)
    out["c4_cold"]["oracle_seconds"] = t_or
print(json.dumps(out, indent=1))
