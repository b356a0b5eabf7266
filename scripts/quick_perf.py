"""Developer probe (not the bench contract): kernel time and achieved FP64 rate for C2 / C3 with the variant knobs."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lexls_b200 as lx
from lexls_b200 import problems

def run_lse(name, batch, reps=5):
    w = problems.workload(name, batch)
    data, _ = problems.generate(w)
    with lx.LexLSEBatch(w.nVar, w.ObjDim, batch=w.batch) as s:
        for k, d in enumerate(data):
            s.setData(k, d)
        ms = []
        for _ in range(reps):
            s.solve(); ms.append(s.last_solve_ms())
        c = s.get_counters(); info = s.kernel_info()
    fl = c[:, 7].astype(np.float64).sum() * 1e3
    best = min(ms[1:])
    return dict(workload=w.name, batch=w.batch, ms=best, problems_per_s=w.batch / best * 1e3, tflops=fl / best * 1e-9, info=info)

def run_lsi(name, batch, reps=4, warm=True):
    w = problems.workload(name, batch)
    data, vi = problems.generate(w)
    with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch) as s:
        for k, d in enumerate(data):
            s.setData(k, d, var_index=vi) if w.ObjType[k] == 1 else s.setData(k, d)
        ms = []
        for _ in range(reps):
            s.solve(); ms.append(s.last_solve_ms())
        c = s.get_counters(); info = s.kernel_info()
        out = dict(workload=w.name + "_cold", batch=w.batch, ms=min(ms[1:]), nfact=float(c[:, 0].mean()),
                   tflops=c[:, 7].astype(np.float64).sum() * 1e3 / min(ms[1:]) * 1e-9, info=info)
        out["problems_per_s"] = w.batch / out["ms"] * 1e3
        res = [out]
        if warm:
            act = [s.getActiveCtr(k) for k in range(len(w.ObjDim))]
            x = s.get_x()
            d2 = problems.perturb(w, data)
            for k, d in enumerate(d2):
                s.setData(k, d, var_index=vi) if w.ObjType[k] == 1 else s.setData(k, d)
            for k, a in enumerate(act):
                s.set_active_guess(k, a)
            s.set_x0(x)
            ms = []
            for _ in range(reps):
                s.solve(); ms.append(s.last_solve_ms())
            c = s.get_counters()
            o2 = dict(workload=w.name + "_warm", batch=w.batch, ms=min(ms[1:]), nfact=float(c[:, 0].mean()),
                      tflops=c[:, 7].astype(np.float64).sum() * 1e3 / min(ms[1:]) * 1e-9)
            o2["problems_per_s"] = w.batch / o2["ms"] * 1e3
            res.append(o2)
    return res

if __name__ == "__main__":
    print(json.dumps(lx.measure_fp64_peak(0)))
    which = sys.argv[1:] or ["c2", "c3"]
    if "c2" in which:
        print(json.dumps(run_lse("c2", 65536)))
    if "c2_60" in which:
        print(json.dumps(run_lse("c2_60", 16384)))
    if "c3" in which:
        for r in run_lsi("c3", 16384):
            print(json.dumps(r))
    if "c4" in which:
        for r in run_lsi("c4", 1024, reps=3, warm=False):
            print(json.dumps(r))
    if "c2_60" in which and False:
        pass
