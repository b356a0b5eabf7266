"""C4 (LexLSI n = 200, 10 x 40 rows, 1,024 problems, cold start): one solve with a bounded factorization count, for the ncu
DRAM-traffic passes of profiles/r2_c4_traffic.json.   usage: [LEXLS_B200_LG_WARPS=..] [LEXLS_B200_LG_GRID=..] python scripts/c4_traffic.py [max_fact]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lexls_b200 as lx
from lexls_b200 import problems

cap = int(sys.argv[1]) if len(sys.argv) > 1 else 40
w = problems.workload("c4")
data, vi = problems.generate(w)
with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch) as s:
    s.setParameters(lx.Params(max_number_of_factorizations=cap))
    for k, d in enumerate(data):
        s.setData(k, d)
    s.solve()
    ms = s.last_solve_ms()
    c = s.get_counters()
    info = s.kernel_info()
print(json.dumps({"kernel": info["name"], "grid": info["grid"], "ms": ms, "factorizations_per_problem": float(c[:, 0].mean()),
                  "problems_x_factorizations_per_s": float(c[:, 0].sum()) / (ms * 1e-3), "tflops": float(c[:, 7].astype(np.float64).sum()) * 1e3 / (ms * 1e-3) / 1e12,
                  "env": {k: v for k, v in os.environ.items() if k.startswith("LEXLS_B200_")}}))
