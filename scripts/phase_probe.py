"""Developer probe: C2 LexLSE with a batch of a few problems per resident warp (all warps of an SM start in phase)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lexls_b200 as lx
from lexls_b200 import problems
batch = int(sys.argv[1])
w = problems.workload("c2", batch)
data, _ = problems.generate(w)
with lx.LexLSEBatch(w.nVar, w.ObjDim, batch=w.batch) as s:
    for k, d in enumerate(data):
        s.setData(k, d)
    ms = []
    for _ in range(6):
        s.solve(); ms.append(s.last_solve_ms())
    print(json.dumps(dict(batch=batch, ms=min(ms[1:]), problems_per_s=batch / min(ms[1:]) * 1e3, info=s.kernel_info())))
