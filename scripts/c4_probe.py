import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lexls_b200 as lx
from lexls_b200 import problems
w = problems.workload("c4", 296)
data, vi = problems.generate(w)
with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch) as s:
    s.setParameters(lx.Params(max_number_of_factorizations=12))
    for k, d in enumerate(data):
        s.setData(k, d)
    for _ in range(2):
        s.solve(); print(s.last_solve_ms(), s.kernel_info()["name"])
    c = s.get_counters(); print(c[:, 0].mean(), c[:, 5].mean())
