"""Small solves of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lexls_b200 as lx
from lexls_b200 import problems

def lse(name, batch):
    w = problems.workload(name, batch)
    data, _ = problems.generate(w)
    with lx.LexLSEBatch(w.nVar, w.ObjDim, batch=w.batch) as s:
        s.request_outputs(lexqr=True)
        for k, d in enumerate(data):
            s.setData(k, d)
        s.solve()
        return float(np.abs(s.get_x()).sum())

def lsi(name, batch):
    w = problems.workload(name, batch)
    data, vi = problems.generate(w)
    with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch) as s:
        s.request_outputs(multipliers=True, lexqr=True)
        for k, d in enumerate(data):
            s.setData(k, d, var_index=vi) if w.ObjType[k] == 1 else s.setData(k, d)
        s.solve()
        return float(np.abs(s.get_x()).sum())

if __name__ == "__main__":
    print("tiny_lse", lse("tiny_lse", 16))
    print("c2", lse("c2", 48))
    print("tiny_lsi", lsi("tiny_lsi", 16))
    print("c3", lsi("c3", 24))
