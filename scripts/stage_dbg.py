import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import lexls_b200 as lx
from lexls_b200 import problems
w = problems.workload("c3", 16384)
data, vi = problems.generate(w)
nl = len(w.ObjDim)
with lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch) as s:
    for k, d in enumerate(data):
        s.setData(k, d, var_index=vi) if w.ObjType[k] == 1 else s.setData(k, d)
    s.solve()
    guess, x0 = [s.getActiveCtr(k) for k in range(nl)], s.get_x()
d2 = problems.perturb(w, data)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
d2 = [pin(a) for a in d2]; guess = [pin(g) for g in guess]; x0 = pin(x0)
x = pin(np.empty_like(x0))
with lx.PipelinedBatch(lx.LexLSIBatch, w.nVar, w.ObjDim, w.ObjType, batch=w.batch, chunks=4) as pb:
    for r in range(6):
        torch.cuda.synchronize(); t = time.perf_counter()
        pb.solve_host(d2, x, var_index=vi, guess=guess, x0=x0)
        torch.cuda.synchronize(); print("round", r, "%.2f ms" % ((time.perf_counter() - t) * 1e3), file=sys.stderr)
