"""BASELINE configs[0]: the reference's own fixture tests/test_01.dat (n = 88, 5 levels 74 simple bounds / 33 / 3 / 2 / 97) as
a batch of 1 (latency of one solve through the C-ABI, device time by CUDA events) and replicated 4,096 times (throughput)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lexls_b200 as lx
from lexls_b200 import datfile

h = datfile.read_hierarchy(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "test_01.dat"))
out = {}
for B in (1, 4096):
    data = [np.ascontiguousarray(np.broadcast_to(np.ascontiguousarray(M.T)[None], (B,) + M.T.shape)) for M in h.objectives]
    with lx.LexLSIBatch(h.nVar, h.ObjDim, h.ObjType, batch=B) as s:
        for k, d in enumerate(data):
            s.setData(k, d, var_index=h.var_index) if h.ObjType[k] == 1 else s.setData(k, d)
        ms, wall = [], []
        for _ in range(5):
            t0 = time.perf_counter(); s.solve(); wall.append((time.perf_counter() - t0) * 1e3); ms.append(s.last_solve_ms())
        x, st, cnt = s.get_x(), s.get_status(), s.get_counters()
        info = s.kernel_info()
    rel = float(np.linalg.norm(x[0] - h.solution) / np.linalg.norm(h.solution))
    out["batch_%d" % B] = dict(kernel=info["name"], device_ms=min(ms), host_wall_ms=min(wall), factorizations=int(cnt[0, 0]),
                              status=int(st[0]), rel_err_vs_file_solution=rel, problems_per_s=B / min(ms) * 1e3,
                              all_identical=bool((x == x[0]).all()))
print(json.dumps(out))
