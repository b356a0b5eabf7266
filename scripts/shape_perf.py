"""Throughput of one LexLSE (or, with a leading `lsi`, LexLSI cold start: general levels, half of the rows inequalities) shape
with random data: python scripts/shape_perf.py [lsi] <nVar> <batch> <dim0> <dim1> ...
(kernel variant overrides through LEXLS_B200_WARPS / LEXLS_B200_LG_WARPS / LEXLS_B200_FORCE_LODGLOBAL)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lexls_b200 as lx

lsi = sys.argv[1] == "lsi"
av = sys.argv[2:] if lsi else sys.argv[1:]
n, B = int(av[0]), int(av[1])
dims = [int(a) for a in av[2:]]
rng = np.random.default_rng(3)
with (lx.LexLSIBatch(n, dims, [0] * len(dims), batch=B) if lsi else lx.LexLSEBatch(n, dims, batch=B)) as s:
    for k, m in enumerate(dims):
        if lsi:
            d = rng.standard_normal((B, n + 2, m))
            d[:, n + 1, :] = d[:, n, :] + (rng.random((B, m)) * (rng.random(m) < 0.5)[None, :])  # ub = lb (equality) or lb + U(0,1)
            s.setData(k, d)
        else:
            s.setData(k, rng.standard_normal((B, n + 1, m)))
    ms = []
    for _ in range(4):
        s.solve(); ms.append(s.last_solve_ms())
    info = s.kernel_info()
    extra = dict(mean_factorizations=float(s.get_counters()[:, 0].mean())) if lsi else {}
print(json.dumps(dict(lsi=lsi, **extra, n=n, dims=dims, batch=B, ms=min(ms[1:]), problems_per_s=B / min(ms[1:]) * 1e3, kernel=info["name"],
                      per_sm=info["problems_per_sm"], smem=info["smem_bytes"])))
