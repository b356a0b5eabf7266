"""BASELINE configs[4]: a 1,048,576-problem LexLSE batch (C2 shape) on one GPU — device-generated data through set_data_dev
(no 20 GB host copy), one solve, a residual property check on a sample."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lexls_b200 as lx

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
n, dims = 40, [20, 20, 20]
g = torch.Generator(device="cuda").manual_seed(1)
lev = [torch.randn((B, n + 1, m), dtype=torch.float64, device="cuda", generator=g) for m in dims]
torch.cuda.synchronize()
with lx.LexLSEBatch(n, dims, batch=B) as s:
    for k, t in enumerate(lev):
        s.setData_dev(k, t.data_ptr())
    ms = []
    for _ in range(3):
        s.solve(); ms.append(s.last_solve_ms())
    x = s.get_x(); st = s.get_status(); ranks = s.get_ranks()
    info = s.kernel_info()
idx = np.random.default_rng(0).choice(B, 2048, replace=False)
worst = 0.0
for k in range(2):  # levels 1 and 2 are satisfied exactly by a full-rank problem (40 equations, 40 unknowns)
    A = lev[k][idx].cpu().numpy()
    r = np.einsum("bcm,bc->bm", A[:, :n, :], x[idx]) - A[:, n, :]
    worst = max(worst, float(np.abs(r).max()))
print(json.dumps(dict(batch=B, ms=min(ms), problems_per_s=B / min(ms) * 1e3, kernel=info["name"], all_solved=bool((st == 0).all()),
                      full_rank=int(((ranks[:, 0] == 20) & (ranks[:, 1] == 20)).sum()), max_abs_residual_levels_1_2=worst)))
