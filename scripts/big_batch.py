"""BASELINE configs[4]: a 1,048,576-problem batch (C2 LexLSE shape or C3 LexLSI shape, cold start) sharded over the GPUs of
one box — contiguous ranges of the batch, no data-path collective; only the timing is reduced (max over ranks).
Device-generated data through set_data_dev (no 20 GB host copy), three solves, a property check on a sample.
usage: python scripts/big_batch.py [lse|lsi] [total batch]          (1 GPU)
       python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/big_batch.py ..."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lexls_b200 as lx

kind = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1] in ("lse", "lsi") else "lse"
nums = [a for a in sys.argv[1:] if a.isdigit()]
total = int(nums[0]) if nums else 1 << 20
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
from lexls_b200.shard import shard_range
lo, hi = shard_range(total, world, rank)
B = hi - lo
g = torch.Generator(device="cuda").manual_seed(1 + rank)
rnd = lambda *s: torch.randn(s, dtype=torch.float64, device="cuda", generator=g)
uni = lambda *s: torch.rand(s, dtype=torch.float64, device="cuda", generator=g)
out = {}
if kind == "lse":
    n, dims = 40, [20, 20, 20]
    lev = [rnd(B, n + 1, m) for m in dims]
    torch.cuda.synchronize()
    with lx.LexLSEBatch(n, dims, batch=B, device=dev) as s:
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
    out = dict(all_solved=bool((st == 0).all()), full_rank=int(((ranks[:, 0] == 20) & (ranks[:, 1] == 20)).sum()),
               max_abs_residual_levels_1_2=worst)
else:
    # C3 shape: 38 joint limits (simple bounds), 12 equalities, 6 inequalities, 6 equalities, posture I x = x_ref
    n, dims, types = 38, [38, 12, 6, 6, 38], [1, 0, 0, 0, 0]
    lev = []
    for k, m in enumerate(dims[1:], 1):
        t = torch.empty((B, n + 2, m), dtype=torch.float64, device="cuda")
        if k == 4:
            t[:, :n, :] = torch.eye(n, m, dtype=torch.float64, device="cuda")
            t[:, n, :] = 0.1 * rnd(B, m)
            t[:, n + 1, :] = t[:, n, :]
        else:
            t[:, :n, :] = rnd(B, n, m)
            if k == 2:
                t[:, n + 1, :] = uni(B, m)
                t[:, n, :] = t[:, n + 1, :] - uni(B, m)
            else:
                t[:, n, :] = rnd(B, m)
                t[:, n + 1, :] = t[:, n, :]
        lev.append(t)
    sb = torch.stack([-0.5 - uni(B, 38), 0.5 + uni(B, 38)], dim=1).cpu().numpy()
    torch.cuda.synchronize()
    with lx.LexLSIBatch(n, dims, types, batch=B, device=dev) as s:
        s.setData(0, sb, var_index=np.arange(38, dtype=np.uint32))
        for k, t in enumerate(lev, 1):
            s.setData_dev(k, t.data_ptr())
        ms = []
        for _ in range(3):
            s.solve(); ms.append(s.last_solve_ms())
        x = s.get_x(); st = s.get_status(); cnt = s.get_counters()
        info = s.kernel_info()
    idx = np.random.default_rng(0).choice(B, 2048, replace=False)
    inside = bool(((x[idx] >= sb[idx, 0] - 1e-9) & (x[idx] <= sb[idx, 1] + 1e-9)).all())  # level 0 is always feasible
    A = lev[0][idx].cpu().numpy()
    r = np.einsum("bcm,bc->bm", A[:, :n, :], x[idx]) - A[:, n, :]
    out = dict(status_counts={int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))},
               mean_factorizations=float(cnt[:, 0].mean()), joint_limits_respected_on_sample=inside,
               median_abs_residual_level_1=float(np.median(np.abs(r))))
t = torch.tensor([min(ms)], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # timing only: the job is as slow as its slowest rank
    agg = [None] * world
    dist.all_gather_object(agg, out)
if rank == 0:
    ms_all = float(t.item())
    print(json.dumps(dict(kind=kind, batch=total, n_gpus=world, batch_per_gpu=B, ms=ms_all, problems_per_s=total / ms_all * 1e3,
                          kernel=info["name"], scaling="strong", rank0=out)))
if world > 1:
    dist.destroy_process_group()
