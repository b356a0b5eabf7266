#!/usr/bin/env python
"""bench.py — headline benchmark of the batched LexLSE / LexLSI hot path (BASELINE.json metric: problems solved / s).

  python bench.py --gpus N --steps K --warmup W [--workload c2|c3|c4] [--impl reference]

A "step" is one pass of the hot path over one resident batch of synthetic problems (one solve_batch launch per GPU).
Default workload = BASELINE.json configs[1] (batched LexLSE, 3 levels, n=40, rows 20/20/20, 65,536 problems per GPU);
`--workload c3` is configs[2] (batched LexLSI, humanoid-IK shape, 5 levels, n=38, 16,384 problems per GPU, warm-started),
`--workload c4` configs[3] (LexLSI n=200, 10 x 40 rows, 1,024 problems per GPU).  The default line also carries C3, C4 and the
1,048,576-problem strong-scaling sweep of configs[4] as complete objects under "also".
With N > 1 (torchrun, one rank per GPU) ONE global batch of N x batch_per_gpu problems is sharded by contiguous ranges
(lexls_b200.shard.shard_range) — weak scaling, no data-path collective; timing = max over ranks, device events on the launch
stream.  The end-to-end legs (host buffers in, x out through the public API) end with the gather of x to rank 0.
`--impl reference` times the reference's CPU algorithm (the oracle: a scalar C++ restatement — the reference needs Eigen,
which this image does not have) with all host threads on bounded samples of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "problems_solved_per_sec"
UNIT = "problems/s"


# ---------------------------------------------------------------------------------------------------------------------
def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c2_60"])
    ap.add_argument("--batch", type=int, default=None, help="problems per GPU (default: the config's batch)")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--e2e-chunks", type=int, default=4,
                    help="end-to-end leg: sub-batches whose host->device copies overlap the previous chunk's solve (1 = plain calls)")
    ap.add_argument("--cpu-sample", type=int, default=None)
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workload line items")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(
                    self.nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.ok:
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------------------
def make_workload(name, batch, rank):
    from lexls_b200 import problems
    w = problems.workload(name, batch)
    data, vi = problems.generate(w, start=rank * w.batch, count=w.batch)
    return w, data, vi


def cpu_run(w, data, vi, sample, threads, guess=None, x0=None, params=None):
    """The reference's CPU algorithm (oracle restatement) on the first `sample` problems; returns seconds."""
    import oracle
    sl = [d[:sample] for d in data]
    op = oracle.OracleParams(**params) if params else None
    t0 = time.perf_counter()
    if w.kind == "lse":
        oracle.lse_solve_batch(w.nVar, w.ObjDim, sl, threads=threads)
    else:
        vib = np.tile(vi, (sample, 1)) if vi is not None else None
        oracle.lsi_solve_batch(w.nVar, w.ObjDim, w.ObjType, sl, var_index=vib,
                               guess=[g[:sample] for g in guess] if guess is not None else None,
                               x0=x0[:sample] if x0 is not None else None, threads=threads, params=op)
    return time.perf_counter() - t0


class Runner:
    """Owns one resident batch on one GPU and knows how to run a device-resident step and an end-to-end step."""

    def __init__(self, w, data, vi, device, torch, params=None):
        import lexls_b200 as lx
        self.lx, self.w, self.vi, self.torch = lx, w, vi, torch
        self.P = len(w.ObjDim)
        self.params = params
        # strict=False: input validation is deferred to the solve — one synchronisation per batch, not one per setData
        if w.kind == "lse":
            self.s = lx.LexLSEBatch(w.nVar, w.ObjDim, batch=w.batch, device=device, strict=False)
        else:
            self.s = lx.LexLSIBatch(w.nVar, w.ObjDim, w.ObjType, batch=w.batch, device=device, strict=False)
            if params:
                self.s.setParameters(lx.Params(**params))
        sp = torch.cuda.current_stream().cuda_stream
        assert sp != 0, "bench must run on a non-default stream (NULL selects the handle's own stream)"
        self.s.set_stream(sp)
        self.guess = self.x0 = None
        self.cold_data = data
        self._set(data)
        if w.kind == "lsi":
            # warm-start workload: cold solve, then the perturbed problem with the previous active set + x
            from lexls_b200 import problems
            self.s.solve()
            self.cold_ms = self.s.last_solve_ms()
            self.cold_counters = self.s.get_counters()
            self.cold_status = self.s.get_status()
            self.guess = [self.s.getActiveCtr(k) for k in range(self.P)]
            self.x0 = self.s.get_x()
            data = problems.perturb(w, data)
            self._set(data)
            for k, g in enumerate(self.guess):
                self.s.set_active_guess(k, g)
            self.s.set_x0(self.x0)
        self.data = data
        # pinned host copies for the end-to-end leg
        self.pinned = []
        for d in data:
            t = torch.empty(d.shape, dtype=torch.float64, pin_memory=True)
            t.numpy()[...] = d
            self.pinned.append(t)
        self.x_pinned = torch.empty((w.batch, w.nVar), dtype=torch.float64, pin_memory=True)
        self.h2d_bytes = int(sum(d.nbytes for d in data))
        self.d2h_bytes = int(self.x_pinned.numel() * 8)
        if w.kind == "lsi":
            self.guess_pinned = []
            for g in self.guess:
                t = torch.empty(g.shape, dtype=torch.uint8, pin_memory=True)
                t.numpy()[...] = g
                self.guess_pinned.append(t)
            self.x0_pinned = torch.empty(self.x0.shape, dtype=torch.float64, pin_memory=True)
            self.x0_pinned.numpy()[...] = self.x0
            self.h2d_bytes += int(sum(g.nbytes for g in self.guess) + self.x0.nbytes)

    def _set(self, data):
        for k, d in enumerate(data):
            if self.w.kind == "lsi" and self.w.ObjType[k] == 1:
                self.s.setData(k, d, var_index=self.vi)
            else:
                self.s.setData(k, d)

    def step(self):
        self.s.solve(sync=False)

    def make_device_inputs(self):
        torch = self.torch
        self.dev = [torch.from_numpy(np.ascontiguousarray(d)).cuda() for d in self.data]
        if self.w.kind == "lsi":
            self.dev_guess = [torch.from_numpy(g).cuda() for g in self.guess]
            self.dev_x0 = torch.from_numpy(self.x0).cuda()
            self.dev_vi = torch.from_numpy(np.ascontiguousarray(self.vi).astype(np.int32)).cuda() if self.vi is not None else None
        torch.cuda.synchronize()

    def free_device_inputs(self):
        self.dev = self.dev_guess = self.dev_x0 = self.dev_vi = None

    def e2e_step_device_inputs(self):
        """Inputs already in device memory (same caller layout): set_data_dev, solve, x into pinned host memory."""
        import ctypes as C
        from lexls_b200.capi import check, lib
        s = self.s
        for k, t in enumerate(self.dev):
            if self.w.kind == "lsi" and self.w.ObjType[k] == 1:
                m = self.w.ObjDim[0]
                check(lib().lexls_b200_set_simple_bounds_dev(s._h, C.cast(C.c_void_p(self.dev_vi.data_ptr()), C.POINTER(C.c_uint32)), C.c_uint64(0),
                                                             C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_double)), C.c_uint32(max(m, 1)), C.c_uint64(2 * m)))
            else:
                s.setData_dev(k, t.data_ptr())
        if self.w.kind == "lsi":
            for k, g in enumerate(self.dev_guess):
                check(lib().lexls_b200_set_active_guess_dev(s._h, C.c_uint32(k), C.cast(C.c_void_p(g.data_ptr()), C.POINTER(C.c_uint8))))
            check(lib().lexls_b200_set_x0_dev(s._h, C.cast(C.c_void_p(self.dev_x0.data_ptr()), C.POINTER(C.c_double))))
        s.solve(sync=False)
        s.get_x(out=self.x_pinned.numpy())

    def make_pipeline(self, chunks, device):
        """The end-to-end leg through lexls_b200.PipelinedBatch (public API): same inputs, `chunks` sub-batches."""
        lx = self.lx
        cls = lx.LexLSEBatch if self.w.kind == "lse" else lx.LexLSIBatch
        self.pipe = lx.PipelinedBatch(cls, self.w.nVar, self.w.ObjDim, self.w.ObjType if self.w.kind == "lsi" else None,
                                      batch=self.w.batch, device=device, chunks=chunks)
        if self.params and self.w.kind == "lsi":
            self.pipe.setParameters(lx.Params(**self.params))
        self.pipe_host = [t.numpy() for t in self.pinned]
        self.pipe_guess = [g.numpy() for g in self.guess_pinned] if self.w.kind == "lsi" else None
        self.pipe_x0 = self.x0_pinned.numpy() if self.w.kind == "lsi" else None

    def e2e_step_pipelined(self):
        self.pipe.solve_host(self.pipe_host, self.x_pinned.numpy(), var_index=self.vi, guess=self.pipe_guess, x0=self.pipe_x0)

    def e2e_step(self):
        """What a caller does per batch: host data in (pinned), solve, x out (pinned)."""
        for k, t in enumerate(self.pinned):
            if self.w.kind == "lsi" and self.w.ObjType[k] == 1:
                self.s.setData(k, t.numpy(), var_index=self.vi)
            else:
                self.s.setData(k, t.numpy())
        if self.w.kind == "lsi":
            for k, g in enumerate(self.guess_pinned):
                self.s.set_active_guess(k, g.numpy())
            self.s.set_x0(self.x0_pinned.numpy())
        self.s.solve(sync=False)
        self.s.get_x(out=self.x_pinned.numpy())


def timed(torch, dist, world, fn, steps):
    """barrier + sync, `steps` calls of fn, sync + barrier; device time by CUDA events on the current stream; max over ranks."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    return ms


def flops_of(counters):
    return float(counters[:, 7].astype(np.float64).sum()) * 1e3


def load_traffic(kernel_name, workload):
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get("%s:%s" % (workload, kernel_name))
    except Exception:
        return None


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# ---------------------------------------------------------------------------------------------------------------------
CPU_SAMPLE = {"c2": 32768, "c2_60": 8192, "c3": 8192, "c4": 64}
C4_MAX_FACT = 1000  # C4 cold needs 310-380 factorizations per problem: with the default cap of 200 nothing would be SOLVED


def h2d_ceiling(torch, dist, world, nbytes, reps=3):
    """What the host link allows: every rank copies `nbytes` from pinned host memory to its GPU at the same time (one plain
    cudaMemcpyAsync, no kernel).  Returns GB/s per rank (max time over ranks)."""
    nbytes = int(min(nbytes, 2 << 30))
    src = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dst.copy_(src, non_blocking=True)
    ms = timed(torch, dist, world, lambda: dst.copy_(src, non_blocking=True), reps)
    del src, dst
    return nbytes * reps / (ms * 1e-3) / 1e9


def measure(name, args, torch, dist, rank, world, local, steps, e2e_steps, batch=None, with_cpu=True, params=None):
    """One workload on all ranks: device-resident value, roofline, end-to-end legs, clocks; rank 0 returns the dict."""
    import lexls_b200 as lx
    from lexls_b200.shard import gather_to_rank0, shard_range
    from lexls_b200 import problems
    w0 = problems.workload(name, batch)
    global_batch = w0.batch * world                      # ONE global batch ...
    lo, hi = shard_range(global_batch, world, rank)     # ... sharded by contiguous problem ranges (lexls_b200/shard.py)
    w = problems.workload(name, hi - lo)
    data, vi = problems.generate(w, start=lo, count=hi - lo)
    r = Runner(w, data, vi, local, torch, params=params)
    r.x_all = torch.empty((global_batch, w.nVar), dtype=torch.float64, pin_memory=True) if (world > 1 and rank == 0) else None
    info = r.s.kernel_info()
    for _ in range(max(args.warmup, 3)):
        r.step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(torch, dist, world, r.step, steps)
    clocks = sampler.stop()
    counters, status = r.s.get_counters(), r.s.get_status()
    ms_per_step = ms / steps
    value = global_batch * steps / (ms * 1e-3)

    # ---- end to end: host buffers in, x out, the final gather of the shards' results to rank 0 inside the timed region ------
    def with_gather(fn):
        def step():
            fn()
            if world > 1:
                gather_to_rank0(r.x_pinned, global_batch, out=r.x_all)
        return step
    legs = {}
    for _ in range(2):
        with_gather(r.e2e_step)()
    legs["plain"] = timed(torch, dist, world, with_gather(r.e2e_step), e2e_steps)
    chunks = args.e2e_chunks
    if chunks > 1:
        r.make_pipeline(chunks, local)
        for _ in range(2):
            with_gather(r.e2e_step_pipelined)()
        legs["pipelined"] = timed(torch, dist, world, with_gather(r.e2e_step_pipelined), e2e_steps)
        r.pipe.close()
    # inputs already on the device (a caller that produces its problems on the GPU): set_data_dev, solve, x out + gather
    r.make_device_inputs()
    for _ in range(2):
        with_gather(r.e2e_step_device_inputs)()
    legs["device_inputs"] = timed(torch, dist, world, with_gather(r.e2e_step_device_inputs), e2e_steps)
    r.free_device_inputs()
    link_gbs = h2d_ceiling(torch, dist, world, r.h2d_bytes)
    apis = {"plain": "LexLS%sBatch: setData%s, solve, get_x" % ("E" if w.kind == "lse" else "I", "" if w.kind == "lse" else ", set_active_guess, set_x0"),
            "pipelined": "PipelinedBatch.solve_host, %d chunks" % chunks}
    best = min((k for k in legs if k != "device_inputs"), key=lambda k: legs[k])
    rate = lambda t: global_batch * e2e_steps / (t * 1e-3)

    out = None
    if rank == 0:
        peaks = load_peaks()
        fp = lx.measure_fp64_peak(local)
        fl = flops_of(counters) * world  # every rank runs the same shapes: algorithmic flops of the whole job
        achieved = flops_of(counters) / (ms_per_step * 1e-3) / 1e12
        in_bytes = float(sum(d.nbytes for d in r.data))
        out_bytes = float(w.batch * (w.nVar * 8 + w.nCtr * 8 + 48))
        hbm_gbs = (in_bytes + out_bytes) / (ms_per_step * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": w.name, "kind": w.kind, "nVar": w.nVar, "levels": w.ObjDim, "batch_per_gpu": w.batch,
                       "global_batch": global_batch, "seed": w.seed, "warm_start": w.kind == "lsi",
                       "parallelism": "one global batch of %d problems sharded by contiguous ranges (lexls_b200.shard.shard_range), "
                                      "one rank per GPU, no data-path collective; e2e ends with the gather of x to rank 0" % global_batch,
                       "l2": "inputs larger than L2: %.0f MB of problem records per GPU are re-read from HBM every step" % (in_bytes / 1e6),
                       "kernel": info},
            "clocks": clocks,
            "e2e": {"value": rate(legs[best]), "unit": UNIT, "h2d_bytes_per_step": r.h2d_bytes, "d2h_bytes_per_step": r.d2h_bytes,
                    "steps": e2e_steps, "ms_per_step": legs[best] / e2e_steps, "api": apis[best],
                    "plain_calls_value": rate(legs["plain"]),
                    "pipelined_value": rate(legs["pipelined"]) if "pipelined" in legs else None,
                    "device_inputs_value": rate(legs["device_inputs"]),
                    "device_inputs_api": "set_data_dev (device pointers), solve, get_x into pinned host memory",
                    "numa": ("rank bound to the %d CPUs local to its GPU before its pinned buffers were allocated" % args.numa_cpus) if getattr(args, "numa_cpus", 0) else "not bound",
                    "final_gather": "lexls_b200.shard.gather_to_rank0(x): shards GPU to GPU over NCCL into one device buffer on rank 0, one D2H into a pinned [global_batch, nVar] buffer; inside the timed region" if world > 1 else "single rank: x lands in the caller's pinned buffer",
                    "host_link": {"h2d_gbs_per_gpu_all_ranks_copying": link_gbs,
                                  "ceiling_problems_per_s": global_batch / (r.h2d_bytes / (link_gbs * 1e9)),
                                  "note": "a plain pinned cudaMemcpyAsync of the step's input bytes on every rank at once: the "
                                          "host-buffer legs cannot be faster than this"}},
            "gpu_launches": steps * info["launches_per_solve"],
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp["dfma_tflops"], "unit": "TFLOP/s",
                         "frac": achieved / fp["dfma_tflops"], "traffic": load_traffic(info["name"], w.name),
                         "peak_source": "measured here: DFMA chain micro-benchmark (MEASURED_PEAKS.json has no FP64 entry); "
                                        "DMMA mma.sync f64 chain = %.1f TFLOP/s" % fp["dmma_tflops"],
                         "algorithmic_flops_per_launch": flops_of(counters), "algorithmic_flops_whole_job_per_step": fl,
                         "kernel": info["name"],
                         "hbm": {"achieved": hbm_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                                 "frac": (hbm_gbs / peaks["hbm_gbs"]) if peaks.get("hbm_gbs") else None,
                                 "algorithmic_bytes_per_launch": in_bytes + out_bytes}},
            "solve_stats": {"status_counts": {str(int(k)): int((status == k).sum()) for k in np.unique(status)},
                            "status_solved_frac": float((status == 0).mean()),
                            "mean_factorizations": float(counters[:, 0].mean()), "mean_rank": float(counters[:, 5].mean())},
        }
        if w.kind == "lsi":
            cold_tf = flops_of(r.cold_counters) / (r.cold_ms * 1e-3) / 1e12
            out["cold_start"] = {"value": global_batch / (r.cold_ms * 1e-3), "unit": UNIT, "ms": r.cold_ms,
                                 "mean_factorizations": float(r.cold_counters[:, 0].mean()),
                                 "status_counts": {str(int(k)): int((r.cold_status == k).sum()) for k in np.unique(r.cold_status)},
                                 "achieved_tflops": cold_tf, "roofline_frac": cold_tf / fp["dfma_tflops"],
                                 "note": "one untimed-warm-up-free solve from scratch (no guess, no x0); rank 0's shard time, all ranks run concurrently"}
            if world == 1 and with_cpu:
                cs = min({"c3": 2048, "c4": 32}.get(name, 512), w.batch)
                sec = cpu_run(w, r.cold_data, vi, cs, host_threads(), None, None, params=params)
                out["cold_start"]["cpu_baseline"] = {"value": cs / sec, "unit": UNIT, "cores": host_threads(), "kind": "port",
                                                     "sample": "first %d problems, cold start, %.2f s wall" % (cs, sec)}
        if world == 1 and with_cpu:
            threads = host_threads()
            sample = min(args.cpu_sample or CPU_SAMPLE.get(name, 4096), w.batch)
            sec = cpu_run(w, r.data, vi, sample, threads, r.guess, r.x0, params=params)
            out["cpu_baseline"] = {"value": sample / sec, "unit": UNIT, "cores": threads, "kind": "port",
                                   "sample": "first %d problems of the same workload; SCALAR C++ restatement of the reference's algorithm "
                                             "(-O3 -march=native, fixed summation orders; the reference itself needs Eigen, absent "
                                             "here: an Eigen build would be faster), one problem per thread on %d host threads, %.2f s wall"
                                             % (sample, threads, sec)}
    r.s.close()
    return out


def strong_scaling_c5(args, torch, dist, rank, world, local):
    """BASELINE configs[4]: 1,048,576 LexLSE problems (C2 shape) sharded over the ranks — strong scaling, device-resident.
    Problem p carries the data of problem p mod 65,536 (generated once, tiled on the device), fed through set_data_dev."""
    import lexls_b200 as lx
    from lexls_b200 import problems
    from lexls_b200.shard import shard_range
    total = 1 << 20
    lo, hi = shard_range(total, world, rank)
    n = hi - lo
    w = problems.workload("c2", 65536)
    data, _ = problems.generate(w)
    with lx.LexLSEBatch(w.nVar, w.ObjDim, batch=n, device=local, strict=False) as s:
        s.set_stream(torch.cuda.current_stream().cuda_stream)
        keep = []
        for k, d in enumerate(data):
            t = torch.from_numpy(d).cuda()
            reps = (n + 65535) // 65536
            t = t.repeat(reps, 1, 1)[:n].contiguous()
            keep.append(t)
            s.setData_dev(k, t.data_ptr())
        torch.cuda.synchronize()
        del keep
        step = lambda: s.solve(sync=False)
        for _ in range(3):
            step()
        k = 10
        ms = timed(torch, dist, world, step, k)
        st = s.get_status()
    if rank != 0:
        return None
    return {"workload": "c5_lexlse_1m_strong", "global_batch": total, "batch_per_gpu": n, "n_gpus": world, "scaling": "strong",
            "value": total * k / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / k, "status_solved_frac": float((st == 0).mean()),
            "data": "problem p = problem (p mod 65,536) of the C2 generator, tiled on the device, set_data_dev"}


def one_process_multi_gpu(args, torch, world, name="c2"):
    """The product's own multi-GPU entry (lexls_b200_create_batch_multi: one host thread + stream per device inside ONE process)
    driven by rank 0 over all `world` GPUs of the box while the other ranks wait: host buffers in, x out of one pinned buffer."""
    import lexls_b200 as lx
    from lexls_b200 import problems
    w = problems.workload(name, None)
    B = w.batch * world
    w = problems.workload(name, B)
    data, vi = problems.generate(w)
    pinned = []
    for d in data:
        t = torch.empty(d.shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = d
        pinned.append(t)
    x = torch.empty((B, w.nVar), dtype=torch.float64, pin_memory=True)
    with lx.LexLSEBatch(w.nVar, w.ObjDim, batch=B, devices=list(range(world)), strict=False) as s:
        def step():
            for k, t in enumerate(pinned):
                s.setData(k, t.numpy())
            s.solve(sync=False)
            s.get_x(out=x.numpy())
        for _ in range(2):
            step()
        k = 5
        t0 = time.perf_counter()
        for _ in range(k):
            step()
        sec = time.perf_counter() - t0
        ms_dev = s.last_solve_ms()
        st = s.get_status()
        shards = s.shards()
    return {"api": "LexLSEBatch(devices=[0..%d]) = lexls_b200_create_batch_multi: setData, solve, get_x" % (world - 1),
            "workload": w.name, "global_batch": B, "n_gpus": world, "shards": [c for _, _, c in shards],
            "e2e_value": B * k / sec, "unit": UNIT, "ms_per_step": sec / k * 1e3, "slowest_shard_kernel_ms": ms_dev,
            "device_resident_value": B / (ms_dev * 1e-3), "status_solved_frac": float((st == 0).mean()),
            "timing": "host wall clock around %d steps (the call blocks until x is in the caller's buffer)" % k}


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — lexls_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # every rank copies through its own host link at once: keep its pinned buffers on the GPU's NUMA node
        from lexls_b200.shard import bind_to_gpu_numa_node
        args.numa_cpus = bind_to_gpu_numa_node(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import lexls_b200 as lx

    # a dedicated (non-default) stream: the library launches on it and the torch events below are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    e2e_steps = args.e2e_steps if args.e2e_steps is not None else max(3, min(args.steps, 10))
    par4 = dict(max_number_of_factorizations=C4_MAX_FACT)
    out = measure(args.workload, args, torch, dist, rank, world, local, args.steps, e2e_steps, batch=args.batch,
                  params=par4 if args.workload == "c4" else None)

    # the other BASELINE configs as complete objects on the same line (every rank takes part; rank 0 keeps the result).
    # A watchdog guarantees the headline line: if the extras hang (a rank died inside a collective), rank 0 prints what it has.
    also = {}
    done = threading.Event()

    def watchdog():
        if not done.wait(timeout=float(os.environ.get("LEXLS_BENCH_EXTRA_TIMEOUT", "420"))) and rank == 0:
            out["also"] = dict(also, error="extras timed out")
            print(json.dumps(out), flush=True)
            os._exit(0)
    if rank == 0 and not args.no_extra:
        threading.Thread(target=watchdog, daemon=True).start()
    if not args.no_extra and args.workload == "c2":
        def guarded(key, fn):
            try:
                res = fn()
                if rank == 0:
                    also[key] = res
            except Exception as e:  # the headline line must still print
                if rank == 0:
                    also[key] = {"error": repr(e)}
                if world > 1:
                    raise
        # C3 — the north-star shape (LexLSI, 5 levels, n = 38, warm start), first class: value, roofline, e2e, cpu_baseline
        guarded("c3_lexlsi_warm", lambda: measure("c3", args, torch, dist, rank, world, local, max(20, args.steps // 2), min(e2e_steps, 5)))
        # C4 — LexLSI n = 200, 10 x 40 rows, cold start until PROBLEM_SOLVED (cap raised to 1,000 factorizations)
        guarded("c4_lexlsi", lambda: measure("c4", args, torch, dist, rank, world, local, 3, 2, params=par4))
        # C5 — 1,048,576 problems, strong scaling
        guarded("c5_strong", lambda: strong_scaling_c5(args, torch, dist, rank, world, local))
        if world > 1:
            # the in-process multi-GPU entry of the C-ABI, driven by rank 0 over all GPUs while the others wait
            dist.barrier()
            if rank == 0:
                try:
                    also["one_process_multi_gpu"] = one_process_multi_gpu(args, torch, world)
                except Exception as e:
                    also["one_process_multi_gpu"] = {"error": repr(e)}
            dist.barrier()
    done.set()
    if rank == 0:
        if also:
            out["also"] = also
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    w, data, vi = make_workload(args.workload, None, 0)
    threads = host_threads()
    sample = args.cpu_sample or {"c2": 4096, "c2_60": 2048, "c3": 1024, "c4": 64}.get(args.workload, 1024)
    sample = min(sample, w.batch)
    params = dict(max_number_of_factorizations=C4_MAX_FACT) if args.workload == "c4" else None
    guess = x0 = None
    if w.kind == "lsi":
        # same warm-start workload as the GPU arm: cold solve of the sample (untimed), then the perturbed problems
        import oracle
        from lexls_b200 import problems
        sl = [d[:sample] for d in data]
        vib = np.tile(vi, (sample, 1)) if vi is not None else None
        cold = oracle.lsi_solve_batch(w.nVar, w.ObjDim, w.ObjType, sl, var_index=vib, threads=threads,
                                      params=oracle.OracleParams(**params) if params else None)
        guess, x0 = cold["active"], cold["x"]
        data = problems.perturb(w, data)
    steps, warm = max(args.steps, 1), max(args.warmup, 1)
    # bound the whole run to a few minutes
    t_probe = cpu_run(w, data, vi, sample, threads, guess, x0, params)
    budget = 120.0
    if t_probe * (steps + warm) > budget:
        steps = max(1, int(budget / t_probe) - warm)
    for _ in range(warm - 1):
        cpu_run(w, data, vi, sample, threads, guess, x0, params)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_run(w, data, vi, sample, threads, guess, x0, params)
    sec = time.perf_counter() - t0
    value = sample * steps / sec
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": sec / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": w.name, "kind": w.kind, "nVar": w.nVar, "levels": w.ObjDim, "batch_per_gpu": w.batch,
                   "warm_start": w.kind == "lsi", "step": "bounded sample of %d problems of the workload per step" % sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d problems per step x %d steps, one problem per thread on %d host threads; CPU oracle = SCALAR C++ "
                                   "restatement of the reference algorithm, -O3 -march=native, fixed summation orders (the reference's "
                                   "only dependency, Eigen, is absent, so its own build is impossible here; an Eigen build would be faster)"
                                   % (sample, steps, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
