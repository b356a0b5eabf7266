"""Multi-GPU sharding of a batch: problems are independent, so a batch is split into contiguous ranges,
one per GPU, with no data-path collective; the only exchange is the final gather of results.

Two ways to use several GPUs:
  * inside ONE process: :class:`ShardedBatch` owns one handle (and stream) per device, launches all shards asynchronously
    and concatenates the results;
  * one process per GPU (torchrun): every rank solves ``shard_range(batch, world, rank)`` with a plain
    LexLSEBatch / LexLSIBatch and calls :func:`gather_to_rank0` on the result arrays.
"""
import os

import numpy as np


def bind_to_gpu_numa_node(device):
    """Pins the calling process to the CPUs that are local to `device` (NVML's ideal CPU affinity) so that the pinned host
    buffers it allocates afterwards are first-touched on the GPU's NUMA node: with several ranks copying at once the host
    link otherwise saturates on one socket's memory.  Returns the number of CPUs bound to, or 0 when NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device))
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def shard_range(batch, world, rank):
    """Contiguous range [lo, hi) of rank `rank`; the first (batch % world) ranks get one extra problem."""
    base, extra = divmod(int(batch), int(world))
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def shard_ranges(batch, world):
    return [shard_range(batch, world, r) for r in range(world)]


def tapered_ranges(batch, chunks):
    """Chunk ranges for a copy-bound pipeline: equal chunks, but the last two are a half and a quarter — what remains after
    the last copy is the last chunk's solve, so that chunk should be small; the earlier ones amortise the per-chunk overheads.
    Measured (4 chunks, 1 x B200): C2 2.65 -> 2.72 M problems/s (96 % of the link ceiling), C3 2.16 -> 2.27 M."""
    chunks = max(1, min(int(chunks), int(batch)))
    if chunks < 3 or batch < 8192:  # a small batch: the small chunks would not fill the GPU (C4: 1,024 problems, 296 in flight)
        return shard_ranges(batch, chunks)
    w = [1.0] * chunks
    w[-2], w[-1] = 0.5, 0.25
    tot, acc, cuts = sum(w), 0.0, [0]
    for x in w[:-1]:
        acc += x
        cuts.append(min(int(batch), max(cuts[-1] + 1, int(round(batch * acc / tot)))))
    cuts.append(int(batch))
    return [(cuts[i], cuts[i + 1]) for i in range(chunks)]


_GATHER_CACHE = {}


def gather_to_rank0(local, batch, group=None, out=None):
    """Final gather over torch.distributed (any backend): rank 0 returns the [batch, ...] array, other ranks None.

    `local` is this rank's [rows, ...] result (numpy array, or a torch tensor — pinned host memory makes the staging copy
    asynchronous).  With NCCL the shards travel GPU to GPU (NVLink) into one device buffer on rank 0 and come down in ONE
    device-to-host copy; staging and receive buffers are cached per shape, `out` (a pinned [batch, ...] torch tensor or numpy
    array over pinned memory) receives the result without a further copy."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    t = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local))
    orig = t.dtype
    if orig == torch.uint32:   # collectives lack unsigned 32-bit: ship the bits as int32
        t = t.view(torch.int32)
    tail = tuple(t.shape[1:])
    sizes = [hi - lo for lo, hi in shard_ranges(batch, world)]
    smax = max(sizes)
    nccl = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu")
    key = (str(dev), t.dtype, smax, tail, world, rank == 0)
    bufs = _GATHER_CACHE.get(key)
    if bufs is None:
        send = torch.zeros((smax,) + tail, dtype=t.dtype, device=dev)  # padded: gather needs equal shapes
        recv = torch.empty((world, smax) + tail, dtype=t.dtype, device=dev) if rank == 0 else None
        bufs = _GATHER_CACHE[key] = (send, recv)
    send, recv = bufs
    send[:t.shape[0]].copy_(t, non_blocking=True)
    dist.gather(send, list(recv.unbind(0)) if rank == 0 else None, dst=0, group=group)
    if rank != 0:
        return None
    if out is None:
        res = torch.empty((int(batch),) + tail, dtype=t.dtype, pin_memory=nccl)
    else:
        res = out if isinstance(out, torch.Tensor) else torch.from_numpy(out)
        if orig == torch.uint32:
            res = res.view(torch.int32)
    lo = 0
    for r, n in enumerate(sizes):
        res[lo:lo + n].copy_(recv[r, :n], non_blocking=True)
        lo += n
    if nccl:
        torch.cuda.current_stream().synchronize()
    a = res.numpy()
    return a.view(np.uint32) if orig == torch.uint32 else a


class ShardedBatch:
    """One logical batch over several GPUs of one box, inside one process."""

    def __init__(self, cls, nVar, ObjDim, ObjType=None, batch=1, devices=(0,)):
        self.devices = list(devices)
        self.ranges = shard_ranges(batch, len(self.devices))
        self.batch = int(batch)
        self.parts = []
        for dev, (lo, hi) in zip(self.devices, self.ranges):
            if hi > lo:
                kw = dict(batch=hi - lo, device=dev, strict=False)
                self.parts.append((lo, hi, cls(nVar, ObjDim, ObjType, **kw) if ObjType is not None else cls(nVar, ObjDim, **kw)))

    def close(self):
        for _, _, s in self.parts:
            s.close()
        self.parts = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def setParameters(self, p):
        for _, _, s in self.parts:
            s.setParameters(p)

    def setData(self, level, data, **kw):
        for lo, hi, s in self.parts:
            s.setData(level, data[lo:hi], **kw)

    def call_sliced(self, name, level, arr):
        for lo, hi, s in self.parts:
            getattr(s, name)(level, arr[lo:hi]) if level is not None else getattr(s, name)(arr[lo:hi])

    def solve(self):
        for _, _, s in self.parts:
            s.solve(sync=False)      # all devices run concurrently
        for _, _, s in self.parts:
            s.synchronize()

    def gather(self, name, *args):
        return np.concatenate([getattr(s, name)(*args) for _, _, s in self.parts], axis=0)


class PipelinedBatch:
    """One logical batch on ONE GPU, split into `chunks` sub-batches with their own handle and CUDA stream, each driven by
    its own host thread: the host->device copy of one chunk overlaps the solve and the result read of the previous ones.
    The copy phases are serialised by a lock (the PCIe link is the shared resource; interleaved copies would all finish
    late), the solves and reads are not.  For callers whose problem data lives in host memory (pinned for full speed)."""

    def __init__(self, cls, nVar, ObjDim, ObjType=None, batch=1, device=0, chunks=4, taper=True):
        import threading
        from concurrent.futures import ThreadPoolExecutor
        self.ObjType = list(ObjType) if ObjType is not None else None
        self.nObj = len(ObjDim)
        chunks = max(1, min(int(chunks), int(batch)))
        self.ranges = tapered_ranges(batch, chunks) if (taper and os.environ.get("LEXLS_B200_PIPE_TAPER", "1") != "0") else shard_ranges(batch, chunks)
        self.parts = []
        for lo, hi in self.ranges:
            kw = dict(batch=hi - lo, device=device, strict=False)
            self.parts.append((lo, hi, cls(nVar, ObjDim, ObjType, **kw) if ObjType is not None else cls(nVar, ObjDim, **kw)))
        for _, _, s in self.parts:
            s.set_async_copies(True)  # solve_host keeps the caller's arrays alive until every chunk has delivered its x
        self._lock = threading.Lock()
        self._pool = ThreadPoolExecutor(max_workers=len(self.parts))

    def close(self):
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
        for _, _, s in self.parts:
            s.close()
        self.parts = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def setParameters(self, p):
        for _, _, s in self.parts:
            s.setParameters(p)

    def solve_host(self, data, x_out, var_index=None, guess=None, x0=None):
        """data[k]: host arrays of the whole batch ([B, cols, m_k]); x_out: [B, nVar] host array that receives x.
        LexLSI warm start: guess[k] [B, m_k] uint8 and x0 [B, nVar].  Returns when every chunk is done."""

        def work(part):
            lo, hi, s = part
            with self._lock:  # one chunk's inputs at a time over the link
                for k, d in enumerate(data):
                    if self.ObjType is not None and self.ObjType[k] == 1:
                        s.setData(k, d[lo:hi], var_index=var_index)
                    else:
                        s.setData(k, d[lo:hi])
                if guess is not None:
                    for k, g in enumerate(guess):
                        s.set_active_guess(k, g[lo:hi])
                if x0 is not None:
                    s.set_x0(x0[lo:hi])
                (s.validate if os.environ.get("LEXLS_B200_PIPE_WAIT") == "validate" else s.wait_copies)()  # the one wait of this chunk: its copies are done before the next chunk gets the link (the
                #                  repacks are deferred to the solve below; bad inputs surface with get_x)
            s.solve(sync=False)
            s.get_x(out=x_out[lo:hi])

        for f in [self._pool.submit(work, p) for p in self.parts]:
            f.result()

    def gather(self, name, *args):
        return np.concatenate([getattr(s, name)(*args) for _, _, s in self.parts], axis=0)
