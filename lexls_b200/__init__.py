"""lexls_b200 — B200-native batched lexicographic least-squares (LexLSE / LexLSI).

The compute path is hand-written CUDA for sm_100a behind a C-ABI (include/lexls_b200.h, lexls_b200/csrc/).
This package is the thin host side: a ctypes binding, batch front-ends mirroring the reference's LexLSE / LexLSI
call sequence, synthetic workload generators and the `.dat` fixture reader.  There is no CPU fallback.
"""
from .capi import LexLSError, Params, SO_PATH  # noqa: F401
from .batch import LexLSEBatch, LexLSIBatch, measure_fp64_peak  # noqa: F401
from . import datfile, problems  # noqa: F401
from .shard import PipelinedBatch, ShardedBatch  # noqa: F401

__all__ = ["LexLSEBatch", "LexLSIBatch", "PipelinedBatch", "ShardedBatch", "Params", "LexLSError", "measure_fp64_peak", "datfile", "problems", "SO_PATH"]
