"""The drop-in C++ headers (include/lexls/*.h) compile against the C-ABI as a user of the reference would use them."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    from lexls_b200 import build
    so = build.build()
    exe = os.path.join(str(tmp_path), "test_dropin")
    cmd = ["g++", "-std=c++11", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp"),
           "-o", exe, so, "-Wl,-rpath," + os.path.dirname(so)]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_dropin_headers_compile_as_cxx11(tmp_path):
    exe = _build(tmp_path)
    import torch
    if not torch.cuda.is_available():
        # without a GPU the facade must fail loudly (LexLS::Exception), not fall back to a CPU path
        res = subprocess.run([exe], capture_output=True, text=True)
        assert res.returncode == 2 and "no CPU fallback" in res.stdout


@pytest.mark.gpu
def test_dropin_known_answers_on_gpu(tmp_path):
    exe = _build(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0 and "OK" in res.stdout, res.stdout + res.stderr
