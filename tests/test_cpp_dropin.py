"""The drop-in C++ headers (include/lexls/*.h) compile against the C-ABI as a user of the reference would use them."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, name="test_dropin"):
    from lexls_b200 import build
    so = build.build()
    exe = os.path.join(str(tmp_path), name)
    cmd = ["g++", "-std=c++11", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", name + ".cpp"),
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


def test_batch_header_compiles_as_cxx11(tmp_path):
    """include/lexls/batch.h (LexLS::solveBatch, LexLSIBatch, LexLSEBatch) builds against the C-ABI; no GPU -> loud failure."""
    exe = _build(tmp_path, "test_batch")
    import torch
    if not torch.cuda.is_available():
        res = subprocess.run([exe], capture_output=True, text=True)
        assert res.returncode == 2 and "no CPU fallback" in res.stdout


@pytest.mark.gpu
def test_cpp_solve_batch_equals_single_problem_solves(tmp_path):
    exe = _build(tmp_path, "test_batch")
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0 and "OK" in res.stdout, res.stdout + res.stderr


def test_tools_header_and_numerical_error_replay_compile(tmp_path, golden_dir):
    """include/lexls/tools.h (HierarchyFileProcessor::import) and the get_xStar / get_data / get_lexqr surface build; without a
    GPU the run fails loudly after the import succeeded."""
    exe = _build(tmp_path, "test_numerical_error_replay")
    import torch
    if not torch.cuda.is_available():
        res = subprocess.run([exe, os.path.join(golden_dir, "test_01.dat")], capture_output=True, text=True)
        assert res.returncode == 2 and "no CPU fallback" in res.stdout


@pytest.mark.gpu
def test_numerical_error_replay_on_test_01_dat(tmp_path, golden_dir):
    """The reference's tests/test_numerical_error.cpp call sequence on its fixture tests/test_01.dat: LexLSI solve, get_xStar,
    get_data, get_lexqr, getActiveCtr_order, then LexLSE on the final working set — reproduced bit for bit (errors exactly 0)."""
    exe = _build(tmp_path, "test_numerical_error_replay")
    res = subprocess.run([exe, os.path.join(golden_dir, "test_01.dat")], capture_output=True, text=True)
    assert res.returncode == 0 and "OK" in res.stdout, res.stdout + res.stderr
    assert "error(x)     = 0.000000000000000000e+00" in res.stdout
