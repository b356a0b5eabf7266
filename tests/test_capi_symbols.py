"""CPU: the C-ABI shared library loads and exports every symbol include/lexls_b200.h declares (no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lexls_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lexls_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from lexls_b200 import build, capi
    build.build()
    lib = capi.lib()
    declared = _declared()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), "missing export: " + name
    assert sorted(capi.SYMBOLS) == declared, "capi.SYMBOLS and the header disagree"
    assert lib.lexls_b200_abi_version() == 3


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import lexls_b200 as lx
    with pytest.raises(lx.LexLSError) as e:
        lx.LexLSEBatch(4, [2, 2], batch=1)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under lexls_b200/ or include/ may import, include or link it."""
    bad = []
    for base in ("lexls_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"(^|\W)(import oracle|from oracle|oracle/|liblexls_oracle|lexls_oracle\.hpp)", txt):
                        # comments that merely mention the oracle's file name are fine; code references are not
                        for line in txt.splitlines():
                            s = line.strip()
                            if re.search(r"(import oracle|from oracle|liblexls_oracle|#include.*oracle)", s) and not s.startswith(("//", "#", "*")):
                                bad.append((f, s))
    assert not bad, bad


def test_header_is_plain_c99(tmp_path):
    """include/lexls_b200.h is the FFI boundary: it must compile as C (no C++ types, no torch types)."""
    import subprocess
    src = tmp_path / "hc.c"
    src.write_text('#include "lexls_b200.h"\nint main(void) { lexls_b200_params_t p; lexls_b200_batch_t *h = 0; (void)p; (void)h; return 0; }\n')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"), "-fsyntax-only", str(src)],
                   check=True, capture_output=True, text=True)
