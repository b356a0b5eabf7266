"""Developer probe: builds variants of the library that differ in compile-time switches of ONE translation unit (the other
objects come from the regular build) into lexls_b200/variants/<name>.so; LEXLS_B200_LIB=<path> makes capi.py load one.

    python scripts/build_variants.py kernels_lse.cu cq=-DLX_LSE_COMPACT_QR=true "cqg=-DLX_LSE_COMPACT_QR=true -DLX_LSE_COMPACT_GS=true"
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lexls_b200 import build as B


def main():
    unit = sys.argv[1]
    variants = dict(a.split("=", 1) for a in sys.argv[2:])
    B.build()
    out = os.path.join(B.HERE, "variants")
    os.makedirs(out, exist_ok=True)
    others = [os.path.join(B.OBJDIR, u + ".o") for u in B.UNITS if u != unit]

    def one(item):
        name, flags = item
        obj = os.path.join(out, name + "_" + unit + ".o")
        cmd = ["nvcc"] + B.NVCC_FLAGS + flags.split() + ["-c", "-o", obj, os.path.join(B.HERE, "csrc", unit)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            return name, r.stderr[-3000:]
        so = os.path.join(out, name + ".so")
        r2 = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", so, obj] + others,
                            capture_output=True, text=True)
        os.remove(obj)
        regs = [l for l in r.stderr.splitlines() if "registers" in l]
        return name, (r2.stderr[-2000:] if r2.returncode else "ok " + so + " | " + " / ".join(x.split("Used ")[1].split(",")[0] for x in regs[:6]))

    with ThreadPoolExecutor(max_workers=4) as ex:
        for name, msg in ex.map(one, variants.items()):
            print(name, msg)


if __name__ == "__main__":
    main()
