"""Print the file:name:lo-hi arguments scripts/ncu_by_range.py takes, one range per device function of the kernel sources
(a function's range runs to the line before the next definition).  usage: python scripts/ncu_ranges.py"""
import os, re
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lexls_b200", "csrc")
out = []
for f in ("lexls_device.cuh", "lexls_panel.cuh"):
    lines = open(os.path.join(root, f)).read().splitlines()
    starts = []
    for i, ln in enumerate(lines, 1):
        if ln.startswith((" ", "\t", "//", "#", "}")):
            continue
        ms = re.match(r"(?:template\s*<[^>]*>\s*)?struct\s+(\w+)", ln)
        if ms and not ln.rstrip().endswith(";"):
            starts.append((i, ms.group(1)))
            continue
        if "(" not in ln:
            continue
        if "__device__" not in ln and "__global__" not in ln:
            continue
        m = re.search(r"([A-Za-z_][A-Za-z_0-9]*)\s*\(", ln.split("__launch_bounds__")[0] if "__global__" not in ln else
                      re.sub(r"__launch_bounds__\([^)]*\)(?:\))?", "", ln))
        if m:
            starts.append((i, m.group(1)))
    for k, (lo, name) in enumerate(starts):
        hi = (starts[k + 1][0] - 1) if k + 1 < len(starts) else len(lines)
        out.append("%s:%s:%d-%d" % (f, name, lo, hi))
print(" ".join(out))
