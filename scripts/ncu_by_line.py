"""Aggregate an `ncu --page source --csv` SASS dump by CUDA source line, using nvdisasm -g line info of the cubin.
usage: ncu_by_line.py <rep.ncu-rep> <kernel-substring e.g. 'lse_kernelILi1ELb0'> [top]"""
import csv, re, subprocess, sys, os, collections, tempfile
rep, ksub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lexls_b200", "liblexls_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
sass = "".join(subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
               for f in sorted(os.listdir(tmp)) if f.endswith(".cubin"))
# walk: function sections ".text.<mangled>" ; line markers '//## File "...", line N' ; instructions '/*0010*/'
addr2line = {}
cur_fn, cur_line = None, None
for ln in sass.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        cur_fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m and cur_fn and ksub in cur_fn:
        addr2line[int(m.group(1), 16)] = cur_line
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
isrc = hdr.index("Source")
base = None
agg = collections.defaultdict(lambda: [0, 0])
tot = [0, 0]
for r in rows[2:]:
    if len(r) <= isamp: continue
    try: a = int(r[ia], 16)
    except ValueError: continue
    if base is None: base = a
    key = addr2line.get(a - base, ("?", 0))
    n, s = int(float(r[ii] or 0)), int(float(r[isamp] or 0))
    agg[key][0] += n; agg[key][1] += s; tot[0] += n; tot[1] += s
print("total inst %d samples %d" % tuple(tot))
src = {}
for (f, l), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    if f not in src:
        p = os.path.join(os.path.dirname(so), "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][l - 1].strip() if 0 < l <= len(src[f]) else ""
    print("%5.1f%% samp %5.1f%% inst  %s:%d  %s" % (100.0 * s / max(tot[1], 1), 100.0 * n / max(tot[0], 1), f, l, text[:110]))
