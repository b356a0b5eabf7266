"""Aggregate ncu SASS stats into phases given by source line ranges of lexls_device.cuh (ranges passed as name:lo-hi ...)."""
import csv, re, subprocess, sys, os, collections, tempfile
rep, ksub = sys.argv[1], sys.argv[2]
phases = []
for a in sys.argv[3:]:
    name, rng = a.split(":"); lo, hi = rng.split("-"); phases.append((name, int(lo), int(hi)))
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lexls_b200", "liblexls_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
sass = "".join(subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
               for f in sorted(os.listdir(tmp)) if f.endswith(".cubin"))
addr2line, cur_fn, cur = {}, None, None
for ln in sass.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m: cur_fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        f, l, rest = os.path.basename(m.group(1)), int(m.group(2)), m.group(3)
        # prefer the outermost location inside lexls_device.cuh (inlined-at chain)
        locs = [(f, l)] + [(os.path.basename(a), int(b)) for a, b in re.findall(r'inlined at "([^"]+)", line (\d+)', rest)]
        cur = locs; continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S+)", ln)
    if m and cur_fn and ksub in cur_fn:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None; agg = collections.defaultdict(lambda: [0, 0]); tot = [0, 0]; ops = collections.defaultdict(lambda: collections.Counter())
for r in rows[2:]:
    try: a = int(r[ia], 16)
    except Exception: continue
    if base is None: base = a
    locs, op = addr2line.get(a - base, ([("?", 0)], "?"))
    n, s = int(float(r[ii] or 0)), int(float(r[isamp] or 0))
    ph = "other"
    for f, l in locs:
        if f == "lexls_device.cuh":
            for name, lo, hi in phases:
                if lo <= l <= hi: ph = name; break
            if ph != "other": break
    agg[ph][0] += n; agg[ph][1] += s; tot[0] += n; tot[1] += s
    ops[ph][op.split(".")[0]] += n
print("total inst %d samples %d" % tuple(tot))
for ph, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    top = ", ".join("%s %.0f%%" % (o, 100.0 * c / max(n, 1)) for o, c in ops[ph].most_common(7))
    print("%-10s %5.1f%% samp %5.1f%% inst   [%s]" % (ph, 100.0 * s / tot[1], 100.0 * n / tot[0], top))
