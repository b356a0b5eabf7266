"""Aggregate ncu SASS stats (instructions executed, pc samples) into named source ranges.
usage: python scripts/ncu_by_range.py report.ncu-rep <kernel substring> file:name:lo-hi ...   (innermost inlined location is used)"""
import csv, re, subprocess, sys, os, collections, tempfile
rep, ksub = sys.argv[1], sys.argv[2]
ranges = []
for a in sys.argv[3:]:
    f, name, rng = a.split(":"); lo, hi = rng.split("-"); ranges.append((f, name, int(lo), int(hi)))
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lexls_b200", "liblexls_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
sass = "".join(subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
               for f in sorted(os.listdir(tmp)) if f.endswith(".cubin"))
addr2loc, cur_fn, cur = {}, None, None
for ln in sass.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m: cur_fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S+)", ln)
    if m and cur_fn and ksub in cur_fn:
        op = m.group(2)
        if op.startswith("@"):
            op = ln.split("*/")[1].split()[1]
        addr2loc[int(m.group(1), 16)] = (cur, op)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None; agg = collections.defaultdict(lambda: [0, 0]); tot = [0, 0]; ops = collections.defaultdict(collections.Counter)
for r in rows[2:]:
    try: a = int(r[ia], 16)
    except Exception: continue
    if base is None: base = a
    loc, op = addr2loc.get(a - base, (("?", 0), "?"))
    n, s = int(float(r[ii] or 0)), int(float(r[isamp] or 0))
    ph = "other:%s" % loc[0]
    for f, name, lo, hi in ranges:
        if loc[0] == f and lo <= loc[1] <= hi: ph = name; break
    agg[ph][0] += n; agg[ph][1] += s; tot[0] += n; tot[1] += s
    ops[ph][op.split(".")[0]] += n
nprob = float(os.environ.get("NPROB", "1"))
print("total inst %d samples %d  (per problem: %.0f inst)" % (tot[0], tot[1], tot[0] / nprob))
for ph, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    top = ", ".join("%s %.0f%%" % (o, 100.0 * c / max(n, 1)) for o, c in ops[ph].most_common(8))
    print("%-14s %5.1f%% samp %5.1f%% inst %8.0f inst/prob  [%s]" % (ph, 100.0 * s / tot[1], 100.0 * n / tot[0], n / nprob, top))
