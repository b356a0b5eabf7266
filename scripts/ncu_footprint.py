"""Code footprint of a kernel's executed path by source function: distinct SASS instructions executed at least once
(x 16 B), and how often they run per problem.   usage: ncu_footprint.py <rep.ncu-rep> <kernel-substring> <problems>"""
import csv, re, subprocess, sys, os, collections, tempfile
rep, ksub, nprob = sys.argv[1], sys.argv[2], float(sys.argv[3])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "lexls_b200", "liblexls_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
sass = "".join(subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
               for f in sorted(os.listdir(tmp)) if f.endswith(".cubin"))
addr2line, cur_fn, cur_line = {}, None, None
for ln in sass.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        cur_fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m and cur_fn and ksub in cur_fn:
        addr2line[int(m.group(1), 16)] = cur_line
# function ranges from the sources: the same ranges scripts/ncu_by_range.py uses (scripts/ncu_ranges.py)
fn_of = {}
for item in subprocess.run([sys.executable, os.path.join(root, "scripts", "ncu_ranges.py")], capture_output=True, text=True).stdout.split():
    f, name, rng = item.split(":")
    lo, hi = rng.split("-")
    for i in range(int(lo), int(hi) + 1):
        fn_of[(f, i)] = name
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in rows[2:]:
    if len(r) <= isamp: continue
    try: a = int(r[ia], 16)
    except ValueError: continue
    if base is None: base = a
    n = int(float(r[ii] or 0))
    if n == 0: continue
    key = addr2line.get(a - base)
    fn = fn_of.get(key, "other:" + (key[0] if key else "?")) if key else "?"
    agg[fn][0] += 1; agg[fn][1] += n; agg[fn][2] += int(float(r[isamp] or 0))
tot = sum(v[0] for v in agg.values()); tn = sum(v[1] for v in agg.values()); ts = sum(v[2] for v in agg.values())
print("executed footprint %.1f KB (%d instr), %.0f warp-instr per problem" % (tot * 16 / 1024, tot, tn / nprob))
for fn, (c, n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-28s %6.1f KB  %5.1f%% of footprint  %8.0f inst/prob  reuse %5.1fx  %5.1f%% samples" % (fn, c * 16 / 1024, 100.0 * c / tot, n / nprob, n / nprob / c / 2, 100.0 * s / max(ts, 1)))
