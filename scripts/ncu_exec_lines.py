"""Executed code footprint of a kernel by SOURCE LINE: distinct SASS instructions executed at least once, warp
instructions per problem, stall samples.   usage: ncu_exec_lines.py <rep.ncu-rep> <kernel-substring> <problems> [top] [outer]
(the .so in the tree must be the one the report was captured with; "outer" attributes inlined code to the outermost
statement — the line of the kernel body or of a non-inlined function that the code was expanded for)"""
import csv, re, subprocess, sys, os, collections, tempfile
rep, ksub, nprob = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
outer = len(sys.argv) > 5 and sys.argv[5] == "outer"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("LEXLS_B200_LIB") or os.path.join(root, "lexls_b200", "liblexls_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
addr2line = {}
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"): continue
    cur_fn = cur = None; fresh = True
    for ln in subprocess.run(["nvdisasm", "-gi" if outer else "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m: cur_fn = m.group(1); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue  # the last line of a chain is the outermost
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m and cur_fn and ksub in cur_fn: addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for r in rows[2:]:
    if len(r) <= isamp: continue
    try: a = int(r[ia], 16)
    except ValueError: continue
    if base is None: base = a
    key = addr2line.get(a - base)
    n = int(float(r[ii] or 0))
    agg[key][3] += 1
    if n == 0: continue
    agg[key][0] += 1; agg[key][1] += n; agg[key][2] += int(float(r[isamp] or 0))
tot = sum(v[0] for v in agg.values()); tn = sum(v[1] for v in agg.values()); ts = sum(v[2] for v in agg.values()); tall = sum(v[3] for v in agg.values())
print("executed footprint %.1f KB (%d of %d instr), %.0f warp-instr per problem" % (tot * 16 / 1024, tot, tall, tn / nprob))
src = {}
for key, (c, n, s, al) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, l = key if key else ("?", 0)
    if f not in src:
        p = os.path.join(root, "lexls_b200", "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
    print("%5d exec-instr %5.1f%%  %7.0f inst/prob  %4.1f%% samp  %s:%d  %s" % (c, 100.0 * c / tot, n / nprob, 100.0 * s / max(ts, 1), f, l, text))
