"""Static code size of one kernel attributed to the OUTERMOST source line (nvdisasm -gi inline chains): which statement of
the kernel body (or of a non-inlined function) the code was expanded for.   usage: sass_callers.py <obj-or-so> <kernel-substring> [top]"""
import re, subprocess, sys, os, collections, tempfile
path, ksub = os.path.abspath(sys.argv[1]), sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", path], cwd=tmp, capture_output=True)
outer = collections.Counter(); total = 0
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"): continue
    cur_fn = None; chain = []; fresh = True
    for ln in subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m: cur_fn = m.group(1); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            if fresh: chain = []; fresh = False
            chain.append((os.path.basename(m.group(1)), int(m.group(2)))); continue
        if cur_fn and ksub in cur_fn and re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            fresh = True
            if chain: outer[chain[-1]] += 1
            total += 1
src = {}
print("kernel *%s*: %d SASS instructions = %.1f KB" % (ksub, total, total * 16 / 1024))
for (f, l), c in outer.most_common(top):
    if f not in src:
        p = os.path.join(root, "lexls_b200", "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][l - 1].strip()[:110] if 0 < l <= len(src[f]) else ""
    print("%5d  %5.1f%%  %s:%d  %s" % (c, 100.0 * c / total, f, l, text))
