"""Static code size of one kernel by source line (nvdisasm -g line tables, inlined callee lines included).
usage: sass_lines.py <kernel-substring> [top]"""
import re, subprocess, sys, os, collections, tempfile
ksub = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("LEXLS_B200_LIB") or os.path.join(root, "lexls_b200", "liblexls_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cnt = collections.Counter(); total = 0
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"): continue
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_fn = cur = None
    for ln in sass.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m: cur_fn = m.group(1); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if cur_fn and ksub in cur_fn and re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            cnt[cur] += 1; total += 1
src = {}
print("kernel *%s*: %d SASS instructions = %.1f KB" % (ksub, total, total * 16 / 1024))
for (f, l), c in cnt.most_common(top):
    if f not in src:
        p = os.path.join(root, "lexls_b200", "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][l - 1].strip()[:100] if 0 < l <= len(src[f]) else ""
    print("%5d  %5.1f%%  %s:%d  %s" % (c, 100.0 * c / total, f, l, text))
