"""Opcode histogram of one (sub)function in an nvdisasm dump: python scripts/sass_fn.py dump.sass <substring> [--dump]"""
import re, sys, collections
txt = open(sys.argv[1]).read().splitlines()
sub = sys.argv[2]
start = None
out = []
for i, ln in enumerate(txt):
    m = re.match(r'^(\S+):\s*$', ln)
    if m and not ln.startswith('.L'):
        if start is not None:
            break
        if sub in m.group(1) and m.group(1).startswith('$'):
            start = i
        continue
    if start is not None:
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(.*?);', ln)
        if m:
            out.append(m.group(1).strip())
        elif re.match(r'^\.L_x_\d+:', ln) and '--dump' in sys.argv:
            out.append(ln.strip())
if '--dump' in sys.argv:
    print("\n".join(out)); sys.exit()
h = collections.Counter()
for o in out:
    o = re.sub(r'^@!?U?P\d+\s+', '', o)
    h[o.split()[0].split('.')[0]] += 1
print(len(out), "instructions")
for k, v in h.most_common(40):
    print("%6d %s" % (v, k))
