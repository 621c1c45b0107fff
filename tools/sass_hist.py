"""Opcode histogram per kernel of a built library (cuobjdump -sass): python tools/sass_hist.py lib.so [kernel-regex] [top]"""
import collections, re, subprocess, sys
lib = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().replace("kvae::", "")
        fn = re.sub(r"\(.*", "", fn)
        hist[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        hist[fn][m.group(1)] += 1
for fn, h in hist.items():
    if pat and not pat.search(fn):
        continue
    tot = sum(h.values())
    print(f"{fn}: {tot} instructions; " + ", ".join(f"{k} {v}" for k, v in h.most_common(top)))
