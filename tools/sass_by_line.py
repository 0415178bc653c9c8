"""Static SASS instruction count per source line (and per source function range) of one kernel: where does the code size go?

    python tools/sass_by_line.py <mangled-name substring> [--top N]
"""
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "jagua-rs_b200", "libjagua_cuda.so")
want = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=tmp, capture_output=True)
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin") or "sm_100" not in f:
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    fn, cur = None, None
    per = {}
    for ln in txt.splitlines():
        s = ln.strip()
        if s.startswith(".text.") and s.endswith(":"):
            fn = s[6:-1]
            cur = None
            continue
        g = re.match(r'//## File "(.*)", line (\d+)', s)
        if g:
            cur = (os.path.basename(g.group(1)), int(g.group(2)))
            continue
        if fn and want in fn and re.match(r"/\*[0-9a-f]{4,}\*/", s):
            per.setdefault(fn, Counter())[cur] += 1
    for fn, c in per.items():
        tot = sum(c.values())
        print(fn[:100], tot, "instructions")
        byfile = Counter()
        for (fl, l), n in c.items():
            byfile[fl] += n
        print("  by file:", dict(byfile))
        for (fl, l), n in c.most_common(top):
            print(f"  {n:5d}  {fl}:{l}")
