"""SASS instruction count per kernel of libjagua_cuda.so (the grid kernels are bound by instruction fetch once their hot
loops outgrow the instruction cache: keep an eye on this after every kernel change).

    python tools/sass_sizes.py [substring]
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "jagua-rs_b200", "libjagua_cuda.so")
want = sys.argv[1] if len(sys.argv) > 1 else ""
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=tmp, capture_output=True)
rows = []
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin") or "sm_100" not in f:
        continue
    txt = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    name, cnt = None, {}
    for ln in txt.splitlines():
        s = ln.strip()
        if s.startswith(".text.") and s.endswith(":"):
            name = s[6:-1]
            continue
        if name and re.match(r"/\*[0-9a-f]{4,}\*/", s):
            cnt[name] = cnt.get(name, 0) + 1
    rows += [(c, n) for n, c in cnt.items()]
for c, n in sorted(rows):
    if want in n:
        d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
        print(f"{c:6d}  {c * 16 / 1024:6.1f} KB  {d[:110]}")
