"""Aggregates an ncu SASS source page per CUDA source line (no GUI needed).

    python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> [<launch-id>] [--top N]

SASS rows of `ncu --page source` are matched, by instruction offset, with `nvdisasm -g` line annotations
of the cubins inside jagua-rs_b200/libjagua_cuda.so (compile with -lineinfo).
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "jagua-rs_b200", "libjagua_cuda.so")


def line_maps(kernel_mangled_sub):
    """{function: {offset: ((file, line), sass)}} for every function whose mangled name contains the substring
    (template instantiations are separate functions, each with offsets from 0)."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=tmp, capture_output=True)
    maps = {}
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin") or "sm_100" not in f:
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur_fn, cur_line = None, None
        for ln in txt.splitlines():
            s = ln.strip()
            if s.startswith(".text.") and s.endswith(":"):
                cur_fn = s[6:-1]
                cur_line = None
                continue
            g = re.match(r'//## File "(.*)", line (\d+)', s)
            if g:
                cur_line = (os.path.basename(g.group(1)), int(g.group(2)))
                continue
            g = re.match(r"/\*([0-9a-f]{4,})\*/\s+(.*)", s)
            if g and cur_fn and kernel_mangled_sub in cur_fn:
                maps.setdefault(cur_fn, {})[int(g.group(1), 16)] = (cur_line, g.group(2))
    return maps


def pick_map(maps, offsets_ops):
    """The instantiation that was profiled: same instruction count and the most matching opcodes."""
    best, best_score = {}, -1
    for fn, m in maps.items():
        if len(m) != len(offsets_ops):
            continue
        score = sum(1 for off, op in offsets_ops if off in m and m[off][1].replace("@", " ").split()[-0:1] and op in m[off][1])
        if score > best_score:
            best, best_score = m, score
    if best_score < 0 and maps:
        best = next(iter(maps.values()))
    return best


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    top = 40
    if "--top" in sys.argv:
        top = int(sys.argv[sys.argv.index("--top") + 1])
    args = ["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre]
    pos = [a for a in sys.argv[3:] if not a.startswith("--") and a.isdigit()]
    if pos and "--top" not in sys.argv[sys.argv.index(pos[0]) - 1:sys.argv.index(pos[0])]:
        args += ["--launch-skip", pos[0], "--launch-count", "1"]
    out = subprocess.run(args, capture_output=True, text=True).stdout
    blocks = out.split('"Kernel Name",')
    blk = blocks[1]
    name = blk.split("\n", 1)[0]
    rows = list(csv.reader(io.StringIO(blk.split("\n", 1)[1])))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    sub = re.search(r"(k_[a-z0-9_]+)", name).group(1)
    srows = [r for r in rows[1:] if len(r) >= len(hdr) and r[0].startswith("0x")]
    base0 = int(srows[0][0], 16)

    def opcode(text):
        toks = text.split()
        toks = [t for t in toks if not t.startswith("@")]
        return toks[0] if toks else ""

    lm = pick_map(line_maps(sub), [(int(r[0], 16) - base0, opcode(r[ix["Source"]])) for r in srows])
    base = None
    agg = defaultdict(lambda: [0.0, 0.0, 0.0])
    tot = [0.0, 0.0, 0.0]
    for r in rows[1:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        a = int(r[0], 16)
        base = a if base is None else base
        ie, te, sm = float(r[ix["Instructions Executed"]] or 0), float(r[ix["Thread Instructions Executed"]] or 0), float(r[ix["# Samples"]] or 0)
        loc = lm.get(a - base, ((None, 0), ""))[0] or ("?", 0)
        for k, v in enumerate((ie, te, sm)):
            agg[loc][k] += v
            tot[k] += v
    print(name[:120])
    print(f"total warp instr {tot[0]:.0f}, thread instr {tot[1]:.0f} (avg {tot[1] / max(tot[0], 1):.1f} thr/instr), samples {tot[2]:.0f}")
    src_cache = {}
    for loc, (ie, te, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        f, l = loc
        text = ""
        for d in ("jagua-rs_b200/csrc",):
            p = os.path.join(ROOT, d, f)
            if os.path.exists(p):
                if p not in src_cache:
                    src_cache[p] = open(p).read().splitlines()
                if 0 < l <= len(src_cache[p]):
                    text = src_cache[p][l - 1].strip()
        print(f"{100 * ie / tot[0]:5.1f}% instr {100 * sm / max(tot[2], 1):5.1f}% smpl {te / max(ie, 1):5.1f} thr | {f}:{l} | {text[:90]}")


if __name__ == "__main__":
    main()
