#!/usr/bin/env python3
"""Per-source-line instruction and stall-sample totals of one kernel in an ncu report (needs -lineinfo in the build).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep ssao_ao [--top 40] [--lib diligentfx_b200/lib/libdfx_b200.so]

ncu's CSV source page carries the counters per SASS instruction but no line numbers; `nvdisasm --print-line-info` of the
library's cubin carries the line of every instruction. Both list the kernel's instructions in the same order, so they
are zipped by position (the build must be the one that was profiled).
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def ncu_sass(rep: str, pattern: str):
    text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{pattern}"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    name = rows[0][1]
    hdr = rows[1]
    ie, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    out = []
    for r in rows[2:]:
        if len(r) <= ie or r[0] == "Kernel Name":
            break  # first matching launch only
        if r[0] == "Address":
            continue
        out.append((r[1].strip(), int(r[ie] or 0), int(r[ismp] or 0)))
    return name, out


def disasm_lines(lib: str, mangled_substr: str):
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], check=True, cwd=td, capture_output=True)
        for cubin in sorted(os.listdir(td)):
            text = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, cubin)], capture_output=True, text=True).stdout
            lines, cur, inside = [], None, False
            for ln in text.splitlines():
                m = re.match(r"\s*\.text\.(\S+):", ln)
                if m:
                    inside = mangled_substr in m.group(1)
                    continue
                if not inside:
                    continue
                m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
                if m:
                    cur = (os.path.basename(m.group(1)), int(m.group(2)))
                    continue
                if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
                    lines.append(cur)
            if lines:
                return lines
    return []


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("kernel", help="regex for ncu -k; also the substring looked up in the mangled names")
    ap.add_argument("--mangled", help="substring of the mangled kernel name if it differs from the regex")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--lib", default="diligentfx_b200/lib/libdfx_b200.so")
    a = ap.parse_args()
    name, sass = ncu_sass(a.rep, a.kernel)
    where = disasm_lines(a.lib, a.mangled or a.kernel)
    if len(where) != len(sass):
        print(f"warning: {len(sass)} profiled instructions vs {len(where)} in {a.lib}: not the profiled build?", file=sys.stderr)
    inst, smp = collections.Counter(), collections.Counter()
    for (txt, e, s), w in zip(sass, where):
        inst[w] += e
        smp[w] += s
    ti, ts = sum(inst.values()) or 1, sum(smp.values()) or 1
    print(f"{name[:100]}\n{ti} warp instructions, {ts} samples\n inst%  smp%  line")
    src_cache = {}
    for w, e in inst.most_common(a.top):
        text = ""
        if w:
            path = os.path.join("diligentfx_b200/csrc", w[0])
            if path not in src_cache and os.path.exists(path):
                src_cache[path] = open(path).read().splitlines()
            if path in src_cache and w[1] - 1 < len(src_cache[path]):
                text = src_cache[path][w[1] - 1].strip()[:110]
        print(f"{e / ti * 100:5.1f} {smp[w] / ts * 100:5.1f}  {w[0] if w else '?'}:{w[1] if w else 0}  {text}")
