#!/usr/bin/env python3
"""cuemu — DEVELOPMENT TOOL, not part of the product (see README.md).

Compiles diligentfx_b200/csrc/*.cu for the HOST against tools/cuemu/include (a stand-in for the slice of CUDA the kernels
use) into tools/cuemu/_build/libdfx_b200_emu.so, which exports the same C-ABI as libdfx_b200.so. Two textual rewrites are
applied to copies of the sources under _build/src/ (git-ignored):
  * `kernel<<<grid, block, smem, stream>>>(args)`  ->  `::cuemu::launch(grid, block, [&]() { kernel(args); })`
  * the three inline-PTX MUFU forms of dfx_common.cuh  ->  calls of their host equivalents

    python tools/cuemu/build_emu.py [-D NAME=VALUE ...]     # extra macros reach the kernels exactly like build.py's
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "diligentfx_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libdfx_b200_emu.so")

_ASM = re.compile(r'asm\("(\w+)\.approx\.ftz\.f32 %0, %1;"\s*:\s*"=f"\((\w+)\)\s*:\s*"f"\((\w+)\)\);')


def _balanced(text: str, start: int, open_ch: str = "(", close_ch: str = ")") -> int:
    """index just past the bracket that closes the one at text[start]"""
    depth = 0
    for i in range(start, len(text)):
        if text[i] == open_ch:
            depth += 1
        elif text[i] == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced brackets")


def rewrite_launches(text: str) -> str:
    out, pos = [], 0
    while True:
        k = text.find("<<<", pos)
        if k < 0:
            out.append(text[pos:])
            return "".join(out)
        # kernel name (with optional template arguments) ends right before <<<
        j = k
        if text[j - 1] == ">":                                   # template arguments: walk back to the matching <
            depth, j = 0, k - 1
            while True:
                depth += {">": 1, "<": -1}.get(text[j], 0)
                if depth == 0:
                    break
                j -= 1
        m = re.search(r"[\w:]+$", text[pos:j])
        name_start = pos + m.start()
        kernel = text[name_start:k]
        e = text.find(">>>", k)
        cfg = text[k + 3:e]
        a0 = text.index("(", e)
        a1 = _balanced(text, a0)
        parts, depth, cur = [], 0, ""
        for ch in cfg:                                           # split the launch configuration at top-level commas
            depth += {"(": 1, ")": -1}.get(ch, 0)
            if ch == "," and depth == 0:
                parts.append(cur)
                cur = ""
            else:
                cur += ch
        parts.append(cur)
        out.append(text[pos:name_start])
        out.append(f"::cuemu::launch(dim3({parts[0].strip()}), dim3({parts[1].strip()}), [&]() {{ {kernel}{text[a0:a1]}; }})")
        pos = a1


def rewrite(text: str) -> str:
    text = _ASM.sub(lambda m: f"{m.group(2)} = ::cuemu::{m.group(1)}_approx_ftz({m.group(3)});", text)
    text = text.replace('#include "../../include/dfx_b200.h"', f'#include "{os.path.join(ROOT, "include", "dfx_b200.h")}"')
    text = text.replace('#include "_gen/', f'#include "{CSRC}/_gen/')
    return rewrite_launches(text)


def build(defines: list[str] | None = None, asan: bool = False) -> str:
    sys.path.insert(0, ROOT)
    from diligentfx_b200 import build as product_build
    product_build.build()                                        # generates csrc/_gen (the blue-noise tables) as a side effect
    src_dir = os.path.join(OUT, "src")
    os.makedirs(src_dir, exist_ok=True)
    for f in os.listdir(CSRC):
        if f.endswith((".cu", ".cuh")):
            open(os.path.join(src_dir, f.replace(".cu", ".cpp") if f.endswith(".cu") else f), "w").write(rewrite(open(os.path.join(CSRC, f)).read()))
    # the TMA / mbarrier primitives are inline PTX: the host build uses its own statement of what they do
    open(os.path.join(src_dir, "dfx_tma.cuh"), "w").write(open(os.path.join(HERE, "include", "dfx_tma_emu.cuh")).read())
    srcs = [os.path.join(src_dir, f.replace(".cu", ".cpp")) for f in product_build.SOURCES]
    out = LIB.replace(".so", "_asan.so") if asan else LIB      # --asan: every plane access of every kernel checked by AddressSanitizer
    cmd = ["g++", "-std=c++17", "-O1" if asan else "-O2", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-w", f"-I{os.path.join(HERE, 'include')}",
           *(["-fsanitize=address", "-fno-omit-frame-pointer", "-g"] if asan else []),
           *[f"-D{d}" for d in (defines or [])], *srcs, os.path.join(HERE, "cuemu_runtime.cpp"), "-o", out]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    print(build(defs, asan="--asan" in sys.argv[1:]))
