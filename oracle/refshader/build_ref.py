"""REFERENCE-SHADER RUNNER — TEST INFRASTRUCTURE ONLY.

Compiles the reference's own PostProcess pixel shaders (HLSL, read IN PLACE from /root/reference/Shaders) as C++ against
oracle/refshader/hlsl.hpp and links them, each behind a small per-pass harness (oracle/refshader/harness/*.inc, this
repository's code), into oracle/_ref/librefshaders.so. The flattened, rewritten shader text only ever exists in memory: it
is piped to g++ on stdin, never written into the repository (oracle/_ref/ holds object files and the library, and is
git-ignored).

Textual rewrites applied to the HLSL (the language differences a header cannot absorb):
  * `#include "X"` resolved by base name inside the Shaders tree and inlined (each file once per translation unit)
  * `cbuffer Name { ... }` -> its member declarations (constant buffers become namespace-scope objects the harness fills)
  * semantics `: SV_Target0`, `: SV_Position`, ... removed; `[branch]`-style attributes removed
  * parameter qualifiers: `in T x` -> `T x`, `out T x` / `inout T x` -> `T& x`
  * unsuffixed floating literals get an `f` (HLSL literals are fp32)
  * `Texture2D name` -> `Texture2D<float4> name`; `__cplusplus` hidden so the shared structure headers take their HLSL side

Usage: python oracle/refshader/build_ref.py [--force] [--only PASS[,PASS...]] [--keep-going]
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.dirname(HERE)
OUT_DIR = os.path.join(ORACLE, "_ref")
LIB = os.path.join(OUT_DIR, "librefshaders.so")
REF_SHADERS = os.environ.get("DFX_REFERENCE_SHADERS", "/root/reference/Shaders")
REF_EXTRA = [os.path.join(os.path.dirname(REF_SHADERS), "Hydrogent", "shaders")]  # HnPostProcess.psh (the compose step)
CXXFLAGS = ["-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-fvisibility=hidden", "-pthread", "-w", f"-I{HERE}",
            f"-I{os.path.join(os.path.dirname(ORACLE), 'include')}"]

SEMANTICS = r"SV_\w+|NORMALIZED_XY|INSTANCE_ID"
_FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")


def _index() -> dict[str, str]:
    idx: dict[str, str] = {}
    for top in [REF_SHADERS, *REF_EXTRA]:
        for root, _, files in os.walk(top):
            for f in files:
                idx.setdefault(f, os.path.join(root, f))
    return idx


def rewrite(text: str) -> str:
    text = text.replace("\r\n", "\n")
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)   # comments first: qualifiers are matched across line ends
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r"cbuffer\s+\w+\s*\{(.*?)\}\s*;?", lambda m: m.group(1), text, flags=re.S)
    text = re.sub(r"\[(?:branch|flatten|loop|unroll(?:\s*\(\s*\w+\s*\))?|earlydepthstencil)\]", "", text)
    text = re.sub(r":\s*(?:" + SEMANTICS + r")\b", "", text)
    text = re.sub(r"(?<=[(,])(\s*)(?:in\s+)?(?:inout|out)\s+((?:const\s+)?[A-Za-z_]\w*\s+\w+\s*\[)", r"\1\2", text)  # arrays are passed by reference in C++ already
    text = re.sub(r"(?<=[(,])(\s*)(?:in\s+)?(?:inout|out)\s+((?:const\s+)?[A-Za-z_]\w*(?:<\w+>)?)\s+(?=[A-Za-z_])", r"\1\2& ", text)
    text = re.sub(r"(?<=[(,])(\s*)in\s+(?=[A-Za-z_])", r"\1", text)
    # the same qualifiers on a parameter that starts its own line (after an #if / #endif inside the parameter list)
    text = re.sub(r"(?m)^(\s*)(?:in\s+)?(?:inout|out)\s+((?:const\s+)?[A-Za-z_]\w*(?:<\w+>)?)\s+(?=[A-Za-z_]\w*\s*[,)])", r"\1\2& ", text)
    text = re.sub(r"(?m)^(\s*)in\s+(?=(?:const\s+)?[A-Za-z_]\w*(?:<\w+>)?\s+[A-Za-z_]\w*\s*[,)])", r"\1", text)
    text = re.sub(r"\b(Texture2D|TextureCube)\s+(?=[A-Za-z_])", r"\1<float4> ", text)
    text = text.replace("__cplusplus", "DFX_REFSH_HIDDEN_CPLUSPLUS")
    lines = []
    for line in text.split("\n"):
        if not line.lstrip().startswith(("#include", "#if", "#elif", "#line", "#error", "#pragma")):
            line = _FLOAT_LIT.sub(lambda m: m.group(1) + "f", line)
        lines.append(line)
    return "\n".join(lines)


def flatten(name: str, idx: dict[str, str], seen: set[str]) -> str:
    if name in seen:
        return f"// (already inlined: {name})\n"
    seen.add(name)
    path = idx.get(name)
    if path is None:
        raise FileNotFoundError(f"{name} not found under {REF_SHADERS}")
    out = [f"// ---- begin {os.path.relpath(path, os.path.dirname(REF_SHADERS))}\n"]
    for line in rewrite(open(path, encoding="utf-8", errors="replace").read()).split("\n"):
        m = re.match(r'\s*#\s*include\s+"([^"]+)"', line)
        out.append(flatten(os.path.basename(m.group(1)), idx, seen) if m else line + "\n")
    out.append(f"// ---- end {name}\n")
    return "".join(out)


def translation_unit(p: dict, idx: dict[str, str]) -> str:
    ns = "refsh_ns_" + p["name"]
    parts = ['#include "prelude.hpp"\n']
    parts += [f"#define {k} {v}\n" for k, v in p.get("macros", {}).items()]
    parts.append(f"namespace {ns} {{\nusing namespace hlsl;\n")
    parts.append(open(os.path.join(HERE, "core_defs.inc")).read())
    seen: set[str] = set()
    for shader in ([p["shader"]] if isinstance(p["shader"], str) else p["shader"]):  # several files: a function library + what its harness needs
        parts.append(flatten(shader, idx, seen))
    parts.append(open(os.path.join(HERE, "harness", "common.inc")).read())
    parts.append(f"#define REFSH_ENTRY refsh_{p['name']}\n")
    parts.append(open(os.path.join(HERE, "harness", p["harness"])).read())
    parts.append(f"\n}} // namespace {ns}\n")
    parts.append(f'extern "C" __attribute__((visibility("default"))) int refsh_{p["name"]}(const refsh_args* a) {{ return {ns}::entry(a); }}\n')
    return "".join(parts)


def passes() -> list[dict]:
    sys.path.insert(0, HERE)
    from manifest import PASSES
    return PASSES


def _compile(p: dict, idx: dict[str, str], force: bool) -> tuple[str, str, str]:
    tu = translation_unit(p, idx)
    digest = hashlib.sha256((tu + open(os.path.join(HERE, "hlsl.hpp")).read() + open(os.path.join(HERE, "prelude.hpp")).read()).encode()).hexdigest()
    obj = os.path.join(OUT_DIR, "obj", p["name"] + ".o")
    stamp = obj + ".sha"
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
        return p["name"], obj, ""
    r = subprocess.run(["g++", *CXXFLAGS, "-x", "c++", "-c", "-", "-o", obj], input=tu, text=True, capture_output=True)
    if r.returncode != 0:
        return p["name"], "", r.stderr
    open(stamp, "w").write(digest)
    return p["name"], obj, ""


def build(force: bool = False, only: list[str] | None = None, keep_going: bool = False, dump: str | None = None) -> str:
    if not os.path.isdir(REF_SHADERS):
        raise FileNotFoundError(f"{REF_SHADERS} is absent: librefshaders.so can only be (re)built where the reference is mounted")
    os.makedirs(os.path.join(OUT_DIR, "obj"), exist_ok=True)
    idx = _index()
    todo = [p for p in passes() if not only or p["name"] in only]
    if dump:
        sys.stdout.write(translation_unit(next(p for p in todo if p["name"] == dump), idx))
        return ""
    objs, failed = [], []
    with cf.ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 4)) as ex:
        for name, obj, err in ex.map(lambda p: _compile(p, idx, force), todo):
            if err:
                failed.append(name)
                sys.stderr.write(f"==== {name}: compile failed\n{err[:6000]}\n")
            else:
                objs.append(obj)
    if failed and not keep_going:
        raise RuntimeError("reference shaders that did not compile: " + ", ".join(failed))
    if not only:
        subprocess.check_call(["g++", "-shared", "-pthread", "-o", LIB, *objs])
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--keep-going", action="store_true")
    ap.add_argument("--dump", default="", help="print the translation unit of one pass (debugging; do not commit the output)")
    a = ap.parse_args()
    print(build(a.force, [s for s in a.only.split(",") if s], a.keep_going, a.dump or None))
