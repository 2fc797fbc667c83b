"""cuemu — DEVELOPMENT TOOL. Dry-runs bench.py's GPU arm on the host build (tiny frame, numbers meaningless): catches Python-level
breakage of the bench (argument handling, JSON keys, pipeline sequencing) on a machine without a GPU.

    python tools/cuemu/run_bench_emu.py --width 160 --height 96 --steps 2 --warmup 1
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tools.cuemu import plugin  # noqa: E402

plugin.pytest_configure(None)
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[1:]
runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")
