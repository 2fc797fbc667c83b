"""Keeps tools/cuemu (the development tool that compiles the kernel SOURCES for the host, tools/cuemu/README.md) working: the host
build must compile from the current csrc/*.cu, and a handful of the `-m gpu` parity tests must pass on it through the plugin,
in a separate process. This is not a CPU path of the product - the package never loads that library, and this test does not
replace any `-m gpu` run; it makes sure the tool is there when a kernel edit needs checking on a machine without a GPU."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_build_of_the_kernel_sources_passes_parity_tests(built):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "cuemu", "build_emu.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    env = dict(os.environ, CUEMU_THREADS=str(min(32, os.cpu_count() or 1)))
    env.pop("DFX_LIB", None)
    cmd = [sys.executable, "-m", "pytest", "-p", "tools.cuemu.plugin", "-q", "-x", "-m", "gpu", os.path.join(ROOT, "tests", "test_parity_gpu.py"),
           os.path.join(ROOT, "tests", "cuemu_extra_check.py"), "-k", "256x144 and (ssr_intersect or ssr_spatial or taa or bloom or ssao_ambient) or bilateral_filter"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1800, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-500:]
