"""The built library really contains what DESIGN.md 4.1 says it uses: TMA tile loads (UTMALDG) with mbarrier waits (SYNCS), thread-block-cluster
barriers (UCGABAR) and warp shuffles in the kernels named there - checked on the SASS of the in-tree .so (cuobjdump; CPU only). The table
this produces for a build is profiles/r2_sass_evidence.md."""
import os
import re
import shutil
import subprocess

import pytest

from diligentfx_b200 import build


@pytest.fixture(scope="module")
def sass(built):
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    text = subprocess.run([tool, "-sass", build.LIB], capture_output=True, text=True, check=True).stdout
    per, fn = {}, None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            per[fn] = []
        elif fn is not None:
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                per[fn].append(m.group(1))
    return per


def _count(per, kernel_substring, mnemonic_prefix, exclude=None):
    return {fn: sum(op.startswith(mnemonic_prefix) for op in ops) for fn, ops in per.items() if kernel_substring in fn and not (exclude and exclude in fn)}


def test_sm100a_only(sass):
    out = subprocess.run([shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump", "-lelf", build.LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_tma_tile_kernels_issue_bulk_tensor_loads(sass):
    tma = _count(sass, "pyramid_tile_kernel", "UTMALDG")
    with_tma = {fn: n for fn, n in tma.items() if "Lb1E" in fn}   # <Op, true>
    assert len(with_tma) == 3 and all(n >= 1 for n in with_tma.values()), with_tma
    assert all(n == 0 for fn, n in tma.items() if "Lb1E" not in fn)
    assert all(n >= 1 for n in _count(sass, "pyramid_tile_kernel", "SYNCS").values() if n) and sum(_count(sass, "pyramid_tile_kernel", "SYNCS").values()) >= 3
    spatial = _count(sass, "ssao_spatial_tile_kernel", "UTMALDG")
    assert spatial and all(n == 2 for n in spatial.values()), spatial   # depth + resampled AO windows


def test_cluster_kernels_use_the_cluster_barrier(sass):
    for kernel in ("bloom_tail_kernel", "pyramid_tail_kernel"):
        bars = _count(sass, kernel, "UCGABAR")
        assert bars and all(n >= 2 for n in bars.values()), (kernel, bars)   # arrive + wait


def test_streaming_kernels_are_shuffle_kernels(sass):
    for kernel in ("bloom_down2x_stream_kernel", "bloom_up2x_stream_kernel", "bloom_levels_kernel"):
        sh = _count(sass, kernel, "SHFL")
        assert sh and all(n >= 40 for n in sh.values()), (kernel, sh)
    for kernel in ("bloom_down2x_stream_kernel", "bloom_up2x_stream_kernel"):   # no shared memory, no CTA barrier: warps are independent
        assert all(n == 0 for n in _count(sass, kernel, "BAR").values()), kernel
