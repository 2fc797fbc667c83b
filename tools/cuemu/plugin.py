"""cuemu — DEVELOPMENT TOOL (see README.md). A pytest plugin that points the pass-level parity tests of tests/ at the HOST
build of the kernel sources (tools/cuemu/_build/libdfx_b200_emu.so) with CPU tensors standing in for device memory:

    python -m pytest -p tools.cuemu.plugin tests/test_parity_gpu.py -m gpu -k "<pass-level tests>"

Only this plugin loads the emulated library; the package and the ordinary test runs never do.
"""
import inspect
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "_build", "libdfx_b200_emu.so")


def pytest_configure(config):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not os.path.exists(LIB):
        raise RuntimeError(f"{LIB} is missing: python tools/cuemu/build_emu.py")
    import ctypes as C

    import torch

    from diligentfx_b200 import capi
    capi._lib = None
    capi.load(LIB)                                                # every capi user of this process now talks to the host build
    # capi.plane_of insists on CUDA tensors; in the emulator device memory IS host memory
    src = inspect.getsource(capi.plane_of).replace("assert t.is_cuda and t.is_contiguous()", "assert t.is_contiguous()")
    ns = dict(vars(capi))
    exec(src, ns)
    capi.plane_of = ns["plane_of"]
    capi.pyramid_of.__globals__["plane_of"] = ns["plane_of"]

    torch.Tensor.cuda = lambda self, *a, **k: self.clone()        # the few direct .cuda() calls in the pass-level tests
    torch.cuda.synchronize = lambda *a, **k: None

    import helpers

    class HostDev(helpers.Dev):
        def up(self, a, dtype=None):
            t = torch.from_numpy(np.ascontiguousarray(a, np.float32 if dtype is None else dtype)).clone()
            self.keep.append(t)
            return t

        def mask(self, a):
            t = torch.from_numpy(np.ascontiguousarray(a != 0, np.uint8)).clone()
            self.keep.append(t)
            return t

        def empty(self, h, w, ch=1, fill=None, dtype=None):
            shape = (h, w) if ch == 1 else (h, w, ch)
            dt = dtype or torch.float32
            t = torch.zeros(shape, dtype=dt) if fill is None else torch.full(shape, fill, dtype=dt)
            self.keep.append(t)
            return t

        def cameras(self, curr, prev):
            buf = (capi.CameraAttribs * 2)(curr, prev)
            t = torch.frombuffer(bytearray(bytes(buf)), dtype=torch.uint8)
            self.keep.append(t)
            return C.c_void_p(t.data_ptr())

        def sync(self):
            pass

        @staticmethod
        def host(t):
            return t.detach().numpy().astype(np.float32)

    helpers.Dev = HostDev
