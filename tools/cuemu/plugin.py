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
LIB = os.path.join(HERE, "_build", "libdfx_b200_emu_asan.so" if os.environ.get("CUEMU_ASAN") else "libdfx_b200_emu.so")


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
    alias = os.path.join(HERE, "_build", "libdfx_b200.so")         # tests/test_cpp_shim.py links its C++ driver with -ldfx_b200 from dirname(capi.LIB_PATH)
    if os.path.islink(alias) or os.path.exists(alias):
        os.remove(alias)
    os.symlink(os.path.basename(LIB), alias)
    capi.LIB_PATH = alias
    # capi.plane_of insists on CUDA tensors; in the emulator device memory IS host memory
    # CUEMU_PITCH_PAD=N: the tests' planes get N extra (poisoned) texels per row, i.e. a row pitch larger than width * texel
    # size, the way an application's own allocations may be laid out (the tests themselves only use tight planes)
    pad = int(os.environ.get("CUEMU_PITCH_PAD", "0"))
    src = inspect.getsource(capi.plane_of).replace("assert t.is_cuda and t.is_contiguous()", "assert t.stride(1) == (1 if t.dim() == 2 else t.shape[2])")
    src = src.replace("return Plane(t.data_ptr(), w * bpp, w, h, f, flags)", "return Plane(t.data_ptr(), t.stride(0) * t.element_size(), w, h, f, flags)")
    assert "t.stride(0) * t.element_size()" in src
    ns = dict(vars(capi))
    exec(src, ns)
    capi.plane_of = ns["plane_of"]
    capi.pyramid_of.__globals__["plane_of"] = ns["plane_of"]

    def _padded(t):
        if not pad:
            return t
        store = torch.full((t.shape[0], t.shape[1] + pad) + tuple(t.shape[2:]), 7 if t.dtype == torch.uint8 else 12345.0, dtype=t.dtype)
        view = store[:, :t.shape[1]]
        view.copy_(t)
        return view

    torch.Tensor.cuda = lambda self, *a, **k: self.clone()        # the few direct .cuda() calls in the pass-level tests
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    torch.Tensor.cpu = lambda self, *a, **k: self.clone()         # a device -> host copy is a copy: tests keep the result while the plane is reused
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.is_available = lambda: True

    def _on_cpu(fn):                                              # factory calls written with device="cuda"
        def wrapped(*a, **k):
            dev = k.get("device")
            if dev is not None and (str(dev).startswith("cuda")):
                k["device"] = "cpu"
            return fn(*a, **k)
        return wrapped

    _to = torch.Tensor.to

    def _to_cpu(self, *a, **k):                                   # .to(torch.device("cuda", n)): device memory is host memory here
        a = tuple("cpu" if (isinstance(x, (str, torch.device)) and str(x).startswith("cuda")) else x for x in a)
        if "device" in k and str(k["device"]).startswith("cuda"):
            k["device"] = "cpu"
        out = _to(self, *a, **k)
        return out.clone() if out is self else out

    torch.Tensor.to = _to_cpu

    for _name in ("empty", "zeros", "ones", "full", "tensor", "rand", "randn", "arange", "empty_like", "zeros_like", "full_like"):
        setattr(torch, _name, _on_cpu(getattr(torch, _name)))
    torch.cuda.current_device = lambda: 0
    torch.cuda.set_device = lambda *a, **k: None

    # The chain driver (diligentfx_b200/chain.py) sequences its passes over CUDA streams and events; the host build executes
    # every call synchronously, so streams and events are inert objects and the chain's planes live on the CPU device.
    class _Event:
        def __init__(self, *a, **k):
            self._t = 0.0

        def record(self, *a, **k):
            import time
            self._t = time.perf_counter()

        def wait(self, *a, **k):
            pass

        def synchronize(self):
            pass

        def query(self):
            return True

        def elapsed_time(self, other):
            return max(1e-3, (other._t - self._t) * 1e3)          # host wall clock: enough for code that divides by it

    class _Stream:
        cuda_stream = 0

        def __init__(self, *a, **k):
            pass

        def wait_event(self, e):
            pass

        def wait_stream(self, s):
            pass

        def record_event(self, e=None):
            return e or _Event()

        def synchronize(self):
            pass

    class _StreamCtx:
        def __init__(self, s):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    class _DeviceCtx:                                             # torch.cuda.device(dev)
        def __init__(self, *a, **k):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    torch.cuda.device = _DeviceCtx
    torch.cuda.can_device_access_peer = lambda a, b: bool(os.environ.get("CUEMU_IPC"))
    _as_tensor = torch.as_tensor

    def _as_tensor_host(obj, *a, **k):                            # strips.PeerSlab views its own allocation through __cuda_array_interface__
        iface = getattr(obj, "__cuda_array_interface__", None)
        if iface is not None:
            return torch.frombuffer((C.c_uint8 * iface["shape"][0]).from_address(iface["data"][0]), dtype=torch.uint8)
        if "device" in k and str(k["device"]).startswith("cuda"):
            k["device"] = "cpu"
        return _as_tensor(obj, *a, **k)

    torch.as_tensor = _as_tensor_host
    _main = _Stream()
    torch.cuda.Event, torch.cuda.Stream, torch.cuda.stream = _Event, _Stream, _StreamCtx
    torch.cuda.current_stream = lambda *a, **k: _main
    from diligentfx_b200 import chain as chain_mod
    _init = chain_mod.PostProcessChain.__init__

    def _cpu_init(self, width, height, config=None, device=None, **kw):
        _init(self, width, height, config, device=torch.device("cpu"), **kw)

    chain_mod.PostProcessChain.__init__ = _cpu_init

    # CUEMU_SPLIT_ROWS=K: every pass-level call that takes a row range runs as two calls, [y0, K) and [K, y1) - what a
    # row-strip decomposition does (K must be a multiple of 64, the strip granularity of DESIGN.md section 7)
    split = int(os.environ.get("CUEMU_SPLIT_ROWS", "0"))
    if split:
        real = capi._lib

        class _SplitRows:
            def __getattr__(self, name):
                fn = getattr(real, name)
                if not name.startswith("dfx_pass_"):
                    return fn

                def call(*args):
                    if args and isinstance(args[-1], capi.Rows) and args[-1].y0 < split < args[-1].y1:
                        r = args[-1]
                        st = fn(*args[:-1], capi.Rows(r.y0, split))
                        return st if st != 0 else fn(*args[:-1], capi.Rows(split, r.y1))
                    return fn(*args)
                return call

        capi._lib = _SplitRows()

    import helpers

    class HostDev(helpers.Dev):
        def up(self, a, dtype=None):
            t = _padded(torch.from_numpy(np.ascontiguousarray(a, np.float32 if dtype is None else dtype)).clone())
            self.keep.append(t)
            return t

        def mask(self, a):
            t = _padded(torch.from_numpy(np.ascontiguousarray(a != 0, np.uint8)).clone())
            self.keep.append(t)
            return t

        def empty(self, h, w, ch=1, fill=None, dtype=None):
            shape = (h, w) if ch == 1 else (h, w, ch)
            dt = dtype or torch.float32
            t = _padded(torch.zeros(shape, dtype=dt) if fill is None else torch.full(shape, fill, dtype=dt))
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
