"""An emulated device for the CPU-only build box: the product's Python layer (pyro2_b200: Pyro, the solvers, the
multigrid classes, the handles) runs unchanged, but every C-ABI call it makes lands in the host-compiled kernel
libraries of tests/emu/ (the same .cu sources built with g++ against the CUDA execution-model emulator) and its
"device" tensors live in host memory.

TEST INFRASTRUCTURE ONLY.  It exists so that the bodies of the `-m gpu` tests -- public API in, reference fixtures
out -- can be rehearsed here, where there is no GPU: host-side logic (parameter plumbing, boundary hooks, time-step
control, problem setups) and kernel indexing are then both checked before a GPU run is spent on them.  Nothing in
pyro2_b200/ imports this file; the product has no CPU path and fails loudly without a CUDA device.

    with emu_device.emulated_device():
        p = Pyro("compressible"); p.initialize_problem(...); p.single_step()

What is patched while the context is active (and restored afterwards):
  * pyro2_b200._lib.lib()         -> a facade that routes each p2b_* symbol to the emulator library built from the
                                     .cu file that defines it (the sweep goes to the warp emulator of sweep_task.cuh)
  * pyro2_b200._lib.stream_ptr()  -> NULL (the emulator runs launches synchronously)
  * ops.require_cuda()            -> no-op; the default device of grids -> cpu
  * torch.zeros / torch.empty / Tensor.to / Tensor.cuda: a "cuda" device argument means host memory;
    Tensor.is_cuda -> True; torch.cuda.synchronize -> no-op (CUDA-graph capture then fails and the callers fall back
    to eager launches, which is their documented behaviour)
"""
import contextlib
import ctypes as C
from unittest import mock

import torch

import emu_util

RESIDENT_WARPS = 148 * 12   # the device's warp slots as choose_seglen() sees them; small grids then split into segments


class EmuLibrary:
    """stands in for the ctypes handle of libpyro2b200.so"""
    ROUTES = (("p2b_mg_", emu_util.load_mg_emu), ("p2b_shared_", emu_util.load_mg_emu), ("p2b_flow_", emu_util.load_flow_emu), ("p2b_lm_", emu_util.load_lm_emu),
              ("p2b_fill_hse", emu_util.load_bc_emu), ("p2b_fill_ambient", emu_util.load_bc_emu),
              ("p2b_fill_ghost", emu_util.load_ghost_emu), ("p2b_cfl_wavemax", emu_util.load_ghost_emu),
              ("p2b_device_sms", emu_util.load_ghost_emu), ("p2b_slab_", emu_util.load_ghost_emu))

    is_emulated = True

    def __init__(self):
        self.calls = {}          # symbol -> number of calls, for the tests to see what really ran
        self._used = []
        self._info = (0, 0, 0)

    def _counted(self, name, f):
        def call(*a):
            self.calls[name] = self.calls.get(name, 0) + 1
            rec = getattr(_CAPTURE, "graph", None)
            if rec is not None and name not in ("p2b_last_error", "p2b_mg_result"):
                # stream capture: the launch is recorded with its argument VALUES (as a CUDA graph freezes kernel
                # arguments) and runs only when the graph is replayed
                rec.calls.append((f, a))
                return 0
            return f(*a)
        return call

    def __getattr__(self, name):
        if name == "p2b_compressible_sweep":
            return self._counted(name, self._sweep)
        if name == "p2b_sweep_info":
            return self._sweep_info
        if name == "p2b_sweep_uses_tensor_map":
            return lambda: 0
        if name == "p2b_test_fastmath":
            lib = emu_util.load_sweep_emu()
            return lambda op, a, b, out, n, stream: lib.emu_test_fastmath(op, a, b, out, n)
        if name == "p2b_last_error":
            return lambda: b"; ".join([lib.p2b_last_error() or b"" for lib in self._used] + [getattr(self, "_sweep_error", b"")])
        if name == "p2b_version":
            return lambda: 1
        for prefix, loader in self.ROUTES:
            if name.startswith(prefix):
                lib = loader()
                if lib not in self._used:
                    self._used.append(lib)
                return self._counted(name, getattr(lib, name))
        raise AttributeError(name)

    def _sweep(self, uin, uout, g_ref, prm_ref, dt, scratch, stream):   # pylint: disable=unused-argument
        """p2b_compressible_sweep over the warp emulator: the argument checks, the copy into the kernel's argument
        struct and the work decomposition are the device launch's own (csrc/sweep_args.cuh, compiled into the
        emulator library); the scratch words come back the same way (wave-speed maxima in [0], [1], status in [3])"""
        lib = emu_util.load_sweep_emu()
        why = C.c_char_p()
        rc = lib.emu_compressible_sweep_abi(uin, uout, g_ref, prm_ref, dt, scratch, RESIDENT_WARPS, C.byref(why))
        if rc:
            self._sweep_error = why.value or b"sweep refused"
            return -1
        ntasks, seglen = C.c_int(), C.c_int()
        lib.emu_sweep_decomposition(g_ref, RESIDENT_WARPS, C.byref(ntasks), C.byref(seglen))
        self._info = (ntasks.value, RESIDENT_WARPS, seglen.value)
        return 0

    def _sweep_info(self, a, b, c):
        for ref, val in zip((a, b, c), self._info):
            ref._obj.value = val
        return 0


def _is_cuda(dev):
    return dev is not None and (dev == "cuda" or (isinstance(dev, str) and dev.startswith("cuda"))
                                or (isinstance(dev, torch.device) and dev.type == "cuda"))


def _host_device(fn):
    def wrapped(*a, **k):
        if _is_cuda(k.get("device")):
            k["device"] = "cpu"
        return fn(*a, **k)
    return wrapped


import threading
_CAPTURE = threading.local()


class EmuGraph:
    """stand-in for torch.cuda.CUDAGraph on the emulated device: library calls made inside `with torch.cuda.graph(g)` are
    recorded, not executed; g.replay() executes them with the recorded argument values.  Catches what a real graph
    would: a replayed cycle that depends on a by-value argument which a later call meant to change."""

    def __init__(self):
        self.calls = []

    def replay(self):
        for f, a in self.calls:
            rc = f(*a)
            assert not rc, rc


@contextlib.contextmanager
def _emu_graph_capture(g, *a, **k):
    assert getattr(_CAPTURE, "graph", None) is None
    _CAPTURE.graph = g
    try:
        yield
    finally:
        _CAPTURE.graph = None


class EmuStream:
    """stand-in for torch.cuda.Stream / Event on the emulated device: the emulated kernels run synchronously, so every
    ordering call is a no-op"""

    def __init__(self, *a, **k):
        pass

    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass

    def record(self, stream=None):
        pass

    def synchronize(self):
        pass

    cuda_stream = None


def _host_tensor_from_pointer(ptr, nelem):
    """ops.tensor_from_pointer over host ("device") memory"""
    return torch.frombuffer((C.c_double * nelem).from_address(ptr), dtype=torch.float64)


@contextlib.contextmanager
def emulated_device():
    """run the enclosed product code against the emulator libraries; yields the EmuLibrary (see .calls)"""
    from pyro2_b200 import _lib, mg_handle, ops
    from pyro2_b200.mesh import patch
    facade = EmuLibrary()
    real_to = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple("cpu" if _is_cuda(x) else x for x in a)
        if _is_cuda(k.get("device")):
            k["device"] = "cpu"
        return real_to(self, *a, **k)
    with contextlib.ExitStack() as st:
        for target, attr, new in (
                (_lib, "lib", lambda: facade), (_lib, "stream_ptr", lambda: None),
                (ops, "require_cuda", lambda: None), (mg_handle, "require_cuda", lambda: None),
                (ops, "tensor_from_pointer", _host_tensor_from_pointer),
                (patch, "_default_device", lambda: torch.device("cpu")),
                (torch, "zeros", _host_device(torch.zeros)), (torch, "empty", _host_device(torch.empty)),
                (torch, "as_tensor", _host_device(torch.as_tensor)),
                (torch.Tensor, "to", to), (torch.Tensor, "cuda", lambda self, *a, **k: self),
                (torch.cuda, "synchronize", lambda *a, **k: None),
                (torch.cuda, "CUDAGraph", EmuGraph), (torch.cuda, "graph", _emu_graph_capture),
                (torch.cuda, "Stream", EmuStream), (torch.cuda, "Event", EmuStream),
                (torch.cuda, "current_stream", lambda *a, **k: EmuStream()),
                (torch.cuda, "stream", lambda s: contextlib.nullcontext()),
                (torch.Tensor, "is_pinned", lambda self, *a, **k: True)):
            st.enter_context(mock.patch.object(target, attr, new))
        st.enter_context(mock.patch.object(torch.Tensor, "is_cuda", property(lambda self: True)))
        yield facade
