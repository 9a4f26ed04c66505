"""pytest plugin for the EMULATED run of the GPU tests (python -m pytest -p cudaemu_plugin -m gpu ..., PYTHONPATH=tests/cudaemu):
loads tests/cudaemu/_build/libzkb200emu.so - every source of libzkb200.so compiled for the host against the stand-in CUDA
runtime - in place of the product library, and makes torch's "cuda" tensors host tensors whose memory is registered with the
stand-in as device memory.  The tests then run unchanged: same C ABI calls, same comparisons with the oracle."""
import ctypes
import importlib.util
import os

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("cudaemu_build", os.path.join(HERE, "build.py"))
_build = importlib.util.module_from_spec(spec)
spec.loader.exec_module(_build)
_SO = _build.build_full()
os.environ["ZKB200_LIB"] = _SO
os.environ.setdefault("ZKB200_LANES", "2")
os.environ.setdefault("ZKB200_STAGE_THREADS", "1")

import torch  # noqa: E402

_lib = ctypes.CDLL(_SO, mode=ctypes.RTLD_GLOBAL)
_lib.emu_register.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
_DEVICE = 2
_keep = []          # registered buffers stay alive: a freed and reused address must not be taken for device memory


def _as_device(t):
    t = t.contiguous()
    if t.numel():
        _lib.emu_register(t.data_ptr(), t.numel() * t.element_size(), _DEVICE)
        _keep.append(t)
    return t


def _wrap_factory(fn):
    def f(*a, **kw):
        dev = kw.pop("device", None)
        t = fn(*a, **kw)
        return _as_device(t) if dev is not None and str(dev).startswith("cuda") else t
    return f


def _pinned(t):
    t = t.clone().contiguous()
    if t.numel():
        _lib.emu_register(t.data_ptr(), t.numel() * t.element_size(), 1)      # cudaMemoryTypeHost
        _keep.append(t)
    return t


torch.Tensor.pin_memory = lambda self, *a, **kw: _pinned(self)
torch.cuda.is_available = lambda: True
torch.cuda.device_count = lambda: 1
torch.Tensor.cuda = lambda self, *a, **kw: _as_device(self.clone())
for _name in ("zeros", "full", "empty", "ones"):
    setattr(torch, _name, _wrap_factory(getattr(torch, _name)))
_empty_like = torch.empty_like
torch.empty_like = lambda t, **kw: _as_device(_empty_like(t, **{k: v for k, v in kw.items() if k != "device"}))


def pytest_configure(config):
    # generated kernels are CUDA binaries: the emulated run uses the data-driven K3 / K5 kernels
    from ziren_b200 import _ffi
    _ffi.lib().zkb200_set_option(b"quotient_codegen", 0)
    _ffi.lib().zkb200_set_option(b"logup_codegen", 0)
