import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_ffi
    oracle_ffi.lib()
    return oracle_ffi


@pytest.fixture(scope="session")
def host():
    """The product's arithmetic / row-filler headers (ziren_b200/csrc/*.cuh) compiled for the host by
    tests/hostcheck/hostcheck.cpp: the CUDA kernels run exactly these expressions."""
    import ctypes
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    src = os.path.join(here, "hostcheck", "hostcheck.cpp")
    out_dir = os.path.join(here, "hostcheck", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhostcheck.so")
    csrc = os.path.join(ROOT, "ziren_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("kb31.cuh", "poseidon2.cuh", "tracegen.cuh", "tracegen_keccak.cuh", "tracegen_global.cuh", "derive.cuh", "machine_dev.h", "lane_pool.h", "p2_rc.inc")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I" + csrc, "-x", "c++", src, "-o", so])
    return ctypes.CDLL(so)
