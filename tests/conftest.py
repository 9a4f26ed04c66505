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


# GPU run order (`pytest -m gpu -x`): the prover's own parity tests first, then the row fillers that have run bit-exact on a
# B200, then what was written after the round's GPU budget was spent and has only been checked on the host (DESIGN.md
# "f3 chip by chip", column "B200 run") - a first-run failure there must not cut the established tests short.
_GPU_ORDER = [("test_gpu_parity", "test_zz_cpp_host", "test_tracegen_keccak"),
              ("test_tracegen", "test_tracegen_mul", "test_zz_tracegen_mem_instr", "test_zz_tracegen_memory_local")]


def pytest_collection_modifyitems(config, items):
    def rank(item):
        mod = os.path.splitext(os.path.basename(str(item.fspath)))[0]
        for r, mods in enumerate(_GPU_ORDER):
            if mod in mods:
                return r
        return len(_GPU_ORDER)
    items.sort(key=rank)       # stable: file and definition order inside a rank
