"""Build-time step: generate and compile (NVRTC, sm_100a, no GPU needed) the per-chip constraint kernels
of the machines the tests and bench.py use, into the in-tree kernel cache next to libzkb200.so
(`ziren_b200/_kernel_cache/`, which travels with the library).  A chip that is not in the cache is simply
compiled the first time it is proved (csrc/quotient_codegen.cpp)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi, synthetic


def machines():
    yield "mini", synthetic.mini_case().machine
    yield "edge", synthetic.edge_case().machine
    yield "noprep", synthetic.noprep_case().machine
    yield "fibonacci", synthetic.fibonacci_core_case(log_cpu=8).machine
    yield "core", synthetic.core_case(log_cpu=9).machine
    yield "keccak", synthetic.keccak_case(log_cpu=8).machine
    yield "compress", synthetic.compress_case(log_max=9).machine
    # the nine narrow core chips with their restated Air::eval (tests/test_tracegen.py: shards proved from event records)
    from . import tracegen as tg
    narrow = {n: np.zeros((16, tg.width(n)), np.uint32) for n in ("AddSub", "ShiftLeft", "Lt", "ShiftRight", "Bitwise", "CloClz", "Branch", "Jump", "MovCond")}
    yield "alu", synthetic.alu_case(narrow, with_lookup_pair=True).machine
    # the real Global chip (synthetic.py _global_chip: septic-extension constraints of degree three)
    yield "global", synthetic.global_case(np.zeros((16, 99), np.uint32)).machine
    yield "div-rem", synthetic.chips_case({"DivRem": np.zeros((16, 106), np.uint32)}).machine
    yield "syscall-precompile", synthetic.syscall_precompile_case(np.zeros((16, 11), np.uint32), np.zeros((16, 99), np.uint32)).machine
    yield "syscall-instrs", synthetic.syscall_instrs_case(np.zeros((16, 77), np.uint32), np.zeros(8, np.uint32), np.zeros(8, np.uint32), 0).machine
    yield "memory-global", synthetic.memory_global_case(np.zeros((16, 111), np.uint32), np.zeros((16, 111), np.uint32),
                                                        np.zeros((16, 99), np.uint32), 0, 0,
                                                        syscall_rows=np.zeros((16, 11), np.uint32)).machine
    # the real KeccakSponge chip (ziren_b200/keccak_air.py: 3788 constraints, 357 lookups): minutes of NVRTC time once
    from . import keccak_sponge
    yield "keccak-real", synthetic.keccak_real_case(keccak_sponge.synthetic_blocks(1, 1), None, log_cpu=10).machine


def precompile(verbose: bool = False) -> int:
    total = 0
    for name, m in machines():
        desc = np.ascontiguousarray(m.descriptor(), dtype=np.uint32)
        n = C.c_size_t()
        rc = _ffi.lib().zkb200_codegen_compile_check(desc.ctypes.data_as(_ffi.u32p), desc.size, C.byref(n))
        if rc < 0:
            raise RuntimeError(f"constraint-kernel generation failed for machine {name}: "
                               + _ffi.lib().zkb200_last_error(None).decode())
        total += rc
        if verbose:
            print(f"  {name}: {rc} chips, {n.value} cubin bytes")
    return total


if __name__ == "__main__":
    print(precompile(verbose=True), "constraint kernels in the cache")
