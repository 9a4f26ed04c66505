"""ctypes binding of libzkb200.so (the C ABI of include/zkb200.h).  There is deliberately no
fallback: if the CUDA library is missing or fails to load, importing the product path raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZKB200_LIB", os.path.join(_DIR, "libzkb200.so"))   # override: A/B builds of the same ABI

u32p = C.POINTER(C.c_uint32)


class Trace(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("height", C.c_size_t), ("width", C.c_size_t),
                ("flags", C.c_uint32), ("n_events", C.c_size_t)]


TRACE_COL_MAJOR, TRACE_EVENTS, TRACE_DERIVED = 1, 2, 4


class Table(C.Structure):
    """zkb200_table (include/zkb200.h): a resident table, device pointers, column-major Montgomery."""
    _fields_ = [("chip", C.c_char_p), ("prep", C.c_void_p), ("main_trace", C.c_void_p), ("height", C.c_size_t)]


def build(verbose: bool = False) -> None:
    """Compile every CUDA source for sm_100a into ziren_b200/libzkb200.so (in-tree)."""
    subprocess.run(["make", "-C", os.path.join(_DIR, "csrc"), "-j8"], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)


_lib = None

# name -> (restype, argtypes): every symbol include/zkb200.h declares
SIGNATURES = {
    "zkb200_ctx_create": (C.c_int, [C.c_int, u32p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "zkb200_ctx_create_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, u32p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "zkb200_ctx_num_devices": (C.c_int, [C.c_void_p]),
    "zkb200_set_option": (C.c_int, [C.c_char_p, C.c_long]),
    "zkb200_quotient_launch_counts": (None, [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
    "zkb200_codegen_compile_check": (C.c_int, [u32p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "zkb200_h2d_probe": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.POINTER(C.c_float)]),
    "zkb200_ctx_destroy": (None, [C.c_void_p]),
    "zkb200_last_error": (C.c_char_p, [C.c_void_p]),
    "zkb200_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "zkb200_setup": (C.c_int, [C.c_void_p, C.POINTER(Trace), C.c_int, C.c_uint32, u32p, u32p, C.POINTER(C.c_void_p)]),
    "zkb200_pk_free": (None, [C.c_void_p]),
    "zkb200_pk_initial_challenger": (C.c_int, [C.c_void_p, u32p]),
    "zkb200_commit": (C.c_int, [C.c_void_p, C.POINTER(Trace), C.c_int, u32p, C.c_size_t, u32p, C.POINTER(C.c_void_p)]),
    "zkb200_shard_free": (None, [C.c_void_p]),
    "zkb200_shard_device": (C.c_int, [C.c_void_p]),
    "zkb200_open": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, u32p, C.POINTER(u32p), C.POINTER(C.c_size_t)]),
    "zkb200_prove_shard": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Trace), C.c_int, u32p, C.c_size_t, u32p,
                                     C.POINTER(u32p), C.POINTER(C.c_size_t)]),
    "zkb200_free": (None, [C.c_void_p]),
    "zkb200_set_profile": (None, [C.c_void_p, C.c_int]),
    "zkb200_launch_count": (C.c_ulonglong, []),
    "zkb200_last_stage_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]),
    "zkb200_coset_lde": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_size_t, C.c_uint, C.c_uint32]),
    "zkb200_ntt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_size_t, C.c_int, C.c_int]),
    "zkb200_mmcs_root": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint), C.POINTER(C.c_size_t), C.c_int, u32p]),
    "zkb200_poseidon2_permute_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "zkb200_derive_multiplicities": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_ulonglong)]),
    "zkb200_permutation_trace": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_size_t, u32p, u32p, C.c_void_p, u32p]),
    "zkb200_quotient": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, u32p, u32p, u32p, u32p,
                                  u32p, u32p, C.c_size_t, C.c_void_p]),
    "zkb200_fri_fold": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, u32p, C.c_void_p, C.c_void_p]),
    "zkb200_grind": (C.c_int, [C.c_void_p, u32p, C.c_uint, u32p]),
    "zkb200_transpose": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int]),
    "zkb200_convert": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "zkb200_alu_trace_width": (C.c_int, [C.c_char_p]),
    "zkb200_generate_alu_trace": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.c_int]),
    "zkb200_keccak_sponge_trace_width": (C.c_int, []),
    "zkb200_generate_keccak_sponge_trace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.c_int]),
    "zkb200_sync": (C.c_int, [C.c_void_p]),
}


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback for the zkb200 hot path)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib
