"""ctypes bindings of the CPU oracle (oracle/libzkoracle.so) and of the reference-header shim
(oracle/_ref/libzkref.so).  TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs.  Canonical-form uint32 everywhere."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
u32p = C.POINTER(C.c_uint32)


def build(quiet: bool = True):
    subprocess.run(["make", "-C", _DIR], check=True, stdout=subprocess.DEVNULL if quiet else None)


def _load(path):
    if not os.path.exists(path):
        build()
    return C.CDLL(path)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load(os.path.join(_DIR, "libzkoracle.so"))
        _lib.zko_last_error.restype = C.c_char_p
        _lib.zko_machine_new.restype = C.c_void_p
        _lib.zko_setup.restype = C.c_void_p
        _lib.zko_two_adic_generator.restype = C.c_uint32
        _lib.zko_grind.restype = C.c_uint32
    return _lib


def ref_lib():
    """The reference's own kb31_t / Poseidon2 C++ headers behind a C shim; None if not built."""
    p = os.path.join(_DIR, "_ref", "libzkref.so")
    if not os.path.exists(p):
        return None
    l = C.CDLL(p)
    for f in ("ref_kb31_mul", "ref_kb31_add", "ref_kb31_sub", "ref_kb31_inv", "ref_kb31_to_monty", "ref_kb31_from_monty"):
        getattr(l, f).restype = C.c_uint32
        getattr(l, f).argtypes = [C.c_uint32] * (1 if f.endswith(("inv", "monty")) else 2)
    return l


def ref_core_lib():
    """The reference's own ALU row fillers (crates/core/machine/include/*.hpp) behind a C shim; None if not built."""
    p = os.path.join(_DIR, "_ref", "libzkref_core.so")
    if not os.path.exists(p):
        return None
    l = C.CDLL(p)
    l.ref_alu_num_cols.restype = C.c_uint
    return l


ALU_CHIPS = ("AddSub", "Bitwise", "Lt", "ShiftLeft", "ShiftRight", "CloClz", "Branch", "Jump", "MovCond")


def alu_width(chip):
    return lib().zko_alu_width(ALU_CHIPS.index(chip))


def alu_trace(chip, events, height):
    """events: (n, 7) uint32, {pc, next_pc, opcode, hi, a, b, c} (AluEvent) or {pc, next_pc, next_next_pc, opcode,
    a, b, c} (Branch / Jump); returns (height, width) canonical rows."""
    ev = _a(events).reshape(-1, 7)
    w = alu_width(chip)
    out = np.zeros((int(height), w), np.uint32)
    if lib().zko_alu_trace(ALU_CHIPS.index(chip), _p(ev), C.c_size_t(ev.shape[0]), C.c_size_t(int(height)), _p(out)):
        raise RuntimeError(err())
    return out


def ref_alu_rows(chip, events):
    """Rows of the reference's own C++ event_to_row (Montgomery words), or None without oracle/_ref."""
    l = ref_core_lib()
    if l is None:
        return None
    ev = _a(events).reshape(-1, 7)
    cid = ALU_CHIPS.index(chip)
    out = np.zeros((ev.shape[0], l.ref_alu_num_cols(cid)), np.uint32)
    if l.ref_alu_event_to_rows(cid, _p(ev), C.c_size_t(ev.shape[0]), _p(out)):
        raise RuntimeError("reference row filler failed")
    return out


MUL_WIDTH, COMP_EVENT_WORDS = 58, 16


def mul_trace(events, height):
    """events: (n, 16) uint32 CompAluEvent records; (height, 58) canonical rows of the Mul chip."""
    ev = _a(events).reshape(-1, COMP_EVENT_WORDS)
    out = np.zeros((int(height), MUL_WIDTH), np.uint32)
    if lib().zko_mul_trace(_p(ev), C.c_size_t(ev.shape[0]), C.c_size_t(int(height)), _p(out)):
        raise RuntimeError(err())
    return out


def ref_mul_rows(events):
    """Rows of the reference's own mul.hpp event_to_row (Montgomery words), or None without oracle/_ref."""
    l = ref_core_lib()
    if l is None or not hasattr(l, "ref_mul_event_to_rows"):
        return None
    ev = _a(events).reshape(-1, COMP_EVENT_WORDS)
    out = np.zeros((ev.shape[0], MUL_WIDTH), np.uint32)
    if l.ref_mul_event_to_rows(_p(ev), C.c_size_t(ev.shape[0]), _p(out)):
        raise RuntimeError("reference row filler failed")
    return out


MEMINSTR_WIDTH = 79


def mem_instr_trace(events, height):
    """events: (n, 16) uint32 MemInstrEvent records; (height, 79) canonical rows of the MemoryInstrs chip."""
    ev = _a(events).reshape(-1, COMP_EVENT_WORDS)
    out = np.zeros((int(height), MEMINSTR_WIDTH), np.uint32)
    if lib().zko_mem_instr_trace(_p(ev), C.c_size_t(ev.shape[0]), C.c_size_t(int(height)), _p(out)):
        raise RuntimeError(err())
    return out


def ref_mem_instr_rows(events):
    """Rows of the reference's own memory_instrs.hpp event_to_row (Montgomery words), or None without oracle/_ref."""
    l = ref_core_lib()
    if l is None or not hasattr(l, "ref_mem_instr_event_to_rows"):
        return None
    ev = _a(events).reshape(-1, COMP_EVENT_WORDS)
    out = np.zeros((ev.shape[0], MEMINSTR_WIDTH), np.uint32)
    if l.ref_mem_instr_event_to_rows(_p(ev), C.c_size_t(ev.shape[0]), _p(out)):
        raise RuntimeError("reference row filler failed")
    return out


MEMLOCAL_WIDTH, MEMLOCAL_ENTRY_WIDTH = 56, 14


def memory_local_trace(events, height):
    """events: (n, 7) uint32 MemoryLocalEvent records, four per row; (height, 56) canonical rows of the MemoryLocal chip."""
    ev = _a(events).reshape(-1, 7)
    out = np.zeros((int(height), MEMLOCAL_WIDTH), np.uint32)
    if lib().zko_memory_local_trace(_p(ev), C.c_size_t(ev.shape[0]), C.c_size_t(int(height)), _p(out)):
        raise RuntimeError(err())
    return out


def ref_memory_local_entries(events):
    """One SingleMemoryLocal (14 Montgomery words) per event from the reference's own memory_local.hpp, or None without
    oracle/_ref."""
    l = ref_core_lib()
    if l is None or not hasattr(l, "ref_memory_local_entries"):
        return None
    ev = _a(events).reshape(-1, 7)
    out = np.zeros((ev.shape[0], MEMLOCAL_ENTRY_WIDTH), np.uint32)
    if l.ref_memory_local_entries(_p(ev), C.c_size_t(ev.shape[0]), _p(out)):
        raise RuntimeError("reference row filler failed")
    return out


CPU_WIDTH, CPU_EVENT_WORDS = 67, 28


def cpu_trace(events, height):
    """events: (n, 28) uint32 zkb200_cpu_event records; (height, 67) canonical rows of the Cpu chip."""
    ev = _a(events).reshape(-1, CPU_EVENT_WORDS)
    out = np.zeros((int(height), CPU_WIDTH), np.uint32)
    if lib().zko_cpu_trace(_p(ev), C.c_size_t(ev.shape[0]), C.c_size_t(int(height)), _p(out)):
        raise RuntimeError(err())
    return out


def ref_cpu_rows(events):
    """Rows of the reference's own cpu.hpp event_to_row (Montgomery words), or None without oracle/_ref."""
    l = ref_core_lib()
    if l is None or not hasattr(l, "ref_cpu_event_to_rows"):
        return None
    ev = _a(events).reshape(-1, CPU_EVENT_WORDS)
    out = np.zeros((ev.shape[0], CPU_WIDTH), np.uint32)
    if l.ref_cpu_event_to_rows(_p(ev), C.c_size_t(ev.shape[0]), _p(out)):
        raise RuntimeError("reference row filler failed")
    return out


MISC_WIDTH, MISC_EVENT_WORDS = 72, 15


def misc_trace(events, height):
    """events: (n, 15) uint32 MiscEvent records; (height, 72) canonical rows of the MiscInstrs chip."""
    ev = _a(events).reshape(-1, MISC_EVENT_WORDS)
    out = np.zeros((int(height), MISC_WIDTH), np.uint32)
    if lib().zko_misc_trace(_p(ev), C.c_size_t(ev.shape[0]), C.c_size_t(int(height)), _p(out)):
        raise RuntimeError(err())
    return out


def ref_misc_rows(events):
    """Rows of the reference's own misc_instrs.hpp event_to_row (Montgomery words), or None without oracle/_ref."""
    l = ref_core_lib()
    if l is None or not hasattr(l, "ref_misc_event_to_rows"):
        return None
    ev = _a(events).reshape(-1, MISC_EVENT_WORDS)
    out = np.zeros((ev.shape[0], MISC_WIDTH), np.uint32)
    if l.ref_misc_event_to_rows(_p(ev), C.c_size_t(ev.shape[0]), _p(out)):
        raise RuntimeError("reference row filler failed")
    return out


ROW_CHIPS = ("DivRem", "SyscallCore", "SyscallPrecompile", "SyscallInstrs")
MEMGLOBAL_WIDTH = 111


def chip_trace_width(chip):
    return lib().zko_chip_trace_width(chip.encode())


def chip_event_words(chip):
    return lib().zko_chip_event_words(chip.encode())


def chip_trace(chip, events, height):
    """DivRem (CompAluEvent, 16 words) / SyscallCore, SyscallPrecompile, SyscallInstrs (SyscallEvent, 14 words): (height, width)
    canonical rows, zero padding rows."""
    w, ew = chip_trace_width(chip), chip_event_words(chip)
    if w < 0:
        raise ValueError(f"oracle: no row filler for chip {chip}")
    ev = _a(events).reshape(-1, ew)
    out = np.zeros((int(height), w), np.uint32)
    if lib().zko_chip_trace(chip.encode(), _p(ev), C.c_size_t(ev.shape[0]), C.c_size_t(int(height)), _p(out)):
        raise RuntimeError(err())
    return out


def memory_global_trace(events, previous_addr, height):
    """events: (n, 4) uint32 address-sorted MemoryInitializeFinalizeEvent records {addr, value, shard, timestamp};
    previous_addr: the public values' previous_init / previous_finalize address; (height, 111) canonical rows."""
    ev = _a(events).reshape(-1, 4)
    out = np.zeros((int(height), MEMGLOBAL_WIDTH), np.uint32)
    if lib().zko_memory_global_trace(_p(ev), C.c_size_t(ev.shape[0]), C.c_uint32(int(previous_addr)), C.c_size_t(int(height)), _p(out)):
        raise RuntimeError(err())
    return out


def ref_chip_rows(chip, events, event_words):
    """Rows of the reference's own div_rem.hpp / syscall.hpp / syscall_instrs.hpp / memory_global.hpp event_to_row (Montgomery
    words), or None without oracle/_ref."""
    l = ref_core_lib()
    if l is None or not hasattr(l, "ref_chip_event_to_rows"):
        return None
    l.ref_chip_num_cols.restype = C.c_uint
    w = l.ref_chip_num_cols(chip.encode())
    ev = _a(events).reshape(-1, event_words)
    out = np.zeros((ev.shape[0], w), np.uint32)
    if l.ref_chip_event_to_rows(chip.encode(), _p(ev), C.c_size_t(ev.shape[0]), _p(out)):
        raise RuntimeError("reference row filler failed")
    return out


GLOBAL_WIDTH, GLOBAL_EVENT_WORDS = 99, 8


def global_trace(events, height):
    """events: (n, 8) uint32 GlobalLookupEvent records {message[7], is_receive | kind << 8}; (height, 99) canonical rows of the
    Global chip."""
    ev = _a(events).reshape(-1, GLOBAL_EVENT_WORDS)
    out = np.zeros((int(height), GLOBAL_WIDTH), np.uint32)
    if lib().zko_global_trace(_p(ev), C.c_size_t(ev.shape[0]), C.c_size_t(int(height)), _p(out)):
        raise RuntimeError(err())
    return out


SEPTIC_OPS = {"mul": 0, "inv": 1, "sqrt": 2, "frobenius": 3, "double_frobenius": 4, "curve_formula": 5, "curve_add": 6}


def _septic_call(fn, op, a, b):
    a = _a(a)
    b = _a(b) if b is not None else None
    out = np.zeros(14 if op == "curve_add" else 7, np.uint32)
    rc = fn(SEPTIC_OPS[op], _p(a), _p(b) if b is not None else None, _p(out))
    if rc < 0:
        raise RuntimeError("septic op failed")
    return None if rc else out


def septic_op(op, a, b=None):
    """F_p^7 / septic-curve primitives of oracle/tracegen_global.h on canonical words; None for the root of a non-square."""
    return _septic_call(lib().zko_septic_op, op, a, b)


def ref_septic_op(op, a, b=None):
    """The same through the reference's kb31_septic_extension_t.hpp; raises LookupError without oracle/_ref."""
    l = ref_core_lib()
    if l is None or not hasattr(l, "ref_septic_op"):
        raise LookupError("oracle/_ref not built")
    return _septic_call(l.ref_septic_op, op, a, b)


KS_WIDTH, KS_REC_WORDS = 3531, 384


def keccak_sponge_trace(blocks, height):
    """blocks: (n, 384) uint32 block records (include/zkb200.h zkb200_keccak_block); (height, 3531) canonical rows."""
    b = _a(blocks).reshape(-1, KS_REC_WORDS)
    out = np.zeros((int(height), KS_WIDTH), np.uint32)
    if lib().zko_keccak_sponge_trace(_p(b), C.c_size_t(b.shape[0]), C.c_size_t(int(height)), _p(out)):
        raise RuntimeError(err())
    return out


def mem_access(value, shard, ts, prev_shard, prev_ts):
    out = np.zeros(9, np.uint32)
    lib().zko_mem_access(C.c_uint32(value), C.c_uint32(shard), C.c_uint32(ts), C.c_uint32(prev_shard), C.c_uint32(prev_ts), _p(out))
    return out


def ref_mem_access(value, shard, ts, prev_shard, prev_ts):
    """MemoryAccessCols of the reference's own memory.hpp populate_read (Montgomery words), or None without oracle/_ref."""
    l = ref_core_lib()
    if l is None or not hasattr(l, "ref_mem_read_cols"):
        return None
    out = np.zeros(9, np.uint32)
    l.ref_mem_read_cols(C.c_uint32(value), C.c_uint32(shard), C.c_uint32(ts), C.c_uint32(prev_shard), C.c_uint32(prev_ts), _p(out))
    return out


def _a(x):
    return np.ascontiguousarray(x, dtype=np.uint32)


def _p(x):
    return x.ctypes.data_as(u32p)


def err():
    return lib().zko_last_error().decode()


def num_threads():
    return lib().zko_num_threads()


def set_num_threads(n):
    lib().zko_set_num_threads(int(n))


def permute(state):
    s = _a(state).copy()
    lib().zko_poseidon2_permute(_p(s))
    return s


def permute_batch(states):
    s = _a(states).copy()
    lib().zko_poseidon2_permute_batch(_p(s), C.c_size_t(s.size // 16))
    return s


def hash(x):
    x = _a(x)
    out = np.zeros(8, np.uint32)
    lib().zko_hash(_p(x), C.c_size_t(x.size), _p(out))
    return out


def compress(l, r):
    l, r = _a(l), _a(r)
    out = np.zeros(8, np.uint32)
    lib().zko_compress(_p(l), _p(r), _p(out))
    return out


def ef_mul(a, b):
    a, b = _a(a), _a(b)
    out = np.zeros(4, np.uint32)
    lib().zko_ef_mul(_p(a), _p(b), _p(out))
    return out


def ef_inv(a):
    a = _a(a)
    out = np.zeros(4, np.uint32)
    lib().zko_ef_inv(_p(a), _p(out))
    return out


def two_adic_generator(bits):
    return lib().zko_two_adic_generator(C.c_uint(bits))


def dft(mat, inverse=False):
    m = _a(mat).copy()
    lib().zko_dft(_p(m), C.c_size_t(m.shape[0]), C.c_size_t(m.shape[1]), C.c_int(int(inverse)))
    return m


def coset_lde(mat, added_bits=1, shift=3):
    m = _a(mat)
    out = np.zeros((m.shape[0] << added_bits, m.shape[1]), np.uint32)
    lib().zko_coset_lde(_p(m), C.c_size_t(m.shape[0]), C.c_size_t(m.shape[1]), C.c_uint(added_bits), C.c_uint32(shift), _p(out))
    return out


def _mat_args(mats):
    mats = [_a(m) for m in mats]
    n = len(mats)
    ptrs = (u32p * n)(*[_p(m) for m in mats])
    hs = (C.c_size_t * n)(*[m.shape[0] for m in mats])
    ws = (C.c_size_t * n)(*[m.shape[1] for m in mats])
    return mats, n, ptrs, hs, ws


def mmcs_root(mats):
    mats, n, ptrs, hs, ws = _mat_args(mats)
    out = np.zeros(8, np.uint32)
    lib().zko_mmcs_root(C.c_int(n), ptrs, hs, ws, _p(out))
    return out


def pcs_commit_root(mats, shifts=None, log_blowup=1):
    mats, n, ptrs, hs, ws = _mat_args(mats)
    out = np.zeros(8, np.uint32)
    sh = _a(shifts) if shifts is not None else None
    lib().zko_pcs_commit_root(C.c_int(n), ptrs, hs, ws, _p(sh) if sh is not None else None, C.c_uint(log_blowup), _p(out))
    return out


def fri_fold(vals, beta, ro_next=None):
    v = _a(vals)
    m = v.size // 4
    out = np.zeros((m // 2, 4), np.uint32)
    b = _a(beta)
    r = _a(ro_next) if ro_next is not None else None
    lib().zko_fri_fold(_p(v), C.c_size_t(m), _p(b), _p(r) if r is not None else None, _p(out))
    return out


def challenger_script(state34, ops, vals):
    st, ops, vals = _a(state34).copy(), _a(ops), _a(vals)
    out = np.zeros(int((ops != 0).sum()) + 1, np.uint32)
    lib().zko_challenger_script(_p(st), _p(ops), _p(vals), C.c_size_t(ops.size), _p(out))
    return st, out[:-1]


def grind(state34, bits):
    st = _a(state34).copy()
    w = lib().zko_grind(_p(st), C.c_uint(bits))
    return st, w


def _named(named):
    names = [n.encode() for n in named]
    mats, n, ptrs, hs, ws = _mat_args(list(named.values()))
    cn = (C.c_char_p * n)(*names)
    return mats, n, cn, ptrs, hs, ws


class OracleMachine:
    def __init__(self, machine):
        d = _a(machine.descriptor())
        self.h = lib().zko_machine_new(_p(d), C.c_size_t(d.size))
        if not self.h:
            raise RuntimeError("oracle: " + err())
        self.machine = machine
        self.pk = None

    def chip_info(self, name):
        out = np.zeros(3, np.uint32)
        assert lib().zko_chip_info(C.c_void_p(self.h), name.encode(), _p(out)) == 0
        return dict(perm_width_ef=int(out[0]), num_constraints=int(out[1]), log_quotient_degree=int(out[2]))

    def setup(self, prep: dict, pc_start=0, init_gsum=None):
        mats, n, cn, ptrs, hs, ws = _named(prep)
        commit = np.zeros(8, np.uint32)
        gs = _a(init_gsum) if init_gsum is not None else np.array([637514027, 1595065213, 1998064738, 72333738, 1211544370, 822986770, 1518535784, 1604177449, 90440090, 259343427, 140470264, 1162099742, 941559812, 1064053343], np.uint32)   # SepticDigest::zero()
        self.pk = lib().zko_setup(C.c_void_p(self.h), C.c_int(n), cn, ptrs, hs, ws, C.c_uint32(pc_start), _p(gs), _p(commit))
        if not self.pk:
            raise RuntimeError("oracle: " + err())
        return commit

    def initial_challenger(self):
        st = np.zeros(34, np.uint32)
        lib().zko_pk_initial_challenger(C.c_void_p(self.pk), _p(st))
        return st

    def commit_shard(self, traces: dict):
        mats, n, cn, ptrs, hs, ws = _named(traces)
        out = np.zeros(8, np.uint32)
        if lib().zko_commit_shard(C.c_void_p(self.h), C.c_int(n), cn, ptrs, hs, ws, _p(out)):
            raise RuntimeError("oracle: " + err())
        return out

    def prove_shard(self, traces: dict, public_values, challenger=None):
        mats, n, cn, ptrs, hs, ws = _named(traces)
        pv = _a(public_values)
        st = _a(challenger).copy() if challenger is not None else self.initial_challenger()
        outp, outn = u32p(), C.c_size_t()
        rc = lib().zko_prove_shard(C.c_void_p(self.h), C.c_void_p(self.pk), C.c_int(n), cn, ptrs, hs, ws, _p(pv),
                                   C.c_size_t(pv.size), _p(st), C.byref(outp), C.byref(outn))
        if rc:
            raise RuntimeError("oracle: " + err())
        proof = np.ctypeslib.as_array(outp, shape=(outn.value,)).copy()
        lib().zko_free(outp)
        return proof, st

    def verify_shard(self, proof, challenger=None):
        p = _a(proof)
        st = _a(challenger) if challenger is not None else self.initial_challenger()
        rc = lib().zko_verify_shard(C.c_void_p(self.h), C.c_void_p(self.pk), _p(p), C.c_size_t(p.size), _p(st))
        return (rc == 0), (err() if rc else "")

    def permutation_trace(self, chip, prep, main, alpha, beta):
        main = _a(main)
        info = self.chip_info(chip)
        out = np.zeros((main.shape[0], 4 * info["perm_width_ef"]), np.uint32)
        ls = np.zeros(4, np.uint32)
        pp = _a(prep) if prep is not None else None
        a, b = _a(alpha), _a(beta)
        rc = lib().zko_permutation_trace(C.c_void_p(self.h), chip.encode(), _p(pp) if pp is not None else None, _p(main),
                                         C.c_size_t(main.shape[0]), _p(a), _p(b), _p(out), _p(ls))
        if rc:
            raise RuntimeError("oracle: " + err())
        return out, ls

    def derive_multiplicities(self, receiver, receiver_prep, senders, main_width):
        """The multiplicity columns of a receive-only table from the rows of its senders (zko_derive_multiplicities).
        receiver_prep: (height, prep_width) canonical; senders: [(chip, prep or None, main)] canonical rows.  Returns
        ((height, main_width) canonical, number of lookups)."""
        rp = _a(receiver_prep)
        out = np.zeros((rp.shape[0], int(main_width)), np.uint32)
        k = len(senders)
        names = (C.c_char_p * max(1, k))(*[s[0].encode() for s in senders])
        preps = [(_a(s[1]) if s[1] is not None else None) for s in senders]
        mains = [_a(s[2]) for s in senders]
        pp = (C.c_void_p * max(1, k))(*[(p.ctypes.data if p is not None else None) for p in preps])
        mp = (C.c_void_p * max(1, k))(*[m.ctypes.data for m in mains])
        hs = (C.c_size_t * max(1, k))(*[m.shape[0] for m in mains])
        n = C.c_ulonglong(0)
        rc = lib().zko_derive_multiplicities(C.c_void_p(self.h), receiver.encode(), _p(rp), C.c_size_t(rp.shape[0]), C.c_int(k), names, pp, mp,
                                             hs, _p(out), C.byref(n))
        if rc:
            raise RuntimeError("oracle: " + err())
        return out, int(n.value)

    def quotient_values(self, chip, log_n, prep_lde, main_lde, perm_lde, perm_alpha, perm_beta, local_sum, global_sum,
                        alpha, pub):
        info = self.chip_info(chip)
        q = 1 << (log_n + info["log_quotient_degree"])
        out = np.zeros((q, 4), np.uint32)
        pl = _a(prep_lde) if prep_lde is not None else None
        ml, el = _a(main_lde), _a(perm_lde)
        args = [_a(x) for x in (perm_alpha, perm_beta, local_sum, global_sum, alpha, pub)]
        rc = lib().zko_quotient_values(C.c_void_p(self.h), chip.encode(), C.c_uint(log_n), _p(pl) if pl is not None else None,
                                       _p(ml), _p(el), *[_p(x) for x in args[:5]], _p(args[5]), C.c_size_t(args[5].size), _p(out))
        if rc:
            raise RuntimeError("oracle: " + err())
        return out
