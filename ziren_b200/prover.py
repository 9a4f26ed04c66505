"""Host-side mirror of the reference's `MachineProver` plugin trait over the zkb200 C ABI.

`B200Prover` has the trait's methods with the same names and meaning
(crates/stark/src/prover.rs:30-184): `setup`, `pk_to_device`, `commit`, `open`, `prove`.
Where the reference returns Rust structs it returns thin Python holders; proofs are the flat
canonical "ZKPF" word array of include/zkb200.h (`ziren_b200.proof.parse` decodes it).
All numeric work happens in libzkb200.so; this module only marshals pointers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from . import field as kb
from .air import Machine

u32p = _ffi.u32p


def _data_ptr(x) -> int:
    """Pointer of a host numpy array or a torch tensor (host or CUDA)."""
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()  # torch.Tensor


def _shape(x):
    return tuple(x.shape)


def _np32(x) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=np.uint32)


def _p(a: np.ndarray):
    return a.ctypes.data_as(u32p)


class ZkbError(RuntimeError):
    """`MachineProver::Error` (crates/stark/src/prover.rs:41): carries zkb200_last_error."""


class ProvingKey:
    """`DeviceProvingKey: MachineProvingKey<SC>` (crates/stark/src/prover.rs:187-199)."""

    def __init__(self, prover, handle, commit):
        self._prover, self._h, self.commit = prover, handle, commit

    def preprocessed_commit(self) -> np.ndarray:
        return self.commit

    def observe_into(self) -> np.ndarray:
        """Challenger image after `pk.observe_into(fresh challenger)` (prover.rs:714-721)."""
        st = np.zeros(34, np.uint32)
        _ffi.lib().zkb200_pk_initial_challenger(self._h, _p(st))
        return st

    def free(self):
        if self._h:
            _ffi.lib().zkb200_pk_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ShardMainData:
    """`ShardMainData<SC, DeviceMatrix, DeviceProverData>` (crates/stark/src/types.rs:16-22)."""

    def __init__(self, handle, main_commit, public_values):
        self._h, self.main_commit, self.public_values = handle, main_commit, public_values

    def device(self) -> int:
        """CUDA device the shard was committed on (a multi-GPU prover routes shards)."""
        return _ffi.lib().zkb200_shard_device(self._h)

    def free(self):
        if self._h:
            _ffi.lib().zkb200_shard_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class EventTrace:
    """A table handed to `commit` as the chip's EVENT RECORDS (ZKB200_TRACE_EVENTS, include/zkb200.h): the library's row
    filler generates it on the device (MachineAir::generate_trace moved into the commit).  `events`: (n, words) uint32
    records (numpy, pinned/pageable torch tensor or CUDA tensor): 7-word ALU / control-flow events
    (ziren_b200/tracegen.py) or 384-word KeccakSponge block records (ziren_b200/keccak_sponge.py)."""

    def __init__(self, events, log_height: int, width: int):
        self.events, self.log_height, self.width = events, int(log_height), int(width)
        shp = _shape(events)
        self.n_events = int(shp[0]) if len(shp) else 0


class DerivedTrace:
    """A receive-only table (Byte, Program) that is NOT handed over (ZKB200_TRACE_DERIVED): `prove_shard` counts its
    multiplicity columns on the device from the shard's other tables (K7).  `height`: the height of its preprocessed trace."""

    def __init__(self, height: int, width: int):
        self.height, self.width = int(height), int(width)


class ColMajorTrace:
    """A device-resident COLUMN-MAJOR table (ZKB200_TRACE_COL_MAJOR): `data` is a CUDA tensor of width x height words."""

    def __init__(self, data, height: int, width: int):
        self.data, self.height, self.width = data, int(height), int(width)


class B200Prover:
    """`impl MachineProver<KoalaBearPoseidon2, A> for B200Prover`.  `device`: one GPU index, a list of
    indices, or -1 for every visible GPU — one prover object whose commit() routes each shard to the
    least-loaded device (what prove_with_context's worker threads share, prove.rs:487-521)."""

    def __init__(self, machine: Machine, device=0):
        self.machine_ = machine
        desc = _np32(machine.descriptor())
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            ids = (C.c_int * len(device))(*device)
            rc = _ffi.lib().zkb200_ctx_create_multi(ids, len(device), _p(desc), desc.size, C.byref(h))
        else:
            rc = _ffi.lib().zkb200_ctx_create(device, _p(desc), desc.size, C.byref(h))
        if rc:
            raise ZkbError(_ffi.lib().zkb200_last_error(None).decode())
        self._h = h
        self.device = device

    # -- plumbing -------------------------------------------------------------------------------
    def _check(self, rc):
        if rc:
            raise ZkbError(_ffi.lib().zkb200_last_error(self._h).decode())

    def _traces(self, named):
        """named: dict name -> row-major (height, width) uint32/int32 Montgomery array
        (numpy, pinned/pageable torch tensor, or CUDA torch tensor)."""
        arr = (_ffi.Trace * len(named))()
        keep = []
        for i, (name, t) in enumerate(named.items()):
            b = name.encode()
            keep.append((b, t))
            if isinstance(t, EventTrace):
                arr[i] = _ffi.Trace(b, _data_ptr(t.events) if t.n_events else None, 1 << t.log_height, t.width,
                                    _ffi.TRACE_EVENTS, t.n_events)
            elif isinstance(t, DerivedTrace):
                arr[i] = _ffi.Trace(b, None, t.height, t.width, _ffi.TRACE_DERIVED, 0)
            elif isinstance(t, ColMajorTrace):
                arr[i] = _ffi.Trace(b, _data_ptr(t.data), t.height, t.width, _ffi.TRACE_COL_MAJOR, 0)
            else:
                h, w = _shape(t)
                arr[i] = _ffi.Trace(b, _data_ptr(t), h, w, 0, 0)
        return arr, keep

    def machine(self) -> Machine:
        return self.machine_

    def num_devices(self) -> int:
        return _ffi.lib().zkb200_ctx_num_devices(self._h)

    def stream_ptr(self) -> int:
        return _ffi.lib().zkb200_ctx_stream(self._h)

    def sync(self):
        self._check(_ffi.lib().zkb200_sync(self._h))

    def close(self):
        if self._h:
            _ffi.lib().zkb200_ctx_destroy(self._h)
            self._h = None

    # -- MachineProver ----------------------------------------------------------------------------
    def setup(self, preprocessed: dict, pc_start: int = 0, initial_global_cumulative_sum=None) -> ProvingKey:
        """`setup` + `pk_to_device` (prover.rs:49-63): commit preprocessed traces on the GPU."""
        arr, keep = self._traces(preprocessed)
        gs = _np32(initial_global_cumulative_sum) if initial_global_cumulative_sum is not None else kb.SEPTIC_DIGEST_ZERO.copy()
        commit = np.zeros(8, np.uint32)
        h = C.c_void_p()
        self._check(_ffi.lib().zkb200_setup(self._h, arr, len(preprocessed), pc_start, _p(gs), _p(commit), C.byref(h)))
        return ProvingKey(self, h, commit)

    pk_to_device = setup

    def commit(self, traces: dict, public_values) -> ShardMainData:
        """`commit(record, traces)` (prover.rs:258-292)."""
        arr, keep = self._traces(traces)
        pv = _np32(public_values)
        commit = np.zeros(8, np.uint32)
        h = C.c_void_p()
        self._check(_ffi.lib().zkb200_commit(self._h, arr, len(traces), _p(pv), pv.size, _p(commit), C.byref(h)))
        return ShardMainData(h, commit, pv)

    def open(self, pk: ProvingKey, data: ShardMainData, challenger: np.ndarray | None = None):
        """`open(pk, data, challenger)` (prover.rs:298-653).  Returns (proof words, challenger)."""
        st = _np32(challenger).copy() if challenger is not None else pk.observe_into()
        out, n = u32p(), C.c_size_t()
        self._check(_ffi.lib().zkb200_open(self._h, pk._h, data._h, _p(st), C.byref(out), C.byref(n)))
        proof = np.ctypeslib.as_array(out, shape=(n.value,)).copy()
        _ffi.lib().zkb200_free(out)
        return proof, st

    def prove_shard(self, pk: ProvingKey, traces: dict, public_values, challenger: np.ndarray | None = None):
        """commit + open for one record: the loop body of `prove` (prover.rs:681-688)."""
        if any(isinstance(t, DerivedTrace) for t in traces.values()):
            # the commit of derived tables reads the proving key's preprocessed traces: one call (zkb200_prove_shard)
            arr, keep = self._traces(traces)
            pv = _np32(public_values)
            st = _np32(challenger).copy() if challenger is not None else pk.observe_into()
            out, n = u32p(), C.c_size_t()
            self._check(_ffi.lib().zkb200_prove_shard(self._h, pk._h, arr, len(traces), _p(pv), pv.size, _p(st), C.byref(out), C.byref(n)))
            proof = np.ctypeslib.as_array(out, shape=(n.value,)).copy()
            _ffi.lib().zkb200_free(out)
            return proof, st
        data = self.commit(traces, public_values)
        try:
            return self.open(pk, data, challenger)
        finally:
            data.free()

    def prove(self, pk: ProvingKey, records: list):
        """`prove(pk, records, challenger)` (prover.rs:660-693): every shard starts from a clone
        of the post-observe_into challenger.  records: list of (traces, public_values)."""
        base = pk.observe_into()
        return [self.prove_shard(pk, tr, pv, base)[0] for tr, pv in records]

    # -- profiling ------------------------------------------------------------------------------
    @staticmethod
    def launch_count() -> int:
        """CUDA kernels launched by libzkb200 in this process so far."""
        return int(_ffi.lib().zkb200_launch_count())

    def set_profile(self, on: bool):
        _ffi.lib().zkb200_set_profile(self._h, int(on))

    def last_stage_times(self) -> dict:
        n = _ffi.lib().zkb200_last_stage_times(self._h, None, None, 0)
        names = (C.c_char_p * n)()
        ms = (C.c_float * n)()
        _ffi.lib().zkb200_last_stage_times(self._h, names, ms, n)
        out = {}
        for i in range(n):
            out[names[i].decode()] = out.get(names[i].decode(), 0.0) + float(ms[i])
        return out

    # -- kernel-level entry points (device pointers, column-major, Montgomery) ------------------
    def coset_lde(self, inp, out, log_n, width, log_blowup=1, shift=3):
        self._check(_ffi.lib().zkb200_coset_lde(self._h, _data_ptr(inp), _data_ptr(out), log_n, width, log_blowup, shift))

    def ntt(self, inp, out, log_n, width, inverse=False, bitrev_out=False):
        self._check(_ffi.lib().zkb200_ntt(self._h, _data_ptr(inp), _data_ptr(out), log_n, width, int(inverse), int(bitrev_out)))

    def mmcs_root(self, mats, log_heights, widths) -> np.ndarray:
        n = len(mats)
        ptrs = (C.c_void_p * n)(*[_data_ptr(m) for m in mats])
        lh = (C.c_uint * n)(*log_heights)
        ws = (C.c_size_t * n)(*widths)
        root = np.zeros(8, np.uint32)
        self._check(_ffi.lib().zkb200_mmcs_root(self._h, ptrs, lh, ws, n, _p(root)))
        return root

    def poseidon2_permute_batch(self, states, n):
        self._check(_ffi.lib().zkb200_poseidon2_permute_batch(self._h, _data_ptr(states), n))

    def permutation_trace(self, chip, prep, main, height, alpha, beta, out) -> np.ndarray:
        a, b = _np32(alpha), _np32(beta)
        ls = np.zeros(4, np.uint32)
        self._check(_ffi.lib().zkb200_permutation_trace(self._h, chip.encode(), _data_ptr(prep) if prep is not None else None,
                                                        _data_ptr(main), height, _p(a), _p(b), _data_ptr(out), _p(ls)))
        return ls

    def derive_multiplicities(self, receiver: str, receiver_prep, receiver_height: int, senders, out) -> int:
        """K7: the multiplicity columns of a receive-only table (Byte, Program, range tables) from the rows of the tables that
        send to it.  senders: [(chip, prep or None, main, height)], CUDA tensors column-major Montgomery; out: CUDA tensor of
        receiver_height x main_width words, column-major.  Returns the number of lookups counted."""
        arr = (_ffi.Table * max(1, len(senders)))()
        for i, (chip, prep, main, height) in enumerate(senders):
            arr[i] = _ffi.Table(chip.encode(), _data_ptr(prep) if prep is not None else None, _data_ptr(main) if height else None, int(height))
        n = C.c_ulonglong(0)
        self._check(_ffi.lib().zkb200_derive_multiplicities(self._h, receiver.encode(), _data_ptr(receiver_prep), int(receiver_height),
                                                            C.cast(arr, C.c_void_p), len(senders), _data_ptr(out), C.byref(n)))
        return int(n.value)

    def quotient(self, chip, log_n, prep_lde, main_lde, perm_lde, perm_alpha, perm_beta, local_sum, global_sum, alpha, pub, out):
        args = [_np32(x) for x in (perm_alpha, perm_beta, local_sum, global_sum, alpha, pub)]
        self._check(_ffi.lib().zkb200_quotient(self._h, chip.encode(), log_n, _data_ptr(prep_lde) if prep_lde is not None else None,
                                               _data_ptr(main_lde), _data_ptr(perm_lde), *[_p(x) for x in args[:5]], _p(args[5]),
                                               args[5].size, _data_ptr(out)))

    def fri_fold(self, inp, m, beta, ro_next, out):
        b = _np32(beta)
        self._check(_ffi.lib().zkb200_fri_fold(self._h, _data_ptr(inp), m, _p(b), _data_ptr(ro_next) if ro_next is not None else None,
                                               _data_ptr(out)))

    def grind(self, challenger, bits) -> int:
        st = _np32(challenger)
        w = C.c_uint32()
        self._check(_ffi.lib().zkb200_grind(self._h, _p(st), bits, C.byref(w)))
        return w.value

    def transpose(self, inp, out, height, width, to_colmajor=True):
        self._check(_ffi.lib().zkb200_transpose(self._h, _data_ptr(inp), _data_ptr(out), height, width, int(to_colmajor)))

    def convert(self, data, n, to_montgomery=True):
        self._check(_ffi.lib().zkb200_convert(self._h, _data_ptr(data), n, int(to_montgomery)))

    # ---- trace generation ----------------------------------------------
    def generate_keccak_sponge_trace(self, blocks, log_height: int, out, col_major: bool = False):
        """`MachineAir::generate_trace` of the KeccakSponge chip on the GPU.  `blocks`: (n, 384) uint32 block records
        (ziren_b200/keccak_sponge.py; numpy or a CUDA tensor); `out`: CUDA tensor of 2^log_height x 3531 words."""
        n = int(_shape(blocks)[0]) if len(_shape(blocks)) else 0
        self._check(_ffi.lib().zkb200_generate_keccak_sponge_trace(self._h, _data_ptr(blocks) if n else None, n, int(log_height),
                                                                   _data_ptr(out), int(col_major)))

    def generate_alu_trace(self, chip: str, events, log_height: int, out, col_major: bool = False):
        """`MachineAir::generate_trace` of an ALU chip (AddSub, Bitwise, Lt, ShiftLeft, ShiftRight,
        CloClz) on the GPU.  `events`: the record's `Vec<AluEvent>` as (n, 7) uint32 words (numpy or a
        CUDA tensor, see ziren_b200/tracegen.py); `out`: CUDA tensor of 2^log_height x width words."""
        n = int(_shape(events)[0]) if len(_shape(events)) else 0
        self._check(_ffi.lib().zkb200_generate_alu_trace(self._h, chip.encode(), _data_ptr(events) if n else None, n,
                                                         int(log_height), _data_ptr(out), int(col_major)))

