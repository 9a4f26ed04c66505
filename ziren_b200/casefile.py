"""Case files for the compiled host driver examples/prove_shard.cpp (format documented there): a
machine descriptor, its preprocessed traces and one or more records (main traces + public values),
everything the C++ mirror of the MachineProver trait needs to prove shards without Python."""
from __future__ import annotations

import numpy as np

from . import field as kb

MAGIC_CASE, MAGIC_PROOFS = 0x53434B5A, 0x4F504B5A     # "ZKCS", "ZKPO"


def _trace_words(name: str, rows: np.ndarray) -> list[np.ndarray]:
    b = name.encode()
    pad = (-len(b)) % 4
    h, w = rows.shape
    head = np.array([len(b)], np.uint32)
    nb = np.frombuffer(b + b"\0" * pad, dtype=np.uint32)
    dims = np.array([h & 0xFFFFFFFF, h >> 32, w], np.uint32)
    return [head, nb, dims, np.ascontiguousarray(kb.to_monty(rows), dtype=np.uint32).ravel()]


def write_case(path: str, machine, prep: dict, records: list, pc_start: int = 0, initial_global_sum=None) -> None:
    """records: list of (traces dict name -> canonical rows, public values)."""
    parts = [np.array([MAGIC_CASE, 1], np.uint32)]
    desc = np.ascontiguousarray(machine.descriptor(), dtype=np.uint32)
    parts += [np.array([desc.size], np.uint32), desc, np.array([pc_start], np.uint32)]
    parts.append(kb.SEPTIC_DIGEST_ZERO.copy() if initial_global_sum is None else np.asarray(initial_global_sum, np.uint32))
    parts.append(np.array([len(prep)], np.uint32))
    for name, rows in prep.items():
        parts += _trace_words(name, rows)
    parts.append(np.array([len(records)], np.uint32))
    for traces, pv in records:
        parts.append(np.array([len(traces)], np.uint32))
        for name, rows in traces.items():
            parts += _trace_words(name, rows)
        pv = np.asarray(pv, np.uint32)
        parts += [np.array([pv.size], np.uint32), pv]
    with open(path, "wb") as f:
        for p in parts:
            f.write(np.ascontiguousarray(p, dtype="<u4").tobytes())


def read_proofs(path: str):
    """-> (preprocessed commitment[8], list of ZKPF proof word arrays)"""
    w = np.fromfile(path, dtype="<u4")
    assert w[0] == MAGIC_PROOFS and w[1] == 1, "not a ZKPO v1 file"
    commit, n, at, out = w[2:10].copy(), int(w[10]), 11, []
    for _ in range(n):
        k = int(w[at])
        out.append(w[at + 1:at + 1 + k].copy())
        at += 1 + k
    return commit, out
