"""CPU tests of the product's arithmetic headers (ziren_b200/csrc/kb31.cuh, poseidon2.cuh) compiled
for the host: the CUDA kernels run exactly these expressions (pipe-pinned adds, plus-form Montgomery
reduction, shift-based 2^-k multiplies), so they are pinned here against the reference-header golden
vectors (tests/golden/ref_vectors.json) and against the oracle on random and edge inputs."""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest

from ziren_b200 import field as kb

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = json.load(open(os.path.join(HERE, "golden", "ref_vectors.json")))
P = kb.P


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _permute(host, states):
    a = np.ascontiguousarray(states, dtype=np.uint32).copy()
    host.hostcheck_permute(_p(a), ctypes.c_size_t(a.size // 16))
    return a


def test_permutation_matches_reference_header_vectors(host):
    for v in GOLD["poseidon2"]:
        assert _permute(host, v["in"]).tolist() == v["out"]


def test_permutation_matches_oracle_on_random_and_edge_states(host, oracle):
    rng = np.random.default_rng(0xC0FFEE)
    states = rng.integers(0, P, size=(256, 16), dtype=np.uint32)
    # edge residues: 0, 1, p-1, values whose low 3/4/8/24 bits are all ones or all zeros (the
    # shift-based 2^-k multiplies split on exactly those bits)
    edges = [0, 1, 2, P - 1, P - 2, (1 << 24) - 1, 1 << 24, (1 << 8) - 1, 1 << 8, 7, 8, 15, 16, (P - 1) >> 1, (P + 1) >> 1,
             0x7effffff, 0x7f000000]
    for i, e in enumerate(edges):
        states[i, :] = e
        states[32 + i, i % 16] = e
    got = _permute(host, states)
    for s, g in zip(states, got.reshape(-1, 16)):
        assert oracle.permute(s.tolist()).tolist() == g.tolist()


def test_field_ops_match_reference_header_vectors(host):
    a, b = np.array(GOLD["a"], np.uint32), np.array(GOLD["b"], np.uint32)
    out = np.zeros_like(a)
    for op, key in [(0, "mul"), (1, "add"), (2, "sub"), (3, "inv"), (4, "to_monty")]:
        host.hostcheck_field(op, _p(a), _p(b), _p(out), ctypes.c_size_t(a.size))
        assert out.tolist() == GOLD[key], key


def test_field_ops_on_edges(host):
    vals = np.array([0, 1, 2, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, 0x7effffff, 1 << 24, (1 << 24) - 1], np.uint32)
    a, b = [x.ravel().copy() for x in np.meshgrid(vals, vals)]
    out = np.zeros_like(a)
    host.hostcheck_field(0, _p(a), _p(b), _p(out), ctypes.c_size_t(a.size))
    assert out.tolist() == ((a.astype(np.uint64) * b.astype(np.uint64)) % P).tolist()
    host.hostcheck_field(5, _p(a), None, _p(out), ctypes.c_size_t(a.size))
    assert all((2 * int(h)) % P == int(x) for h, x in zip(out, a))


def test_ef_mul_and_lazy_accumulator_match_oracle(host, oracle):
    rng = np.random.default_rng(7)
    for _ in range(20):
        a = rng.integers(0, P, 4, dtype=np.uint32)
        b = rng.integers(0, P, 4, dtype=np.uint32)
        out = np.zeros(4, np.uint32)
        host.hostcheck_ef_mul(_p(a), _p(b), _p(out))
        assert out.tolist() == oracle.ef_mul(a.tolist(), b.tolist()).tolist()
    # sums of EF x base products with worst-case operands (p-1) in every position and ragged lengths
    for n in (1, 3, 4, 5, 8, 13):
        for fill in (None, P - 1):
            w = rng.integers(0, P, 4 * n, dtype=np.uint32) if fill is None else np.full(4 * n, fill, np.uint32)
            x = rng.integers(0, P, n, dtype=np.uint32) if fill is None else np.full(n, fill, np.uint32)
            out = np.zeros(4, np.uint32)
            host.hostcheck_efacc(_p(w), _p(x), ctypes.c_size_t(n), _p(out))
            want = [sum(int(w[4 * i + j]) * int(x[i]) for i in range(n)) % P for j in range(4)]
            assert out.tolist() == want
