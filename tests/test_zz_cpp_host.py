"""The compiled host side above the C ABI: include/zkb200.hpp (C++ mirror of the MachineProver trait,
crates/stark/src/prover.rs:30-184) driven by examples/prove_shard.cpp on a case file, without Python
in the proving process."""
import os
import subprocess

import numpy as np
import pytest

from ziren_b200 import casefile, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "examples", "prove_shard")


@pytest.fixture(scope="module")
def driver():
    subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True, stdout=subprocess.DEVNULL)
    return BIN


def test_case_file_layout(tmp_path):
    case = synthetic.mini_case(seed=3)
    path = str(tmp_path / "c.zkcase")
    casefile.write_case(path, case.machine, case.prep, [(case.traces, case.public_values)] * 2)
    w = np.fromfile(path, dtype="<u4")
    desc = case.machine.descriptor()
    assert w[0] == casefile.MAGIC_CASE and w[1] == 1 and w[2] == desc.size
    assert np.array_equal(w[3:3 + desc.size], desc)
    assert w[3 + desc.size + 15] == len(case.prep)


def test_proofs_file_round_trip(tmp_path):
    proofs = [np.arange(10, dtype=np.uint32), np.arange(3, dtype=np.uint32) + 7]
    words = [np.array([casefile.MAGIC_PROOFS, 1], np.uint32), np.arange(8, dtype=np.uint32) + 100, np.array([len(proofs)], np.uint32)]
    for p in proofs:
        words += [np.array([p.size], np.uint32), p]
    path = str(tmp_path / "p.out")
    np.concatenate(words).astype("<u4").tofile(path)
    commit, got = casefile.read_proofs(path)
    assert commit.tolist() == list(range(100, 108)) and [g.tolist() for g in got] == [p.tolist() for p in proofs]


def test_driver_builds_and_fails_loudly_without_a_gpu(driver, tmp_path):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present: covered by the gpu test below")
    except ImportError:
        pass
    case = synthetic.mini_case(seed=3)
    path = str(tmp_path / "c.zkcase")
    casefile.write_case(path, case.machine, case.prep, [(case.traces, case.public_values)])
    r = subprocess.run([driver, path, str(tmp_path / "p.out")], capture_output=True, text=True)
    assert r.returncode == 1 and "CUDA" in r.stderr          # no CPU fallback behind the ABI
    r = subprocess.run([driver], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


@pytest.mark.gpu
def test_cpp_driver_proofs_match_oracle(driver, tmp_path, oracle):
    a, b = synthetic.mini_case(seed=3), synthetic.mini_case(seed=3)
    # two records of the same machine with different main traces
    b.traces = synthetic.mini_case(seed=4).traces
    b.public_values = synthetic.mini_case(seed=4).public_values
    path, out = str(tmp_path / "c.zkcase"), str(tmp_path / "p.out")
    casefile.write_case(path, a.machine, a.prep, [(a.traces, a.public_values), (b.traces, b.public_values)])
    r = subprocess.run([driver, path, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    commit, proofs = casefile.read_proofs(out)
    om = oracle.OracleMachine(a.machine)
    assert np.array_equal(commit, om.setup(a.prep))
    assert len(proofs) == 2
    for case, got in ((a, proofs[0]), (b, proofs[1])):
        want, _ = om.prove_shard(case.traces, case.public_values)
        ok, err = om.verify_shard(got)
        assert ok, err
        assert np.array_equal(got, want)
