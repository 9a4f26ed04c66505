"""The WHOLE library on the CPU: tests/cudaemu/build.py compiles every source of libzkb200.so (all kernels, the prover, the C
ABI) for the host against the stand-in CUDA runtime (tests/cudaemu/cuda_runtime.h: CUDA threads as fibers, barriers by arrival
count, warp shuffles, dynamic shared memory, poisoned shared / device memory), and the `-m gpu` tests run UNCHANGED against
that build in a child pytest process (plugin tests/cudaemu/cudaemu_plugin.py: the emulated library in place of the product's,
torch "cuda" tensors = registered host buffers, data-driven K3 / K5 kernels because generated ones are CUDA binaries).

What this is: a second, hardware-free execution of the product's own source text, bit-exact against the oracle - the whole
shard proof included.  What it is not: evidence about races between threads of a phase, memory-model effects or speed; the
B200 runs are that.  Only a selection runs here (an emulated proof of the mini shard takes about 20 s); any GPU test can be run
the same way:  PYTHONPATH=tests/cudaemu python -m pytest -p cudaemu_plugin -m gpu tests/<file>::<test>
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_emulated(*selection, timeout=900):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.path.join(ROOT, "tests", "cudaemu") + os.pathsep + env.get("PYTHONPATH", "")
    env.pop("ZKB200_LIB", None)
    r = subprocess.run([sys.executable, "-m", "pytest", "-p", "cudaemu_plugin", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider", *selection],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-25:])
    assert r.returncode == 0, tail
    return tail


def test_whole_shard_proof_is_bit_exact_under_emulation():
    """commit + open of the mini machine's shard (every stage of the path: layout, LDE, Merkle, LogUp, quotient, opening, FRI,
    grind, queries), word for word the oracle's proof and accepted by its verifier"""
    out = _run_emulated("tests/test_gpu_parity.py::test_shard_proof_bit_exact_and_verifies")
    assert "1 passed" in out, out


def test_kernel_level_entry_points_under_emulation():
    out = _run_emulated("tests/test_gpu_parity.py::test_poseidon2_permute_batch", "tests/test_gpu_parity.py::test_commit_matches_oracle",
                        "tests/test_gpu_parity.py::test_grind_matches_oracle", "tests/test_gpu_parity.py::test_fri_fold_matches_oracle")
    assert "passed" in out and "failed" not in out, out


def test_derive_multiplicities_gpu_cases_under_emulation():
    """K7 through the C ABI (zkb200_derive_multiplicities) and shards proved with ZKB200_TRACE_DERIVED tables (Byte and Program
    never handed over): the test cases the B200 run will execute"""
    out = _run_emulated("tests/test_zzzz_derive.py", "-k", "not real_keccak")      # that one takes 100 s emulated; green, see profiles/
    assert "6 passed" in out, out


@pytest.mark.parametrize("selection", [
    "tests/test_zzy_tracegen_more.py::test_syscall_instrs_shard_proves_from_event_records",
    "tests/test_zzz_tracegen_global.py::test_global_shard_proves_bit_exact",
])
def test_event_record_shards_prove_bit_exact_under_emulation(selection):
    """Tables handed to zkb200_commit as EVENT RECORDS (row fillers inside the commit; for Global the lift and the curve-point
    scan): chips whose GPU cases were written after the round's GPU budget was spent"""
    out = _run_emulated(selection, timeout=1500)
    assert "passed" in out and "failed" not in out, out
