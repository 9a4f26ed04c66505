#!/usr/bin/env python3
"""Headline benchmark: MIPS cycles proved per second (BASELINE.json `metric`).

A "step" proves ONE shard of the named workload through the reference-facing C ABI
(`zkb200_commit` + `zkb200_open`, i.e. MachineProver::{commit, open}).  Default workload at N=1 is
BASELINE.json configs[1] ("examples/keccak-precompile, 2^20-row shards"): a core shard with a 2^20-row Cpu table
(synthetic core tables: no guest ELF can be built here, see ziren_b200/synthetic.py) whose area is dominated by the
reference's REAL KeccakSponge precompile chip - its 3531-column layout, its restated Air::eval (3 788 constraints,
357 lookups; ziren_b200/keccak_air.py), 2^18 rows from the row filler on well-formed sponge events.
cycles per shard = rows of the Cpu table.

  value : device-resident inputs (row-major traces already in HBM), timed with CUDA events on the
          prover's stream, max over ranks.
  e2e   : the same calls with HOST inputs, H2D and the proof's D2H inside the timed region: pinned event records for
          the chip whose table the library generates on the device (ZKB200_TRACE_EVENTS), pinned rows for every other
          table; `e2e.uploaded_keccak_rows` is the same proof with that chip's rows uploaded instead.
  roofline / cpu_baseline : see DESIGN.md §Measurement.

Multi-GPU (torchrun, one rank per GPU): shards are independent, every rank proves
its own shards, no data-path collective; NCCL only gathers the 8-word commitments and the timings.
`--impl reference` times the CPU oracle (the reference itself cannot be built here: no Rust) with
all host threads on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from ziren_b200 import field as kb  # noqa: E402
from ziren_b200 import synthetic  # noqa: E402

METRIC = "mips_cycles_proved_per_sec"
UNIT = "cycles/s"


def bind_to_gpu_numa_node(torch, index):
    """At N > 1 every rank stages 5 GB per shard from pinned host memory: keep the rank's threads
    and its pinned pages on the NUMA node its GPU hangs off (first-touch placement), so the ranks do
    not all pull their uploads through one socket.  Best effort: returns the CPU list it bound to."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if len(cpus) >= 4:
            os.sched_setaffinity(0, cpus)
            return txt
    except Exception:
        pass
    return None


def keccak_blocks(args, lc):
    """Event records that fill the KeccakSponge table of the keccak-real workload: four-block sponges, 24 rows per block,
    as many as fit 2^(lc-2) rows (2^20: 2730 events = 10920 blocks = 262080 of 262144 rows)."""
    from ziren_b200 import keccak_sponge as ksp
    rows = 1 << (lc - 2)
    per = 4 if rows >= 96 else 1
    return ksp.synthetic_blocks(max(1, rows // (24 * per)), per, seed=args.rank, shard=1)


def host_mem_available_gb():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def make_case(args, sample=False, keccak_rows="oracle"):
    synthetic.CONSTRAINTS_PER_GROUP = args.constraint_density
    w = args.workload
    lc = args.sample_log_cpu if sample else args.log_cpu
    if w == "keccak-real":
        # the REAL KeccakSponge chip (3531 columns, restated Air::eval, 357 lookups).  Its rows come from a row filler:
        # the CUDA kernel in the GPU arm (keccak_rows = None: filled in by main()), the oracle in the CPU arm
        blocks = keccak_blocks(args, lc)
        rows = None
        if keccak_rows == "oracle":
            from oracle import oracle_ffi as o
            rows = o.keccak_sponge_trace(blocks, 1 << (lc - 2))
        case = synthetic.keccak_real_case(blocks, rows, log_cpu=lc, seed=0xC0FFEE + args.rank)
        case.blocks = blocks
        return case
    if w == "keccak":
        return synthetic.keccak_case(log_cpu=lc, seed=0xC0FFEE + args.rank)
    if w == "core":
        return synthetic.core_case(log_cpu=lc, seed=0xC0FFEE + args.rank)
    if w == "fibonacci":
        return synthetic.fibonacci_core_case(log_cpu=lc, seed=0xC0FFEE + args.rank)
    if w == "compress":
        return synthetic.compress_case(log_max=lc, seed=0xC0FFEE + args.rank)
    raise SystemExit(f"unknown workload {w}")


def workload_config(args, case, sample=False):
    lc = args.sample_log_cpu if sample else args.log_cpu
    names = {"keccak": "examples/keccak-precompile-like synthetic shard (Cpu 2^%d rows, KeccakSponge 2^%d x 4259 cols)" % (lc, lc - 2),
             "keccak-real": "examples/keccak-precompile-like shard (Cpu 2^%d rows, synthetic core tables) with the REAL KeccakSponge chip: 2^%d rows x "
                            "3531 main + 720 permutation columns (cost 4259 per row as mips_costs.json), restated Air::eval with 3788 "
                            "constraints and 357 lookups, rows from the row filler on well-formed four-block sponge events" % (lc, lc - 2),
             "core": "tendermint-like maximal core shard (maximal_shapes.json[21][1] scaled to Cpu 2^%d)" % lc,
             "fibonacci": "examples/fibonacci-like single core shard (Cpu 2^%d)" % lc,
             "compress": "recursion-compress-like inner proof (shrink shape, tallest table 2^%d, Poseidon2Wide 313 columns), "
                         "setup (preprocessed commit) inside every proof as crates/prover/src/lib.rs:809-832 does" % lc}
    return {"workload": names[args.workload], "cycles_per_shard": case.cycles, "cells_per_shard": case.cells,
            "trace_bytes_per_shard": case.trace_bytes, "fri": {"log_blowup": 1, "num_queries": 84, "pow_bits": 16},
            "constraints_per_6_columns": args.constraint_density,
            "l2_policy": "inputs larger than L2 (trace bytes >> 126 MB); fresh shard allocations every step",
            "parallelism": f"shard-per-gpu x{args.gpus}", "shards_in_flight_per_gpu": {"value": args.value_threads, "e2e": args.e2e_threads},
            **({"total_shards": args.shards, "shards_per_rank": len(range(args.rank, args.shards, args.world))} if args.shards else {})}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [int(s[0]) for s in self.samples if s and s[0].isdigit()]
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args):
    """CPU arm: the oracle port (kind 'port'), all host threads, bounded sample of the workload."""
    from oracle import oracle_ffi as o
    if args.rank != 0:
        return
    t_gen = time.perf_counter()
    case = make_case(args, sample=True)
    t_gen = time.perf_counter() - t_gen
    om = o.OracleMachine(case.machine)
    om.setup(case.prep)
    o.set_num_threads(os.cpu_count())     # torchrun exports OMP_NUM_THREADS=1; use every host core
    cores = o.num_threads()
    # SAME configuration as the GPU arm (one full shard is about 90 s of CPU work at 2^20 cycles), so the sample is
    # bounded by the number of shards, not by their size: one untimed-free pass, `steps` capped at one shard
    same = args.sample_log_cpu == args.log_cpu
    steps = 1 if same else max(1, args.steps)
    t0 = time.perf_counter()
    for _ in range(steps):
        if args.workload == "compress":
            om.setup(case.prep)           # setup is part of every compress proof (crates/prover/src/lib.rs:809-810)
        proof, _ = om.prove_shard(case.traces, case.public_values)
    dt = time.perf_counter() - t0
    val = (case.cycles if case.cycles else 1) * steps / dt
    cfg = workload_config(args, case, sample=True)
    metric, unit = (METRIC, UNIT) if case.cycles else ("compress_inner_proofs_per_sec", "proofs/s")
    line = {"metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus, "steps": steps, "warmup": 0,
            "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 (KoalaBear 31-bit prime field)", "data": "synthetic", "impl": "reference", "config": cfg,
            "same_config_as_gpu_arm": same,
            **({"keccak_trace_generation_s_not_in_value": t_gen} if args.workload == "keccak-real" else {}),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{steps} whole shard(s) of " + ("the bench configuration itself" if same else f"the same machine scaled to Cpu 2^{args.sample_log_cpu}") +
                                       " (requested steps/warmup are capped: the CPU needs ~90 s per 2^20-cycle shard); "
                                       "C++ oracle (OpenMP), not Plonky3's AVX code: the Rust reference cannot be built here"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="keccak-real", choices=["keccak", "keccak-real", "core", "fibonacci", "compress"])
    ap.add_argument("--log-cpu", type=int, default=20)
    ap.add_argument("--sample-log-cpu", type=int, default=None,
                    help="size of the CPU arm's shard (default: the bench configuration itself, capped at 2^20)")
    ap.add_argument("--shards", type=int, default=0,
                    help="strong scaling: prove this many shards in total, split round-robin over the ranks "
                         "(BASELINE configs[2]: --workload core --log-cpu 21 --shards 18; configs[4]: --workload compress --log-cpu 18 --shards 127)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--constraint-density", type=int, default=2,
                    help="constraints per 6-column group of the synthetic tables: 2 (default) ... 12 (about 2 per column)")
    ap.add_argument("--no-pageable", action="store_true", help="skip the pageable-host e2e figure")
    ap.add_argument("--no-verify", action="store_true", help="skip the oracle verifier's check of the timed proof")
    ap.add_argument("--value-threads", type=int, default=3,
                    help="host threads proving device-resident shards concurrently in the `value` arm (compute lanes)")
    ap.add_argument("--e2e-threads", type=int, default=4,
                    help="host threads calling commit/open concurrently in the e2e arm (the reference keeps "
                         "shard_batch_size shards in flight, prove.rs:487-521): uploads of one shard overlap the open of another")
    ap.add_argument("--stages", action="store_true", help="print per-stage device times to stderr")
    args = ap.parse_args()
    args.rank = int(os.environ.get("RANK", "0"))
    args.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.sample_log_cpu is None:
        # the CPU arm proves ONE shard of the same machine: the bench configuration itself for the synthetic workloads
        # (about 75 s per 2^20-cycle shard on 16 cores); the real KeccakSponge chip costs the scalar oracle about four
        # times as much per cycle (3 788 constraints, 357 lookups per row: about 80 s per 2^20-cycle shard on 16 cores) and
        # about 35 GB of host memory: the bench configuration itself on a host with >= 96 GB available, else the same
        # machine at a quarter of the height (about 20 s)
        roomy = host_mem_available_gb() >= 96.0
        args.sample_log_cpu = min(args.log_cpu, 20 if (args.workload != "keccak-real" or roomy) else 18)

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from ziren_b200.prover import B200Prover

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(args.local_rank)
    distributed = args.world > 1
    numa_cpus = bind_to_gpu_numa_node(torch, args.local_rank) if distributed else None
    if distributed:
        # NCCL prints its version banner on stdout when NCCL_DEBUG is VERSION/INFO: keep stdout to the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", args.local_rank))

    case = make_case(args, keccak_rows=None)
    prover = B200Prover(case.machine, device=args.local_rank)
    stream = torch.cuda.ExternalStream(prover.stream_ptr(), device=torch.device("cuda", args.local_rank))
    prep_monty = {k: kb.to_monty(v) for k, v in case.prep.items()}
    pk = prover.setup(prep_monty)
    base_ch = pk.observe_into()
    per_proof_setup = args.workload == "compress"      # crates/prover/src/lib.rs:809-810: setup inside every compress proof
    units_per_shard = case.cycles if case.cycles else 1
    metric, unit = (METRIC, UNIT) if case.cycles else ("compress_inner_proofs_per_sec", "proofs/s")
    # strong scaling (--shards S): S shards in total, rank r proves shards r, r + world, ...
    my_steps = len(range(args.rank, args.shards, args.world)) if args.shards else args.steps

    real = args.workload == "keccak-real"
    keccak_dev = None
    if real:
        # MachineAir::generate_trace of the KeccakSponge chip on the device (zkb200_generate_keccak_sponge_trace): the
        # row-major table every arm below proves; the placeholder only carries its shape into the cell / byte counts
        from ziren_b200 import keccak_sponge as ksp
        from ziren_b200.prover import EventTrace
        log_hk = args.log_cpu - 2
        keccak_dev = torch.empty((1 << log_hk, ksp.WIDTH), dtype=torch.int32, device="cuda")
        prover.generate_keccak_sponge_trace(case.blocks, log_hk, keccak_dev, col_major=False)
        case.traces["KeccakSponge"] = np.broadcast_to(np.zeros(1, np.uint32), (1 << log_hk, ksp.WIDTH))
    # inputs: Montgomery row-major, once in pinned host memory (e2e arm), once resident in HBM
    host_tr = {}
    cfg = workload_config(args, case)
    if numa_cpus:
        cfg["host_numa_binding"] = "each rank bound to its GPU's local_cpulist (rank 0: %s)" % numa_cpus
    cells, shapes = case.cells, {k: v.shape for k, v in case.traces.items()}
    for k in list(case.traces):
        if real and k == "KeccakSponge":
            host_tr[k] = keccak_dev.cpu().pin_memory()
        else:
            host_tr[k] = torch.from_numpy(kb.to_monty(case.traces[k]).view(np.int32)).pin_memory()
        case.traces[k] = None          # the canonical copy is not needed any more (host RAM at 8 ranks)
    dev_tr = {k: (keccak_dev if (real and k == "KeccakSponge") else v.cuda()) for k, v in host_tr.items()}
    h2d_bytes = sum(4 * v.numel() for v in host_tr.values())
    gen_tr = None
    if real:
        # e2e with trace generation moved into the commit: the chip's EVENT RECORDS cross PCIe (pinned), not its rows
        ev_pinned = torch.from_numpy(case.blocks.view(np.int32)).pin_memory()
        gen_tr = {k: (EventTrace(ev_pinned, log_hk, ksp.WIDTH) if k == "KeccakSponge" else v) for k, v in host_tr.items()}
        h2d_bytes_gen = h2d_bytes - 4 * host_tr["KeccakSponge"].numel() + 4 * ev_pinned.numel()
    torch.cuda.synchronize()

    def prove(traces):
        if per_proof_setup:
            pkk = prover.setup(prep_monty)
            proof, _ = prover.prove_shard(pkk, traces, case.public_values, pkk.observe_into())
            pkk.free()
            return proof
        proof, _ = prover.prove_shard(pk, traces, case.public_values, base_ch)
        return proof

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(traces, steps, nthreads=1):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        proofs = [None] * steps
        if nthreads <= 1:
            for i in range(steps):
                proofs[i] = prove(traces)
        else:
            def worker(t):
                for i in range(t, steps, nthreads):
                    proofs[i] = prove(traces)
            ths = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
        proof = proofs[-1]
        prover.sync()
        e1.record(stream)
        e1.synchronize()
        barrier()
        ms = e0.elapsed_time(e1)
        if distributed:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, proof

    for _ in range(args.warmup):
        prove(dev_tr)
    if args.value_threads > 1:
        timed(dev_tr, args.value_threads, args.value_threads)
    sampler = ClockSampler(args.local_rank)
    sampler.start()
    launches0 = prover.launch_count()
    ms_dev, proof = timed(dev_tr, my_steps, args.value_threads)
    launches = prover.launch_count() - launches0
    # e2e inputs.  keccak-real: what a caller of this library hands over - pinned EVENT RECORDS for the chip that has a row
    # filler on the device (KeccakSponge: ZKB200_TRACE_EVENTS, its table is generated inside the commit) and pinned rows for
    # every other table; the arm that uploads the chip's rows instead is timed as well and reported next to it.
    e2e_in = gen_tr if real else host_tr
    timed(e2e_in, max(args.warmup, args.e2e_threads), args.e2e_threads)      # warm the staging/pool paths of the e2e arm
    ms_e2e, proof2 = timed(e2e_in, my_steps, args.e2e_threads)
    # one shard in flight: what the reference's own GPU options ask for (shard_batch_size = 1,
    # crates/stark/src/opts.rs:83-110) - upload, layout change, LDE and leaf hashing overlap INSIDE the shard
    ms_e2e_1, proof3 = timed(e2e_in, my_steps, 1)
    ms_up = ms_up_1 = None
    if real:
        timed(host_tr, max(args.warmup, args.e2e_threads), args.e2e_threads)
        ms_up, proof5 = timed(host_tr, my_steps, args.e2e_threads)
        ms_up_1, proof6 = timed(host_tr, my_steps, 1)
        assert np.array_equal(proof, proof5) and np.array_equal(proof, proof6)
    clocks = sampler.stop()
    assert np.array_equal(proof, proof2) and np.array_equal(proof, proof3)
    # pageable host memory (what RowMajorMatrix.values and the record's event vectors are, prover.rs:258-262): rows staged
    # through the pinned ring, event records by one DMA
    ms_e2e_pageable = None
    if args.world == 1 and not args.no_pageable:
        pageable = {k: np.array(v.numpy(), copy=True) for k, v in host_tr.items()}
        if real:
            pageable["KeccakSponge"] = EventTrace(np.array(case.blocks, copy=True), log_hk, ksp.WIDTH)
        timed(pageable, 1, 1)
        ms_e2e_pageable, proof4 = timed(pageable, max(2, args.steps // 3), 1)
        assert np.array_equal(proof, proof4)
        del pageable
    d2h_bytes = int(proof.size) * 4

    # the proof that was timed is checked: the oracle's verifier (restated from crates/stark/src/verifier.rs
    # and the recursion circuit) must accept it at the bench configuration, and reject a corrupted copy
    verified = None
    if args.rank == 0 and not args.no_verify:
        from oracle import oracle_ffi as o
        om = o.OracleMachine(case.machine)
        o.set_num_threads(os.cpu_count())
        om.setup(case.prep)
        ok, err = om.verify_shard(proof)
        bad = proof.copy()
        bad[proof.size // 2] ^= 1
        verified = bool(ok) and not om.verify_shard(bad)[0]
        if not ok:
            print("oracle verifier REJECTED the timed proof:", err, file=sys.stderr)

    stages = None
    if args.rank == 0:
        prover.set_profile(True)
        prove(dev_tr)
        stages = prover.last_stage_times()
        prover.set_profile(False)
        if args.stages:
            print("stage ms:", json.dumps(stages), file=sys.stderr)

    # commitments are the only thing the ranks exchange (NCCL all_gather of 8 words)
    if distributed:
        c = torch.from_numpy(proof[2:10].astype(np.int64)).cuda()
        allc = [torch.empty_like(c) for _ in range(args.world)]
        dist.all_gather(allc, c)

    roofline = roofline_other = cpu_base = None
    if args.rank == 0:
        k3_bytes = sum(8.0 * shapes[c.name][0] * (c.prep_width + c.main_width + 4 * c.perm_width_ef) + 32.0 * shapes[c.name][0]
                       for c in case.machine.chips if c.name in shapes)
        rl = stage_rooflines(shapes, stages or {}, k3_bytes=k3_bytes, k3_dram_ratio=3.49 if real else None)
        ranked = sorted(rl.values(), key=lambda r: -r["ms"])
        if ranked:
            roofline, roofline_other = ranked[0], ranked[1:]
        if not args.no_cpu_baseline and args.world == 1:      # reported at N=1 only (other ranks would contend for the cores)
            cpu_base = measure_cpu_baseline(args)

    if args.rank == 0:
        total_cycles = units_per_shard * (args.shards if args.shards else args.steps * args.gpus)
        e2e_uploaded = None
        if real:
            e2e_uploaded = {"value": total_cycles / (ms_up / 1e3), "unit": unit, "ms_per_step": ms_up / max(my_steps, 1),
                            "h2d_bytes_per_step": h2d_bytes, "host_threads_in_flight": args.e2e_threads,
                            "one_shard_in_flight": {"value": total_cycles / (ms_up_1 / 1e3), "ms_per_step": ms_up_1 / max(my_steps, 1)},
                            "what": "the same proof with the KeccakSponge ROWS uploaded from pinned host memory (trace generation left on "
                                    "the host, as the synthetic workload of earlier rounds had to)"}
        line = {"metric": metric, "value": total_cycles / (ms_dev / 1e3), "unit": unit, "n_gpus": args.gpus,
                "steps": my_steps, "warmup": args.warmup, "ms_per_step": ms_dev / max(my_steps, 1), "higher_is_better": True,
                "scaling": "strong" if args.shards else "weak", "vs_baseline": None, "dtype": "u32 (KoalaBear 31-bit prime field)", "data": "synthetic",
                "config": cfg,
                "e2e": {"value": total_cycles / (ms_e2e / 1e3), "unit": unit, "h2d_bytes_per_step": h2d_bytes_gen if real else h2d_bytes,
                        "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / max(my_steps, 1),
                        "host_threads_in_flight": args.e2e_threads,
                        **({"inputs": "pinned host memory: EVENT RECORDS of the KeccakSponge chip (zkb200_keccak_block, ZKB200_TRACE_EVENTS: "
                                      "MachineAir::generate_trace runs on the device inside zkb200_commit, SURVEY.md section 8 row f3) and "
                                      "row-major rows of every other table; the proof is word for word the one of the other arms",
                            "uploaded_keccak_rows": e2e_uploaded} if real else {}),
                        "one_shard_in_flight": {"value": total_cycles / (ms_e2e_1 / 1e3), "ms_per_step": ms_e2e_1 / max(my_steps, 1)},
                        "pageable_host_one_shard_in_flight": None if ms_e2e_pageable is None else
                        {"value": units_per_shard * max(2, args.steps // 3) / (ms_e2e_pageable / 1e3), "ms_per_step": ms_e2e_pageable / max(2, args.steps // 3)}},
                "verified": verified, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_other": roofline_other,
                "cpu_baseline": cpu_base,
                "stage_ms": stages, "cells_per_sec": cells * (args.shards if args.shards else args.steps * args.gpus) / (ms_dev / 1e3)}
        print(json.dumps(line))
    pk.free()
    prover.close()
    if distributed:
        dist.barrier()          # ranks leave together (rank 0 may still have been timing the CPU baseline)
        dist.destroy_process_group()


def stage_rooflines(trace_shapes, stages, log_blowup=1, k3_bytes=None, k3_dram_ratio=None):
    """HBM rooflines of the two dominant kernel families from the per-stage CUDA-event times of a
    live profiled step (zkb200_set_profile): K2 = Merkle build of the main commit, K1 = coset LDE of
    the main commit.  Algorithmic bytes per SURVEY.md section 8d."""
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, which = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        peak, which = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    shapes = [(h << log_blowup, w) for h, w in trace_shapes.values()]
    hmax = max(h for h, _ in shapes)
    inj_heights = sorted({h for h, _ in shapes if h < hmax})
    merkle_bytes = sum(4.0 * h * w for h, w in shapes) + 32.0 * hmax + 96.0 * (hmax - 1) + 64.0 * sum(inj_heights)
    perms = sum(h * (-(-w // 8)) for h, w in shapes) + (hmax - 1) + sum(inj_heights)
    lde_bytes = sum(12.0 * h * w for h, w in trace_shapes.values())
    out = {}
    # DRAM traffic / algorithmic bytes of the family's kernels, from the committed ncu captures
    # (profiles/r01_ncu_full_summary_v2.txt: leaf_hash_kernel 2^19 x 512 moves 1.0926 GB for 1.0905 GB algorithmic;
    # profiles/r02_ncu_hot_kernels.txt: the three lean launches of one LDE piece move 2.28 GB for 0.805 GB, the
    # A->B->C intermediates go through HBM)
    ncu_ratio = {"k2_merkle": 1.0019, "k1_lde": 2.83}       # r02: 0.495 + 0.751 + 1.031 GB moved for 0.805 GB algorithmic (2^16 x 1024 piece)
    for key, stage, alg, kern in (("k2_merkle", "commit_main_merkle", merkle_bytes, "merkle_build: leaf_hash_kernel (Poseidon2 sponge over the rows of every height) + compress_kernel"),
                                  ("k1_lde", "commit_main_lde", lde_bytes, "coset_lde_batch: ntt_strided_lean_kernel + ntt_contig_lean_kernel")):
        ms = stages.get(stage)
        if not ms:
            continue
        ach = alg / (ms / 1e3) / 1e9
        out[key] = {"bound": "hbm", "kernel": kern, "achieved": ach, "peak": peak, "peak_source": which, "unit": "GB/s",
                    "frac": ach / peak, "traffic": alg * ncu_ratio[key],
                    "traffic_source": "algorithmic bytes x the dram__bytes ratio of the ncu --set full capture under profiles/",
                    "ms": ms, "algorithmic_bytes": alg,
                    "share_of_step": ms / max(sum(stages.values()), 1e-9)}
    if k3_bytes and stages.get("quotient"):
        ms = stages["quotient"]
        ach = k3_bytes / (ms / 1e3) / 1e9
        out["k3_quotient"] = {"bound": "hbm", "kernel": "quotient_values: generated per-chip constraint kernels qk (NVRTC) + LogUp constraints",
                              "achieved": ach, "peak": peak, "peak_source": which, "unit": "GB/s", "frac": ach / peak,
                              # the Keccak chip's kernel moves 7.79 GB of DRAM for 2.23 GB algorithmic (profiles/r02_qk_keccak_ncu.txt);
                              # it is 90 % of the stage's bytes, the ratio is applied to the whole stage
                              "traffic": k3_bytes * k3_dram_ratio if k3_dram_ratio else None,
                              "traffic_source": "algorithmic bytes x the dram__bytes ratio of the ncu --set full capture of the KeccakSponge kernel" if k3_dram_ratio else None,
                              "ms": ms, "algorithmic_bytes": k3_bytes, "share_of_step": ms / max(sum(stages.values()), 1e-9),
                              "note": "4*2n*(P+M+4E) + 16*2n bytes per chip (SURVEY.md section 8d); on the real KeccakSponge chip the kernel "
                                      "executes 471 k warp instructions per row-warp (3 788 constraints = 66 k node evaluations, 357 lookups): "
                                      "issue slots 42 % busy, fma-heavy pipe 50 %, DRAM 3.5 x the algorithmic bytes because every constraint "
                                      "family re-loads its columns (profiles/r02_qk_keccak_ncu.txt)"}
    if "k1_lde" in out:
        # arithmetic floor of the coset LDE under the same instruction prices: per input element one inverse and two
        # forward transforms = 1.5 log2(n) butterflies (Shoup product 8.2 + add 2.95 + sub about 1.3 cycles) and three
        # more Shoup products (four-step twiddles, coset scale); weighted by the elements of every table
        cyc = sum(h * w * (1.5 * max(1, int(h).bit_length() - 1) * 12.45 + 3 * 8.2) for h, w in trace_shapes.values())
        floor_ms = cyc / (148 * 4 * 32 * 1.965e9) * 1e3
        out["k1_lde"]["issue_roofline"] = {"model_floor_ms": floor_ms, "achieved_ms": out["k1_lde"]["ms"], "frac": floor_ms / out["k1_lde"]["ms"],
                                           "source": "butterfly arithmetic only, prices of tools/sweep/pipe_probe.cu"}
    if "k2_merkle" in out:
        rate = perms / (stages["commit_main_merkle"] / 1e3)
        out["k2_merkle"]["poseidon2_Gperm_per_s"] = rate / 1e9
        # the bound that does apply: integer issue.  Hard limit = one warp instruction per scheduler cycle; the kernel
        # executes about 4 450 warp instructions per permutation (ncu smsp__inst_executed / permutations,
        # profiles/r02_ncu_hot_kernels.txt), so frac = instructions / scheduler cycles spent = the issue-slot utilisation
        # (ncu: 2.45 instructions per cycle per SM of 4).  148 SMs x 4 schedulers x 32 lanes at 1965 MHz.
        lanes_hz = 148 * 4 * 32 * 1.965e9
        cyc = lanes_hz / rate
        out["k2_merkle"]["issue_roofline"] = {"warp_instructions_per_permutation": 4450, "scheduler_cycles_per_permutation": cyc,
                                              "frac": 4450 / cyc, "limit": "1 warp instruction per scheduler per cycle",
                                              "source": "profiles/r02_ncu_hot_kernels.txt (leaf_hash_kernel)"}
        out["k2_merkle"]["note"] = "integer-issue bound, not HBM bound: fma-heavy pipe 71 %, dominant stall math_pipe_throttle; see profiles/README.md"
    return out


def measure_cpu_baseline(args):
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--workload", args.workload, "--log-cpu", str(args.log_cpu), "--sample-log-cpu", str(args.sample_log_cpu),
                          "--constraint-density", str(args.constraint_density)],
                         capture_output=True, text=True, env={**os.environ, "RANK": "0", "WORLD_SIZE": "1"})
    try:
        return json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
    except Exception:
        return {"error": (out.stderr or out.stdout)[-300:]}


if __name__ == "__main__":
    main()
