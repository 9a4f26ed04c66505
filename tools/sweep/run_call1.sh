#!/bin/bash
# one gpurun call: Poseidon2 variant sweep, GPU parity tests, kernel microbenches, bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt 2>&1
for b in tools/sweep/build/p2_*; do timeout 60 $b 21 8; done > gpurun_out/c1_p2_sweep.jsonl 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c1_pytest.log
( timeout 120 python tools/microbench.py permute --log-n 22
  timeout 120 python tools/microbench.py mmcs --log-n 19 --width 512
  timeout 120 python tools/microbench.py lde --log-n 18 --width 512 ) > gpurun_out/c1_micro.jsonl 2>&1
timeout 500 python bench.py --stages > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
tail -3 gpurun_out/c1_pytest.log; cat gpurun_out/c1_p2_sweep.jsonl | cut -c1-160; cat gpurun_out/c1_micro.jsonl | cut -c1-250; cut -c1-400 gpurun_out/c1_bench.json
