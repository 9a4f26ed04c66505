#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/sweep/build/pipe_probe > gpurun_out/c2_pipe.jsonl 2>&1
for b in tools/sweep/build/p2_*; do timeout 60 $b 21 8; done > gpurun_out/c2_p2_sweep.jsonl 2>&1
( timeout 120 python tools/microbench.py permute --log-n 22
  timeout 120 python tools/microbench.py mmcs --log-n 19 --width 512 ) > gpurun_out/c2_micro.jsonl 2>&1
timeout 500 python bench.py --stages --no-cpu-baseline > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
cat gpurun_out/c2_pipe.jsonl | cut -c1-200
