#!/bin/bash
mkdir -p gpurun_out
( ZKB200_NTT_STRIDED_TILE_LOG=12 timeout 100 python tools/microbench.py lde --log-n 18 --width 512
  ZKB200_NTT_CONTIG_TILE_LOG=11 timeout 100 python tools/microbench.py lde --log-n 18 --width 512
  ZKB200_NTT_CONTIG_TILE_LOG=13 timeout 100 python tools/microbench.py lde --log-n 18 --width 512 ) > gpurun_out/c9_micro.jsonl 2>&1
timeout 300 python bench.py --no-cpu-baseline --value-threads 4 --e2e-threads 5 > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
cat gpurun_out/c9_micro.jsonl | cut -c1-120; cut -c1-120 gpurun_out/c9_bench.json; tail -2 gpurun_out/c9_bench.err
