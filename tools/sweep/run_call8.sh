#!/bin/bash
mkdir -p gpurun_out
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'ntt_strided_kernel|ntt_contig_kernel' -c 3 -o gpurun_out/c8_lde -f python tools/microbench.py lde --log-n 18 --width 256 --reps 1 --warmup 0 > gpurun_out/c8_lde.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'leaf_hash_kernel' -c 1 -o gpurun_out/c8_leaf -f python tools/microbench.py mmcs --log-n 19 --width 512 --reps 1 --warmup 0 > gpurun_out/c8_leaf.log 2>&1
tail -2 gpurun_out/c8_lde.log gpurun_out/c8_leaf.log
