#!/bin/bash
# launch list of one warm proof + full captures of the non-hash kernels of the largest chip
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c5_launches.csv python tools/one_step.py 20 2 > gpurun_out/c5_one_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'eval_columns|reduce_matrix|quotient_kernel|ntt_strided|ntt_contig|transpose' -s 200 -c 120 -o gpurun_out/c5_full -f python tools/one_step.py 20 2 > gpurun_out/c5_full.log 2>&1
ls -la gpurun_out/ | tail -5
