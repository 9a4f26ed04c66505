#!/bin/bash
set -u
mkdir -p gpurun_out
# launch list of two warm proofs at the bench size (cheap: one metric)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches.csv python tools/one_step.py 20 2 > gpurun_out/r02_launches.log 2>&1
tail -3 gpurun_out/r02_launches.log
# full sections for the hot kernels at a reduced size (replays save/restore the written buffers)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'reduce_matrix|eval_columns_v2|^qk$|ntt_contig_lean|ntt_strided_lean|leaf_absorb|leaf_hash_kernel|compress_kernel|transpose_kernel' -c 260 -f -o gpurun_out/r02_full python tools/one_step.py 18 1 > gpurun_out/r02_full.log 2>&1
tail -3 gpurun_out/r02_full.log
ls -la gpurun_out/r02_full.ncu-rep
