#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/c6_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c6_pytest.log
( timeout 120 python tools/microbench.py lde --log-n 18 --width 512
  timeout 120 python tools/microbench.py lde --log-n 20 --width 128
  timeout 120 python tools/microbench.py ntt --log-n 20 --width 128 ) > gpurun_out/c6_micro.jsonl 2>&1
timeout 400 python bench.py --stages --no-cpu-baseline > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'eval_columns_kernel|reduce_matrix_kernel|quotient_kernel' -c 16 -o gpurun_out/c6_open -f python tools/one_step.py 18 1 > gpurun_out/c6_ncu.log 2>&1
tail -2 gpurun_out/c6_pytest.log; cat gpurun_out/c6_micro.jsonl | cut -c1-200; tail -3 gpurun_out/c6_ncu.log
