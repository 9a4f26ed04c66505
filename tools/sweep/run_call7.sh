#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/c7_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c7_pytest.log
timeout 400 python bench.py --stages --no-cpu-baseline > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
tail -2 gpurun_out/c7_pytest.log; tail -2 gpurun_out/c7_bench.err | cut -c1-900
