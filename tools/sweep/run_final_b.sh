#!/bin/bash
# last verification of the round: GPU parity tests (incl. trace generation), smoke, trace-generation
# throughput, default bench line
mkdir -p gpurun_out
T=fb
timeout 240 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
timeout 60 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${T}_smoke.log
( timeout 60 python tools/microbench.py tracegen --chip ShiftRight --log-n 22
  timeout 60 python tools/microbench.py tracegen --chip ShiftRight --log-n 22 --col-major
  timeout 60 python tools/microbench.py tracegen --chip AddSub --log-n 23
  timeout 60 python tools/microbench.py tracegen --chip Lt --log-n 22 --col-major ) > gpurun_out/${T}_tracegen.jsonl 2>&1
timeout 300 python bench.py --stages > gpurun_out/${T}_bench_ours.json 2> gpurun_out/${T}_bench_ours.err
tail -4 gpurun_out/${T}_pytest.log; tail -2 gpurun_out/${T}_smoke.log; cut -c1-260 gpurun_out/${T}_tracegen.jsonl; cut -c1-200 gpurun_out/${T}_bench_ours.json
