#!/bin/bash
# round-end measurement set: GPU parity tests, both bench arms, per-launch time list of one warm proof
mkdir -p gpurun_out
T=${1:-fa}
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
timeout 200 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout 500 python bench.py --stages > gpurun_out/${T}_bench_ours.json 2> gpurun_out/${T}_bench_ours.err
if [ "$2" = "ncu" ]; then
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python tools/one_step.py 20 2 > gpurun_out/${T}_one_step.log 2>&1
fi
tail -2 gpurun_out/${T}_pytest.log; cut -c1-150 gpurun_out/${T}_bench_reference.json; cut -c1-150 gpurun_out/${T}_bench_ours.json
