#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/h2d_probe.py 16668 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s5_pytest.log 2>&1; tail -3 gpurun_out/s5_pytest.log
timeout 600 python bench.py --stages --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err; tail -2 gpurun_out/s5_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s5_bench.json'))
print('value',d['value'],d['ms_per_step'],'e2e',json.dumps(d['e2e']),'verified',d['verified'])
PY
ZKB200_PULL_CTAS=16 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-verify --no-pageable > gpurun_out/s5_bench_p16.json 2> gpurun_out/s5_bench_p16.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s5_bench_p16.json'))
print('PULL16: value',d['value'],d['ms_per_step'],'e2e',json.dumps(d['e2e']))
PY
