#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s11_pytest.log 2>&1; tail -3 gpurun_out/s11_pytest.log
run() { name=$1; shift
  timeout 600 python bench.py --stages --steps 6 --warmup 3 --no-cpu-baseline --no-verify --no-pageable "$@" > gpurun_out/s11_bench_$name.json 2> gpurun_out/s11_bench_$name.err
  python - "$name" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/s11_bench_%s.json'%sys.argv[1]))
    e=d['e2e']
    print(sys.argv[1],'value',round(d['ms_per_step'],1),'e2e',round(e['ms_per_step'],1),'e2e1',round(e['one_shard_in_flight']['ms_per_step'],1),'open',round(d['stage_ms']['open_reduce'],2))
except Exception as ex:
    print(sys.argv[1],'FAILED',ex); print(open('gpurun_out/s11_bench_%s.err'%sys.argv[1]).read()[-800:])
PY
}
run split
ZKB200_PULL_SPLIT=0 run nosplit
timeout 600 python tools/ntt_sweep.py > gpurun_out/r02_ntt_sweep.jsonl 2>&1; tail -2 gpurun_out/r02_ntt_sweep.jsonl | cut -c1-200
