#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s10_pytest.log 2>&1; tail -3 gpurun_out/s10_pytest.log
timeout 600 python tools/stage_ab.py 20 qk_block=128 quotient_codegen=0 > gpurun_out/s10_stage_ab.jsonl 2>&1; cut -c1-400 gpurun_out/s10_stage_ab.jsonl
run() { name=$1; shift
  timeout 600 python bench.py --stages --steps 6 --warmup 3 --no-cpu-baseline --no-verify --no-pageable "$@" > gpurun_out/s10_bench_$name.json 2> gpurun_out/s10_bench_$name.err
  python - "$name" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/s10_bench_%s.json'%sys.argv[1]))
    e=d['e2e']
    print(sys.argv[1],'value',round(d['ms_per_step'],1),'e2e',round(e['ms_per_step'],1),'e2e1',round(e['one_shard_in_flight']['ms_per_step'],1))
except Exception as ex:
    print(sys.argv[1],'FAILED',ex); print(open('gpurun_out/s10_bench_%s.err'%sys.argv[1]).read()[-800:])
PY
}
run pull32
ZKB200_PULL_CTAS=24 run pull24
ZKB200_PULL_CTAS=16 run pull16
