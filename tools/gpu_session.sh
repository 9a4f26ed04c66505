#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s6_pytest.log 2>&1; tail -3 gpurun_out/s6_pytest.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --stages --steps 6 --warmup 3 --no-cpu-baseline --no-verify --no-pageable > gpurun_out/s6_bench_$name.json 2> gpurun_out/s6_bench_$name.err
  python - "$name" <<'PY'
import json,sys
d=json.load(open('gpurun_out/s6_bench_%s.json'%sys.argv[1]))
e=d['e2e']
print(sys.argv[1],'value',round(d['ms_per_step'],1),'e2e4',round(e['ms_per_step'],1),'e2e1',round(e['one_shard_in_flight']['ms_per_step'],1), 'open',round(d['stage_ms']['open_reduce'],2),'lde',round(d['stage_ms']['commit_main_lde'],2))
PY
}
run pull32 A=1
run excl8 ZKB200_PULL_EXCLUSIVE=1
run excl4 ZKB200_PULL_EXCLUSIVE=1 ZKB200_PULL_CTAS=4
run dma2d ZKB200_UPLOAD=dma2d
run dma ZKB200_UPLOAD=dma
