#!/bin/bash
# BASELINE.json configs on ONE B200: [0] fibonacci 2^16, [2] core 2^21 x 18 shards, [3] NTT sweep, [4] compress x 16,
# and the default bench line (configs[1]) with the same-config CPU arm.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
B="--no-pageable"
timeout 600 python bench.py --workload fibonacci --log-cpu 16 --steps 24 --warmup 3 $B > gpurun_out/r02_cfg0_fibonacci.json 2> gpurun_out/r02_cfg0.err; tail -1 gpurun_out/r02_cfg0.err
timeout 900 python bench.py --workload core --log-cpu 21 --shards 18 --warmup 3 --no-cpu-baseline $B > gpurun_out/r02_cfg2_core_1gpu.json 2> gpurun_out/r02_cfg2.err; tail -1 gpurun_out/r02_cfg2.err
timeout 900 python bench.py --workload compress --log-cpu 18 --shards 16 --warmup 3 --no-cpu-baseline $B > gpurun_out/r02_cfg4_compress_1gpu.json 2> gpurun_out/r02_cfg4.err; tail -1 gpurun_out/r02_cfg4.err
timeout 600 python tools/ntt_sweep.py > gpurun_out/r02_ntt_sweep.jsonl 2>&1; tail -3 gpurun_out/r02_ntt_sweep.jsonl
timeout 900 python bench.py --stages > gpurun_out/r02_bench_ours.json 2> gpurun_out/r02_bench_ours.err; tail -1 gpurun_out/r02_bench_ours.err
for f in r02_cfg0_fibonacci r02_cfg2_core_1gpu r02_cfg4_compress_1gpu r02_bench_ours; do
python - $f <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/%s.json'%sys.argv[1]))
    print(sys.argv[1], d['metric'], round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), 'e2e1', round(d['e2e']['one_shard_in_flight']['ms_per_step'],2), 'verified', d.get('verified'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
done
