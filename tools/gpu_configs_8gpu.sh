#!/bin/bash
# BASELINE.json configs[2] and [4] over N GPUs of one box (torchrun, one rank per GPU), plus the default weak-scaling line.
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
B="--no-pageable --no-cpu-baseline"
timeout 900 $TR bench.py --gpus $N --workload core --log-cpu 21 --shards 18 --warmup 2 $B > gpurun_out/r02_cfg2_core_${N}gpu.json 2> gpurun_out/r02_cfg2_${N}.err; tail -1 gpurun_out/r02_cfg2_${N}.err
timeout 900 $TR bench.py --gpus $N --workload compress --log-cpu 18 --shards 127 --warmup 2 $B > gpurun_out/r02_cfg4_compress_${N}gpu.json 2> gpurun_out/r02_cfg4_${N}.err; tail -1 gpurun_out/r02_cfg4_${N}.err
timeout 900 $TR bench.py --gpus $N --steps 6 --warmup 3 $B > gpurun_out/r02_bench_ours_${N}gpu.json 2> gpurun_out/r02_bench_${N}.err; tail -1 gpurun_out/r02_bench_${N}.err
timeout 600 python -m pytest tests -m gpu -x -q -k "multi_device" > gpurun_out/r02_multidev_${N}.log 2>&1; tail -2 gpurun_out/r02_multidev_${N}.log
for f in r02_cfg2_core_${N}gpu r02_cfg4_compress_${N}gpu r02_bench_ours_${N}gpu; do
python - $f <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d['metric'], d['scaling'], round(d['value'],2), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],2), 'verified', d.get('verified'))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
done
