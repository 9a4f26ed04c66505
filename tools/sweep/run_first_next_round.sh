#!/bin/bash
# First GPU call of the next round (ROUND_NOTES.md): everything that was written after the last GPU run
# of round 1, in one box visit.  About 4 GPU-minutes.
#   gpurun --timeout 900 -- 'bash tools/sweep/run_first_next_round.sh'
mkdir -p gpurun_out
T=n1
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
timeout 300 python bench.py --stages --no-cpu-baseline > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err
# opt-in eval_columns variant: parity first, then its stage time (open_reduce) against the default's
ZKB200_EVAL_V2=1 timeout 300 python -m pytest tests -m gpu -q -k "proof or shard or commit or edge or reference_shapes" > gpurun_out/${T}_pytest_evalv2.log 2>&1
echo "pytest exit $?" >> gpurun_out/${T}_pytest_evalv2.log
ZKB200_EVAL_V2=1 timeout 300 python bench.py --stages --no-cpu-baseline --steps 6 > gpurun_out/${T}_bench_evalv2.json 2> gpurun_out/${T}_bench_evalv2.err
# opt-in NTT variant: parity of the transforms, then the LDE microbench both ways
ZKB200_NTT_PRETWIDDLE=1 timeout 200 python -m pytest tests -m gpu -q -k "lde or ntt or commit or shard_proof" > gpurun_out/${T}_pytest_pretw.log 2>&1
echo "pytest exit $?" >> gpurun_out/${T}_pytest_pretw.log
( timeout 60 python tools/microbench.py lde --log-n 18 --width 512
  ZKB200_NTT_PRETWIDDLE=1 timeout 60 python tools/microbench.py lde --log-n 18 --width 512 ) > gpurun_out/${T}_lde_pretw.jsonl 2>&1
tail -3 gpurun_out/${T}_pytest_pretw.log; cut -c1-120 gpurun_out/${T}_lde_pretw.jsonl
tail -3 gpurun_out/${T}_pytest.log gpurun_out/${T}_pytest_evalv2.log
python - <<'PY'
import json
for tag in ("default", "evalv2"):
    try:
        d = json.loads(open(f"gpurun_out/n1_bench_{tag}.json").read().strip().splitlines()[-1])
        print(tag, "value ms", round(d["ms_per_step"], 1), "e2e ms", round(d["e2e"]["ms_per_step"], 1), "open_reduce", round(d["stage_ms"]["open_reduce"], 2))
    except Exception as e:
        print(tag, "failed:", e)
PY
