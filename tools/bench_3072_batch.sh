#!/bin/bash
# args: tag batch [VAR=value...]
tag=$1; b=$2; shift; shift
env "$@" timeout 500 python bench.py --workload 3072v --steps 2 --warmup 2 --batch $b --distinct 8 --no-secondary --no-cpu-baseline --exe-matrices 0 --e2e-chunks 0 > gpurun_out/q3_$tag.json 2> gpurun_out/q3_$tag.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/q3_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["value"], 2), {k: round(v, 1) for k, v in d["stage_ms_per_step"].items()})
except Exception as e:
    print("$tag failed", e, open("gpurun_out/q3_$tag.err").read()[-800:])
PY
