#!/bin/bash
# usage: tools/quick_bench.sh <tag> [VAR=value ...]  -- short 512v bench (no secondary / cpu baseline / exe leg), prints the stage split
tag=$1; shift
env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline --exe-matrices 0 > gpurun_out/qb_$tag.json 2> gpurun_out/qb_$tag.err
echo "$tag rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/qb_$tag.json"))
    print("$tag", round(d["value"], 1), {k: round(v, 1) for k, v in d["stage_ms_per_step"].items()}, d["phase_cycle_share"])
except Exception as e:
    print("$tag failed", e, open("gpurun_out/qb_$tag.err").read()[-1500:])
PY
