#!/bin/bash
# development aid: the 3072v workload alone; args: tag [VAR=value ...]
tag=$1; shift
env "$@" timeout 600 python bench.py --workload 3072v --steps 2 --warmup 2 --no-secondary --no-cpu-baseline --exe-matrices 0 > gpurun_out/q3_$tag.json 2> gpurun_out/q3_$tag.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/q3_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["value"], 2), {k: round(v, 1) for k, v in d["stage_ms_per_step"].items()}, d["phase_cycle_share"])
except Exception as e:
    print("$tag failed", e, open("gpurun_out/q3_$tag.err").read()[-1500:])
PY
