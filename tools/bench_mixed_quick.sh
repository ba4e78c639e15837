#!/bin/bash
# development aid: the mixed workload alone (N=1); args: tag M [VAR=value ...]
tag=$1; M=$2; shift; shift
env "$@" timeout 900 python bench.py --workload mixed --mixed-m $M --no-cpu-baseline > gpurun_out/qm_$tag.json 2> gpurun_out/qm_$tag.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/qm_$tag.json").read().strip().splitlines()[-1])
    m = d["mixed"]
    print("$tag", round(m["value"], 1), m["rank0_stage_ms"], "failed", m["failed"])
except Exception as e:
    print("$tag failed", e, open("gpurun_out/qm_$tag.err").read()[-1500:])
PY
