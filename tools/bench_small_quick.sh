#!/bin/bash
# development aid: the 64n workload alone; args: tag [VAR=value ...]
tag=$1; shift
env "$@" timeout 300 python bench.py --workload 64n --steps 5 --warmup 3 --no-secondary --no-cpu-baseline --exe-matrices 0 > gpurun_out/qs_$tag.json 2> gpurun_out/qs_$tag.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/qs_$tag.json"))
    print("$tag", round(d["value"]), {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()})
except Exception as e:
    print("$tag failed", e, open("gpurun_out/qs_$tag.err").read()[-800:])
PY
