#!/bin/bash
# development aid: paladin-gpu on N copies of 16 distinct 512x512 files with per-batch stage times; args: N [extra exe flags]
N=${1:-592}; shift
D=$(mktemp -d /tmp/paladin_exe_XXXX)
python - "$D" "$N" <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import bench
d, n = sys.argv[1], int(sys.argv[2])
ii, jj, vals = bench.synth_coo(512, 16, 3)
for k in range(16):
    open(os.path.join(d, "d%d.matrix" % k), "wb").write(bench.mm_text(512, ii, jj, vals[k]))
names = []
for k in range(n):
    nm = "m%05d.matrix" % k
    os.link(os.path.join(d, "d%d.matrix" % (k % 16)), os.path.join(d, nm))
    names.append(nm)
open(os.path.join(d, "listing"), "w").write("\n".join(names))
PY
PALADIN_GPU_TIMING=1 paladin_b200/lib/paladin-gpu --listing=$D/listing --rootdir=$D --right --gpus=1 "$@" 2>&1 | grep -E "batch@|^max|^avg|load imb" 
rm -rf $D
