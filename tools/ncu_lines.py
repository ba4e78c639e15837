#!/usr/bin/env python
"""Development aid: joins an ncu '--page source --csv' SASS dump with 'nvdisasm -g -c' line info of the same
cubin and prints the hottest source lines (warp stall samples, instructions executed) of one kernel.
usage: ncu_lines.py <ncu_source.csv> <nvdisasm.txt> <kernel-substring> [top]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis_txt, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# address -> (file, line) from nvdisasm
addr2line = {}
infunc = False
cur = ("?", 0)
for ln in open(dis_txt, errors="ignore"):
    if ln.startswith("//---------------------") and ".text." in ln:
        infunc = kern in ln
        cur = ("?", 0)
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0])
tot_s = tot_i = 0
base = None
for r in rows[2:]:
    try:
        a = int(r[ci["Address"]], 16) if r[ci["Address"]].startswith("0x") else int(r[ci["Address"]])
        s = int(r[ci["# Samples"]])
        ie = int(r[ci["Instructions Executed"]])
    except Exception:
        continue
    if base is None:
        base = a
    key = addr2line.get(a - base, (("?", 0), ""))[0]
    agg[key][0] += s
    agg[key][1] += ie
    tot_s += s
    tot_i += ie
print("total samples %d, instructions executed %d" % (tot_s, tot_i))
for key, (s, ie) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%6.2f%% samples  %6.2f%% instr   %s:%d" % (100.0 * s / max(tot_s, 1), 100.0 * ie / max(tot_i, 1), key[0], key[1]))
