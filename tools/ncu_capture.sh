#!/bin/bash
# Round-2 ncu evidence for K1 (run under gpurun on ONE GPU): a launch list of a short bench run, then ONE
# `--set full` capture of exactly the kernels of the last timed tick, exported as a raw CSV.
#   tools/ncu_capture.sh <tag> [ENV=VALUE ...]        e.g. tools/ncu_capture.sh r02_sort   /   r02_index GCRA_ADAPTIVE=0
# Outputs under gpurun_out/: <tag>_launches.csv, <tag>_k1_full.ncu-rep, <tag>_k1_full_raw.csv
set -u
tag=$1; shift
ARGS="bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-sweep --sustain-sec 0"
env "$@" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${tag}_launches.csv python $ARGS > gpurun_out/${tag}_launches.log 2>&1
read skip count <<< $(python - gpurun_out/${tag}_launches.csv <<'PY'
import csv, sys
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("==")))
names = [r["Kernel Name"] for r in rows]
starts = [i for i, n in enumerate(names) if n.startswith("void gcra::ingest_kernel") or n.startswith("void gcra::probe_kernel")
          or "ingest_kernel<" in n or "probe_kernel<" in n]
# the last start is the serial phase tick, the one before it the last timed tick
a, b = starts[-2], starts[-1]
print(a, b - a)
PY
)
echo "last timed tick: launches $skip .. +$count"
env "$@" timeout 900 ncu --set full --clock-control none --import-source on --launch-skip $skip --launch-count $count -f -o gpurun_out/${tag}_k1_full python $ARGS > gpurun_out/${tag}_full.log 2>&1
ncu -i gpurun_out/${tag}_k1_full.ncu-rep --page raw --csv > gpurun_out/${tag}_k1_full_raw.csv 2>/dev/null
wc -l gpurun_out/${tag}_k1_full_raw.csv
