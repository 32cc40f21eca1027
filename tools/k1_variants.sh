#!/bin/bash
# A/B runs of the N=1 kernel-only bench under different env / library variants (gpurun helper).
# usage: tools/k1_variants.sh "name:ENV=.. ENV=.." ...
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  env $envs timeout 300 python bench.py --steps 24 --warmup 5 --no-cpu --no-e2e > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/var_%s.json" % name).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(name, "G/s=%.2f" % (d["value"] / 1e9), "ms/tick=%.4f" % d["ms_per_step"], "serial", {k: round(v, 4) for k, v in r["serial_tick_phase_ms"].items()},
          "detail", {k: round(v, 4) for k, v in (r.get("serial_tick_detail_ms") or {}).items()}, "residue=%.3f" % r.get("residue_fraction", -1), "drains", r.get("pipeline_drains"))
except Exception as e:
    print(name, "FAILED", e, open("gpurun_out/var_%s.err" % name).read()[-500:])
PY
done
