#!/bin/bash
# bench.py under several environment settings (one per argument, "VAR=val[,VAR=val]"), headline numbers only
set -u
mkdir -p gpurun_out
for cfg in "$@"; do
  tag=$(echo "$cfg" | tr ',=' '__')
  env $(echo "$cfg" | tr ',' ' ') timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-batch --no-structured ${BENCH_ARGS:-} > gpurun_out/sweep_$tag.json 2> gpurun_out/sweep_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/sweep_$tag.json"))
    s = d.get("spmv", {}).get("variants", {})
    print("$cfg", "value %.1f e2e %.1f ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["detail"]["phase_ms_per_step"], d["detail"]["ring"],
          "frac %.3f" % d["roofline"]["frac"], {k: round(v["ms_flushed"] * 1e3, 1) for k, v in s.items()})
except Exception as e:
    print("$cfg", "failed", e, open("gpurun_out/sweep_$tag.err").read()[-400:])
PY
done
