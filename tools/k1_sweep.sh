#!/bin/bash
# K1 tuning knob sweep: prints value / ms per step for each PMK_K1_MINB
for m in 4 3; do
  PMK_K1_MINB=$m timeout 300 python bench.py --steps 100 --warmup 5 --pipeline-iters 0 --no-cpu-baseline 2>/dev/null > /tmp/k1_$m.json
  python - "$m" <<'PY'
import json, sys
m = sys.argv[1]
d = json.loads(open(f"/tmp/k1_{m}.json").read().strip().splitlines()[-1])
print("minb", m, d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
PY
done
