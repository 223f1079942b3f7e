#!/bin/bash
# A/B comparison of library builds on the GPU box: tools/ab_bench.sh <label>=<path.so> ...  (development aid)
mkdir -p gpurun_out
for spec in "$@"; do
  label="${spec%%=*}"; so="${spec#*=}"
  GPD_B200_LIB="$so" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-preprocess > gpurun_out/ab_$label.json 2> gpurun_out/ab_$label.err
  python - "$label" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/ab_{sys.argv[1]}.json"))
print(sys.argv[1], round(d["value"]), round(d["e2e"]["value"]), d["stage_ms_per_step"], {k: v["ms_per_step"] for k, v in d["kernels"].items()})
PY
done
