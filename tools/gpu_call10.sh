#!/bin/bash
mkdir -p gpurun_out
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5) > gpurun_out/smoke.log 2>&1
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log 2>&1
bash tools/ab_bench.sh overlap=gpd_b200/libgpd_b200.so > gpurun_out/ab.log 2>&1
(GPD_B200_OVERLAP=0 bash tools/ab_bench.sh serial=gpd_b200/libgpd_b200.so) > gpurun_out/ab_serial.log 2>&1
bash tools/ab_bench.sh overlap2=gpd_b200/libgpd_b200.so > gpurun_out/ab2.log 2>&1
cat gpurun_out/smoke.log; tail -6 gpurun_out/pytest_gpu.log; cat gpurun_out/ab.log gpurun_out/ab_serial.log gpurun_out/ab2.log
python - <<'PY'
import json
for f in ("overlap","serial"):
    d=json.load(open(f"gpurun_out/ab_{f}.json")); print(f, d["ms_per_step"], d["e2e"]["value"], d["stage_ms_per_step"])
PY
