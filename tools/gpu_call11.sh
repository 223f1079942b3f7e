#!/bin/bash
# final evidence set of round 2 on one GPU: smoke, all GPU tests, the default bench line + the reference arm, the ncu launch list,
# one full capture of the hot kernels, the k_images2 phase probe
mkdir -p gpurun_out
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5) > gpurun_out/smoke.log 2>&1
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/pytest_gpu.log 2>&1
timeout 300 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-preprocess > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_images|k_hands|k_conv1|k_conv2|k_ip1|k_frames" -c 24 -o gpurun_out/r2_final -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-preprocess --samples 16000 > gpurun_out/ncu.log 2>&1
timeout 200 python tools/phase_probe.py > gpurun_out/phase.log 2>&1
cat gpurun_out/smoke.log; tail -8 gpurun_out/pytest_gpu.log; cat gpurun_out/phase.log
python - <<'PY'
import json
for f in ("bench_default","bench_reference"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), d.get("e2e",{}).get("value"), d.get("ms_per_step"), d.get("stage_ms_per_step"), d.get("roofline"), d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline",{}).get("cores"))
    except Exception as e: print(f, "FAILED", e)
PY
