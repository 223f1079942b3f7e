#!/bin/bash
# round-end evidence set on one GPU: all GPU tests, the default bench line, the ncu launch list and one full capture
mkdir -p gpurun_out
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5) > gpurun_out/smoke.log 2>&1
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/pytest_gpu.log 2>&1
bash tools/ab_bench.sh cur=gpd_b200/libgpd_b200.so > gpurun_out/ab.log 2>&1
(GPD_B200_IMAGES_KERNEL=1 bash tools/ab_bench.sh general=gpd_b200/libgpd_b200.so) > gpurun_out/ab_general.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-preprocess > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_images|k_hands|k_conv1|k_conv2|k_ip1|k_frames" -c 9 -o gpurun_out/r2_round_end -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-preprocess --samples 16000 > gpurun_out/ncu.log 2>&1
timeout 200 python tools/phase_probe.py > gpurun_out/phase.log 2>&1
cat gpurun_out/smoke.log; tail -8 gpurun_out/pytest_gpu.log; cat gpurun_out/ab.log gpurun_out/ab_general.log; cat gpurun_out/phase.log
python - <<'PY'
import json
for f in ("bench_default","bench_reference"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), d.get("e2e",{}).get("value"), d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline",{}).get("cores"), d.get("cpu_baseline",{}).get("runs_samples_per_s"))
    except Exception as e: print(f, "FAILED", e)
PY
