#!/bin/bash
# GPU session 3 (1 GPU): new k_images (P16) + conv1 bulk copy + conv2 converter warps: smoke, stats, tests, bench, ncu
mkdir -p gpurun_out
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15) > gpurun_out/smoke.log 2>&1
(timeout 400 python tools/gpu_check.py 2>&1 | tail -80) > gpurun_out/gpu_check.log 2>&1
for t in tests/test_gpu_parity.py tests/test_gpu_preprocess.py tests/test_host_cpp.py tests/test_weights_io.py; do
  (timeout 600 python -m pytest $t -m gpu -q 2>&1 | tail -40) > gpurun_out/pytest_$(basename $t .py).log 2>&1
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-preprocess > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err
timeout 200 python tools/phase_probe.py > gpurun_out/phase.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_images|k_conv2|k_conv1" -c 3 -o gpurun_out/r2_call3 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-preprocess --samples 16000 > gpurun_out/ncu.log 2>&1
cat gpurun_out/smoke.log; tail -30 gpurun_out/gpu_check.log
for t in test_gpu_parity test_gpu_preprocess test_host_cpp test_weights_io; do echo "== $t"; tail -6 gpurun_out/pytest_$t.log; done
for f in bench_n1 bench_c5_n1; do python - $f <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/{sys.argv[1]}.json"))
    print(sys.argv[1], round(d["value"]), round(d["e2e"]["value"]), d["stage_ms_per_step"], {k: v["ms_per_step"] for k, v in d["kernels"].items()}, d["gpu_launches"])
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1500:])
PY
done
cat gpurun_out/phase.log
