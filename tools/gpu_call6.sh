#!/bin/bash
mkdir -p gpurun_out
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5) > gpurun_out/smoke.log 2>&1
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -30) > gpurun_out/pytest_test_gpu_parity.log 2>&1
(timeout 300 python tools/gpu_check.py 2>&1 | grep -E "===|images|scores|flags") > gpurun_out/gpu_check.log 2>&1
bash tools/ab_bench.sh cur=gpd_b200/libgpd_b200.so > gpurun_out/ab.log 2>&1
(GPD_B200_IMAGES_KERNEL=1 bash tools/ab_bench.sh general=gpd_b200/libgpd_b200.so) > gpurun_out/ab_general.log 2>&1
timeout 200 python tools/phase_probe.py > gpurun_out/phase.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_images2" -c 1 -o gpurun_out/r2_call6 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-preprocess --samples 16000 > gpurun_out/ncu.log 2>&1
cat gpurun_out/smoke.log; tail -12 gpurun_out/pytest_test_gpu_parity.log
cat gpurun_out/gpu_check.log | head -30; cat gpurun_out/ab.log gpurun_out/ab_general.log; cat gpurun_out/phase.log
