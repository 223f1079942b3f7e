#!/bin/bash
# GPU session 4 (1 GPU): k_images with DPX assemble / ballot occupancy; NT = 512 vs 1024; device clustering test
mkdir -p gpurun_out
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5) > gpurun_out/smoke.log 2>&1
for t in tests/test_gpu_parity.py tests/test_host_cpp.py; do
  (timeout 600 python -m pytest $t -m gpu -q 2>&1 | tail -30) > gpurun_out/pytest_$(basename $t .py).log 2>&1
done
(GPD_B200_LIB=build/ab/nt1024.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -8) > gpurun_out/pytest_nt1024.log 2>&1
bash tools/ab_bench.sh nt512=gpd_b200/libgpd_b200.so nt1024=build/ab/nt1024.so > gpurun_out/ab.log 2>&1
timeout 200 python tools/phase_probe.py > gpurun_out/phase.log 2>&1
GPD_B200_LIB=build/ab/nt1024.so timeout 200 python tools/phase_probe.py > gpurun_out/phase_1024.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_images" -c 1 -o gpurun_out/r2_call4 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-preprocess --samples 16000 > gpurun_out/ncu.log 2>&1
cat gpurun_out/smoke.log
for t in test_gpu_parity test_host_cpp nt1024; do echo "== $t"; tail -6 gpurun_out/pytest_$t.log; done
cat gpurun_out/ab.log; cat gpurun_out/phase.log; cat gpurun_out/phase_1024.log
