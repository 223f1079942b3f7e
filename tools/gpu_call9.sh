#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "reevaluate or kernels_agree or dense" 2>&1 | tail -8) > gpurun_out/pytest_sel.log 2>&1
for v in nt384 nt256; do (GPD_B200_LIB=build/ab/$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "kernels_agree or krylon or synthetic_table" 2>&1 | tail -4) > gpurun_out/pytest_$v.log 2>&1; done
bash tools/ab_bench.sh cur=gpd_b200/libgpd_b200.so nt384=build/ab/nt384.so nt256=build/ab/nt256.so > gpurun_out/ab.log 2>&1
GPD_B200_LIB=build/ab/nt384.so timeout 200 python tools/phase_probe.py > gpurun_out/phase_384.log 2>&1
GPD_B200_LIB=build/ab/nt256.so timeout 200 python tools/phase_probe.py > gpurun_out/phase_256.log 2>&1
tail -4 gpurun_out/pytest_sel.log gpurun_out/pytest_nt384.log gpurun_out/pytest_nt256.log; cat gpurun_out/ab.log; cat gpurun_out/phase_384.log gpurun_out/phase_256.log
