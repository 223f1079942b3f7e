#!/bin/bash
mkdir -p gpurun_out
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5) > gpurun_out/smoke.log 2>&1
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/pytest_gpu.log 2>&1
bash tools/ab_bench.sh cur=gpd_b200/libgpd_b200.so > gpurun_out/ab.log 2>&1
cat gpurun_out/smoke.log; tail -30 gpurun_out/pytest_gpu.log; cat gpurun_out/ab.log
