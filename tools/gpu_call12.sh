#!/bin/bash
# A/B: two MMA issuer warps per CTA in conv1 / conv2 (default build) against one (build/ab/mw1.so); classifier parity of the default
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "classifier or tensor_core or krylon or synthetic_table or two_view" 2>&1 | tail -6) > gpurun_out/pytest_mw2.log 2>&1
bash tools/ab_bench.sh mw2=gpd_b200/libgpd_b200.so mw1=build/ab/mw1.so mw2b=gpd_b200/libgpd_b200.so > gpurun_out/ab_mw.log 2>&1
tail -4 gpurun_out/pytest_mw2.log; cat gpurun_out/ab_mw.log
