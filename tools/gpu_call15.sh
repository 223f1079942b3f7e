#!/bin/bash
# conv1 with 5 / 6 / 7 epilogue groups (TMEM accumulator buffers): classifier parity of the default + bench of each
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "classifier or tensor_core or krylon or other_channel" 2>&1 | tail -6) > gpurun_out/pytest_c1.log 2>&1
bash tools/ab_bench.sh ng6=gpd_b200/libgpd_b200.so ng5=build/ab/ng5.so ng7=build/ab/ng7.so > gpurun_out/ab_c1.log 2>&1
tail -4 gpurun_out/pytest_c1.log; cat gpurun_out/ab_c1.log
