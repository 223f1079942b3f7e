#!/bin/bash
# conv2 with 65 K-chunks (paired taps for channels 16..19) + two MMA issuer warps: classifier parity + bench
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "classifier or tensor_core or krylon or synthetic_table or two_view or stage_entry or other_channel" 2>&1 | tail -6) > gpurun_out/pytest_c2.log 2>&1
bash tools/ab_bench.sh c2k65=gpd_b200/libgpd_b200.so > gpurun_out/ab_c2.log 2>&1
tail -4 gpurun_out/pytest_c2.log; cat gpurun_out/ab_c2.log
