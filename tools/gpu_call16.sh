#!/bin/bash
# image stage: per-count reciprocal table for the cell means, saturating conversion instead of the clamp: parity + bench
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/pytest_img.log 2>&1
bash tools/ab_bench.sh img=gpd_b200/libgpd_b200.so > gpurun_out/ab_img.log 2>&1
tail -4 gpurun_out/pytest_img.log; cut -c1-420 gpurun_out/ab_img.log
