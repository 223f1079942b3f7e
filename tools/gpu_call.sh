#!/bin/bash
# One GPU-box session (development aid): GPU tests, A/B bench of library builds, k_images phase probe, ncu capture.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/pytest_gpu.log 2>&1
bash tools/ab_bench.sh "$@" > gpurun_out/ab.log 2>&1
timeout 200 python tools/phase_probe.py > gpurun_out/phase.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_images|k_conv2" -c 4 -o gpurun_out/r2_call1 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-preprocess --samples 16000 > gpurun_out/ncu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/ab.log; cat gpurun_out/phase.log
