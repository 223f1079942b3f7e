#!/bin/bash
# GPU session 2 (2 GPUs): tests, 1-GPU bench, launch list, multi-GPU parity + 2-GPU bench lines (configs 3, 4, 5)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/pytest_gpu.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-preprocess > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-preprocess > gpurun_out/ncu_bench.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tools/multi_gpu_check.py > gpurun_out/multi_check.log 2>&1
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2.err
timeout 300 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err
timeout 300 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err
tail -4 gpurun_out/pytest_gpu.log; grep -h "multi_gpu_check" gpurun_out/multi_check.log | tail -4
for f in bench_n1 bench_n2 bench_c4_n1 bench_c4_n2 bench_c5_n1; do python - $f <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/{sys.argv[1]}.json"))
    print(sys.argv[1], round(d["value"]), round(d["e2e"]["value"]), d["stage_ms_per_step"], d.get("parity_check"), d["gpu_launches"])
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1500:])
PY
done
