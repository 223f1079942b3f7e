#!/bin/bash
# multi-GPU session (8 GPUs): in-library sharding parity at N = 8, config 4 (1 M samples, strong scaling) at N = 8 / 4 / 2,
# config 5 (two views, 12 channels) at N = 4, config 3 (weak) at N = 8; the 2-GPU CLI test
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/multi_check_n8.log 2>&1
(timeout 300 python -m pytest tests/test_host_cpp.py -m gpu -q -k "two_gpus" 2>&1 | tail -5) > gpurun_out/pytest_cli_2gpu.log 2>&1
run() { # name nproc args...
  local name=$1 np=$2; shift 2
  timeout 400 $TR --nproc-per-node $np --master-port $((29520 + RANDOM % 200)) bench.py --gpus $np "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
}
run bench_c4_n8 8 --config 4 --steps 3 --warmup 3 --no-cpu-baseline
run bench_c4_n4 4 --config 4 --steps 3 --warmup 3 --no-cpu-baseline
run bench_c4_n2 2 --config 4 --steps 3 --warmup 3 --no-cpu-baseline
timeout 300 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err
run bench_c5_n4 4 --config 5 --steps 3 --warmup 3 --no-cpu-baseline
run bench_c3_n8 8 --steps 5 --warmup 3 --no-cpu-baseline
grep -h "multi_gpu_check" gpurun_out/multi_check_n8.log | sort | tail -16; tail -3 gpurun_out/pytest_cli_2gpu.log
for f in bench_c4_n1 bench_c4_n2 bench_c4_n4 bench_c4_n8 bench_c5_n4 bench_c3_n8; do python - $f <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/{sys.argv[1]}.json"))
    pc=d.get("parity_check") or {}
    print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["stage_ms_per_step"], "parity", pc.get("flags_bit_equal"), pc.get("scores_bit_equal"), pc.get("ranks_covered"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1200:])
PY
done
