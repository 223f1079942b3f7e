#!/bin/bash
# multi-GPU session, final build (8 GPUs): the scaling runs the driver makes — config 3 (weak) at N = 1, 2, 4, 8 back to back on one
# box — plus config 4 (strong) at N = 8 and the in-library parity check at N = 8
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { # name nproc args...
  local name=$1 np=$2; shift 2
  timeout 300 $TR --nproc-per-node $np --master-port $((29520 + RANDOM % 200)) bench.py --gpus $np "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
}
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-preprocess > gpurun_out/final_c3_n1.json 2> gpurun_out/final_c3_n1.err
run final_c3_n8 8 --steps 5 --warmup 3 --no-cpu-baseline
run final_c3_n2 2 --steps 5 --warmup 3 --no-cpu-baseline
run final_c3_n4 4 --steps 5 --warmup 3 --no-cpu-baseline
run final_c4_n8 8 --config 4 --steps 3 --warmup 3 --no-cpu-baseline
timeout 200 $TR --nproc-per-node 8 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/multi_check_n8_final.log 2>&1
grep -h "multi_gpu_check" gpurun_out/multi_check_n8_final.log | sort | tail -10
for f in final_c3_n1 final_c3_n2 final_c3_n4 final_c3_n8 final_c4_n8; do python - $f <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/{sys.argv[1]}.json"))
    pc=d.get("parity_check") or {}
    print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", d["ms_per_step"], d["stage_ms_per_step"], "parity", pc.get("flags_bit_equal"), pc.get("scores_bit_equal"), pc.get("ranks_covered"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1200:])
PY
done
