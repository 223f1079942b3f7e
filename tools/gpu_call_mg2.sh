#!/bin/bash
# multi-GPU session, final build (8 GPUs, short): config 3 (weak scaling) at N = 1 and N = 8 on the same box (resident and end to
# end, parity_check inside the bench line)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-preprocess > gpurun_out/final_c3_n1.json 2> gpurun_out/final_c3_n1.err
timeout 250 $TR --nproc-per-node 8 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_c3_n8.json 2> gpurun_out/final_c3_n8.err
for f in final_c3_n1 final_c3_n8; do python - $f <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/{sys.argv[1]}.json"))
    pc=d.get("parity_check") or {}
    print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", d["ms_per_step"], "parity", pc.get("flags_bit_equal"), pc.get("scores_bit_equal"), pc.get("ranks_covered"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1200:])
PY
done
