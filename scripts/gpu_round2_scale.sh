#!/bin/bash
# round 2 final, 8 GPUs of one box: strong scaling of BASELINE.json configs[3] (65 536 chains in total) at 1 / 2 / 4 / 8 GPUs
# and the weak-scaled 8-GPU line.  Usage: gpurun --gpus 8 -- bash scripts/gpu_round2_scale.sh
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/{n}.json").read().splitlines() if l.startswith("{")][-1])
    print(n, {k: d.get(k) for k in ("value","ms_per_step","n_gpus","scaling","rhat_max")}, "e2e", d.get("e2e") and round(d["e2e"]["value"]))
except Exception as e:
    print(n, "ERR", e); print(open(f"gpurun_out/{n}.err").read()[-1200:])
PY
}
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_r02_final_strong_1.json 2> gpurun_out/scale_r02_final_strong_1.err; show scale_r02_final_strong_1
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_r02_final_strong_$n.json 2> gpurun_out/scale_r02_final_strong_$n.err; show scale_r02_final_strong_$n
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --scaling weak > gpurun_out/scale_r02_final_weak_8.json 2> gpurun_out/scale_r02_final_weak_8.err; show scale_r02_final_weak_8
