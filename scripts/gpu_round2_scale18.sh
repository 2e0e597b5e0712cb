#!/bin/bash
# final build, 8 GPUs of one box: configs[3] strong-scaled at 1 and 8 GPUs (the 2- and 4-GPU points: scale_r02_final_strong_{2,4}.json)
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
timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/scale_r02_final2_strong_1.json 2> gpurun_out/scale_r02_final2_strong_1.err; show scale_r02_final2_strong_1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/scale_r02_final2_strong_8.json 2> gpurun_out/scale_r02_final2_strong_8.err; show scale_r02_final2_strong_8
