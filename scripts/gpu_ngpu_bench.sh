#!/bin/bash
# chain-sharded bench at N GPUs (weak scaling): usage gpu_ngpu_bench.sh N
N=${1:-4}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['n_gpus','value','ms_per_step','leapfrog_steps_per_sec','rhat_max','gpu_launches']}, d['e2e'])
PY
tail -2 gpurun_out/bench_${N}gpu.err
