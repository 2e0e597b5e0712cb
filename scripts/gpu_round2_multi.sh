#!/bin/bash
# round 2, 8 GPUs of one box: strong / weak scaling of BASELINE.json configs[3], configs[4] (10 M x 64 row-sharded, 64 chains),
# and the 2-rank row-shard parity script.  Usage: gpurun --gpus 8 -- bash scripts/gpu_round2_multi.sh
mkdir -p gpurun_out
tr() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) "$@"; }
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
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_r02_strong_1.json 2> gpurun_out/scale_r02_strong_1.err; show scale_r02_strong_1
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_r02_strong_$n.json 2> gpurun_out/scale_r02_strong_$n.err; show scale_r02_strong_$n
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --scaling weak > gpurun_out/scale_r02_weak_8.json 2> gpurun_out/scale_r02_weak_8.err; show scale_r02_weak_8
# configs[4]: 8 x 1.25 M rows
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29660 scripts/gpu_cfg5_shard.py > gpurun_out/cfg5_8gpu.log 2> gpurun_out/cfg5_8gpu.err
grep '^{' gpurun_out/cfg5_8gpu.log | tail -1 > gpurun_out/cfg5_8gpu.json; cut -c1-1500 gpurun_out/cfg5_8gpu.json
grep -E "NCCL INFO.*(Algo|algo|NVLS|Connected|nvls|Ring|Tree)" gpurun_out/cfg5_8gpu.log gpurun_out/cfg5_8gpu.err | head -12 > gpurun_out/cfg5_8gpu_nccl.txt; head -6 gpurun_out/cfg5_8gpu_nccl.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_row_shard.py > gpurun_out/row_shard_2gpu.json 2> gpurun_out/row_shard_2gpu.err; tail -1 gpurun_out/row_shard_2gpu.json | cut -c1-900
