#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/gpu_cfg5_shard.py 2> gpurun_out/cfg5.err | tee gpurun_out/probe_cfg5_2gpu.json
tail -5 gpurun_out/cfg5.err
