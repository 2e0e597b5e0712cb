#!/bin/bash
# two B200s: row-sharded parity test (NCCL all-reduce inside the library) + chain-sharded bench at N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "row_sharded" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 700 gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
