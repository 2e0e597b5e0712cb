#!/bin/bash
# two B200s: the row-sharded parity tests (NCCL all-reduce inside the library; skipped on a 1-GPU box)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "row_shard or sharded" > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_2gpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_row_shard.py > gpurun_out/row_shard_2gpu.json 2> gpurun_out/row_shard_2gpu.err; echo "row_shard rc=$?"; tail -1 gpurun_out/row_shard_2gpu.json | cut -c1-700
