#!/bin/bash
# round 2, call 10 (2 GPUs): the N>1 bench path (weak-scaling cfg3 headline + the sharded cfg4 schedule) and the sharded schedule's bit-identity check
set -u
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/tools/pipeline_sharded.py --width 1000 --height 562 --views 6 --src 3 --check --out gpurun_out/r02_pipeline_sharded_check_2gpu.json 2>&1 | grep -v Warning | tail -4
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_10_bench_2gpu.json 2> gpurun_out/r2_10_bench_2gpu.err
tail -c 3500 gpurun_out/r2_10_bench_2gpu.json; echo; grep -v "Warning\|warn\|^$\|\*\*\*" gpurun_out/r2_10_bench_2gpu.err | tail -8
