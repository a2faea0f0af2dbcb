#!/bin/bash
# 2 GPUs: the bench line under torchrun with k_sweep_q (one reference view per rank + the sharded cfg4 schedule)
set -u
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_40_bench_2gpu.json 2> gpurun_out/r2_40_bench_2gpu.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_40_bench_2gpu.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','n_gpus','iter_ms','ms_per_step','setup_broadcast_ms','parity_bits_equal')})
    print(d.get('per_rank')); print(d['secondary']['cfg4'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r2_40_bench_2gpu.err').read()[-3000:])
PY
