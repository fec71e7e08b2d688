#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/r2q_bench_n$N.log 2>&1
tail -n 5 gpurun_out/r2q_bench_n$N.log | cut -c1-600
nproc; free -g | head -2
