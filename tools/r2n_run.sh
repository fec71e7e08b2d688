#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r2n_bench_n2.log 2>&1
tail -n 5 gpurun_out/r2n_bench_n2.log | cut -c1-2500
