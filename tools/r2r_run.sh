#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
summ() { tail -1 $1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print(j['n_gpus'], j['value'], j['e2e']['value']); print(json.dumps({k:{kk:v.get(kk) for kk in ('ms','mpix_s','bytes_equal_to_one_gpu_run','pixels_equal_to_one_gpu_run','lossless')} for k,v in j['strong'].items()}))" || tail -5 $1; }
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-drop-in > gpurun_out/r2r_n1.log 2>&1; summ gpurun_out/r2r_n1.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-drop-in > gpurun_out/r2r_n2.log 2>&1; summ gpurun_out/r2r_n2.log
