#!/bin/bash
# usage (on a GPU box with N GPUs): tools/bench_n.sh N [extra bench args] -- the bench line of an N-rank run, into gpurun_out/bench_nN.log
N=${1:-2}; shift
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 "$@" ) > gpurun_out/bench_n$N.log 2>&1
grep '^{' gpurun_out/bench_n$N.log | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print(j['n_gpus'], j['value'], j['e2e']['value']); print(json.dumps({k:{kk:v.get(kk) for kk in ('ms','mpix_s','host_gather_ms','bytes_equal_to_one_gpu_run','pixels_equal_to_one_gpu_run')} for k,v in j['strong'].items()}))" || tail -5 gpurun_out/bench_n$N.log
