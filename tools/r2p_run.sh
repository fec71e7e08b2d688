#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-drop-in > gpurun_out/r2p_bench.log 2>&1
tail -1 gpurun_out/r2p_bench.log | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print(j['value'], j['e2e']['value']); print(json.dumps({k:{kk:v.get(kk) for kk in ('ms','mpix_s','lossless')} for k,v in j['strong'].items()}))" || tail -5 gpurun_out/r2p_bench.log
