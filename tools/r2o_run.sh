#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
{
echo "== parity dynamic"; GB200_T1_DEC_DYNAMIC=1 timeout 900 python -m pytest tests/test_gpu_lanes.py tests/test_gpu_pipeline.py tests/test_gpu_stages.py -x -q -m gpu 2>&1 | tail -4
for wl in c4x30 c3; do
  echo "== $wl static"; timeout 600 python tools/t1_bench.py $wl 3
  echo "== $wl dynamic"; GB200_T1_DEC_DYNAMIC=1 timeout 600 python tools/t1_bench.py $wl 3
  for l in 4 8; do echo "== $wl dynamic lanes $l"; GB200_T1_DEC_LANES=$l GB200_T1_DEC_DYNAMIC=1 timeout 600 python tools/t1_bench.py $wl 3; done
done
echo "== c2 mq lanes 2"; GB200_T1_MQ_LANES=2 timeout 300 python tools/t1_bench.py c2 5
echo "== c2 mq lanes 1"; GB200_T1_MQ_LANES=1 timeout 300 python tools/t1_bench.py c2 5
echo "== c2 mq default"; timeout 300 python tools/t1_bench.py c2 5
echo "== c4 mq lanes 2"; GB200_T1_MQ_LANES=2 timeout 300 python tools/t1_bench.py c4 5
echo "== c4 default"; timeout 300 python tools/t1_bench.py c4 5
echo "== strong (2 workers)"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-drop-in 2>&1 | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print(json.dumps({k:{kk:v.get(kk) for kk in ('ms','mpix_s')} for k,v in j['strong'].items()}))"
} > gpurun_out/r2o.log 2>&1
grep -v "^$" gpurun_out/r2o.log | tail -50
