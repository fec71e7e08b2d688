#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_lanes.py -x -q -m gpu 2>&1 | tail -15
echo "== lockstep suite"; GB200_T1_MQ_LOCKSTEP=1 timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_stages.py tests/test_gpu_packed.py -x -q -m gpu -k "not c5" 2>&1 | tail -5
for wl in c2 c1 c4 c4x30 c3; do
  echo "== $wl default"; timeout 600 python tools/t1_bench.py $wl 3
  echo "== $wl lockstep"; GB200_T1_MQ_LOCKSTEP=1 timeout 600 python tools/t1_bench.py $wl 3
done
} > gpurun_out/r2l.log 2>&1
grep -v "^$" gpurun_out/r2l.log | tail -40
