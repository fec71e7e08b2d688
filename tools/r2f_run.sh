#!/bin/bash
# one gpurun call: parity of the new Tier-1 variants, then their timings
cd /root/repo
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_lanes.py tests/test_gpu_stages.py -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/r2f_pytest.log 2>&1
( GB200_T1_DEC_UNIFORM=1 timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_stages.py -x -q -m gpu -k "not c5" 2>&1 | tail -15 ) > gpurun_out/r2f_pytest_uniform.log 2>&1
{
for wl in c2 c1 c4; do
  echo "== $wl default"; timeout 300 python tools/t1_bench.py $wl 5
  echo "== $wl uniform"; GB200_T1_DEC_UNIFORM=1 timeout 300 python tools/t1_bench.py $wl 5
  for ch in 1 2 4 8; do for sd in 1 2 4; do
    [ $ch == 1 ] && [ $sd != 1 ] && continue
    echo "== $wl chunks $ch sides $sd"; GB200_T1_ENC_CHUNKS=$ch GB200_T1_ENC_SIDES=$sd timeout 300 python tools/t1_bench.py $wl 5
  done; done
done
for wl in c4x30 c3; do
  echo "== $wl default"; timeout 600 python tools/t1_bench.py $wl 3
  echo "== $wl uniform"; GB200_T1_DEC_UNIFORM=1 timeout 600 python tools/t1_bench.py $wl 3
  echo "== $wl chunks 1"; GB200_T1_ENC_CHUNKS=1 timeout 600 python tools/t1_bench.py $wl 3
  echo "== $wl chunks 4 sides 2"; GB200_T1_ENC_CHUNKS=4 GB200_T1_ENC_SIDES=2 timeout 600 python tools/t1_bench.py $wl 3
done
} > gpurun_out/r2f_t1.log 2>&1
tail -5 gpurun_out/r2f_pytest.log gpurun_out/r2f_pytest_uniform.log; grep -v "^$" gpurun_out/r2f_t1.log | tail -80
