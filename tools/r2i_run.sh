#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_lanes.py -x -q -m gpu 2>&1 | tail -5 ) > gpurun_out/r2i_pytest.log 2>&1
( GB200_T1_DEC_UNIFORM=1 timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_stages.py tests/test_gpu_packed.py -x -q -m gpu -k "not c5" 2>&1 | tail -8 ) > gpurun_out/r2i_pytest_uniform.log 2>&1
{
for wl in c2 c1 c4; do
  echo "== $wl default"; timeout 300 python tools/t1_bench.py $wl 5
  echo "== $wl uniform"; GB200_T1_DEC_UNIFORM=1 timeout 300 python tools/t1_bench.py $wl 5
  for v in $VARIANTS; do
  echo "== $wl uniform $v"; GB200_LIB=$PWD/grokimagecompression_b200/libgrok_b200_$v.so GB200_T1_DEC_UNIFORM=1 timeout 300 python tools/t1_bench.py $wl 5
  done
done
} > gpurun_out/r2i_t1.log 2>&1
tail -n 3 gpurun_out/r2i_pytest.log gpurun_out/r2i_pytest_uniform.log; grep -v "^$" gpurun_out/r2i_t1.log | tail -40
