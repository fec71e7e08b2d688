#!/bin/bash
# variants of the decode kernel: parity (lane test + stage tests) and Tier-1 timing per variant
cd /root/repo
mkdir -p gpurun_out
{
for v in $VARIANTS; do
  export GB200_LIB=$PWD/grokimagecompression_b200/libgrok_b200_$v.so
  echo "== parity $v"; timeout 600 python -m pytest tests/test_gpu_lanes.py tests/test_gpu_stages.py tests/test_gpu_pipeline.py -x -q -m gpu -k "not c5 and not full" 2>&1 | tail -2
  for wl in $WLS; do echo "== $wl $v"; timeout 300 python tools/t1_bench.py $wl 5; done
done
unset GB200_LIB
for wl in $WLS; do echo "== $wl default"; timeout 300 python tools/t1_bench.py $wl 5; done
} > gpurun_out/r2k_t1.log 2>&1
grep -v "^$" gpurun_out/r2k_t1.log | tail -60
