#!/bin/bash
# usage (on the GPU box): VARIANTS="a b" WLS="c2 c1" bash tools/variant_run.sh -- parity (lane / stage / pipeline tests) and Tier-1 timing of
# grokimagecompression_b200/libgrok_b200_<variant>.so (tools/build_variant.sh), then of the default library
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for v in $VARIANTS; do
  export GB200_LIB=$PWD/grokimagecompression_b200/libgrok_b200_$v.so
  echo "== parity $v"; timeout 600 python -m pytest tests/test_gpu_lanes.py tests/test_gpu_stages.py tests/test_gpu_pipeline.py -x -q -m gpu -k "not c5 and not full" 2>&1 | tail -2
  for wl in $WLS; do echo "== $wl $v"; timeout 300 python tools/t1_bench.py $wl 5; done
done
unset GB200_LIB
for wl in $WLS; do echo "== $wl default"; timeout 300 python tools/t1_bench.py $wl 5; done
} > gpurun_out/variant_run.log 2>&1
grep -v "^$" gpurun_out/variant_run.log | tail -60
