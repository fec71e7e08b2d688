#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
# memcheck of the Tier-1 paths (small image, every lane variant and the uniform decoder)
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_lanes.py -x -q -m gpu -k "True" 2>&1 | tail -25 ) > gpurun_out/r2m_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2m_memcheck.log
# per-line profile of the current uniform decoder
python tools/t1_bench.py c2 1 > gpurun_out/r2m_plain.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:"t1_decode" -c 1 -f -o gpurun_out/r2m_uni python tools/t1_bench.py c2 1 > gpurun_out/r2m_ncu.log 2>&1
ncu -i gpurun_out/r2m_uni.ncu-rep --page raw --csv > gpurun_out/r2m_uni_raw.csv 2>/dev/null
ncu -i gpurun_out/r2m_uni.ncu-rep --page source --csv > gpurun_out/r2m_uni_src.csv 2>/dev/null
rm -f gpurun_out/r2m_uni.ncu-rep
tail -8 gpurun_out/r2m_memcheck.log; cat gpurun_out/r2m_plain.log
