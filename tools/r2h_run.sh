#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
export GB200_T1_DEC_UNIFORM=1
python tools/t1_bench.py c2 1 > gpurun_out/r2h_plain.log 2>&1 || { echo fail; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name regex:"t1_decode" -c 2 -f -o gpurun_out/r2h_uni python tools/t1_bench.py c2 1 > gpurun_out/r2h_ncu.log 2>&1
ncu -i gpurun_out/r2h_uni.ncu-rep --page raw --csv > gpurun_out/r2h_uni_raw.csv 2>/dev/null
ncu -i gpurun_out/r2h_uni.ncu-rep --page source --csv > gpurun_out/r2h_uni_src.csv 2>/dev/null
cat gpurun_out/r2h_plain.log; tail -3 gpurun_out/r2h_ncu.log
