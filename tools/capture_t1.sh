#!/bin/bash
# Tier-1 part of tools/capture_profiles.sh only (launch list of a bench step + full capture of the Tier-1 / MCT kernels), for a
# change that leaves the wavelet kernels alone.   usage: tools/capture_t1.sh <tag>
set -u
tag=${1:-r02c}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong --no-drop-in"
$B > gpurun_out/${tag}_bench_plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/${tag}_bench_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_launches.log 2>&1
python tools/t1_bench.py c2 1 > gpurun_out/${tag}_t1_plain.log 2>&1 || { echo "t1_bench failed"; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name regex:"t1_|mct3|dcshift" -c 12 -f -o gpurun_out/${tag}_t1_full python tools/t1_bench.py c2 1 > gpurun_out/${tag}_ncu_t1.log 2>&1
ncu -i gpurun_out/${tag}_t1_full.ncu-rep --page raw --csv > gpurun_out/${tag}_t1_raw.csv 2>/dev/null
rm -f gpurun_out/${tag}_t1_full.ncu-rep
cat gpurun_out/${tag}_t1_plain.log
