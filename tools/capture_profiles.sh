#!/bin/bash
# Run on the GPU box (gpurun): the ncu evidence kept under profiles/ for this round.  Every capture follows a plain run of the
# same command that exited 0.  Outputs go to gpurun_out/; profiles/update_constants.py turns them into the summaries and the
# source-hashed constants bench.py reads (t1_issue.json, dwt_traffic.json).
#   usage: tools/capture_profiles.sh <tag>      e.g. r02
set -u
tag=${1:-r02}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong --no-drop-in"
$B > gpurun_out/${tag}_bench_plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/${tag}_bench_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_launches.log 2>&1
python tools/t1_bench.py c2 1 > gpurun_out/${tag}_t1_plain.log 2>&1 || { echo "t1_bench failed"; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name regex:"t1_|mct3|dcshift" -c 12 -f -o gpurun_out/${tag}_t1_full python tools/t1_bench.py c2 1 > gpurun_out/${tag}_ncu_t1.log 2>&1
ncu -i gpurun_out/${tag}_t1_full.ncu-rep --page raw --csv > gpurun_out/${tag}_t1_raw.csv 2>/dev/null
# the wavelet kernels of one configs[1] image (9/7) and one configs[2] image (5/3): DRAM traffic per launch
for w in c2 c3; do
  DWT_BENCH_WARMUP=0 python tools/dwt_bench.py $w 1 "rows=0,unroll=1" > gpurun_out/${tag}_dwt_${w}_plain.log 2>&1
  # one image forward (5 level launches), then one image inverse (5)
  DWT_BENCH_WARMUP=0 ncu --set full --clock-control none --kernel-name regex:dwt_ -c 10 -f -o gpurun_out/${tag}_dwt_${w} python tools/dwt_bench.py $w 1 "rows=0,unroll=1" > gpurun_out/${tag}_ncu_dwt_${w}.log 2>&1
  ncu -i gpurun_out/${tag}_dwt_${w}.ncu-rep --page raw --csv > gpurun_out/${tag}_dwt_${w}_raw.csv 2>/dev/null
done
rm -f gpurun_out/${tag}_dwt_c2.ncu-rep gpurun_out/${tag}_dwt_c3.ncu-rep   # the raw csv is what is kept
ls -la gpurun_out | tail -20
