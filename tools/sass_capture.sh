#!/bin/bash
# SASS-level ncu capture (instruction counts + stall samples per SASS line) of one launch of a kernel at a config
# usage: gpurun -- 'bash tools/sass_capture.sh TAG CONFIG KERNEL_REGEX'      -> gpurun_out/TAG_src.csv, TAG_raw.csv
TAG=$1; CFG=$2; K=$3
CMD="python bench.py --config $CFG --steps 128 --warmup 64 --no-cpu-baseline --no-sweep --min-seconds 0.01 --e2e-steps 4"
ncu --set full --clock-control none --import-source on -k regex:$K -s 40 -c 1 -o gpurun_out/$TAG -f $CMD > gpurun_out/${TAG}.log 2>&1
ncu -i gpurun_out/$TAG.ncu-rep --page source --csv > gpurun_out/${TAG}_src.csv 2>/dev/null
ncu -i gpurun_out/$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
rm -f gpurun_out/$TAG.ncu-rep
ls -la gpurun_out/${TAG}_src.csv
