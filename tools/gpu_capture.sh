#!/bin/bash
# Round profile capture, run on the GPU box via gpurun:   gpurun -- 'bash tools/gpu_capture.sh r01b c2'
# 1. plain bench (must exit 0)  2. ncu launch list of the same command  3. ncu --set full of two launches per kernel.
# Outputs land in gpurun_out/; tools/summarize_ncu.py turns them into the tracked summaries under profiles/.
set -u
TAG=${1:-r02}; CFG=${2:-c3}
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --config $CFG --steps 128 --warmup 64 --no-cpu-baseline --no-sweep --min-seconds 0.01 --e2e-steps 8"
$CMD > $OUT/${TAG}_${CFG}_plain.json 2> $OUT/${TAG}_${CFG}_plain.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_${CFG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $OUT/${TAG}_${CFG}_launches.csv $CMD > $OUT/${TAG}_${CFG}_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'policy_tc_kernel|env_kernel|policy_small|policy_large|policy_attn' -s 200 -c 6 -o $OUT/${TAG}_${CFG}_full -f $CMD > $OUT/${TAG}_${CFG}_ncu_f.log 2>&1
ncu -i $OUT/${TAG}_${CFG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_${CFG}_raw.csv 2>/dev/null
# gpurun brings back at most 64 MiB: the report itself stays on the box unless KEEP_REP=1
[ "${KEEP_REP:-0}" = "1" ] || rm -f $OUT/${TAG}_${CFG}_full.ncu-rep
echo "capture done: $TAG $CFG"; ls -la $OUT | grep ${TAG}_${CFG}
