#!/bin/bash
# experiment: lanes per env of env_kernel (CM_ENV_GROUP; 0 = the library's own choice) and library variants      usage: gpurun -- 'bash tools/env_group_sweep.sh'
mkdir -p gpurun_out
run() {  # config group [lib]
  CM_ENV_GROUP=$2 COM_MARL_B200_LIB=$3 python bench.py --config $1 --no-cpu-baseline --no-sweep --e2e-steps 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 group $2 $3', 'value %.1fM' % (d['value']/1e6), 'ms/step %.4f' % d['ms_per_step'], 'policy %.4f' % d['roofline']['ms_per_launch'], 'env %.4f (alone %.4f)' % (d['roofline_env']['ms_per_launch'], d['roofline_env'].get('ms_per_launch_alone', 0)))"
}
{
for spec in "c3 0" "c3 4" "c3 16" "c4 0" "c4 16" "c5 0" "c1 0" "c2 0"; do run $spec; done
W4=$PWD/com_marl_b200/lib/variants/lib_env_w4.so
[ -f $W4 ] && for spec in "c3 0" "c4 0" "c5 0"; do run $spec $W4; done
python -m pytest tests/test_gpu_env.py tests/test_gpu_fullsize.py tests/test_gpu_rollout.py -m gpu -x -q 2>&1 | tail -3
} | tee gpurun_out/env_group_sweep.txt
CMD="python bench.py --config c3 --steps 128 --warmup 64 --no-cpu-baseline --no-sweep --min-seconds 0.01 --e2e-steps 4"
ncu --set full --clock-control none --import-source on -k regex:policy_tc_kernel -s 40 -c 1 -o gpurun_out/pol_c3_src -f $CMD > gpurun_out/pol_c3_src.log 2>&1
ncu -i gpurun_out/pol_c3_src.ncu-rep --page source --csv > gpurun_out/pol_c3_src.csv 2>/dev/null
ncu -i gpurun_out/pol_c3_src.ncu-rep --page raw --csv > gpurun_out/pol_c3_raw.csv 2>/dev/null
rm -f gpurun_out/pol_c3_src.ncu-rep
ls -la gpurun_out | tail -5
