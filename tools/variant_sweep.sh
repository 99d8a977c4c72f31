#!/bin/bash
# experiment: library variants built under com_marl_b200/lib/variants/ (COM_MARL_B200_LIB selects one)      usage: gpurun -- 'bash tools/variant_sweep.sh "c3 c4" lib_a.so lib_b.so'
mkdir -p gpurun_out
CFGS=$1; shift
run() {  # config lib
  COM_MARL_B200_LIB=$2 python bench.py --config $1 --no-cpu-baseline --no-sweep --e2e-steps 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 $(basename $2)', 'value %.1fM' % (d['value']/1e6), 'ms/step %.4f' % d['ms_per_step'], 'policy %.4f' % d['roofline']['ms_per_launch'], 'env %.4f' % d['roofline_env']['ms_per_launch'])"
}
{
for c in $CFGS; do for v in "$@"; do run $c $PWD/com_marl_b200/lib/variants/$v; done; done
} | tee gpurun_out/variant_sweep.txt
