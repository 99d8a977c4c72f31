#!/bin/bash
# quick device-side numbers for a config: value, ms/step, per-kernel ms      usage: bash tools/quick_bench.sh c2 [extra bench args]
CFG=${1:-c2}; shift
python bench.py --config $CFG --no-cpu-baseline --e2e-steps 20 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$CFG', 'value %.1fM' % (d['value']/1e6), 'ms/step %.4f' % d['ms_per_step'], 'policy %.4f' % d['roofline']['ms_per_launch'], 'env %.4f' % d['roofline_env']['ms_per_launch'], 'e2e %.1fM' % (d['e2e']['value']/1e6), 'frac %.4f' % d['roofline']['frac'], d['clocks'])"
