#!/bin/bash
# Round validation on the GPU box:   gpurun --timeout 900 -- 'bash tools/round_capture.sh r02o'
# GPU tests, smoke(), the default bench line (the driver's command), the reference arm, then the ncu launch list + full summary of C3.
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; tail -2 $OUT/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 400 $OUT/${TAG}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2> /dev/null; tail -c 300 $OUT/${TAG}_bench_reference_arm.json
bash tools/gpu_capture.sh $TAG c3 | tail -3
