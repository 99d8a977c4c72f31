python -m pytest tests -m gpu -x -q > gpurun_out/r01g_pytest.log 2>&1; tail -2 gpurun_out/r01g_pytest.log
bash tools/cent_capture.sh 2>&1 | tail -5
bash tools/cent_capture_full.sh > /dev/null 2>&1
for c in c2 c3 c4 c5; do timeout 200 python tools/policy_kinds_bench.py $c 2>/dev/null | tail -3; done > gpurun_out/r01g_policy_kinds.jsonl
wc -l gpurun_out/r01g_policy_kinds.jsonl
python bench.py > gpurun_out/r01g_c2_bench.json 2> gpurun_out/r01g_c2_bench.err; tail -c 300 gpurun_out/r01g_c2_bench.json
