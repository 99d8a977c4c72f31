timeout 300 python -m pytest tests/test_gpu_policy_cent.py -x -q 2>&1 | tail -3
for c in c3 c4 c5; do timeout 120 python tools/policy_kinds_bench.py $c --kinds cent 2>&1 | tail -1; done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:policy_cent -s 40 -c 6 --csv --log-file gpurun_out/r01g_cent_c3_launches.csv python tools/policy_kinds_bench.py c3 --kinds cent > gpurun_out/r01g_cent_ncu.log 2>&1
grep -c policy_cent gpurun_out/r01g_cent_c3_launches.csv
