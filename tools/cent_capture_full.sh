# ncu --set full of one policy_cent_l1_tc_kernel launch (C3 sizes) with source-level stall samples
ncu --set full --clock-control none --import-source on -k regex:policy_cent_l1_tc -s 8 -c 1 -o gpurun_out/r01g_cent_l1_full -f python tools/policy_kinds_bench.py c3 --kinds cent > gpurun_out/r01g_cent_full.log 2>&1
ncu -i gpurun_out/r01g_cent_l1_full.ncu-rep --page raw --csv > gpurun_out/r01g_cent_l1_raw.csv 2>/dev/null
ncu -i gpurun_out/r01g_cent_l1_full.ncu-rep --page source --csv > gpurun_out/r01g_cent_l1_source.csv 2>/dev/null
ls -la gpurun_out | grep r01g
