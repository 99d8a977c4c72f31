CM_CENT_DEBUG=14 ncu --set full --clock-control none --import-source on -k regex:policy_cent_l1_tc -s 8 -c 1 -o gpurun_out/r01g_cent_l1_skel -f python tools/policy_kinds_bench.py c3 --kinds cent > gpurun_out/r01g_cent_skel.log 2>&1
ncu -i gpurun_out/r01g_cent_l1_skel.ncu-rep --page source --csv > gpurun_out/r01g_cent_l1_skel_source.csv 2>/dev/null
ls -la gpurun_out | grep skel
