mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_field_fwd|tc_dgrad' -s 2 -c 2 -o gpurun_out/ncu_r2b_tc -f python tools/prof_step.py 2 > gpurun_out/s7_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/s7_ncu.log
