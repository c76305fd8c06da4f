mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest.log; tail -15 gpurun_out/s2_pytest.log
timeout 600 python tools/psnr_parity.py --steps 2000 --seeds 3 --out gpurun_out/psnr_parity_r2.json > gpurun_out/s2_psnr.log 2>&1; echo "psnr rc=$?"; tail -3 gpurun_out/s2_psnr.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2a.csv python tools/prof_step.py 4 > gpurun_out/s2_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2a_full.csv python tools/prof_step.py 4 full > gpurun_out/s2_ncu1f.log 2>&1; echo "ncu1f rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_field_fwd|tc_dgrad|tc_wgrad' -s 3 -c 3 -o gpurun_out/ncu_r2a_tc -f python tools/prof_step.py 3 > gpurun_out/s2_ncu2.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out
