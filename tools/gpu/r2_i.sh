mkdir -p gpurun_out
timeout 300 python tools/wg16_check.py gpurun_out/wg16_check_b.json > gpurun_out/s18_wg16.log 2>&1; echo "wg16 rc=$?"; tail -4 gpurun_out/s18_wg16.log | cut -c1-400
timeout 200 python tools/step_time.py bias_in_wgrad 2>/dev/null | tail -1 | tee gpurun_out/s18_ab.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/s18_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/s18_pytest.log | tail -8
