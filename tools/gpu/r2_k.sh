mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s20_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|passed|failed|^E  " gpurun_out/s20_pytest.log | tail -12
timeout 200 python tools/step_time.py one_adam 2>/dev/null | tail -1 | tee gpurun_out/s20_ab.log
timeout 100 python tools/prof_host.py 2>/dev/null | head -3
