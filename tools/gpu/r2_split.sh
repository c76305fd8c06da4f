mkdir -p gpurun_out
timeout 300 python tools/fwd_split_check.py gpurun_out/fwd_split_check.json > gpurun_out/s6_split.log 2>&1; echo "split rc=$?"; tail -8 gpurun_out/s6_split.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "render or trainer" > gpurun_out/s6_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s6_pytest.log
timeout 200 python tools/step_time.py > gpurun_out/s6_step.log 2>&1; tail -4 gpurun_out/s6_step.log
