mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize_ref.py -m gpu -q -k "phong or infer_occ or invariant" > gpurun_out/s19_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|passed|failed|^E  " gpurun_out/s19_pytest.log | tail -20
