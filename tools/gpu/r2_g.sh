mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s15_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/s15_pytest.log | tail -12
timeout 200 python tools/step_time.py dgrad_overlap 2>/dev/null | tail -1 | tee gpurun_out/s15_ab.log
NNB_LIB_PATH=$PWD/nope_nerf_b200/libnnb_prof.so timeout 300 python tools/tc_prof_bwd.py > gpurun_out/s15_prof_bwd.log 2>&1; echo "prof_bwd rc=$?"; grep -v "^{" gpurun_out/s15_prof_bwd.log | grep -A18 "tc_dgrad MMA"
