mkdir -p gpurun_out
NNB_LIB_PATH=$PWD/nope_nerf_b200/libnnb_prof.so timeout 90 python tools/tc_prof_bwd.py > gpurun_out/s17_prof_bwd.log 2>&1; echo "prof_bwd rc=$?"; grep -v "^{" gpurun_out/s17_prof_bwd.log | tail -22
timeout 120 python tools/prof_step.py 200 > gpurun_out/s17_eager200.log 2>&1; echo "eager200 rc=$?"; tail -1 gpurun_out/s17_eager200.log
timeout 200 python tools/step_time.py dgrad_aux 2>/dev/null | tail -1 | tee gpurun_out/s17_ab.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/s17_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/s17_pytest.log | tail -8
