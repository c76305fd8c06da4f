mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/s12_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s12_pytest.log
timeout 200 python tools/step_time.py nhalf_both 2>/dev/null | tail -1 | tee gpurun_out/s12_ab.log
NNB_LIB_PATH=$PWD/nope_nerf_b200/libnnb_prof.so timeout 300 python tools/tc_prof_bwd.py > gpurun_out/s12_prof_bwd.log 2>&1; echo "prof_bwd rc=$?"; grep -v "^{" gpurun_out/s12_prof_bwd.log | tail -22
