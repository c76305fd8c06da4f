mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "render_vs_reference or trainer_step or full_size or ragged or full_frame" > gpurun_out/s11_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s11_pytest.log
timeout 200 python tools/step_time.py lean 2>/dev/null | tail -1 | tee gpurun_out/s11_ab.log
NNB_LIB_PATH=$PWD/nope_nerf_b200/libnnb_prof.so timeout 300 python tools/tc_prof_mma.py > gpurun_out/s11_prof_fwd.log 2>&1; echo "prof_fwd rc=$?"; tail -42 gpurun_out/s11_prof_fwd.log
