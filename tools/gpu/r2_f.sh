mkdir -p gpurun_out
for v in nope_nerf_b200 nnb_reg144; do
  NNB_LIB_PATH=$PWD/nope_nerf_b200/lib$v.so timeout 200 python tools/step_time.py $v 2>/dev/null | tail -1 | tee -a gpurun_out/s14_ab.log
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "render_vs_reference or trainer_step or full_size or ragged or full_frame" > gpurun_out/s14_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s14_pytest.log
