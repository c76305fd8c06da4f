mkdir -p gpurun_out
./tools/micro/mma_ts > gpurun_out/s5_mma_ts.log 2>&1; echo "mma_ts rc=$?"; cat gpurun_out/s5_mma_ts.log
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/s5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s5_pytest.log; tail -12 gpurun_out/s5_pytest.log
NNB_LIB_PATH=$PWD/nope_nerf_b200/libnnb_prof.so timeout 300 python tools/tc_prof_mma.py > gpurun_out/s5_prof_fwd.log 2>&1; echo "prof_fwd rc=$?"; tail -40 gpurun_out/s5_prof_fwd.log
NNB_LIB_PATH=$PWD/nope_nerf_b200/libnnb_prof.so timeout 300 python tools/tc_prof_bwd.py > gpurun_out/s5_prof_bwd.log 2>&1; echo "prof_bwd rc=$?"; grep -v "^{" gpurun_out/s5_prof_bwd.log | tail -30
timeout 900 python tools/psnr_parity.py --steps 2000 --seeds 6 --out gpurun_out/psnr_parity_r2.json > gpurun_out/s5_psnr.log 2>&1; echo "psnr rc=$?"; tail -1 gpurun_out/s5_psnr.log
