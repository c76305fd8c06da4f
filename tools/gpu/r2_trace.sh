mkdir -p gpurun_out
NNB_LIB_PATH=$PWD/nope_nerf_b200/libnnb_prof.so timeout 300 python tools/tc_prof_mma.py > gpurun_out/s9_prof_fwd.log 2>&1; echo "prof_fwd rc=$?"; cat gpurun_out/s9_prof_fwd.log | tail -80
