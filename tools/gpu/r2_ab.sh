mkdir -p gpurun_out
./tools/micro/mbar_pingpong > gpurun_out/s8_pingpong.log 2>&1; cat gpurun_out/s8_pingpong.log
for v in nope_nerf_b200 nnb_nohint nnb_old256 nnb_old256_nohint; do
  NNB_LIB_PATH=$PWD/nope_nerf_b200/lib$v.so timeout 200 python tools/step_time.py $v 2>/dev/null | tail -1 | tee -a gpurun_out/s8_ab.log
done
