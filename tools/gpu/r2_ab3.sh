mkdir -p gpurun_out; rm -f gpurun_out/s23_ab.log
for v in nope_nerf_b200 nnb_nosplit nope_nerf_b200 nnb_nosplit nope_nerf_b200 nnb_nosplit; do
  NNB_LIB_PATH=$PWD/nope_nerf_b200/lib$v.so timeout 200 python tools/step_time.py $v 2>/dev/null | tail -1 | tee -a gpurun_out/s23_ab.log
done
