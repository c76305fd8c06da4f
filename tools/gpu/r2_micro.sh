mkdir -p gpurun_out
./tools/micro/mma_commit > gpurun_out/s10_mma_commit.log 2>&1; cat gpurun_out/s10_mma_commit.log
