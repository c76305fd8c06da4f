mkdir -p gpurun_out
timeout 330 python tools/psnr_parity.py --steps 2000 --seeds 5 --out gpurun_out/psnr_parity_r2_final.json > gpurun_out/s25_psnr.log 2>&1; echo "psnr rc=$?"; tail -1 gpurun_out/s25_psnr.log | cut -c1-1500
