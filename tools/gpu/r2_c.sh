mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs -x > gpurun_out/s6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s6_pytest.log; tail -15 gpurun_out/s6_pytest.log
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/s6_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','sustained','burst')}); print(d['e2e']); print(d['roofline']['kernel_ms']); print(d.get('full_loss')); print(d.get('c3')); print(d.get('fwd_only')); print(d['clocks'])
PY
tail -3 gpurun_out/s6_bench.err
NNB_LIB_PATH=$PWD/nope_nerf_b200/libnnb_prof.so timeout 300 python tools/tc_prof_mma.py > gpurun_out/s6_prof_fwd.log 2>&1; echo "prof_fwd rc=$?"; tail -40 gpurun_out/s6_prof_fwd.log
NNB_LIB_PATH=$PWD/nope_nerf_b200/libnnb_prof.so timeout 300 python tools/tc_prof_bwd.py > gpurun_out/s6_prof_bwd.log 2>&1; echo "prof_bwd rc=$?"; grep -v "^{" gpurun_out/s6_prof_bwd.log | tail -24
timeout 200 python tools/prof_host.py > gpurun_out/s6_prof_host.log 2>&1; echo "prof_host rc=$?"; head -50 gpurun_out/s6_prof_host.log
timeout 900 python tools/psnr_parity.py --steps 2000 --seeds 6 --arms eager,eager_simt,eager_exact,eager_torchadam --out gpurun_out/psnr_parity_arms.json > gpurun_out/s6_psnr.log 2>&1; echo "psnr rc=$?"; tail -1 gpurun_out/s6_psnr.log
