mkdir -p gpurun_out
timeout 300 python tools/wg16_check.py gpurun_out/wg16_check.json > gpurun_out/s3_wg16.log 2>&1; echo "wg16 rc=$?"; tail -8 gpurun_out/s3_wg16.log
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/s3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest.log; tail -25 gpurun_out/s3_pytest.log
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/s3_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','sustained','burst')}); print(d['e2e']); print(d['roofline']['kernel_ms']); print(d.get('full_loss')); print(d.get('c3')); print(d['clocks'])
PY
tail -3 gpurun_out/s3_bench.err
