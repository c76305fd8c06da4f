# consolidated N=1 records of the round (product library: split experiment compiled out)
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_r2_final_n1.json 2> gpurun_out/bench_r2_final_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_final_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','sustained','burst','gpu_launches')}); print(d['e2e']['value'], d['e2e']['ms_per_step']); print(d['roofline']['kernel_ms'], d['roofline']['frac']); print(d.get('full_loss',{}).get('ms_per_step')); print(d.get('c3',{}).get('ms_per_step')); print(d.get('fwd_only',{}).get('value')); print(d.get('exact_wgrad_planes',{}).get('ms_per_step')); print(d.get('reference_cuda',{}).get('sec_per_step')); print(d['clocks'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_final.csv python tools/prof_step.py 4 > gpurun_out/s21_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_field_fwd|tc_dgrad|tc_wgrad16' -s 3 -c 3 -o gpurun_out/ncu_r2_final_tc -f python tools/prof_step.py 3 > gpurun_out/s21_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/s24_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s24_pytest.log
