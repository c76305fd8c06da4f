# validation of the load-before-store epilogue order + the consolidated N=1 records of the round
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s21_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; grep -E "^FAILED|passed|failed|^E  " gpurun_out/s21_pytest.log | tail -12
timeout 200 python tools/step_time.py 2>/dev/null | tail -1 | tee gpurun_out/s21_step.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python bench.py > gpurun_out/bench_r2_final_n1.json 2> gpurun_out/bench_r2_final_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_final_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','sustained','burst','gpu_launches')}); print(d['e2e']); print(d['roofline']); print(d.get('full_loss')); print(d.get('c3')); print(d.get('fwd_only')); print(d.get('exact_wgrad_planes')); print(d.get('reference_cuda')); print(d.get('cpu_baseline')); print(d['clocks'])
PY
tail -3 gpurun_out/bench_r2_final_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_final.csv python tools/prof_step.py 4 > gpurun_out/s21_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_field_fwd|tc_dgrad|tc_wgrad16' -s 3 -c 3 -o gpurun_out/ncu_r2_final_tc -f python tools/prof_step.py 3 > gpurun_out/s21_ncu2.log 2>&1; echo "ncu2 rc=$?"
