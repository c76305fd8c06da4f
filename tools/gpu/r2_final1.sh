mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_r2b_n1.json 2> gpurun_out/bench_r2b_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2b_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','sustained','burst','gpu_launches')}); print(d['e2e']); print(d['roofline']); print(d.get('full_loss')); print(d.get('c3')); print(d.get('fwd_only')); print(d.get('exact_wgrad_planes')); print(d.get('reference_cuda')); print(d.get('cpu_baseline')); print(d['clocks'])
PY
tail -3 gpurun_out/bench_r2b_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2b.csv python tools/prof_step.py 4 > gpurun_out/s16_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_field_fwd|tc_dgrad|tc_wgrad16' -s 3 -c 3 -o gpurun_out/ncu_r2c_tc -f python tools/prof_step.py 3 > gpurun_out/s16_ncu2.log 2>&1; echo "ncu2 rc=$?"
