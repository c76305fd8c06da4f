# bench only, on every GPU of the box (8-GPU validation of the peer-memory exchange: weak, strong and C5 records)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/s5_bench_n$N.json 2> gpurun_out/s5_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/s5_bench_n$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','sustained','n_gpus')}); print(d.get('strong')); print(d.get('c5')); print(d['e2e']['value'], d['e2e']['ms_per_step'])
PY
tail -5 gpurun_out/s5_bench_n$N.err
