# replica bench at N GPUs as the driver launches it ($1 = N)
N=${1:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 2000 --warmup 200 > gpurun_out/r2_bench_G_${N}gpu.json 2> gpurun_out/r2_bench_G_${N}gpu.err; echo "rc=$?"
python - <<PY
import json
lines=[l for l in open('gpurun_out/r2_bench_G_${N}gpu.json').read().strip().splitlines()]
print('stdout lines:', len(lines))
d=json.loads(lines[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, d['config']['chains_per_gpu'], d['allgather'], 'e2e', d['e2e']['value'])
PY
