for c in 8 32; do
export CUDA_DEVICE_MAX_CONNECTIONS=$c
timeout 600 python bench.py --workload T --steps 4000 --warmup 200 --no-cpu-baseline --no-ref-gpu > gpurun_out/bench_T_conn$c.json 2> gpurun_out/bench_T_conn$c.err
timeout 900 python bench.py --workload G --start bomb --steps 1500 --warmup 100 --no-cpu-baseline --no-ref-gpu > gpurun_out/bench_G_conn$c.json 2> gpurun_out/bench_G_conn$c.err
python - <<PY
import json
for w in ("T","G"):
    d=json.loads(open('gpurun_out/bench_%s_conn$c.json'%w).read().strip().splitlines()[-1])
    print('conn $c', w, '8 chains:', round(d['value']), 'ms/8-chain step', round(d['ms_per_step'],4), 'single:', round(d['single_chain']['mid']['value']))
PY
done
