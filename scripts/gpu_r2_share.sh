for sh in 1 2 4 8; do
IG_GPU_SHARE=$sh timeout 600 python bench.py --workload T --steps 4000 --warmup 200 --no-cpu-baseline --no-ref-gpu > gpurun_out/bench_T_share$sh.json 2> gpurun_out/bench_T_share$sh.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_T_share$sh.json').read().strip().splitlines()[-1])
print('share $sh', '8 chains:', round(d['value']), 'ms/8-chain step', round(d['ms_per_step'],4), 'single:', round(d['single_chain']['mid']['value']), 'ratio', round(d['value']/d['single_chain']['mid']['value'],2))
PY
done
for sh in 1 4; do
IG_GPU_SHARE=$sh timeout 900 python bench.py --workload G --start bomb --steps 1000 --warmup 100 --no-cpu-baseline --no-ref-gpu > gpurun_out/bench_G_share$sh.json 2> gpurun_out/bench_G_share$sh.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_G_share$sh.json').read().strip().splitlines()[-1])
print('G share $sh', '8 chains:', round(d['value']), 'ms/8-chain step', round(d['ms_per_step'],4), 'single:', round(d['single_chain']['mid']['value']), 'ratio', round(d['value']/d['single_chain']['mid']['value'],2))
PY
done
