timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests_n.log 2>&1; tail -3 gpurun_out/r2_tests_n.log
python bench.py --workload G --start true --steps 600 --warmup 50 --chains 1 --no-cpu-baseline --no-ref-gpu 2> gpurun_out/bench_G_dedup.err | tail -1 > gpurun_out/bench_G_true_dedup.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_G_true_dedup.json'))
m=d['single_chain']['assembled']
print('assembled: ms/step', round(m['ms_per_step'],4), 'value', round(m['value']), 'scoring ms', round(m['kernels']['scoring']['ms_per_step'],4), 'nuis pair', round(m['with_nuisance']['ms_per_step_e2e'],4))
PY
