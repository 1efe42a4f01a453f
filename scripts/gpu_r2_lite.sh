timeout 1200 python -m pytest tests -m gpu -q -x -k "cached or facade or incremental or run_cycle or checkpoint or determinism" > gpurun_out/r2_tests_f.log 2>&1; tail -4 gpurun_out/r2_tests_f.log
python bench.py --workload G --start both --steps 1500 --warmup 100 --chains 1 --no-cpu-baseline --no-ref-gpu 2> gpurun_out/bench_G_lite.err | tail -1 > gpurun_out/bench_G_lite.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_G_lite.json'))
for st in ('mid','assembled'):
    m=d['single_chain'][st]
    print(st, 'ms/step', round(m['ms_per_step'],4), 'e2e ms', round(m['e2e']['ms_per_step'],4), 'nuis:', {k:(round(v,4) if isinstance(v,float) else v) for k,v in m['with_nuisance'].items() if k!='note'})
PY
