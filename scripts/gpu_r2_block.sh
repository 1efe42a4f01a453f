timeout 900 python -m pytest tests -m gpu -q -x -k "run_cycle or device_rng or replicas or chains or determinism or checkpoint or incremental" > gpurun_out/r2_tests_g.log 2>&1; tail -4 gpurun_out/r2_tests_g.log
for bs in 1 8 32; do
IG_BLOCK_STEPS=$bs timeout 600 python bench.py --workload T --steps 4000 --warmup 200 --no-cpu-baseline --no-ref-gpu > gpurun_out/bench_T_block$bs.json 2> gpurun_out/bench_T_block$bs.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_T_block$bs.json').read().strip().splitlines()[-1])
print('block $bs', '8 chains:', round(d['value']), 'ms/8-chain step', round(d['ms_per_step'],4), 'single:', round(d['single_chain']['mid']['value']), 'ratio', round(d['value']/d['single_chain']['mid']['value'],2))
PY
done
