timeout 1200 python -m pytest tests -m gpu -q -x -k "streaming or facade or cached or medium_level" > gpurun_out/r2_tests_e.log 2>&1; tail -6 gpurun_out/r2_tests_e.log
for sr in 0 default; do
  if [ $sr = 0 ]; then export IG_STREAM=0; else unset IG_STREAM; fi
  python bench.py --workload G --start bomb --steps 1500 --warmup 100 --chains 1 --no-cpu-baseline --no-ref-gpu 2> gpurun_out/bench_G_mid_stream_$sr.err | tail -1 > gpurun_out/bench_G_mid_stream_$sr.json
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_G_mid_stream_$sr.json'))
m=d['single_chain']['mid']
print('IG_STREAM=$sr', 'ms/step', m['ms_per_step'], 'e2e ms', m['e2e']['ms_per_step'], 'nuis e2e ms', m['with_nuisance']['ms_per_step_e2e'], 'lnz ms', m['with_nuisance']['k_full_lnz_ms_per_call'], {k: round(v,1) for k,v in m['kernel_us_per_step'].items()})
PY
done
