# same-box A/B of the scoring kernel's knobs on the T workload: bash scripts/gpu_sweep_T.sh "IG_SPARSE_DIV=0" "IG_SPARSE_DIV=4" ...
for cfg in "$@"; do
  for rep in 1 2; do
  env $cfg python bench.py --steps 3000 --warmup 300 --no-cpu-baseline --no-ref-gpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$cfg', round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']), round(d['e2e_cycle_api']['value']), d['e2e_cycle_api']['device_ms_per_step'])"
  done
done
