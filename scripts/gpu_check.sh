# parity suite + short T bench (+ optional G benches with: bash scripts/gpu_check.sh G)
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 3000 --warmup 300 --no-cpu-baseline --no-ref-gpu 2>&1 | tail -1 > gpurun_out/bT.json
python -c "
import json; d=json.load(open('gpurun_out/bT.json')); print('T', d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e_cycle_api']['value'], d['e2e_cycle_api']['device_ms_per_step'])"
if [ "$1" = "G" ]; then bash scripts/gpu_bench_G.sh; fi
