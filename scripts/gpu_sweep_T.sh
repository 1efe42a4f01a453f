timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for d in 1 2 4 8; do IG_GS_DIV=$d python bench.py --steps 2000 --warmup 300 --no-cpu-baseline --no-ref-gpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('gsdiv',$d, d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e_cycle_api']['value'])"; done
