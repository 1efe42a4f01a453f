# last GPU batch of round 2: the new -m gpu test, the default bench line (workload G), the ncu launch list of assembled-G steps, Y3
timeout 120 python -m pytest tests/test_pyramid_load.py -m gpu -q 2>&1 | tail -2
timeout 420 python bench.py > gpurun_out/r2_bench_G_1gpu_e.json 2> gpurun_out/r2_bench_G_e.err; echo "G rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 700 --csv --log-file gpurun_out/r2_launches_G_assembled.csv \
  python bench.py --workload G --start true --steps 20 --warmup 3 --chains 1 --nuisance-steps 10 --no-cpu-baseline --no-ref-gpu > gpurun_out/l.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/r2_launches_G_assembled.csv > gpurun_out/r2_launches_G_assembled_summary.txt 2>&1; tail -25 gpurun_out/r2_launches_G_assembled_summary.txt
timeout 200 python bench.py --workload Y3 --steps 2000 --warmup 200 > gpurun_out/r2_bench_Y3_1gpu.json 2> gpurun_out/r2_bench_Y3.err; echo "Y3 rc=$?"
python - <<PY
import json
for f in ('r2_bench_G_1gpu_e','r2_bench_Y3_1gpu'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'roofline', d['roofline']['kernel'], round(d['roofline']['frac'],5), 'share', round(d['roofline'].get('share_of_step',0),3))
    for st,m in d['single_chain'].items():
        print('   ', st, 'dev', round(m['value']), round(m['ms_per_step'],4), 'e2e', round(m['e2e']['value']), 'nuis e2e ms', round(m['with_nuisance']['ms_per_step_e2e'],4), 'lnz', round(m['with_nuisance']['k_full_lnz_ms_per_call'],4), 'score ms', round(m['kernels']['scoring']['ms_per_step'],4), 'frac', round(m['kernels']['scoring']['frac_hbm'],4), round(m['kernels']['k_full_lnz']['frac_hbm'],4))
PY
