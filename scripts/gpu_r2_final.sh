# final measurements of the round: default bench (G), T, rigid mode at assembled G
python bench.py > gpurun_out/r2_bench_G_c.json 2> gpurun_out/r2_bench_G_c.err; echo "G rc=$?"
python bench.py --workload T --steps 3000 --warmup 300 > gpurun_out/r2_bench_T_c.json 2> gpurun_out/r2_bench_T_c.err; echo "T rc=$?"
python bench.py --workload G --start true --mode rigid --steps 600 --warmup 50 --chains 1 --no-cpu-baseline --no-ref-gpu 2> gpurun_out/bench_G_rigid.err | tail -1 > gpurun_out/r2_bench_G_true_rigid_c.json
python - <<PY
import json
for f in ('r2_bench_G_c','r2_bench_T_c','r2_bench_G_true_rigid_c'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))
    for st,m in d['single_chain'].items():
        print('   ', st, 'dev', round(m['value']), round(m['ms_per_step'],4), 'e2e', round(m['e2e']['value']), round(m['e2e']['ms_per_step'],4), 'nuis e2e ms', round(m['with_nuisance']['ms_per_step_e2e'],4), 'lnz', round(m['with_nuisance']['k_full_lnz_ms_per_call'],4), 'score ms', round(m['kernels']['scoring']['ms_per_step'],4), 'frac', round(m['kernels']['scoring']['frac_hbm'],4), round(m['kernels']['k_full_lnz']['frac_hbm'],4))
PY
