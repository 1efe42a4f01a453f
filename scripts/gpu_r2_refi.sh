timeout 900 python -m pytest tests/test_ref_unmodified.py -m gpu -q -x > gpurun_out/r2_tests_refi.log 2>&1; tail -25 gpurun_out/r2_tests_refi.log
timeout 900 python bench.py --workload T --steps 2000 --warmup 200 > gpurun_out/bench_T_r2.json 2> gpurun_out/bench_T_r2.err; echo "rc=$?"; tail -c 500 gpurun_out/bench_T_r2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_T_r2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','allgather')})
print('ref ii', d.get('ref_gpu_baseline'))
print('ref i', d.get('ref_gpu_baseline_unmodified_class'))
for st,m in d['single_chain'].items():
    print(st, m['value'], m['ms_per_step'], m['e2e'], m['with_nuisance'])
print(d['roofline'])
PY
