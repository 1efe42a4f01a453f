# cached likelihood records: tests, timing on G (both paths), ncu of the stream kernel
timeout 900 python -m pytest tests -m gpu -q -x -k "cached or replays or facade or nuis or incremental or checkpoint" > gpurun_out/r2_tests_d.log 2>&1; tail -5 gpurun_out/r2_tests_d.log
python scripts/gpu_r2_full_lnz.py --workload G --states init,true > gpurun_out/r2_lnz_cache_G.json 2> gpurun_out/r2_lnz_cache_G.err; cat gpurun_out/r2_lnz_cache_G.json
IG_LNZ_CACHE=0 python scripts/gpu_r2_full_lnz.py --workload G --states true > gpurun_out/r2_lnz_nocache_G.json 2>> gpurun_out/r2_lnz_cache_G.err; cat gpurun_out/r2_lnz_nocache_G.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lnz_stream -s 2 -c 1 -f -o gpurun_out/r2_lnz_stream_G_true python scripts/gpu_ncu_target.py --workload G --state true --steps 2 --nuis 4 > gpurun_out/ncu_lnz_stream.log 2>&1
tail -2 gpurun_out/ncu_lnz_stream.log
