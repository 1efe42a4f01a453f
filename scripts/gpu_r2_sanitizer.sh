# compute-sanitizer over the round-2 kernels on small levels (memcheck, then racecheck)
K='cached_likelihood_records_with_circular or (cached_likelihood_records_equal and toy) or (streaming_path_equals and toy) or streaming_path_random or large_level_code_paths or chains_sharing or bin_contacts_edge or build_matches_reference or run_cycle_equals or device_rng'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests -m gpu -q -x -k "$K" > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2_sanitizer_$tool.log | tail -4
done
