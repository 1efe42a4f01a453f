# device-side timeline of one step's kernels (warm, graph replay); the _tl library is built by: nvcc ... -DIG_TIMELINE
# needs instagraal_b200/libinstagraal_b200_tl.so (bash scripts/build_timeline_lib.sh)
for wl in ${WLS:-T}; do IG_B200_LIB=$PWD/instagraal_b200/libinstagraal_b200_tl.so python scripts/timeline.py $wl 2000 > gpurun_out/timeline_$wl.txt 2>&1; cat gpurun_out/timeline_$wl.txt | grep -v Warning | tail -32; done
