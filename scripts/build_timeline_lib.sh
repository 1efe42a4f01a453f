# builds the instrumented variant of the library used by scripts/gpu_timeline.sh (run here, before gpurun; the .so travels)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC -shared \
     -DIG_TIMELINE instagraal_b200/csrc/ig_kernels.cu -o instagraal_b200/libinstagraal_b200_tl.so
