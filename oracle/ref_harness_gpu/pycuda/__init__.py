"""TEST / BASELINE INFRASTRUCTURE ONLY (oracle/ref_harness_gpu).

A minimal stand-in for the ``pycuda`` package on top of cuda-python's driver bindings, so that the *unmodified* reference
``instagraal.cuda_lib_gl_single.sampler`` class (a copy of /root/reference/src/instagraal under the git-ignored
baseline/_ref/, made by build()) runs on a real GPU: pycuda itself is not installable here (no network, no wheel).
Device memory is device memory, every kernel launch is a cuLaunchKernel of the reference's own kernels
(oracle/_ref/ref_kernels.cubin = kernel_sparse_adapt.cu compiled for sm_100a by oracle/Makefile with the five macro values
of CL:1526-1537).  What differs from the real pycuda: ``SourceModule`` loads that prebuilt cubin instead of calling nvcc at
run time (not timed by the bench), and ``gpuarray.max / sum`` reduce on the host after one device-to-host copy (pycuda
launches a reduction kernel and copies 4 bytes back).  bench.py times this as B-ref route (i).
"""
