"""TEST INFRASTRUCTURE ONLY (oracle/ref_harness).

A minimal stand-in for the ``pycuda`` package so that the *unmodified* reference
``instagraal.cuda_lib_gl_single.sampler`` class (imported from /root/reference/src, never copied)
can run in this GPU-less container.  "Device memory" is host memory, and every kernel launch is
executed by ``oracle/_ref/libref_cpu.so`` -- the reference's own kernel_sparse_adapt.cu compiled
in place for the CPU (see oracle/Makefile, oracle/ref_emu/).  Used only by
``oracle/make_golden.py`` to generate the golden vectors under tests/golden/.
"""
