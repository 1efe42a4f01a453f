// TEST INFRASTRUCTURE ONLY: stub of <curand_kernel.h> for the CPU emulation of the reference
// kernels.  The reference only uses curand for OpenGL particle jitter (gl_update_pos), which
// never feeds the chain, so a trivial LCG is enough.
#pragma once
#include "cuda_shim.h"
struct curandState { unsigned long long s; unsigned long long pad[5]; };  // 48 bytes like XORWOW
typedef curandState curandStateXORWOW;
static inline void curand_init(unsigned long long seed, unsigned long long seq, unsigned long long off, curandState* st) {
    st->s = seed * 6364136223846793005ULL + seq * 1442695040888963407ULL + off + 1;
}
static inline float curand_normal(curandState* st) {
    st->s = st->s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (float)((double)(st->s >> 11) / 9007199254740992.0 - 0.5);
}
