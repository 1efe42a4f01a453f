// TEST INFRASTRUCTURE ONLY (oracle/): fiber scheduler that executes a CUDA-style kernel
// (compiled by g++ through cuda_shim.h) block by block on one CPU thread.
// Every CUDA thread of a block is a fiber with its own stack; __syncthreads() yields to the
// scheduler, which resumes the block's fibers round-robin => correct barrier semantics and a
// deterministic (thread-index) order for shared-memory atomics.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include "cuda_shim.h"

emu_uint3 threadIdx, blockIdx, blockDim, gridDim;

typedef void (*kfn_t)(uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t,
                      double, double, double, double,
                      uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t,
                      uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t);

static const size_t STACK_BYTES = 96 * 1024;
static const int MAX_THREADS = 1024;
static char* g_stacks = nullptr;

struct Fiber { void* sp; int alive; };
static Fiber g_fibers[MAX_THREADS];
static void* g_sched_sp;
static int g_cur;
static kfn_t g_fn;
static uint64_t g_ia[22];
static double g_fa[4];

extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

static void fiber_entry() {
    g_fn(g_ia[0], g_ia[1], g_ia[2], g_ia[3], g_ia[4], g_ia[5], g_fa[0], g_fa[1], g_fa[2], g_fa[3],
         g_ia[6], g_ia[7], g_ia[8], g_ia[9], g_ia[10], g_ia[11], g_ia[12], g_ia[13],
         g_ia[14], g_ia[15], g_ia[16], g_ia[17], g_ia[18], g_ia[19], g_ia[20], g_ia[21]);
    Fiber* f = &g_fibers[g_cur];
    f->alive = 0;
    emu_switch(&f->sp, g_sched_sp);
    abort();
}

void emu_syncthreads(void) {
    Fiber* f = &g_fibers[g_cur];
    emu_switch(&f->sp, g_sched_sp);
}

static void init_fiber(int t) {
    char* top = g_stacks + (size_t)(t + 1) * STACK_BYTES;
    uint64_t* L = (uint64_t*)(((uintptr_t)top - 64) & ~(uintptr_t)15);
    L[1] = 0;                          // fake return address of fiber_entry
    L[0] = (uint64_t)(uintptr_t)&fiber_entry;
    uint64_t* sp = L - 6;              // r15 r14 r13 r12 rbx rbp
    for (int i = 0; i < 6; i++) sp[i] = 0;
    g_fibers[t].sp = sp;
    g_fibers[t].alive = 1;
}

// int_args: integer-class arguments in declaration order (pointers, ints, 64-bit ints);
// fp_bits: raw 32-bit patterns of float arguments (placed in the low lanes of xmm0..3).
extern "C" int emu_launch(void* fn, unsigned gx, unsigned gy, unsigned bx, unsigned by, unsigned bz,
                          int n_int, const uint64_t* int_args, int n_fp, const uint32_t* fp_bits) {
    if (n_int > 22 || n_fp > 4) return -1;
    unsigned nthreads = bx * by * bz;
    if (nthreads == 0 || nthreads > (unsigned)MAX_THREADS) return -2;
    if (!g_stacks) {
        g_stacks = (char*)mmap(nullptr, STACK_BYTES * MAX_THREADS, PROT_READ | PROT_WRITE,
                               MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (g_stacks == (char*)MAP_FAILED) return -3;
    }
    g_fn = (kfn_t)fn;
    memset(g_ia, 0, sizeof(g_ia));
    for (int i = 0; i < n_int; i++) g_ia[i] = int_args[i];
    for (int i = 0; i < 4; i++) g_fa[i] = 0.0;
    for (int i = 0; i < n_fp; i++) { uint64_t b = fp_bits[i]; memcpy(&g_fa[i], &b, 8); }
    gridDim = {gx, gy, 1};
    blockDim = {bx, by, bz};
    for (unsigned byi = 0; byi < gy; byi++)
        for (unsigned bxi = 0; bxi < gx; bxi++) {
            blockIdx = {bxi, byi, 0};
            for (unsigned t = 0; t < nthreads; t++) init_fiber((int)t);
            unsigned alive = nthreads;
            while (alive) {
                alive = 0;
                for (unsigned t = 0; t < nthreads; t++) {
                    if (!g_fibers[t].alive) continue;
                    g_cur = (int)t;
                    threadIdx = {t % bx, (t / bx) % by, t / (bx * by)};
                    emu_switch(&g_sched_sp, g_fibers[t].sp);
                    alive += g_fibers[t].alive;
                }
            }
        }
    return 0;
}
