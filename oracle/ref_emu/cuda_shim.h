// TEST INFRASTRUCTURE ONLY (oracle/): SIMT-on-CPU shim that lets g++ compile the
// reference's CUDA kernel file *in place* (/root/reference/src/instagraal/kernels/
// kernel_sparse_adapt.cu) into oracle/_ref/libref_cpu.so.  Nothing here is product code.
// One OS thread runs every CUDA thread of a block as a fiber (emu_runtime.cpp), so
// __shared__ == thread_local static and atomics are plain read-modify-write.
#pragma once
#include <math.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <cmath>
#include <cstdlib>

#define __global__
#define __device__
#define __host__
#define __shared__ thread_local
#define __inline__ inline
#define __forceinline__ inline
#ifndef __restrict__
#define __restrict__ __restrict
#endif

struct emu_uint3 { unsigned x, y, z; };
extern emu_uint3 threadIdx, blockIdx, blockDim, gridDim;
static const int warpSize = 32;

struct int2 { int x, y; };
struct int3 { int x, y, z; };
struct int4 { int x, y, z, w; };
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };

extern "C" void emu_syncthreads(void);
static inline void __syncthreads(void) { emu_syncthreads(); }

static inline int atomicAdd(int* a, int v) { int o = *a; *a = o + v; return o; }
static inline unsigned atomicAdd(unsigned* a, unsigned v) { unsigned o = *a; *a = o + v; return o; }
static inline float atomicAdd(float* a, float v) { float o = *a; *a = o + v; return o; }
static inline double atomicAdd(double* a, double v) { double o = *a; *a = o + v; return o; }

static inline float __int2float_rn(int x) { return (float)x; }
static inline double __int2double_rn(int x) { return (double)x; }
static inline int __float2int_rd(float x) { return (int)floorf(x); }
static inline int __shfl_down_sync(unsigned, int v, unsigned, int = 32) { return v; }

static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline float min(float a, float b) { return fminf(a, b); }
static inline double max(double a, double b) { return fmax(a, b); }
static inline double min(double a, double b) { return fmin(a, b); }
// CUDA's global-namespace math overloads for float arguments
using std::abs;
using std::pow;
using std::exp;
using std::floor;
using std::log10;
using std::sqrt;
using std::cos;
using std::sin;
