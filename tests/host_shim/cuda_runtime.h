// TEST INFRASTRUCTURE -- a host stand-in for <cuda_runtime.h>: just enough of the CUDA device vocabulary to compile
// csrc/core.cuh and csrc/grid.cuh (the per-state geometry code and the broad-phase builders) for the CPU with g++
// -ffp-contract=off, so that the CPU suite can run the very device functions against the oracle, bit for bit, before any
// GPU time is spent.  Every *_rn intrinsic is the corresponding single IEEE operation; fmaf is the C fmaf.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(x) alignas(x)

struct float4 { float x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
struct dim3 { unsigned x = 1, y = 1, z = 1; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static thread_local dim3 blockIdx, threadIdx, blockDim, gridDim;

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline int __float2int_rd(float a) { return (int)floorf(a); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline float __uint_as_float(unsigned v) { float f; memcpy(&f, &v, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned v; memcpy(&v, &f, 4); return v; }
using std::max;
using std::min;
