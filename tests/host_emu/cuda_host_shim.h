// cuda_host_shim.h -- TEST INFRASTRUCTURE.  Lets g++ compile qubatron_b200/csrc's kernel headers for the host so that
// the logic of the traversal (not its code generation) can be checked against the oracle without a GPU: one thread
// at a time, CUDA built-ins stated in plain C++.  Nothing under qubatron_b200/ includes or links this.
#pragma once
#include <cuda_runtime.h> // vector types only (host-compilable)
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#undef __global__
#define __global__ static inline
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __shared__
#define __shared__ static // one CTA at a time: a function-local static array stands in for the CTA's shared array
#undef __constant__
#define __constant__ static
#undef __device__
#define __device__
#undef __forceinline__
#define __forceinline__ inline
#undef __noinline__
#define __noinline__

struct EmuIdx
{
    unsigned x, y, z;
};
static EmuIdx threadIdx, blockIdx, blockDim, gridDim;
using std::max;
using std::min;

template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int      __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float    __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float    __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s)
{
    s &= 31u;
    return s ? (hi << s) | (lo >> (32u - s)) : hi;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned) (((unsigned long long) a * b) >> 32); }
static inline int  __popc(unsigned v) { return __builtin_popcount(v); }
static inline int  __clz(int v) { return v ? __builtin_clz((unsigned) v) : 32; }
static inline int  __float2int_rn(float f) { return (int) lrintf(f); } // round to nearest even (default mode)
static inline void __syncthreads() {}
static inline void __syncwarp() {}
static inline void __threadfence_system() {}
static inline void __nanosleep(unsigned) {}
static inline void __trap() { abort(); }
static inline long long clock64() { return 0; }
// warp exchanges: a thread runs alone.  The frame store is kept on its scalar path by the harness (odd pitch), the
// counter reductions are not emulated (COUNT = false)
template <class T>
static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T>
static inline T __shfl_xor_sync(unsigned, T, int) { return T(0); }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p += v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
