// octree_ptx_host.h -- TEST INFRASTRUCTURE: qubatron_b200/csrc/octree_ptx.cuh's wrappers in plain C++ (see
// cuda_host_shim.h).  Shared-space addresses are 32-bit offsets from an anchor in the library's static data.
#pragma once
#include "cuda_host_shim.h"

extern "C" int qb_emu_dyn_shared[]; // the CTA's dynamic shared memory (emu.cpp)
#define QB_DYN_SHARED(T, name) T* name = (T*) qb_emu_dyn_shared

namespace qb
{
namespace ptx
{

// the device takes MUFU.RCP (~1 ulp) and refines it; here the seed is the correctly rounded reciprocal, which the
// same refinement leaves (almost always) unchanged.  Only the IEEE-division mode uses it.
inline float rcp_approx_ftz(float d) { return 1.0f / d; }

typedef unsigned long long f2;
inline f2 f2pack(float lo, float hi) { return (f2) __float_as_uint(lo) | ((f2) __float_as_uint(hi) << 32); }
inline void f2unpack(f2 v, float& lo, float& hi)
{
    lo = __uint_as_float((unsigned) v);
    hi = __uint_as_float((unsigned) (v >> 32));
}
#define QB_EMU_F2OP(name, op)                                                                                         \
    inline f2 name(f2 a, f2 b)                                                                                        \
    {                                                                                                                 \
        float al, ah, bl, bh;                                                                                         \
        f2unpack(a, al, ah);                                                                                          \
        f2unpack(b, bl, bh);                                                                                          \
        volatile float l = al op bl, h = ah op bh; /* each half rounded once, never contracted */                     \
        return f2pack(l, h);                                                                                          \
    }
QB_EMU_F2OP(f2add, +)
QB_EMU_F2OP(f2sub, -)
QB_EMU_F2OP(f2mul, *)
#undef QB_EMU_F2OP

static char emu_shared_anchor;
inline unsigned shared_addr(const void* p) { return (unsigned) (int) ((const char*) p - &emu_shared_anchor); }
inline void     keep_in_register(unsigned&) {}
inline unsigned* emu_shared_ptr(unsigned addr) { return (unsigned*) (&emu_shared_anchor + (int) addr); }
template <unsigned OFF>
inline void sts_ordered(unsigned addr, unsigned v) { *emu_shared_ptr(addr + OFF) = v; }
template <unsigned OFF>
inline unsigned lds_ordered(unsigned addr) { return *emu_shared_ptr(addr + OFF); }
inline unsigned lds_table(unsigned addr) { return *emu_shared_ptr(addr); }
inline int bfind(unsigned x) { return 31 - __builtin_clz(x); }
inline unsigned prmt(unsigned a, unsigned b, unsigned sel) // default mode: nibble k of sel picks byte 0-7 of {b, a}
{
    const unsigned long long src = (unsigned long long) a | ((unsigned long long) b << 32);
    unsigned                 r   = 0;
    for (int k = 0; k < 4; k++)
    {
        const unsigned s = (sel >> (4 * k)) & 0xfu;
        unsigned       byte = (unsigned) (src >> (8 * (s & 7u))) & 0xffu;
        if (s & 8u) byte = (byte & 0x80u) ? 0xffu : 0u; // sign replication
        r |= byte << (8 * k);
    }
    return r;
}
inline float select_by_kind(float w1, float w2, float w3, int kind) { return kind == 1 ? w1 : (kind == 2 ? w2 : w3); }
inline void     st_release_sys(unsigned* p, unsigned v) { *(volatile unsigned*) p = v; }
inline unsigned ld_acquire_sys(const unsigned* p) { return *(const volatile unsigned*) p; }

} // namespace ptx
} // namespace qb
