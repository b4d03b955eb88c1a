// zkm_arith.cuh -- 32-bit-limb carry-chain primitives for sm_100a.
//
// On the device every primitive is one PTX instruction with the .cc carry flag
// (ptxas fuses adjacent mad.lo.cc / madc.hi.cc pairs into IMAD.WIDE.U32.X with a
// predicate carry, so an n-limb Montgomery product issues ~2n^2+n IMAD.WIDEs on
// the FMA pipe -- checked with cuobjdump, see profiles/).
//
// When compiled by a host compiler with -DZKM_HOST_EMU the same primitives are
// emulated with a thread-local carry flag, so the exact limb schedules used by
// the kernels can be unit-tested on a machine without a GPU.  That build is a
// TEST vehicle only; nothing in the shipped library is built with ZKM_HOST_EMU.
#pragma once
#include <stdint.h>

#if defined(ZKM_HOST_EMU)
#define ZKM_HD inline
#define ZKM_DEV inline
#define ZKM_CONST_ARRAY(name, n) static const uint32_t name[n]
#define ZKM_UNROLL
#define ZKM_CEXPR inline constexpr
#else
#define ZKM_HD __host__ __device__ __forceinline__
#define ZKM_DEV __device__ __forceinline__
#define ZKM_CONST_ARRAY(name, n) static __device__ __constant__ uint32_t name[n]
#define ZKM_UNROLL _Pragma("unroll")
#define ZKM_CEXPR __host__ __device__ __forceinline__ constexpr
#endif

namespace zkm {
namespace ptx {

#if defined(ZKM_HOST_EMU)
static thread_local uint32_t g_cc = 0;
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + g_cc; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + g_cc; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; g_cc = (uint32_t)((t >> 32) & 1); return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - g_cc; g_cc = (uint32_t)((t >> 32) & 1); return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - g_cc; }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c + g_cc; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)mul_hi(a, b) + c; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)mul_hi(a, b) + c + g_cc; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return mul_hi(a, b) + c + g_cc; }
#else
ZKM_DEV uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKM_DEV uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKM_DEV uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKM_DEV uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKM_DEV uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKM_DEV uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKM_DEV uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
ZKM_DEV uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
ZKM_DEV uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZKM_DEV uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZKM_DEV uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZKM_DEV uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZKM_DEV uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#endif

}  // namespace ptx
}  // namespace zkm
