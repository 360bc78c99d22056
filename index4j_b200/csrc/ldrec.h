// FMGPU_LD256(p): fetch one 32-byte record.  On the device a single 256-bit load (LDG.E.256) that bypasses L1
// allocation (records are gathered, not streamed); on the host (layout tests) a plain read.
#pragma once
#include "layout.h"

#if defined(__CUDACC__)
namespace fmgpu {
__device__ __forceinline__ Rec32 ld256(const Rec32* p) {
    Rec32 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p));
    return r;
}
// same load with an L2 eviction priority: evict_last for records worth keeping resident (the compact cell table),
// evict_first for records that stream through (level sectors of a 247 MB table)
__device__ __forceinline__ Rec32 ld256_keep(const Rec32* p) {
    Rec32 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ Rec32 ld256_stream(const Rec32* p) {
    Rec32 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p));
    return r;
}
// one 8-byte (block, symbol) cell
__device__ __forceinline__ Cell8 ldcell(const Cell8* p) {
    Cell8 c;
    asm volatile("ld.global.nc.L1::no_allocate.v2.b32 {%0,%1}, [%2];" : "=r"(c.value), "=r"(c.info) : "l"(p));
    return c;
}
// the same through L1 (allocating): 4 cells share a sector, 16 a line
__device__ __forceinline__ Cell8 ldcell_l1(const Cell8* p) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    Cell8 c;
    c.value = v.x;
    c.info = v.y;
    return c;
}
__device__ __forceinline__ Cell8 ldcell_keep(const Cell8* p) {
    Cell8 c;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v2.b32 {%0,%1}, [%2];" : "=r"(c.value), "=r"(c.info) : "l"(p));
    return c;
}
}  // namespace fmgpu
#endif
// L2 eviction priorities of the backward search's two record families (experiment knob, -DCOUNT_L2_HINTS=n):
// 1 = cells evict_last + occurrence records evict_first, 2 = cells evict_last only, 3 = occurrence records evict_first only,
// 4 = cells through L1 (allocating loads)
#ifndef COUNT_L2_HINTS
#define COUNT_L2_HINTS 3  // measured best together with the persisting window on the cell table (fmgpu.cu l2_window)
#endif
#if defined(__CUDA_ARCH__)
#if COUNT_L2_HINTS == 1 || COUNT_L2_HINTS == 2
#define FMGPU_LDCELL(p) ::fmgpu::ldcell_keep(p)
#elif COUNT_L2_HINTS == 4
#define FMGPU_LDCELL(p) ::fmgpu::ldcell_l1(p)
#else
#define FMGPU_LDCELL(p) ::fmgpu::ldcell(p)
#endif
#if COUNT_L2_HINTS == 1 || COUNT_L2_HINTS == 3
#define FMGPU_LD256_OCC(p) ::fmgpu::ld256_stream(p)
#else
#define FMGPU_LD256_OCC(p) ::fmgpu::ld256(p)
#endif
#define FMGPU_LD256(p) ::fmgpu::ld256(p)
#define FMGPU_LDG32(p) __ldg(p)
#define FMGPU_LDG16(p) __ldg(p)
#else
#define FMGPU_LDCELL(p) (*(p))
#define FMGPU_LD256_OCC(p) (*(p))
#define FMGPU_LD256(p) (*(p))
#define FMGPU_LDG32(p) (*(p))
#define FMGPU_LDG16(p) (*(p))
#endif
