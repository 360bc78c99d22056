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
}  // namespace fmgpu
#endif
#if defined(__CUDA_ARCH__)
#define FMGPU_LD256(p) ::fmgpu::ld256(p)
#define FMGPU_LDG32(p) __ldg(p)
#define FMGPU_LDG16(p) __ldg(p)
#else
#define FMGPU_LD256(p) (*(p))
#define FMGPU_LDG32(p) (*(p))
#define FMGPU_LDG16(p) (*(p))
#endif
