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
}  // namespace fmgpu
#endif
#if defined(__CUDA_ARCH__)
#define FMGPU_LD256(p) ::fmgpu::ld256(p)
#define FMGPU_LDG32(p) __ldg(p)
#else
#define FMGPU_LD256(p) (*(p))
#define FMGPU_LDG32(p) (*(p))
#endif
