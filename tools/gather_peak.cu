// Micro-benchmark: how many independent random 32-byte sectors per second can this GPU gather from HBM?
// (the practical roofline of a dependent-gather workload whose unit of access is one DRAM sector)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_peak tools/gather_peak.cu && ./gather_peak
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

struct alignas(32) Rec32 { uint32_t w[8]; };
__device__ __forceinline__ Rec32 ld256(const Rec32* p) {
    Rec32 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p));
    return r;
}
// ILP independent chains per thread; each chain is DEPENDENT (next index derived from the loaded data), like an LF / rank walk
template <int ILP>
__global__ void k_gather(const Rec32* buf, uint64_t n_rec, int iters, uint32_t* out) {
    uint64_t idx[ILP];
    uint32_t acc = 0;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int k = 0; k < ILP; ++k) idx[k] = __umul64hi((tid + 1) * 0x9E3779B97F4A7C15ULL + k * 0xBF58476D1CE4E5B9ULL, n_rec);
    for (int it = 0; it < iters; ++it) {
        Rec32 r[ILP];
#pragma unroll
        for (int k = 0; k < ILP; ++k) r[k] = ld256(buf + idx[k]);
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            acc += r[k].w[1];
            // range reduction by multiply-high (a 64-bit '%' costs ~100 instructions and made v1 of this tool ALU-bound)
            idx[k] = __umul64hi((idx[k] + r[k].w[0] + it) * 6364136223846793005ULL + 1442695040888963407ULL, n_rec);
        }
    }
    out[tid] = acc;
}
template <int ILP>
void run(const Rec32* buf, uint64_t n_rec, uint32_t* out, int threads_per_sm, int sms) {
    const int block = 256, grid = sms * threads_per_sm / block, iters = 200;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_gather<ILP><<<grid, block>>>(buf, n_rec, 20, out);
    cudaEventRecord(e0);
    k_gather<ILP><<<grid, block>>>(buf, n_rec, iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double sectors = (double)grid * block * ILP * iters;
    printf("{\"buffer_gb\": %.2f, \"threads_per_sm\": %d, \"ilp\": %d, \"gsectors_per_s\": %.2f, \"gb_per_s\": %.1f}\n",
           n_rec * 32 / 1e9, threads_per_sm, ILP, sectors / ms / 1e6, sectors * 32 / ms / 1e6);
}
int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    for (double gb : {0.37, 0.73, 4.0}) {
        const uint64_t n_rec = (uint64_t)(gb * 1e9 / 32);
        Rec32* buf;
        uint32_t* out;
        cudaMalloc(&buf, n_rec * 32);
        cudaMemset(buf, 1, n_rec * 32);
        cudaMalloc(&out, (size_t)p.multiProcessorCount * 2048 * 4);
        for (int t : {512, 1024, 2048}) {
            run<1>(buf, n_rec, out, t, p.multiProcessorCount);
            run<2>(buf, n_rec, out, t, p.multiProcessorCount);
            run<4>(buf, n_rec, out, t, p.multiProcessorCount);
            run<8>(buf, n_rec, out, t, p.multiProcessorCount);
        }
        cudaFree(buf);
        cudaFree(out);
    }
    return 0;
}
