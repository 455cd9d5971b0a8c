// Microbenchmark: does the tcgen05.ld SHAPE change the tensor-memory read throughput?  32 registers per thread each:
// 32x32b.x32 (row per thread, what the epilogues use), 16x256b.x8, 16x128b.x16, 16x64b.x32.  4 and 8 warps, load + wait per iteration.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define REGS32 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}"
#define OUTS32 "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),"=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
template <int S> __device__ __forceinline__ void ld(uint32_t a, uint32_t (&r)[32]) {
    if (S == 0) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " REGS32 ", [%32];" : OUTS32 : "r"(a) : "memory");
    if (S == 1) asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 " REGS32 ", [%32];" : OUTS32 : "r"(a) : "memory");
    if (S == 2) asm volatile("tcgen05.ld.sync.aligned.16x128b.x16.b32 " REGS32 ", [%32];" : OUTS32 : "r"(a) : "memory");
    if (S == 3) asm volatile("tcgen05.ld.sync.aligned.16x64b.x32.b32 " REGS32 ", [%32];" : OUTS32 : "r"(a) : "memory");
}
template <int S> __global__ void k(int iters, long long* out, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
    uint32_t r[32], acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) { ld<S>(base + (it & 1) * 64, r); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += r[it & 31]; }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}
template <int S> void run(const char* name, long long* out, uint32_t* sink) {
    const int iters = 4096;
    for (int warps : {1, 4, 8}) {
        k<S><<<1, warps * 32>>>(iters, out, sink);
        cudaError_t e = cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
        printf("%-12s warps %d: %.1f cycles per 4 KB access per warp, %.1f B/clk per SM %s\n", name, warps, (double)c / iters, (double)iters * warps * 4096 / c, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
}
int main() {
    long long* out; uint32_t* sink;
    cudaMalloc(&out, 64); cudaMalloc(&sink, 4096);
    run<0>("32x32b.x32", out, sink); run<1>("16x256b.x8", out, sink); run<2>("16x128b.x16", out, sink); run<3>("16x64b.x32", out, sink);
    return 0;
}
