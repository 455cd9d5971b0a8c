// Microbenchmark: tensor-memory read / write throughput per SM (tcgen05.ld / tcgen05.st 32x32b.x32) as a function of the number of
// warps issuing.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_bw ldtm_bw.cu && ./ldtm_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
          "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        :: "r"(taddr), "r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]),"r"(r[8]),"r"(r[9]),"r"(r[10]),"r"(r[11]),"r"(r[12]),"r"(r[13]),"r"(r[14]),"r"(r[15]),
           "r"(r[16]),"r"(r[17]),"r"(r[18]),"r"(r[19]),"r"(r[20]),"r"(r[21]),"r"(r[22]),"r"(r[23]),"r"(r[24]),"r"(r[25]),"r"(r[26]),"r"(r[27]),"r"(r[28]),"r"(r[29]),"r"(r[30]),"r"(r[31]) : "memory");
}
// mode 0: ld + wait per iteration (latency-exposed, as the epilogues do); 1: 4 loads in flight then wait; 2: st + wait
__global__ void k(int iters, int mode, long long* out, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);     // lane quarter = warp % 4; column block by warp group
    uint32_t r[32], acc = 0;
    for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
    tmem_st32(base, r); tmem_st32(base + 32, r); tmem_st32(base + 64, r); tmem_st32(base + 96, r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    if (mode == 0) {
        for (int it = 0; it < iters; ++it) { tmem_ld32(base + (it & 3) * 32, r); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += r[it & 31]; }
    } else if (mode == 1) {
        uint32_t r1[32], r2[32], r3[32];
        for (int it = 0; it < iters; it += 4) {
            tmem_ld32(base, r); tmem_ld32(base + 32, r1); tmem_ld32(base + 64, r2); tmem_ld32(base + 96, r3);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += r[it & 31] + r1[it & 31] + r2[it & 31] + r3[it & 31];
        }
    } else {
        for (int it = 0; it < iters; ++it) { r[it & 31] += it; tmem_st32(base + (it & 3) * 32, r); asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + r[3];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}
int main() {
    long long* out; uint32_t* sink;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&sink, 1024 * 1024 * 4);
    const int iters = 4096;
    for (int mode = 0; mode < 3; ++mode)
        for (int warps : {1, 4, 8, 16}) {
            k<<<1, warps * 32>>>(iters, mode, out, sink);
            cudaError_t e = cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
            const double bytes = (double)iters * warps * 32 * 32 * 4;
            printf("mode %d (%s) warps %2d: %.1f cycles per x32 access per warp, %.1f B/clk per SM  %s\n", mode, mode == 0 ? "ld+wait" : mode == 1 ? "4 ld in flight" : "st+wait", warps,
                   (double)c / iters, bytes / c, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
