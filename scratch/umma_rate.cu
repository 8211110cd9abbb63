// tcgen05 tf32 MMA issue-rate microbenchmark: cycles per MMA for various N, A from smem or TMEM
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t a, uint32_t lbo, uint32_t sbo, uint32_t lt) {
    return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)lt << 61);
}
__global__ void k(long long *out, int N, int a_tmem, int iters, int ndst, int nacc) {
    extern __shared__ unsigned char sraw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)sraw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar; __shared__ uint32_t slot;
    int tid = threadIdx.x;
    for (int i = tid; i < 48 * 1024 / 4; i += blockDim.x) ((float *)smem)[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tb = slot;
    // issuers: lane 0 of warps 1..n_issuers; issuer w uses accumulator (w-1)
    const int w = tid >> 5;
#ifdef ELECT
    uint32_t elected = 0;
    if (w >= 1 && w <= nacc) asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(elected));
    if (elected) {
#else
    if ((tid & 31) == 0 && w >= 1 && w <= nacc) {
#endif
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
        uint64_t da = make_desc(smem_u32(smem), 16, 1024, 2);
        uint64_t db = make_desc(smem_u32(smem + 16384), 16, 1024, 2);
        const uint32_t td0 = (ndst == 0) ? tb : tb + (uint32_t)((w - 1) * ndst * N);   // issuer w cycles over ndst accumulators of its own; ndst = 0: every issuer accumulates into the SAME accumulator
        __shared__ uint64_t bars[4];
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[w - 1])));
        asm volatile("fence.mbarrier_init.release.cluster;");
        long long t0 = clock64();
        for (int i = 0; i < iters; i += 8) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint32_t td = td0 + (uint32_t)((ndst ? u % ndst : 0) * N);
                if (a_tmem) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" :: "r"(td), "r"(tb + 448), "l"(db), "r"(idesc), "r"(1u) : "memory");
                else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" :: "r"(td), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bars[w - 1])) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" :: "r"(smem_u32(&bars[w - 1])), "r"(0u) : "memory");
        long long t1 = clock64();
        if (blockIdx.x == 0 && w == 1) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tb)); }
}
int main() {
    long long *d; cudaMalloc(&d, 8); long long h;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    int iters = 4096;
    for (int at = 0; at < 2; at++) for (int N : {32, 64, 96, 128, 192, 256}) for (int nacc : {1, 2, 3, 4}) for (int ndst : {0, 1, 2}) {
        if (nacc * ndst * N > 448) continue;
        k<<<148, 192, 64 * 1024>>>(d, N, at, iters, ndst, nacc);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("tf32 A=%s N=%3d issuers=%d accumulators/issuer=%d: %s %.1f cycles per MMA per issuer, %.1f per MMA overall\n", at ? "tmem" : "smem", N, nacc, ndst, cudaGetErrorString(e), (double)h / iters, (double)h / iters / nacc);
    }
    return 0;
}
