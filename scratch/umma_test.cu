// minimal tcgen05 tf32 MMA test: D[128 x 32] = A[128 x 8] * B[8 x 32]
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t a, uint32_t lbo, uint32_t sbo, uint32_t lt) {
    return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)lt << 61);
}
__global__ void k(float *out, int a_mn_major, int use_mask_form) {
    extern __shared__ unsigned char sraw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)sraw + 1023) & ~(uintptr_t)1023);
    float *A = (float *)smem;              // 4 KB: MN-major: [m atom 4][k8 8][32 n] swizzled; K-major: [128 rows][32 B]... use SW128 K-major: rows of 128B, only first 32B used
    float *B = (float *)(smem + 16384);    // K-major SW64: [32 rows][64 B]
    __shared__ uint64_t bar; __shared__ uint32_t slot;
    int tid = threadIdx.x;
    // logical A[m][k] = (m % 7) + 0.5f * k ; B[k][n] = (n % 5) - 0.25f * k
    for (int i = tid; i < 16384 / 4; i += blockDim.x) { A[i] = 0.f; }
    for (int i = tid; i < 4096 / 4; i += blockDim.x) { B[i] = 0.f; }
    __syncthreads();
    for (int i = tid; i < 128 * 8; i += blockDim.x) {
        int m = i / 8, kk = i % 8;
        float v = (float)(m % 7) + 0.5f * kk;
        int off;
        if (a_mn_major) { int atom = m / 32, byte = (m % 32) * 4, c32 = byte >> 5, rest = byte & 31, k4 = kk & 3; off = atom * 512 + (kk >> 2) * 2048 + k4 * 128 + ((c32 ^ k4) << 5) + rest; }
        else { int r8 = m % 8, grp = m / 8; int chunk = kk / 4; off = grp * 1024 + r8 * 128 + ((chunk ^ r8) << 4) + (kk % 4) * 4; }
        *(float *)(smem + off) = v;
    }
    for (int i = tid; i < 32 * 8; i += blockDim.x) {
        int n = i / 8, kk = i % 8;
        float v = (float)(n % 5) - 0.25f * kk;
        int r8 = n % 8, grp = n / 8; int chunk = kk / 4;
        int off = grp * 512 + r8 * 64 + ((chunk ^ ((r8 >> 1) & 3)) << 4) + (kk % 4) * 4;
        *(float *)(smem + 16384 + off) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" :: "r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tb = slot;
    if (tid == 32) {
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | (0u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        uint64_t da = a_mn_major ? make_desc(smem_u32(smem), 512, 2048, 1) : make_desc(smem_u32(smem), 16, 1024, 2);
        uint64_t db = make_desc(smem_u32(smem + 16384), 16, 512, 4);
        if (use_mask_form) {
            uint32_t z = 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5,%6,%7,%8}, p;\n\t}"
                         :: "r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(0u), "r"(z), "r"(z), "r"(z), "r"(z) : "memory");
        } else {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         :: "r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    if (tid >= 128) {
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" :: "r"(smem_u32(&bar)), "r"(0u) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        int q = (tid >> 5) - 4, lane = tid & 31;
        for (int c0 = 0; c0 < 32; c0 += 16) {
            uint32_t r[16];
            uint32_t ta = tb + ((uint32_t)(q * 32) << 16) + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(ta));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int t = 0; t < 16; t++) out[(q * 32 + lane) * 32 + c0 + t] = __uint_as_float(r[t]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" :: "r"(tb)); }
}
int main() {
    float *d; cudaMalloc(&d, 128 * 32 * 4);
    float h[128 * 32];
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int mn = 0; mn < 2; mn++) for (int mf = 0; mf < 2; mf++) {
        cudaMemset(d, 0xff, sizeof(h));
        k<<<1, 256, 32 * 1024>>>(d, mn, mf);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double maxerr = 0; int nz = 0;
        for (int m = 0; m < 128; m++) for (int n = 0; n < 32; n++) {
            double ref = 0; for (int kk = 0; kk < 8; kk++) ref += ((m % 7) + 0.5 * kk) * ((n % 5) - 0.25 * kk);
            double err = fabs(h[m * 32 + n] - ref); if (err > maxerr) maxerr = err; if (h[m * 32 + n] != 0) nz++;
        }
        printf("a_mn_major=%d mask_form=%d: %s maxerr=%g nonzero=%d  D[1][1]=%g D[5][3]=%g D[100][17]=%g\n", mn, mf, cudaGetErrorString(e), maxerr, nz, h[33], h[5 * 32 + 3], h[100 * 32 + 17]);
    }
    return 0;
}
