// pgroup_tc.cu -- tensor-core kernel for the pattern-grouped format (kernel K4 of SURVEY.md 2.2):
// tcgen05.mma kind::tf32 with the 3xTF32 error-compensated split, accumulators in TMEM.
//
// Per pattern group g (see pgroup.cu) the product  Y[rows[g], :] = V_g (G x K) . X[cols[g], :] (K x N)
// is a true small GEMM (G = 96..512 output channels, K = 9*Cin+1 = 865..4609, N = batch), which is what
// north_star allows on the tensor pipe.  Orientation: the BATCH is the UMMA M dimension (always a full
// 128 lanes), the group's rows are the UMMA N dimension (any multiple of 16 up to 256, so G = 96 or 192
// wastes nothing):
//
//     D[128 batch][Gp] (TMEM, fp32)  +=  A[128 batch][8 k] (smem, MN-major)  .  B[8 k][Gp] (smem, K-major)
//
//   * B = weight block V_g: pre-split at pack time into hi = v & 0xffffe000 (exactly representable in
//     TF32) and lo = v - hi, stored K-contiguous, loaded by TMA (SWIZZLE_64B boxes of 16 k x Gp rows).
//   * A = gathered activation rows X[cols[g][k], n0:n0+128]: four producer warps read each 512-byte row
//     segment with coalesced LDG.128, split it into hi/lo in registers and store both into the UMMA
//     canonical MN-major SWIZZLE_128B_BASE32B layout -- the only MN-major layout tcgen05 accepts for 32-bit
//     operands (atom = 32 batch x 4 k = 512 B, 32-byte chunk c of row k lands at chunk c ^ (k & 3)) -- then
//     fence.proxy.async + mbarrier arrive.
//   * one elected thread issues, per 8-k step, hi.hi + lo.hi + hi.lo (fp32 accumulate in TMEM): the
//     dropped lo.lo term is below 2^-22 relative, so results stay inside the fp32 rtol 1e-4 parity bar,
//     which plain TF32 (10-bit mantissa) would not.
//   * stages of 16 k, ring of kStages (full/empty mbarriers, tcgen05.commit frees a stage), epilogue:
//     tcgen05.ld 32x32b.x16 -> ReLU -> Y[rows[g][j]][n] (32 lanes = 32 consecutive batch columns = 128 B).
//
// One CTA per (group, 128*NB batch columns); NB = 2 batch tiles share every weight stage.
#include "common.cuh"
#include <cuda.h>
#include <string.h>

namespace {

constexpr int kThreads = 256;          // warp 0: TMA(B)  warp 1: MMA  warp 2: TMEM alloc  warp 3: idle  warps 4-7: A producers + epilogue
constexpr int kProducerThreads = 128;
constexpr int KS = 16;                 // k per stage
constexpr int BM = 128;                // batch columns per UMMA (M)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// D[tmem] (+)= A[smem desc] . B[smem desc], kind::tf32, issued by one thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// shared-memory matrix descriptors (cute::UMMA::SmemDescriptor, version 1 = Blackwell)
//   bits [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version | [61,64) layout type
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)layout_type << 61);
}
constexpr uint32_t kLayoutSW128B32 = 1, kLayoutSW64 = 4;

// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, A MN-major, B K-major
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int NB, bool RELU>
__global__ void __launch_bounds__(kThreads, 1)
pg_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
             const int32_t *__restrict__ rows, const int32_t *__restrict__ cols,
             int G, int Gp, int K_pad, int chunks_per_group, int n_stages, uint32_t tmem_cols,
             const float *__restrict__ X, int64_t ldx, float *__restrict__ Y, int64_t ldy, int64_t n_vecs)
{
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment for the swizzle atoms
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const int a_tile_bytes = 4 * 4 * 512;                        // [k-atom 4][batch-atom 4][512] per batch tile, hi or lo
    const int a_stage_bytes = NB * 2 * a_tile_bytes;             // NB batch tiles x (hi, lo)
    const int b_plane_bytes = Gp * KS * 4;                       // Gp rows x 64 B
    const int stage_bytes = a_stage_bytes + 2 * b_plane_bytes;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + (size_t)n_stages * stage_bytes);
    uint64_t *empty_bar = full_bar + n_stages;
    uint64_t *accum_bar = empty_bar + n_stages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t g = blockIdx.x / chunks_per_group;             // pattern group
    const int gchunk = (int)(blockIdx.x - g * chunks_per_group); // 256-row chunk of a very tall group
    const int row0 = gchunk * 256;                               // first group row handled here
    const int64_t nbase = (int64_t)blockIdx.y * (BM * NB);
    const int n_ksteps = K_pad / KS;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < n_stages; s++) { mbar_init(&full_bar[s], kProducerThreads + 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    } else if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: weight block planes (hi, lo), 16 k x Gp rows per stage =====
        if (lane == 0) {
            const int grow = (int)(g * G) + row0;
            for (int ks = 0; ks < n_ksteps; ks++) {
                const int s = ks % n_stages;
                const uint32_t ph = (uint32_t)(ks / n_stages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                unsigned char *bs = smem + (size_t)s * stage_bytes + a_stage_bytes;
                mbar_arrive_expect_tx(&full_bar[s], 2u * (uint32_t)b_plane_bytes);
                tma_load_2d(bs, &map_hi, ks * KS, grow, &full_bar[s]);
                tma_load_2d(bs + b_plane_bytes, &map_lo, ks * KS, grow, &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc(BM, Gp);
            for (int ks = 0; ks < n_ksteps; ks++) {
                const int s = ks % n_stages;
                const uint32_t ph = (uint32_t)(ks / n_stages) & 1u;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t sb = sa + a_stage_bytes;
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    const uint32_t a_hi = sa + b * 2 * a_tile_bytes, a_lo = a_hi + a_tile_bytes;
#pragma unroll
                    for (int kk = 0; kk < KS / 8; kk++) {
                        // A: MN-major SWIZZLE_128B_BASE32B (the only MN-major layout for 32-bit operands): atom = 32 batch x 4 k
                        // (512 B), LBO = 512 (next 32 batch columns), SBO = 2048 (next 4 k); one MMA (8 k) spans two k-atoms
                        const uint64_t da_hi = make_desc(a_hi + kk * 4096, 512, 2048, kLayoutSW128B32);
                        const uint64_t da_lo = make_desc(a_lo + kk * 4096, 512, 2048, kLayoutSW128B32);
                        // B: K-major SW64 (64-byte rows), 8-row groups 512 B apart; second k step = +32 B
                        const uint64_t db_hi = make_desc(sb + kk * 32, 16, 512, kLayoutSW64);
                        const uint64_t db_lo = make_desc(sb + b_plane_bytes + kk * 32, 16, 512, kLayoutSW64);
                        const uint32_t d = tmem_base + (uint32_t)(b * Gp);
                        umma_tf32(d, da_hi, db_hi, idesc, (ks > 0 || kk > 0) ? 1u : 0u);
                        umma_tf32(d, da_lo, db_hi, idesc, 1u);
                        umma_tf32(d, da_hi, db_lo, idesc, 1u);
                    }
                }
                umma_commit(&empty_bar[s]);                    // frees the stage when these MMAs retire
            }
            umma_commit(accum_bar);
        }
    } else if (warp >= 4) {
        // ===== A producers: gather X rows, split hi/lo, store in UMMA MN-major SW128 layout =====
        const int pt = tid - 128;                               // 0..127
        const int chunk = pt & 31;                              // 16-byte chunk inside the 512-byte row segment
        const int rsub = pt >> 5;                               // 0..3
        const int32_t *__restrict__ cg = cols + g * (int64_t)K_pad;
        const int m_atom = chunk >> 3, j = chunk & 7;
        for (int ks = 0; ks < n_ksteps; ks++) {
            const int s = ks % n_stages;
            const uint32_t ph = (uint32_t)(ks / n_stages) & 1u;
            float4 v[NB][4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int k = rsub + 4 * i;                     // 0..15 inside the stage
                const int32_t c = __ldg(cg + ks * KS + k);
                const float *__restrict__ xr = X + (int64_t)c * ldx;
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    const int64_t n = nbase + b * BM + chunk * 4;
                    v[b][i] = (n < n_vecs) ? __ldg(reinterpret_cast<const float4 *>(xr + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            mbar_wait(&empty_bar[s], ph ^ 1u);
            unsigned char *as = smem + (size_t)s * stage_bytes;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int k = rsub + 4 * i;
                const int k4 = k & 3;
                const int off = (k >> 2) * 2048 + m_atom * 512 + k4 * 128 + ((((j >> 1) ^ k4) << 5) | ((j & 1) << 4));
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    const float4 x = v[b][i];
                    float4 hi, lo;
                    hi.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); lo.x = x.x - hi.x;
                    hi.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); lo.y = x.y - hi.y;
                    hi.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); lo.z = x.z - hi.z;
                    hi.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); lo.w = x.w - hi.w;
                    *reinterpret_cast<float4 *>(as + b * 2 * a_tile_bytes + off) = hi;
                    *reinterpret_cast<float4 *>(as + b * 2 * a_tile_bytes + a_tile_bytes + off) = lo;
                }
            }
            fence_proxy_async();                                // generic-proxy stores -> visible to the tensor core (async proxy)
            mbar_arrive(&full_bar[s]);
        }

        // ===== epilogue: TMEM -> registers -> ReLU -> Y =====
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int q = warp - 4;                                 // TMEM lane quarter owned by this warp
        const int g_valid = min(Gp, G - row0);
        const int32_t *__restrict__ rg = rows + g * (int64_t)G + row0;
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const int64_t n = nbase + b * BM + q * 32 + lane;
            for (int c0 = 0; c0 < Gp; c0 += 16) {
                uint32_t r[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * Gp + c0);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                               "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (n < n_vecs) {
#pragma unroll
                    for (int t = 0; t < 16; t++) {
                        if (c0 + t < g_valid) {
                            float y = __uint_as_float(r[t]);
                            if (RELU) y = fmaxf(y, 0.0f);
                            Y[(int64_t)__ldg(rg + c0 + t) * ldy + n] = y;
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(tmem_cols));
    }
}

// hi/lo split of a packed value array (weights): hi exactly representable in TF32
__global__ void split_tf32_kernel(const float *__restrict__ v, int64_t n, float *__restrict__ hi, float *__restrict__ lo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float x = v[i];
        const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        hi[i] = h;
        lo[i] = x - h;
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

template <int NB>
int launch_tc(const CUtensorMap *maps, const int32_t *rows, const int32_t *cols, int64_t n_groups, int G, int K_pad,
              const float *X, int64_t ldx, float *Y, int64_t ldy, int64_t n_vecs, bool relu, cudaStream_t s)
{
    const int chunks_per_group = (G + 255) / 256;
    const int Gp = (G > 256) ? 256 : ((G + 15) / 16) * 16;
    const int a_stage = NB * 2 * 8192, b_plane = Gp * KS * 4, stage = a_stage + 2 * b_plane;
    int n_stages = (int)((227 * 1024 - 1024 - 256) / stage);
    if (n_stages > 6) n_stages = 6;
    KN_REQUIRE(n_stages >= 2, "spmm_pg_tc: stage does not fit shared memory");
    const size_t smem = (size_t)n_stages * stage + 1024 + 256;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < NB * Gp) tmem_cols <<= 1;
    KN_REQUIRE(tmem_cols <= 512, "spmm_pg_tc: accumulator does not fit TMEM");
    const int64_t gx = n_groups * chunks_per_group, gy = kn_cdiv(n_vecs, BM * NB);
    KN_REQUIRE(gx <= 0x7fffffffLL && gy <= 65535, "spmm_pg_tc: grid too large");
    static bool configured = false;
    if (!configured) {
        KN_CUDA(cudaFuncSetAttribute(pg_tc_kernel<NB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        KN_CUDA(cudaFuncSetAttribute(pg_tc_kernel<NB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    dim3 grid((unsigned)gx, (unsigned)gy);
    if (relu) pg_tc_kernel<NB, true><<<grid, kThreads, smem, s>>>(maps[0], maps[1], rows, cols, G, Gp, K_pad, chunks_per_group, n_stages, tmem_cols, X, ldx, Y, ldy, n_vecs);
    else      pg_tc_kernel<NB, false><<<grid, kThreads, smem, s>>>(maps[0], maps[1], rows, cols, G, Gp, K_pad, chunks_per_group, n_stages, tmem_cols, X, ldx, Y, ldy, n_vecs);
    KN_CHECK_LAUNCH();
    return KN_OK;
}
}  // namespace

KN_API int kn_pg_tc_split(const float *vals, int64_t n, float *vals_hi, float *vals_lo, void *stream) {
    KN_REQUIRE(n >= 0, "pg_tc_split: negative length");
    if (n == 0) return KN_OK;
    KN_REQUIRE(vals && vals_hi && vals_lo, "pg_tc_split: null pointer");
    split_tf32_kernel<<<(unsigned)kn_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(vals, n, vals_hi, vals_lo);
    KN_CHECK_LAUNCH();
    return KN_OK;
}

KN_API int kn_pg_tc_tensormaps(const float *vals_hi, const float *vals_lo, int64_t n_rows_total, int32_t G, int32_t K_pad, void *maps_out_host) {
    KN_REQUIRE(vals_hi && vals_lo && maps_out_host, "pg_tc_tensormaps: null pointer");
    KN_REQUIRE(n_rows_total > 0 && G > 0 && K_pad > 0 && K_pad % KS == 0, "pg_tc_tensormaps: bad shape");
    PFN_encodeTiled enc = get_encode();
    if (!enc) { kn_set_error("cuTensorMapEncodeTiled is not available from this driver"); return KN_ERR_UNSUPPORTED; }
    const int Gp = (G > 256) ? 256 : ((G + 15) / 16) * 16;
    const float *planes[2] = {vals_hi, vals_lo};
    for (int i = 0; i < 2; i++) {
        CUtensorMap m;                                         // 64-byte aligned local; the caller's buffer need not be
        cuuint64_t dims[2] = {(cuuint64_t)K_pad, (cuuint64_t)n_rows_total};
        cuuint64_t strides[1] = {(cuuint64_t)K_pad * sizeof(float)};
        cuuint32_t box[2] = {(cuuint32_t)KS, (cuuint32_t)Gp};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)planes[i], dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { kn_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return KN_ERR_CUDA; }
        memcpy(reinterpret_cast<unsigned char *>(maps_out_host) + i * sizeof(CUtensorMap), &m, sizeof(CUtensorMap));
    }
    return KN_OK;
}

KN_API int kn_spmm_pg_tc_f32(const void *maps_host, const int32_t *rows, const int32_t *cols, int64_t n_groups, int32_t G, int32_t K_pad,
                             const float *X, int64_t ldx, float *Y, int64_t ldy, int64_t n_vecs, uint32_t flags, void *stream) {
    KN_REQUIRE(n_groups >= 0 && G > 0 && K_pad > 0 && K_pad % KS == 0, "spmm_pg_tc: bad shape (G=%d K_pad=%d)", G, K_pad);
    KN_REQUIRE(n_vecs >= 0 && ldx >= n_vecs && ldy >= n_vecs, "spmm_pg_tc: bad leading dimension");
    if (n_groups == 0 || n_vecs == 0) return KN_OK;
    KN_REQUIRE(maps_host && rows && cols && X && Y, "spmm_pg_tc: null pointer");
    KN_REQUIRE(n_vecs % 4 == 0 && ldx % 4 == 0 && (((uintptr_t)X) & 15) == 0, "spmm_pg_tc: n_vecs and ldx must be multiples of 4, X 16-byte aligned");
    CUtensorMap maps[2];
    memcpy(maps, maps_host, 2 * sizeof(CUtensorMap));
    const bool relu = (flags & KN_SPMM_RELU) != 0;
    const int Gp = (G > 256) ? 256 : ((G + 15) / 16) * 16;
    // two batch tiles per CTA share every weight stage when the accumulators fit TMEM and the batch is wide enough
    if (2 * Gp <= 512 && n_vecs > BM)
        return launch_tc<2>(maps, rows, cols, n_groups, G, K_pad, X, ldx, Y, ldy, n_vecs, relu, (cudaStream_t)stream);
    return launch_tc<1>(maps, rows, cols, n_groups, G, K_pad, X, ldx, Y, ldy, n_vecs, relu, (cudaStream_t)stream);
}
