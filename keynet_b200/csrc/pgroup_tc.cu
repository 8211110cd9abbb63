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
//   * A = gathered activation rows X[cols[g][k], n0:n0+128], kept in TENSOR MEMORY (lane = batch element,
//     column = k): eight producer warps read coalesced 128-byte row segments, split them into hi/lo in
//     registers and write both with tcgen05.st into a TMEM ring; the MMA takes A from TMEM, so shared memory
//     carries only the weight ring.  (A first version staged A in shared memory, in the MN-major
//     SWIZZLE_128B_BASE32B layout -- the only MN-major layout tcgen05 accepts for 32-bit operands; it was
//     bound by the 128 B/clk shared-memory port: 7 KB of operands per 48-cycle N=96 MMA.)
//   * one elected thread issues, per 8-k step, hi.hi + lo.hi + hi.lo (fp32 accumulate in TMEM): the
//     dropped lo.lo term is below 2^-22 relative, so results stay inside the fp32 rtol 1e-4 parity bar,
//     which plain TF32 (10-bit mantissa) would not.
//   * stages of 16 k, ring of kStages (full/empty mbarriers, tcgen05.commit frees a stage), epilogue:
//     tcgen05.ld 32x32b.x16 -> ReLU -> Y[rows[g][j]][n] (32 lanes = 32 consecutive batch columns = 128 B).
//
// One CTA per (group, 128*NB batch columns); NB = 2 batch tiles share every weight stage.
#include "tc_common.cuh"
#include <string.h>
#include <stdlib.h>

namespace {
using namespace kn_tc;

constexpr int kThreads = 384;          // warp 0: TMA(B)  warps 1,3: MMA issuers  warp 2: TMEM alloc  warps 4-11: A producers + epilogue
constexpr int kProducerThreads = 256;
constexpr int KS = 16;                 // k per stage
constexpr int BM = 128;                // batch columns per UMMA (M)
constexpr int kSuperTiles = 8;         // rasterisation super-tile: 8 x 128 = 1024 batch columns (sweep 4..32: 4/8 best by ~1 %)

// Shared memory holds only the weight ring; the activation operand lives in TMEM:
//   TMEM columns [0, NB*Gp)                      fp32 accumulators, one Gp-wide block per batch tile
//   TMEM columns [a0 + (sa*NB + b)*32, +32)      A ring: 16 hi columns then 16 lo columns (lane = batch element, column = k)
// A tf32 UMMA with N = 96 needs 4 KB of A and 3 KB of B per 48 cycles; with A in shared memory that is 146 B/clk, above
// the 128 B/clk shared-memory port, so the tensor pipe starves.  From TMEM the A operand costs no shared-memory
// bandwidth and the producers' stores (tcgen05.st) bypass shared memory as well.
// phase timestamps of one CTA (debug aid, read back with kn_debug_tc_timing): entry, prologue done, first stage full,
// last MMA issued, accumulators ready, epilogue stored, teardown
__device__ long long g_tc_timing[8];
__device__ int g_tc_timing_cta = -1;
#define KN_STAMP(i) do { if ((int)blockIdx.x == g_tc_timing_cta) g_tc_timing[i] = clock64(); } while (0)

template <int NB, bool RELU, bool DUAL, int CS, bool PEERS>
__global__ void __launch_bounds__(kThreads, 1)
pg_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
             const int32_t *__restrict__ rows, const int32_t *__restrict__ cols, const int32_t *__restrict__ group_k, const int32_t *__restrict__ block_of,
             int G, int Gp, int K_pad, int chunks_per_group, int64_t n_items, int64_t n_tiles, int n_b, int n_a, uint32_t a0, int super_tiles,
             const float *__restrict__ X, int64_t ldx, float *__restrict__ Y, int64_t ldy, int64_t n_vecs, const __grid_constant__ KnPeers peers)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const int b_plane_bytes = Gp * KS * 4;                       // Gp rows x 64 B
    const int stage_bytes = 2 * b_plane_bytes;                   // hi plane, lo plane
    int32_t *s_cols = reinterpret_cast<int32_t *>(smem + (size_t)n_b * stage_bytes);        // the group's gather list, K_pad entries
    uint64_t *fullB = reinterpret_cast<uint64_t *>(smem + (size_t)n_b * stage_bytes + (size_t)K_pad * 4);
    uint64_t *emptyB = fullB + n_b;
    uint64_t *fullA = emptyB + n_b;
    uint64_t *emptyA = fullA + n_a;
    uint64_t *accum_bar = emptyA + n_a;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const KnRaster rt = kn_raster(blockIdx.x, n_items, n_tiles, super_tiles);
    const int64_t g = rt.item / chunks_per_group;                // pattern group
    const int gchunk = (int)(rt.item - g * chunks_per_group);    // 256-row chunk of a very tall group
    const int row0 = gchunk * 256;                               // first group row handled here
    const int64_t nbase = rt.tile * (BM * NB);
    // groups of one class share K_pad (storage) but loop only over their own K (edge / corner pixels have fewer taps)
    const int n_ksteps = group_k ? (__ldg(group_k + g) + KS - 1) / KS : K_pad / KS;
    const uint32_t crank = (CS > 1) ? cluster_ctarank() : 0u;
    // DUAL (Gp <= 128): the hi and lo weight planes are adjacent in shared memory, so ONE UMMA with N = 2*Gp computes
    // x_hi.[w_hi ; w_lo] into two accumulator blocks (summed in the epilogue): 2 instructions per k-step instead of 3.
    const int acc_w = DUAL ? 2 * Gp : Gp;                        // accumulator columns per batch tile

    if (tid == 0) KN_STAMP(0);
    if (warp == 0 && lane == 0) {
        // NB MMA issuers (one per batch tile): each commits once per stage to both rings and once to accum_bar
        // CS > 1: the CTAs of a cluster (same group, adjacent batch tiles) share every weight stage by TMA multicast, so a
        // stage is free only when the issuers of ALL cluster CTAs have retired it
        for (int s = 0; s < n_b; s++) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], NB * CS); }
        for (int s = 0; s < n_a; s++) { mbar_init(&fullA[s], kProducerThreads); mbar_init(&emptyA[s], NB); }
        mbar_init(accum_bar, NB);
        fence_barrier_init();
    } else if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // the column list is read once per CTA: a gather then never waits for an index load before it can issue
    for (int i = tid; i < n_ksteps * KS; i += kThreads) s_cols[i] = __ldg(cols + g * (int64_t)K_pad + i);
    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_all();                              // partner barriers are initialised before any multicast
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) KN_STAMP(1);

    if (warp == 0) {
        // ===== TMA producer: weight block planes (hi, lo), 16 k x Gp rows per stage =====
        if (elect_one()) {
            const int grow = (int)((block_of ? (int64_t)__ldg(block_of + g) : g) * G) + row0;      // rows of the group's unique value block
            int s = 0; uint32_t ph = 0;                          // ring position / phase kept incrementally (no runtime div/mod)
            for (int ks = 0; ks < n_ksteps; ks++) {
                mbar_wait(&emptyB[s], ph ^ 1u);
                unsigned char *bs = smem + (size_t)s * stage_bytes;
                mbar_arrive_expect_tx(&fullB[s], 2u * (uint32_t)b_plane_bytes);
                if (CS == 1) {
                    tma_load_2d(bs, &map_hi, ks * KS, grow, &fullB[s]);
                    tma_load_2d(bs + b_plane_bytes, &map_lo, ks * KS, grow, &fullB[s]);
                } else {
                    // this CTA fetches rows [rank*Gp/CS, (rank+1)*Gp/CS) of both planes and multicasts them to the cluster
                    const int part = Gp / CS, r0 = (int)crank * part;
                    tma_load_2d_mcast(bs + r0 * KS * 4, &map_hi, ks * KS, grow + r0, &fullB[s], (uint16_t)((1u << CS) - 1u));
                    tma_load_2d_mcast(bs + b_plane_bytes + r0 * KS * 4, &map_lo, ks * KS, grow + r0, &fullB[s], (uint16_t)((1u << CS) - 1u));
                }
                if (++s == n_b) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1 || (warp == 3 && NB == 2)) {
        // ===== MMA issuers: one elected thread per batch tile (warp 1 -> batch tile 0, warp 3 -> batch tile 1) =====
        if (elect_one()) {
            const int b = (warp == 1) ? 0 : 1;
            const uint32_t idesc = make_idesc(BM, Gp);
            const uint32_t idesc2 = make_idesc(BM, 2 * Gp);
            const uint32_t d = tmem_base + (uint32_t)(b * acc_w);
            const uint64_t desc0 = make_desc(smem_u32(smem), 16, 512, kLayoutSW64);
            int sb = 0, sa = 0; uint32_t pb = 0, pa = 0;
            for (int ks = 0; ks < n_ksteps; ks++) {
                mbar_wait(&fullB[sb], pb);
                mbar_wait(&fullA[sa], pa);
                tc_fence_after();
                if (ks == 0 && warp == 1) KN_STAMP(2);
                const uint32_t ta = tmem_base + a0 + (uint32_t)((sa * NB + b) * 32);
                // B: K-major SW64 (64-byte rows), 8-row groups 512 B apart; stage / plane / k-step only move the
                // 14-bit start-address field (units of 16 B), so the descriptor is one 64-bit add away from desc0
                const uint64_t dstage = desc0 + (uint64_t)(((uint32_t)sb * (uint32_t)stage_bytes) >> 4);
#pragma unroll
                for (int kk = 0; kk < KS / 8; kk++) {
                    const uint64_t db_hi = dstage + (uint64_t)(kk * 2);
                    const uint64_t db_lo = db_hi + (uint64_t)(b_plane_bytes >> 4);
                    if (DUAL) {
                        umma_tf32_ts(d, ta + kk * 8, db_hi, idesc2, (ks > 0 || kk > 0) ? 1u : 0u); // x_hi . [w_hi ; w_lo]
                        umma_tf32_ts(d, ta + 16 + kk * 8, db_hi, idesc, 1u);                       // x_lo . w_hi
                    } else {
                        umma_tf32_ts(d, ta + kk * 8, db_hi, idesc, (ks > 0 || kk > 0) ? 1u : 0u);  // x_hi . w_hi
                        umma_tf32_ts(d, ta + 16 + kk * 8, db_hi, idesc, 1u);                       // x_lo . w_hi
                        umma_tf32_ts(d, ta + kk * 8, db_lo, idesc, 1u);                            // x_hi . w_lo
                    }
                }
                if (CS == 1) umma_commit(&emptyB[sb]);         // both rings are released when every issuer's MMAs retire
                else umma_commit_mcast(&emptyB[sb], (uint16_t)((1u << CS) - 1u));
                umma_commit(&emptyA[sa]);
                if (++sb == n_b) { sb = 0; pb ^= 1u; }
                if (++sa == n_a) { sa = 0; pa ^= 1u; }
            }
            umma_commit(accum_bar);
            if (warp == 1) KN_STAMP(3);
        }
    } else if (warp >= 4) {
        // ===== A producers: gather X, split hi/lo, tcgen05.st into the TMEM A ring =====
        // thread = one batch element (TMEM lane); it loads X[cols[k]][n] for the k's of its stage share -- each warp
        // load is one coalesced 128-byte row segment -- and runs PD stages ahead of the TMEM stores (register ring).
        const int q = warp & 3;                                 // TMEM lane quarter of this warp
        const int sel = (warp - 4) >> 2;                        // NB == 2: batch tile; NB == 1: k half of the stage
        constexpr int KW = (NB == 2) ? 16 : 8;                  // k values per thread and stage
        constexpr int PD = 4;                                   // prefetch distance in stages
        const int b_mine = (NB == 2) ? sel : 0;
        const int k0 = (NB == 2) ? 0 : sel * 8;
        const int64_t n = nbase + b_mine * BM + q * 32 + lane;
        const bool n_ok = n < n_vecs;
        // address of X[c][n] = xaddr + c * (ldx * 4): one IMAD.WIDE.U32 + one ld.global.nc per element.  Batch lanes past
        // n_vecs are clamped to the last valid column: their TMEM rows only feed outputs the epilogue never stores.
        const uint64_t xaddr = reinterpret_cast<uint64_t>(X) + (uint64_t)(n_ok ? n : (n_vecs - 1)) * 4ull;
        const uint32_t ldxb = (uint32_t)ldx * 4u;
        float buf[PD][KW];

        auto gather = [&](float (&v)[KW], int ks) {
            const int4 *cp = reinterpret_cast<const int4 *>(s_cols + ks * KS + k0);             // 16-byte aligned
#pragma unroll
            for (int i4 = 0; i4 < KW / 4; i4++) {
                const int4 c = cp[i4];                          // warp-broadcast LDS.128: 4 column indices
                const uint32_t cc[4] = {(uint32_t)c.x, (uint32_t)c.y, (uint32_t)c.z, (uint32_t)c.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint64_t a;
                    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"(cc[j]), "r"(ldxb), "l"(xaddr));
                    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v[4 * i4 + j]) : "l"(a));
                }
            }
        };
        int psa = 0; uint32_t ppa = 0;                           // producer ring position / phase
        auto publish = [&](const float (&v)[KW]) {
            const int sa = psa;
            uint32_t hi[KW], lo[KW];
#pragma unroll
            for (int i = 0; i < KW; i++) {
                hi[i] = __float_as_uint(v[i]) & 0xFFFFE000u;
                lo[i] = __float_as_uint(v[i] - __uint_as_float(hi[i]));
            }
            mbar_wait(&emptyA[sa], ppa ^ 1u);
            tc_fence_after();
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + a0 + (uint32_t)((sa * NB + b_mine) * 32 + k0);
            tmem_store<KW>(ta, hi);
            tmem_store<KW>(ta + 16, lo);
        };
        // completing a stage (wait for the TMEM stores, then signal the issuers) is deferred until the next
        // gather has been issued, so the store latency overlaps with load issue instead of serialising every stage
        auto finish = [&]() {
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(&fullA[psa]);
            if (++psa == n_a) { psa = 0; ppa ^= 1u; }
        };
#pragma unroll
        for (int d = 0; d < PD; d++)
            if (d < n_ksteps) gather(buf[d], d);
        for (int ks0 = 0; ks0 < n_ksteps; ks0 += PD) {
#pragma unroll
            for (int d = 0; d < PD; d++) {
                const int ks = ks0 + d;
                if (ks < n_ksteps) {
                    publish(buf[d]);
                    if (ks + PD < n_ksteps) gather(buf[d], ks + PD);
                    finish();
                }
            }
        }

        // ===== epilogue: TMEM -> registers -> ReLU -> Y =====
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        if (tid == 128) KN_STAMP(4);
        const int half = (warp - 4) >> 2;                       // two warps per lane quarter: even / odd 16-column chunks
        const int g_valid = min(Gp, G - row0);
        const int32_t *__restrict__ rg = rows + g * (int64_t)G + row0;
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const int64_t ne = nbase + b * BM + q * 32 + lane;
            for (int c0 = half * 16; c0 < Gp; c0 += 32) {
                uint32_t r[16], r2[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * acc_w + c0);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                               "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                             : "r"(taddr));
                if (DUAL) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                                 : "=r"(r2[0]), "=r"(r2[1]), "=r"(r2[2]), "=r"(r2[3]), "=r"(r2[4]), "=r"(r2[5]), "=r"(r2[6]), "=r"(r2[7]),
                                   "=r"(r2[8]), "=r"(r2[9]), "=r"(r2[10]), "=r"(r2[11]), "=r"(r2[12]), "=r"(r2[13]), "=r"(r2[14]), "=r"(r2[15])
                                 : "r"(taddr + (uint32_t)Gp));
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (ne < n_vecs) {
                    if constexpr (!PEERS) {
#pragma unroll
                        for (int t = 0; t < 16; t++) {
                            if (c0 + t < g_valid) {
                                float y = __uint_as_float(r[t]);
                                if (DUAL) y += __uint_as_float(r2[t]);
                                if (RELU) y = fmaxf(y, 0.0f);
                                Y[(int64_t)__ldg(rg + c0 + t) * ldy + ne] = y;
                            }
                        }
                    } else {
                        // fused all-gather (separate instantiation: this code costs the plain epilogue ~10 %): row ids and
                        // need masks first, then one pass of NVLink stores per peer
                        int32_t yrow[16];
                        unsigned pmask[16];
#pragma unroll
                        for (int t = 0; t < 16; t++) yrow[t] = (c0 + t < g_valid) ? __ldg(rg + c0 + t) : -1;
#pragma unroll
                        for (int t = 0; t < 16; t++) pmask[t] = (yrow[t] >= 0) ? kn_peer_mask(peers, yrow[t]) : 0u;
#pragma unroll
                        for (int t = 0; t < 16; t++) {
                            float y = __uint_as_float(r[t]);
                            if (DUAL) y += __uint_as_float(r2[t]);
                            if (RELU) y = fmaxf(y, 0.0f);
                            r[t] = __float_as_uint(y);
                        }
                        for (int p = 0; p < peers.n; p++) {
                            float *__restrict__ yb = peers.y[p];
#pragma unroll
                            for (int t = 0; t < 16; t++)
                                if ((pmask[t] >> p) & 1u) yb[(int64_t)yrow[t] * ldy + ne] = __uint_as_float(r[t]);
                        }
                    }
                }
            }
        }
    }

    if (tid == 128) KN_STAMP(5);
    tc_fence_before();
    __syncthreads();
    if (tid == 0) KN_STAMP(6);
    if (CS > 1) cluster_sync_all();                              // no CTA leaves while a partner may still signal its barriers
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base));
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

// batch columns (in 128-column tiles) per raster super-tile: small enough that the X rows shared by neighbouring groups
// (3x3 windows re-read every row 9 times, one image row of groups apart) are still L2-resident when they are re-read
static int tc_super_tiles() {
    static const int v = getenv("KN_TC_SUPER") ? atoi(getenv("KN_TC_SUPER")) : kSuperTiles;
    return v > 0 ? v : kSuperTiles;
}

template <int NB, bool DUAL, int CS>
int launch_tc(const CUtensorMap *maps, const int32_t *rows, const int32_t *cols, const int32_t *group_k, const int32_t *block_of, int64_t n_groups, int G, int K_pad,
              const float *X, int64_t ldx, float *Y, int64_t ldy, int64_t n_vecs, bool relu, cudaStream_t s)
{
    const int chunks_per_group = (G + 255) / 256;
    const int Gp = (G > 256) ? 256 : ((G + 15) / 16) * 16;
    const int stage = 2 * Gp * KS * 4;
    int n_b = (int)((220 * 1024 - (int64_t)K_pad * 4) / stage);   // weight ring in shared memory, next to the column list
    if (n_b > 8) n_b = 8;
    const uint32_t a0 = (uint32_t)(NB * (DUAL ? 2 * Gp : Gp));   // activation ring in TMEM, after the accumulators
    int n_a = (int)((512 - a0) / (NB * 32));
    if (n_a > 6) n_a = 6;
    KN_REQUIRE(n_b >= 2 && n_a >= 2, "spmm_pg_tc: rings do not fit (Gp=%d NB=%d)", Gp, NB);
    const size_t smem = (size_t)n_b * stage + (size_t)K_pad * 4 + 1024 + 512;
    const int64_t gx = n_groups * chunks_per_group, gy = kn_cdiv(n_vecs, BM * NB);
    KN_REQUIRE(gx * gy <= 0x7fffffffLL, "spmm_pg_tc: grid too large");
    KN_ONCE_PER_DEVICE {
        KN_CUDA(cudaFuncSetAttribute(pg_tc_kernel<NB, true, DUAL, CS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        KN_CUDA(cudaFuncSetAttribute(pg_tc_kernel<NB, false, DUAL, CS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        KN_CUDA(cudaFuncSetAttribute(pg_tc_kernel<NB, true, DUAL, CS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        KN_CUDA(cudaFuncSetAttribute(pg_tc_kernel<NB, false, DUAL, CS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    const CUtensorMap *m = maps + (CS > 1 ? 2 : 0);              // maps[2..3]: boxes of Gp/2 rows for the multicast halves
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(gx * gy));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int n_b_ = n_b, n_a_ = n_a;
    const int super_ = tc_super_tiles() / NB > 0 ? tc_super_tiles() / NB : 1;
    const KnPeers peers = kn_current_peers();
#define KN_TC_LAUNCH(R, P) KN_CUDA(cudaLaunchKernelEx(&cfg, pg_tc_kernel<NB, R, DUAL, CS, P>, m[0], m[1], rows, cols, group_k, block_of, G, Gp, K_pad, \
                                                     chunks_per_group, gx, gy, n_b_, n_a_, a0, super_, X, ldx, Y, ldy, n_vecs, peers))
    if (peers.n > 0) { if (relu) KN_TC_LAUNCH(true, true); else KN_TC_LAUNCH(false, true); }
    else             { if (relu) KN_TC_LAUNCH(true, false); else KN_TC_LAUNCH(false, false); }
#undef KN_TC_LAUNCH
    return KN_OK;
}

// cluster of 2 along the batch-tile axis whenever the tiles pair up inside a raster super-tile
template <int NB, bool DUAL>
int launch_tc_auto(const CUtensorMap *maps, const int32_t *rows, const int32_t *cols, const int32_t *group_k, const int32_t *block_of, int64_t n_groups, int G, int K_pad,
                   const float *X, int64_t ldx, float *Y, int64_t ldy, int64_t n_vecs, bool relu, cudaStream_t s)
{
    const int Gp = (G > 256) ? 256 : ((G + 15) / 16) * 16;
    const int64_t n_tiles = kn_cdiv(n_vecs, BM * NB);
    const int64_t S = tc_super_tiles() / NB > 0 ? tc_super_tiles() / NB : 1;
    const bool pairable = (n_tiles % 2 == 0) && (S % 2 == 0) && (Gp % 16 == 0) && ((n_tiles % S) % 2 == 0);
    if (pairable) return launch_tc<NB, DUAL, 2>(maps, rows, cols, group_k, block_of, n_groups, G, K_pad, X, ldx, Y, ldy, n_vecs, relu, s);
    return launch_tc<NB, DUAL, 1>(maps, rows, cols, group_k, block_of, n_groups, G, K_pad, X, ldx, Y, ldy, n_vecs, relu, s);
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
    const float *planes[4] = {vals_hi, vals_lo, vals_hi, vals_lo};
    for (int i = 0; i < 4; i++) {
        CUtensorMap m;                                         // 64-byte aligned local; the caller's buffer need not be
        cuuint64_t dims[2] = {(cuuint64_t)K_pad, (cuuint64_t)n_rows_total};
        cuuint64_t strides[1] = {(cuuint64_t)K_pad * sizeof(float)};
        cuuint32_t box[2] = {(cuuint32_t)KS, (cuuint32_t)(i < 2 ? Gp : Gp / 2)};   // maps 2,3: half boxes for 2-CTA multicast
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)planes[i], dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { kn_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return KN_ERR_CUDA; }
        memcpy(reinterpret_cast<unsigned char *>(maps_out_host) + i * sizeof(CUtensorMap), &m, sizeof(CUtensorMap));
    }
    return KN_OK;
}

KN_API int kn_spmm_pg_tc_f32(const void *maps_host, const int32_t *rows, const int32_t *cols, const int32_t *group_k, const int32_t *block_of, int64_t n_groups, int32_t G, int32_t K_pad,
                             const float *X, int64_t ldx, float *Y, int64_t ldy, int64_t n_vecs, uint32_t flags, const kn_peers *peers_arg, void *stream) {
    KnPeersScope peers_scope(peers_arg);
    if (!peers_scope.ok) return KN_ERR_INVALID_ARGUMENT;
    KN_REQUIRE(n_groups >= 0 && G > 0 && K_pad > 0 && K_pad % KS == 0, "spmm_pg_tc: bad shape (G=%d K_pad=%d)", G, K_pad);
    KN_REQUIRE(n_vecs >= 0 && ldx >= n_vecs && ldy >= n_vecs, "spmm_pg_tc: bad leading dimension");
    if (n_groups == 0 || n_vecs == 0) return KN_OK;
    KN_REQUIRE(maps_host && rows && cols && X && Y, "spmm_pg_tc: null pointer");
    KN_REQUIRE(n_vecs % 4 == 0 && ldx % 4 == 0 && (((uintptr_t)X) & 15) == 0, "spmm_pg_tc: n_vecs and ldx must be multiples of 4, X 16-byte aligned");
    KN_REQUIRE(ldx * 4 < 0xffffffffLL, "spmm_pg_tc: leading dimension in bytes must fit 32 bits");
    CUtensorMap maps[4];
    memcpy(maps, maps_host, 4 * sizeof(CUtensorMap));
    const bool relu = (flags & KN_SPMM_RELU) != 0;
    const int Gp = (G > 256) ? 256 : ((G + 15) / 16) * 16;
    // two batch tiles per CTA share every weight stage when accumulators + a >= 2-deep activation ring fit TMEM
    static const int use_dual = getenv("KN_TC_DUAL") ? atoi(getenv("KN_TC_DUAL")) : 1;   // experiment switch
    if (use_dual && Gp <= 96 && n_vecs > BM)          // 2 tiles x 2*Gp accumulator columns + a 2-deep activation ring fit the 512 TMEM columns
        return launch_tc_auto<2, true>(maps, rows, cols, group_k, block_of, n_groups, G, K_pad, X, ldx, Y, ldy, n_vecs, relu, (cudaStream_t)stream);
    if (use_dual && Gp <= 128)
        return launch_tc_auto<1, true>(maps, rows, cols, group_k, block_of, n_groups, G, K_pad, X, ldx, Y, ldy, n_vecs, relu, (cudaStream_t)stream);
    if (2 * Gp + 2 * 64 <= 512 && n_vecs > BM)
        return launch_tc_auto<2, false>(maps, rows, cols, group_k, block_of, n_groups, G, K_pad, X, ldx, Y, ldy, n_vecs, relu, (cudaStream_t)stream);
    return launch_tc_auto<1, false>(maps, rows, cols, group_k, block_of, n_groups, G, K_pad, X, ldx, Y, ldy, n_vecs, relu, (cudaStream_t)stream);
}

// debug: select the CTA whose phase timestamps are recorded (-1 = none) / read them back (7 x clock64)
KN_API int kn_debug_tc_timing(int32_t cta, int64_t *out_host) {
    if (out_host) {
        long long t[8];
        KN_CUDA(cudaMemcpyFromSymbol(t, g_tc_timing, sizeof(t)));
        for (int i = 0; i < 8; i++) out_host[i] = (int64_t)t[i];
    }
    int c = cta;
    KN_CUDA(cudaMemcpyToSymbol(g_tc_timing_cta, &c, sizeof(int)));
    return KN_OK;
}
