// pgtile_tc.cu -- spatially tiled tensor-core kernel for keyed convolution layers with few output channels (G <= 128).
//
// pgroup_tc.cu runs one output pixel per CTA: for every 16-k stage it pulls 8 KB of gathered activations AND the pixel's
// weight slab (hi + lo planes) from L2.  For G <= 128 that is ~64 B/clk/SM at full tensor rate, ~9.5 KB/clk over the chip,
// against an L2 throughput cap of ~6.3 KB/clk: measured 49 % (G = 64) / 62 % (G = 128) of the 3xTF32 tensor ceiling, i.e.
// those layers are bound by L2 -> SM traffic, not by the tensor pipe (VGG16 conv1_2 / conv2_x, AllConvNet conv2 / conv3).
//
// Here one CTA owns a TILE of th x tw neighbouring output pixels (2x2 for G <= 96, 1x2 for G <= 128) x 128 batch columns:
//   * the union of the tile's input positions ((th-1)*stride+P) x ((tw-1)*stride+Q) is gathered ONCE per 16-channel chunk:
//     a 3x3 / stride-1 2x2 tile reads 16 positions instead of 4 x 9 = 36 (each gathered stage feeds up to 4 pixels' MMAs);
//   * under permutation-only keys every pixel multiplies the SAME weight matrix, so the P*Q tap slabs of a channel chunk
//     are loaded ONCE per chunk by TMA into a slab ring and reused by every pixel of the tile at its own position
//     (tap of pixel t at union position p = p - t*stride); taps that fall outside the image are simply not issued;
//   * every pixel has its own fp32 accumulator in TMEM and its own MMA-issuer thread (a thread issues one tcgen05.mma per
//     ~117 cycles, a 128 x 64 x 8 tf32 MMA occupies the pipe for ~33: one issuer per accumulator keeps the pipe fed);
//   * A operand (gathered activations, hi/lo split in registers) in a TMEM ring exactly as in pgroup_tc.cu; 3xTF32
//     (hi.hi + lo.hi + hi.lo), bias as one extra stage reading the homogeneous row.
// L2 -> SM bytes per (pixel, 16-channel chunk): G = 64: 108 KB -> 50 KB, G = 96: 126 -> 59, G = 128: 216 -> 120.
//
// Weight layout (built by the host): Wt[G][K_pad] TAP-major, k = tap * C + c, bias at k = n_taps * C, zero padded to a
// multiple of 16; hi / lo TF32 planes + TMA descriptors with boxes of 16 k x Gp rows (kn_pg_tc_split / kn_pg_tc_tensormaps).
// Tile tables (kn_conv2d_tiles_index): cols[tile][position][C] = activation row of every (position, channel) under the
// layer's input key (-1 = position outside the image), rows[tile][pixel][G] = output row of every (pixel, channel).
#include "tc_common.cuh"
#include <string.h>
#include <stdlib.h>

namespace {
using namespace kn_tc;

constexpr int kThreads = 512;          // warp 0: TMA (weight slabs)  warps 1-6: MMA issuers  warp 7: TMEM alloc  warps 8-15: A producers + epilogue
constexpr int kProducerThreads = 256;
constexpr int kProducerWarp0 = 8;
constexpr int KS = 16;                 // k per stage = channels per chunk
constexpr int BM = 128;                // batch columns per CTA (UMMA M)
constexpr int kMaxT = 6;               // output pixels per tile (issuer warps)
constexpr int kMaxPos = 32;            // union positions per tile
constexpr int kMaxStages = 512;
constexpr int kBiasPos = 255;
constexpr int kRawStageBytes = KS * BM * 4;      // one gathered stage: 16 rows x 128 batch columns, fp32

struct TileGeom {
    int C, G, Gp, th, tw, T, stride, P, Q, n_taps, uh, uw, U_pos, n_chunks;
    int n_slots, n_a, n_raw;           // weight-slab ring (shared memory), activation ring (TMEM) and raw gather ring (shared memory) depths
    uint32_t a0;                       // first TMEM column of the activation ring (after the T accumulators)
    int slab_bytes, plane_bytes;
};

template <bool RELU, bool PEERS>
__global__ void __launch_bounds__(kThreads, 1)
pg_tile_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                  const int32_t *__restrict__ tile_cols, const int32_t *__restrict__ tile_rows, int32_t bias_col,
                  const TileGeom geo, int64_t n_sp_tiles, int64_t n_btiles, int super_tiles,
                  const float *__restrict__ X, int64_t ldx, float *__restrict__ Y, int64_t ldy, int64_t n_vecs, const __grid_constant__ KnPeers peers)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    float *s_raw = reinterpret_cast<float *>(smem + (size_t)geo.n_slots * geo.slab_bytes);          // [n_raw][16 rows][128 batch] gathered fp32
    unsigned char *sp = smem + (size_t)geo.n_slots * geo.slab_bytes + (size_t)geo.n_raw * kRawStageBytes;
    int32_t *s_cols = reinterpret_cast<int32_t *>(sp);                      sp += (size_t)geo.U_pos * geo.C * 4;
    int32_t *s_tap = reinterpret_cast<int32_t *>(sp);                       sp += (size_t)kMaxT * kMaxPos * 4;     // [t][p] -> tap or -1
    int32_t *s_valid = reinterpret_cast<int32_t *>(sp);                     sp += (size_t)kMaxPos * 4;
    uint16_t *s_stage = reinterpret_cast<uint16_t *>(sp);                   sp += (size_t)kMaxStages * 2;         // (chunk << 8) | position
    int32_t *s_nstages = reinterpret_cast<int32_t *>(sp);                   sp += 16;
    uint64_t *fullB = reinterpret_cast<uint64_t *>(sp);
    uint64_t *emptyB = fullB + geo.n_slots;
    uint64_t *fullA = emptyB + geo.n_slots;
    uint64_t *emptyA = fullA + geo.n_a;
    uint64_t *accum_bar = emptyA + geo.n_a;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const KnRaster rt = kn_raster(blockIdx.x, n_sp_tiles, n_btiles, super_tiles);
    const int64_t tile = rt.item;
    const int64_t nbase = rt.tile * BM;
    const int T = geo.T, C = geo.C, Gp = geo.Gp, U_pos = geo.U_pos;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < geo.n_slots; s++) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], T); }
        for (int s = 0; s < geo.n_a; s++) { mbar_init(&fullA[s], kProducerThreads); mbar_init(&emptyA[s], T); }
        mbar_init(accum_bar, T);
        fence_barrier_init();
    } else if (warp == 7) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // the tile's gather table, the (pixel, position) -> tap table and the list of in-image stages
    const int32_t *__restrict__ tc = tile_cols + tile * (int64_t)U_pos * C;
    for (int i = tid; i < U_pos * C; i += kThreads) s_cols[i] = __ldg(tc + i);
    for (int i = tid; i < kMaxT * kMaxPos; i += kThreads) {
        const int t = i / kMaxPos, p = i - t * kMaxPos;
        int tap = -1;
        if (t < T && p < U_pos) {
            const int ty = t / geo.tw, tx = t - ty * geo.tw, py = p / geo.uw, px = p - py * geo.uw;
            const int dy = py - ty * geo.stride, dx = px - tx * geo.stride;
            if (dy >= 0 && dy < geo.P && dx >= 0 && dx < geo.Q) tap = dy * geo.Q + dx;
        }
        s_tap[i] = tap;
    }
    if (tid < kMaxPos) s_valid[tid] = (tid < U_pos && __ldg(tc + (int64_t)tid * C) >= 0) ? 1 : 0;
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int cc = 0; cc < geo.n_chunks; cc++)
            for (int p = 0; p < U_pos; p++)
                if (s_valid[p]) s_stage[n++] = (uint16_t)((cc << 8) | p);
        s_stage[n++] = (uint16_t)((geo.n_chunks << 8) | kBiasPos);          // bias: the homogeneous row against the bias slab
        *s_nstages = n;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_stages = *s_nstages;
    const int n_slabs = geo.n_chunks * geo.n_taps + 1;

    if (warp == 0) {
        // ===== TMA producer: weight slabs (chunk, tap) in order, then the bias slab =====
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int j = 0; j < n_slabs; j++) {
                const int cc = j / geo.n_taps, tap = j - cc * geo.n_taps;
                const int k0 = (j == n_slabs - 1) ? geo.n_taps * C : tap * C + cc * KS;
                mbar_wait(&emptyB[s], ph ^ 1u);
                unsigned char *bs = smem + (size_t)s * geo.slab_bytes;
                mbar_arrive_expect_tx(&fullB[s], (uint32_t)geo.slab_bytes);
                tma_load_2d(bs, &map_hi, k0, 0, &fullB[s]);
                tma_load_2d(bs + geo.plane_bytes, &map_lo, k0, 0, &fullB[s]);
                if (++s == geo.n_slots) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp >= 1 && warp <= kMaxT) {
        // ===== MMA issuers: one elected thread per output pixel of the tile =====
        const int t = warp - 1;
        if (lane == 0 && t < T) {
            const uint32_t idesc = make_idesc(BM, Gp);
            const uint32_t d = tmem_base + (uint32_t)(t * Gp);
            const uint64_t desc0 = make_desc(smem_u32(smem), 16, 512, kLayoutSW64);
            int sa = 0; uint32_t pa = 0;
            uint32_t started = 0;
            auto slab_of = [&](int j, int &slot, uint32_t &ph) { const int w = j / geo.n_slots; slot = j - w * geo.n_slots; ph = (uint32_t)(w & 1); };
            auto issue = [&](int slot) {
                const uint32_t ta = tmem_base + geo.a0 + (uint32_t)(sa * 32);
                const uint64_t dslab = desc0 + (uint64_t)(((uint32_t)slot * (uint32_t)geo.slab_bytes) >> 4);
#pragma unroll
                for (int kk = 0; kk < KS / 8; kk++) {
                    const uint64_t db_hi = dslab + (uint64_t)(kk * 2);
                    const uint64_t db_lo = db_hi + (uint64_t)(geo.plane_bytes >> 4);
                    umma_tf32_ts(d, ta + kk * 8, db_hi, idesc, started);       // x_hi . w_hi
                    started = 1u;
                    umma_tf32_ts(d, ta + 16 + kk * 8, db_hi, idesc, 1u);       // x_lo . w_hi
                    umma_tf32_ts(d, ta + kk * 8, db_lo, idesc, 1u);            // x_hi . w_lo
                }
            };
            for (int cc = 0; cc < geo.n_chunks; cc++) {
                for (int p = 0; p < U_pos; p++) {
                    const int tap = s_tap[t * kMaxPos + p];
                    const bool valid = s_valid[p] != 0;
                    int slot = 0; uint32_t pb = 0;
                    if (tap >= 0) {
                        slab_of(cc * geo.n_taps + tap, slot, pb);
                        mbar_wait(&fullB[slot], pb);
                    }
                    if (valid) {
                        mbar_wait(&fullA[sa], pa);
                        tc_fence_after();
                        if (tap >= 0) { issue(slot); umma_commit(&emptyB[slot]); }
                        umma_commit(&emptyA[sa]);
                        if (++sa == geo.n_a) { sa = 0; pa ^= 1u; }
                    } else if (tap >= 0) {
                        mbar_arrive(&emptyB[slot]);                // tap outside the image: release the slab unused
                    }
                }
            }
            {   // bias stage
                int slot; uint32_t pb;
                slab_of(n_slabs - 1, slot, pb);
                mbar_wait(&fullB[slot], pb);
                mbar_wait(&fullA[sa], pa);
                tc_fence_after();
                issue(slot);
                umma_commit(&emptyB[slot]);
                umma_commit(&emptyA[sa]);
            }
            umma_commit(accum_bar);
        }
    } else if (warp >= kProducerWarp0) {
        // ===== A producers: gather X rows of one (position, channel chunk), split hi/lo, tcgen05.st into the TMEM ring =====
        // The gathers go through a ring of RAW stages in shared memory filled by cp.async (16 B per lane = one 512 B row
        // segment per warp instruction): up to n_raw - 1 stages (8 KB each) are in flight per SM, tracked by commit groups.
        // (Register prefetch -- ld.global into a 4-stage register ring -- ran ONE stage per memory latency, ~1100 cycles:
        // the loads of all ring stages shared scoreboards, so waiting for the oldest stage waited for the youngest too.)
        const int pw = warp - kProducerWarp0;                   // producer warp 0..7
        const int q = warp & 3;                                 // TMEM lane quarter
        const int sel = pw >> 2;                                // k half of the stage
        constexpr int KW = 8;
        const int k0 = sel * KW;
        const int n_raw = geo.n_raw;
        const int64_t ncol = nbase + lane * 4;                  // first batch column of this lane's 16-byte chunk
        const bool col_ok = ncol < n_vecs;                      // n_vecs % 4 == 0: a chunk is entirely inside or outside
        const float *__restrict__ xcol = X + (col_ok ? ncol : 0);
        auto issue_gather = [&](int i) {
            if (i < n_stages) {
                const int st = s_stage[i];
                const int cc = st >> 8, p = st & 255;
                float *dst = s_raw + (size_t)(i % n_raw) * (KS * BM);
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    const int r = pw * 2 + rr;                  // row of the stage handled by this warp
                    const int32_t c = (p == kBiasPos) ? bias_col : s_cols[p * C + cc * KS + r];
                    const float *src = xcol + (int64_t)c * ldx;
                    const unsigned d32 = smem_u32(dst + r * BM + lane * 4);
                    const int bytes = col_ok ? 16 : 0;          // src-size 0 => 16 bytes of zeros
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(d32), "l"(src), "r"(bytes));
                }
            }
            asm volatile("cp.async.commit_group;\n" ::);
        };
        int psa = 0; uint32_t ppa = 0;
        for (int i = 0; i < n_raw - 1; i++) issue_gather(i);
        for (int i = 0; i < n_stages; i++) {
            // stage i has landed for this thread's copies; the barrier makes every producer's copies visible and tells that
            // all of them are done reading stage i-1, whose slot the next gather overwrites
            switch (n_raw - 2) {
                case 1: asm volatile("cp.async.wait_group 1;\n" ::: "memory"); break;
                case 2: asm volatile("cp.async.wait_group 2;\n" ::: "memory"); break;
                case 3: asm volatile("cp.async.wait_group 3;\n" ::: "memory"); break;
                case 4: asm volatile("cp.async.wait_group 4;\n" ::: "memory"); break;
                case 5: asm volatile("cp.async.wait_group 5;\n" ::: "memory"); break;
                case 6: asm volatile("cp.async.wait_group 6;\n" ::: "memory"); break;
                case 7: asm volatile("cp.async.wait_group 7;\n" ::: "memory"); break;
                case 8: asm volatile("cp.async.wait_group 8;\n" ::: "memory"); break;
                case 9: asm volatile("cp.async.wait_group 9;\n" ::: "memory"); break;
                case 10: asm volatile("cp.async.wait_group 10;\n" ::: "memory"); break;
                case 11: asm volatile("cp.async.wait_group 11;\n" ::: "memory"); break;
                case 12: asm volatile("cp.async.wait_group 12;\n" ::: "memory"); break;
                case 13: asm volatile("cp.async.wait_group 13;\n" ::: "memory"); break;
                case 14: asm volatile("cp.async.wait_group 14;\n" ::: "memory"); break;
                default: asm volatile("cp.async.wait_group 0;\n" ::: "memory"); break;
            }
            asm volatile("bar.sync 1, 256;\n" ::: "memory");
            issue_gather(i + n_raw - 1);
            const float *src = s_raw + (size_t)(i % n_raw) * (KS * BM) + q * 32 + lane;
            uint32_t hi[KW], lo[KW];
#pragma unroll
            for (int j = 0; j < KW; j++) {
                const float v = src[(k0 + j) * BM];
                hi[j] = __float_as_uint(v) & 0xFFFFE000u;
                lo[j] = __float_as_uint(v - __uint_as_float(hi[j]));
            }
            mbar_wait(&emptyA[psa], ppa ^ 1u);
            tc_fence_after();
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + geo.a0 + (uint32_t)(psa * 32 + k0);
            tmem_store<KW>(ta, hi);
            tmem_store<KW>(ta + 16, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(&fullA[psa]);
            if (++psa == geo.n_a) { psa = 0; ppa ^= 1u; }
        }
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");

        // ===== epilogue: TMEM -> registers -> ReLU -> Y rows of every pixel of the tile =====
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int half = sel;                                    // two warps per lane quarter: even / odd 16-column chunks
        const int64_t ne = nbase + q * 32 + lane;
        for (int t = 0; t < T; t++) {
            const int32_t *__restrict__ rg = tile_rows + (tile * T + t) * (int64_t)geo.G;
            for (int c0 = half * 16; c0 < Gp; c0 += 32) {
                uint32_t r[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * Gp + c0);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                               "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (ne < n_vecs) {
                    if constexpr (!PEERS) {
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            if (c0 + j < geo.G) {
                                float y = __uint_as_float(r[j]);
                                if (RELU) y = fmaxf(y, 0.0f);
                                Y[(int64_t)__ldg(rg + c0 + j) * ldy + ne] = y;
                            }
                        }
                    } else {
                        int32_t yrow[16];
                        unsigned pmask[16];
#pragma unroll
                        for (int j = 0; j < 16; j++) yrow[j] = (c0 + j < geo.G) ? __ldg(rg + c0 + j) : -1;
#pragma unroll
                        for (int j = 0; j < 16; j++) pmask[j] = (yrow[j] >= 0) ? kn_peer_mask(peers, yrow[j]) : 0u;
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            float y = __uint_as_float(r[j]);
                            if (RELU) y = fmaxf(y, 0.0f);
                            r[j] = __float_as_uint(y);
                        }
                        for (int pp = 0; pp < peers.n; pp++) {
                            float *__restrict__ yb = peers.y[pp];
#pragma unroll
                            for (int j = 0; j < 16; j++)
                                if ((pmask[j] >> pp) & 1u) yb[(int64_t)yrow[j] * ldy + ne] = __uint_as_float(r[j]);
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 7) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base));
    }
}

// ---- tile tables from the layer geometry -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv_tiles_index_kernel(kn_conv2d_desc d, const int32_t *__restrict__ tile_origin, int64_t n_tiles, int th, int tw,
                        const int32_t *__restrict__ row_of_src, const int32_t *__restrict__ col_map,
                        int32_t *__restrict__ tile_cols, int32_t *__restrict__ tile_rows)
{
    const int Uo = d.U / d.stride, Vo = d.V / d.stride;
    const int uh = (th - 1) * d.stride + d.P, uw = (tw - 1) * d.stride + d.Q;
    const int ph = (d.P - 1) / 2, qh = (d.Q - 1) / 2;
    const int U_pos = uh * uw, T = th * tw;
    for (int64_t tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
        const int o = tile_origin[tl];
        const int oy = o / Vo, ox = o - oy * Vo;
        for (int i = threadIdx.x; i < U_pos * d.C; i += blockDim.x) {
            const int p = i / d.C, c = i - p * d.C;
            const int py = p / uw, px = p - py * uw;
            const int iy = oy * d.stride - ph + py, ix = ox * d.stride - qh + px;
            int32_t col = -1;
            if (iy >= 0 && iy < d.U && ix >= 0 && ix < d.V) {
                const int32_t cs = c * d.U * d.V + iy * d.V + ix;
                col = col_map ? __ldg(col_map + cs) : cs;
            }
            tile_cols[tl * (int64_t)U_pos * d.C + i] = col;
        }
        for (int i = threadIdx.x; i < T * d.M; i += blockDim.x) {
            const int t = i / d.M, m = i - t * d.M;
            const int ty = t / tw, tx = t - ty * tw;
            const int64_t s = (int64_t)m * Uo * Vo + (int64_t)(oy + ty) * Vo + (ox + tx);
            tile_rows[tl * (int64_t)T * d.M + i] = row_of_src ? row_of_src[s] : (int32_t)s;
        }
    }
}

int tile_super_tiles() {
    static const int v = getenv("KN_TILE_SUPER") ? atoi(getenv("KN_TILE_SUPER")) : 8;
    return v > 0 ? v : 8;
}
}  // namespace

KN_API int kn_conv2d_tiles_index(const kn_conv2d_desc *desc, const int32_t *tile_origin, int64_t n_tiles, int32_t th, int32_t tw,
                                 const int32_t *row_of_src, const int32_t *col_map, int32_t *tile_cols, int32_t *tile_rows, void *stream) {
    KN_REQUIRE(desc && desc->C > 0 && desc->M > 0 && desc->stride > 0 && (desc->P % 2) == 1 && (desc->Q % 2) == 1, "conv_tiles: bad descriptor");
    KN_REQUIRE(n_tiles >= 0 && th > 0 && tw > 0, "conv_tiles: bad tile shape");
    if (n_tiles == 0) return KN_OK;
    KN_REQUIRE(tile_origin && tile_cols && tile_rows, "conv_tiles: null pointer");
    const int64_t cap = (int64_t)kn_sm_count() * 8;
    conv_tiles_index_kernel<<<(unsigned)(n_tiles < cap ? n_tiles : cap), 256, 0, (cudaStream_t)stream>>>(*desc, tile_origin, n_tiles, th, tw, row_of_src, col_map, tile_cols, tile_rows);
    KN_CHECK_LAUNCH();
    return KN_OK;
}

KN_API int kn_spmm_tile_tc_f32(const void *maps_host, const int32_t *tile_cols, const int32_t *tile_rows, int32_t bias_col, int64_t n_tiles,
                               int32_t C, int32_t G, int32_t th, int32_t tw, int32_t stride, int32_t P, int32_t Q,
                               const float *X, int64_t ldx, float *Y, int64_t ldy, int64_t n_vecs, uint32_t flags, const kn_peers *peers_arg, void *stream) {
    KnPeersScope peers_scope(peers_arg);
    if (!peers_scope.ok) return KN_ERR_INVALID_ARGUMENT;
    KN_REQUIRE(n_tiles >= 0 && C > 0 && C % KS == 0 && G > 0 && th > 0 && tw > 0 && stride > 0 && P > 0 && Q > 0, "spmm_tile_tc: bad shape (C=%d G=%d)", C, G);
    KN_REQUIRE(n_vecs >= 0 && ldx >= n_vecs && ldy >= n_vecs, "spmm_tile_tc: bad leading dimension");
    if (n_tiles == 0 || n_vecs == 0) return KN_OK;
    KN_REQUIRE(maps_host && tile_cols && tile_rows && X && Y, "spmm_tile_tc: null pointer");
    KN_REQUIRE(n_vecs % 4 == 0 && ldx % 4 == 0 && (((uintptr_t)X) & 15) == 0, "spmm_tile_tc: n_vecs and ldx must be multiples of 4, X 16-byte aligned");
    KN_REQUIRE(ldx * 4 < 0xffffffffLL, "spmm_tile_tc: leading dimension in bytes must fit 32 bits");
    TileGeom g;
    g.C = C; g.G = G; g.Gp = ((G + 15) / 16) * 16; g.th = th; g.tw = tw; g.T = th * tw; g.stride = stride; g.P = P; g.Q = Q; g.n_taps = P * Q;
    g.uh = (th - 1) * stride + P; g.uw = (tw - 1) * stride + Q; g.U_pos = g.uh * g.uw; g.n_chunks = C / KS;
    KN_REQUIRE(g.T <= kMaxT && g.U_pos <= kMaxPos && g.Gp <= 256, "spmm_tile_tc: tile too large (T=%d positions=%d)", g.T, g.U_pos);
    KN_REQUIRE(g.n_chunks * g.U_pos + 1 <= kMaxStages && g.n_chunks < 255, "spmm_tile_tc: too many stages");
    g.a0 = (uint32_t)(g.T * g.Gp);
    g.n_a = (int)((512 - g.a0) / 32);
    if (g.n_a > 8) g.n_a = 8;
    KN_REQUIRE(g.a0 <= 512 && g.n_a >= 2, "spmm_tile_tc: accumulators of %d pixels x %d rows leave no room for the activation ring", g.T, g.Gp);
    g.plane_bytes = g.Gp * KS * 4;
    g.slab_bytes = 2 * g.plane_bytes;
    const size_t fixed = (size_t)g.U_pos * C * 4 + (size_t)kMaxT * kMaxPos * 4 + kMaxPos * 4 + kMaxStages * 2 + 16 + 1024 /*align*/ + 1024 /*barriers*/;
    const int64_t avail = 226 * 1024 - (int64_t)fixed;
    // shared memory: the weight slabs of at least one channel chunk (+1 so the next chunk can start loading), the rest split
    // between the raw gather ring (bytes in flight towards this SM) and more weight slabs
    int n_raw = (int)((avail - (int64_t)(g.n_taps + 1) * g.slab_bytes) / kRawStageBytes);
    if (n_raw > 16) n_raw = 16;
    KN_REQUIRE(n_raw >= 3, "spmm_tile_tc: the weight slabs of one channel chunk do not fit shared memory (Gp=%d taps=%d)", g.Gp, g.n_taps);
    int n_slots = (int)((avail - (int64_t)n_raw * kRawStageBytes) / g.slab_bytes);
    if (n_slots > 3 * g.n_taps) n_slots = 3 * g.n_taps;
    KN_REQUIRE(n_slots >= g.n_taps + 1, "spmm_tile_tc: the weight slabs of one channel chunk do not fit shared memory (Gp=%d taps=%d)", g.Gp, g.n_taps);
    KN_REQUIRE((size_t)(2 * n_slots + 2 * g.n_a + 2) * 8 <= 1024, "spmm_tile_tc: too many barriers");
    g.n_slots = n_slots;
    g.n_raw = n_raw;
    const size_t smem = (size_t)n_slots * g.slab_bytes + (size_t)n_raw * kRawStageBytes + fixed;
    const int64_t n_btiles = kn_cdiv(n_vecs, BM);
    KN_REQUIRE(n_tiles * n_btiles <= 0x7fffffffLL, "spmm_tile_tc: grid too large");
    KN_ONCE_PER_DEVICE {
        KN_CUDA(cudaFuncSetAttribute(pg_tile_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        KN_CUDA(cudaFuncSetAttribute(pg_tile_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        KN_CUDA(cudaFuncSetAttribute(pg_tile_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        KN_CUDA(cudaFuncSetAttribute(pg_tile_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    CUtensorMap maps[4];
    memcpy(maps, maps_host, 4 * sizeof(CUtensorMap));
    const KnPeers peers = kn_current_peers();
    const bool relu = (flags & KN_SPMM_RELU) != 0;
    const int super_ = tile_super_tiles();
    const dim3 grid((unsigned)(n_tiles * n_btiles));
    cudaStream_t s = (cudaStream_t)stream;
#define KN_TILE_LAUNCH(R, PP) pg_tile_tc_kernel<R, PP><<<grid, kThreads, smem, s>>>(maps[0], maps[1], tile_cols, tile_rows, bias_col, g, n_tiles, n_btiles, super_, X, ldx, Y, ldy, n_vecs, peers)
    if (peers.n > 0) { if (relu) KN_TILE_LAUNCH(true, true); else KN_TILE_LAUNCH(false, true); }
    else             { if (relu) KN_TILE_LAUNCH(true, false); else KN_TILE_LAUNCH(false, false); }
#undef KN_TILE_LAUNCH
    KN_CHECK_LAUNCH();
    return KN_OK;
}
