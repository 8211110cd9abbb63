// pgcluster.cu -- pattern groups clustered by spatial neighbourhood, gathered rows staged once in shared memory.
//
// The small-group kernel (pg_small_kernel, pgroup.cu) gathers every X row a group reads straight from L2: a 3x3
// convolution re-reads each input row 9 times, and at LeNet sizes (G = 6..16 output channels, K = 10..55 taps) the
// product is bound by L2 -> SM gather bandwidth, not by HBM or the FMA pipe.  Neighbouring output pixels share most
// of their taps -- also under permutation / gain keys, which only relabel rows and columns -- so the builder
// (sparse.PatternGroups, using the pixel of the underlying Toeplitz row as the hint) bundles the groups of a tile of
// output pixels into a CLUSTER with one union column list:
//
//     cluster c:  ucols[cl_uptr[c] .. cl_uptr[c+1])     union of the columns its groups read (<= u_max rows)
//                 groups cl_gptr[c] .. cl_gptr[c+1])    lidx[g][k] = byte offset of column k in the staged tile
//                 valsT[block][k][GM]                   value block, k-major (4 rows per uniform 128-bit load)
//
// CTA = (cluster, 128 batch columns): the union rows are staged once with cp.async (512 B per row, coalesced), then
// every warp walks groups of the cluster: per k one conflict-free LDS.128 of the staged row feeds 4*GM FFMA.  The
// L2 -> SM traffic drops from K rows per group to (union / groups) rows per group (conv 3x3, 7x7 pixel tile: 10 -> 1.7).
// Pool layers (every row its own pattern, G = 1) are clustered per (channel, pixel tile) the same way.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int TN = 128;          // batch columns per CTA (lane l owns columns 4l .. 4l+3)

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, bool valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 16 : 0;                     // src-size 0 => 16 bytes of zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(dst), "l"(gmem_src), "r"(bytes));
}

template <int GM, bool RELU>
__global__ void __launch_bounds__(kThreads, (GM > 8) ? 2 : 3)
pg_cluster_kernel(const int32_t *__restrict__ cl_gptr, const int32_t *__restrict__ cl_uptr, const int32_t *__restrict__ ucols,
                  const int32_t *__restrict__ rows, const int32_t *__restrict__ lidx, const float *__restrict__ valsT,
                  const int32_t *__restrict__ group_k, const int32_t *__restrict__ block_of, int64_t n_clusters, int G, int K_pad, int u_max, int n_chunks, int gm_total,
                  const float *__restrict__ X, int64_t ldx, float *__restrict__ Y, int64_t ldy, int64_t n_vecs, const __grid_constant__ KnPeers peers)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // clusters fastest: the CTAs in flight work on one batch tile, so the halo rows two clusters share hit in L2
    const int64_t tile = blockIdx.x / n_clusters, c = blockIdx.x - tile * n_clusters;
    const int64_t n0 = tile * TN + lane * 4;
    const bool ok = n0 < n_vecs;
    const float *__restrict__ xb = X + (ok ? n0 : 0);

    // ---- stage: union rows of X (512 B each) and the cluster's lidx table.  The row indices are fetched by one
    //      coalesced load per warp and broadcast by shuffle -- a dependent index load per row would serialise the
    //      whole stage behind L2 latency.
    const int u0 = __ldg(cl_uptr + c), U = __ldg(cl_uptr + c + 1) - u0;
    const int g_beg = __ldg(cl_gptr + c), g_end = __ldg(cl_gptr + c + 1);
    constexpr int kWarps = kThreads / 32;
    {
        const int u_mine = warp + kWarps * lane;                               // lane i holds the index of this warp's i-th row
        const int my_col = (u_mine < U) ? __ldg(ucols + u0 + u_mine) : 0;
        const int n_mine = (U - warp + kWarps - 1) / kWarps;                   // rows this warp stages (<= 28 <= 32)
        for (int i = 0; i < n_mine; i++) {
            const int col = __shfl_sync(0xffffffffu, my_col, i);
            cp_async16(smem_raw + (size_t)(warp + kWarps * i) * (TN * 4) + lane * 16, xb + (int64_t)col * ldx, ok);
        }
        int32_t *s_lidx = reinterpret_cast<int32_t *>(smem_raw + (size_t)u_max * (TN * 4));
        const int n_l = (g_end - g_beg) * K_pad;                               // multiple of 32 ints: whole 16-byte chunks
        const int32_t *__restrict__ src = lidx + (int64_t)g_beg * K_pad;
        for (int i = threadIdx.x * 4; i < n_l; i += kThreads * 4) cp_async16(s_lidx + i, src + i, true);
    }
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();

    const unsigned char *__restrict__ sx = smem_raw + lane * 16;
    const int32_t *__restrict__ s_lidx = reinterpret_cast<const int32_t *>(smem_raw + (size_t)u_max * (TN * 4));
    // work item = (group, chunk of GM rows): tall groups (first conv layer of an RGB network: G = 64 / 96 output channels over
    // 28 taps) are walked GM rows at a time against the same staged rows
    const int n_items = (g_end - g_beg) * n_chunks;
    for (int item = warp; item < n_items; item += kWarps) {
        const int g = g_beg + item / n_chunks;
        const int r0 = (item - (g - g_beg) * n_chunks) * GM;
        const int K = group_k ? __ldg(group_k + g) : K_pad;
        const int64_t blk = block_of ? (int64_t)__ldg(block_of + g) : g;
        const int32_t *__restrict__ li = s_lidx + (g - g_beg) * K_pad;
        const float *__restrict__ vt = valsT + blk * (int64_t)K_pad * gm_total + r0;
        float acc[GM][4];
#pragma unroll
        for (int r = 0; r < GM; r++) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0f; }
#pragma unroll 4
        for (int k = 0; k < K; k++) {
            const float4 x = *reinterpret_cast<const float4 *>(sx + li[k]);
#pragma unroll
            for (int r2 = 0; r2 < GM; r2 += 2) {                                                   // GM even: exact row count for G = 6
                const float2 a = __ldg(reinterpret_cast<const float2 *>(vt + k * gm_total + r2)); // warp-uniform: one L1 sector
                acc[r2 + 0][0] = fmaf(a.x, x.x, acc[r2 + 0][0]); acc[r2 + 0][1] = fmaf(a.x, x.y, acc[r2 + 0][1]); acc[r2 + 0][2] = fmaf(a.x, x.z, acc[r2 + 0][2]); acc[r2 + 0][3] = fmaf(a.x, x.w, acc[r2 + 0][3]);
                acc[r2 + 1][0] = fmaf(a.y, x.x, acc[r2 + 1][0]); acc[r2 + 1][1] = fmaf(a.y, x.y, acc[r2 + 1][1]); acc[r2 + 1][2] = fmaf(a.y, x.z, acc[r2 + 1][2]); acc[r2 + 1][3] = fmaf(a.y, x.w, acc[r2 + 1][3]);
            }
        }
        if (ok) {
#pragma unroll
            for (int r4 = 0; r4 < GM; r4 += 2) {                              // row ids / need masks of 2 rows, then their stores
                int32_t yrows[2];
                unsigned pmask[2];
#pragma unroll
                for (int i = 0; i < 2; i++) yrows[i] = (r0 + r4 + i < G) ? __ldg(rows + (int64_t)g * G + r0 + r4 + i) : 0;
                if (peers.n != 0) {
#pragma unroll
                    for (int i = 0; i < 2; i++) pmask[i] = kn_peer_mask(peers, yrows[i]);
                }
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const int r = r4 + i;
                    if (r0 + r < G) {
                        float4 o = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
                        if (RELU) { o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f); o.z = fmaxf(o.z, 0.0f); o.w = fmaxf(o.w, 0.0f); }
                        const int64_t yoff = (int64_t)yrows[i] * ldy + n0;
                        if (peers.n == 0) *reinterpret_cast<float4 *>(Y + yoff) = o;
                        else KN_FOR_EACH_DEST(peers, Y, pmask[i], yb) *reinterpret_cast<float4 *>(yb + yoff) = o;
                    }
                }
            }
        }
    }
}

template <int GM>
int launch_cluster(const int32_t *cl_gptr, const int32_t *cl_uptr, const int32_t *ucols, const int32_t *rows, const int32_t *lidx, const float *valsT,
                   const int32_t *group_k, const int32_t *block_of, int64_t n_clusters, int G, int K_pad, int u_max, int g_max,
                   const float *X, int64_t ldx, float *Y, int64_t ldy, int64_t n_vecs, bool relu, cudaStream_t s)
{
    const int n_chunks = (G + GM - 1) / GM, gm_total = n_chunks * GM;           // G <= 16: one chunk of GM = G rounded to even
    const int64_t n_tiles = kn_cdiv(n_vecs, TN), gx = n_clusters * n_tiles;
    KN_REQUIRE(gx <= 0x7fffffffLL, "spmm_cg: grid too large");
    const size_t smem = (size_t)u_max * TN * sizeof(float) + (size_t)g_max * K_pad * sizeof(int32_t);
    KN_REQUIRE(smem <= (size_t)KN_CG_MAX_UNION * TN * sizeof(float), "spmm_cg: staged tile + index table of %zu bytes do not fit", smem);
    KN_ONCE_PER_DEVICE {
        KN_CUDA(cudaFuncSetAttribute(pg_cluster_kernel<GM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, KN_CG_MAX_UNION * TN * (int)sizeof(float)));
        KN_CUDA(cudaFuncSetAttribute(pg_cluster_kernel<GM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, KN_CG_MAX_UNION * TN * (int)sizeof(float)));
    }
    if (relu) pg_cluster_kernel<GM, true><<<(unsigned)gx, kThreads, smem, s>>>(cl_gptr, cl_uptr, ucols, rows, lidx, valsT, group_k, block_of, n_clusters, G, K_pad, u_max, n_chunks, gm_total, X, ldx, Y, ldy, n_vecs, kn_current_peers());
    else      pg_cluster_kernel<GM, false><<<(unsigned)gx, kThreads, smem, s>>>(cl_gptr, cl_uptr, ucols, rows, lidx, valsT, group_k, block_of, n_clusters, G, K_pad, u_max, n_chunks, gm_total, X, ldx, Y, ldy, n_vecs, kn_current_peers());
    KN_CHECK_LAUNCH();
    return KN_OK;
}
}  // namespace

KN_API int kn_spmm_cg_f32(const int32_t *cl_gptr, const int32_t *cl_uptr, const int32_t *ucols, const int32_t *rows, const int32_t *lidx, const float *valsT,
                          const int32_t *group_k, const int32_t *block_of, int64_t n_clusters, int32_t G, int32_t K_pad, int32_t u_max, int32_t g_max,
                          const float *X, int64_t ldx, float *Y, int64_t ldy, int64_t n_vecs, uint32_t flags, const kn_peers *peers_arg, void *stream) {
    KnPeersScope peers_scope(peers_arg);
    if (!peers_scope.ok) return KN_ERR_INVALID_ARGUMENT;
    KN_REQUIRE(n_clusters >= 0 && G > 0 && G <= 256 && K_pad > 0, "spmm_cg: bad shape (G=%d K_pad=%d)", G, K_pad);
    KN_REQUIRE(g_max > 0 && K_pad % 32 == 0, "spmm_cg: bad g_max / K_pad");
    KN_REQUIRE(u_max > 0 && u_max <= KN_CG_MAX_UNION, "spmm_cg: union of %d columns does not fit the staging tile (max %d)", u_max, KN_CG_MAX_UNION);
    KN_REQUIRE(n_vecs >= 0 && ldx >= n_vecs && ldy >= n_vecs, "spmm_cg: bad leading dimension");
    if (n_clusters == 0 || n_vecs == 0) return KN_OK;
    KN_REQUIRE(cl_gptr && cl_uptr && ucols && rows && lidx && valsT && X && Y, "spmm_cg: null pointer");
    KN_REQUIRE(n_vecs % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && (((uintptr_t)X | (uintptr_t)Y | (uintptr_t)valsT) & 15) == 0,
               "spmm_cg: n_vecs, ldx, ldy must be multiples of 4 and X, Y, valsT 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const bool relu = (flags & KN_SPMM_RELU) != 0;
#define KN_CG(GM) return launch_cluster<GM>(cl_gptr, cl_uptr, ucols, rows, lidx, valsT, group_k, block_of, n_clusters, G, K_pad, u_max, g_max, X, ldx, Y, ldy, n_vecs, relu, s)
    if (G > 16) KN_CG(16);            // chunks of 16 rows: valsT[block][K_pad][G rounded up to 16]
    switch ((G + 1) / 2) {            // GM = G rounded up to even: valsT[block][K_pad][GM]
        case 1: KN_CG(2);
        case 2: KN_CG(4);
        case 3: KN_CG(6);
        case 4: KN_CG(8);
        case 5: KN_CG(10);
        case 6: KN_CG(12);
        case 7: KN_CG(14);
        default: KN_CG(16);
    }
#undef KN_CG
}
