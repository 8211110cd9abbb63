// keyedconv.cu -- fused key compile of a conv / linear layer under monomial keys: W_hat = A . toeplitz(conv) . Ainv written
// straight from the filter weights, ONE pass, ONE column sort per output pixel.
//
// The reference builds the Toeplitz matrix (keynet/sparse.py:122-203) and folds the keys in with two scipy SpGEMMs
// (keynet/layer.py:35,59,70).  With one entry per key row (permutation / gain keys and their products) that is
//     W_hat[r, col_map[c]] = fl32(fl32(a_r * W[s, c]) * ainv_c),   s = A.perm[r],
// exact zeros dropped, columns ascending (canonical CSR).  All M rows of one output pixel (one per output channel) read
// the SAME input taps, so they share one column set and therefore one sorted order: the CTA that owns a pixel maps its
// C*np*nq (+1 bias) columns through col_map once, sorts (new column, tap) pairs once in shared memory, and then its
// warps stream the M rows out -- coalesced 128 B stores of indices and values, zeros squeezed out with a ballot.
// Nothing but the finished CSR ever touches HBM: the un-keyed Toeplitz matrix (another 8 B per entry written and read
// back) and the per-row sorts of the two-kernel path (csrc/toeplitz.cu + csrc/keycompile.cu) are gone.
// Algorithmic bytes: nnz * 8 (indices + data written once) + (R+1) * 8 (row pointers); the weights (<= 9.4 MB) and the key
// vectors stay in L2.
//
// The same geometry also yields the PATTERN-GROUP execution format directly (kn_conv2d_groups_*): rows / column list /
// value block of every output pixel without going through a CSR at all -- for VGG16 (15 G stored entries = 120 GB as CSR)
// the whole keyed network is then built from < 1 GB of writes.  Under permutation-only keys the value block of a pixel
// depends only on which taps are in bounds (interior, edges, corners), so blocks are shared by construction: the
// reference's "unique tiles" (keynet/sparse.py:553-568,690-779) without hashing anything.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kSortCap = 8192;      // taps per output pixel held in shared memory (3 x 32 KB: column, tap, column scale)

struct PixGeom {
    int u, v;            // input coordinate of the window centre
    int p0, np, q0, nq;  // first valid tap offset and number of valid taps per axis
};

__device__ __forceinline__ PixGeom pix_geom(const kn_conv2d_desc &d, int pix) {
    PixGeom g;
    const int Vo = d.V / d.stride;
    const int ku = pix / Vo, kv = pix - ku * Vo;
    g.u = ku * d.stride; g.v = kv * d.stride;
    const int ph = (d.P - 1) / 2, qh = (d.Q - 1) / 2;
    const int plo = max(-ph, -g.u), phi = min(ph, d.U - 1 - g.u);
    const int qlo = max(-qh, -g.v), qhi = min(qh, d.V - 1 - g.v);
    g.p0 = plo; g.np = max(0, phi - plo + 1);
    g.q0 = qlo; g.nq = max(0, qhi - qlo + 1);
    return g;
}

// tap e of a pixel, in the Toeplitz row's own order (channel, kernel row, kernel column == ascending source column):
// source column and offset of the weight inside one output channel's [C][P][Q] slab
__device__ __forceinline__ void tap_of(const kn_conv2d_desc &d, const PixGeom &g, int e, int32_t &col_src, int32_t &widx) {
    const int taps = g.np * g.nq;
    const int c = e / taps, t = e - c * taps;
    const int ip = t / g.nq, iq = t - ip * g.nq;
    const int p = g.p0 + ip, q = g.q0 + iq;
    col_src = c * d.U * d.V + (g.u + p) * d.V + (g.v + q);
    widx = (c * d.P + (p + (d.P - 1) / 2)) * d.Q + (q + (d.Q - 1) / 2);
}

__device__ __forceinline__ float keyed(float w, float a, float ai, bool has_a, bool has_ai) {
    float t = w;
    if (has_a) t = __fmul_rn(a, t);          // left product first, exactly like A.dot(W).dot(Ainv)
    if (has_ai) t = __fmul_rn(t, ai);
    return t;
}

// ---- count: stored entries of every compiled row (exact zeros dropped unless keep_zeros) ------------------------------
__global__ void __launch_bounds__(kThreads)
keyed_conv_count_kernel(kn_conv2d_desc d, const float *__restrict__ weight, const float *__restrict__ bias,
                        const int32_t *__restrict__ pix, int64_t n_groups, const int32_t *__restrict__ row_of_src,
                        const float *__restrict__ row_scale, const float *__restrict__ col_scale, int keep_zeros,
                        int64_t *__restrict__ row_nnz)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int UoVo = (d.U / d.stride) * (d.V / d.stride);
    const int64_t R_src = (int64_t)d.M * UoVo;
    const int K_src = d.C * d.U * d.V;
    const int CPQ = d.C * d.P * d.Q;
    const int64_t n_rows = n_groups * d.M;
    for (int64_t i = (int64_t)blockIdx.x * kWarps + warp; i < n_rows; i += (int64_t)gridDim.x * kWarps) {
        const int64_t g = i / d.M;
        const int m = (int)(i - g * d.M);
        const int px = pix ? pix[g] : (int)g;
        const int64_t s = (int64_t)m * UoVo + px;
        const int64_t r = row_of_src ? row_of_src[s] : s;
        if (r < 0) continue;                                 // warp-uniform
        const PixGeom geo = pix_geom(d, px);
        const int K_main = d.C * geo.np * geo.nq;
        const float a = row_scale ? row_scale[r] : 1.0f;
        int cnt = 0;
        if (keep_zeros) cnt = (lane == 0) ? K_main + (d.has_bias ? 1 : 0) : 0;
        else {
            for (int e = lane; e < K_main; e += 32) {
                int32_t cs, wi;
                tap_of(d, geo, e, cs, wi);
                const float v = keyed(__ldg(weight + (int64_t)m * CPQ + wi), a, col_scale ? __ldg(col_scale + cs) : 1.0f, row_scale != nullptr, col_scale != nullptr);
                cnt += (v != 0.0f) ? 1 : 0;
            }
            if (d.has_bias && lane == 0)
                cnt += (keyed(__ldg(bias + m), a, col_scale ? __ldg(col_scale + K_src) : 1.0f, row_scale != nullptr, col_scale != nullptr) != 0.0f) ? 1 : 0;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        if (lane == 0) row_nnz[r] = cnt;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {               // homogeneous row e_last
        const int64_t r = row_of_src ? row_of_src[R_src] : R_src;
        if (r >= 0) {
            const float v = keyed(1.0f, row_scale ? row_scale[r] : 1.0f, col_scale ? col_scale[K_src] : 1.0f, row_scale != nullptr, col_scale != nullptr);
            row_nnz[r] = (keep_zeros || v != 0.0f) ? 1 : 0;
        }
    }
}

// ---- normalised bitonic sort of (key, payload) pairs in shared memory, ascending keys, any n ---------------------------
__device__ __forceinline__ void bitonic_sort_kp(int32_t *keys, int32_t *pay, int n) {
    if (n < 2) return;
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int k = 2; k <= np2; k <<= 1) {
        for (int i = threadIdx.x; i < np2 / 2; i += blockDim.x) {
            const int blk = i / (k / 2), off = i - blk * (k / 2);
            const int a = blk * k + off, b = blk * k + (k - 1 - off);
            if (b < n) {
                const int32_t ka = keys[a], kb = keys[b];
                if (ka > kb) { keys[a] = kb; keys[b] = ka; const int32_t t = pay[a]; pay[a] = pay[b]; pay[b] = t; }
            }
        }
        __syncthreads();
        for (int j = k / 4; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np2 / 2; i += blockDim.x) {
                const int a = 2 * j * (i / j) + (i % j), b = a + j;
                if (b < n) {
                    const int32_t ka = keys[a], kb = keys[b];
                    if (ka > kb) { keys[a] = kb; keys[b] = ka; const int32_t t = pay[a]; pay[a] = pay[b]; pay[b] = t; }
                }
            }
            __syncthreads();
        }
    }
}

// ---- fill: one CTA per output pixel ------------------------------------------------------------------------------------
// BIG = false: the pixel's (column, tap) list is built and sorted in shared memory (up to kSortCap taps).
// BIG = true : a single pixel with more taps than that (a dense linear layer: 25 089 columns for VGG16 fc6): the list lives in a
//              global scratch buffer, is sorted once by CTA 0 of a preceding launch (sort_only), and every CTA then streams a
//              slice of the M rows from it.
template <bool BIG>
__global__ void __launch_bounds__(kThreads)
keyed_conv_fill_kernel(kn_conv2d_desc d, const float *__restrict__ weight, const float *__restrict__ bias,
                       const int32_t *__restrict__ pix, int64_t n_groups, const int32_t *__restrict__ row_of_src,
                       const int32_t *__restrict__ col_map, const float *__restrict__ row_scale, const float *__restrict__ col_scale,
                       int keep_zeros, const int64_t *__restrict__ out_indptr, int32_t *__restrict__ out_indices, float *__restrict__ out_data,
                       int32_t *scratch, int cap, int sort_only)
{
    extern __shared__ int32_t smem[];
    int32_t *s_key = BIG ? scratch : smem;                   // new column of every tap (sorted ascending)
    int32_t *s_tap = BIG ? scratch + cap : smem + kSortCap;  // weight offset of the tap inside a channel slab; -1 = bias entry
    float *s_cs = reinterpret_cast<float *>(BIG ? scratch + 2 * cap : smem + 2 * kSortCap);       // column scale of the tap (only with col_scale)
    __shared__ int s_unsorted;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int UoVo = (d.U / d.stride) * (d.V / d.stride);
    const int64_t R_src = (int64_t)d.M * UoVo;
    const int K_src = d.C * d.U * d.V;
    const int CPQ = d.C * d.P * d.Q;

    // pixels are dealt out in CONTIGUOUS blocks: the rows of one output channel for consecutive pixels are consecutive rows of
    // the CSR, so every warp's output stream stays inside a few pages (pixel-interleaved CTAs touched 2 * M far-apart pages per
    // pixel -- rows of different channels lie hundreds of MB apart -- and ran at 0.4 TB/s, bound by address translation)
    const int64_t per_cta = BIG ? n_groups : (n_groups + gridDim.x - 1) / gridDim.x;
    const int64_t g_begin = BIG ? 0 : (int64_t)blockIdx.x * per_cta;
    const int64_t g_end = BIG ? n_groups : (g_begin + per_cta < n_groups ? g_begin + per_cta : n_groups);
    for (int64_t g = g_begin; g < g_end; g++) {
        const int px = pix ? pix[g] : (int)g;
        const PixGeom geo = pix_geom(d, px);
        const int K_main = d.C * geo.np * geo.nq;
        const int K = K_main + (d.has_bias ? 1 : 0);
        if (!BIG || sort_only) {
        if (threadIdx.x == 0) s_unsorted = 0;
        __syncthreads();
        int unsorted = 0;
        for (int e = threadIdx.x; e < K; e += kThreads) {
            int32_t cs, wi;
            if (e < K_main) tap_of(d, geo, e, cs, wi);
            else { cs = K_src; wi = -1; }
            s_key[e] = col_map ? __ldg(col_map + cs) : cs;
            s_tap[e] = col_scale ? e : wi;       // with column scales the payload is the tap number (scale and offset looked up after the sort)
        }
        __syncthreads();
        for (int e = threadIdx.x + 1; e < K; e += kThreads) unsorted |= (s_key[e - 1] > s_key[e]) ? 1 : 0;
        if (unsorted) s_unsorted = 1;
        __syncthreads();
        if (s_unsorted) bitonic_sort_kp(s_key, s_tap, K);     // identity / monotone column maps skip the sort
        if (col_scale) {
            // payload = tap number: turn it into (weight offset, column scale) in sorted order
            for (int i = threadIdx.x; i < K; i += kThreads) {
                const int e = s_tap[i];
                int32_t cs, wi;
                if (e < K_main) tap_of(d, geo, e, cs, wi);
                else { cs = K_src; wi = -1; }
                s_cs[i] = __ldg(col_scale + cs);
                s_tap[i] = wi;
            }
        }
        __syncthreads();
        }
        if (BIG && sort_only) return;                           // (single pixel: every thread of the one CTA leaves together)
        // stream the M rows of this pixel: one warp per row, 32 sorted taps per step (BIG: the rows are spread over the grid)
        for (int m = BIG ? (int)blockIdx.x * kWarps + warp : warp; m < d.M; m += BIG ? (int)gridDim.x * kWarps : kWarps) {
            const int64_t s = (int64_t)m * UoVo + px;
            const int64_t r = row_of_src ? row_of_src[s] : s;
            if (r < 0) continue;
            const float a = row_scale ? row_scale[r] : 1.0f;
            const float *__restrict__ wm = weight + (int64_t)m * CPQ;
            const float bm = d.has_bias ? __ldg(bias + m) : 0.0f;
            int64_t out = out_indptr[r];
#pragma unroll 4
            for (int i0 = 0; i0 < K; i0 += 32) {
                const int i = i0 + lane;
                float v = 0.0f;
                bool keep = false;
                int32_t key = 0;
                if (i < K) {
                    const int wi = s_tap[i];
                    key = s_key[i];
                    v = keyed(wi >= 0 ? __ldg(wm + wi) : bm, a, col_scale ? s_cs[i] : 1.0f, row_scale != nullptr, col_scale != nullptr);
                    keep = keep_zeros || v != 0.0f;
                }
                const unsigned mask = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const int64_t pos = out + __popc(mask & ((1u << lane) - 1u));
                    out_indices[pos] = key;
                    out_data[pos] = v;
                }
                out += __popc(mask);
            }
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {               // homogeneous row e_last
        const int64_t r = row_of_src ? row_of_src[R_src] : R_src;
        if (r >= 0 && out_indptr[r + 1] > out_indptr[r]) {
            out_indices[out_indptr[r]] = col_map ? col_map[K_src] : K_src;
            out_data[out_indptr[r]] = keyed(1.0f, row_scale ? row_scale[r] : 1.0f, col_scale ? col_scale[K_src] : 1.0f, row_scale != nullptr, col_scale != nullptr);
        }
    }
}

// ---- fill, row-order variant ------------------------------------------------------------------------------------------------
// The per-pixel kernel above writes the M rows of a pixel, which lie M * Uo*Vo rows apart in the CSR: its 2*M output streams
// per CTA jump between pages hundreds of MB apart (~1 TB/s).  Here the rows are written IN ROW ORDER -- one warp per row,
// consecutive warps on consecutive rows, one sequential output stream per array for the whole layer -- and the sorted
// (column, tap) list of the row's pixel comes from a table:
//   * a column map (permuted input key): kn lists kernel = the per-pixel sort above, run once per pixel into a scratch buffer
//     [pixel][K] that stays L2-resident while the M channel rows of the layer pass over it;
//   * no column map (identity input key: every conv that follows a keyed ReLU): column = pixel base + relative column of the
//     tap, so ONE list per border class (which taps are in bounds: <= 9 classes for a 3x3 kernel) serves every pixel.
// The bias entry (homogeneous column: always the largest column) is appended by the row kernel.
__global__ void __launch_bounds__(kThreads)
keyed_conv_lists_kernel(kn_conv2d_desc d, const int32_t *__restrict__ pix, int64_t n_groups, const int32_t *__restrict__ col_map,
                        const float *__restrict__ col_scale, int K2, int2 *__restrict__ lists, float *__restrict__ list_scale)
{
    extern __shared__ int32_t smem[];
    int32_t *s_key = smem;
    int32_t *s_tap = smem + kSortCap;
    __shared__ int s_unsorted;
    for (int64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const int px = pix ? pix[g] : (int)g;
        const PixGeom geo = pix_geom(d, px);
        const int K = d.C * geo.np * geo.nq;
        if (threadIdx.x == 0) s_unsorted = 0;
        __syncthreads();
        for (int e = threadIdx.x; e < K; e += kThreads) {
            int32_t cs, wi;
            tap_of(d, geo, e, cs, wi);
            s_key[e] = col_map ? __ldg(col_map + cs) : cs;
            s_tap[e] = e;
        }
        __syncthreads();
        int unsorted = 0;
        for (int e = threadIdx.x + 1; e < K; e += kThreads) unsorted |= (s_key[e - 1] > s_key[e]) ? 1 : 0;
        if (unsorted) s_unsorted = 1;
        __syncthreads();
        if (s_unsorted) bitonic_sort_kp(s_key, s_tap, K);
        __syncthreads();
        for (int i = threadIdx.x; i < K; i += kThreads) {
            int32_t cs, wi;
            tap_of(d, geo, s_tap[i], cs, wi);
            lists[g * (int64_t)K2 + i] = make_int2(s_key[i], wi);
            if (list_scale) list_scale[g * (int64_t)K2 + i] = __ldg(col_scale + cs);
        }
        __syncthreads();
    }
}

// lists of the border classes (identity input key): class (p0, np, q0, nq) = which taps are in bounds; entries are
// (column relative to the pixel's base column u*V + v, weight offset), ascending by construction.  Class index =
// ((p0+ph)*P + np-1) * Q*Q + (q0+qh)*Q + nq-1.
__device__ __forceinline__ int class_index(const kn_conv2d_desc &d, const PixGeom &g) {
    return (((g.p0 + (d.P - 1) / 2) * d.P + (g.np - 1)) * d.Q + (g.q0 + (d.Q - 1) / 2)) * d.Q + (g.nq - 1);
}

__global__ void __launch_bounds__(kThreads)
conv_class_lists_kernel(kn_conv2d_desc d, int K2, int2 *__restrict__ lists)
{
    const int ph = (d.P - 1) / 2, qh = (d.Q - 1) / 2;
    const int cls = blockIdx.x;
    const int inq = cls % d.Q, iq0 = (cls / d.Q) % d.Q, inp = (cls / (d.Q * d.Q)) % d.P, ip0 = cls / (d.Q * d.Q * d.P);
    const int p0 = ip0 - ph, np = inp + 1, q0 = iq0 - qh, nq = inq + 1;
    if (p0 + np - 1 > ph || q0 + nq - 1 > qh) return;            // not a window of the kernel
    const int taps = np * nq, K = d.C * taps;
    for (int e = threadIdx.x; e < K; e += kThreads) {
        const int c = e / taps, t = e - c * taps;
        const int ip = t / nq, iq = t - ip * nq;
        const int p = p0 + ip, q = q0 + iq;
        lists[cls * (int64_t)K2 + e] = make_int2(c * d.U * d.V + p * d.V + q, (c * d.P + (p + ph)) * d.Q + (q + qh));
    }
}

// zero weights among the in-bounds taps of every (output channel, border class): with permutation-only keys the stored entries
// of a row are its taps minus these (the exact zeros scipy's SpGEMM drops), so the count pass needs no per-entry work
__global__ void __launch_bounds__(kThreads)
conv_class_zeros_kernel(kn_conv2d_desc d, const float *__restrict__ weight, const float *__restrict__ bias, int n_cls, int32_t *__restrict__ zeros)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ph = (d.P - 1) / 2, qh = (d.Q - 1) / 2;
    const int CPQ = d.C * d.P * d.Q;
    for (int64_t i = (int64_t)blockIdx.x * kWarps + warp; i < (int64_t)d.M * n_cls; i += (int64_t)gridDim.x * kWarps) {
        const int m = (int)(i / n_cls), cls = (int)(i - (int64_t)m * n_cls);
        const int inq = cls % d.Q, iq0 = (cls / d.Q) % d.Q, inp = (cls / (d.Q * d.Q)) % d.P, ip0 = cls / (d.Q * d.Q * d.P);
        const int p0 = ip0 - ph, np = inp + 1, q0 = iq0 - qh, nq = inq + 1;
        int cnt = 0;
        if (p0 + np - 1 <= ph && q0 + nq - 1 <= qh) {
            const int taps = np * nq, K = d.C * taps;
            for (int e = lane; e < K; e += 32) {
                const int c = e / taps, t = e - c * taps;
                const int ip = t / nq, iq = t - ip * nq;
                cnt += (__ldg(weight + (int64_t)m * CPQ + (c * d.P + (p0 + ip + ph)) * d.Q + (q0 + iq + qh)) == 0.0f) ? 1 : 0;
            }
            if (lane == 0 && d.has_bias) cnt += (__ldg(bias + m) == 0.0f) ? 1 : 0;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        if (lane == 0) zeros[i] = cnt;
    }
}

__global__ void __launch_bounds__(kThreads)
keyed_conv_count_class_kernel(kn_conv2d_desc d, const int32_t *__restrict__ pix, int64_t n_groups, const int32_t *__restrict__ row_of_src,
                              int n_cls, const int32_t *__restrict__ zeros, int keep_zeros, int64_t *__restrict__ row_nnz)
{
    const int UoVo = (d.U / d.stride) * (d.V / d.stride);
    const int64_t n_items = (int64_t)d.M * n_groups;
    for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(it / n_groups);
        const int64_t g = it - (int64_t)m * n_groups;
        const int px = pix ? pix[g] : (int)g;
        const int64_t s = (int64_t)m * UoVo + px;
        const int64_t r = row_of_src ? row_of_src[s] : s;
        if (r < 0) continue;
        const PixGeom geo = pix_geom(d, px);
        const int K = d.C * geo.np * geo.nq + (d.has_bias ? 1 : 0);
        row_nnz[r] = K - (keep_zeros ? 0 : zeros[(int64_t)m * n_cls + class_index(d, geo)]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t R_src = (int64_t)d.M * UoVo;
        const int64_t r = row_of_src ? row_of_src[R_src] : R_src;
        if (r >= 0) row_nnz[r] = 1;
    }
}

__global__ void __launch_bounds__(kThreads)
keyed_conv_rows_kernel(kn_conv2d_desc d, const float *__restrict__ weight, const float *__restrict__ bias,
                       const int32_t *__restrict__ pix, int64_t n_groups, const int32_t *__restrict__ row_of_src,
                       const float *__restrict__ row_scale, int keep_zeros,
                       const int2 *__restrict__ lists, const float *__restrict__ list_scale, int class_mode,
                       const int32_t *__restrict__ col_map, const float *__restrict__ col_scale, int K2,
                       const int64_t *__restrict__ out_indptr, int32_t *__restrict__ out_indices, float *__restrict__ out_data)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int UoVo = (d.U / d.stride) * (d.V / d.stride);
    const int64_t R_src = (int64_t)d.M * UoVo;
    const int CPQ = d.C * d.P * d.Q;
    const int K_src = d.C * d.U * d.V;
    const int32_t key_last = col_map ? __ldg(col_map + K_src) : K_src;          // homogeneous column: always the largest
    const float scale_last = col_scale ? __ldg(col_scale + K_src) : 1.0f;
    const int has_col_scale = col_scale != nullptr;
    const int64_t n_items = (int64_t)d.M * n_groups;
    // consecutive warps take consecutive rows (channel-major, pixels of the list in order): sequential output streams
    for (int64_t it = (int64_t)blockIdx.x * kWarps + warp; it < n_items; it += (int64_t)gridDim.x * kWarps) {
        const int m = (int)(it / n_groups);
        const int64_t g = it - (int64_t)m * n_groups;
        const int px = pix ? pix[g] : (int)g;
        const int64_t s = (int64_t)m * UoVo + px;
        const int64_t r = row_of_src ? row_of_src[s] : s;
        if (r < 0) continue;
        const PixGeom geo = pix_geom(d, px);
        const int K = d.C * geo.np * geo.nq;
        const int64_t lg = class_mode ? class_index(d, geo) : g;
        const int2 *__restrict__ lst = lists + lg * (int64_t)K2;
        const float *__restrict__ lsc = list_scale ? list_scale + lg * (int64_t)K2 : nullptr;
        const int32_t kb = class_mode ? geo.u * d.V + geo.v : 0;
        const float a = row_scale ? row_scale[r] : 1.0f;
        const float *__restrict__ wm = weight + (int64_t)m * CPQ;
        int64_t out = out_indptr[r];
#pragma unroll 4
        for (int i0 = 0; i0 < K; i0 += 32) {
            const int i = i0 + lane;
            float v = 0.0f;
            bool keep = false;
            int32_t key = 0;
            if (i < K) {
                const int2 e = lst[i];
                key = e.x + kb;
                v = keyed(__ldg(wm + e.y), a, lsc ? lsc[i] : 1.0f, row_scale != nullptr, has_col_scale != 0);
                keep = keep_zeros || v != 0.0f;
            }
            const unsigned mask = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int64_t pos = out + __popc(mask & ((1u << lane) - 1u));
                out_indices[pos] = key;
                out_data[pos] = v;
            }
            out += __popc(mask);
        }
        if (d.has_bias && lane == 0) {
            const float v = keyed(__ldg(bias + m), a, scale_last, row_scale != nullptr, has_col_scale != 0);
            if (keep_zeros || v != 0.0f) { out_indices[out] = key_last; out_data[out] = v; }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {               // homogeneous row e_last
        const int64_t r = row_of_src ? row_of_src[R_src] : R_src;
        if (r >= 0 && out_indptr[r + 1] > out_indptr[r]) {
            out_indices[out_indptr[r]] = key_last;
            out_data[out_indptr[r]] = keyed(1.0f, row_scale ? row_scale[r] : 1.0f, scale_last, row_scale != nullptr, has_col_scale != 0);
        }
    }
}

// ---- pattern groups straight from the geometry -------------------------------------------------------------------------
// rows[g][M] (compiled row of every output channel of pixel g), cols[g][K_pad] (new column of every tap, Toeplitz order,
// padding repeats the first column), group_k[g]
__global__ void __launch_bounds__(kThreads)
conv_groups_index_kernel(kn_conv2d_desc d, const int32_t *__restrict__ pix, int64_t n_groups, const int32_t *__restrict__ row_of_src,
                         const int32_t *__restrict__ col_map, int K_pad, int32_t *__restrict__ rows, int32_t *__restrict__ cols, int32_t *__restrict__ group_k)
{
    const int UoVo = (d.U / d.stride) * (d.V / d.stride);
    const int K_src = d.C * d.U * d.V;
    for (int64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const int px = pix ? pix[g] : (int)g;
        const PixGeom geo = pix_geom(d, px);
        const int K_main = d.C * geo.np * geo.nq;
        const int K = K_main + (d.has_bias ? 1 : 0);
        int32_t first;
        {
            int32_t cs, wi;
            if (K_main > 0) tap_of(d, geo, 0, cs, wi); else cs = K_src;
            first = col_map ? __ldg(col_map + cs) : cs;
        }
        for (int e = threadIdx.x; e < K_pad; e += kThreads) {
            int32_t c = first;
            if (e < K) {
                int32_t cs, wi;
                if (e < K_main) tap_of(d, geo, e, cs, wi); else cs = K_src;
                c = col_map ? __ldg(col_map + cs) : cs;
            }
            cols[g * (int64_t)K_pad + e] = c;
        }
        for (int m = threadIdx.x; m < d.M; m += kThreads) {
            const int64_t s = (int64_t)m * UoVo + px;
            rows[g * (int64_t)d.M + m] = row_of_src ? row_of_src[s] : (int32_t)s;
        }
        if (threadIdx.x == 0) group_k[g] = K;
    }
}

// value blocks vals[b][M][K_pad] of the pixels block_pix[b]: one block per pixel when the keys carry gains
// (scaled != 0: row / column scales applied like the compile), one per border class otherwise
__global__ void __launch_bounds__(kThreads)
conv_groups_values_kernel(kn_conv2d_desc d, const float *__restrict__ weight, const float *__restrict__ bias,
                          const int32_t *__restrict__ block_pix, int64_t n_blocks, const int32_t *__restrict__ row_of_src,
                          const float *__restrict__ row_scale, const float *__restrict__ col_scale, int K_pad, float *__restrict__ vals)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int UoVo = (d.U / d.stride) * (d.V / d.stride);
    const int K_src = d.C * d.U * d.V;
    const int CPQ = d.C * d.P * d.Q;
    const int64_t n_rows = n_blocks * d.M;
    for (int64_t i = (int64_t)blockIdx.x * kWarps + warp; i < n_rows; i += (int64_t)gridDim.x * kWarps) {
        const int64_t b = i / d.M;
        const int m = (int)(i - b * d.M);
        const int px = block_pix[b];
        const PixGeom geo = pix_geom(d, px);
        const int K_main = d.C * geo.np * geo.nq;
        const int K = K_main + (d.has_bias ? 1 : 0);
        float a = 1.0f;
        if (row_scale) {
            const int64_t s = (int64_t)m * UoVo + px;
            const int64_t r = row_of_src ? row_of_src[s] : s;
            a = (r >= 0) ? row_scale[r] : 0.0f;
        }
        float *__restrict__ v = vals + i * (int64_t)K_pad;
        for (int e = lane; e < K_pad; e += 32) {
            float x = 0.0f;
            if (e < K) {
                int32_t cs, wi;
                if (e < K_main) tap_of(d, geo, e, cs, wi); else { cs = K_src; wi = -1; }
                x = keyed(wi >= 0 ? __ldg(weight + (int64_t)m * CPQ + wi) : __ldg(bias + m), a, col_scale ? __ldg(col_scale + cs) : 1.0f, row_scale != nullptr, col_scale != nullptr);
            }
            v[e] = x;
        }
    }
}

int check_desc(const kn_conv2d_desc *d) {
    KN_REQUIRE(d != nullptr, "keyed_conv: null descriptor");
    KN_REQUIRE(d->C > 0 && d->U > 0 && d->V > 0 && d->M > 0, "keyed_conv: non-positive shape");
    KN_REQUIRE(d->P > 0 && d->Q > 0 && (d->P % 2) == 1 && (d->Q % 2) == 1, "keyed_conv: kernel must be odd (P=%d Q=%d)", d->P, d->Q);
    KN_REQUIRE(d->stride > 0 && !d->depthwise, "keyed_conv: stride must be positive, depthwise layers use the CSR path");
    KN_REQUIRE((int64_t)d->C * d->U * d->V < 0x7fffffffLL && (int64_t)d->M * (d->U / d->stride) * (d->V / d->stride) < 0x7fffffffLL, "keyed_conv: index exceeds int32");
    return KN_OK;
}

int grid_for(int64_t items, int per_cta, int ctas_per_sm) {
    const int64_t want = kn_cdiv(items, per_cta);
    const int64_t cap = (int64_t)kn_sm_count() * ctas_per_sm;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}
}  // namespace

KN_API int kn_keyed_conv2d_count(const kn_conv2d_desc *desc, const float *weight, const float *bias, const int32_t *pix, int64_t n_groups,
                                 const int32_t *row_of_src, const float *row_scale, const float *col_scale, int32_t keep_zeros,
                                 int64_t *row_nnz, void *stream) {
    int rc = check_desc(desc); if (rc) return rc;
    KN_REQUIRE(n_groups >= 0, "keyed_conv: negative group count");
    KN_REQUIRE(weight && row_nnz, "keyed_conv: null pointer");
    KN_REQUIRE(!desc->has_bias || bias, "keyed_conv: has_bias set but bias is null");
    cudaStream_t s = (cudaStream_t)stream;
    if (!row_scale && !col_scale && n_groups > 64) {
        // permutation-only keys: a row stores its taps minus the zero weights among them, a function of (channel, border class)
        const int n_cls = desc->P * desc->P * desc->Q * desc->Q;
        int32_t *zeros = nullptr;
        KN_CUDA(cudaMallocAsync((void **)&zeros, sizeof(int32_t) * (size_t)desc->M * n_cls, s));
        conv_class_zeros_kernel<<<grid_for((int64_t)desc->M * n_cls, kWarps, 16), kThreads, 0, s>>>(*desc, weight, bias, n_cls, zeros);
        keyed_conv_count_class_kernel<<<grid_for((int64_t)desc->M * n_groups, kThreads, 16), kThreads, 0, s>>>(*desc, pix, n_groups, row_of_src, n_cls, zeros, keep_zeros, row_nnz);
        const cudaError_t e = cudaGetLastError();
        KN_CUDA(cudaFreeAsync(zeros, s));
        if (e != cudaSuccess) { kn_set_error("keyed_conv: launch failed: %s", cudaGetErrorString(e)); return KN_ERR_CUDA; }
        return KN_OK;
    }
    keyed_conv_count_kernel<<<grid_for(n_groups * desc->M, kWarps, 16), kThreads, 0, s>>>(*desc, weight, bias, pix, n_groups, row_of_src, row_scale, col_scale, keep_zeros, row_nnz);
    KN_CHECK_LAUNCH();
    return KN_OK;
}

KN_API int kn_keyed_conv2d_fill(const kn_conv2d_desc *desc, const float *weight, const float *bias, const int32_t *pix, int64_t n_groups,
                                const int32_t *row_of_src, const int32_t *col_map, const float *row_scale, const float *col_scale, int32_t keep_zeros,
                                const int64_t *out_indptr, int32_t *out_indices, float *out_data, void *stream) {
    int rc = check_desc(desc); if (rc) return rc;
    KN_REQUIRE(n_groups >= 0, "keyed_conv: negative group count");
    KN_REQUIRE(weight && out_indptr && out_indices && out_data, "keyed_conv: null pointer");
    KN_REQUIRE(!desc->has_bias || bias, "keyed_conv: has_bias set but bias is null");
    const int64_t K_max = (int64_t)desc->C * desc->P * desc->Q + (desc->has_bias ? 1 : 0);
    cudaStream_t s = (cudaStream_t)stream;
    if (K_max > kSortCap) {
        // more taps per pixel than the shared-memory sort holds: supported for ONE pixel (dense linear layers), whose list is
        // sorted once in a stream-ordered scratch buffer and then shared by the whole grid
        if (n_groups != 1 || K_max > (1 << 24)) { kn_set_error("keyed_conv: %lld taps per output pixel exceed the shared-memory sort (%d): use the two-kernel path", (long long)K_max, kSortCap); return KN_ERR_UNSUPPORTED; }
        int32_t *scratch = nullptr;
        KN_CUDA(cudaMallocAsync((void **)&scratch, (size_t)K_max * 12, s));
        keyed_conv_fill_kernel<true><<<1, kThreads, 0, s>>>(*desc, weight, bias, pix, n_groups, row_of_src, col_map, row_scale, col_scale, keep_zeros,
                                                             out_indptr, out_indices, out_data, scratch, (int)K_max, 1);
        keyed_conv_fill_kernel<true><<<grid_for(desc->M, kWarps, 8), kThreads, 0, s>>>(*desc, weight, bias, pix, n_groups, row_of_src, col_map, row_scale, col_scale, keep_zeros,
                                                                                        out_indptr, out_indices, out_data, scratch, (int)K_max, 0);
        const cudaError_t e = cudaGetLastError();
        KN_CUDA(cudaFreeAsync(scratch, s));
        if (e != cudaSuccess) { kn_set_error("keyed_conv: launch failed: %s", cudaGetErrorString(e)); return KN_ERR_CUDA; }
        return KN_OK;
    }
    // ---- row-order path: a table of sorted (column, tap) lists + one warp per row, rows written in order
    {
        const int K2 = desc->C * desc->P * desc->Q;
        const bool per_pixel = (col_map != nullptr) || (col_scale != nullptr);
        int2 *lists = nullptr; float *list_scale = nullptr;
        const size_t smem = (size_t)kSortCap * 4 * 2;
        KN_ONCE_PER_DEVICE {
            KN_CUDA(cudaFuncSetAttribute(keyed_conv_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortCap * 4 * 2));
            // the scratch lists come from the stream-ordered pool: keep its memory between calls (the default threshold of 0
            // returns it to the driver at every synchronisation, i.e. hundreds of MB are re-created per layer)
            int dev = 0; cudaMemPool_t pool;
            if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                uint64_t thr = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            }
        }
        if (per_pixel) {
            // permuted / scaled input key: one sorted list per pixel, L2-resident while the M channel rows pass over it
            KN_CUDA(cudaMallocAsync((void **)&lists, (size_t)n_groups * K2 * sizeof(int2), s));
            if (col_scale) KN_CUDA(cudaMallocAsync((void **)&list_scale, (size_t)n_groups * K2 * sizeof(float), s));
            keyed_conv_lists_kernel<<<grid_for(n_groups, 1, 3), kThreads, smem, s>>>(*desc, pix, n_groups, col_map, col_scale, K2, lists, list_scale);
        } else {
            // identity input key: one list per border class
            const int n_cls = desc->P * desc->P * desc->Q * desc->Q;
            KN_CUDA(cudaMallocAsync((void **)&lists, (size_t)n_cls * K2 * sizeof(int2), s));
            conv_class_lists_kernel<<<n_cls, kThreads, 0, s>>>(*desc, K2, lists);
        }
        keyed_conv_rows_kernel<<<grid_for((int64_t)desc->M * n_groups, kWarps, 8), kThreads, 0, s>>>(*desc, weight, bias, pix, n_groups, row_of_src, row_scale, keep_zeros,
                                                                                                    lists, list_scale, per_pixel ? 0 : 1, col_map, col_scale, K2,
                                                                                                    out_indptr, out_indices, out_data);
        const cudaError_t e = cudaGetLastError();
        KN_CUDA(cudaFreeAsync(lists, s));
        if (list_scale) KN_CUDA(cudaFreeAsync(list_scale, s));
        if (e != cudaSuccess) { kn_set_error("keyed_conv: launch failed: %s", cudaGetErrorString(e)); return KN_ERR_CUDA; }
        return KN_OK;
    }
}

KN_API int kn_conv2d_groups_index(const kn_conv2d_desc *desc, const int32_t *pix, int64_t n_groups, const int32_t *row_of_src, const int32_t *col_map,
                                  int32_t K_pad, int32_t *rows, int32_t *cols, int32_t *group_k, void *stream) {
    int rc = check_desc(desc); if (rc) return rc;
    KN_REQUIRE(n_groups >= 0 && K_pad > 0, "conv_groups: bad shape");
    if (n_groups == 0) return KN_OK;
    KN_REQUIRE(rows && cols && group_k, "conv_groups: null pointer");
    KN_REQUIRE((int64_t)desc->C * desc->P * desc->Q + (desc->has_bias ? 1 : 0) <= K_pad, "conv_groups: K_pad smaller than the tap count");
    conv_groups_index_kernel<<<grid_for(n_groups, 1, 8), kThreads, 0, (cudaStream_t)stream>>>(*desc, pix, n_groups, row_of_src, col_map, K_pad, rows, cols, group_k);
    KN_CHECK_LAUNCH();
    return KN_OK;
}

KN_API int kn_conv2d_groups_values(const kn_conv2d_desc *desc, const float *weight, const float *bias, const int32_t *block_pix, int64_t n_blocks,
                                   const int32_t *row_of_src, const float *row_scale, const float *col_scale, int32_t K_pad, float *vals, void *stream) {
    int rc = check_desc(desc); if (rc) return rc;
    KN_REQUIRE(n_blocks >= 0 && K_pad > 0, "conv_groups: bad shape");
    if (n_blocks == 0) return KN_OK;
    KN_REQUIRE(weight && block_pix && vals, "conv_groups: null pointer");
    KN_REQUIRE(!desc->has_bias || bias, "conv_groups: has_bias set but bias is null");
    conv_groups_values_kernel<<<grid_for(n_blocks * desc->M, kWarps, 16), kThreads, 0, (cudaStream_t)stream>>>(*desc, weight, bias, block_pix, n_blocks, row_of_src, row_scale, col_scale, K_pad, vals);
    KN_CHECK_LAUNCH();
    return KN_OK;
}
