// keycompile.cu -- key compile W_hat = A . W . Ainv for monomial keys (kernel K2 of SURVEY.md 2.2).
//
// The reference runs two scipy SpGEMMs (keynet/layer.py:35,59,70).  When A and Ainv have one
// entry per row (permutation, diagonal gain, or their product) that is an integer reindexing
// plus two fp32 multiplies per stored value:
//     row gather by A          -> done upstream (row_ids of the Toeplitz / linear builders)
//     c' = col_map[c]          -> column of the single entry in row c of Ainv
//     v' = fl32(fl32(a_r * v) * ainv_c)        left product first, exactly like A.dot(W).dot(Ainv)
//     v' == 0 entries dropped  -> scipy's csr_matmat never stores an exact zero
//     columns sorted ascending -> canonical form (what `sort_indices()` gives on the reference)
// Multiplies use __fmul_rn (no contraction, denormals kept), so stored values are bit-identical
// to scipy's.
//
// One CTA compiles one row at a time: entries are scaled, compacted into shared memory and
// sorted by new column with a normalised bitonic network (all compare-exchanges ascending, so a
// non power-of-two length needs no padding writes).  Rows longer than the shared-memory budget
// are ranked with a column bitmap in shared memory (position = number of set bits below the new
// column; O(nnz + n_cols/32) per row, used for the dense fc rows), and if the matrix is too wide
// even for that, the bitonic network runs in place on the output row in global memory.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kSmemCap = 8192;   // entries per row held in shared memory (64 KB: key + value)
constexpr int kBitmapWords = kSmemCap;              // the same 64 KB viewed as bitmap words + word prefix
constexpr int64_t kBitmapMaxCols = (int64_t)kBitmapWords * 32;

__device__ __forceinline__ float keyed_value(float v, const float *row_scale, const float *col_scale, int64_t r, int32_t c) {
    float t = v;
    if (row_scale) t = __fmul_rn(row_scale[r], t);
    if (col_scale) t = __fmul_rn(t, col_scale[c]);
    return t;
}

__global__ void __launch_bounds__(kThreads)
keycompile_count_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices, const float *__restrict__ data,
                        int64_t n_rows, const float *__restrict__ row_scale, const float *__restrict__ col_scale,
                        int keep_zeros, int64_t *__restrict__ row_nnz)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = kThreads / 32;
    for (int64_t r = (int64_t)blockIdx.x * wpb + warp; r < n_rows; r += (int64_t)gridDim.x * wpb) {
        const int64_t beg = indptr[r], end = indptr[r + 1];
        int cnt = 0;
        for (int64_t e = beg + lane; e < end; e += 32)
            cnt += (keep_zeros || keyed_value(data[e], row_scale, col_scale, r, indices[e]) != 0.0f) ? 1 : 0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        if (lane == 0) row_nnz[r] = cnt;
    }
}

// ---- keys with a bias column (affine photometric keys [[D, b],[0, 1]], keynet/sparse.py:99-119) --------------------
// A = [[diag(a) P, ba],[0, 1]] on the left adds ba[r] to the row's last-column entry (W's last row is e_last);
// Ainv = [[diag(ai) Pi, bi],[0, 1]] on the right sends every entry t = fl(a_r * w) at column c < last to column
// col_map[c] with value fl(t * ai[c]) AND adds fl(t * bi[c]) to the last column.  The last-column value of a row is
// therefore  fl(a_r * w_last) + ba[r] + sum_c fl(t_c * bi[c])  -- a reduction over the row, evaluated by BOTH the
// count and the fill kernel with this one block-wide routine so they agree bit for bit on whether it is zero.
// (The summation order differs from scipy's traversal order: the last column matches the reference to fp32
// rounding, every other entry and all indices stay bit-exact.)
__device__ __forceinline__ float biased_last_value(const int32_t *__restrict__ indices, const float *__restrict__ data, int64_t beg, int64_t end,
                                                   int64_t r, int32_t last_in, const float *row_scale, const float *row_bias,
                                                   const float *col_bias, float *s_red /*[kThreads]*/)
{
    float part = 0.0f;
    for (int64_t e = beg + threadIdx.x; e < end; e += blockDim.x) {
        const int32_t c = indices[e];
        float t = data[e];
        if (row_scale) t = __fmul_rn(row_scale[r], t);
        if (c == last_in) part = __fadd_rn(part, t);
        else if (col_bias) part = __fadd_rn(part, __fmul_rn(t, col_bias[c]));
    }
    s_red[threadIdx.x] = part;
    __syncthreads();
    for (int off = kThreads / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) s_red[threadIdx.x] = __fadd_rn(s_red[threadIdx.x], s_red[threadIdx.x + off]);
        __syncthreads();
    }
    float v = s_red[0];
    __syncthreads();
    if (row_bias) v = __fadd_rn(v, row_bias[r]);
    return v;
}

__global__ void __launch_bounds__(kThreads)
keycompile_count_bias_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices, const float *__restrict__ data,
                             int64_t n_rows, const float *__restrict__ row_scale, const float *__restrict__ col_scale,
                             const float *__restrict__ row_bias, const float *__restrict__ col_bias, int32_t last_in,
                             int keep_zeros, int64_t *__restrict__ row_nnz)
{
    __shared__ float s_red[kThreads];
    __shared__ int s_cnt;
    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const int64_t beg = indptr[r], end = indptr[r + 1];
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        int cnt = 0;
        for (int64_t e = beg + threadIdx.x; e < end; e += blockDim.x) {
            const int32_t c = indices[e];
            if (c != last_in) cnt += (keep_zeros || keyed_value(data[e], row_scale, col_scale, r, c) != 0.0f) ? 1 : 0;
        }
        if (cnt) atomicAdd(&s_cnt, cnt);
        const float last = biased_last_value(indices, data, beg, end, r, last_in, row_scale, row_bias, col_bias, s_red);
        if (threadIdx.x == 0) row_nnz[r] = s_cnt + ((keep_zeros || last != 0.0f) ? 1 : 0);
        __syncthreads();
    }
}

// Rows with at most 32 entries (pooling layers: k*k taps; permutation / gain matrices: one entry): one WARP per row, the
// entries sorted by ranking -- every lane counts the kept entries with a smaller new column through 32 shuffles.  (One CTA
// per row, as below, spends 256 threads and a dozen barriers on 9 entries: VGG16's pooling layers have 1.5 M such rows.)
constexpr int kShortRow = 32;

__global__ void __launch_bounds__(kThreads)
keycompile_fill_short_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices, const float *__restrict__ data,
                             int64_t n_rows, const int32_t *__restrict__ col_map,
                             const float *__restrict__ row_scale, const float *__restrict__ col_scale, int keep_zeros,
                             const int64_t *__restrict__ out_indptr, int32_t *__restrict__ out_indices, float *__restrict__ out_data)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = kThreads / 32;
    for (int64_t r = (int64_t)blockIdx.x * wpb + warp; r < n_rows; r += (int64_t)gridDim.x * wpb) {
        const int64_t beg = indptr[r];
        const int n = (int)(indptr[r + 1] - beg);
        if (n > kShortRow || n == 0) continue;               // warp-uniform; long rows belong to the CTA kernel
        const int64_t obeg = out_indptr[r];
        int32_t cn = 0x7fffffff;
        float v = 0.0f;
        bool keep = false;
        if (lane < n) {
            const int32_t c = indices[beg + lane];
            v = keyed_value(data[beg + lane], row_scale, col_scale, r, c);
            keep = keep_zeros || v != 0.0f;
            cn = col_map ? col_map[c] : c;
        }
        int rank = 0;
#pragma unroll
        for (int j = 0; j < kShortRow; j++) {
            const int32_t cj = __shfl_sync(0xffffffffu, cn, j);
            const int kj = __shfl_sync(0xffffffffu, keep ? 1 : 0, j);
            rank += (kj && cj < cn) ? 1 : 0;
        }
        if (keep) { out_indices[obeg + rank] = cn; out_data[obeg + rank] = v; }
    }
}

__global__ void __launch_bounds__(kThreads)
keycompile_fill_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices, const float *__restrict__ data,
                       int64_t n_rows, int64_t n_cols, const int32_t *__restrict__ col_map,
                       const float *__restrict__ row_scale, const float *__restrict__ col_scale,
                       const float *__restrict__ row_bias, const float *__restrict__ col_bias, int32_t last_in, int keep_zeros,
                       const int64_t *__restrict__ out_indptr, int32_t *__restrict__ out_indices, float *__restrict__ out_data)
{
    extern __shared__ unsigned char smem_raw[];
    int32_t *s_key = reinterpret_cast<int32_t *>(smem_raw);
    float *s_val = reinterpret_cast<float *>(smem_raw + sizeof(int32_t) * kSmemCap);
    __shared__ int s_count;
    __shared__ float s_red[kThreads];
    const bool biased = (row_bias != nullptr) || (col_bias != nullptr);   // last_in = old last column, or -1 when no bias is involved

    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const int64_t beg = indptr[r], end = indptr[r + 1];
        const int64_t obeg = out_indptr[r];
        int64_t n_out = out_indptr[r + 1] - obeg;
        if (n_out == 0) continue;                               // block-uniform
        if (!biased && end - beg <= kShortRow) continue;        // written by keycompile_fill_short_kernel
        if (biased) {
            // the last-column entry is a row reduction; it is the largest column, so it goes straight to the row's end
            const float last = biased_last_value(indices, data, beg, end, r, last_in, row_scale, row_bias, col_bias, s_red);
            if (keep_zeros || last != 0.0f) {
                if (threadIdx.x == 0) { out_indices[obeg + n_out - 1] = (int32_t)(n_cols - 1); out_data[obeg + n_out - 1] = last; }
                n_out -= 1;
            }
            if (n_out == 0) continue;
        }
        const bool in_smem = n_out <= kSmemCap;
        if (!in_smem && n_cols <= kBitmapMaxCols) {
            // ---- long row, moderately wide matrix: rank entries through a column bitmap ----------
            unsigned *s_bits = reinterpret_cast<unsigned *>(s_key);
            int *s_pref = reinterpret_cast<int *>(s_val);
            const int n_words = (int)((n_cols + 31) >> 5);
            for (int w = threadIdx.x; w < n_words; w += blockDim.x) s_bits[w] = 0u;
            __syncthreads();
            for (int64_t e = beg + threadIdx.x; e < end; e += blockDim.x) {
                const int32_t c = indices[e];
                if (biased && c == last_in) continue;
                if (keep_zeros || keyed_value(data[e], row_scale, col_scale, r, c) != 0.0f) {
                    const int32_t cn = col_map ? col_map[c] : c;
                    atomicOr(&s_bits[cn >> 5], 1u << (cn & 31));
                }
            }
            __syncthreads();
            // exclusive prefix of per-word popcounts: sequential chunk per thread + scan of chunk sums
            const int chunk = (n_words + blockDim.x - 1) / blockDim.x;
            const int w0 = threadIdx.x * chunk, w1 = min(n_words, w0 + chunk);
            int local = 0;
            for (int w = w0; w < w1; w++) { s_pref[w] = local; local += __popc(s_bits[w]); }
            __shared__ int s_chunk[kThreads];
            s_chunk[threadIdx.x] = local;
            __syncthreads();
            if (threadIdx.x == 0) { int run = 0; for (int t = 0; t < kThreads; t++) { const int v = s_chunk[t]; s_chunk[t] = run; run += v; } }
            __syncthreads();
            const int off = s_chunk[threadIdx.x];
            for (int w = w0; w < w1; w++) s_pref[w] += off;
            __syncthreads();
            for (int64_t e = beg + threadIdx.x; e < end; e += blockDim.x) {
                const int32_t c = indices[e];
                if (biased && c == last_in) continue;
                const float v = keyed_value(data[e], row_scale, col_scale, r, c);
                if (keep_zeros || v != 0.0f) {
                    const int32_t cn = col_map ? col_map[c] : c;
                    const int pos = s_pref[cn >> 5] + __popc(s_bits[cn >> 5] & ((1u << (cn & 31)) - 1u));
                    out_indices[obeg + pos] = cn; out_data[obeg + pos] = v;
                }
            }
            __syncthreads();
            continue;
        }
        if (threadIdx.x == 0) s_count = 0;
        __syncthreads();
        for (int64_t e = beg + threadIdx.x; e < end; e += blockDim.x) {
            const int32_t c = indices[e];
            if (biased && c == last_in) continue;
            const float v = keyed_value(data[e], row_scale, col_scale, r, c);
            if (keep_zeros || v != 0.0f) {
                const int pos = atomicAdd(&s_count, 1);         // order is irrelevant: sorted next
                const int32_t cn = col_map ? col_map[c] : c;
                if (in_smem) { s_key[pos] = cn; s_val[pos] = v; }
                else { out_indices[obeg + pos] = cn; out_data[obeg + pos] = v; }
            }
        }
        __syncthreads();
        const int n = (int)n_out;
        if (in_smem) {
            bitonic_sort_pairs(s_key, s_val, n);
            for (int i = threadIdx.x; i < n; i += blockDim.x) { out_indices[obeg + i] = s_key[i]; out_data[obeg + i] = s_val[i]; }
        } else {
            __threadfence_block();
            bitonic_sort_pairs(out_indices + obeg, out_data + obeg, n);
        }
        __syncthreads();
    }
}

__global__ void gather_rows_count_kernel(const int64_t *__restrict__ indptr, const int64_t *__restrict__ row_ids, int64_t n_rows, int64_t *__restrict__ row_nnz) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const int64_t s = row_ids ? row_ids[i] : i;
    row_nnz[i] = indptr[s + 1] - indptr[s];
}

__global__ void __launch_bounds__(kThreads)
gather_rows_fill_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices, const float *__restrict__ data,
                        const int64_t *__restrict__ row_ids, int64_t n_rows,
                        const int64_t *__restrict__ out_indptr, int32_t *__restrict__ out_indices, float *__restrict__ out_data)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = kThreads / 32;
    for (int64_t i = (int64_t)blockIdx.x * wpb + warp; i < n_rows; i += (int64_t)gridDim.x * wpb) {
        const int64_t s = row_ids ? row_ids[i] : i;
        const int64_t beg = indptr[s], n = indptr[s + 1] - beg, obeg = out_indptr[i];
        for (int64_t e = lane; e < n; e += 32) { out_indices[obeg + e] = indices[beg + e]; out_data[obeg + e] = data[beg + e]; }
    }
}

int row_grid(int64_t n_rows, int rows_per_cta) {
    const int64_t want = kn_cdiv(n_rows, rows_per_cta);
    const int64_t cap = (int64_t)kn_sm_count() * 8;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}
}  // namespace

KN_API int kn_keycompile_count(const int64_t *indptr, const int32_t *indices, const float *data, int64_t n_rows,
                               const float *row_scale, const float *col_scale, const float *row_bias, const float *col_bias, int64_t n_cols_in,
                               int32_t keep_zeros, int64_t *row_nnz, void *stream) {
    KN_REQUIRE(n_rows >= 0, "keycompile: negative row count");
    if (n_rows == 0) return KN_OK;
    KN_REQUIRE(indptr && indices && data && row_nnz, "keycompile: null pointer");
    if (row_bias || col_bias) {
        KN_REQUIRE(n_cols_in > 0, "keycompile: bias keys need the input column count");
        keycompile_count_bias_kernel<<<row_grid(n_rows, 1), kThreads, 0, (cudaStream_t)stream>>>(indptr, indices, data, n_rows, row_scale, col_scale,
                                                                                             row_bias, col_bias, (int32_t)(n_cols_in - 1), keep_zeros, row_nnz);
        KN_CHECK_LAUNCH();
        return KN_OK;
    }
    keycompile_count_kernel<<<row_grid(n_rows, kThreads / 32), kThreads, 0, (cudaStream_t)stream>>>(indptr, indices, data, n_rows, row_scale, col_scale, keep_zeros, row_nnz);
    KN_CHECK_LAUNCH();
    return KN_OK;
}

KN_API int kn_keycompile_fill(const int64_t *indptr, const int32_t *indices, const float *data, int64_t n_rows, int64_t n_cols,
                              const int32_t *col_map, const float *row_scale, const float *col_scale,
                              const float *row_bias, const float *col_bias, int64_t n_cols_in, int32_t keep_zeros,
                              const int64_t *out_indptr, int32_t *out_indices, float *out_data, void *stream) {
    KN_REQUIRE(n_rows >= 0, "keycompile: negative row count");
    if (n_rows == 0) return KN_OK;
    KN_REQUIRE(indptr && indices && data && out_indptr && out_indices && out_data, "keycompile: null pointer");
    KN_REQUIRE(out_indices != indices && out_data != data, "keycompile: in-place compile is not supported");
    const size_t smem = (size_t)kSmemCap * (sizeof(int32_t) + sizeof(float));
    KN_ONCE_PER_DEVICE {
        KN_CUDA(cudaFuncSetAttribute(keycompile_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (!(row_bias || col_bias)) {
        keycompile_fill_short_kernel<<<row_grid(n_rows, kThreads / 32), kThreads, 0, (cudaStream_t)stream>>>(indptr, indices, data, n_rows, col_map, row_scale, col_scale, keep_zeros,
                                                                                                     out_indptr, out_indices, out_data);
        KN_CHECK_LAUNCH();
    }
    keycompile_fill_kernel<<<row_grid(n_rows, 1), kThreads, smem, (cudaStream_t)stream>>>(indptr, indices, data, n_rows, n_cols, col_map, row_scale, col_scale, row_bias, col_bias,
                                                                                         (row_bias || col_bias) ? (int32_t)(n_cols_in - 1) : -1, keep_zeros,
                                                                                         out_indptr, out_indices, out_data);
    KN_CHECK_LAUNCH();
    return KN_OK;
}

KN_API int kn_csr_gather_rows_count(const int64_t *indptr, const int64_t *row_ids, int64_t n_rows, int64_t *row_nnz, void *stream) {
    KN_REQUIRE(n_rows >= 0, "gather_rows: negative row count");
    if (n_rows == 0) return KN_OK;
    KN_REQUIRE(indptr && row_nnz, "gather_rows: null pointer");
    gather_rows_count_kernel<<<(unsigned)kn_cdiv(n_rows, 256), 256, 0, (cudaStream_t)stream>>>(indptr, row_ids, n_rows, row_nnz);
    KN_CHECK_LAUNCH();
    return KN_OK;
}

KN_API int kn_csr_gather_rows_fill(const int64_t *indptr, const int32_t *indices, const float *data,
                                   const int64_t *row_ids, int64_t n_rows,
                                   const int64_t *out_indptr, int32_t *out_indices, float *out_data, void *stream) {
    KN_REQUIRE(n_rows >= 0, "gather_rows: negative row count");
    if (n_rows == 0) return KN_OK;
    KN_REQUIRE(indptr && indices && data && out_indptr && out_indices && out_data, "gather_rows: null pointer");
    gather_rows_fill_kernel<<<row_grid(n_rows, kThreads / 32), kThreads, 0, (cudaStream_t)stream>>>(indptr, indices, data, row_ids, n_rows, out_indptr, out_indices, out_data);
    KN_CHECK_LAUNCH();
    return KN_OK;
}
