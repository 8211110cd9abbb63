// spgemm.cu -- C = A . B for CSR matrices: the key compile W_hat = A . W . Ainv for GENERAL keys.
//
// Permutation / gain / bias keys have one entry per row and compile by reindexing (keycompile.cu).  The reference's
// other key families -- Givens-rotation orthogonal blocks (keynet/sparse.py:288-309), doubly stochastic blocks with
// their dense inverses (:335-353), block-diagonal repeats of both (keynet/system.py:398-410) -- have several entries per
// row, and the reference runs scipy's csr_matmat twice (keynet/layer.py:35,59,70).  This is that product on the GPU:
//
//   kn_spgemm_bound   ub[r] = sum over the entries (r, j) of A of nnz(B[j, :])          (expansion size of row r)
//   kn_spgemm_rows    one CTA per row: expand the products a_rj * b_jc into the row's slice of a scratch CSR, sort them
//                     by column (bitonic network, in shared memory up to 8192 products, in place in global memory
//                     beyond), add duplicates in fp32 in sorted order, drop exact zeros (csr_matmat never stores one),
//                     compact to the front of the slice; row_nnz[r] = entries kept
//   kn_csr_compact    scratch slices -> final CSR (after an exclusive scan of row_nnz)
//
// Result: canonical CSR (sorted columns).  Entries with one or two contributions -- all of them for Givens keys with few
// rotations -- are bit-identical to scipy's; longer sums differ by the order of the fp32 additions.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kSmemCap = 8192;   // products per row sorted in shared memory (64 KB: column + value)

__global__ void __launch_bounds__(kThreads)
spgemm_bound_kernel(const int64_t *__restrict__ a_indptr, const int32_t *__restrict__ a_indices, int64_t n_rows,
                    const int64_t *__restrict__ b_indptr, int64_t *__restrict__ ub)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = kThreads / 32;
    for (int64_t r = (int64_t)blockIdx.x * wpb + warp; r < n_rows; r += (int64_t)gridDim.x * wpb) {
        int64_t cnt = 0;
        for (int64_t e = a_indptr[r] + lane; e < a_indptr[r + 1]; e += 32) {
            const int32_t j = a_indices[e];
            cnt += b_indptr[j + 1] - b_indptr[j];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        if (lane == 0) ub[r] = cnt;
    }
}

// merge the sorted products k[0..n), v[0..n) of one row: duplicates summed, zeros dropped, result written to
// (ok, ov)[0..count); (ok, ov) may alias (k, v) when both are global (compaction only moves entries forward)
template <typename KeyPtr, typename ValPtr>
__device__ __forceinline__ int merge_sorted_row(KeyPtr k, ValPtr v, int n, int32_t *__restrict__ ok, float *__restrict__ ov, int *s_scan) {
    int run_off = 0;
    for (int base = 0; base < n; base += kThreads) {
        const int i = base + threadIdx.x;
        int flag = 0; int32_t key = 0; float s = 0.0f;
        if (i < n) {
            key = k[i];
            if (i == 0 || k[i - 1] != key) {
                s = v[i];
                for (int j = i + 1; j < n && k[j] == key; j++) s = __fadd_rn(s, v[j]);
                flag = (s != 0.0f) ? 1 : 0;
            }
        }
        __syncthreads();                                     // every read of this tile is done before anything is overwritten
        // block-wide exclusive scan of flag
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        int incl = flag;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += t; }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        if (threadIdx.x == 0) { int run = 0; for (int w = 0; w < kThreads / 32; w++) { const int t = s_scan[w]; s_scan[w] = run; run += t; } s_scan[kThreads / 32] = run; }
        __syncthreads();
        if (flag) { const int pos = run_off + s_scan[warp] + incl - 1; ok[pos] = key; ov[pos] = s; }
        run_off += s_scan[kThreads / 32];
        __syncthreads();
    }
    return run_off;
}

__global__ void __launch_bounds__(kThreads)
spgemm_rows_kernel(const int64_t *__restrict__ a_indptr, const int32_t *__restrict__ a_indices, const float *__restrict__ a_data, int64_t n_rows,
                   const int64_t *__restrict__ b_indptr, const int32_t *__restrict__ b_indices, const float *__restrict__ b_data,
                   const int64_t *__restrict__ tmp_ptr, int32_t *__restrict__ tmp_indices, float *__restrict__ tmp_data, int64_t *__restrict__ row_nnz)
{
    extern __shared__ unsigned char smem_raw[];
    int32_t *s_key = reinterpret_cast<int32_t *>(smem_raw);
    float *s_val = reinterpret_cast<float *>(smem_raw + sizeof(int32_t) * kSmemCap);
    __shared__ int s_count;
    __shared__ int s_scan[kThreads / 32 + 1];

    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const int64_t t0 = tmp_ptr[r];
        const int64_t L = tmp_ptr[r + 1] - t0;
        if (L == 0) { if (threadIdx.x == 0) row_nnz[r] = 0; continue; }            // block-uniform
        const bool in_smem = L <= kSmemCap;
        int32_t *__restrict__ gk = tmp_indices + t0;
        float *__restrict__ gv = tmp_data + t0;
        if (threadIdx.x == 0) s_count = 0;
        __syncthreads();
        // ---- expand: every entry (r, j) of A contributes a_rj * B[j, :]; slots are reserved with one atomic per entry
        //      (the order inside the buffer is irrelevant, it is sorted next)
        for (int64_t e = a_indptr[r] + threadIdx.x; e < a_indptr[r + 1]; e += kThreads) {
            const int32_t j = a_indices[e];
            const float a = a_data[e];
            const int64_t bb = b_indptr[j];
            const int lb = (int)(b_indptr[j + 1] - bb);
            const int pos = atomicAdd(&s_count, lb);
            for (int t = 0; t < lb; t++) {
                const int32_t c = b_indices[bb + t];
                const float p = __fmul_rn(a, b_data[bb + t]);
                if (in_smem) { s_key[pos + t] = c; s_val[pos + t] = p; }
                else { gk[pos + t] = c; gv[pos + t] = p; }
            }
        }
        __syncthreads();
        const int n = (int)L;
        int kept;
        if (in_smem) {
            bitonic_sort_pairs(s_key, s_val, n);
            kept = merge_sorted_row(s_key, s_val, n, gk, gv, s_scan);
        } else {
            __threadfence_block();
            bitonic_sort_pairs(gk, gv, n);
            __threadfence_block();
            kept = merge_sorted_row(gk, gv, n, gk, gv, s_scan);
        }
        if (threadIdx.x == 0) row_nnz[r] = kept;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads)
csr_compact_kernel(const int64_t *__restrict__ tmp_ptr, const int32_t *__restrict__ tmp_indices, const float *__restrict__ tmp_data, int64_t n_rows,
                   const int64_t *__restrict__ out_indptr, int32_t *__restrict__ out_indices, float *__restrict__ out_data)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = kThreads / 32;
    for (int64_t r = (int64_t)blockIdx.x * wpb + warp; r < n_rows; r += (int64_t)gridDim.x * wpb) {
        const int64_t src = tmp_ptr[r], dst = out_indptr[r], n = out_indptr[r + 1] - dst;
        for (int64_t i = lane; i < n; i += 32) { out_indices[dst + i] = tmp_indices[src + i]; out_data[dst + i] = tmp_data[src + i]; }
    }
}

int row_grid(int64_t n_rows, int per_block) {
    const int64_t want = kn_cdiv(n_rows, per_block);
    const int64_t cap = (int64_t)kn_sm_count() * 16;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}
}  // namespace

KN_API int kn_spgemm_bound(const int64_t *a_indptr, const int32_t *a_indices, int64_t n_rows, const int64_t *b_indptr, int64_t *ub, void *stream) {
    KN_REQUIRE(n_rows >= 0, "spgemm_bound: negative row count");
    if (n_rows == 0) return KN_OK;
    KN_REQUIRE(a_indptr && b_indptr && ub, "spgemm_bound: null pointer");
    spgemm_bound_kernel<<<row_grid(n_rows, kThreads / 32), kThreads, 0, (cudaStream_t)stream>>>(a_indptr, a_indices, n_rows, b_indptr, ub);
    KN_CHECK_LAUNCH();
    return KN_OK;
}

KN_API int kn_spgemm_rows(const int64_t *a_indptr, const int32_t *a_indices, const float *a_data, int64_t n_rows,
                          const int64_t *b_indptr, const int32_t *b_indices, const float *b_data,
                          const int64_t *tmp_ptr, int32_t *tmp_indices, float *tmp_data, int64_t *row_nnz, void *stream) {
    KN_REQUIRE(n_rows >= 0, "spgemm_rows: negative row count");
    if (n_rows == 0) return KN_OK;
    KN_REQUIRE(a_indptr && b_indptr && tmp_ptr && row_nnz, "spgemm_rows: null pointer");
    const size_t smem = (size_t)kSmemCap * (sizeof(int32_t) + sizeof(float));
    KN_ONCE_PER_DEVICE {
        KN_CUDA(cudaFuncSetAttribute(spgemm_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    spgemm_rows_kernel<<<row_grid(n_rows, 1), kThreads, smem, (cudaStream_t)stream>>>(a_indptr, a_indices, a_data, n_rows, b_indptr, b_indices, b_data,
                                                                                        tmp_ptr, tmp_indices, tmp_data, row_nnz);
    KN_CHECK_LAUNCH();
    return KN_OK;
}

KN_API int kn_csr_compact(const int64_t *tmp_ptr, const int32_t *tmp_indices, const float *tmp_data, int64_t n_rows,
                          const int64_t *out_indptr, int32_t *out_indices, float *out_data, void *stream) {
    KN_REQUIRE(n_rows >= 0, "csr_compact: negative row count");
    if (n_rows == 0) return KN_OK;
    KN_REQUIRE(tmp_ptr && out_indptr, "csr_compact: null pointer");
    csr_compact_kernel<<<row_grid(n_rows, kThreads / 32), kThreads, 0, (cudaStream_t)stream>>>(tmp_ptr, tmp_indices, tmp_data, n_rows, out_indptr, out_indices, out_data);
    KN_CHECK_LAUNCH();
    return KN_OK;
}
