// toeplitz.cu -- closed-form CSR construction of the layer matrices (kernel K3 of SURVEY.md 2.2).
//
// The reference emits COO triplets from a 6-deep serial numba loop, materialises 3 arrays of
// Uo*Vo*C*M*P*Q entries and converts them with scipy (keynet/sparse.py:122-203).  Here every
// row of the Toeplitz matrix is written directly in CSR order from its closed form:
//
//   source row s = (m, ku, kv) -> u = ku*stride, v = kv*stride
//   valid taps:   p in [p0, p1) with 0 <= u+p < U,   q in [q0, q1) with 0 <= v+q < V
//   entry order:  c ascending, then p, then q   ==  ascending column  c*U*V + (u+p)*V + (v+q)
//   then the bias entry in column C*U*V; source row R = M*Uo*Vo is the homogeneous row e_last.
//
// `row_ids` selects and orders the source rows, which is how an output permutation key (row
// gather) and a row shard are applied without ever materialising the un-keyed matrix rows that
// the rank does not own.  One warp writes one row; lanes stride over the row's entries so index
// and value stores are coalesced.
#include "common.cuh"

namespace {

constexpr int kWarps = 8;

struct RowGeom {
    int m, u, v;        // output channel, top-left input coordinate of the window centre
    int p0, np, q0, nq; // first valid tap offset and number of valid taps per axis
    bool last;          // homogeneous row
};

__device__ __forceinline__ RowGeom row_geom(const kn_conv2d_desc &d, int64_t s) {
    RowGeom g;
    const int Uo = d.U / d.stride, Vo = d.V / d.stride;
    const int64_t R = (int64_t)d.M * Uo * Vo;
    g.last = (s >= R);
    if (g.last) { g.m = g.u = g.v = g.p0 = g.np = g.q0 = g.nq = 0; return g; }
    g.m = (int)(s / ((int64_t)Uo * Vo));
    const int rem = (int)(s - (int64_t)g.m * Uo * Vo);
    const int ku = rem / Vo, kv = rem - ku * Vo;
    g.u = ku * d.stride; g.v = kv * d.stride;
    const int ph = (d.P - 1) / 2, qh = (d.Q - 1) / 2;      // odd kernels: taps -ph..+ph
    const int plo = max(-ph, -g.u), phi = min(ph, d.U - 1 - g.u);
    const int qlo = max(-qh, -g.v), qhi = min(qh, d.V - 1 - g.v);
    g.p0 = plo; g.np = max(0, phi - plo + 1);
    g.q0 = qlo; g.nq = max(0, qhi - qlo + 1);
    return g;
}

__global__ void toeplitz_count_kernel(kn_conv2d_desc d, const int64_t *__restrict__ row_ids, int64_t n_rows, int64_t *__restrict__ row_nnz) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const int64_t s = row_ids ? row_ids[i] : i;
    const RowGeom g = row_geom(d, s);
    int64_t n;
    if (g.last) n = 1;
    else n = (int64_t)(d.depthwise ? 1 : d.C) * g.np * g.nq + (d.has_bias ? 1 : 0);
    row_nnz[i] = n;
}

__global__ void __launch_bounds__(kWarps * 32)
toeplitz_fill_kernel(kn_conv2d_desc d, const float *__restrict__ weight, const float *__restrict__ bias,
                     const int64_t *__restrict__ row_ids, int64_t n_rows,
                     const int64_t *__restrict__ indptr, int32_t *__restrict__ indices, float *__restrict__ data)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t i = (int64_t)blockIdx.x * kWarps + warp; i < n_rows; i += (int64_t)gridDim.x * kWarps) {
        const int64_t s = row_ids ? row_ids[i] : i;
        const RowGeom g = row_geom(d, s);
        const int64_t out = indptr[i];
        const int32_t K = d.C * d.U * d.V;
        if (g.last) {
            if (lane == 0) { indices[out] = K; data[out] = 1.0f; }
            continue;
        }
        const int taps = g.np * g.nq;
        const int nch = d.depthwise ? 1 : d.C;
        const int n_main = nch * taps;
        const int ph = (d.P - 1) / 2, qh = (d.Q - 1) / 2;
        for (int e = lane; e < n_main; e += 32) {
            const int cc = e / taps, t = e - cc * taps;
            const int ip = t / g.nq, iq = t - ip * g.nq;
            const int p = g.p0 + ip, q = g.q0 + iq;
            const int c = d.depthwise ? g.m : cc;
            const int32_t col = c * d.U * d.V + (g.u + p) * d.V + (g.v + q);
            const int64_t widx = d.depthwise ? ((int64_t)g.m * d.P + (p + ph)) * d.Q + (q + qh)
                                             : (((int64_t)g.m * d.C + c) * d.P + (p + ph)) * d.Q + (q + qh);
            indices[out + e] = col;
            data[out + e] = __ldg(weight + widx);
        }
        if (d.has_bias && lane == 0) { indices[out + n_main] = K; data[out + n_main] = __ldg(bias + g.m); }
    }
}

// ---- nn.Linear: [[W, b],[0, 1]] with zeros dropped ------------------------------------------
__global__ void __launch_bounds__(kWarps * 32)
linear_count_kernel(const float *__restrict__ weight, const float *__restrict__ bias, int64_t n_out, int64_t n_in,
                    const int64_t *__restrict__ row_ids, int64_t n_rows, int64_t *__restrict__ row_nnz)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t i = (int64_t)blockIdx.x * kWarps + warp; i < n_rows; i += (int64_t)gridDim.x * kWarps) {
        const int64_t s = row_ids ? row_ids[i] : i;
        if (s >= n_out) { if (lane == 0) row_nnz[i] = 1; continue; }
        const float *__restrict__ w = weight + s * n_in;
        int cnt = 0;
        for (int64_t c = lane; c < n_in; c += 32) cnt += (w[c] != 0.0f) ? 1 : 0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        if (lane == 0) row_nnz[i] = cnt + ((bias != nullptr && bias[s] != 0.0f) ? 1 : 0);
    }
}

__global__ void __launch_bounds__(kWarps * 32)
linear_fill_kernel(const float *__restrict__ weight, const float *__restrict__ bias, int64_t n_out, int64_t n_in,
                   const int64_t *__restrict__ row_ids, int64_t n_rows,
                   const int64_t *__restrict__ indptr, int32_t *__restrict__ indices, float *__restrict__ data)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t i = (int64_t)blockIdx.x * kWarps + warp; i < n_rows; i += (int64_t)gridDim.x * kWarps) {
        const int64_t s = row_ids ? row_ids[i] : i;
        int64_t out = indptr[i];
        if (s >= n_out) { if (lane == 0) { indices[out] = (int32_t)n_in; data[out] = 1.0f; } continue; }
        const float *__restrict__ w = weight + s * n_in;
        for (int64_t c0 = 0; c0 < n_in; c0 += 32) {               // ordered compaction keeps columns ascending
            const int64_t c = c0 + lane;
            const float x = (c < n_in) ? w[c] : 0.0f;
            const unsigned mask = __ballot_sync(0xffffffffu, x != 0.0f);
            if (x != 0.0f) {
                const int64_t pos = out + __popc(mask & ((1u << lane) - 1u));
                indices[pos] = (int32_t)c; data[pos] = x;
            }
            out += __popc(mask);
        }
        if (lane == 0 && bias != nullptr && bias[s] != 0.0f) { indices[out] = (int32_t)n_in; data[out] = bias[s]; }
    }
}

int check_desc(const kn_conv2d_desc *d) {
    KN_REQUIRE(d != nullptr, "toeplitz: null descriptor");
    KN_REQUIRE(d->C > 0 && d->U > 0 && d->V > 0 && d->M > 0, "toeplitz: non-positive shape");
    KN_REQUIRE(d->P > 0 && d->Q > 0 && (d->P % 2) == 1 && (d->Q % 2) == 1, "toeplitz: kernel must be odd (P=%d Q=%d)", d->P, d->Q);
    KN_REQUIRE(d->stride > 0, "toeplitz: stride must be positive");
    KN_REQUIRE(!d->depthwise || d->M == d->C, "toeplitz: depthwise needs M == C");
    KN_REQUIRE((int64_t)d->C * d->U * d->V < 0x7fffffffLL, "toeplitz: column index exceeds int32");
    return KN_OK;
}

int grid_for_rows(int64_t n_rows) {
    const int64_t want = kn_cdiv(n_rows, kWarps);
    const int64_t cap = (int64_t)kn_sm_count() * 16;      // persistent-ish: a few CTAs per SM, grid-stride over rows
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}
}  // namespace

KN_API int kn_toeplitz_conv2d_count(const kn_conv2d_desc *desc, const int64_t *row_ids, int64_t n_rows, int64_t *row_nnz, void *stream) {
    int rc = check_desc(desc); if (rc) return rc;
    KN_REQUIRE(n_rows >= 0, "toeplitz: negative row count");
    if (n_rows == 0) return KN_OK;
    KN_REQUIRE(row_nnz != nullptr, "toeplitz: null row_nnz");
    toeplitz_count_kernel<<<(unsigned)kn_cdiv(n_rows, 256), 256, 0, (cudaStream_t)stream>>>(*desc, row_ids, n_rows, row_nnz);
    KN_CHECK_LAUNCH();
    return KN_OK;
}

KN_API int kn_toeplitz_conv2d_fill(const kn_conv2d_desc *desc, const float *weight, const float *bias,
                                   const int64_t *row_ids, int64_t n_rows,
                                   const int64_t *indptr, int32_t *indices, float *data, void *stream) {
    int rc = check_desc(desc); if (rc) return rc;
    KN_REQUIRE(n_rows >= 0, "toeplitz: negative row count");
    if (n_rows == 0) return KN_OK;
    KN_REQUIRE(weight && indptr && indices && data, "toeplitz: null pointer");
    KN_REQUIRE(!desc->has_bias || bias, "toeplitz: has_bias set but bias is null");
    toeplitz_fill_kernel<<<grid_for_rows(n_rows), kWarps * 32, 0, (cudaStream_t)stream>>>(*desc, weight, bias, row_ids, n_rows, indptr, indices, data);
    KN_CHECK_LAUNCH();
    return KN_OK;
}

KN_API int kn_linear_count(const float *weight, const float *bias, int64_t n_out, int64_t n_in,
                           const int64_t *row_ids, int64_t n_rows, int64_t *row_nnz, void *stream) {
    KN_REQUIRE(n_out > 0 && n_in > 0 && n_in < 0x7fffffffLL, "linear: bad shape");
    KN_REQUIRE(n_rows >= 0, "linear: negative row count");
    if (n_rows == 0) return KN_OK;
    KN_REQUIRE(weight && row_nnz, "linear: null pointer");
    linear_count_kernel<<<grid_for_rows(n_rows), kWarps * 32, 0, (cudaStream_t)stream>>>(weight, bias, n_out, n_in, row_ids, n_rows, row_nnz);
    KN_CHECK_LAUNCH();
    return KN_OK;
}

KN_API int kn_linear_fill(const float *weight, const float *bias, int64_t n_out, int64_t n_in,
                          const int64_t *row_ids, int64_t n_rows,
                          const int64_t *indptr, int32_t *indices, float *data, void *stream) {
    KN_REQUIRE(n_out > 0 && n_in > 0 && n_in < 0x7fffffffLL, "linear: bad shape");
    KN_REQUIRE(n_rows >= 0, "linear: negative row count");
    if (n_rows == 0) return KN_OK;
    KN_REQUIRE(weight && indptr && indices && data, "linear: null pointer");
    linear_fill_kernel<<<grid_for_rows(n_rows), kWarps * 32, 0, (cudaStream_t)stream>>>(weight, bias, n_out, n_in, row_ids, n_rows, indptr, indices, data);
    KN_CHECK_LAUNCH();
    return KN_OK;
}
