// layout.cu -- activation layout helpers: the homogeneous coordinate (reference
// keynet/torch.py:65-77) fused with the batch-major <-> feature-major transpose that the
// reference performs as `x_affine.t()` / `.t()` around every SpMM (keynet/layer.py:92).
// The layer chain itself stays feature-major ([dim+1][N]) from encrypt to the last layer.
#include "common.cuh"

namespace {
constexpr int kTileDim = 32;

// images [n_vecs][dim] -> X [dim+1][ldx], X[dim][:] = 1
__global__ void __launch_bounds__(kTileDim * 8)
affine_to_linear_t_kernel(const float *__restrict__ img, int64_t n_vecs, int64_t dim, float *__restrict__ X, int64_t ldx) {
    __shared__ float tile[kTileDim][kTileDim + 1];
    const int64_t d0 = (int64_t)blockIdx.x * kTileDim, n0 = (int64_t)blockIdx.y * kTileDim;
    for (int j = threadIdx.y; j < kTileDim; j += 8) {           // read: rows = batch, cols = features (coalesced)
        const int64_t n = n0 + j, d = d0 + threadIdx.x;
        tile[j][threadIdx.x] = (n < n_vecs && d < dim) ? img[n * dim + d] : 0.0f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < kTileDim; j += 8) {           // write: rows = features, cols = batch (coalesced)
        const int64_t d = d0 + j, n = n0 + threadIdx.x;
        if (n < n_vecs && d <= dim) X[d * ldx + n] = (d == dim) ? 1.0f : tile[threadIdx.x][j];
    }
}

// sensor.encrypt() in one pass for a monomial image key (permutation x gain [+ bias column]): the homogenised image column d
// (d == dim: the constant 1) lands in row row_of_col[d] scaled by scale_of_col[d], plus row_bias[row] when the key has a
// bias column -- i.e. Y = A . affine_to_linear(images)^T without materialising X.  Values are rounded like the SpMM's
// (one product, then one addition).
constexpr int kEncN = 128;        // batch columns per CTA: every keyed row receives 512 contiguous bytes
__global__ void __launch_bounds__(256)
encrypt_monomial_t_kernel(const float *__restrict__ img, int64_t n_vecs, int64_t dim, const int32_t *__restrict__ row_of_col,
                          const float *__restrict__ scale_of_col, const float *__restrict__ row_bias, float *__restrict__ Y, int64_t ldy) {
    __shared__ float tile[kTileDim][kEncN + 1];                 // [feature][batch], odd stride: both phases conflict-free
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t d0 = (int64_t)blockIdx.x * kTileDim, n0 = (int64_t)blockIdx.y * kEncN;
    const int64_t d = d0 + lane;
#pragma unroll 4
    for (int j = warp; j < kEncN; j += 8) {                     // read: one 128-byte piece of image n0+j per warp
        const int64_t n = n0 + j;
        tile[lane][j] = (n < n_vecs && d < dim) ? __ldg(img + n * dim + d) : 1.0f;     // d == dim: homogeneous coordinate
    }
    __syncthreads();
    for (int j = warp; j < kTileDim; j += 8) {                  // write: 512 bytes of keyed row row_of_col[d0+j]
        const int64_t dj = d0 + j;
        if (dj > dim) break;                                    // warp-uniform
        const int64_t r = __ldg(row_of_col + dj);
        const float sc = __ldg(scale_of_col + dj);
        const float b = (row_bias != nullptr && dj != dim) ? __ldg(row_bias + r) : 0.0f;
        float *__restrict__ yr = Y + r * ldy + n0;
#pragma unroll
        for (int i = 0; i < kEncN / 32; i++) {
            const int nn = lane + 32 * i;
            if (n0 + nn < n_vecs) {
                float y = __fmul_rn(sc, tile[j][nn]);
                if (row_bias != nullptr && dj != dim) y = __fadd_rn(y, b);
                yr[nn] = y;
            }
        }
    }
}

// X [dim+1][ldx] -> out [n_vecs][dim]; counts vectors whose last coordinate is not ~1
__global__ void __launch_bounds__(kTileDim * 8)
linear_to_affine_t_kernel(const float *__restrict__ X, int64_t ldx, int64_t n_vecs, int64_t dim, float *__restrict__ out,
                          float atol, int32_t *__restrict__ bad_count) {
    __shared__ float tile[kTileDim][kTileDim + 1];
    const int64_t d0 = (int64_t)blockIdx.x * kTileDim, n0 = (int64_t)blockIdx.y * kTileDim;
    for (int j = threadIdx.y; j < kTileDim; j += 8) {
        const int64_t d = d0 + j, n = n0 + threadIdx.x;
        tile[j][threadIdx.x] = (n < n_vecs && d < dim) ? X[d * ldx + n] : 0.0f;
    }
    if (blockIdx.x == 0 && threadIdx.y == 0 && bad_count != nullptr) {
        const int64_t n = n0 + threadIdx.x;
        if (n < n_vecs) {
            const float h = X[dim * ldx + n];
            if (!(fabsf(h - 1.0f) <= atol)) atomicAdd(bad_count, 1);   // also catches NaN
        }
    }
    __syncthreads();
    for (int j = threadIdx.y; j < kTileDim; j += 8) {
        const int64_t n = n0 + j, d = d0 + threadIdx.x;
        if (n < n_vecs && d < dim) out[n * dim + d] = tile[threadIdx.x][j];
    }
}
}  // namespace

KN_API int kn_affine_to_linear_t(const float *images, int64_t n_vecs, int64_t dim, float *X, int64_t ldx, void *stream) {
    KN_REQUIRE(n_vecs >= 0 && dim >= 0 && ldx >= n_vecs, "affine_to_linear: bad shape");
    if (n_vecs == 0) return KN_OK;
    KN_REQUIRE(images && X, "affine_to_linear: null pointer");
    const int64_t gx = kn_cdiv(dim + 1, kTileDim), gy = kn_cdiv(n_vecs, kTileDim);
    KN_REQUIRE(gy <= 65535, "affine_to_linear: batch too large for one call (%lld)", (long long)n_vecs);
    affine_to_linear_t_kernel<<<dim3((unsigned)gx, (unsigned)gy), dim3(kTileDim, 8), 0, (cudaStream_t)stream>>>(images, n_vecs, dim, X, ldx);
    KN_CHECK_LAUNCH();
    return KN_OK;
}

KN_API int kn_encrypt_monomial_t(const float *images, int64_t n_vecs, int64_t dim, const int32_t *row_of_col, const float *scale_of_col,
                                 const float *row_bias, float *Y, int64_t ldy, void *stream) {
    KN_REQUIRE(n_vecs >= 0 && dim >= 0 && ldy >= n_vecs, "encrypt_monomial: bad shape");
    if (n_vecs == 0) return KN_OK;
    KN_REQUIRE(images && row_of_col && scale_of_col && Y, "encrypt_monomial: null pointer");
    const int64_t gx = kn_cdiv(dim + 1, kTileDim), gy = kn_cdiv(n_vecs, kEncN);
    KN_REQUIRE(gy <= 65535, "encrypt_monomial: batch too large for one call (%lld)", (long long)n_vecs);
    encrypt_monomial_t_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, (cudaStream_t)stream>>>(images, n_vecs, dim, row_of_col, scale_of_col, row_bias, Y, ldy);
    KN_CHECK_LAUNCH();
    return KN_OK;
}

KN_API int kn_linear_to_affine_t(const float *X, int64_t ldx, int64_t n_vecs, int64_t dim, float *out,
                                 float atol, int32_t *bad_count_dev, void *stream) {
    KN_REQUIRE(n_vecs >= 0 && dim >= 0 && ldx >= n_vecs, "linear_to_affine: bad shape");
    if (n_vecs == 0) return KN_OK;
    KN_REQUIRE(X && out, "linear_to_affine: null pointer");
    const int64_t gx = kn_cdiv(dim > 0 ? dim : 1, kTileDim), gy = kn_cdiv(n_vecs, kTileDim);
    KN_REQUIRE(gy <= 65535, "linear_to_affine: batch too large for one call (%lld)", (long long)n_vecs);
    linear_to_affine_t_kernel<<<dim3((unsigned)gx, (unsigned)gy), dim3(kTileDim, 8), 0, (cudaStream_t)stream>>>(X, ldx, n_vecs, dim, out, atol, bad_count_dev);
    KN_CHECK_LAUNCH();
    return KN_OK;
}
