// scan.cu -- int64 prefix sum used by the two-phase CSR builders (count -> scan -> fill).
// Three passes (tile sums, scan of tile sums, tile rescan); the inputs are row counts, i.e. at
// most a few million elements, so this is launch-latency bound, not bandwidth bound.
#include "common.cuh"

namespace {
constexpr int kThreads = 256;
constexpr int kItems = 8;
constexpr int kTile = kThreads * kItems;

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t *total, int64_t *s_warp /*[8]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int64_t warp_off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) {
        const int64_t sw = s_warp[w];
        if (w < warp) warp_off += sw;
        tot += sw;
    }
    __syncthreads();
    *total = tot;
    return warp_off + incl - v;
}

__global__ void __launch_bounds__(kThreads) tile_sum_kernel(const int64_t *__restrict__ in, int64_t n, int64_t *__restrict__ tile_sums) {
    __shared__ int64_t s_warp[kThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
    int64_t v = 0;
#pragma unroll
    for (int i = 0; i < kItems; i++) if (base + i < n) v += in[base + i];
    int64_t tot;
    block_exclusive_scan(v, &tot, s_warp);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kThreads) scan_tile_sums_kernel(int64_t *__restrict__ tile_sums, int64_t n_tiles) {
    __shared__ int64_t s_warp[kThreads / 32];
    int64_t carry = 0;
    for (int64_t base = 0; base < n_tiles; base += kThreads) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = (i < n_tiles) ? tile_sums[i] : 0;
        int64_t tot;
        const int64_t excl = block_exclusive_scan(v, &tot, s_warp);
        if (i < n_tiles) tile_sums[i] = carry + excl;
        carry += tot;
    }
}

__global__ void __launch_bounds__(kThreads) tile_scan_kernel(const int64_t *in, int64_t n, const int64_t *__restrict__ tile_offsets, int64_t *out) {
    __shared__ int64_t s_warp[kThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
    int64_t vals[kItems];
    int64_t v = 0;
#pragma unroll
    for (int i = 0; i < kItems; i++) { vals[i] = (base + i < n) ? in[base + i] : 0; v += vals[i]; }
    int64_t tot;
    int64_t run = block_exclusive_scan(v, &tot, s_warp) + tile_offsets[blockIdx.x];
    // block_exclusive_scan ends with __syncthreads(): every thread of this tile has read its inputs,
    // so writing out[i+1] is safe even when `in` aliases `out + 1`
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = 0;
#pragma unroll
    for (int i = 0; i < kItems; i++) {
        run += vals[i];
        if (base + i < n) out[base + i + 1] = run;
    }
}
}  // namespace

KN_API int kn_exclusive_scan_i64(const int64_t *in, int64_t *out, int64_t n, void *stream) {
    KN_REQUIRE(n >= 0, "scan: negative length");
    KN_REQUIRE(out != nullptr, "scan: null output");
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) { KN_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t), s)); return KN_OK; }
    KN_REQUIRE(in != nullptr, "scan: null input");
    const int64_t n_tiles = kn_cdiv(n, kTile);
    int64_t *tile_sums = nullptr;
    KN_CUDA(cudaMallocAsync((void **)&tile_sums, (size_t)n_tiles * sizeof(int64_t), s));
    tile_sum_kernel<<<(unsigned)n_tiles, kThreads, 0, s>>>(in, n, tile_sums);
    scan_tile_sums_kernel<<<1, kThreads, 0, s>>>(tile_sums, n_tiles);
    tile_scan_kernel<<<(unsigned)n_tiles, kThreads, 0, s>>>(in, n, tile_sums, out);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(tile_sums, s);
    if (e != cudaSuccess) { kn_set_error("scan launch failed: %s", cudaGetErrorString(e)); return KN_ERR_CUDA; }
    return KN_OK;
}
