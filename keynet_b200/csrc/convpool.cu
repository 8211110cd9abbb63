// convpool.cu -- fused  keyed conv (+ReLU)  ->  keyed average pooling  for small layers (LeNet-sized):
//     Y = W_pool . relu(W_conv . X)
// with the intermediate (all conv outputs of an image: 4 705 rows for LeNet conv1, the largest activation of the network)
// held in SHARED MEMORY instead of being written to HBM by one launch and read back by the next.  Algorithmic bytes of the pair
// (per layer): (C_in + 2 R_conv + R_pool) * N * 4; bytes this kernel moves: (C_in + R_pool) * N * 4 -- for LeNet
// conv1 -> pool1 that is 1 962 rows instead of 11 372.
//
// Valid for permutation-only keys (reference PermutationKeynet: the key of a layer followed by ReLU is a permutation, which
// commutes with ReLU; the conv's output key and the pool's input key cancel, so the intermediate needs no key at all):
//   xrow[c*U*V + y*V + x]  = activation row holding input pixel (c, y, x) under the conv's input key (last entry: homogeneous row)
//   yrow[m*Up*Vp + py*Vp + px] = output row of pooled pixel (m, py, px) under the pool's output key (last entry: homogeneous row)
// One CTA = NB batch columns of the whole image: (1) gather the NB-column slab of X into smem, (2) direct convolution from
// smem to smem (+bias, ReLU), register tile = all / 4 output channels x NB columns per pixel, (3) pooling from smem, rows
// scattered to Y.  fp32 FMA throughout (these layers have 1..6 input channels: no tensor-core shape).
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int kThreads = 256;

template <int NB, int CG>
__global__ void __launch_bounds__(kThreads)
convpool_kernel(kn_conv2d_desc d, const float *__restrict__ weight, const float *__restrict__ bias, const int32_t *__restrict__ xrow,
                int pk, int pstride, float pool_w, const int32_t *__restrict__ yrow,
                const float *__restrict__ X, int64_t ldx, float *__restrict__ Y, int64_t ldy, int64_t n_vecs)
{
    extern __shared__ float smem_f[];
    const int Uo = d.U / d.stride, Vo = d.V / d.stride;          // conv output size
    const int Up = Uo / pstride, Vp = Vo / pstride;              // pooled size
    const int n_in = d.C * d.U * d.V, n_mid = d.M * Uo * Vo, n_out = d.M * Up * Vp;
    const int CPQ = d.C * d.P * d.Q;
    float *xs = smem_f;                                          // [n_in][NB]
    float *ms = xs + (size_t)n_in * NB;                          // [n_mid][NB]
    const int Mp = (d.M + CG - 1) / CG * CG;                     // output channels padded to whole register groups (zero weights)
    float *ws = ms + (size_t)n_mid * NB;                         // [Mp][CPQ] then bias[Mp]
    const int64_t n0 = (int64_t)blockIdx.x * NB;
    const int tid = threadIdx.x;
    typedef float4 V4;

    // ---- (1) weights and the NB-column slab of X -> smem
    for (int i = tid; i < Mp * CPQ; i += kThreads) ws[i] = (i < d.M * CPQ) ? __ldg(weight + i) : 0.0f;
    for (int i = tid; i < Mp; i += kThreads) ws[Mp * CPQ + i] = (i < d.M) ? __ldg(bias + i) : 0.0f;
    for (int i = tid; i < n_in; i += kThreads) {
        const float *src = X + (int64_t)__ldg(xrow + i) * ldx + n0;
#pragma unroll
        for (int j = 0; j < NB; j += 4) {
            V4 v = (n0 + j < n_vecs) ? *reinterpret_cast<const V4 *>(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<V4 *>(xs + (size_t)i * NB + j) = v;
        }
    }
    __syncthreads();

    // ---- (2) direct convolution smem -> smem, bias, ReLU.  item = (conv pixel, group of CG output channels)
    const int ph = (d.P - 1) / 2, qh = (d.Q - 1) / 2;
    const int n_groups = (d.M + CG - 1) / CG;
    for (int it = tid; it < Uo * Vo * n_groups; it += kThreads) {
        const int grp = it / (Uo * Vo), pix = it - grp * (Uo * Vo);
        const int ky = pix / Vo, kx = pix - ky * Vo;
        const int y = ky * d.stride, x = kx * d.stride;
        const int m0 = grp * CG;
        float acc[CG][NB];
#pragma unroll
        for (int g = 0; g < CG; g++) {
            const float b = ws[Mp * CPQ + m0 + g];
#pragma unroll
            for (int j = 0; j < NB; j++) acc[g][j] = b;
        }
        for (int c = 0; c < d.C; c++) {
            for (int p = -ph; p <= ph; p++) {
                const int yy = y + p;
                if (yy < 0 || yy >= d.U) continue;
                for (int q = -qh; q <= qh; q++) {
                    const int xx = x + q;
                    if (xx < 0 || xx >= d.V) continue;
                    float xv[NB];
                    const float *xp = xs + (size_t)((c * d.U + yy) * d.V + xx) * NB;
#pragma unroll
                    for (int j = 0; j < NB; j += 4) { const V4 v = *reinterpret_cast<const V4 *>(xp + j); xv[j] = v.x; xv[j + 1] = v.y; xv[j + 2] = v.z; xv[j + 3] = v.w; }
                    const int wi = (c * d.P + (p + ph)) * d.Q + (q + qh);
#pragma unroll
                    for (int g = 0; g < CG; g++) {
                        const float w = ws[(m0 + g) * CPQ + wi];
#pragma unroll
                        for (int j = 0; j < NB; j++) acc[g][j] = fmaf(w, xv[j], acc[g][j]);
                    }
                }
            }
        }
#pragma unroll
        for (int g = 0; g < CG; g++) {
            if (m0 + g < d.M) {
                float *mp = ms + (size_t)((m0 + g) * Uo * Vo + pix) * NB;
#pragma unroll
                for (int j = 0; j < NB; j += 4)
                    *reinterpret_cast<V4 *>(mp + j) = make_float4(fmaxf(acc[g][j], 0.f), fmaxf(acc[g][j + 1], 0.f), fmaxf(acc[g][j + 2], 0.f), fmaxf(acc[g][j + 3], 0.f));
            }
        }
    }
    __syncthreads();

    // ---- (3) average pooling (centred pk x pk windows, zero padding, divisor pk*pk) smem -> Y rows
    const int hk = (pk - 1) / 2;
    for (int it = tid; it < n_out; it += kThreads) {
        const int m = it / (Up * Vp), pp = it - m * (Up * Vp);
        const int py = pp / Vp, px = pp - py * Vp;
        float s[NB];
#pragma unroll
        for (int j = 0; j < NB; j++) s[j] = 0.0f;
        for (int dy = -hk; dy <= hk; dy++) {
            const int yy = py * pstride + dy;
            if (yy < 0 || yy >= Uo) continue;
            for (int dx = -hk; dx <= hk; dx++) {
                const int xx = px * pstride + dx;
                if (xx < 0 || xx >= Vo) continue;
                const float *mp = ms + (size_t)((m * Uo + yy) * Vo + xx) * NB;
#pragma unroll
                for (int j = 0; j < NB; j += 4) { const V4 v = *reinterpret_cast<const V4 *>(mp + j); s[j] = fmaf(pool_w, v.x, s[j]); s[j + 1] = fmaf(pool_w, v.y, s[j + 1]); s[j + 2] = fmaf(pool_w, v.z, s[j + 2]); s[j + 3] = fmaf(pool_w, v.w, s[j + 3]); }
            }
        }
        float *dst = Y + (int64_t)__ldg(yrow + it) * ldy + n0;
#pragma unroll
        for (int j = 0; j < NB; j += 4)
            if (n0 + j < n_vecs) *reinterpret_cast<V4 *>(dst + j) = make_float4(s[j], s[j + 1], s[j + 2], s[j + 3]);
    }
    // homogeneous coordinate: passes through both layers unchanged
    if (tid < NB && n0 + tid < n_vecs) Y[(int64_t)__ldg(yrow + n_out) * ldy + n0 + tid] = X[(int64_t)__ldg(xrow + n_in) * ldx + n0 + tid];
}

template <int NB, int CG>
int launch_convpool(const kn_conv2d_desc *d, const float *w, const float *b, const int32_t *xrow, int pk, int ps, float pw, const int32_t *yrow,
                    const float *X, int64_t ldx, float *Y, int64_t ldy, int64_t n_vecs, cudaStream_t s) {
    const int Uo = d->U / d->stride, Vo = d->V / d->stride;
    const int Mp = (d->M + CG - 1) / CG * CG;
    const size_t smem = ((size_t)d->C * d->U * d->V * NB + (size_t)d->M * Uo * Vo * NB + (size_t)Mp * (d->C * d->P * d->Q + 1)) * sizeof(float);
    KN_REQUIRE(smem <= 227 * 1024, "convpool: image of %zu bytes per CTA does not fit shared memory", smem);
    KN_ONCE_PER_DEVICE {
        KN_CUDA(cudaFuncSetAttribute(convpool_kernel<NB, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    convpool_kernel<NB, CG><<<(unsigned)kn_cdiv(n_vecs, NB), kThreads, smem, s>>>(*d, w, b, xrow, pk, ps, pw, yrow, X, ldx, Y, ldy, n_vecs);
    KN_CHECK_LAUNCH();
    return KN_OK;
}
}  // namespace

KN_API int kn_convpool_f32(const kn_conv2d_desc *desc, const float *weight, const float *bias, const int32_t *xrow,
                           int32_t pool_k, int32_t pool_stride, float pool_w, const int32_t *yrow,
                           const float *X, int64_t ldx, float *Y, int64_t ldy, int64_t n_vecs, void *stream) {
    KN_REQUIRE(desc && desc->C > 0 && desc->M > 0 && desc->U > 0 && desc->V > 0 && desc->stride > 0 && (desc->P % 2) == 1 && (desc->Q % 2) == 1 && desc->has_bias && !desc->depthwise,
               "convpool: bad descriptor");
    KN_REQUIRE(pool_k > 0 && (pool_k % 2) == 1 && pool_stride > 0, "convpool: pooling window must be odd");
    KN_REQUIRE((desc->U / desc->stride) % pool_stride == 0 && (desc->V / desc->stride) % pool_stride == 0, "convpool: pooled size must divide the conv output");
    KN_REQUIRE(n_vecs >= 0 && ldx >= n_vecs && ldy >= n_vecs, "convpool: bad leading dimension");
    if (n_vecs == 0) return KN_OK;
    KN_REQUIRE(weight && bias && xrow && yrow && X && Y, "convpool: null pointer");
    KN_REQUIRE(n_vecs % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && (((uintptr_t)X | (uintptr_t)Y) & 15) == 0, "convpool: n_vecs, ldx, ldy must be multiples of 4 and X, Y 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const int Uo = desc->U / desc->stride, Vo = desc->V / desc->stride;
    const size_t per_col = ((size_t)desc->C * desc->U * desc->V + (size_t)desc->M * Uo * Vo) * sizeof(float);
    static const int force_nb = getenv("KN_CONVPOOL_NB") ? atoi(getenv("KN_CONVPOOL_NB")) : 0;
    const bool nb8 = (n_vecs % 8 == 0) && (force_nb == 8 || (force_nb == 0 && per_col * 8 + 8192 <= 110 * 1024));   // two CTAs per SM with 8 columns each
    const int cg = (desc->M % 6 == 0 && desc->M <= 12) ? 6 : 4;                      // register tile: output channels per item
    if (cg == 6) {
        if (nb8) return launch_convpool<8, 6>(desc, weight, bias, xrow, pool_k, pool_stride, pool_w, yrow, X, ldx, Y, ldy, n_vecs, s);
        return launch_convpool<4, 6>(desc, weight, bias, xrow, pool_k, pool_stride, pool_w, yrow, X, ldx, Y, ldy, n_vecs, s);
    }
    if (nb8) return launch_convpool<8, 4>(desc, weight, bias, xrow, pool_k, pool_stride, pool_w, yrow, X, ldx, Y, ldy, n_vecs, s);
    return launch_convpool<4, 4>(desc, weight, bias, xrow, pool_k, pool_stride, pool_w, yrow, X, ldx, Y, ldy, n_vecs, s);
}
