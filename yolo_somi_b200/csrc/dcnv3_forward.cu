// dcnv3_forward.cu -- DCNv3 core forward for sm_100a.
//
//   out[n,ho,wo,g,c] = sum_p mask[n,ho,wo,g,p] * bilinear(value[n,:,:,g,c], loc(n,ho,wo,g,p))
//
// Replaces the reference's dcnv3_im2col_gpu_kernel + dcnv3_im2col_bilinear
// (models/ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh:216-275, 32-80), which runs one thread per
// output CHANNEL and therefore re-reads the 18 offsets + 9 masks and redoes the coordinate math
// gc times per (pixel, group), gathering 2-byte scalars.
//
// Kernel `fwd_gather` (this file): one thread owns VEC contiguous channels (one 128-bit access)
// of one (n, ho, wo, g); offsets are read as (dx,dy) pairs; the four corner reads of a tap are
// issued together (predicated, zero for out-of-map corners) and folded with mask*bilinear weights
// in fp32.  Output is written exactly once (no zero-fill pass, unlike dcnv3_cuda.cu:55-57).
#include "dcnv3_common.cuh"
#include "dcnv3_launch.h"

namespace dcnv3 {

constexpr int kFwdThreads = 256;

// KH/KW > 0: compile-time kernel extent (fully unrolled tap loop); 0: run-time extent.
template <typename T, int VEC, int KH, int KW>
__global__ void __launch_bounds__(kFwdThreads)
fwd_gather(const T *__restrict__ value, const T *__restrict__ offset, const T *__restrict__ mask,
           T *__restrict__ out, const Geom q, const long long n_threads, const int vec_per_group) {
    const long long t = (long long)blockIdx.x * kFwdThreads + threadIdx.x;
    if (t >= n_threads) return;
    const int cv = (int)(t % vec_per_group);
    const long long pg = t / vec_per_group;  // flat (n, ho, wo, g)
    const int g = (int)(pg % q.G);
    const long long pix = pg / q.G;
    const int wo = (int)(pix % q.Wo);
    const long long row = pix / q.Wo;
    const int ho = (int)(row % q.Ho);
    const int n = (int)(row / q.Ho);

    const int kh = KH ? KH : q.kh, kw = KW ? KW : q.kw;
    const int P = kh * kw;
    const int C = q.G * q.gc;
    const T *img = value + (size_t)n * q.H * q.W * C + g * q.gc + cv * VEC;
    const T *off = offset + pg * P * 2;
    const T *msk = mask + pg * P;

    const float base_w = axis_base(wo, kw, q.sw, q.pw, q.dw, q.sigma);
    const float base_h = axis_base(ho, kh, q.sh, q.ph, q.dh, q.sigma);

    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;

#pragma unroll
    for (int i = 0; i < kw; ++i) {
#pragma unroll
        for (int j = 0; j < kh; ++j) {
            const int p = i * kh + j;
            const float2 d = load_pair(off + 2 * p);
            const float m = to_f32(__ldg(msk + p));
            const float loc_w = base_w + ((float)(i * q.dw) + d.x) * q.sigma;
            const float loc_h = base_h + ((float)(j * q.dh) + d.y) * q.sigma;
            const Tap tp = make_tap(loc_h, loc_w, q.H, q.W);
            if (!tp.inside) continue;
            const T *c1 = img + ((ptrdiff_t)tp.h0 * q.W + tp.w0) * C;
            float v1[VEC], v2[VEC], v3[VEC], v4[VEC];
            ChanVec<T, VEC>::load(c1, tp.tl, v1);
            ChanVec<T, VEC>::load(c1 + C, tp.tr, v2);
            ChanVec<T, VEC>::load(c1 + (ptrdiff_t)q.W * C, tp.bl, v3);
            ChanVec<T, VEC>::load(c1 + (ptrdiff_t)q.W * C + C, tp.br, v4);
            const float w1 = tp.hh * tp.hw * m, w2 = tp.hh * tp.lw * m;
            const float w3 = tp.lh * tp.hw * m, w4 = tp.lh * tp.lw * m;
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                acc[v] += w1 * v1[v] + w2 * v2[v] + w3 * v3[v] + w4 * v4[v];
        }
    }
    ChanVec<T, VEC>::store(out + pg * q.gc + cv * VEC, acc);
}

template <typename T, int VEC>
static cudaError_t launch_typed(const void *value, const void *offset, const void *mask, void *out,
                                const Geom &q, cudaStream_t stream) {
    const int vec_per_group = q.gc / VEC;
    const long long n_threads = (long long)q.N * q.Ho * q.Wo * q.G * vec_per_group;
    if (n_threads == 0) return cudaSuccess;
    const long long blocks = (n_threads + kFwdThreads - 1) / kFwdThreads;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    const T *v = static_cast<const T *>(value), *o = static_cast<const T *>(offset),
            *m = static_cast<const T *>(mask);
    T *y = static_cast<T *>(out);
    if (q.kh == 3 && q.kw == 3)
        fwd_gather<T, VEC, 3, 3><<<(unsigned)blocks, kFwdThreads, 0, stream>>>(v, o, m, y, q, n_threads, vec_per_group);
    else
        fwd_gather<T, VEC, 0, 0><<<(unsigned)blocks, kFwdThreads, 0, stream>>>(v, o, m, y, q, n_threads, vec_per_group);
    return cudaGetLastError();
}

template <typename T, int VEC>
static cudaError_t launch_by_alignment(const void *value, const void *offset, const void *mask,
                                       void *out, const Geom &q, cudaStream_t stream) {
    // the 128-bit path needs gc % VEC == 0 and 16-byte aligned value/out base pointers
    const bool vec_ok = (q.gc % VEC == 0) && (((uintptr_t)value | (uintptr_t)out) % 16 == 0);
    return vec_ok ? launch_typed<T, VEC>(value, offset, mask, out, q, stream)
                  : launch_typed<T, 1>(value, offset, mask, out, q, stream);
}

cudaError_t launch_forward(const void *value, const void *offset, const void *mask, void *out,
                           const Geom &q, int dtype, cudaStream_t stream) {
    switch (dtype) {
    case 0: return launch_by_alignment<float, 4>(value, offset, mask, out, q, stream);
    case 1: return launch_by_alignment<__half, 8>(value, offset, mask, out, q, stream);
    default: return launch_by_alignment<__nv_bfloat16, 8>(value, offset, mask, out, q, stream);
    }
}

}  // namespace dcnv3
