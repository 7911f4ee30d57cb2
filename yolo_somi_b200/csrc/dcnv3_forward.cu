// dcnv3_forward.cu -- DCNv3 core forward for sm_100a.
//
//   out[n,ho,wo,g,c] = sum_p mask[n,ho,wo,g,p] * bilinear(value[n,:,:,g,c], loc(n,ho,wo,g,p))
//
// Replaces the reference's dcnv3_im2col_gpu_kernel + dcnv3_im2col_bilinear
// (models/ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh:216-275, 32-80), which runs one thread per
// output CHANNEL and therefore re-reads the 18 offsets + 9 masks and redoes the coordinate math
// gc times per (pixel, group), gathering 2-byte scalars.
//
// Kernel `fwd_gather`: one thread owns NV 16-byte chunks (NV*8 bf16/fp16 or NV*4 fp32 channels)
// of one (n, ho, wo, g) -- a whole group at C=256/G=16 -- so coordinates, corner predicates and
// mask*bilinear weights are computed once per sampling point.  Offsets are read as (dx,dy) pairs;
// the corner reads of a tap are 128-bit and issued together (zero for out-of-map corners); 16-bit
// data is never unpacked: each element is folded into the fp32 accumulator with one (fast) or two
// (weight split hi+lo) FHFMA.  The output is written exactly once, no zero-fill pass
// (the reference zero-fills, dcnv3_cuda.cu:55-57).
// Kernel `fwd_gather_scalar`: generic fallback, one thread per channel (any group_channels,
// any alignment).
#include "dcnv3_common.cuh"
#include "dcnv3_launch.h"

#include <algorithm>
#include <cstdlib>

namespace dcnv3 {

constexpr int kFwdThreads = 128;

// Index decomposition of a flat (n, ho, wo, g) id.
struct PixelGroup {
    int n, ho, wo, g;
};
__device__ __forceinline__ PixelGroup split_pg(long long pg, const Geom &q) {
    PixelGroup r;
    r.g = (int)(pg % q.G);
    const long long pix = pg / q.G;
    r.wo = (int)(pix % q.Wo);
    const long long row = pix / q.Wo;
    r.ho = (int)(row % q.Ho);
    r.n = (int)(row / q.Ho);
    return r;
}

// KH/KW > 0: compile-time kernel extent (fully unrolled tap loop); 0: run-time extent.
// Grid: x covers the (ho, wo, g, part) threads of one image (32-bit index math), y = image.
template <typename T, int NV, int KH, int KW, bool FAST>
__global__ void __launch_bounds__(kFwdThreads)
fwd_gather(const T *__restrict__ value, const T *__restrict__ offset, const T *__restrict__ mask,
           T *__restrict__ out, const Geom q, const unsigned thr_per_image, const unsigned thr_per_group,
           const int n0) {
    constexpr int E = Chunk<T>::kElems;  // channels per 16-byte chunk
    const unsigned t = blockIdx.x * kFwdThreads + threadIdx.x;
    if (t >= thr_per_image) return;
    const int n = n0 + blockIdx.y;
    const unsigned part = t % thr_per_group;
    const unsigned pgl = t / thr_per_group;  // (ho, wo, g) within the image
    const unsigned g = pgl % q.G;
    const unsigned pix = pgl / q.G;
    const int wo = pix % q.Wo, ho = pix / q.Wo;

    const int kh = KH ? KH : q.kh, kw = KW ? KW : q.kw;
    const int P = kh * kw;
    const int C = q.G * q.gc;
    const int row_stride = q.W * C;  // elements; H*row_stride < 2^31 (checked by the C ABI)
    const size_t pg = (size_t)n * ((size_t)q.Ho * q.Wo * q.G) + pgl;
    const T *img = value + (size_t)n * q.H * row_stride + g * q.gc + part * (NV * E);
    const T *off = offset + pg * (P * 2);
    const T *msk = mask + pg * P;

    const float base_w = axis_base(wo, kw, q.sw, q.pw, q.dw, q.sigma);
    const float base_h = axis_base(ho, kh, q.sh, q.ph, q.dh, q.sigma);

    float acc[NV * E];
#pragma unroll
    for (int v = 0; v < NV * E; ++v) acc[v] = 0.f;

#pragma unroll
    for (int i = 0; i < kw; ++i) {
#pragma unroll
        for (int j = 0; j < kh; ++j) {
            const int p = i * kh + j;
            const float2 d = load_pair(off + 2 * p);
            const float m = to_f32(__ldg(msk + p));
            const float loc_w = base_w + ((float)(i * q.dw) + d.x) * q.sigma;
            const float loc_h = base_h + ((float)(j * q.dh) + d.y) * q.sigma;
            const ClampedTap tp = make_clamped_tap(loc_h, loc_w, q.H, q.W);
            if (!tp.inside) continue;
            const T *r_lo = img + tp.row_lo * row_stride, *r_hi = img + tp.row_hi * row_stride;
            const int c_lo = tp.col_lo * C, c_hi = tp.col_hi * C;
            uint4 v1[NV], v2[NV], v3[NV], v4[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                v1[k] = __ldg(reinterpret_cast<const uint4 *>(r_lo + c_lo + k * E));
                v2[k] = __ldg(reinterpret_cast<const uint4 *>(r_lo + c_hi + k * E));
                v3[k] = __ldg(reinterpret_cast<const uint4 *>(r_hi + c_lo + k * E));
                v4[k] = __ldg(reinterpret_cast<const uint4 *>(r_hi + c_hi + k * E));
            }
            const float fy_lo = tp.hh * tp.top * m, fy_hi = tp.lh * tp.bot * m;
            const float fx_lo = tp.hw * tp.lef, fx_hi = tp.lw * tp.rig;
            const Weight<T, FAST> w1(fy_lo * fx_lo), w2(fy_lo * fx_hi), w3(fy_hi * fx_lo), w4(fy_hi * fx_hi);
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                axpy<T, FAST>(acc + k * E, v1[k], w1);
                axpy<T, FAST>(acc + k * E, v2[k], w2);
                axpy<T, FAST>(acc + k * E, v3[k], w3);
                axpy<T, FAST>(acc + k * E, v4[k], w4);
            }
        }
    }
    T *dst = out + pg * q.gc + part * (NV * E);
#pragma unroll
    for (int k = 0; k < NV; ++k) *reinterpret_cast<uint4 *>(dst + k * E) = pack<T>(acc + k * E);
}

template <typename T>
__global__ void __launch_bounds__(256)
fwd_gather_scalar(const T *__restrict__ value, const T *__restrict__ offset,
                  const T *__restrict__ mask, T *__restrict__ out, const Geom q,
                  const long long n_threads) {
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t >= n_threads) return;
    const int c = (int)(t % q.gc);
    const long long pg = t / q.gc;
    const PixelGroup id = split_pg(pg, q);
    const int P = q.kh * q.kw;
    const int C = q.G * q.gc;
    const int row_stride = q.W * C;
    const T *img = value + (size_t)id.n * q.H * row_stride + id.g * q.gc + c;
    const float base_w = axis_base(id.wo, q.kw, q.sw, q.pw, q.dw, q.sigma);
    const float base_h = axis_base(id.ho, q.kh, q.sh, q.ph, q.dh, q.sigma);
    float acc = 0.f;
    for (int i = 0; i < q.kw; ++i)
        for (int j = 0; j < q.kh; ++j) {
            const int p = i * q.kh + j;
            const float dx = to_f32(__ldg(offset + (pg * P + p) * 2));
            const float dy = to_f32(__ldg(offset + (pg * P + p) * 2 + 1));
            const float m = to_f32(__ldg(mask + pg * P + p));
            const Tap tp = make_tap(base_h + ((float)(j * q.dh) + dy) * q.sigma,
                                    base_w + ((float)(i * q.dw) + dx) * q.sigma, q.H, q.W);
            if (!tp.inside) continue;
            const T *c1 = img + (tp.h0 * row_stride + tp.w0 * C);
            const float v1 = tp.tl ? to_f32(__ldg(c1)) : 0.f;
            const float v2 = tp.tr ? to_f32(__ldg(c1 + C)) : 0.f;
            const float v3 = tp.bl ? to_f32(__ldg(c1 + row_stride)) : 0.f;
            const float v4 = tp.br ? to_f32(__ldg(c1 + row_stride + C)) : 0.f;
            acc += (tp.hh * tp.hw * v1 + tp.hh * tp.lw * v2 + tp.lh * tp.hw * v3 + tp.lh * tp.lw * v4) * m;
        }
    out[t] = from_f32<T>(acc);
}

// ---------------------------------------------------------------------------------------------
bool fast_weights_requested() {
    // 16-bit I/O only.  Default ("fast"): the bilinear*mask coefficient of a corner is rounded to the
    // I/O dtype (2^-9 relative for bf16, 2^-12 for fp16) and folded with ONE FHFMA per element.
    // DCNV3_WEIGHTS=split keeps it fp32-accurate (hi+lo parts, two FHFMA per element).
    const char *e = std::getenv("DCNV3_WEIGHTS");
    return !(e && e[0] == 's');
}

template <typename T, int NV, bool FAST>
static cudaError_t launch_vec(const T *v, const T *o, const T *m, T *y, const Geom &q, cudaStream_t stream) {
    constexpr int E = Chunk<T>::kElems;
    const unsigned thr_per_group = q.gc / (NV * E);
    const long long per_image = (long long)q.Ho * q.Wo * q.G * thr_per_group;  // < 2^31 (C ABI check)
    const unsigned blocks = (unsigned)((per_image + kFwdThreads - 1) / kFwdThreads);
    for (int n0 = 0; n0 < q.N; n0 += 65535) {  // gridDim.y limit
        const dim3 grid(blocks, (unsigned)std::min(65535, q.N - n0));
        if (q.kh == 3 && q.kw == 3)
            fwd_gather<T, NV, 3, 3, FAST><<<grid, kFwdThreads, 0, stream>>>(v, o, m, y, q, (unsigned)per_image, thr_per_group, n0);
        else
            fwd_gather<T, NV, 0, 0, FAST><<<grid, kFwdThreads, 0, stream>>>(v, o, m, y, q, (unsigned)per_image, thr_per_group, n0);
    }
    return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_typed(const void *value, const void *offset, const void *mask, void *out,
                                const Geom &q, cudaStream_t stream) {
    constexpr int E = Chunk<T>::kElems;
    const T *v = static_cast<const T *>(value), *o = static_cast<const T *>(offset),
            *m = static_cast<const T *>(mask);
    T *y = static_cast<T *>(out);
    const long long n_groups = (long long)q.N * q.Ho * q.Wo * q.G;
    if (n_groups == 0) return cudaSuccess;
    // the 128-bit path needs whole chunks per group and 16-byte aligned value/out
    const bool vec_ok = (q.gc % E == 0) && (((uintptr_t)value | (uintptr_t)out) % 16 == 0);
    if (!vec_ok) {
        const long long n_threads = n_groups * q.gc;
        const long long blocks = (n_threads + 255) / 256;
        if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        fwd_gather_scalar<T><<<(unsigned)blocks, 256, 0, stream>>>(v, o, m, y, q, n_threads);
        return cudaGetLastError();
    }
    // one 16-byte chunk per lane keeps every warp-wide gather sector-efficient (adjacent lanes share
    // 32-byte sectors); DCNV3_NV=2 (development knob) gives a lane two chunks instead
    const char *nv_env = std::getenv("DCNV3_NV");
    const bool two = nv_env && nv_env[0] == '2' && (q.gc / E) % 2 == 0;
    if constexpr (sizeof(T) == 2) {
        if (fast_weights_requested())
            return two ? launch_vec<T, 2, true>(v, o, m, y, q, stream) : launch_vec<T, 1, true>(v, o, m, y, q, stream);
    }
    return two ? launch_vec<T, 2, false>(v, o, m, y, q, stream) : launch_vec<T, 1, false>(v, o, m, y, q, stream);
}

cudaError_t launch_forward(const void *value, const void *offset, const void *mask, void *out,
                           const Geom &q, int dtype, cudaStream_t stream) {
    cudaError_t err = cudaSuccess;
    if (try_launch_forward_gs(value, offset, mask, out, q, dtype, fast_weights_requested(), stream, &err))
        return err;
    if (try_launch_forward_tile(value, offset, mask, out, q, dtype, fast_weights_requested(), stream, &err))
        return err;
    if (try_launch_forward_mma(value, offset, mask, out, q, dtype, stream, &err)) return err;
    switch (dtype) {
    case 0: return launch_typed<float>(value, offset, mask, out, q, stream);
    case 1: return launch_typed<__half>(value, offset, mask, out, q, stream);
    default: return launch_typed<__nv_bfloat16>(value, offset, mask, out, q, stream);
    }
}

}  // namespace dcnv3
