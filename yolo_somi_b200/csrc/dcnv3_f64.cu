// dcnv3_f64.cu -- DCNv3 core forward and backward for fp64 I/O (dtype tag DCNV3_F64).
//
// The reference dispatches double as well as float / half (AT_DISPATCH_FLOATING_TYPES_AND_HALF,
// models/ops_dcnv3/src/cuda/dcnv3_cuda.cu:69,147) and its own test script drives the extension in double first
// (models/ops_dcnv3/test.py:33-57,88-150: forward and all three gradients against dcnv3_core_pytorch in double, for
// group_channels in {1, 16, 30, 32, 64, 71, 1025}).  This file exists so that script runs unmodified on top of this
// library; it is a correctness path, not a tuned one: all arithmetic in double (the reference's opmath_t for double),
// any group_channels, any kernel / stride / pad / dilation.
//
//   forward   one thread per (n, ho, wo, g, c); the channel is the fastest thread index, so the four corner reads
//             of a warp are contiguous runs (dcnv3_im2col_cuda.cuh:216-275, 32-80).
//   backward  one WARP per (n, ho, wo, g): the lanes stride over the group's channels; per point they add their
//             corner shares to grad_value with fp64 atomics (zeroed by the launcher) and reduce the three channel
//             sums (grad_mask, grad_offset x / y) with shuffles; lane 0 writes them once (:82-147, 278-370).
//             A point that fails the range test writes zeros (the reference relies on its zero-filled outputs).
// offset_scale arrives as a float through the C ABI, as it does through the reference's own signature
// (src/dcnv3.h:27: `const float offset_scale`).
#include "dcnv3_launch.h"

namespace dcnv3 {
namespace f64 {

struct TapD {
    bool inside, tl, tr, bl, br;
    int h0, w0;
    double lh, lw, hh, hw;
};

__device__ __forceinline__ TapD make_tap_d(double loc_h, double loc_w, int H, int W) {
    TapD t;
    t.inside = loc_h > -1.0 && loc_w > -1.0 && loc_h < (double)H && loc_w < (double)W;   // :262-263
    const double fh = floor(loc_h), fw = floor(loc_w);
    t.h0 = (int)fh;
    t.w0 = (int)fw;
    t.lh = loc_h - fh;
    t.lw = loc_w - fw;
    t.hh = 1.0 - t.lh;
    t.hw = 1.0 - t.lw;
    const bool top = t.h0 >= 0, bot = t.h0 + 1 <= H - 1, lef = t.w0 >= 0, rig = t.w0 + 1 <= W - 1;
    t.tl = t.inside && top && lef;
    t.tr = t.inside && top && rig;
    t.bl = t.inside && bot && lef;
    t.br = t.inside && bot && rig;
    return t;
}

// (c - pad + o * stride) - c * sigma with c = (dil * (k - 1)) >> 1   (:232-236,249-252)
__device__ __forceinline__ double anchor(int o, int k, int stride, int pad, int dil, double sigma) {
    const int c = (dil * (k - 1)) >> 1;
    return (double)(c - pad + o * stride) - (double)c * sigma;
}

struct PixelGroupD { int n, ho, wo, g; };
__device__ __forceinline__ PixelGroupD split(long long pg, const Geom &q) {
    PixelGroupD r;
    r.g = (int)(pg % q.G);
    const long long pix = pg / q.G;
    r.wo = (int)(pix % q.Wo);
    const long long row = pix / q.Wo;
    r.ho = (int)(row % q.Ho);
    r.n = (int)(row / q.Ho);
    return r;
}

__global__ void __launch_bounds__(256)
fwd_f64(const double *__restrict__ value, const double *__restrict__ offset, const double *__restrict__ mask,
        double *__restrict__ out, const Geom q, const long long n_threads) {
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t >= n_threads) return;
    const int c = (int)(t % q.gc);
    const long long pg = t / q.gc;
    const PixelGroupD id = split(pg, q);
    const int P = q.kh * q.kw, C = q.G * q.gc;
    const long long row_stride = (long long)q.W * C;
    const double sigma = (double)q.sigma;
    const double *img = value + (size_t)id.n * q.H * row_stride + id.g * q.gc + c;
    const double base_w = anchor(id.wo, q.kw, q.sw, q.pw, q.dw, sigma), base_h = anchor(id.ho, q.kh, q.sh, q.ph, q.dh, sigma);
    double acc = 0.0;
    for (int i = 0; i < q.kw; ++i)
        for (int j = 0; j < q.kh; ++j) {
            const int p = i * q.kh + j;
            const double dx = __ldg(offset + (pg * P + p) * 2), dy = __ldg(offset + (pg * P + p) * 2 + 1);
            const double m = __ldg(mask + pg * P + p);
            const TapD tp = make_tap_d(base_h + ((double)(j * q.dh) + dy) * sigma, base_w + ((double)(i * q.dw) + dx) * sigma, q.H, q.W);
            if (!tp.inside) continue;
            const double *c1 = img + ((long long)tp.h0 * row_stride + (long long)tp.w0 * C);
            const double v1 = tp.tl ? __ldg(c1) : 0.0, v2 = tp.tr ? __ldg(c1 + C) : 0.0;
            const double v3 = tp.bl ? __ldg(c1 + row_stride) : 0.0, v4 = tp.br ? __ldg(c1 + row_stride + C) : 0.0;
            acc += (tp.hh * tp.hw * v1 + tp.hh * tp.lw * v2 + tp.lh * tp.hw * v3 + tp.lh * tp.lw * v4) * m;
        }
    out[t] = acc;
}

__global__ void __launch_bounds__(256)
bwd_f64(const double *__restrict__ value, const double *__restrict__ offset, const double *__restrict__ mask,
        const double *__restrict__ grad_out, double *__restrict__ grad_value, double *__restrict__ grad_offset,
        double *__restrict__ grad_mask, const Geom q, const long long n_groups) {
    const int lane = threadIdx.x & 31;
    const long long pg = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (pg >= n_groups) return;
    const PixelGroupD id = split(pg, q);
    const int P = q.kh * q.kw, C = q.G * q.gc;
    const long long row_stride = (long long)q.W * C;
    const double sigma = (double)q.sigma;
    const size_t img_off = (size_t)id.n * q.H * row_stride + id.g * q.gc;
    const double *img = value + img_off;
    double *gimg = grad_value + img_off;
    const double *go = grad_out + pg * q.gc;
    const double base_w = anchor(id.wo, q.kw, q.sw, q.pw, q.dw, sigma), base_h = anchor(id.ho, q.kh, q.sh, q.ph, q.dh, sigma);
    for (int i = 0; i < q.kw; ++i)
        for (int j = 0; j < q.kh; ++j) {
            const int p = i * q.kh + j;
            const double dx = __ldg(offset + (pg * P + p) * 2), dy = __ldg(offset + (pg * P + p) * 2 + 1);
            const double m = __ldg(mask + pg * P + p);
            const TapD tp = make_tap_d(base_h + ((double)(j * q.dh) + dy) * sigma, base_w + ((double)(i * q.dw) + dx) * sigma, q.H, q.W);
            double gm = 0.0, gx = 0.0, gy = 0.0;
            if (tp.inside) {
                const long long o1 = (long long)tp.h0 * row_stride + (long long)tp.w0 * C;
                const double w1 = tp.hh * tp.hw, w2 = tp.hh * tp.lw, w3 = tp.lh * tp.hw, w4 = tp.lh * tp.lw;
                for (int c = lane; c < q.gc; c += 32) {
                    const double g = __ldg(go + c), gmk = g * m;
                    const double v1 = tp.tl ? __ldg(img + o1 + c) : 0.0, v2 = tp.tr ? __ldg(img + o1 + C + c) : 0.0;
                    const double v3 = tp.bl ? __ldg(img + o1 + row_stride + c) : 0.0;
                    const double v4 = tp.br ? __ldg(img + o1 + row_stride + C + c) : 0.0;
                    gm += g * (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
                    gx += (-tp.hh * v1 + tp.hh * v2 - tp.lh * v3 + tp.lh * v4) * gmk;
                    gy += (-tp.hw * v1 - tp.lw * v2 + tp.hw * v3 + tp.lw * v4) * gmk;
                    if (tp.tl) atomicAdd(gimg + o1 + c, w1 * gmk);
                    if (tp.tr) atomicAdd(gimg + o1 + C + c, w2 * gmk);
                    if (tp.bl) atomicAdd(gimg + o1 + row_stride + c, w3 * gmk);
                    if (tp.br) atomicAdd(gimg + o1 + row_stride + C + c, w4 * gmk);
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    gm += __shfl_xor_sync(0xffffffffu, gm, d);
                    gx += __shfl_xor_sync(0xffffffffu, gx, d);
                    gy += __shfl_xor_sync(0xffffffffu, gy, d);
                }
            }
            if (lane == 0) {
                grad_offset[(pg * P + p) * 2] = sigma * gx;
                grad_offset[(pg * P + p) * 2 + 1] = sigma * gy;
                grad_mask[pg * P + p] = gm;
            }
        }
}

}  // namespace f64

cudaError_t launch_forward_f64(const void *value, const void *offset, const void *mask, void *out, const Geom &q,
                               cudaStream_t stream) {
    const long long n_threads = (long long)q.N * q.Ho * q.Wo * q.G * q.gc;
    if (n_threads == 0) return cudaSuccess;
    const long long blocks = (n_threads + 255) / 256;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    f64::fwd_f64<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const double *>(value), static_cast<const double *>(offset),
                                                       static_cast<const double *>(mask), static_cast<double *>(out), q, n_threads);
    return cudaGetLastError();
}

cudaError_t launch_backward_f64(const void *value, const void *offset, const void *mask, const void *grad_out,
                                void *grad_value, void *grad_offset, void *grad_mask, const Geom &q, cudaStream_t stream) {
    const size_t v_bytes = (size_t)q.N * q.H * q.W * q.G * q.gc * sizeof(double);
    if (v_bytes) {
        const cudaError_t e = cudaMemsetAsync(grad_value, 0, v_bytes, stream);
        if (e != cudaSuccess) return e;
    }
    const long long n_groups = (long long)q.N * q.Ho * q.Wo * q.G;
    if (n_groups == 0) return cudaSuccess;
    const long long blocks = (n_groups + 7) / 8;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    f64::bwd_f64<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const double *>(value), static_cast<const double *>(offset),
                                                       static_cast<const double *>(mask), static_cast<const double *>(grad_out),
                                                       static_cast<double *>(grad_value), static_cast<double *>(grad_offset),
                                                       static_cast<double *>(grad_mask), q, n_groups);
    return cudaGetLastError();
}

}  // namespace dcnv3
