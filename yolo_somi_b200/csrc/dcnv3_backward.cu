// dcnv3_backward.cu -- DCNv3 core backward for sm_100a.
//
// For upstream gradient g = grad_out[n,ho,wo,g,c] and every sampling point p of (n,ho,wo,g):
//   grad_value[corner_k] += w_k * m * g                                   (scatter, 4 corners)
//   grad_mask[p]          = sum_c g * (w1 v1 + w2 v2 + w3 v3 + w4 v4)
//   grad_offset[p].x      = sigma * m * sum_c g * (hh (v2 - v1) + lh (v4 - v3))
//   grad_offset[p].y      = sigma * m * sum_c g * (hw (v3 - v1) + lw (v4 - v2))
// (models/ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh:82-147; points failing the range test give 0).
//
// Replaces the reference's dcnv3_col2im_* family (:278-839), which launches one block of
// `group_channels` threads per (n,ho,wo,g) (16 threads at C=256/G=16), lets thread 0 add up the
// per-channel partials serially behind 18 __syncthreads, issues 4 scalar fp32 atomics per
// (point, channel), and needs zero-filled fp32 gradient buffers plus three cast passes for half
// (dcnv3_cuda.cu:126-133,168-173).
//
// Kernel `bwd_scatter` (this file): one thread owns VEC contiguous channels of one (n,ho,wo,g).
//   * the channel sums of grad_mask / grad_offset are folded to four per-corner dot products
//     d_k = sum_c g_c v_k,c in registers, then combined across the gc/VEC lanes of the group with
//     xor-shuffles; lane 0 of the group writes grad_offset / grad_mask once, in the I/O dtype --
//     no atomics, no zero-fill, deterministic;
//   * grad_value contributions leave as 128-bit vector reductions (REDG.E.ADD.F32x4) into an fp32
//     accumulator: grad_value itself for fp32 I/O, a scratch plane for 16-bit I/O which one cast
//     pass then narrows;
//   * DETERMINISTIC mode accumulates grad_value in 64-bit fixed point (integer adds commute, so
//     the result does not depend on arrival order), scaled by a power of two derived on the device
//     from max|grad_out|.
// Kernel `bwd_scatter_warp`: generic fallback (any group_channels, any alignment): one warp per
// (n,ho,wo,g), lanes stride over channels, full-warp shuffle reduction.
#include "dcnv3_common.cuh"
#include "dcnv3_launch.h"

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <type_traits>

namespace dcnv3 {

constexpr int kBwdThreads = 256;
constexpr int kFixedBits = 40;  // fixed-point fraction: contributions are scaled to ~2^40 * |g|/max|g|

// ---------------------------------------------------------------------------------------------
// grad_value accumulators
struct AccumF32 {
    float *buf;
    template <int VEC>
    __device__ __forceinline__ void add(size_t idx, const float (&c)[VEC]) const {
        if constexpr (VEC % 4 == 0) {
#pragma unroll
            for (int v = 0; v < VEC; v += 4)
                atomicAdd(reinterpret_cast<float4 *>(buf + idx + v),
                          make_float4(c[v], c[v + 1], c[v + 2], c[v + 3]));
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) atomicAdd(buf + idx + v, c[v]);
        }
    }
};

struct AccumFixed {
    unsigned long long *buf;
    float scale;  // power of two
    template <int VEC>
    __device__ __forceinline__ void add(size_t idx, const float (&c)[VEC]) const {
#pragma unroll
        for (int v = 0; v < VEC; ++v)
            atomicAdd(buf + idx + v, (unsigned long long)__float2ll_rn(c[v] * scale));
    }
};

// Power-of-two scale for the fixed-point accumulator: 2^(kFixedBits - ceil(log2(amax))).
__device__ __forceinline__ float fixed_scale(const unsigned *amax_bits) {
    const float amax = __uint_as_float(*amax_bits);
    if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
    int e;
    frexpf(amax, &e);  // amax = f * 2^e, f in [0.5, 1)
    return ldexpf(1.f, kFixedBits - e);
}

template <typename T>
__global__ void __launch_bounds__(256)
absmax_kernel(const T *__restrict__ x, size_t n, unsigned *__restrict__ out_bits) {
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = fabsf(to_f32(x[i]));
        m = (v > m && isfinite(v)) ? v : m;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    // non-negative floats order like their bit patterns; max is order-independent -> deterministic
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));
}

template <typename T>
__global__ void __launch_bounds__(256)
narrow_f32_kernel(const float *__restrict__ src, T *__restrict__ dst, size_t n, const unsigned long long *cond = nullptr,
                  unsigned long long thr = 0) {
    // 8 elements per thread per step: two 128-bit loads, one 128-bit store (src/dst 16-byte aligned)
    // back to front: the accumulating kernels work image-major, so the END of the plane is what is still in L2
    asm volatile("griddepcontrol.wait;" ::: "memory");   // (a no-op unless launched as a programmatic dependent)
    if (cond) {                                          // (conditional fall-back of dcnv3_backward_vres.cu)
        unsigned long long count;                        // a volatile load: must not move above the wait
        asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(count) : "l"(cond) : "memory");
        if (count <= thr) return;
    }
    const size_t n8 = n / 8;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n8; j += (size_t)gridDim.x * blockDim.x) {
        const size_t i = n8 - 1 - j;
        const float4 a = __ldcs(reinterpret_cast<const float4 *>(src) + 2 * i);
        const float4 b = __ldcs(reinterpret_cast<const float4 *>(src) + 2 * i + 1);
        if constexpr (sizeof(T) == 2) {
            uint4 o;
            o.x = pack2(a.x, a.y, T()); o.y = pack2(a.z, a.w, T());
            o.z = pack2(b.x, b.y, T()); o.w = pack2(b.z, b.w, T());
            reinterpret_cast<uint4 *>(dst)[i] = o;
        } else {
            reinterpret_cast<float4 *>(dst)[2 * i] = a;
            reinterpret_cast<float4 *>(dst)[2 * i + 1] = b;
        }
    }
    for (size_t i = n8 * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = from_f32<T>(src[i]);
}

template <typename T>
__global__ void __launch_bounds__(256)
narrow_fixed_kernel(const unsigned long long *__restrict__ src, const unsigned *__restrict__ amax_bits,
                    T *__restrict__ dst, size_t n) {
    const double inv = 1.0 / (double)fixed_scale(amax_bits);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = from_f32<T>((float)((double)(long long)src[i] * inv));
}

// ---------------------------------------------------------------------------------------------
// One thread owns NV 16-byte chunks of one (n,ho,wo,g).  PAIR: lanes 2j/2j+1 swap halves of their
// fp32 contributions so that every 128-bit reduction instruction covers whole 32-byte sectors
// (lane 2j sends the even 16-byte pieces of both lanes' contributions, lane 2j+1 the odd ones).
// Grid: x covers the (ho, wo, g, part) threads of one image (32-bit index math), y = image.
template <typename T, int NV, int KH, int KW, typename Accum, bool PAIR>
__global__ void __launch_bounds__(kBwdThreads)
bwd_scatter(const T *__restrict__ value, const T *__restrict__ offset, const T *__restrict__ mask,
            const T *__restrict__ grad_out, Accum accum, const unsigned *__restrict__ amax_bits,
            T *__restrict__ grad_offset, T *__restrict__ grad_mask, const Geom q,
            const unsigned thr_per_image, const unsigned thr_per_group, const int n0) {
    constexpr int E = Chunk<T>::kElems;
    constexpr int CH = NV * E;       // channels per thread
    constexpr int R = CH / 4;        // 16-byte fp32 pieces per corner contribution
    static_assert(!PAIR || (R % 2 == 0), "pair exchange needs an even number of pieces");
    unsigned t = blockIdx.x * kBwdThreads + threadIdx.x;
    // whole lane-groups are live or not (thr_per_image is a multiple of thr_per_group, which
    // divides 32); dead lanes still take part in the shuffles but touch no memory
    const bool live = t < thr_per_image;
    if (!live) t = thr_per_image - 1;
    const int n = n0 + blockIdx.y;
    const unsigned part = t % thr_per_group;
    const unsigned pgl = t / thr_per_group;
    const unsigned g = pgl % q.G;
    const unsigned pix = pgl / q.G;
    const int wo = pix % q.Wo, ho = pix / q.Wo;

    if constexpr (sizeof(accum.buf[0]) == 8) accum.scale = fixed_scale(amax_bits);

    const int kh = KH ? KH : q.kh, kw = KW ? KW : q.kw;
    const int P = kh * kw;
    const int C = q.G * q.gc;
    const int row_stride = q.W * C;
    const size_t pg = (size_t)n * ((size_t)q.Ho * q.Wo * q.G) + pgl;
    const size_t img_base = (size_t)n * q.H * row_stride + g * q.gc + part * CH;
    const T *img = value + img_base;
    const T *off = offset + pg * (P * 2);
    const T *msk = mask + pg * P;
    T *goff = grad_offset + pg * (P * 2);
    T *gmsk = grad_mask + pg * P;

    uint4 gq[NV];   // upstream gradient, packed (for the exact corner dot products)
    float gf[CH];   // and in fp32 (for the grad_value contributions)
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        gq[k] = ldg16(grad_out + pg * q.gc + part * CH + k * E, live);
        unpack<T>(gq[k], gf + k * E);
    }
    // PAIR: the 16-byte pieces of parity (lane & 1) of BOTH lanes' gradients -- `g_even` for the
    // point owned by the even lane of the pair, `g_odd` for the odd lane's point
    float g_even[PAIR ? CH / 2 : 1], g_odd[PAIR ? CH / 2 : 1];
    const int odd = threadIdx.x & 1;
    if constexpr (PAIR) {
#pragma unroll
        for (int r = 0; r < R; r += 2)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                // I keep my piece r+odd and need the partner's piece r+odd; the partner needs my r+!odd
                const float mine = odd ? gf[(r + 1) * 4 + e] : gf[r * 4 + e];
                const float send = odd ? gf[r * 4 + e] : gf[(r + 1) * 4 + e];
                const float theirs = __shfl_xor_sync(0xffffffffu, send, 1);
                g_even[(r / 2) * 4 + e] = odd ? theirs : mine;
                g_odd[(r / 2) * 4 + e] = odd ? mine : theirs;
            }
    }

    const float base_w = axis_base(wo, kw, q.sw, q.pw, q.dw, q.sigma);
    const float base_h = axis_base(ho, kh, q.sh, q.ph, q.dh, q.sigma);

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

            float gm = 0.f, gx = 0.f, gy = 0.f;
            float coef[4] = {0.f, 0.f, 0.f, 0.f};   // w_k * m per corner; 0 = nothing to add
            int at[4] = {0, 0, 0, 0};               // corner offsets inside the image (elements)
            if (tp.inside) {
                const int r_lo = tp.row_lo * row_stride, r_hi = tp.row_hi * row_stride;
                const int c_lo = tp.col_lo * C, c_hi = tp.col_hi * C;
                at[0] = r_lo + c_lo; at[1] = r_lo + c_hi; at[2] = r_hi + c_lo; at[3] = r_hi + c_hi;
                float d1 = 0.f, d2 = 0.f, d3 = 0.f, d4 = 0.f;
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    const uint4 v1 = __ldg(reinterpret_cast<const uint4 *>(img + at[0] + k * E));
                    const uint4 v2 = __ldg(reinterpret_cast<const uint4 *>(img + at[1] + k * E));
                    const uint4 v3 = __ldg(reinterpret_cast<const uint4 *>(img + at[2] + k * E));
                    const uint4 v4 = __ldg(reinterpret_cast<const uint4 *>(img + at[3] + k * E));
                    d1 = dot<T>(gq[k], v1, d1);
                    d2 = dot<T>(gq[k], v2, d2);
                    d3 = dot<T>(gq[k], v3, d3);
                    d4 = dot<T>(gq[k], v4, d4);
                }
                // zero padding lives in the 1-D factors (see ClampedTap)
                const float fy_lo = tp.hh * tp.top, fy_hi = tp.lh * tp.bot;
                const float fx_lo = tp.hw * tp.lef, fx_hi = tp.lw * tp.rig;
                const float w1 = fy_lo * fx_lo, w2 = fy_lo * fx_hi, w3 = fy_hi * fx_lo, w4 = fy_hi * fx_hi;
                gm = w1 * d1 + w2 * d2 + w3 * d3 + w4 * d4;
                gx = m * (fy_lo * (tp.rig * d2 - tp.lef * d1) + fy_hi * (tp.rig * d4 - tp.lef * d3));
                gy = m * (fx_lo * (tp.bot * d3 - tp.top * d1) + fx_hi * (tp.bot * d4 - tp.top * d2));
                if (live) { coef[0] = w1 * m; coef[1] = w2 * m; coef[2] = w3 * m; coef[3] = w4 * m; }
            }
            if constexpr (PAIR) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const size_t mine = img_base + (size_t)at[k];
                    const size_t theirs = (size_t)__shfl_xor_sync(0xffffffffu, (unsigned long long)mine, 1);
                    const float coef_theirs = __shfl_xor_sync(0xffffffffu, coef[k], 1);
                    const float c_even = odd ? coef_theirs : coef[k];   // point owned by the even lane
                    const float c_odd = odd ? coef[k] : coef_theirs;    // point owned by the odd lane
                    const size_t a_even = (odd ? theirs : mine) + odd * 4;
                    const size_t a_odd = (odd ? mine : theirs) + odd * 4;
                    if (c_even != 0.f) {
#pragma unroll
                        for (int r = 0; r < R; r += 2) {
                            float c[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) c[e] = c_even * g_even[(r / 2) * 4 + e];
                            accum.template add<4>(a_even + r * 4, c);
                        }
                    }
                    if (c_odd != 0.f) {
#pragma unroll
                        for (int r = 0; r < R; r += 2) {
                            float c[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) c[e] = c_odd * g_odd[(r / 2) * 4 + e];
                            accum.template add<4>(a_odd + r * 4, c);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (coef[k] != 0.f) {
                        float c[CH];
#pragma unroll
                        for (int v = 0; v < CH; ++v) c[v] = coef[k] * gf[v];
                        accum.template add<CH>(img_base + (size_t)at[k], c);
                    }
                }
            }
            // fold the channel sums across the lanes of this (pixel, group)
            for (int s = thr_per_group >> 1; s > 0; s >>= 1) {
                gm += __shfl_xor_sync(0xffffffffu, gm, s);
                gx += __shfl_xor_sync(0xffffffffu, gx, s);
                gy += __shfl_xor_sync(0xffffffffu, gy, s);
            }
            if (live && part == 0) {
                store_pair(goff + 2 * p, q.sigma * gx, q.sigma * gy);
                gmsk[p] = from_f32<T>(gm);
            }
        }
    }
}

// Generic fallback: one warp per (n,ho,wo,g); lanes stride over the group's channels.
template <typename T, typename Accum>
__global__ void __launch_bounds__(kBwdThreads)
bwd_scatter_warp(const T *__restrict__ value, const T *__restrict__ offset, const T *__restrict__ mask,
                 const T *__restrict__ grad_out, Accum accum, const unsigned *__restrict__ amax_bits,
                 T *__restrict__ grad_offset, T *__restrict__ grad_mask, const Geom q,
                 const long long n_groups) {
    const long long pg = ((long long)blockIdx.x * kBwdThreads + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (pg >= n_groups) return;  // warp-uniform
    const int g = (int)(pg % q.G);
    const long long pix = pg / q.G;
    const int wo = (int)(pix % q.Wo);
    const long long row = pix / q.Wo;
    const int ho = (int)(row % q.Ho);
    const int n = (int)(row / q.Ho);

    if constexpr (sizeof(accum.buf[0]) == 8) accum.scale = fixed_scale(amax_bits);

    const int P = q.kh * q.kw;
    const int C = q.G * q.gc;
    const size_t img_base = (size_t)n * q.H * q.W * C + g * q.gc;
    const T *img = value + img_base;
    const T *gout = grad_out + pg * q.gc;
    const float base_w = axis_base(wo, q.kw, q.sw, q.pw, q.dw, q.sigma);
    const float base_h = axis_base(ho, q.kh, q.sh, q.ph, q.dh, q.sigma);

    for (int i = 0; i < q.kw; ++i)
        for (int j = 0; j < q.kh; ++j) {
            const int p = i * q.kh + j;
            const float dx = to_f32(__ldg(offset + (pg * P + p) * 2));
            const float dy = to_f32(__ldg(offset + (pg * P + p) * 2 + 1));
            const float m = to_f32(__ldg(mask + pg * P + p));
            const float loc_w = base_w + ((float)(i * q.dw) + dx) * q.sigma;
            const float loc_h = base_h + ((float)(j * q.dh) + dy) * q.sigma;
            const Tap tp = make_tap(loc_h, loc_w, q.H, q.W);
            float gm = 0.f, gx = 0.f, gy = 0.f;
            if (tp.inside) {  // warp-uniform
                const ptrdiff_t at = ((ptrdiff_t)tp.h0 * q.W + tp.w0) * C;
                const ptrdiff_t down = (ptrdiff_t)q.W * C;
                const float w1 = tp.hh * tp.hw, w2 = tp.hh * tp.lw, w3 = tp.lh * tp.hw, w4 = tp.lh * tp.lw;
                for (int c = lane; c < q.gc; c += 32) {
                    const float gc_ = to_f32(__ldg(gout + c));
                    float v1[1], v2[1], v3[1], v4[1];
                    ChanVec<T, 1>::load(img + at + c, tp.tl, v1);
                    ChanVec<T, 1>::load(img + at + C + c, tp.tr, v2);
                    ChanVec<T, 1>::load(img + at + down + c, tp.bl, v3);
                    ChanVec<T, 1>::load(img + at + down + C + c, tp.br, v4);
                    gm += gc_ * (w1 * v1[0] + w2 * v2[0] + w3 * v3[0] + w4 * v4[0]);
                    gx += gc_ * m * (tp.hh * (v2[0] - v1[0]) + tp.lh * (v4[0] - v3[0]));
                    gy += gc_ * m * (tp.hw * (v3[0] - v1[0]) + tp.lw * (v4[0] - v2[0]));
                    float t1[1];
                    if (tp.tl) { t1[0] = (w1 * m) * gc_; accum.template add<1>(img_base + at + c, t1); }
                    if (tp.tr) { t1[0] = (w2 * m) * gc_; accum.template add<1>(img_base + at + C + c, t1); }
                    if (tp.bl) { t1[0] = (w3 * m) * gc_; accum.template add<1>(img_base + at + down + c, t1); }
                    if (tp.br) { t1[0] = (w4 * m) * gc_; accum.template add<1>(img_base + at + down + C + c, t1); }
                }
            }
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                gm += __shfl_xor_sync(0xffffffffu, gm, s);
                gx += __shfl_xor_sync(0xffffffffu, gx, s);
                gy += __shfl_xor_sync(0xffffffffu, gy, s);
            }
            if (lane == 0) {
                grad_offset[(pg * P + p) * 2] = from_f32<T>(q.sigma * gx);
                grad_offset[(pg * P + p) * 2 + 1] = from_f32<T>(q.sigma * gy);
                grad_mask[pg * P + p] = from_f32<T>(gm);
            }
        }
}

// ---------------------------------------------------------------------------------------------
static inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

size_t backward_workspace_bytes(const Geom &q, int dtype, unsigned flags) {
    const size_t plane = (size_t)q.N * q.H * q.W * q.G * q.gc;
    if (flags & 1u) return kWorkspaceHeader + plane * sizeof(unsigned long long);
    if (dtype != 0) return kWorkspaceHeader + plane * sizeof(float);
    return 0;
}

int bwd_variant_requested() {
    // development knob: DCNV3_BWD_PAIR=0 disables the sector-coalescing lane-pair exchange
    const char *e = std::getenv("DCNV3_BWD_PAIR");
    return (e && e[0] == '0') ? 0 : 1;
}

template <typename T, int NV, typename Accum, bool PAIR>
static cudaError_t launch_vec(const T *v, const T *o, const T *m, const T *go, Accum accum,
                              const unsigned *amax_bits, T *goff, T *gmsk, const Geom &q,
                              cudaStream_t stream) {
    constexpr int E = Chunk<T>::kElems;
    const unsigned thr_per_group = q.gc / (NV * E);
    const long long per_image = (long long)q.Ho * q.Wo * q.G * thr_per_group;  // < 2^31 (C ABI check)
    const unsigned blocks = (unsigned)((per_image + kBwdThreads - 1) / kBwdThreads);
    for (int n0 = 0; n0 < q.N; n0 += 65535) {  // gridDim.y limit
        const dim3 grid(blocks, (unsigned)std::min(65535, q.N - n0));
        if (q.kh == 3 && q.kw == 3)
            bwd_scatter<T, NV, 3, 3, Accum, PAIR><<<grid, kBwdThreads, 0, stream>>>(
                v, o, m, go, accum, amax_bits, goff, gmsk, q, (unsigned)per_image, thr_per_group, n0);
        else
            bwd_scatter<T, NV, 0, 0, Accum, PAIR><<<grid, kBwdThreads, 0, stream>>>(
                v, o, m, go, accum, amax_bits, goff, gmsk, q, (unsigned)per_image, thr_per_group, n0);
    }
    return cudaGetLastError();
}

template <typename T, typename Accum>
static cudaError_t launch_scatter(const T *v, const T *o, const T *m, const T *go, Accum accum,
                                  const unsigned *amax_bits, T *goff, T *gmsk, const Geom &q,
                                  bool vec_ok, cudaStream_t stream) {
    constexpr int E = Chunk<T>::kElems;
    constexpr bool kF32Accum = sizeof(accum.buf[0]) == 4;
    const long long n_groups = (long long)q.N * q.Ho * q.Wo * q.G;
    if (n_groups == 0) return cudaSuccess;
    const int chunks = q.gc / E;
    // one 16-byte chunk per lane by default (sector-efficient gathers); DCNV3_NV=2: two chunks
    const char *nv_env = std::getenv("DCNV3_NV");
    const int nv = (nv_env && nv_env[0] == '2' && chunks % 2 == 0) ? 2 : 1;
    const int thr_per_group = chunks / nv;
    if (vec_ok && q.gc % E == 0 && is_pow2(thr_per_group) && thr_per_group <= 32) {
        const bool pair = kF32Accum && bwd_variant_requested();
        if (nv == 2) {
            if constexpr (kF32Accum)
                if (pair) return launch_vec<T, 2, Accum, true>(v, o, m, go, accum, amax_bits, goff, gmsk, q, stream);
            return launch_vec<T, 2, Accum, false>(v, o, m, go, accum, amax_bits, goff, gmsk, q, stream);
        }
        if constexpr (kF32Accum && E == 8)
            if (pair) return launch_vec<T, 1, Accum, true>(v, o, m, go, accum, amax_bits, goff, gmsk, q, stream);
        return launch_vec<T, 1, Accum, false>(v, o, m, go, accum, amax_bits, goff, gmsk, q, stream);
    }
    const long long blocks = (n_groups * 32 + kBwdThreads - 1) / kBwdThreads;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    bwd_scatter_warp<T, Accum><<<(unsigned)blocks, kBwdThreads, 0, stream>>>(
        v, o, m, go, accum, amax_bits, goff, gmsk, q, n_groups);
    return cudaGetLastError();
}

// A side stream per device for work that may run beside the main stream inside one call (fork / join
// with events: no host synchronisation, legal under stream capture).  nullptr if creation fails.
struct SideStream {
    cudaStream_t stream;
    cudaEvent_t fork, join;
};
static SideStream *side_stream() {
    constexpr int kMaxDev = 64;
    static SideStream tab[kMaxDev];
    static int state[kMaxDev];   // 0 = not tried, 1 = ready, -1 = failed
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (state[dev] == 0) {
        SideStream s{};
        const bool ok = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) == cudaSuccess &&
                        cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) == cudaSuccess &&
                        cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) == cudaSuccess;
        tab[dev] = s;
        state[dev] = ok ? 1 : -1;
        if (!ok) (void)cudaGetLastError();
    }
    return state[dev] == 1 ? &tab[dev] : nullptr;
}

template <typename T>
static cudaError_t backward_typed(const void *value, const void *offset, const void *mask,
                                  const void *grad_out, void *grad_value, void *grad_offset,
                                  void *grad_mask, void *workspace, const Geom &q, unsigned flags,
                                  cudaStream_t stream) {
    const T *v = static_cast<const T *>(value), *o = static_cast<const T *>(offset),
            *m = static_cast<const T *>(mask), *go = static_cast<const T *>(grad_out);
    T *gv = static_cast<T *>(grad_value), *goff = static_cast<T *>(grad_offset),
      *gmsk = static_cast<T *>(grad_mask);
    const size_t plane = (size_t)q.N * q.H * q.W * q.G * q.gc;
    const size_t n_out = (size_t)q.N * q.Ho * q.Wo * q.G * q.gc;
    const bool vec_ok = (((uintptr_t)value | (uintptr_t)grad_out | (uintptr_t)grad_value) % 16 == 0) &&
                        (((uintptr_t)offset | (uintptr_t)grad_offset) % (2 * sizeof(T)) == 0);
    static const int num_sms = [] {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        return n;
    }();
    const int aux_blocks = num_sms * 8;
    cudaError_t err;
    if (plane == 0) return cudaSuccess;

    if (flags & 1u) {  // deterministic: 64-bit fixed point
        unsigned *amax_bits = static_cast<unsigned *>(workspace);
        auto *acc = reinterpret_cast<unsigned long long *>(static_cast<char *>(workspace) + kWorkspaceHeader);
        if ((err = cudaMemsetAsync(workspace, 0, kWorkspaceHeader + plane * sizeof(unsigned long long), stream)) != cudaSuccess) return err;
        if (n_out) {
            absmax_kernel<T><<<aux_blocks, 256, 0, stream>>>(go, n_out, amax_bits);
            if ((err = cudaGetLastError()) != cudaSuccess) return err;
        }
        if ((err = launch_scatter<T>(v, o, m, go, AccumFixed{acc, 1.f}, amax_bits, goff, gmsk, q, vec_ok, stream)) != cudaSuccess) return err;
        narrow_fixed_kernel<T><<<aux_blocks, 256, 0, stream>>>(acc, amax_bits, gv, plane);
        return cudaGetLastError();
    }
    if constexpr (sizeof(T) == 4) {  // fp32 I/O: accumulate straight into grad_value
        if ((err = cudaMemsetAsync(gv, 0, plane * sizeof(float), stream)) != cudaSuccess) return err;
        if (vec_ok && try_launch_backward_tile(value, offset, mask, grad_out, reinterpret_cast<float *>(gv), grad_offset, grad_mask, q, 0, stream, &err))
            return err;
        return launch_scatter<T>(v, o, m, go, AccumF32{reinterpret_cast<float *>(gv)}, nullptr, goff, gmsk, q, vec_ok, stream);
    } else {  // 16-bit I/O: fp32 scratch plane, then one narrowing pass
        float *acc = reinterpret_cast<float *>(static_cast<char *>(workspace) + kWorkspaceHeader);
        const int dtype_tag = std::is_same<T, __half>::value ? 1 : 2;
        // Default for eligible shapes (gc == 16, G % 8 == 0, 3x3 / stride 1 / dilation 1): the SPLIT backward --
        // channel sums (grad_offset / grad_mask) in the forward's group-slice layout (dcnv3_backward_dots.cu),
        // grad_value as a tcgen05 product with TMEM accumulators (dcnv3_backward_vmma.cu).  The channel-sum
        // kernel also zeroes the value kernel's fp32 plane (a slice per CTA while the CTA waits for its window).
        // DCNV3_BWD=strip selects the fused register-accumulator kernel, DCNV3_VALUE=hmma the HMMA value kernel.
        {
            const char *e = std::getenv("DCNV3_BWD");
            const bool split = !(e && e[0]) || (e[0] == 's' && e[1] == 'p');
            const char *ev = std::getenv("DCNV3_VALUE");
            const bool hmma = ev && ev[0] == 'h';
            // group_channels == 16, maps up to 240 wide: grad_value with the accumulator resident in tensor memory
            // (dcnv3_backward_vres.cu) -- written once in the I/O dtype: no fp32 plane, nothing to zero or to narrow.
            // DCNV3_VALUE=mma keeps the plane form (dcnv3_backward_vmma.cu).
            if (split && vec_ok && !(ev && ev[0]) && backward_vres_preferred(q) && backward_vres_eligible(offset, mask, grad_out, grad_value, q) &&
                backward_vmma_eligible(offset, mask, grad_out, acc, q)) {
                cudaError_t e1 = cudaSuccess, e2 = cudaSuccess;
                // (the channel-sum kernel zeroes the scratch header: the far-point counter lives there)
                auto *counter = static_cast<unsigned long long *>(workspace);
                if (try_launch_backward_dots(value, offset, mask, grad_out, grad_offset, grad_mask, q, dtype_tag, stream, &e1,
                                             static_cast<float *>(workspace), kWorkspaceHeader)) {
                    if (e1 != cudaSuccess) return e1;
                    unsigned long long thr = 0;
                    if (try_launch_backward_vres(offset, mask, grad_out, grad_value, acc, plane * sizeof(float), counter, &thr, q,
                                                 dtype_tag, stream, &e2)) {
                        if (e2 != cudaSuccess) return e2;
                        // More than `thr` points beyond their patch's band (offsets of many pixels): the 16-bit atomics of
                        // the far path would round too often.  The plane form then runs on top and overwrites grad_value;
                        // both kernels read the count on the device and exit at once in the common case.
                        static const bool no_fb = [] { const char *e = std::getenv("DCNV3_VRES_NOFALLBACK"); return e && e[0] == '1'; }();
                        if (no_fb) return cudaSuccess;   // development only: far-heavy inputs then lose accuracy
                        if (!try_launch_backward_vmma(offset, mask, grad_out, acc, q, dtype_tag, stream, &e2, counter, thr))
                            return cudaErrorInvalidConfiguration;
                        if (e2 != cudaSuccess) return e2;
                        if ((err = pdl_launch(true, narrow_f32_kernel<T>, dim3(aux_blocks / 8), dim3(256), 0, stream,
                                              static_cast<const float *>(acc), gv, plane, static_cast<const unsigned long long *>(counter), thr)) != cudaSuccess) return err;
                        return cudaGetLastError();
                    }
                }
            }
            if (split && vec_ok && q.G % 8 == 0 && (hmma || backward_vmma_eligible(offset, mask, grad_out, acc, q))) {
                // The plane is zeroed by the channel-sum kernel itself (a slice per CTA, while the CTA waits for its
                // window): no memset launch, no side stream.  DCNV3_ZERO=side keeps the earlier form (memset on a
                // side stream beside the channel sums, fork / join with events).
                const char *ez = std::getenv("DCNV3_ZERO");
                SideStream *ss = (ez && ez[0] == 's') ? side_stream() : nullptr;
                cudaError_t e1 = cudaSuccess, e2 = cudaSuccess;
                if (ss) {   // fork: memset of the plane || channel sums
                    // (the events are per device, not per call: record + wait must not interleave with another
                    // host thread's; the side stream itself is in order, so a later `join` covers an earlier memset)
                    static std::mutex fork_mu;
                    std::lock_guard<std::mutex> lock(fork_mu);
                    if ((err = cudaEventRecord(ss->fork, stream)) != cudaSuccess) return err;
                    if ((err = cudaStreamWaitEvent(ss->stream, ss->fork, 0)) != cudaSuccess) return err;
                    if ((err = cudaMemsetAsync(acc, 0, plane * sizeof(float), ss->stream)) != cudaSuccess) return err;
                    if ((err = cudaEventRecord(ss->join, ss->stream)) != cudaSuccess) return err;
                }
                const bool dots = try_launch_backward_dots(value, offset, mask, grad_out, grad_offset, grad_mask, q, dtype_tag, stream, &e1,
                                                           ss ? nullptr : acc, ss ? 0 : plane * sizeof(float));
                if (ss) {   // join (also when the channel-sum kernel declined: the plane is needed either way)
                    if ((err = cudaStreamWaitEvent(stream, ss->join, 0)) != cudaSuccess) return err;
                } else if (!dots && (err = cudaMemsetAsync(acc, 0, plane * sizeof(float), stream)) != cudaSuccess) {
                    return err;   // (the channel-sum kernel declined: nobody has zeroed the plane yet)
                }
                if (dots) {
                    if (e1 != cudaSuccess) return e1;
                    // DCNV3_VALUE=band: the dense-band value kernel (experiments/dcnv3_backward_vband.cu, opt-in)
                    const bool band = ev && ev[0] == 'b';
                    if ((band && try_launch_backward_vband(offset, mask, grad_out, acc, q, dtype_tag, stream, &e2)) ||
                        (!hmma && try_launch_backward_vmma(offset, mask, grad_out, acc, q, dtype_tag, stream, &e2)) ||
                        try_launch_backward_vstrip(offset, mask, grad_out, acc, q, dtype_tag, stream, &e2)) {
                        if (e2 != cudaSuccess) return e2;
                        if ((err = pdl_launch(pdl_for(q), narrow_f32_kernel<T>, dim3(aux_blocks), dim3(256), 0, stream,
                                              static_cast<const float *>(acc), gv, plane, static_cast<const unsigned long long *>(nullptr), 0ull)) != cudaSuccess) return err;
                        return cudaGetLastError();
                    }
                }
                // (not reached for eligible shapes) fall through: the fused kernels recompute everything
            } else if ((err = cudaMemsetAsync(acc, 0, plane * sizeof(float), stream)) != cudaSuccess) {
                return err;
            }
        }
        if (!(vec_ok && (try_launch_backward_strip(value, offset, mask, grad_out, acc, grad_offset, grad_mask, q, dtype_tag, stream, &err) ||
                         try_launch_backward_mma2(value, offset, mask, grad_out, acc, grad_offset, grad_mask, q, dtype_tag, stream, &err) ||
                         try_launch_backward_mma(value, offset, mask, grad_out, acc, grad_offset, grad_mask, q, dtype_tag, stream, &err) ||
                         try_launch_backward_tile(value, offset, mask, grad_out, acc, grad_offset, grad_mask, q, dtype_tag, stream, &err))))
            err = launch_scatter<T>(v, o, m, go, AccumF32{acc}, nullptr, goff, gmsk, q, vec_ok, stream);
        if (err != cudaSuccess) return err;
        narrow_f32_kernel<T><<<aux_blocks, 256, 0, stream>>>(acc, gv, plane);
        return cudaGetLastError();
    }
}

cudaError_t launch_backward(const void *value, const void *offset, const void *mask,
                            const void *grad_out, void *grad_value, void *grad_offset,
                            void *grad_mask, void *workspace, const Geom &q, int dtype,
                            unsigned flags, cudaStream_t stream) {
    switch (dtype) {
    case 0: return backward_typed<float>(value, offset, mask, grad_out, grad_value, grad_offset, grad_mask, workspace, q, flags, stream);
    case 1: return backward_typed<__half>(value, offset, mask, grad_out, grad_value, grad_offset, grad_mask, workspace, q, flags, stream);
    default: return backward_typed<__nv_bfloat16>(value, offset, mask, grad_out, grad_value, grad_offset, grad_mask, workspace, q, flags, stream);
    }
}

}  // namespace dcnv3
