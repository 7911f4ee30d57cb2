// dcnv3_dwconv.cu -- the producer of the layer's x1: depthwise k x k
// convolution + LayerNorm over the channels + GELU, channels-last in and out, one pass
// (reference: models/ops_dcnv3/modules/dcnv3.py:276-289,328-329 with build_norm_layer :41-62 --
// NHWC -> permute -> Conv2d(groups=C) in NCHW -> permute back -> LayerNorm -> GELU: five kernels
// and two layout changes, every one a full read + write of the activation tensor).
//
// A CTA owns an 8 x 8 tile of pixels: one TMA box brings the (8+2r) x (8+2r) x C input window into
// shared memory (out-of-map pixels zero-filled = the convolution's zero padding); a group of C/8
// lanes owns a pixel (8 channels = 16 bytes per lane): k*k taps of 128-bit shared loads, fp32
// accumulation, mean / variance by xor-shuffles inside the lane group (two-pass: exact variance of
// the values in registers), exact erf GELU, one 128-bit store.  HBM traffic = read x once, write x1
// once -- the kernel is an HBM-roofline kernel (algorithmic bytes 2 s N H W C).
#include "dcnv3_sm100.h"

#include "dcnv3_launch.h"
#include "dcnv3_tma.cuh"

#include <algorithm>
#include <cstdlib>

namespace dcnv3 {
namespace dwc {

constexpr int kTile = 8;
constexpr int kThreads = 256;

struct Params {
    int N, H, W, C, k, r;
    int tiles_x, tiles_y;
    float eps, inv_c;
};

// exact-form GELU 0.5 x (1 + erf(x / sqrt 2)); erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far
// below the 16-bit output rounding) -- one reciprocal, one exp2 and a degree-5 Horner instead of erff()
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    float t, ex;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-1.4426950408889634f * z * z));
    const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
    const float erf_abs = fmaf(-poly, ex, 1.f);
    return 0.5f * x * (1.f + copysignf(erf_abs, x));
}

// acc[e] += x[e] * w[e] for 8 packed 16-bit pairs (exact products, fp32 accumulation)
template <typename T> __device__ __forceinline__ void fma8(float (&acc)[8], const uint4 &x, const uint4 &w) {
    const uint32_t a[4] = {x.x, x.y, x.z, x.w}, b[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        acc[2 * i] = mix_fma(lo16(a[i]), lo16(b[i]), acc[2 * i], T());
        acc[2 * i + 1] = mix_fma(hi16(a[i]), hi16(b[i]), acc[2 * i + 1], T());
    }
}

template <typename T, int K /* compile-time kernel size (weights live in registers) or 0 */, int CPP /* lanes per pixel = C / 8 */>
__global__ void __launch_bounds__(kThreads)
dwconv_ln_gelu(const __grid_constant__ CUtensorMap tmap, const T *__restrict__ wdw /* [k*k][C], I/O dtype */,
               const float *__restrict__ bdw, const float *__restrict__ gamma, const float *__restrict__ beta,
               T *__restrict__ out, T *__restrict__ conv_out /* pre-LayerNorm values for the backward, or null */,
               const Params pp) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    const int win = kTile + 2 * pp.r;
    constexpr int cpp = CPP;                      // lanes per pixel (8, 16 or 32)
    const int tile = blockIdx.x, tx = tile % pp.tiles_x, ty = tile / pp.tiles_x, n = blockIdx.y;
    const int x0 = tx * kTile, y0 = ty * kTile;
    const unsigned win_bytes = (unsigned)(win * win * pp.C * 2);

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        mbar_expect_tx(&bar, win_bytes);
        tma_load_4d(smem, &tmap, &bar, 0, x0 - pp.r, y0 - pp.r, n);
    }
    // per-lane constants while the box is in flight: this lane's 8 channels of every parameter
    const int cl = tid % cpp;                      // 16-byte chunk of the pixel this lane owns
    const int ch0 = cl * 8;
    float g8[8], b8[8], cb8[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { g8[e] = gamma[ch0 + e]; b8[e] = beta[ch0 + e]; cb8[e] = bdw[ch0 + e]; }
    // a lane's channels are fixed: its taps stay in registers, packed in the I/O dtype (the
    // reference's 16-bit conv multiplies 16-bit weights too); products are exact, sums fp32 (FHFMA)
    uint4 wr[K > 0 ? K * K : 1];
    if constexpr (K > 0) {
#pragma unroll
        for (int t = 0; t < K * K; ++t) wr[t] = __ldg(reinterpret_cast<const uint4 *>(wdw + (size_t)t * pp.C + ch0));
    }
    const uint32_t smem_lane = smem_u32(smem) + ch0 * 2;
    const uint32_t pix_b = pp.C * 2, row_b = win * pix_b;
    __syncthreads();
    mbar_wait(&bar, 0);

    const int ppw = kThreads / cpp;                // pixels processed per pass by the CTA
    for (int p = tid / cpp; p < kTile * kTile; p += ppw) {
        const int px = p % kTile, py = p / kTile;
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = cb8[e];
        const uint32_t base = smem_lane + py * row_b + px * pix_b;
        if constexpr (K > 0) {
#pragma unroll
            for (int j = 0; j < K; ++j)
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    uint4 v;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + j * row_b + i * pix_b));
                    fma8<T>(acc, v, wr[j * K + i]);
                }
        } else {
            for (int j = 0; j < pp.k; ++j) {
                for (int i = 0; i < pp.k; ++i) {
                    uint4 v;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + j * row_b + i * pix_b));
                    fma8<T>(acc, v, __ldg(reinterpret_cast<const uint4 *>(wdw + (size_t)(j * pp.k + i) * pp.C + ch0)));
                }
            }
        }
        // LayerNorm over the C channels of the pixel = over the cpp lanes of this group
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) s += acc[e];
#pragma unroll
        for (int o = cpp >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * pp.inv_c;
        float q = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float d = acc[e] - mean; q += d * d; }
#pragma unroll
        for (int o = cpp >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(fmaf(q, pp.inv_c, pp.eps));
        float y[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = gelu_erf((acc[e] - mean) * rstd * g8[e] + b8[e]);
        const int ox = x0 + px, oy = y0 + py;
        if (ox < pp.W && oy < pp.H) {
            const size_t at = (((size_t)n * pp.H + oy) * pp.W + ox) * pp.C + ch0;
            *reinterpret_cast<uint4 *>(out + at) = pack<T>(y);
            if (conv_out) *reinterpret_cast<uint4 *>(conv_out + at) = pack<T>(acc);
        }
    }
}

// The same for the compile-time kernel sizes as a PERSISTENT kernel: one wave of CTAs, each walking a strided list of
// tiles with two window buffers -- the TMA box of the next tile is in flight while this one is computed.  (The one-shot
// form runs 1600 CTAs of 51 KB at two per SM for cfg2: 5.4 waves, and every CTA starts by waiting for its own box.)
template <typename T, int K, int CPP>
__global__ void __launch_bounds__(kThreads)
dwconv_ln_gelu_persistent(const __grid_constant__ CUtensorMap tmap, const T *__restrict__ wdw, const float *__restrict__ bdw,
                          const float *__restrict__ gamma, const float *__restrict__ beta, T *__restrict__ out,
                          T *__restrict__ conv_out, const Params pp) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar[2];
    const int tid = threadIdx.x;
    const int win = kTile + 2 * pp.r;
    constexpr int cpp = CPP;
    const unsigned win_bytes = (unsigned)(win * win * pp.C * 2);
    const int per_image = pp.tiles_x * pp.tiles_y, total = per_image * pp.N;
    int t = blockIdx.x;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
        if (t < total) {
            const int n = t / per_image, r = t % per_image;
            mbar_expect_tx(&bar[0], win_bytes);
            tma_load_4d(smem, &tmap, &bar[0], 0, (r % pp.tiles_x) * kTile - pp.r, (r / pp.tiles_x) * kTile - pp.r, n);
        }
    }
    const int cl = tid % cpp, ch0 = cl * 8;
    float g8[8], b8[8], cb8[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { g8[e] = gamma[ch0 + e]; b8[e] = beta[ch0 + e]; cb8[e] = bdw[ch0 + e]; }
    uint4 wr[K * K];
#pragma unroll
    for (int i = 0; i < K * K; ++i) wr[i] = __ldg(reinterpret_cast<const uint4 *>(wdw + (size_t)i * pp.C + ch0));
    const uint32_t pix_b = pp.C * 2, row_b = win * pix_b;
    __syncthreads();
    unsigned phases = 0u;                      // bit s: the parity buffer s completes next
    const int ppw = kThreads / cpp;
    for (int stage = 0; t < total; t += gridDim.x, stage ^= 1) {
        const int tn = t + gridDim.x;
        if (tid == 0 && tn < total) {          // the other buffer was released by the barrier that ended the last pass
            const int n2 = tn / per_image, r2 = tn % per_image;
            mbar_expect_tx(&bar[stage ^ 1], win_bytes);
            tma_load_4d(smem + (size_t)(stage ^ 1) * win_bytes, &tmap, &bar[stage ^ 1], 0, (r2 % pp.tiles_x) * kTile - pp.r,
                        (r2 / pp.tiles_x) * kTile - pp.r, n2);
        }
        mbar_wait(&bar[stage], (phases >> stage) & 1u);
        phases ^= 1u << stage;
        const int n = t / per_image, r = t % per_image;
        const int x0 = (r % pp.tiles_x) * kTile, y0 = (r / pp.tiles_x) * kTile;
        const uint32_t smem_lane = smem_u32(smem) + stage * win_bytes + ch0 * 2;
        for (int p = tid / cpp; p < kTile * kTile; p += ppw) {
            const int px = p % kTile, py = p / kTile;
            float acc[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = cb8[e];
            const uint32_t base = smem_lane + py * row_b + px * pix_b;
#pragma unroll
            for (int j = 0; j < K; ++j)
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    uint4 v;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + j * row_b + i * pix_b));
                    fma8<T>(acc, v, wr[j * K + i]);
                }
            float s = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) s += acc[e];
#pragma unroll
            for (int o = cpp >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mean = s * pp.inv_c;
            float q = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) { const float d = acc[e] - mean; q += d * d; }
#pragma unroll
            for (int o = cpp >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            const float rstd = rsqrtf(fmaf(q, pp.inv_c, pp.eps));
            float y[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] = gelu_erf((acc[e] - mean) * rstd * g8[e] + b8[e]);
            const int ox = x0 + px, oy = y0 + py;
            if (ox < pp.W && oy < pp.H) {
                const size_t at = (((size_t)n * pp.H + oy) * pp.W + ox) * pp.C + ch0;
                *reinterpret_cast<uint4 *>(out + at) = pack<T>(y);
                if (conv_out) *reinterpret_cast<uint4 *>(conv_out + at) = pack<T>(acc);
            }
        }
        __syncthreads();                        // everybody is done with this buffer: the next pass may refill it
    }
}

template <typename T>
static int launch(const void *x, const void *wdw_v, const float *bdw, const float *gamma, const float *beta, void *out, void *conv_out,
                  int N, int H, int W, int C, int k, float eps, int dtype, cudaStream_t stream) {
    const T *wdw = static_cast<const T *>(wdw_v);
    Params pp;
    pp.N = N; pp.H = H; pp.W = W; pp.C = C; pp.k = k; pp.r = (k - 1) / 2; pp.eps = eps; pp.inv_c = 1.f / (float)C;
    pp.tiles_x = (W + kTile - 1) / kTile;
    pp.tiles_y = (H + kTile - 1) / kTile;
    const int win = kTile + 2 * pp.r;
    const size_t smem = (size_t)win * win * C * 2;
    if (smem > 200 * 1024 || win > 256 || N > 65535) return DCNV3_E_SHAPE;
    CUtensorMap tmap;
    if (!make_nhwc_tensor_map(&tmap, x, dtype, N, H, W, C, C, win, win)) return DCNV3_E_SHAPE;
    const dim3 grid(pp.tiles_x * pp.tiles_y, N);
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, kThreads, smem, stream>>>(tmap, wdw, bdw, gamma, beta, static_cast<T *>(out), static_cast<T *>(conv_out), pp);
    };
    // k == 3 and two windows within half an SM's shared memory: the persistent, double-buffered form
    const char *eo = std::getenv("DCNV3_DWCONV_ONESHOT");      // (read per call: the tests compare the two forms)
    const bool one_shot = eo && eo[0] == '1';
    if (k == 3 && 2 * smem <= 100 * 1024 + 8 * 1024 && !one_shot) {
        static int num_sms = 0;
        if (num_sms == 0) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        }
        const long long total = (long long)pp.tiles_x * pp.tiles_y * N;
        auto gp = [&](auto kern) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * smem));
            int occ = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, 2 * smem) != cudaSuccess || occ < 1) occ = 1;
            const int ctas = (int)std::min<long long>(total, (long long)num_sms * occ);
            kern<<<ctas, kThreads, 2 * smem, stream>>>(tmap, wdw, bdw, gamma, beta, static_cast<T *>(out), static_cast<T *>(conv_out), pp);
        };
        if (C == 256) gp(dwconv_ln_gelu_persistent<T, 3, 32>); else if (C == 128) gp(dwconv_ln_gelu_persistent<T, 3, 16>);
        else gp(dwconv_ln_gelu_persistent<T, 3, 8>);
        return (int)cudaGetLastError();
    }
    if (k == 3) {
        if (C == 256) go(dwconv_ln_gelu<T, 3, 32>); else if (C == 128) go(dwconv_ln_gelu<T, 3, 16>); else go(dwconv_ln_gelu<T, 3, 8>);
    } else {
        if (C == 256) go(dwconv_ln_gelu<T, 0, 32>); else if (C == 128) go(dwconv_ln_gelu<T, 0, 16>); else go(dwconv_ln_gelu<T, 0, 8>);
    }
    return (int)cudaGetLastError();
}

}  // namespace dwc
}  // namespace dcnv3

extern "C" int dcnv3_dwconv_ln_gelu_sm100(const void *x, const void *w_dw, const float *b_dw, const float *gamma,
                                          const float *beta, void *out, void *conv_out, int N, int H, int W, int C, int k, float eps,
                                          int dtype, void *stream) {
    if (dtype != DCNV3_F16 && dtype != DCNV3_BF16) return DCNV3_E_DTYPE;
    if (N < 0 || H <= 0 || W <= 0 || k <= 0 || k % 2 == 0 || k > 7 || !(C == 64 || C == 128 || C == 256)) return DCNV3_E_SHAPE;
    if (N == 0) return DCNV3_OK;
    if (!x || !w_dw || !b_dw || !gamma || !beta || !out) return DCNV3_E_NULL;
    if (((uintptr_t)x | (uintptr_t)out | (uintptr_t)w_dw | (uintptr_t)conv_out) % 16) return DCNV3_E_ALIGN;
    if (dtype == DCNV3_F16)
        return dcnv3::dwc::launch<__half>(x, w_dw, b_dw, gamma, beta, out, conv_out, N, H, W, C, k, eps, dtype, (cudaStream_t)stream);
    return dcnv3::dwc::launch<__nv_bfloat16>(x, w_dw, b_dw, gamma, beta, out, conv_out, N, H, W, C, k, eps, dtype, (cudaStream_t)stream);
}
