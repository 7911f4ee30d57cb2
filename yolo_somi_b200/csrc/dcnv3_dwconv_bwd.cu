// dcnv3_dwconv_bwd.cu -- backward of the layer's x1 producer (depthwise 3 x 3 convolution + LayerNorm + GELU,
// reference models/ops_dcnv3/modules/dcnv3.py:276-289,328-329) and of the mask soft-max (:331-334), channels-last.
//
// Round 1 left these to PyTorch: LayerNorm + GELU re-run under autograd on the saved convolution output, then
// aten.convolution_backward on permuted views -- per layer call at cfg2 (N = 16, 80 x 80, C = 256, bf16) that was
// 153 us LayerNorm forward + 90 us LayerNorm grad + 114 us dim-0 reductions + 93 + 49 us depthwise dgrad / wgrad +
// 43 us GELU backward + 85 us layout copies + 93 us elementwise glue (profiles/README.md, r2 layer profile).
// Here it is two passes over the activation tensor:
//   ln_gelu_bwd   per pixel (a group of C / 8 lanes, 8 channels = 16 bytes per lane, as the forward): the LayerNorm
//                 statistics from the saved pre-norm values, GELU' and the LayerNorm input gradient du in one go,
//                 du written once in the I/O dtype; the three channel sums (grad_gamma, grad_beta, grad_bias of the
//                 convolution) are kept in registers over the CTA's pixels, combined in shared memory and leave as
//                 one fp32 reduction per channel and CTA;
//   dwconv_bwd    per 8 x 8 tile: du and x windows by TMA (zero fill = the convolution's padding), grad_x = du
//                 correlated with the flipped taps and the tile's share of grad_w = sum_p du[p] x[p + tap] with exact
//                 FHFMA products, 72 fp32 sums per lane, combined across the CTA in the (then dead) window memory.
// Algorithmic bytes: 3 + 3 tensor passes (312 MB at cfg2) against ~14 before.
#include "dcnv3_sm100.h"

#include "dcnv3_launch.h"
#include "dcnv3_tma.cuh"

#include <algorithm>

namespace dcnv3 {
namespace dwcb {

constexpr int kTile = 8;
constexpr int kThreads = 256;

struct Params {
    int N, H, W, C;
    int tiles_x, tiles_y;
    long long pixels;
    float eps, inv_c;
};

// erf by Abramowitz-Stegun 7.1.26 (as the forward), and the Gaussian density for GELU'
__device__ __forceinline__ void gelu_parts(float v, float &cdf, float &pdf) {
    const float z = fabsf(v) * 0.70710678118654752f;
    float t, ex;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-1.4426950408889634f * z * z));
    const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
    const float erf_abs = fmaf(-poly, ex, 1.f);
    cdf = 0.5f * (1.f + copysignf(erf_abs, v));
    pdf = 0.3989422804014327f * ex;                       // exp(-v^2 / 2) / sqrt(2 pi)
}

__device__ __forceinline__ void red_add(float *p, float v) { atomicAdd(p, v); }

// ------------------------------------------------------------------------------------------------ pass 1
template <typename T, int CPP /* lanes per pixel = C / 8 */>
__global__ void __launch_bounds__(kThreads)
ln_gelu_bwd(const T *__restrict__ conv_out, const T *__restrict__ grad_out, const float *__restrict__ gamma,
            const float *__restrict__ beta, T *__restrict__ du_out, float *__restrict__ g_bias, float *__restrict__ g_gamma,
            float *__restrict__ g_beta, const Params pp) {
    constexpr int kGroups = kThreads / CPP;                // pixels per pass of the CTA
    __shared__ float part[3][kGroups][CPP * 8];
    const int tid = threadIdx.x, cl = tid % CPP, grp = tid / CPP, ch0 = cl * 8;
    float g8[8], b8[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { g8[e] = gamma[ch0 + e]; b8[e] = beta[ch0 + e]; }
    float s_b[8] = {}, s_g[8] = {}, s_t[8] = {};
    const long long stride = (long long)gridDim.x * kGroups;
    // (the whole CTA iterates together: the shuffles below always find every lane of the warp)
    // The next pass's two 128-bit loads are issued before this pass's arithmetic (one pixel per lane group and pass,
    // ~330 dependent instructions behind two loads: without the prefetch the kernel waited on them -- 4.8 warps per
    // issue on the long scoreboard, issue-active 49 %, profiles/r2_dwconv_family_ncu_summary.txt).
    uint4 nu = make_uint4(0u, 0u, 0u, 0u), ng = nu;
    {
        const long long p0 = (long long)blockIdx.x * kGroups + grp;
        if ((long long)blockIdx.x * kGroups < pp.pixels) {
            const size_t a0 = (size_t)(p0 < pp.pixels ? p0 : 0) * pp.C + ch0;
            nu = __ldg(reinterpret_cast<const uint4 *>(conv_out + a0));
            if (p0 < pp.pixels) ng = __ldg(reinterpret_cast<const uint4 *>(grad_out + a0));
        }
    }
    for (long long base = (long long)blockIdx.x * kGroups; base < pp.pixels; base += stride) {
        const long long p = base + grp;
        const bool live = p < pp.pixels;
        const size_t at = (size_t)(live ? p : 0) * pp.C + ch0;
        float u[8], gy[8];
        unpack<T>(nu, u);
        unpack<T>(ng, gy);
        if (base + stride < pp.pixels) {
            const long long p2 = base + stride + grp;
            const size_t a2 = (size_t)(p2 < pp.pixels ? p2 : 0) * pp.C + ch0;
            nu = __ldg(reinterpret_cast<const uint4 *>(conv_out + a2));
            ng = p2 < pp.pixels ? __ldg(reinterpret_cast<const uint4 *>(grad_out + a2)) : make_uint4(0u, 0u, 0u, 0u);
        }
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) s += u[e];
#pragma unroll
        for (int o = CPP >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * pp.inv_c;
        float q = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float d = u[e] - mean; q += d * d; }
#pragma unroll
        for (int o = CPP >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(fmaf(q, pp.inv_c, pp.eps));
        float xh[8], dx[8], s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            xh[e] = (u[e] - mean) * rstd;
            const float v = fmaf(xh[e], g8[e], b8[e]);
            float cdf, pdf;
            gelu_parts(v, cdf, pdf);
            const float dv = gy[e] * fmaf(v, pdf, cdf);    // GELU'(v) = Phi(v) + v phi(v)
            s_g[e] = fmaf(dv, xh[e], s_g[e]);
            s_t[e] += dv;
            dx[e] = dv * g8[e];
            s1 += dx[e];
            s2 = fmaf(dx[e], xh[e], s2);
        }
#pragma unroll
        for (int o = CPP >> 1; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        s1 *= pp.inv_c; s2 *= pp.inv_c;
        float du[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            du[e] = rstd * (dx[e] - s1 - xh[e] * s2);
            s_b[e] += du[e];
        }
        if (live) *reinterpret_cast<uint4 *>(du_out + at) = pack<T>(du);
    }
    // the CTA's channel sums: one reduction per channel
#pragma unroll
    for (int e = 0; e < 8; ++e) { part[0][grp][ch0 + e] = s_b[e]; part[1][grp][ch0 + e] = s_g[e]; part[2][grp][ch0 + e] = s_t[e]; }
    __syncthreads();
    for (int i = tid; i < 3 * CPP * 8; i += kThreads) {
        const int which = i / (CPP * 8), c = i % (CPP * 8);
        float acc = 0.f;
#pragma unroll
        for (int g = 0; g < kGroups; ++g) acc += part[which][g][c];
        red_add((which == 0 ? g_bias : which == 1 ? g_gamma : g_beta) + c, acc);
    }
}

// ------------------------------------------------------------------------------------------------ pass 2
// acc[e] += a[e] * b[e] for 8 packed 16-bit pairs (exact products, fp32 accumulation)
template <typename T> __device__ __forceinline__ void fma8(float (&acc)[8], const uint4 &x, const uint4 &w) {
    const uint32_t a[4] = {x.x, x.y, x.z, x.w}, b[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        acc[2 * i] = mix_fma(lo16(a[i]), lo16(b[i]), acc[2 * i], T());
        acc[2 * i + 1] = mix_fma(hi16(a[i]), hi16(b[i]), acc[2 * i + 1], T());
    }
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

template <typename T, int CPP>
__global__ void __launch_bounds__(kThreads, 2)
dwconv_bwd(const __grid_constant__ CUtensorMap tmap_du, const __grid_constant__ CUtensorMap tmap_x,
           const T *__restrict__ wdw /* [9][C], I/O dtype */, T *__restrict__ grad_x, float *__restrict__ grad_w /* [9][C] */,
           const Params pp) {
    constexpr int K = 3, R = 1, kWin = kTile + 2 * R, kGroups = kThreads / CPP, C = CPP * 8;
    constexpr unsigned kWinBytes = kWin * kWin * C * 2;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, cl = tid % CPP, grp = tid / CPP, ch0 = cl * 8;
    const int tile = blockIdx.x, tx = tile % pp.tiles_x, ty = tile / pp.tiles_x, n = blockIdx.y;
    const int x0 = tx * kTile, y0 = ty * kTile;
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        mbar_expect_tx(&bar, 2 * kWinBytes);
        tma_load_4d(smem, &tmap_du, &bar, 0, x0 - R, y0 - R, n);
        tma_load_4d(smem + kWinBytes, &tmap_x, &bar, 0, x0 - R, y0 - R, n);
    }
    // the nine taps of every channel behind the windows (registers are for the 72 grad_w sums)
    for (int i = tid; i < K * K * C / 8; i += kThreads)
        reinterpret_cast<uint4 *>(smem + 2 * kWinBytes)[i] = __ldg(reinterpret_cast<const uint4 *>(wdw) + i);
    const uint32_t du_lane = smem_u32(smem) + ch0 * 2, x_lane = du_lane + kWinBytes, w_lane = du_lane + 2 * kWinBytes;
    constexpr uint32_t pix_b = C * 2, row_b = kWin * pix_b;
    float gw[K * K][8];
#pragma unroll
    for (int t = 0; t < K * K; ++t)
#pragma unroll
        for (int e = 0; e < 8; ++e) gw[t][e] = 0.f;
    __syncthreads();
    mbar_wait(&bar, 0);

    for (int p = grp; p < kTile * kTile; p += kGroups) {
        const int px = p % kTile, py = p / kTile;
        // window coordinates of this pixel: (py + R, px + R)
        const uint32_t ctr = (py + R) * row_b + (px + R) * pix_b;
        const uint4 du_c = lds128(du_lane + ctr);
        float gx[8] = {};
#pragma unroll
        for (int j = 0; j < K; ++j)
#pragma unroll
            for (int i = 0; i < K; ++i) {
                // forward: conv[p] = sum w[j][i] x[p + (j - R, i - R)]  =>  grad_x[q] = sum w[j][i] du[q - (j - R, i - R)]
                fma8<T>(gx, lds128(du_lane + (py + 2 * R - j) * row_b + (px + 2 * R - i) * pix_b), lds128(w_lane + (j * K + i) * pix_b));
                // grad_w[j][i] += du[p] x[p + (j - R, i - R)]
                fma8<T>(gw[j * K + i], lds128(x_lane + (py + j) * row_b + (px + i) * pix_b), du_c);
            }
        const int ox = x0 + px, oy = y0 + py;
        if (ox < pp.W && oy < pp.H)
            *reinterpret_cast<uint4 *>(grad_x + (((size_t)n * pp.H + oy) * pp.W + ox) * C + ch0) = pack<T>(gx);
    }
    // the tile's share of grad_w: combine the pixel groups in the window memory (dead now), one reduction per element
    __syncthreads();
    float *part = reinterpret_cast<float *>(smem);          // [kGroups][9 * C]
#pragma unroll
    for (int t = 0; t < K * K; ++t)
#pragma unroll
        for (int e = 0; e < 8; ++e) part[(grp * K * K + t) * C + ch0 + e] = gw[t][e];
    __syncthreads();
    for (int i = tid; i < K * K * C; i += kThreads) {
        float acc = 0.f;
#pragma unroll
        for (int g = 0; g < kGroups; ++g) acc += part[g * K * K * C + i];
        red_add(grad_w + i, acc);
    }
}

template <typename T>
static int launch(const void *x, const void *conv_out, const void *grad_out, const void *wdw, const float *gamma,
                  const float *beta, void *du, void *grad_x, float *grad_params, int N, int H, int W, int C, float eps, int dtype,
                  cudaStream_t stream) {
    Params pp;
    pp.N = N; pp.H = H; pp.W = W; pp.C = C; pp.eps = eps; pp.inv_c = 1.f / (float)C;
    pp.tiles_x = (W + kTile - 1) / kTile;
    pp.tiles_y = (H + kTile - 1) / kTile;
    pp.pixels = (long long)N * H * W;
    if (N > 65535) return DCNV3_E_SHAPE;
    float *g_w = grad_params, *g_b = grad_params + 9 * C, *g_gamma = g_b + C, *g_beta = g_gamma + C;
    cudaError_t e = cudaMemsetAsync(grad_params, 0, (size_t)12 * C * sizeof(float), stream);
    if (e != cudaSuccess) return (int)e;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int cpp = C / 8, groups = kThreads / cpp;
    const long long want = (pp.pixels + groups - 1) / groups;
    // one wave of resident CTAs, each striding over the pixels (148 x 8 CTAs at 3 resident per SM ran 2.7 waves)
    static int resident[3] = {0, 0, 0};
    const int ri = C == 256 ? 0 : C == 128 ? 1 : 2;
    if (resident[ri] == 0) {
        int occ = 0;
        const void *fn = C == 256 ? (const void *)ln_gelu_bwd<T, 32> : C == 128 ? (const void *)ln_gelu_bwd<T, 16> : (const void *)ln_gelu_bwd<T, 8>;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kThreads, 0) != cudaSuccess || occ < 1) occ = 2;
        resident[ri] = occ;
    }
    const int ctas1 = (int)std::min<long long>(want, (long long)num_sms * resident[ri]);
    const T *co = static_cast<const T *>(conv_out), *go = static_cast<const T *>(grad_out);
    T *dup = static_cast<T *>(du);
    if (C == 256) ln_gelu_bwd<T, 32><<<ctas1, kThreads, 0, stream>>>(co, go, gamma, beta, dup, g_b, g_gamma, g_beta, pp);
    else if (C == 128) ln_gelu_bwd<T, 16><<<ctas1, kThreads, 0, stream>>>(co, go, gamma, beta, dup, g_b, g_gamma, g_beta, pp);
    else ln_gelu_bwd<T, 8><<<ctas1, kThreads, 0, stream>>>(co, go, gamma, beta, dup, g_b, g_gamma, g_beta, pp);
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;

    constexpr int kWin = kTile + 2;
    CUtensorMap tdu, tx;
    if (!make_nhwc_tensor_map(&tdu, du, dtype, N, H, W, C, C, kWin, kWin)) return DCNV3_E_SHAPE;
    if (!make_nhwc_tensor_map(&tx, x, dtype, N, H, W, C, C, kWin, kWin)) return DCNV3_E_SHAPE;
    const size_t smem = std::max((size_t)2 * kWin * kWin * C * 2 + (size_t)9 * C * 2, (size_t)groups * 9 * C * sizeof(float));
    const dim3 grid(pp.tiles_x * pp.tiles_y, N);
    auto go2 = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, kThreads, smem, stream>>>(tdu, tx, static_cast<const T *>(wdw), static_cast<T *>(grad_x), g_w, pp);
    };
    if (C == 256) go2(dwconv_bwd<T, 32>); else if (C == 128) go2(dwconv_bwd<T, 16>); else go2(dwconv_bwd<T, 8>);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ mask soft-max
// grad of the logits from grad of the soft-maxed masks: gl_p = m_p (gm_p - sum_q gm_q m_q) per (pixel, group),
// fp32 arithmetic, one pass (the reference's autograd runs the soft-max backward in fp32 on up-cast copies).
template <typename T>
__global__ void __launch_bounds__(256)
mask_softmax_bwd(const T *__restrict__ g_mask, const T *__restrict__ mask, T *__restrict__ g_logit, long long rows, int P) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const T *gm = g_mask + r * P, *mk = mask + r * P;
    float dot = 0.f;
    for (int p = 0; p < P; ++p) dot = fmaf(to_f32(gm[p]), to_f32(mk[p]), dot);
    T *gl = g_logit + r * P;
    for (int p = 0; p < P; ++p) gl[p] = from_f32<T>(to_f32(mk[p]) * (to_f32(gm[p]) - dot));
}

// The same for P == 9 with 128-bit accesses: a thread owns EIGHT consecutive rows = 72 elements = nine 16-byte words of
// each tensor (the scalar form issues 27 two-byte accesses per row and stalls on the load / store queue: 45 us for
// 2 x 29.5 MB in, 29.5 MB out at cfg2 against ~15 us of HBM time; profiles/r2_dwconv_family_ncu_summary.txt).  Same
// operation order per row as above, so the results are bit-identical.
template <typename T>
__global__ void __launch_bounds__(128)
mask_softmax_bwd_p9x8(const uint4 *__restrict__ g_mask, const uint4 *__restrict__ mask, uint4 *__restrict__ g_logit, long long chunks) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= chunks) return;
    uint32_t gw[36], mw[36];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const uint4 a = __ldcs(g_mask + c * 9 + i), b = __ldg(mask + c * 9 + i);
        gw[4 * i] = a.x; gw[4 * i + 1] = a.y; gw[4 * i + 2] = a.z; gw[4 * i + 3] = a.w;
        mw[4 * i] = b.x; mw[4 * i + 1] = b.y; mw[4 * i + 2] = b.z; mw[4 * i + 3] = b.w;
    }
    float dot[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float d = 0.f;
#pragma unroll
        for (int p = 0; p < 9; ++p) {
            const int e = 9 * j + p;
            const float2 g2 = unpack2(gw[e >> 1], T()), m2 = unpack2(mw[e >> 1], T());
            d = fmaf((e & 1) ? g2.y : g2.x, (e & 1) ? m2.y : m2.x, d);
        }
        dot[j] = d;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int w = 4 * i + k, e0 = 2 * w, e1 = 2 * w + 1;
            const float2 g2 = unpack2(gw[w], T()), m2 = unpack2(mw[w], T());
            o[k] = pack2(m2.x * (g2.x - dot[e0 / 9]), m2.y * (g2.y - dot[e1 / 9]), T());
        }
        g_logit[c * 9 + i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

}  // namespace dwcb
}  // namespace dcnv3

extern "C" int dcnv3_dwconv_ln_gelu_backward_sm100(const void *x, const void *conv_out, const void *grad_out, const void *w_dw,
                                                   const float *gamma, const float *beta, void *du_scratch, void *grad_x,
                                                   float *grad_params, int N, int H, int W, int C, int k, float eps, int dtype,
                                                   void *stream) {
    if (dtype != DCNV3_F16 && dtype != DCNV3_BF16) return DCNV3_E_DTYPE;
    if (N < 0 || H <= 0 || W <= 0 || k != 3 || !(C == 64 || C == 128 || C == 256)) return DCNV3_E_SHAPE;
    if (N == 0) return DCNV3_OK;
    if (!x || !conv_out || !grad_out || !w_dw || !gamma || !beta || !du_scratch || !grad_x || !grad_params) return DCNV3_E_NULL;
    if (((uintptr_t)x | (uintptr_t)conv_out | (uintptr_t)grad_out | (uintptr_t)w_dw | (uintptr_t)du_scratch | (uintptr_t)grad_x) % 16)
        return DCNV3_E_ALIGN;
    if (dtype == DCNV3_F16)
        return dcnv3::dwcb::launch<__half>(x, conv_out, grad_out, w_dw, gamma, beta, du_scratch, grad_x, grad_params, N, H, W, C, eps,
                                           dtype, (cudaStream_t)stream);
    return dcnv3::dwcb::launch<__nv_bfloat16>(x, conv_out, grad_out, w_dw, gamma, beta, du_scratch, grad_x, grad_params, N, H, W, C, eps,
                                              dtype, (cudaStream_t)stream);
}

extern "C" int dcnv3_mask_softmax_backward_sm100(const void *grad_mask, const void *mask, void *grad_logit, long long rows,
                                                 int points, int dtype, void *stream) {
    if (dtype != DCNV3_F16 && dtype != DCNV3_BF16) return DCNV3_E_DTYPE;
    if (rows < 0 || points <= 0) return DCNV3_E_SHAPE;
    if (rows == 0) return DCNV3_OK;
    if (!grad_mask || !mask || !grad_logit) return DCNV3_E_NULL;
    if (points == 9 && rows % 8 == 0 && (((uintptr_t)grad_mask | (uintptr_t)mask | (uintptr_t)grad_logit) % 16) == 0) {
        const long long chunks = rows / 8, vb = (chunks + 127) / 128;
        if (vb > 0x7fffffffLL) return DCNV3_E_TOO_LARGE;
        if (dtype == DCNV3_F16)
            dcnv3::dwcb::mask_softmax_bwd_p9x8<__half><<<(unsigned)vb, 128, 0, (cudaStream_t)stream>>>(
                static_cast<const uint4 *>(grad_mask), static_cast<const uint4 *>(mask), static_cast<uint4 *>(grad_logit), chunks);
        else
            dcnv3::dwcb::mask_softmax_bwd_p9x8<__nv_bfloat16><<<(unsigned)vb, 128, 0, (cudaStream_t)stream>>>(
                static_cast<const uint4 *>(grad_mask), static_cast<const uint4 *>(mask), static_cast<uint4 *>(grad_logit), chunks);
        return (int)cudaGetLastError();
    }
    const long long blocks = (rows + 255) / 256;
    if (blocks > 0x7fffffffLL) return DCNV3_E_TOO_LARGE;
    if (dtype == DCNV3_F16)
        dcnv3::dwcb::mask_softmax_bwd<__half><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            static_cast<const __half *>(grad_mask), static_cast<const __half *>(mask), static_cast<__half *>(grad_logit), rows, points);
    else
        dcnv3::dwcb::mask_softmax_bwd<__nv_bfloat16><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            static_cast<const __nv_bfloat16 *>(grad_mask), static_cast<const __nv_bfloat16 *>(mask),
            static_cast<__nv_bfloat16 *>(grad_logit), rows, points);
    return (int)cudaGetLastError();
}
