// dcnv3_forward_mma.cu -- DCNv3 core forward for 16-bit I/O on the tensor cores.
//
// The SIMT tiled forward (dcnv3_forward_tile.cu) is instruction-issue bound: 292 instructions per
// sampling point, of which 128 are the FHFMAs that fold 4 corners x 16 channels into the accumulator
// (profiles/README.md).  For the 32 output pixels o of one warp and the value cells q of the band
// of the window that those pixels can reach, the whole gather is one small dense product
//
//        out[o, c] = sum_q  A[o, q] * value[q, c],
//        A[o, q]   = sum over the sampling points p of pixel o and their corners k landing on q
//                    of (bilinear weight_k * mask_p)                 (dcnv3_im2col_cuda.cuh:56-79,264-267)
//
// so a thread only drops its pixel's 36 scalar coefficients into its own row of A (no per-channel
// work, no corner-rotation logic), and 17 k-steps x 2 x 2 HMMA m16n8k16 per warp produce the
// 32 x 16 outputs in fp32 accumulators.  The value operand is read by ldmatrix straight from the
// TMA-staged window ([cell][16 ch], zero-filled outside the map = the op's zero padding).
// A is stored in the I/O dtype: every coefficient sum is rounded to bf16/fp16 (2^-9 / 2^-12
// relative) -- the DCNV3_WEIGHTS=fast class of error; products and sums are exact / fp32.
// DCNV3_FWD=tile selects the SIMT kernel with fp32-exact weights instead.
//
// A CTA (4 warps) owns an 8x16 (w x h) tile of output pixels of one (image, group).  Points whose
// corner block leaves the window or the warp's band fall back to clamped global reads with fp32
// FMAs; their partial sums join the MMA result through a small shared-memory patch.
#include "../dcnv3_common.cuh"
#include "../dcnv3_launch.h"
#include "../dcnv3_stage.cuh"
#include "../dcnv3_tma.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace dcnv3 {
namespace fmma {

constexpr int kTileW = 8, kTileH = 16, kThreads = kTileW * kTileH, kWarps = kThreads / 32;
constexpr int kRowsPerWarp = kTileH / kWarps;                  // 4 tile rows (32 pixels) per warp
constexpr int kWinW = 18, kWinH = 26, kCells = kWinW * kWinH;   // value window (cells)
constexpr int kBandH = kRowsPerWarp + (kWinH - kTileH) + 1;     // 15 window rows reachable by a warp
constexpr int kBandCells = kBandH * kWinW;                      // 270
constexpr int kKSteps = (kBandCells + 15) / 16;                 // 17
constexpr int kBandPad = kKSteps * 16;                          // 272
constexpr int kCh = 16, kSliceBytes = 32;
constexpr int kWinCellsPad = (kWarps - 1) * kRowsPerWarp * kWinW + kBandPad;   // 488: last band stays inside
constexpr int kAPitch = kBandPad * 2 + 16;                      // 560 B: 8 rows hit 8 distinct 16-byte slots
constexpr size_t kWinBytes = (size_t)kWinCellsPad * kSliceBytes;
constexpr size_t kABytes = (size_t)kWarps * 32 * kAPitch;
static_assert(kAPitch % 16 == 0 && (kAPitch / 16) % 8 == 3, "A row pitch must spread 8 rows over the 8 bank groups");

struct Params {
    int ox_rel, oy_rel, tiles_x, n0;
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, __nv_bfloat16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, __half) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

static size_t smem_bytes(int P) {
    return kWinBytes + kABytes + (size_t)kThreads * kCh * 4 + (size_t)kThreads * P * 6;
}

template <typename T, int KH, int KW>
__global__ void __launch_bounds__(kThreads)
fwd_mma(const __grid_constant__ CUtensorMap tmap, const T *__restrict__ value,
        const T *__restrict__ offset, const T *__restrict__ mask, T *__restrict__ out,
        const Geom q, const Params tp) {
    constexpr int E = 8;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    const int kh = KH ? KH : q.kh, kw = KW ? KW : q.kw;
    const int P = kh * kw;
    unsigned char *win = smem;                                            // [488][32 B] (468 loaded)
    unsigned char *s_a = smem + kWinBytes;                                // [warp][32 pixels][560 B]
    float *s_patch = reinterpret_cast<float *>(s_a + kABytes);            // [128][16] fallback sums
    uint32_t *s_off = reinterpret_cast<uint32_t *>(s_patch + kThreads * kCh);
    T *s_msk = reinterpret_cast<T *>(s_off + kThreads * P);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_x = blockIdx.x % tp.tiles_x, tile_y = blockIdx.x / tp.tiles_x;
    const int g = blockIdx.y, n = tp.n0 + blockIdx.z;
    const int wo0 = tile_x * kTileW, ho0 = tile_y * kTileH;
    const int wo = wo0 + (tid % kTileW), ho = ho0 + (tid / kTileW);
    const bool live = wo < q.Wo && ho < q.Ho;
    const int ox = wo0 * q.sw + tp.ox_rel, oy = ho0 * q.sh + tp.oy_rel;
    const int C = q.G * q.gc, row_stride = q.W * C, ch0 = g * q.gc;
    const size_t img_pix = (size_t)n * q.Ho * q.Wo;

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, kCells * kSliceBytes);
        tma_load_4d(win, &tmap, &bar, ch0, ox, oy, n);
    }
    // ---- while the box is in flight: zero A and the window tail, stage offsets / masks
    {
        uint4 *z = reinterpret_cast<uint4 *>(s_a);
        for (int i = tid; i < (int)(kABytes / 16); i += kThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
        uint4 *zt = reinterpret_cast<uint4 *>(win + (size_t)kCells * kSliceBytes);
        for (int i = tid; i < (kWinCellsPad - kCells) * 2; i += kThreads) zt[i] = make_uint4(0u, 0u, 0u, 0u);
        stage_offsets_masks<T, KH * KW, kThreads, kTileW>(offset, mask, s_off, s_msk, P, tid, wo0, ho0, q.Wo, q.Ho, q.G, g, img_pix);
    }
    const float base_w = axis_base(wo, kw, q.sw, q.pw, q.dw, q.sigma);
    const float base_h = axis_base(ho, kh, q.sh, q.ph, q.dh, q.sigma);
    const T *img = value + (size_t)n * q.H * row_stride + ch0;
    unsigned char *arow = s_a + ((size_t)warp * 32 + lane) * kAPitch;     // this pixel's row of A
    const int band_cell0 = warp * kRowsPerWarp * kWinW;
    __syncthreads();

    // ---- A build: thread <-> pixel; the value window is not needed yet
    float fb[kCh];           // fallback partial sums (points served from global memory)
    bool used_fb = false;
#pragma unroll
    for (int c = 0; c < kCh; ++c) fb[c] = 0.f;
    if (live) {
#pragma unroll
        for (int i = 0; i < kw; ++i) {
#pragma unroll
            for (int jj = 0; jj < kh; ++jj) {
                const int p = i * kh + jj;
                const float2 d = unpack2(s_off[tid * P + p], T());
                const float m = to_f32(s_msk[tid * P + p]);
                const float loc_w = base_w + ((float)(i * q.dw) + d.x) * q.sigma;
                const float loc_h = base_h + ((float)(jj * q.dh) + d.y) * q.sigma;
                // range test of the reference (dcnv3_im2col_cuda.cuh:262-263); also rejects NaN
                const bool inside = loc_h > -1.f && loc_w > -1.f && loc_h < (float)q.H && loc_w < (float)q.W;
                if (!inside) continue;
                const float fh = floorf(loc_h), fw = floorf(loc_w);
                const float lh = loc_h - fh, lw = loc_w - fw, hh = 1.f - lh, hw = 1.f - lw;
                const int hwin = (int)fh - oy, wwin = (int)fw - ox;
                const int cb = hwin * kWinW + wwin - band_cell0;
                if ((unsigned)hwin < (unsigned)(kWinH - 1) && (unsigned)wwin < (unsigned)(kWinW - 1) &&
                    cb >= 0 && cb < (kBandH - 1) * kWinW - 1) {
                    // the four corner cells are distinct: read all four, then write all four
                    T *e = reinterpret_cast<T *>(arow) + cb;
                    const float a0 = to_f32(e[0]), a1 = to_f32(e[1]), a2 = to_f32(e[kWinW]), a3 = to_f32(e[kWinW + 1]);
                    e[0] = from_f32<T>(a0 + hh * hw * m);
                    e[1] = from_f32<T>(a1 + hh * lw * m);
                    e[kWinW] = from_f32<T>(a2 + lh * hw * m);
                    e[kWinW + 1] = from_f32<T>(a3 + lh * lw * m);
                } else {
                    // ---- fallback: clamped global reads, fp32 weights
                    used_fb = true;
                    const ClampedTap ct = make_clamped_tap(loc_h, loc_w, q.H, q.W);
                    const T *r_lo = img + ct.row_lo * row_stride, *r_hi = img + ct.row_hi * row_stride;
                    const int c_lo = ct.col_lo * C, c_hi = ct.col_hi * C;
                    const float fy_lo = ct.hh * ct.top * m, fy_hi = ct.lh * ct.bot * m;
                    const float fx_lo = ct.hw * ct.lef, fx_hi = ct.lw * ct.rig;
                    const T *corner[4] = {r_lo + c_lo, r_lo + c_hi, r_hi + c_lo, r_hi + c_hi};
                    const float wc[4] = {fy_lo * fx_lo, fy_lo * fx_hi, fy_hi * fx_lo, fy_hi * fx_hi};
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const Weight<T, false> wt(wc[t]);
                        axpy<T, false>(fb, __ldg(reinterpret_cast<const uint4 *>(corner[t])), wt);
                        axpy<T, false>(fb + E, __ldg(reinterpret_cast<const uint4 *>(corner[t] + E)), wt);
                    }
                }
            }
        }
    }
    const bool warp_fb = __any_sync(0xffffffffu, used_fb);
    if (warp_fb) {
#pragma unroll
        for (int c = 0; c < kCh; c += 4)
            *reinterpret_cast<float4 *>(s_patch + tid * kCh + c) = make_float4(fb[c], fb[c + 1], fb[c + 2], fb[c + 3]);
    }
    __syncwarp();
    mbar_wait(&bar, 0);        // value window has landed

    // ---- out (32 x 16) = A (32 x 272) * value band (272 x 16)
    float acc[2][2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
    {
        const int r_in = (lane & 7) + ((lane >> 3) & 1) * 8, kc_in = lane >> 4;
        const uint32_t a_addr = smem_u32(s_a) + (uint32_t)(warp * 32 + r_in) * kAPitch + kc_in * 16;
        const uint32_t b_addr = smem_u32(win) + (uint32_t)(band_cell0 + r_in) * kSliceBytes + kc_in * 16;
#pragma unroll
        for (int ks = 0; ks < kKSteps; ++ks) {
            uint32_t a0[4], a1[4], b[4];
            ldmatrix_x4(a0, a_addr + ks * 32);                               // pixels 0-15, cells 16ks..
            ldmatrix_x4(a1, a_addr + 16 * kAPitch + ks * 32);                // pixels 16-31
            ldmatrix_x4_trans(b, b_addr + ks * 16 * kSliceBytes);            // {n0:k0-7, n0:k8-15, n1:k0-7, n1:k8-15}
            mma16816(acc[0][0], a0, b[0], b[1], T());
            mma16816(acc[0][1], a0, b[2], b[3], T());
            mma16816(acc[1][0], a1, b[0], b[1], T());
            mma16816(acc[1][1], a1, b[2], b[3], T());
        }
    }
    // ---- epilogue: C fragment (row = lane/4 (+8), cols 2*(lane%4)+{0,1}) -> out, 16-bit pairs
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            const int px = warp * 32 + mt * 16 + hrow * 8 + (lane >> 2);   // pixel (thread id) of this row
            const int w = wo0 + (px % kTileW), h = ho0 + (px / kTileW);
            if (w < q.Wo && h < q.Ho) {
                T *dst = out + (img_pix + (size_t)h * q.Wo + w) * C + ch0 + 2 * (lane & 3);
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    float v0 = acc[mt][nt][hrow * 2], v1 = acc[mt][nt][hrow * 2 + 1];
                    if (warp_fb) {
                        const float2 f = *reinterpret_cast<const float2 *>(s_patch + px * kCh + nt * 8 + 2 * (lane & 3));
                        v0 += f.x; v1 += f.y;
                    }
                    *reinterpret_cast<uint32_t *>(dst + nt * 8) = pack2(v0, v1, T());
                }
            }
        }
    }
}

template <typename T>
static bool launch_typed(const void *value, const void *offset, const void *mask, void *out,
                         const Geom &q, int dtype, cudaStream_t stream, cudaError_t *err) {
    if (q.gc != kCh || q.kh > 8 || q.kw > 8 || q.G > 65535) return false;
    if (((uintptr_t)value | (uintptr_t)out) % 16 || (uintptr_t)offset % 4) return false;
    const float span_w = (kTileW - 1) * q.sw + (q.kw - 1) * q.dw * q.sigma;
    const float span_h = (kTileH - 1) * q.sh + (q.kh - 1) * q.dh * q.sigma;
    if (!(q.sigma > 0.f) || span_w + 4 > kWinW - 2 || span_h + 4 > kWinH - 2) return false;
    const int C = q.G * q.gc, P = q.kh * q.kw;
    const size_t smem = smem_bytes(P);
    if (smem > 112 * 1024) return false;
    CUtensorMap tmap;
    if (!make_nhwc_tensor_map(&tmap, value, dtype, q.N, q.H, q.W, C, kCh, kWinW, kWinH)) return false;
    Params tp;
    const int cw = (q.dw * (q.kw - 1)) >> 1, chh = (q.dh * (q.kh - 1)) >> 1;
    const float a_w = (float)(cw - q.pw) - cw * q.sigma, a_h = (float)(chh - q.ph) - chh * q.sigma;
    tp.ox_rel = (int)std::floor(a_w + 0.5f * span_w - 0.5f * (kWinW - 2));
    tp.oy_rel = (int)std::floor(a_h + 0.5f * span_h - 0.5f * (kWinH - 2));
    tp.tiles_x = (q.Wo + kTileW - 1) / kTileW;
    const unsigned tiles = tp.tiles_x * ((q.Ho + kTileH - 1) / kTileH);
    const T *v = static_cast<const T *>(value), *o = static_cast<const T *>(offset),
            *m = static_cast<const T *>(mask);
    T *y = static_cast<T *>(out);
    const bool k33 = q.kh == 3 && q.kw == 3;
    if (k33) cudaFuncSetAttribute(fwd_mma<T, 3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else cudaFuncSetAttribute(fwd_mma<T, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int n0 = 0; n0 < q.N; n0 += 65535) {
        tp.n0 = n0;
        const dim3 grid(tiles, (unsigned)q.G, (unsigned)std::min(65535, q.N - n0));
        if (k33) fwd_mma<T, 3, 3><<<grid, kThreads, smem, stream>>>(tmap, v, o, m, y, q, tp);
        else fwd_mma<T, 0, 0><<<grid, kThreads, smem, stream>>>(tmap, v, o, m, y, q, tp);
    }
    *err = cudaGetLastError();
    return true;
}

}  // namespace fmma

// 16-bit I/O with group_channels == 16; false = not eligible (caller uses the SIMT kernels).
bool try_launch_forward_mma(const void *value, const void *offset, const void *mask, void *out,
                            const Geom &q, int dtype, cudaStream_t stream, cudaError_t *err) {
    // Opt-in (DCNV3_FWD=mma): parity-tested, but at 2 CTAs/SM (72 KB of A tiles per CTA) it is
    // latency-bound and slower than the SIMT tiled kernel on B200 (262 vs 168 us, profiles/README.md)
    const char *e = std::getenv("DCNV3_FWD");
    if (!(e && e[0] == 'm')) return false;
    if ((long long)q.N * q.Ho * q.Wo == 0) return false;
    if (dtype == 1) return fmma::launch_typed<__half>(value, offset, mask, out, q, dtype, stream, err);
    if (dtype == 2) return fmma::launch_typed<__nv_bfloat16>(value, offset, mask, out, q, dtype, stream, err);
    return false;
}

}  // namespace dcnv3
