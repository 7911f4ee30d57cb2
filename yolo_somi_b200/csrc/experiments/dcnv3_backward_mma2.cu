// dcnv3_backward_mma2.cu -- the tensor-core backward of dcnv3_backward_mma.cu split in two launches
// so that each part runs at the occupancy it needs:
//   bwd_dots        grad_offset / grad_mask.  Same structure as fwd_tile (16x16 tile, 26x26 TMA value
//                   window, conflict-free rotated LDS.128 gather, exact FHFMA corner dot products
//                   with the pixel's upstream gradient); ~35 KB of shared memory per CTA, so the
//                   gather is no longer latency-bound at 2 CTAs per SM.
//   bwd_value_mma   grad_value.  Needs only offsets, masks and grad_out: every thread recomputes its
//                   pixel's sampling coordinates (cheap), drops the 36 coefficients w_k*m into its
//                   column of the warp's A tile, then A (272 cells x 32 pixels) x grad_out
//                   (32 x 16 ch) on the tensor cores (HMMA m16n8k16, fp32 accumulate), bands summed
//                   and flushed with 128-bit reductions.  81 KB per CTA (the four 17 KB A tiles).
// See dcnv3_backward_mma.cu for the derivation and the precision note (A is stored in the I/O dtype).
#include "../dcnv3_common.cuh"
#include "../dcnv3_launch.h"
#include "../dcnv3_stage.cuh"
#include "../dcnv3_tma.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace dcnv3 {
namespace mma2 {

// ------------------------------------------------------------------------------------------------
// shared helpers
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
template <typename V> __device__ __forceinline__ void rotate4(V (&x)[4], int r) {  // out[t] = in[(t+r)&3]
    if (r & 1) { const V t = x[0]; x[0] = x[1]; x[1] = x[2]; x[2] = x[3]; x[3] = t; }
    if (r & 2) { V t = x[0]; x[0] = x[2]; x[2] = t; t = x[1]; x[1] = x[3]; x[3] = t; }
}
__device__ __forceinline__ void red_add4(float *p, float4 v) { atomicAdd(reinterpret_cast<float4 *>(p), v); }
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

struct Params {
    int ox_rel, oy_rel, tiles_x, n0;
};

// ================================================================================================
// Kernel 1: grad_offset / grad_mask
namespace dots {
constexpr int kTile = 16, kWin = 26, kThreads = kTile * kTile, kSliceBytes = 32;
constexpr int kWinBytes = kWin * kWin * kSliceBytes;
static_assert(kWin % 4 == 2, "window width must be 2 mod 4");
static size_t smem_bytes(int P) { return kWinBytes + (size_t)kThreads * P * 6; }
}  // namespace dots

template <typename T, int KH, int KW>
__global__ void __launch_bounds__(dots::kThreads)
bwd_dots(const __grid_constant__ CUtensorMap tmap, const T *__restrict__ value,
         const T *__restrict__ offset, const T *__restrict__ mask, const T *__restrict__ grad_out,
         T *__restrict__ grad_offset, T *__restrict__ grad_mask, const Geom q, const Params tp) {
    using namespace dots;
    constexpr int E = 8;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    const int kh = KH ? KH : q.kh, kw = KW ? KW : q.kw;
    const int P = kh * kw;
    unsigned char *win = smem;
    uint32_t *s_off = reinterpret_cast<uint32_t *>(smem + kWinBytes);   // [256][P] (dx,dy) pairs
    T *s_msk = reinterpret_cast<T *>(s_off + kThreads * P);             // [256][P]

    const int tid = threadIdx.x;
    const int tile_x = blockIdx.x % tp.tiles_x, tile_y = blockIdx.x / tp.tiles_x;
    const int g = blockIdx.y, n = tp.n0 + blockIdx.z;
    const int wo0 = tile_x * kTile, ho0 = tile_y * kTile;
    const int wo = wo0 + (tid % kTile), ho = ho0 + (tid / kTile);
    const bool live = wo < q.Wo && ho < q.Ho;
    const int ox = wo0 * q.sw + tp.ox_rel, oy = ho0 * q.sh + tp.oy_rel;
    const int C = q.G * q.gc, row_stride = q.W * C, ch0 = g * q.gc;

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, kWinBytes);
        tma_load_4d(win, &tmap, &bar, ch0, ox, oy, n);
    }
    const size_t img_pix = (size_t)n * q.Ho * q.Wo;
    stage_offsets_masks<T, KH * KW, kThreads, kTile>(offset, mask, s_off, s_msk, P, tid, wo0, ho0, q.Wo, q.Ho, q.G, g, img_pix);
    const int j = tid & 7, half = j & 1;
    // upstream gradient of this thread's pixel: chunk `half` first
    uint4 gq_a = make_uint4(0u, 0u, 0u, 0u), gq_b = gq_a;
    if (live) {
        const T *gp = grad_out + (img_pix + (size_t)ho * q.Wo + wo) * C + ch0;
        gq_a = __ldg(reinterpret_cast<const uint4 *>(gp + half * E));
        gq_b = __ldg(reinterpret_cast<const uint4 *>(gp + (half ^ 1) * E));
    }
    const float base_w = axis_base(wo, kw, q.sw, q.pw, q.dw, q.sigma);
    const float base_h = axis_base(ho, kh, q.sh, q.ph, q.dh, q.sigma);
    const uint32_t win_addr = smem_u32(win) + half * 16;
    const T *img = value + (size_t)n * q.H * row_stride + ch0;

    __syncthreads();
    mbar_wait(&bar, 0);

    if (live) {
#pragma unroll
        for (int i = 0; i < kw; ++i) {
#pragma unroll
            for (int jj = 0; jj < kh; ++jj) {
                const int p = i * kh + jj;
                float gm = 0.f, gx = 0.f, gy = 0.f;
                const float2 d = unpack2(s_off[tid * P + p], T());
                const float m = to_f32(s_msk[tid * P + p]);
                const float loc_w = base_w + ((float)(i * q.dw) + d.x) * q.sigma;
                const float loc_h = base_h + ((float)(jj * q.dh) + d.y) * q.sigma;
                const bool inside = loc_h > -1.f && loc_w > -1.f && loc_h < (float)q.H && loc_w < (float)q.W;
                const float fh = floorf(loc_h), fw = floorf(loc_w);
                const float lh = loc_h - fh, lw = loc_w - fw, hh = 1.f - lh, hw = 1.f - lw;
                const int hwin = (int)fh - oy, wwin = (int)fw - ox;
                if (!inside) {
                } else if ((unsigned)hwin < (unsigned)(kWin - 1) && (unsigned)wwin < (unsigned)(kWin - 1)) {
                    int o[4] = {0, kSliceBytes, kWin * kSliceBytes, kWin * kSliceBytes + kSliceBytes};
                    const int rho = ((j >> 1) - (wwin + 2 * hwin)) & 3;
                    rotate4(o, rho);
                    const uint32_t tl = win_addr + (uint32_t)(hwin * kWin + wwin) * kSliceBytes;
                    float dr[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const uint32_t a = tl + o[t];
                        const uint4 qa = lds128(a), qb = lds128(a ^ 16u);
                        dr[t] = dot<T>(gq_a, qa, 0.f) + dot<T>(gq_b, qb, 0.f);
                    }
                    rotate4(dr, (4 - rho) & 3);
                    gm = hh * hw * dr[0] + hh * lw * dr[1] + lh * hw * dr[2] + lh * lw * dr[3];
                    gx = m * (hh * (dr[1] - dr[0]) + lh * (dr[3] - dr[2]));
                    gy = m * (hw * (dr[2] - dr[0]) + lw * (dr[3] - dr[1]));
                } else {
                    const ClampedTap ct = make_clamped_tap(loc_h, loc_w, q.H, q.W);
                    const int r_lo = ct.row_lo * row_stride, r_hi = ct.row_hi * row_stride;
                    const int c_lo = ct.col_lo * C, c_hi = ct.col_hi * C;
                    const int at[4] = {r_lo + c_lo, r_lo + c_hi, r_hi + c_lo, r_hi + c_hi};
                    const int ea = half * E, eb = (half ^ 1) * E;
                    float dk[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const uint4 qa = __ldg(reinterpret_cast<const uint4 *>(img + at[t] + ea));
                        const uint4 qb = __ldg(reinterpret_cast<const uint4 *>(img + at[t] + eb));
                        dk[t] = dot<T>(gq_a, qa, 0.f) + dot<T>(gq_b, qb, 0.f);
                    }
                    const float fy_lo = ct.hh * ct.top, fy_hi = ct.lh * ct.bot;
                    const float fx_lo = ct.hw * ct.lef, fx_hi = ct.lw * ct.rig;
                    gm = fy_lo * fx_lo * dk[0] + fy_lo * fx_hi * dk[1] + fy_hi * fx_lo * dk[2] + fy_hi * fx_hi * dk[3];
                    gx = m * (fy_lo * (ct.rig * dk[1] - ct.lef * dk[0]) + fy_hi * (ct.rig * dk[3] - ct.lef * dk[2]));
                    gy = m * (fx_lo * (ct.bot * dk[2] - ct.top * dk[0]) + fx_hi * (ct.bot * dk[3] - ct.top * dk[1]));
                }
                s_off[tid * P + p] = pack2(q.sigma * gx, q.sigma * gy, T());
                s_msk[tid * P + p] = from_f32<T>(gm);
            }
        }
    }
    __syncthreads();
    for (int idx = tid; idx < kThreads * P; idx += kThreads) {
        const int px = idx / P, p = idx - px * P;
        const int w = wo0 + (px % kTile), h = ho0 + (px / kTile);
        if (w < q.Wo && h < q.Ho) {
            const size_t pgi = (img_pix + (size_t)h * q.Wo + w) * q.G + g;
            reinterpret_cast<uint32_t *>(grad_offset)[pgi * P + p] = s_off[idx];
            grad_mask[pgi * P + p] = s_msk[idx];
        }
    }
}

// ================================================================================================
// Kernel 2: grad_value = A x grad_out
namespace val {
constexpr int kTileW = 8, kTileH = 16, kThreads = kTileW * kTileH, kWarps = kThreads / 32;
constexpr int kRowsPerWarp = kTileH / kWarps;                 // 4 tile rows (32 pixels) per warp
constexpr int kWinW = 18, kWinH = 26, kCells = kWinW * kWinH;  // accumulator window (cells)
constexpr int kBandH = kRowsPerWarp + (kWinH - kTileH) + 1;    // 15 rows reachable by one warp
constexpr int kBandCells = kBandH * kWinW;                     // 270
constexpr int kMTiles = (kBandCells + 15) / 16;                // 17
constexpr int kBufCells = kMTiles * 16;                        // 272 rows of 64 bytes
constexpr int kCh = 16, kSliceBytes = 32;
constexpr size_t kBufBytes = (size_t)kWarps * kBufCells * 64;  // A tiles / result bands
static size_t smem_bytes(int P) { return kBufBytes + (size_t)kThreads * kSliceBytes + (size_t)kThreads * P * 6; }
// byte offset of A[cell][pixel]: 64-byte rows, 16-byte chunks XOR-swizzled (conflict-free ldmatrix)
__device__ __forceinline__ uint32_t a_elem_off(int cell, int pixel) {
    return (uint32_t)cell * 64u + ((((uint32_t)pixel >> 3) ^ (((uint32_t)cell >> 1) & 3u)) << 4) + (((uint32_t)pixel & 7u) << 1);
}
}  // namespace val

template <typename T, int KH, int KW>
__global__ void __launch_bounds__(val::kThreads)
bwd_value_mma(const T *__restrict__ offset, const T *__restrict__ mask, const T *__restrict__ grad_out,
              float *__restrict__ gv_acc, const Geom q, const Params tp) {
    using namespace val;
    constexpr int E = 8;
    extern __shared__ __align__(128) unsigned char smem[];
    const int kh = KH ? KH : q.kh, kw = KW ? KW : q.kw;
    const int P = kh * kw;
    unsigned char *s_buf = smem;                                            // [warp][272][64 B]
    unsigned char *s_gout = smem + kBufBytes;                               // [128][32 B]
    uint32_t *s_off = reinterpret_cast<uint32_t *>(s_gout + kThreads * kSliceBytes);
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

    {
        uint4 *z = reinterpret_cast<uint4 *>(s_buf);
        for (int i = tid; i < kWarps * kBufCells * 4; i += kThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
        stage_offsets_masks<T, KH * KW, kThreads, kTileW>(offset, mask, s_off, s_msk, P, tid, wo0, ho0, q.Wo, q.Ho, q.G, g, img_pix);
        for (int idx = tid; idx < kThreads * 2; idx += kThreads) {
            const int px = idx >> 1, c = idx & 1;
            const int w = wo0 + (px % kTileW), h = ho0 + (px / kTileW);
            uint4 v = make_uint4(0u, 0u, 0u, 0u);   // pixels outside the map contribute nothing
            if (w < q.Wo && h < q.Ho)
                v = __ldg(reinterpret_cast<const uint4 *>(grad_out + (img_pix + (size_t)h * q.Wo + w) * C + ch0 + c * E));
            *reinterpret_cast<uint4 *>(s_gout + px * kSliceBytes + c * 16) = v;
        }
    }
    const float base_w = axis_base(wo, kw, q.sw, q.pw, q.dw, q.sigma);
    const float base_h = axis_base(ho, kh, q.sh, q.ph, q.dh, q.sigma);
    const size_t img_base = (size_t)n * q.H * row_stride + ch0;
    unsigned char *abuf = s_buf + (size_t)warp * kBufCells * 64;
    const int band_cell0 = warp * kRowsPerWarp * kWinW;
    __syncthreads();

    // ---- A build: thread <-> pixel (column `lane` of the warp's A tile)
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
                const bool inside = loc_h > -1.f && loc_w > -1.f && loc_h < (float)q.H && loc_w < (float)q.W;
                if (!inside) continue;
                const float fh = floorf(loc_h), fw = floorf(loc_w);
                const float lh = loc_h - fh, lw = loc_w - fw, hh = 1.f - lh, hw = 1.f - lw;
                const int h0 = (int)fh, w0 = (int)fw;
                const int hwin = h0 - oy, wwin = w0 - ox;
                const int cb = hwin * kWinW + wwin - band_cell0;
                const float c1 = hh * hw * m, c2 = hh * lw * m, c3 = lh * hw * m, c4 = lh * lw * m;
                if ((unsigned)hwin < (unsigned)(kWinH - 1) && (unsigned)wwin < (unsigned)(kWinW - 1) &&
                    cb >= 0 && cb < (kBandH - 1) * kWinW - 1) {
                    // the four corner cells are distinct: read all four, then write all four
                    T *e0 = reinterpret_cast<T *>(abuf + a_elem_off(cb, lane));
                    T *e1 = reinterpret_cast<T *>(abuf + a_elem_off(cb + 1, lane));
                    T *e2 = reinterpret_cast<T *>(abuf + a_elem_off(cb + kWinW, lane));
                    T *e3 = reinterpret_cast<T *>(abuf + a_elem_off(cb + kWinW + 1, lane));
                    const float a0 = to_f32(*e0), a1 = to_f32(*e1), a2 = to_f32(*e2), a3 = to_f32(*e3);
                    *e0 = from_f32<T>(a0 + c1);
                    *e1 = from_f32<T>(a1 + c2);
                    *e2 = from_f32<T>(a2 + c3);
                    *e3 = from_f32<T>(a3 + c4);
                } else {
                    // corner block leaves the window / band: direct reductions (corners outside the
                    // map are skipped -- zero padding)
                    float gf[kCh];
                    unpack<T>(*reinterpret_cast<const uint4 *>(s_gout + tid * kSliceBytes), gf);
                    unpack<T>(*reinterpret_cast<const uint4 *>(s_gout + tid * kSliceBytes + 16), gf + E);
                    const float cs[4] = {c1, c2, c3, c4};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int y = h0 + (k >> 1), x = w0 + (k & 1);
                        if ((unsigned)y < (unsigned)q.H && (unsigned)x < (unsigned)q.W && cs[k] != 0.f) {
                            float *dst = gv_acc + img_base + (size_t)y * row_stride + (size_t)x * C;
#pragma unroll
                            for (int e = 0; e < kCh; e += 4)
                                red_add4(dst + e, make_float4(cs[k] * gf[e], cs[k] * gf[e + 1], cs[k] * gf[e + 2], cs[k] * gf[e + 3]));
                        }
                    }
                }
            }
        }
    }
    __syncwarp();

    // ---- band = A (272 x 32) * grad_out (32 x 16), result overwrites A m-tile by m-tile
    {
        const uint32_t a_base = smem_u32(abuf);
        uint32_t bf[2][4];   // [k-step][{n0:k0-7, n0:k8-15, n1:k0-7, n1:k8-15}]
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const int px = warp * 32 + ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
            ldmatrix_x4_trans(bf[ks], smem_u32(s_gout) + px * kSliceBytes + (lane >> 4) * 16);
        }
        const int r_in = (lane & 7) + ((lane >> 3) & 1) * 8, kc_in = lane >> 4;
#pragma unroll 4
        for (int mt = 0; mt < kMTiles; ++mt) {
            float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
            const int row = mt * 16 + r_in;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t af[4];
                const uint32_t chunk = (uint32_t)(ks * 2 + kc_in) ^ (((uint32_t)row >> 1) & 3u);
                ldmatrix_x4(af, a_base + (uint32_t)row * 64u + (chunk << 4));
                mma16816(acc0, af, bf[ks][0], bf[ks][1], T());
                mma16816(acc1, af, bf[ks][2], bf[ks][3], T());
            }
            float *r0 = reinterpret_cast<float *>(abuf) + (size_t)(mt * 16 + (lane >> 2)) * kCh + 2 * (lane & 3);
            *reinterpret_cast<float2 *>(r0) = make_float2(acc0[0], acc0[1]);
            *reinterpret_cast<float2 *>(r0 + 8) = make_float2(acc1[0], acc1[1]);
            *reinterpret_cast<float2 *>(r0 + 8 * kCh) = make_float2(acc0[2], acc0[3]);
            *reinterpret_cast<float2 *>(r0 + 8 * kCh + 8) = make_float2(acc1[2], acc1[3]);
        }
    }
    __syncthreads();

    // ---- flush: sum the warps' bands per window cell, add to the global accumulator
    const float *bands = reinterpret_cast<const float *>(s_buf);
    for (int idx = tid; idx < kCells * 4; idx += kThreads) {
        const int cell = idx >> 2, piece = idx & 3;
        const int wy = cell / kWinW, wx = cell % kWinW;
        const int y = oy + wy, x = ox + wx;
        if ((unsigned)y < (unsigned)q.H && (unsigned)x < (unsigned)q.W) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const int by = wy - w * kRowsPerWarp;
                if (by >= 0 && by < kBandH) {
                    const float4 t = *reinterpret_cast<const float4 *>(
                        bands + ((size_t)w * kBufCells + by * kWinW + wx) * kCh + piece * 4);
                    v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
                }
            }
            if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
                red_add4(gv_acc + img_base + (size_t)y * row_stride + (size_t)x * C + piece * 4, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------
static bool window_origin(const Geom &q, int tile_w, int tile_h, int win_w, int win_h, Params *tp) {
    const float span_w = (tile_w - 1) * q.sw + (q.kw - 1) * q.dw * q.sigma;
    const float span_h = (tile_h - 1) * q.sh + (q.kh - 1) * q.dh * q.sigma;
    if (!(q.sigma > 0.f) || span_w + 4 > win_w - 2 || span_h + 4 > win_h - 2) return false;
    const int cw = (q.dw * (q.kw - 1)) >> 1, chh = (q.dh * (q.kh - 1)) >> 1;
    const float a_w = (float)(cw - q.pw) - cw * q.sigma, a_h = (float)(chh - q.ph) - chh * q.sigma;
    tp->ox_rel = (int)std::floor(a_w + 0.5f * span_w - 0.5f * (win_w - 2));
    tp->oy_rel = (int)std::floor(a_h + 0.5f * span_h - 0.5f * (win_h - 2));
    return true;
}

template <typename T>
static bool launch_typed(const void *value, const void *offset, const void *mask, const void *grad_out,
                         float *gv_acc, void *grad_offset, void *grad_mask, const Geom &q, int dtype,
                         cudaStream_t stream, cudaError_t *err) {
    if (q.gc != val::kCh || q.kh > 8 || q.kw > 8 || q.G > 65535) return false;
    if (((uintptr_t)value | (uintptr_t)grad_out | (uintptr_t)gv_acc) % 16) return false;
    if (((uintptr_t)offset | (uintptr_t)grad_offset) % 4) return false;
    Params pd, pv;
    if (!window_origin(q, dots::kTile, dots::kTile, dots::kWin, dots::kWin, &pd)) return false;
    if (!window_origin(q, val::kTileW, val::kTileH, val::kWinW, val::kWinH, &pv)) return false;
    const int C = q.G * q.gc, P = q.kh * q.kw;
    const size_t smem_d = dots::smem_bytes(P), smem_v = val::smem_bytes(P);
    if (smem_d > 100 * 1024 || smem_v > 110 * 1024) return false;
    CUtensorMap tmap;
    if (!make_nhwc_tensor_map(&tmap, value, dtype, q.N, q.H, q.W, C, val::kCh, dots::kWin, dots::kWin)) return false;
    pd.tiles_x = (q.Wo + dots::kTile - 1) / dots::kTile;
    pv.tiles_x = (q.Wo + val::kTileW - 1) / val::kTileW;
    const unsigned tiles_d = pd.tiles_x * ((q.Ho + dots::kTile - 1) / dots::kTile);
    const unsigned tiles_v = pv.tiles_x * ((q.Ho + val::kTileH - 1) / val::kTileH);
    const T *v = static_cast<const T *>(value), *o = static_cast<const T *>(offset),
            *m = static_cast<const T *>(mask), *go = static_cast<const T *>(grad_out);
    T *goff = static_cast<T *>(grad_offset), *gmsk = static_cast<T *>(grad_mask);
    const bool k33 = q.kh == 3 && q.kw == 3;
    if (k33) {
        cudaFuncSetAttribute(bwd_dots<T, 3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d);
        cudaFuncSetAttribute(bwd_value_mma<T, 3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v);
    } else {
        cudaFuncSetAttribute(bwd_dots<T, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d);
        cudaFuncSetAttribute(bwd_value_mma<T, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v);
    }
    for (int n0 = 0; n0 < q.N; n0 += 65535) {
        pd.n0 = pv.n0 = n0;
        const unsigned nz = (unsigned)std::min(65535, q.N - n0);
        const dim3 gd(tiles_d, (unsigned)q.G, nz), gv(tiles_v, (unsigned)q.G, nz);
        if (k33) {
            bwd_value_mma<T, 3, 3><<<gv, val::kThreads, smem_v, stream>>>(o, m, go, gv_acc, q, pv);
            bwd_dots<T, 3, 3><<<gd, dots::kThreads, smem_d, stream>>>(tmap, v, o, m, go, goff, gmsk, q, pd);
        } else {
            bwd_value_mma<T, 0, 0><<<gv, val::kThreads, smem_v, stream>>>(o, m, go, gv_acc, q, pv);
            bwd_dots<T, 0, 0><<<gd, dots::kThreads, smem_d, stream>>>(tmap, v, o, m, go, goff, gmsk, q, pd);
        }
    }
    *err = cudaGetLastError();
    return true;
}

}  // namespace mma2

bool try_launch_backward_mma2(const void *value, const void *offset, const void *mask,
                              const void *grad_out, float *gv_acc, void *grad_offset, void *grad_mask,
                              const Geom &q, int dtype, cudaStream_t stream, cudaError_t *err) {
    const char *e = std::getenv("DCNV3_BWD");   // development knob: DCNV3_BWD=mma2 selects the split form
    if (!(e && e[0] == 'm' && e[1] == 'm' && e[2] == 'a' && e[3] == '2')) return false;
    if ((long long)q.N * q.Ho * q.Wo == 0) return false;
    if (dtype == 1) return mma2::launch_typed<__half>(value, offset, mask, grad_out, gv_acc, grad_offset, grad_mask, q, dtype, stream, err);
    if (dtype == 2) return mma2::launch_typed<__nv_bfloat16>(value, offset, mask, grad_out, gv_acc, grad_offset, grad_mask, q, dtype, stream, err);
    return false;
}

}  // namespace dcnv3
