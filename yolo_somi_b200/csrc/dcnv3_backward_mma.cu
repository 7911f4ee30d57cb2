// dcnv3_backward_mma.cu -- DCNv3 core backward for 16-bit I/O: shared-memory tiled gather for
// grad_offset / grad_mask, and grad_value as a small dense product on the tensor cores.
//
// Why (profiles/README.md): the channel-corner additions of grad_value are 3.7 GB of fp32
// read-modify-write payload.  Sent to L2 as vector reductions they take 1.04 ms; summed in shared
// memory with SIMT read-modify-writes they are bounded by 7.5 GB of shared-memory traffic and, in
// practice, by instruction issue (profiles r1_bwd_tile_*).  But for the 32 output pixels o of one
// warp and the accumulator cells q of that warp's band the whole scatter is one product
//
//        grad_value[q, c] += sum_o  A[q, o] * grad_out[o, c],
//        A[q, o] = sum over the sampling points p of pixel o and their corners k landing on q of
//                  (bilinear weight_k * mask_p)                         (dcnv3_im2col_cuda.cuh:106-140)
//
// so every thread only has to drop its pixel's 36 scalar coefficients into its own column of A
// (thread-exclusive: no atomics, no races), and 17 x 2 x 2 HMMA m16n8k16 per warp do the
// 272 x 32 x 16 product in fp32 accumulators.  A's entries are stored in the I/O dtype (bf16: 2^-9
// relative rounding of each coefficient, the same class of error as DCNV3_WEIGHTS=fast in the
// forward; the products and the sums are exact / fp32).
//
// A CTA (4 warps) owns an 8x16 (w x h) tile of output pixels of one (image, group); a warp owns 4
// rows (32 pixels, lane <-> pixel) and a private 17 KB buffer that first holds its A tile
// [272 cells][32 pixels] (16-byte chunks XOR-swizzled so that ldmatrix is conflict-free) and is then
// overwritten, m-tile by m-tile, with the fp32 result band [272 cells][16 channels].  Bands are
// summed and added to the global fp32 accumulator with 128-bit reductions (cells outside the map
// and all-zero pieces are skipped), as in dcnv3_backward_tile.cu.  The gather of the value window
// (one TMA box, zero-filled outside the map) and the exact-FHFMA corner dot products for
// grad_offset / grad_mask are those of the tiled kernels.  Points whose corner block leaves the
// window / the band fall back to clamped global reads and direct reductions.
#include "dcnv3_common.cuh"
#include "dcnv3_launch.h"
#include "dcnv3_stage.cuh"
#include "dcnv3_tma.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace dcnv3 {
namespace mma {

constexpr int kTileW = 8, kTileH = 16;           // output pixels per tile
constexpr int kThreads = kTileW * kTileH;        // 128: one thread per pixel, 4 warps
constexpr int kWarps = kThreads / 32;
constexpr int kRowsPerWarp = kTileH / kWarps;    // 4 tile rows (32 pixels) per warp
constexpr int kWinW = 18, kWinH = 26;            // window (value pixels == accumulator cells)
constexpr int kCells = kWinW * kWinH;
constexpr int kBandH = kRowsPerWarp + (kWinH - kTileH) + 1;   // 15 window rows reachable by a warp
constexpr int kBandCells = kBandH * kWinW;       // 270
constexpr int kMTiles = (kBandCells + 15) / 16;  // 17
constexpr int kBufCells = kMTiles * 16;          // 272 rows of 64 bytes
constexpr int kSliceBytes = 32;                  // 16 channels of 16-bit data
constexpr int kCh = 16;
static_assert(kWinW % 4 == 2, "window width must be 2 mod 4 (conflict-free corner layout)");
static_assert(kTileW == 8, "a quarter-warp must be one tile row (rotation scheme)");

struct Params {
    int ox_rel, oy_rel, tiles_x, n0;
};

struct Layout {   // bytes; every region 16-byte aligned
    static constexpr size_t win = 0;                                        // [kWinH][kWinW][32 B]
    static constexpr size_t buf = win + (size_t)kCells * kSliceBytes;       // [warp][272][64 B]
    static constexpr size_t gout = buf + (size_t)kWarps * kBufCells * 64;   // [128][32 B]
    static constexpr size_t off = gout + (size_t)kThreads * kSliceBytes;    // [128][P] uint32 pairs
    __host__ __device__ static size_t msk(int P) { return off + (size_t)kThreads * P * 4; }
    __host__ __device__ static size_t total(int P) { return (msk(P) + (size_t)kThreads * P * 2 + 15) & ~(size_t)15; }
};

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
// D(16x8, fp32) += A(16x16) * B(16x8)
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

// byte offset of A[cell][pixel] inside a warp buffer: 64-byte rows, 16-byte chunks swizzled
__device__ __forceinline__ uint32_t a_elem_off(int cell, int pixel) {
    return (uint32_t)cell * 64u + ((((uint32_t)pixel >> 3) ^ (((uint32_t)cell >> 1) & 3u)) << 4) + (((uint32_t)pixel & 7u) << 1);
}

// Everything one thread needs to process the sampling points of its pixel.
template <typename T> struct PointCtx {
    int H, W, C, row_stride;        // map extents (elements)
    int oy, ox;                      // window origin in the map
    int band_cell0;                  // first window cell of this warp's band
    int j, half, lane;               // lane in quarter-warp, first 16-byte chunk, lane in warp
    uint32_t win_addr;               // shared address of the window (+ half * 16)
    unsigned char *abuf;             // this warp's A tile
    const T *img;                    // value + image/group base
    float *gv_img;                   // fp32 accumulator + image/group base
    uint4 gq_a, gq_b;                // upstream gradient of the pixel, chunk `half` / other chunk
    float sigma;
};

// One sampling point at (loc_h, loc_w) with mask m: returns the three channel sums
// (grad_mask, grad_offset_x / sigma, grad_offset_y / sigma) and adds the point's coefficients to
// the A tile (window path) or straight to the global accumulator (fallback).
template <typename T>
__device__ __forceinline__ void process_point(const PointCtx<T> &c, float loc_h, float loc_w, float m,
                                              float &gm, float &gx, float &gy) {
    constexpr int E = 8;
    gm = gx = gy = 0.f;
    // range test of the reference (dcnv3_im2col_cuda.cuh:262-263); also rejects NaN
    const bool inside = loc_h > -1.f && loc_w > -1.f && loc_h < (float)c.H && loc_w < (float)c.W;
    if (!inside) return;
    const float fh = floorf(loc_h), fw = floorf(loc_w);
    const float lh = loc_h - fh, lw = loc_w - fw, hh = 1.f - lh, hw = 1.f - lw;
    const int hwin = (int)fh - c.oy, wwin = (int)fw - c.ox;
    const int cb = hwin * kWinW + wwin - c.band_cell0;   // top-left cell inside the band
    if ((unsigned)hwin < (unsigned)(kWinH - 1) && (unsigned)wwin < (unsigned)(kWinW - 1) && cb >= 0 &&
        cb < (kBandH - 1) * kWinW - 1) {
        // ---- window path; out-of-map corners read zeros and their cells are dropped at flush time
        int o[4] = {0, kSliceBytes, kWinW * kSliceBytes, kWinW * kSliceBytes + kSliceBytes};
        const int rho = ((c.j >> 1) - (wwin + 2 * hwin)) & 3;
        rotate4(o, rho);
        const uint32_t tl = c.win_addr + (uint32_t)(hwin * kWinW + wwin) * kSliceBytes;
        float dr[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const uint32_t a = tl + o[t];
            const uint4 qa = lds128(a), qb = lds128(a ^ 16u);
            dr[t] = dot<T>(c.gq_a, qa, 0.f) + dot<T>(c.gq_b, qb, 0.f);
        }
        rotate4(dr, (4 - rho) & 3);   // back to corner order TL, TR, BL, BR
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
        gm = w1 * dr[0] + w2 * dr[1] + w3 * dr[2] + w4 * dr[3];
        gx = m * (hh * (dr[1] - dr[0]) + lh * (dr[3] - dr[2]));
        gy = m * (hw * (dr[2] - dr[0]) + lw * (dr[3] - dr[1]));
        // drop the four coefficients into this pixel's column of A (thread-exclusive; the four
        // corner cells are distinct: read all four, then write all four)
        T *e0 = reinterpret_cast<T *>(c.abuf + a_elem_off(cb, c.lane));
        T *e1 = reinterpret_cast<T *>(c.abuf + a_elem_off(cb + 1, c.lane));
        T *e2 = reinterpret_cast<T *>(c.abuf + a_elem_off(cb + kWinW, c.lane));
        T *e3 = reinterpret_cast<T *>(c.abuf + a_elem_off(cb + kWinW + 1, c.lane));
        const float a0 = to_f32(*e0), a1 = to_f32(*e1), a2 = to_f32(*e2), a3 = to_f32(*e3);
        *e0 = from_f32<T>(a0 + w1 * m);
        *e1 = from_f32<T>(a1 + w2 * m);
        *e2 = from_f32<T>(a2 + w3 * m);
        *e3 = from_f32<T>(a3 + w4 * m);
    } else {
        // ---- fallback: clamped global reads, direct reductions
        const ClampedTap ct = make_clamped_tap(loc_h, loc_w, c.H, c.W);
        const int r_lo = ct.row_lo * c.row_stride, r_hi = ct.row_hi * c.row_stride;
        const int c_lo = ct.col_lo * c.C, c_hi = ct.col_hi * c.C;
        const int at[4] = {r_lo + c_lo, r_lo + c_hi, r_hi + c_lo, r_hi + c_hi};
        const int ea = c.half * E, eb = (c.half ^ 1) * E;
        float dk[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const uint4 qa = __ldg(reinterpret_cast<const uint4 *>(c.img + at[t] + ea));
            const uint4 qb = __ldg(reinterpret_cast<const uint4 *>(c.img + at[t] + eb));
            dk[t] = dot<T>(c.gq_a, qa, 0.f) + dot<T>(c.gq_b, qb, 0.f);
        }
        const float fy_lo = ct.hh * ct.top, fy_hi = ct.lh * ct.bot;
        const float fx_lo = ct.hw * ct.lef, fx_hi = ct.lw * ct.rig;
        const float wk[4] = {fy_lo * fx_lo, fy_lo * fx_hi, fy_hi * fx_lo, fy_hi * fx_hi};
        gm = wk[0] * dk[0] + wk[1] * dk[1] + wk[2] * dk[2] + wk[3] * dk[3];
        gx = m * (fy_lo * (ct.rig * dk[1] - ct.lef * dk[0]) + fy_hi * (ct.rig * dk[3] - ct.lef * dk[2]));
        gy = m * (fx_lo * (ct.bot * dk[2] - ct.top * dk[0]) + fx_hi * (ct.bot * dk[3] - ct.top * dk[1]));
        float ga[E], gb[E];
        unpack<T>(c.gq_a, ga);
        unpack<T>(c.gq_b, gb);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float cf = wk[t] * m;
            if (cf != 0.f) {
                float *dst = c.gv_img + at[t];
#pragma unroll
                for (int e = 0; e < E; e += 4) {
                    red_add4(dst + ea + e, make_float4(cf * ga[e], cf * ga[e + 1], cf * ga[e + 2], cf * ga[e + 3]));
                    red_add4(dst + eb + e, make_float4(cf * gb[e], cf * gb[e + 1], cf * gb[e + 2], cf * gb[e + 3]));
                }
            }
        }
    }
}

// band = A (272 x 32) * grad_out (32 x 16) for one warp; the fp32 result overwrites A in place,
// m-tile by m-tile ([cell][16 ch] rows of 64 bytes).
template <typename T>
__device__ __forceinline__ void band_mma(unsigned char *abuf, const unsigned char *s_gout, int warp, int lane) {
    const uint32_t a_base = smem_u32(abuf);
    uint32_t bf[2][4];   // [k-step][{n0:k0-7, n0:k8-15, n1:k0-7, n1:k8-15}]
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        const int px = warp * 32 + ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldmatrix_x4_trans(bf[ks], smem_u32(s_gout) + px * kSliceBytes + (lane >> 4) * 16);
    }
    const int r_in = (lane & 7) + ((lane >> 3) & 1) * 8;   // row inside an m-tile this lane addresses
    const int kc_in = lane >> 4;                            // 16-byte k chunk (0/1) inside a k-step
    // A fragments are fetched one m-tile ahead of the HMMAs that consume them
    auto load_a = [&](int mt, uint32_t (&af)[2][4]) {
        const int row = mt * 16 + r_in;
        const uint32_t sw = ((uint32_t)row >> 1) & 3u;
        ldmatrix_x4(af[0], a_base + (uint32_t)row * 64u + (((uint32_t)kc_in ^ sw) << 4));
        ldmatrix_x4(af[1], a_base + (uint32_t)row * 64u + (((uint32_t)(2 + kc_in) ^ sw) << 4));
    };
    uint32_t cur[2][4], nxt[2][4];
    load_a(0, cur);
#pragma unroll
    for (int mt = 0; mt < kMTiles; ++mt) {
        if (mt + 1 < kMTiles) load_a(mt + 1, nxt);
        float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            mma16816(acc0, cur[ks], bf[ks][0], bf[ks][1], T());
            mma16816(acc1, cur[ks], bf[ks][2], bf[ks][3], T());
        }
        // the m-tile's 16 rows are dead now (their fragments are in registers): overwrite them
        float *r0 = reinterpret_cast<float *>(abuf) + (size_t)(mt * 16 + (lane >> 2)) * kCh + 2 * (lane & 3);
        *reinterpret_cast<float2 *>(r0) = make_float2(acc0[0], acc0[1]);
        *reinterpret_cast<float2 *>(r0 + 8) = make_float2(acc1[0], acc1[1]);
        *reinterpret_cast<float2 *>(r0 + 8 * kCh) = make_float2(acc0[2], acc0[3]);
        *reinterpret_cast<float2 *>(r0 + 8 * kCh + 8) = make_float2(acc1[2], acc1[3]);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int e = 0; e < 4; ++e) cur[ks][e] = nxt[ks][e];
    }
}

// Flush: add the four warps' bands to the global fp32 accumulator and (ZERO) leave them cleared.
// Thread <-> one (window column, 16-byte piece): 72 of the 128 threads walk the 26 window rows
// with a fully unrolled loop, so which bands cover a row is known at compile time and the address
// arithmetic is one pointer increment per row (the generic per-cell loop cost 91 instructions per
// 16 bytes flushed and 41 % of the kernel's run time, profiles/README.md).
template <bool ZERO>
__device__ __forceinline__ void flush_bands(float *bands, float *gimg /* accumulator at (n,0,0,ch0) */,
                                            int oy, int ox, int H, int W, int row_stride, int C, int tid) {
    if (tid >= kWinW * 4) return;
    const int wx = tid >> 2, piece = tid & 3;
    const int x = ox + wx;
    const bool x_ok = (unsigned)x < (unsigned)W;
    float *col = bands + wx * kCh + piece * 4;                       // (band 0, row 0) of this column
    float *dst = gimg + (ptrdiff_t)oy * row_stride + (ptrdiff_t)x * C + piece * 4;
    // pass 1: loads, sums and reductions only -- no store in between, so the 60 LDS.128 of a
    // column are independent and pipeline (interleaving the clearing stores serialised them)
#pragma unroll
    for (int wy = 0; wy < kWinH; ++wy) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const int by = wy - w * kRowsPerWarp;                    // compile-time after unrolling
            if (by >= 0 && by < kBandH) {
                const float4 t = *reinterpret_cast<const float4 *>(col + ((size_t)w * kBufCells + by * kWinW) * kCh);
                v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
            }
        }
        if (x_ok && (unsigned)(oy + wy) < (unsigned)H && (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f))
            red_add4(dst, v);
        dst += row_stride;
    }
    // pass 2: clear what was read
    if (ZERO) {
#pragma unroll
        for (int w = 0; w < kWarps; ++w)
#pragma unroll
            for (int by = 0; by < kBandH; ++by)
                if (by + w * kRowsPerWarp < kWinH)
                    *reinterpret_cast<float4 *>(col + ((size_t)w * kBufCells + by * kWinW) * kCh) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <typename T, int KH, int KW>
__global__ void __launch_bounds__(kThreads)
bwd_mma(const __grid_constant__ CUtensorMap tmap, const T *__restrict__ value,
        const T *__restrict__ offset, const T *__restrict__ mask, const T *__restrict__ grad_out,
        float *__restrict__ gv_acc, T *__restrict__ grad_offset, T *__restrict__ grad_mask,
        const Geom q, const Params tp) {
    constexpr int E = 8;                    // channels per 16-byte chunk
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;

    const int kh = KH ? KH : q.kh, kw = KW ? KW : q.kw;
    const int P = kh * kw;
    unsigned char *win = smem + Layout::win;
    unsigned char *s_buf = smem + Layout::buf;
    unsigned char *s_gout = smem + Layout::gout;
    uint32_t *s_off = reinterpret_cast<uint32_t *>(smem + Layout::off);
    T *s_msk = reinterpret_cast<T *>(smem + Layout::msk(P));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_x = blockIdx.x % tp.tiles_x, tile_y = blockIdx.x / tp.tiles_x;
    const int g = blockIdx.y;
    const int n = tp.n0 + blockIdx.z;
    const int wo0 = tile_x * kTileW, ho0 = tile_y * kTileH;
    const int wo = wo0 + (tid % kTileW), ho = ho0 + (tid / kTileW);
    const bool live = wo < q.Wo && ho < q.Ho;
    const int ox = wo0 * q.sw + tp.ox_rel, oy = ho0 * q.sh + tp.oy_rel;
    const int C = q.G * q.gc;
    const int row_stride = q.W * C;
    const int ch0 = g * q.gc;

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, kCells * kSliceBytes);
        tma_load_4d(win, &tmap, &bar, ch0, ox, oy, n);
    }

    // ---- while the box is in flight: zero the A tiles, stage offsets / masks / grad_out
    {
        uint4 *z = reinterpret_cast<uint4 *>(s_buf);
        for (int i = tid; i < kWarps * kBufCells * 4; i += kThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
        const size_t img_pix = (size_t)n * q.Ho * q.Wo;
        stage_offsets_masks<T, KH * KW, kThreads, kTileW>(offset, mask, s_off, s_msk, P, tid, wo0, ho0, q.Wo, q.Ho, q.G, g, img_pix);
        for (int idx = tid; idx < kThreads * 2; idx += kThreads) {   // 2 x 16-byte chunks per pixel
            const int px = idx >> 1, c = idx & 1;
            const int w = wo0 + (px % kTileW), h = ho0 + (px / kTileW);
            uint4 v = make_uint4(0u, 0u, 0u, 0u);   // pixels outside the map contribute nothing
            if (w < q.Wo && h < q.Ho)
                v = __ldg(reinterpret_cast<const uint4 *>(grad_out + (img_pix + (size_t)h * q.Wo + w) * C + ch0 + c * E));
            *reinterpret_cast<uint4 *>(s_gout + px * kSliceBytes + c * 16) = v;
        }
    }

    const int j = tid & 7;                         // lane within the quarter-warp
    const int half = j & 1;                        // 16-byte chunk read first
    const float base_w = axis_base(wo, kw, q.sw, q.pw, q.dw, q.sigma);
    const float base_h = axis_base(ho, kh, q.sh, q.ph, q.dh, q.sigma);
    const uint32_t win_addr = smem_u32(win) + half * 16;
    const size_t img_base = (size_t)n * q.H * row_stride + ch0;   // element index of (n, 0, 0, ch0)
    const T *img = value + img_base;
    unsigned char *abuf = s_buf + (size_t)warp * kBufCells * 64;  // this warp's A tile / result band
    const int band_cell0 = warp * kRowsPerWarp * kWinW;           // first window cell of the band

    __syncthreads();
    mbar_wait(&bar, 0);

    // upstream gradient of this thread's pixel: chunk `half` and the other chunk, packed
    const uint4 gq_a = *reinterpret_cast<const uint4 *>(s_gout + tid * kSliceBytes + half * 16);
    const uint4 gq_b = *reinterpret_cast<const uint4 *>(s_gout + tid * kSliceBytes + (half ^ 1) * 16);

    // ------------------------------------------------------------------ gather + A build
    if (live) {
        PointCtx<T> c;
        c.H = q.H; c.W = q.W; c.C = C; c.row_stride = row_stride; c.oy = oy; c.ox = ox;
        c.band_cell0 = band_cell0; c.j = j; c.half = half; c.lane = lane; c.win_addr = win_addr;
        c.abuf = abuf; c.img = img; c.gv_img = gv_acc + img_base; c.gq_a = gq_a; c.gq_b = gq_b;
        c.sigma = q.sigma;
#pragma unroll 1
        for (int i = 0; i < kw; ++i) {
#pragma unroll
            for (int jj = 0; jj < (KH ? KH : 8); ++jj) {
                if (jj >= kh) break;
                const int p = i * kh + jj;
                const float2 d = unpack2(s_off[tid * P + p], T());
                const float m = to_f32(s_msk[tid * P + p]);
                const float loc_w = base_w + ((float)(i * q.dw) + d.x) * q.sigma;
                const float loc_h = base_h + ((float)(jj * q.dh) + d.y) * q.sigma;
                float gm, gx, gy;
                process_point<T>(c, loc_h, loc_w, m, gm, gx, gy);
                s_off[tid * P + p] = pack2(q.sigma * gx, q.sigma * gy, T());
                s_msk[tid * P + p] = from_f32<T>(gm);
            }
        }
    }
    __syncwarp();
    band_mma<T>(abuf, s_gout, warp, lane);
    __syncthreads();

    // ---- grad_offset / grad_mask: coalesced write-out of the staged values
    {
        const size_t img_pix = (size_t)n * q.Ho * q.Wo;
        for (int idx = tid; idx < kThreads * P; idx += kThreads) {
            const int px = idx / P, p = idx - px * P;
            const int w = wo0 + (px % kTileW), h = ho0 + (px / kTileW);
            if (w < q.Wo && h < q.Ho) {
                const size_t pgi = (img_pix + (size_t)h * q.Wo + w) * q.G + g;
                reinterpret_cast<uint32_t *>(grad_offset)[pgi * P + p] = s_off[idx];
                grad_mask[pgi * P + p] = s_msk[idx];
            }
        }
    }
    flush_bands<false>(reinterpret_cast<float *>(s_buf), gv_acc + img_base, oy, ox, q.H, q.W, row_stride, C, tid);
}

// ================================================================================================
// Persistent form (3x3 kernels): a CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... so
// that every global-memory latency is taken off the critical path:
//   * the NEXT tile's offsets / masks / grad_out are loaded into registers while the current tile
//     is being gathered and parked in the second staging buffer at the end of the iteration;
//   * the NEXT tile's value window is requested (TMA) as soon as the last warp has finished
//     gathering from the window, and lands during the tensor-core product and the flush;
//   * the flush leaves the A tiles / bands zeroed, so there is no separate clearing pass.
// In the one-tile-per-CTA kernel above these waits were 26 % of the run time at 2 CTAs per SM
// (profiles/README.md, bwd_mma source view).
struct PersistParams {
    int ox_rel, oy_rel, tiles_x, tiles_xy, total_tiles;
};

struct TileAt {
    int n, g, wo0, ho0, ox, oy;
};
__device__ __forceinline__ TileAt decode_tile(int t, const Geom &q, const PersistParams &pp) {
    TileAt a;
    const int txy = t % pp.tiles_xy, r = t / pp.tiles_xy;
    a.g = r % q.G;
    a.n = r / q.G;
    a.wo0 = (txy % pp.tiles_x) * kTileW;
    a.ho0 = (txy / pp.tiles_x) * kTileH;
    a.ox = a.wo0 * q.sw + pp.ox_rel;
    a.oy = a.ho0 * q.sh + pp.oy_rel;
    return a;
}

// one tile's staging data, in registers (P = 9)
template <typename T> struct StageRegs {
    uint32_t off[9];
    T msk[9];
    uint4 go[2];
};
template <typename T>
__device__ __forceinline__ void stage_load(StageRegs<T> &r, const TileAt &a, const T *__restrict__ offset,
                                           const T *__restrict__ mask, const T *__restrict__ grad_out,
                                           const Geom &q, int tid) {
    const size_t img_pix = (size_t)a.n * q.Ho * q.Wo;
    const size_t base = (img_pix * q.G + a.g) * 9;
    const uint32_t *obase = reinterpret_cast<const uint32_t *>(offset) + base;
    const T *mbase = mask + base;
#pragma unroll
    for (int it = 0; it < 9; ++it) {
        const unsigned idx = tid + it * kThreads;
        const unsigned px = idx / 9, p = idx - px * 9;
        const unsigned w = a.wo0 + (px % kTileW), h = a.ho0 + (px / kTileW);
        r.off[it] = 0u;
        r.msk[it] = from_f32<T>(0.f);
        if (w < (unsigned)q.Wo && h < (unsigned)q.Ho) {
            const unsigned rel = (h * q.Wo + w) * (unsigned)(q.G * 9) + p;
            r.off[it] = __ldg(obase + rel);
            r.msk[it] = __ldg(mbase + rel);
        }
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int idx = tid + it * kThreads, px = idx >> 1, c = idx & 1;
        const int w = a.wo0 + (px % kTileW), h = a.ho0 + (px / kTileW);
        r.go[it] = make_uint4(0u, 0u, 0u, 0u);   // pixels outside the map contribute nothing
        if (w < q.Wo && h < q.Ho)
            r.go[it] = __ldg(reinterpret_cast<const uint4 *>(
                grad_out + (img_pix + (size_t)h * q.Wo + w) * (q.G * q.gc) + a.g * q.gc + c * 8));
    }
}
template <typename T>
__device__ __forceinline__ void stage_store(const StageRegs<T> &r, uint32_t *s_off, T *s_msk,
                                            unsigned char *s_gout, int tid) {
#pragma unroll
    for (int it = 0; it < 9; ++it) {
        s_off[tid + it * kThreads] = r.off[it];
        s_msk[tid + it * kThreads] = r.msk[it];
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int idx = tid + it * kThreads;
        *reinterpret_cast<uint4 *>(s_gout + (idx >> 1) * kSliceBytes + (idx & 1) * 16) = r.go[it];
    }
}

struct PLayout {   // bytes
    static constexpr size_t win = 0;
    static constexpr size_t buf = win + (size_t)kCells * kSliceBytes;
    static constexpr size_t stage = buf + (size_t)kWarps * kBufCells * 64;
    static constexpr size_t stage_bytes = (size_t)kThreads * kSliceBytes + (size_t)kThreads * 9 * 6;   // gout + off + msk
    static constexpr size_t total = stage + 2 * stage_bytes;
};

template <typename T>
__global__ void __launch_bounds__(kThreads)
bwd_mma_persistent(const __grid_constant__ CUtensorMap tmap, const T *__restrict__ value,
                   const T *__restrict__ offset, const T *__restrict__ mask,
                   const T *__restrict__ grad_out, float *__restrict__ gv_acc,
                   T *__restrict__ grad_offset, T *__restrict__ grad_mask, const Geom q,
                   const PersistParams pp) {
    constexpr int P = 9;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    unsigned char *win = smem + PLayout::win;
    unsigned char *s_buf = smem + PLayout::buf;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = q.G * q.gc, row_stride = q.W * C;
    const int j = tid & 7, half = j & 1;
    unsigned char *abuf = s_buf + (size_t)warp * kBufCells * 64;
    const int band_cell0 = warp * kRowsPerWarp * kWinW;
    auto stage_gout = [&](int b) { return smem + PLayout::stage + (size_t)b * PLayout::stage_bytes; };
    auto stage_off = [&](int b) { return reinterpret_cast<uint32_t *>(stage_gout(b) + kThreads * kSliceBytes); };
    auto stage_msk = [&](int b) { return reinterpret_cast<T *>(stage_off(b) + kThreads * P); };

    int t = blockIdx.x;
    if (t >= pp.total_tiles) return;
    TileAt cur = decode_tile(t, q, pp);
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    {   // prologue: first tile's staging data, zero the A tiles once
        StageRegs<T> r;
        stage_load<T>(r, cur, offset, mask, grad_out, q, tid);
        uint4 *z = reinterpret_cast<uint4 *>(s_buf);
        for (int i = tid; i < kWarps * kBufCells * 4; i += kThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
        stage_store<T>(r, stage_off(0), stage_msk(0), stage_gout(0), tid);
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, kCells * kSliceBytes);
        tma_load_4d(win, &tmap, &bar, cur.g * q.gc, cur.ox, cur.oy, cur.n);
    }

    for (int it = 0;; ++it) {
        const int b = it & 1;
        const int t_next = t + gridDim.x;
        const bool has_next = t_next < pp.total_tiles;
        TileAt nxt = cur;
        StageRegs<T> nr;
        if (has_next) {
            nxt = decode_tile(t_next, q, pp);
            stage_load<T>(nr, nxt, offset, mask, grad_out, q, tid);   // in flight during the gather
        }
        uint32_t *s_off = stage_off(b);
        T *s_msk = stage_msk(b);
        unsigned char *s_gout = stage_gout(b);
        const int wo = cur.wo0 + (tid % kTileW), ho = cur.ho0 + (tid / kTileW);
        const bool live = wo < q.Wo && ho < q.Ho;
        const size_t img_base = (size_t)cur.n * q.H * row_stride + cur.g * q.gc;

        mbar_wait(&bar, it & 1);   // value window of the current tile has landed

        // ---------------------------------------------------------------- gather + A build
        if (live) {
            PointCtx<T> c;
            c.H = q.H; c.W = q.W; c.C = C; c.row_stride = row_stride; c.oy = cur.oy; c.ox = cur.ox;
            c.band_cell0 = band_cell0; c.j = j; c.half = half; c.lane = lane;
            c.win_addr = smem_u32(win) + half * 16;
            c.abuf = abuf; c.img = value + img_base; c.gv_img = gv_acc + img_base;
            c.gq_a = *reinterpret_cast<const uint4 *>(s_gout + tid * kSliceBytes + half * 16);
            c.gq_b = *reinterpret_cast<const uint4 *>(s_gout + tid * kSliceBytes + (half ^ 1) * 16);
            c.sigma = q.sigma;
            const float base_w = axis_base(wo, 3, q.sw, q.pw, q.dw, q.sigma);
            const float base_h = axis_base(ho, 3, q.sh, q.ph, q.dh, q.sigma);
#pragma unroll 1
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
                    const int p = i * 3 + jj;
                    const float2 d = unpack2(s_off[tid * P + p], T());
                    const float m = to_f32(s_msk[tid * P + p]);
                    const float loc_w = base_w + ((float)(i * q.dw) + d.x) * q.sigma;
                    const float loc_h = base_h + ((float)(jj * q.dh) + d.y) * q.sigma;
                    float gm, gx, gy;
                    process_point<T>(c, loc_h, loc_w, m, gm, gx, gy);
                    s_off[tid * P + p] = pack2(q.sigma * gx, q.sigma * gy, T());
                    s_msk[tid * P + p] = from_f32<T>(gm);
                }
            }
        }
        __syncthreads();   // every warp is done with the window (and with its own A columns)
        if (has_next && tid == 0) {
            // order the generic-proxy reads of the window before the async-proxy overwrite
            fence_proxy_async();
            mbar_expect_tx(&bar, kCells * kSliceBytes);
            tma_load_4d(win, &tmap, &bar, nxt.g * q.gc, nxt.ox, nxt.oy, nxt.n);
        }
        // ---------------------------------------------------------------- band = A x grad_out
        band_mma<T>(abuf, s_gout, warp, lane);
        __syncthreads();

        // ---- grad_offset / grad_mask: coalesced write-out of the staged results
        {
            const size_t img_pix = (size_t)cur.n * q.Ho * q.Wo;
            const size_t base = (img_pix * q.G + cur.g) * P;
            uint32_t *ob = reinterpret_cast<uint32_t *>(grad_offset) + base;
            T *mb = grad_mask + base;
#pragma unroll
            for (int k = 0; k < P; ++k) {
                const unsigned idx = tid + k * kThreads;
                const unsigned px = idx / P, p = idx - px * P;
                const unsigned w = cur.wo0 + (px % kTileW), h = cur.ho0 + (px / kTileW);
                if (w < (unsigned)q.Wo && h < (unsigned)q.Ho) {
                    const unsigned rel = (h * q.Wo + w) * (unsigned)(q.G * P) + p;
                    ob[rel] = s_off[idx];
                    mb[rel] = s_msk[idx];
                }
            }
        }
        // ---- flush (leaves the bands zeroed: they are the next tile's A tiles; band rows below
        //      the window and the two pad cells per band never receive a coefficient)
        flush_bands<true>(reinterpret_cast<float *>(s_buf), gv_acc + img_base, cur.oy, cur.ox, q.H, q.W, row_stride, C, tid);
        if (!has_next) break;
        stage_store<T>(nr, stage_off(b ^ 1), stage_msk(b ^ 1), stage_gout(b ^ 1), tid);
        __syncthreads();   // next staging buffer visible, bands zeroed
        cur = nxt;
        t = t_next;
    }
}

template <typename T>
static bool launch_typed(const void *value, const void *offset, const void *mask, const void *grad_out,
                         float *gv_acc, void *grad_offset, void *grad_mask, const Geom &q, int dtype,
                         cudaStream_t stream, cudaError_t *err) {
    if (q.gc != kCh) return false;   // one 16-channel slice == one group (channel sums stay in the CTA)
    if (q.kh > 8) return false;
    if (((uintptr_t)value | (uintptr_t)grad_out | (uintptr_t)gv_acc) % 16) return false;
    if (((uintptr_t)offset | (uintptr_t)grad_offset) % 4) return false;
    const float span_w = (kTileW - 1) * q.sw + (q.kw - 1) * q.dw * q.sigma;
    const float span_h = (kTileH - 1) * q.sh + (q.kh - 1) * q.dh * q.sigma;
    if (!(q.sigma > 0.f) || span_w + 4 > kWinW - 2 || span_h + 4 > kWinH - 2) return false;
    const int C = q.G * q.gc;
    const int P = q.kh * q.kw;
    const size_t smem = Layout::total(P);
    if (smem > 110 * 1024 || q.G > 65535) return false;
    CUtensorMap tmap;
    if (!make_nhwc_tensor_map(&tmap, value, dtype, q.N, q.H, q.W, C, kCh, kWinW, kWinH)) return false;

    Params tp;
    const int cw = (q.dw * (q.kw - 1)) >> 1, chh = (q.dh * (q.kh - 1)) >> 1;
    const float a_w = (float)(cw - q.pw) - cw * q.sigma, a_h = (float)(chh - q.ph) - chh * q.sigma;
    tp.ox_rel = (int)std::floor(a_w + 0.5f * span_w - 0.5f * (kWinW - 2));
    tp.oy_rel = (int)std::floor(a_h + 0.5f * span_h - 0.5f * (kWinH - 2));
    tp.tiles_x = (q.Wo + kTileW - 1) / kTileW;
    const int tiles_y = (q.Ho + kTileH - 1) / kTileH;
    const T *v = static_cast<const T *>(value), *o = static_cast<const T *>(offset),
            *m = static_cast<const T *>(mask), *go = static_cast<const T *>(grad_out);
    T *goff = static_cast<T *>(grad_offset), *gmsk = static_cast<T *>(grad_mask);
    const bool k33 = q.kh == 3 && q.kw == 3;
    const char *pe = std::getenv("DCNV3_BWD_PERSIST");   // development knob: 0 = one tile per CTA
    if (k33 && !(pe && pe[0] == '0') && (long long)tp.tiles_x * tiles_y * q.G * q.N < (1LL << 31)) {
        static int num_sms = 0;
        if (num_sms == 0) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        }
        PersistParams pp;
        pp.ox_rel = tp.ox_rel; pp.oy_rel = tp.oy_rel; pp.tiles_x = tp.tiles_x;
        pp.tiles_xy = tp.tiles_x * tiles_y;
        pp.total_tiles = pp.tiles_xy * q.G * q.N;
        const int ctas = std::min(pp.total_tiles, 2 * num_sms);   // 2 resident CTAs per SM (107 KB each)
        cudaFuncSetAttribute(bwd_mma_persistent<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PLayout::total);
        bwd_mma_persistent<T><<<ctas, kThreads, PLayout::total, stream>>>(tmap, v, o, m, go, gv_acc, goff, gmsk, q, pp);
        *err = cudaGetLastError();
        return true;
    }
    if (k33) cudaFuncSetAttribute(bwd_mma<T, 3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else cudaFuncSetAttribute(bwd_mma<T, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int n0 = 0; n0 < q.N; n0 += 65535) {
        tp.n0 = n0;
        const dim3 grid((unsigned)(tp.tiles_x * tiles_y), (unsigned)q.G, (unsigned)std::min(65535, q.N - n0));
        if (k33) bwd_mma<T, 3, 3><<<grid, kThreads, smem, stream>>>(tmap, v, o, m, go, gv_acc, goff, gmsk, q, tp);
        else bwd_mma<T, 0, 0><<<grid, kThreads, smem, stream>>>(tmap, v, o, m, go, gv_acc, goff, gmsk, q, tp);
    }
    *err = cudaGetLastError();
    return true;
}

}  // namespace mma

// gv_acc: zero-initialised fp32 accumulator with the shape of value.  16-bit I/O with
// group_channels == 16 only; returns false otherwise (caller uses the direct kernel).
bool try_launch_backward_mma(const void *value, const void *offset, const void *mask,
                             const void *grad_out, float *gv_acc, void *grad_offset, void *grad_mask,
                             const Geom &q, int dtype, cudaStream_t stream, cudaError_t *err) {
    const char *e = std::getenv("DCNV3_BWD");   // development knob: DCNV3_BWD=scatter|tile select other kernels
    if (e && (e[0] == 's' || e[0] == 't')) return false;
    if ((long long)q.N * q.Ho * q.Wo == 0) return false;
    if (dtype == 1) return mma::launch_typed<__half>(value, offset, mask, grad_out, gv_acc, grad_offset, grad_mask, q, dtype, stream, err);
    if (dtype == 2) return mma::launch_typed<__nv_bfloat16>(value, offset, mask, grad_out, gv_acc, grad_offset, grad_mask, q, dtype, stream, err);
    return false;
}

}  // namespace dcnv3
