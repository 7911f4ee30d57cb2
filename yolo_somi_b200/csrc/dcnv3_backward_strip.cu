// dcnv3_backward_strip.cu -- DCNv3 core backward for 16-bit I/O, group_channels == 16, 3x3 / stride 1 /
// dilation 1: the tensor-core scatter of dcnv3_backward_mma.cu with the result kept in REGISTERS.
//
// Why (profiles/README.md, r1_v3): in bwd_mma_persistent the product  band = A x grad_out  is
// written to shared memory (13.9 M wavefronts), summed over the four warps' overlapping bands and
// flushed by 72 of 128 threads (7.6 M + 7.4 M wavefronts, 25 % of the run time, three CTA barriers
// per tile).  The kernel sits on the shared-memory pipe (l1tex 69 %), so those bytes are the cost.
//
// How: a warp owns an 8-pixel-wide strip and walks DOWN it four rows at a time.  A step's 32
// pixels (lane <-> pixel) reach the cells of a 14-row x 16-column band; the next step's band is
// the same one shifted by four rows.  So the fp32 accumulator of the product lives in the HMMA
// accumulator registers (14 m-tiles x 2 n-tiles x 4), is shifted by four m-tiles after every step,
// and the four rows that no later step can touch leave straight from the registers to the global
// fp32 accumulator as 64-bit vector reductions.  Nothing of grad_value ever passes through shared
// memory, there is no flush pass and no CTA barrier except the one that recycles the TMA window.
//
//   per step and warp:  wait(cp.async staging) -> gather (window, rotated LDS.128, exact FHFMA dot
//   products) + A build (thread-exclusive 16-bit read-modify-writes) -> grad_offset / grad_mask
//   written out coalesced -> prefetch the next step's offsets / masks / grad_out with cp.async
//   (no registers, lands during the HMMAs) -> 14 x (2 ldmatrix + 4 HMMA + clear the m-tile)
//   -> reductions of the finished rows, accumulator shift.
//
// A CTA is two warps (a 16x16 tile of one (image, group), one 26x26-pixel TMA window, the same box
// as the forward); four CTAs are resident per SM and each warp runs its phases independently of
// the others, so tensor, shared-memory and SIMT phases of different warps overlap.
//
// The band is 16 columns wide (one m-tile per band row), which covers offsets of +-3 sigma px
// around the kernel taps; a point beyond that but inside the window still takes its dot products
// from the window and hands its four coefficients to a small per-warp spill list that the warp
// turns into reductions cooperatively; a point outside the window takes the clamped global path.
//
// Semantics: reference dcnv3_im2col_cuda.cuh:82-147 (col2im bilinear), :278-370 (channel sums).
#include "dcnv3_common.cuh"
#include "dcnv3_launch.h"
#include "dcnv3_tma.cuh"
#include "dcnv3_strip_io.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace dcnv3 {
namespace strip {

constexpr int kSteps = 4;
constexpr int kWarps = 2;
constexpr int kTileW = kStripW * kWarps, kTileH = kPatchH * kSteps;   // 16 x 16 output pixels
constexpr int kThreads = kWarps * 32;
constexpr int kWinW = 26, kWinH = 26;            // value window (2 mod 4: conflict-free corners)
constexpr int kBandW = 16, kBandH = 12;          // cells one step of one warp scatters into (taps +- 3 px)
constexpr int kMTiles = kBandH;                  // one m-tile (16 cells) per band row
constexpr int kSpillCap = 32;
static_assert(kWinW % 4 == 2, "window width must be 2 mod 4");

constexpr int kWinBytes = kWinW * kWinH * kSliceBytes;          // 21632
constexpr int kABytes = kBandH * kBandW * 64;                   // 12288 per warp
constexpr int kSpillBytes = 16 + kSpillCap * 32;                // 272 per warp
constexpr int kSmemBytes = kWinBytes + kWarps * (kABytes + kStageBytes + kSpillBytes) + kIoTblBytes;

struct Params {
    int ox_rel, oy_rel;      // window origin relative to the tile origin
    int bxw, byw;            // band origin of strip 0 / step 0 in window coordinates
    int tiles_x, tiles_xy, total_tiles;
    unsigned long long mask_bytes;   // size of the mask tensor (guards the word-granular staging)
};

struct TileAt {
    int n, g, wo0, ho0, ox, oy;
};
__device__ __forceinline__ TileAt decode_tile(int t, const Geom &q, const Params &pp) {
    TileAt a;
    const int txy = t % pp.tiles_xy, r = t / pp.tiles_xy;
    a.g = r % q.G;
    a.n = r / q.G;
    a.wo0 = (txy % pp.tiles_x) * kTileW;
    a.ho0 = (txy / pp.tiles_x) * kTileH;
    a.ox = a.wo0 + pp.ox_rel;
    a.oy = a.ho0 + pp.oy_rel;
    return a;
}

struct SpillEntry {   // 32 bytes
    int h0, w0, lane, pad;
    float c[4];
};

template <typename V> __device__ __forceinline__ void rotate4(V (&x)[4], int r) {  // out[t] = in[(t+r)&3]
    if (r & 1) { const V t = x[0]; x[0] = x[1]; x[1] = x[2]; x[2] = x[3]; x[3] = t; }
    if (r & 2) { V t = x[0]; x[0] = x[2]; x[2] = t; t = x[1]; x[1] = x[3]; x[3] = t; }
}

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


// byte offset of A[cell][pixel] inside a warp's A tile: 64-byte rows, 16-byte chunks swizzled
__device__ __forceinline__ uint32_t a_elem_off(int cell, int pixel) {
    return (uint32_t)cell * 64u + ((((uint32_t)pixel >> 3) ^ (((uint32_t)cell >> 1) & 3u)) << 4) + (((uint32_t)pixel & 7u) << 1);
}

// ------------------------------------------------------------------------------------------------
// Rare paths, kept out of line so that the hot loop stays small (instruction cache) and is not
// register-allocated for them.  Everything they need that changes once per tile lives in a
// per-thread context in local memory; per-step values come as arguments.
template <typename T> struct SlowCtx {
    int H, W, C, row_stride;
    int oy, ox;                      // window origin in the map
    int lane, j, half;
    uint32_t win_addr;               // shared address of the window (+ half * 16)
    unsigned char *abuf;             // this warp's A tile
    unsigned char *spill;            // this warp's spill list
    const unsigned char *s_gout;     // this warp's staged grad_out [32][32 B]
    const T *img;                    // value + (n, 0, 0, ch0)
    float *gv_img;                   // fp32 accumulator + (n, 0, 0, ch0)
};

// coefficient x grad_out of one lane's pixel straight to the global accumulator
template <typename T>
__device__ __forceinline__ void direct_scatter(const SlowCtx<T> &c, const uint4 &gq_a, const uint4 &gq_b,
                                               const int (&at)[4], const float (&cf)[4]) {
    constexpr int E = 8;
    float ga[E], gb[E];
    unpack<T>(gq_a, ga);
    unpack<T>(gq_b, gb);
    const int ea = c.half * E, eb = (c.half ^ 1) * E;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        if (cf[t] != 0.f) {
            float *dst = c.gv_img + at[t];
#pragma unroll
            for (int e = 0; e < E; e += 4) {
                red_add4(dst + ea + e, make_float4(cf[t] * ga[e], cf[t] * ga[e + 1], cf[t] * ga[e + 2], cf[t] * ga[e + 3]));
                red_add4(dst + eb + e, make_float4(cf[t] * gb[e], cf[t] * gb[e + 1], cf[t] * gb[e + 2], cf[t] * gb[e + 3]));
            }
        }
    }
}

// A point inside the window whose corner block leaves the band: its four coefficients go to the
// warp's spill list (or, if that is full, straight to the global accumulator).
template <typename T>
__device__ __noinline__ void spill_push(const SlowCtx<T> *cp, int h0, int w0, float c0, float c1, float c2, float c3) {
    const SlowCtx<T> &c = *cp;
    unsigned *cnt = reinterpret_cast<unsigned *>(c.spill);
    const unsigned pos = atomicAdd(cnt, 1u);
    if (pos < (unsigned)kSpillCap) {
        SpillEntry *e = reinterpret_cast<SpillEntry *>(c.spill + 16) + pos;
        e->h0 = h0; e->w0 = w0; e->lane = c.lane; e->pad = 0;
        e->c[0] = c0; e->c[1] = c1; e->c[2] = c2; e->c[3] = c3;
    } else {
        const uint4 gq_a = *reinterpret_cast<const uint4 *>(c.s_gout + c.lane * kSliceBytes + c.half * 16);
        const uint4 gq_b = *reinterpret_cast<const uint4 *>(c.s_gout + c.lane * kSliceBytes + (c.half ^ 1) * 16);
        const bool top = h0 >= 0, bot = h0 + 1 < c.H, lef = w0 >= 0, rig = w0 + 1 < c.W;
        const int at[4] = {h0 * c.row_stride + w0 * c.C, h0 * c.row_stride + (w0 + 1) * c.C,
                           (h0 + 1) * c.row_stride + w0 * c.C, (h0 + 1) * c.row_stride + (w0 + 1) * c.C};
        const float cg[4] = {top && lef ? c0 : 0.f, top && rig ? c1 : 0.f, bot && lef ? c2 : 0.f, bot && rig ? c3 : 0.f};
        direct_scatter<T>(c, gq_a, gq_b, at, cg);
    }
}

// One sampling point, any location: channel sums (grad_mask, grad_offset / sigma) and the point's
// four coefficients into the A tile / the spill list / the global accumulator.
template <typename T>
__device__ __noinline__ void slow_point(const SlowCtx<T> *cp, int band_row0, int band_col0, float loc_h,
                                        float loc_w, float m, float *res /* gm, gx, gy */) {
    const SlowCtx<T> &c = *cp;
    res[0] = res[1] = res[2] = 0.f;
    // range test of the reference (dcnv3_im2col_cuda.cuh:262-263); also rejects NaN
    const bool inside = loc_h > -1.f && loc_w > -1.f && loc_h < (float)c.H && loc_w < (float)c.W;
    if (!inside) return;
    const uint4 gq_a = *reinterpret_cast<const uint4 *>(c.s_gout + c.lane * kSliceBytes + c.half * 16);
    const uint4 gq_b = *reinterpret_cast<const uint4 *>(c.s_gout + c.lane * kSliceBytes + (c.half ^ 1) * 16);
    const float fh = floorf(loc_h), fw = floorf(loc_w);
    const float lh = loc_h - fh, lw = loc_w - fw, hh = 1.f - lh, hw = 1.f - lw;
    const int h0 = (int)fh, w0 = (int)fw;
    const int hwin = h0 - c.oy, wwin = w0 - c.ox;
    if ((unsigned)hwin < (unsigned)(kWinH - 1) && (unsigned)wwin < (unsigned)(kWinW - 1)) {
        // ---- window path; out-of-map corners read zeros (TMA fill)
        int o[4] = {0, kSliceBytes, kWinW * kSliceBytes, kWinW * kSliceBytes + kSliceBytes};
        const int rho = ((c.j >> 1) - (wwin + 2 * hwin)) & 3;
        rotate4(o, rho);
        const uint32_t tl = c.win_addr + (uint32_t)(hwin * kWinW + wwin) * kSliceBytes;
        float dr[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const uint32_t a = tl + o[t];
            const uint4 qa = lds128(a), qb = lds128(a ^ 16u);
            dr[t] = dot<T>(gq_a, qa, 0.f) + dot<T>(gq_b, qb, 0.f);
        }
        rotate4(dr, (4 - rho) & 3);   // back to corner order TL, TR, BL, BR
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
        res[0] = w1 * dr[0] + w2 * dr[1] + w3 * dr[2] + w4 * dr[3];
        res[1] = m * (hh * (dr[1] - dr[0]) + lh * (dr[3] - dr[2]));
        res[2] = m * (hw * (dr[2] - dr[0]) + lw * (dr[3] - dr[1]));
        const int br = hwin - band_row0, bc = wwin - band_col0;
        if ((unsigned)br < (unsigned)(kBandH - 1) && (unsigned)bc < (unsigned)(kBandW - 1)) {
            const int cb = br * kBandW + bc;
            T *e0 = reinterpret_cast<T *>(c.abuf + a_elem_off(cb, c.lane));
            T *e1 = reinterpret_cast<T *>(c.abuf + a_elem_off(cb + 1, c.lane));
            T *e2 = reinterpret_cast<T *>(c.abuf + a_elem_off(cb + kBandW, c.lane));
            T *e3 = reinterpret_cast<T *>(c.abuf + a_elem_off(cb + kBandW + 1, c.lane));
            const float a0 = to_f32(*e0), a1 = to_f32(*e1), a2 = to_f32(*e2), a3 = to_f32(*e3);
            *e0 = from_f32<T>(a0 + w1 * m);
            *e1 = from_f32<T>(a1 + w2 * m);
            *e2 = from_f32<T>(a2 + w3 * m);
            *e3 = from_f32<T>(a3 + w4 * m);
        } else {
            spill_push<T>(cp, h0, w0, w1 * m, w2 * m, w3 * m, w4 * m);
        }
    } else {
        // ---- outside the window: clamped global reads, direct reductions
        constexpr int E = 8;
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
            dk[t] = dot<T>(gq_a, qa, 0.f) + dot<T>(gq_b, qb, 0.f);
        }
        const float fy_lo = ct.hh * ct.top, fy_hi = ct.lh * ct.bot;
        const float fx_lo = ct.hw * ct.lef, fx_hi = ct.lw * ct.rig;
        const float wk[4] = {fy_lo * fx_lo, fy_lo * fx_hi, fy_hi * fx_lo, fy_hi * fx_hi};
        res[0] = wk[0] * dk[0] + wk[1] * dk[1] + wk[2] * dk[2] + wk[3] * dk[3];
        res[1] = m * (fy_lo * (ct.rig * dk[1] - ct.lef * dk[0]) + fy_hi * (ct.rig * dk[3] - ct.lef * dk[2]));
        res[2] = m * (fx_lo * (ct.bot * dk[2] - ct.top * dk[0]) + fx_hi * (ct.bot * dk[3] - ct.top * dk[1]));
        const float cf[4] = {wk[0] * m, wk[1] * m, wk[2] * m, wk[3] * m};
        direct_scatter<T>(c, gq_a, gq_b, at, cf);
    }
}

// ------------------------------------------------------------------------------------------------
// Hot path.  Per-step values of one lane; coordinates are WINDOW-relative (the window origin is
// folded into the anchors), so one range test on the floats' bit patterns decides "window hit" and
// the zero fill of the TMA window stands in for the reference's range test (all four corners of a
// point that fails it lie outside the map and read zeros).
template <typename T> struct StepCtx {
    uint32_t win_addr;     // shared address of the window (+ half * 16)
    uint32_t a_lane;       // shared address of the A tile + (lane & 7) * 2
    uint32_t chunk;        // lane >> 3: this pixel's 16-byte chunk of an A row
    int jq;                // (lane & 7) >> 1: rotation phase
    int band_row0, band_col0;   // band origin in window coordinates
    int oy, ox;            // window origin in the map
    float bw, bh;          // window-relative anchors of the pixel
    float sigma;
    uint32_t s_off_lane;   // shared address of this lane's 9 (dx, dy) pairs / results
    uint32_t s_msk_lane;   // shared address of this lane's 9 masks / results
    uint4 gq_a, gq_b;      // upstream gradient of the pixel: chunk `half` / the other chunk
    uint32_t spill;        // shared address of this warp's spill list
    int lane;
};


// The NP points (tap column i, tap rows 0..NP-1) of this lane's pixel at once: independent
// instruction streams for the scheduler.  Returns false, having done nothing, if any of them
// leaves the window.
template <typename T, int NP>
__device__ __forceinline__ bool points_fast(const StepCtx<T> &c, int i) {
    float uw[NP], vw[NP], m[NP];
    bool ok = true;
    const float fi = (float)i;
    const int p0 = i * 3;
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const float2 d = unpack2(lds32(c.s_off_lane + (p0 + q) * 4), T());
        m[q] = f32_of((uint16_t)lds16(c.s_msk_lane + (p0 + q) * 2), T());
        uw[q] = c.bw + (fi + d.x) * c.sigma;
        vw[q] = c.bh + ((float)q + d.y) * c.sigma;
        // 0 <= x < limit on the float's bit pattern: negative values and NaN compare as large unsigned
        ok = ok && __float_as_uint(uw[q]) < __float_as_uint((float)(kWinW - 1)) &&
             __float_as_uint(vw[q]) < __float_as_uint((float)(kWinH - 1));
    }
    if (!ok) return false;
    // points that leave the band need a slot in the spill list: reserve them all now, while nothing
    // has been written yet; if the list is full the whole group takes the general path instead
    uint32_t pos0 = 0;
    int nslot = 0;
    {
        int nmiss = 0;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const int br = (int)floorf(vw[q]) - c.band_row0, bc = (int)floorf(uw[q]) - c.band_col0;
            nmiss += !((unsigned)br < (unsigned)(kBandH - 1) && (unsigned)bc < (unsigned)(kBandW - 1));
        }
        if (nmiss) {
            asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(pos0) : "r"(c.spill), "r"((uint32_t)nmiss) : "memory");
            if (pos0 + (uint32_t)nmiss > (uint32_t)kSpillCap) {
                // reserved slots that exist stay empty (zero coefficients are skipped)
                for (uint32_t k = pos0; k < pos0 + (uint32_t)nmiss && k < (uint32_t)kSpillCap; ++k)
                    asm volatile("st.shared.v4.f32 [%0], {%1,%1,%1,%1};" ::"r"(c.spill + 16 + k * 32 + 16), "f"(0.f) : "memory");
                return false;
            }
        }
    }

    float lh[NP], lw[NP];
    int wc[NP], wr[NP], rho[NP];
    uint4 qa[NP][4], qb[NP][4];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const float fw = floorf(uw[q]), fh = floorf(vw[q]);
        lw[q] = uw[q] - fw;
        lh[q] = vw[q] - fh;
        wc[q] = (int)fw;
        wr[q] = (int)fh;
        rho[q] = (c.jq - (wc[q] + 2 * wr[q])) & 3;
        int o[4] = {0, kSliceBytes, kWinW * kSliceBytes, kWinW * kSliceBytes + kSliceBytes};
        rotate4(o, rho[q]);
        const uint32_t tl = c.win_addr + (uint32_t)(wr[q] * kWinW + wc[q]) * kSliceBytes;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            qa[q][t] = lds128(tl + o[t]);
            qb[q][t] = lds128((tl + o[t]) ^ 16u);
        }
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        float dr[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) dr[t] = dot<T>(c.gq_a, qa[q][t], 0.f) + dot<T>(c.gq_b, qb[q][t], 0.f);
        rotate4(dr, (4 - rho[q]) & 3);   // back to corner order TL, TR, BL, BR
        const float hh = 1.f - lh[q], hw = 1.f - lw[q];
        const float gm = hh * (hw * dr[0] + lw[q] * dr[1]) + lh[q] * (hw * dr[2] + lw[q] * dr[3]);
        // the reference's range test (dcnv3_im2col_cuda.cuh:334) is implied by the window's zero fill except at
        // loc == -1 exactly (row / column -1 reads zeros: grad_mask is 0, the derivative across it is not)
        const float mg = (uw[q] == -1.f - (float)c.ox || vw[q] == -1.f - (float)c.oy) ? 0.f : m[q];
        const float gx = mg * (hh * (dr[1] - dr[0]) + lh[q] * (dr[3] - dr[2]));
        const float gy = mg * (hw * (dr[2] - dr[0]) + lw[q] * (dr[3] - dr[1]));
        sts32(c.s_off_lane + (p0 + q) * 4, pack2(c.sigma * gx, c.sigma * gy, T()));
        sts16(c.s_msk_lane + (p0 + q) * 2, bits16(gm, T()));
    }
    // A build: this pixel's column (thread-exclusive).  The four corner cells of a point are
    // distinct (read all four, then write all four); points go one after the other because two
    // points of a pixel may share a cell.  Rows cb and cb+16 have the same swizzle.
    // A point whose corner block leaves the band hands its coefficients to the warp's spill list.
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const int br = wr[q] - c.band_row0, bc = wc[q] - c.band_col0;
        const float hm = (1.f - lh[q]) * m[q], lm = lh[q] * m[q], hw = 1.f - lw[q];
        if ((unsigned)br < (unsigned)(kBandH - 1) && (unsigned)bc < (unsigned)(kBandW - 1)) {
            const uint32_t c0 = (uint32_t)(br * kBandW + bc), c1 = c0 + 1u;
            const uint32_t e0 = c.a_lane + c0 * 64u + ((c.chunk ^ ((c0 >> 1) & 3u)) << 4);
            const uint32_t e1 = c.a_lane + c1 * 64u + ((c.chunk ^ ((c1 >> 1) & 3u)) << 4);
            const float a0 = f32_of((uint16_t)lds16(e0), T()), a1 = f32_of((uint16_t)lds16(e1), T());
            const float a2 = f32_of((uint16_t)lds16(e0 + kBandW * 64), T()), a3 = f32_of((uint16_t)lds16(e1 + kBandW * 64), T());
            sts16(e0, bits16(a0 + hm * hw, T()));
            sts16(e1, bits16(a1 + hm * lw[q], T()));
            sts16(e0 + kBandW * 64, bits16(a2 + lm * hw, T()));
            sts16(e1 + kBandW * 64, bits16(a3 + lm * lw[q], T()));
        } else {
            const uint32_t e = c.spill + 16 + (pos0 + (uint32_t)nslot) * 32;   // slot reserved above
            ++nslot;
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(e), "r"(wr[q] + c.oy), "r"(wc[q] + c.ox), "r"(c.lane), "r"(0) : "memory");
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(e + 16), "f"(hm * hw), "f"(hm * lw[q]), "f"(lm * hw), "f"(lm * lw[q]) : "memory");
        }
    }
    return true;
}

// Reductions of NR finished band rows, straight from the accumulator fragments:
// lane (gid, tig) holds cells gid / gid+8 of every row and channels 2tig,2tig+1 (+8).
// p = accumulator element (row my0, column mx0 + gid, channel 2 tig); ok0 / ok1: column in the map.
template <int NR>
__device__ __forceinline__ void flush_rows(const float (&acc)[kMTiles][2][4], float *p, int my0, int H,
                                           int row_stride, int C, bool ok0, bool ok1) {
    float *p1 = p + 8 * C;
#pragma unroll
    for (int b = 0; b < NR; ++b) {
        const bool okr = (unsigned)(my0 + b) < (unsigned)H;
        if (okr && ok0) {
            red_add2(p, acc[b][0][0], acc[b][0][1]);
            red_add2(p + 8, acc[b][1][0], acc[b][1][1]);
        }
        if (okr && ok1) {
            red_add2(p1, acc[b][0][2], acc[b][0][3]);
            red_add2(p1 + 8, acc[b][1][2], acc[b][1][3]);
        }
        p += row_stride;
        p1 += row_stride;
    }
}


template <typename T>
__global__ void __launch_bounds__(kThreads, 4)
bwd_strip(const __grid_constant__ CUtensorMap tmap, const T *__restrict__ value,
          const T *__restrict__ offset, const T *__restrict__ mask, const T *__restrict__ grad_out,
          float *__restrict__ gv_acc, T *__restrict__ grad_offset, T *__restrict__ grad_mask,
          const Geom q, const Params pp) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char *win = smem;
    unsigned char *abuf = smem + kWinBytes + warp * kABytes;
    unsigned char *stage = smem + kWinBytes + kWarps * kABytes + warp * kStageBytes;
    unsigned char *spill = smem + kWinBytes + kWarps * (kABytes + kStageBytes) + warp * kSpillBytes;
    uint32_t *io_tbl = reinterpret_cast<uint32_t *>(smem + kWinBytes + kWarps * (kABytes + kStageBytes + kSpillBytes));
    const uint32_t stage_addr = smem_u32(stage);
    const uint32_t a_base = smem_u32(abuf);

    const int C = q.G * q.gc, row_stride = q.W * C;
    const int half = lane & 1;
    const int px_x = lane & 7, px_y = lane >> 3;      // this lane's pixel inside the 8x4 patch

    int t = blockIdx.x;
    if (t >= pp.total_tiles) return;
    TileAt cur = decode_tile(t, q, pp);

    // pixel index of the first pixel of a step's patch
    auto patch_pix = [&](const TileAt &a, int s) -> size_t {
        return ((size_t)a.n * q.Ho + (a.ho0 + s * kPatchH)) * q.Wo + (a.wo0 + warp * kStripW);
    };
    __shared__ __align__(16) IoCtx<T> io;
    if (tid == 0) {
        io.offset = offset; io.mask = mask; io.grad_out = grad_out; io.grad_offset = grad_offset; io.grad_mask = grad_mask;
        io.mask_end = reinterpret_cast<const unsigned char *>(mask) + pp.mask_bytes;
        io.Wo = q.Wo; io.Ho = q.Ho; io.G = q.G; io.C = C;
    }
    build_io_table(io_tbl, q.Wo, q.G * kP, tid, kThreads);
    __syncthreads();

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    {   // prologue: zero this warp's A tile and spill counter, request the first staging data
        for (int i = lane; i < kABytes / 16; i += 32) sts128_zero(a_base + i * 16);
        if (lane == 0) *reinterpret_cast<unsigned *>(spill) = 0u;
        stage_io<T>(&io, io_tbl, stage_addr, 0, 0, 0, 0, 0, patch_pix(cur, 0), cur.g, cur.wo0 + warp * kStripW, cur.ho0, 1);
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, kWinBytes);
        tma_load_4d(win, &tmap, &bar, cur.g * q.gc, cur.ox, cur.oy, cur.n);
    }

    SlowCtx<T> sc;
    sc.H = q.H; sc.W = q.W; sc.C = C; sc.row_stride = row_stride;
    sc.lane = lane; sc.j = lane & 7; sc.half = half;
    sc.win_addr = smem_u32(win) + half * 16;
    sc.abuf = abuf; sc.spill = spill; sc.s_gout = stage + kStageGout;

    float acc[kMTiles][2][4];

    for (int it = 0;; ++it) {
        const int t_next = t + gridDim.x;
        const bool has_next = t_next < pp.total_tiles;
        TileAt nxt = cur;
        if (has_next) nxt = decode_tile(t_next, q, pp);
        {
            const size_t img_base = (size_t)cur.n * q.H * row_stride + cur.g * q.gc;
            sc.oy = cur.oy; sc.ox = cur.ox; sc.img = value + img_base; sc.gv_img = gv_acc + img_base;
        }

#pragma unroll
        for (int b = 0; b < kMTiles; ++b)
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[b][n][e] = 0.f;

        mbar_wait(&bar, it & 1);   // value window of this tile has landed

#pragma unroll 1
        for (int s = 0; s < kSteps; ++s) {
            const int wb = cur.wo0 + warp * kStripW, hb = cur.ho0 + s * kPatchH;
            const int wo = wb + px_x, ho = hb + px_y;
            const bool live = wo < q.Wo && ho < q.Ho;
            const int band_row0 = pp.byw + s * kPatchH, band_col0 = pp.bxw + warp * kStripW;
            const size_t pix = patch_pix(cur, s);
            cp_async_wait_all();
            __syncwarp();

            // ------------------------------------------------------------ gather + A build
            if (live) {
                StepCtx<T> c;
                c.win_addr = smem_u32(win) + half * 16;
                c.a_lane = a_base + (lane & 7) * 2;
                c.chunk = (uint32_t)lane >> 3;
                c.jq = (lane & 7) >> 1;
                c.band_row0 = band_row0; c.band_col0 = band_col0; c.oy = cur.oy; c.ox = cur.ox; c.spill = smem_u32(spill); c.lane = lane;
                const float base_w = axis_base(wo, 3, 1, q.pw, 1, q.sigma);
                const float base_h = axis_base(ho, 3, 1, q.ph, 1, q.sigma);
                c.bw = base_w - (float)cur.ox;
                c.bh = base_h - (float)cur.oy;
                c.sigma = q.sigma;
                // the mask run of this pixel was staged from its enclosing 4-byte words: 0/1 element shift
                const unsigned sh = (unsigned)(((pix + (size_t)(px_y * q.Wo + px_x)) * q.G + cur.g) * kP) & 1u;
                c.s_off_lane = stage_addr + kStageOff + lane * (kP * 4);
                c.s_msk_lane = stage_addr + kStageMsk + lane * (kMskWords * 4) + sh * 2;
                c.gq_a = lds128(stage_addr + kStageGout + lane * kSliceBytes + half * 16);
                c.gq_b = lds128(stage_addr + kStageGout + lane * kSliceBytes + (half ^ 1) * 16);
                auto slow = [&](int p) {
                    const int i = (p * 11) >> 5, jj = p - 3 * i;
                    const float2 d = unpack2(lds32(c.s_off_lane + p * 4), T());
                    const float m = f32_of((uint16_t)lds16(c.s_msk_lane + p * 2), T());
                    float res[3];
                    slow_point<T>(&sc, band_row0, band_col0, base_h + ((float)jj + d.y) * q.sigma,
                                  base_w + ((float)i + d.x) * q.sigma, m, res);
                    sts32(c.s_off_lane + p * 4, pack2(q.sigma * res[1], q.sigma * res[2], T()));
                    sts16(c.s_msk_lane + p * 2, bits16(res[0], T()));
                };
#pragma unroll 1
                for (int i = 0; i < 3; ++i) {
                    if (!points_fast<T, 3>(c, i)) { slow(3 * i); slow(3 * i + 1); slow(3 * i + 2); }
                }
            }
            __syncwarp();

            // ---- spilled coefficients (points beyond the band): cooperative reductions
            {
                const unsigned cnt = min(*reinterpret_cast<volatile unsigned *>(spill), (unsigned)kSpillCap);
                if (cnt) {
                    const SpillEntry *se = reinterpret_cast<const SpillEntry *>(spill + 16);
                    const int corner = lane >> 3, chp = (lane & 7) * 2;
                    for (unsigned e = 0; e < cnt; ++e) {
                        const int hh = se[e].h0 + (corner >> 1), ww = se[e].w0 + (corner & 1);
                        if ((unsigned)hh < (unsigned)q.H && (unsigned)ww < (unsigned)q.W) {
                            const float cf = se[e].c[corner];
                            if (cf != 0.f) {   // (also skips reserved-but-unused slots: zero coefficients)
                                const float2 gf = unpack2(lds32(stage_addr + kStageGout + (se[e].lane & 31) * kSliceBytes + chp * 2), T());
                                red_add2(sc.gv_img + (ptrdiff_t)hh * row_stride + (ptrdiff_t)ww * C + chp, cf * gf.x, cf * gf.y);
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) *reinterpret_cast<volatile unsigned *>(spill) = 0u;
                }
            }

            // ---- B fragments (grad_out of the 32 pixels)
            uint32_t bf[2][4];   // [k-step][{n0:k0-7, n0:k8-15, n1:k0-7, n1:k8-15}]
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const int px = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                ldmatrix_x4_trans(bf[ks], stage_addr + kStageGout + px * kSliceBytes + (lane >> 4) * 16);
            }

            // ---- write this step's grad_offset / grad_mask out, request the next step's inputs
            {
                const bool in_tile = s + 1 < kSteps;
                const TileAt &na = in_tile ? cur : nxt;
                const int ns = in_tile ? s + 1 : 0;
                stage_io<T>(&io, io_tbl, stage_addr, pix, cur.g, wb, hb, 1, patch_pix(na, ns), na.g, na.wo0 + warp * kStripW,
                            na.ho0 + ns * kPatchH, in_tile || has_next);
            }

            if (s + 1 == kSteps) {
                __syncthreads();   // both warps are done with the window
                if (has_next && tid == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(&bar, kWinBytes);
                    tma_load_4d(win, &tmap, &bar, nxt.g * q.gc, nxt.ox, nxt.oy, nxt.n);
                }
            }

            // ------------------------------------------------------------ acc += A x grad_out
            {
                const int r_in = (lane & 7) + ((lane >> 3) & 1) * 8;
                const uint32_t kc_in = (uint32_t)lane >> 4;
                const uint32_t sw = ((uint32_t)r_in >> 1) & 3u;            // (row >> 1) & 3, row = 16 b + r_in
                const uint32_t ra0 = a_base + (uint32_t)r_in * 64u + ((kc_in ^ sw) << 4);
                const uint32_t ra1 = a_base + (uint32_t)r_in * 64u + (((2u + kc_in) ^ sw) << 4);
                const uint32_t za = a_base + lane * 16;
#pragma unroll
                for (int b = 0; b < kMTiles; ++b) {
                    uint32_t a0[4], a1[4];
                    ldmatrix_x4(a0, ra0 + b * 1024);
                    ldmatrix_x4(a1, ra1 + b * 1024);
                    mma16816(acc[b][0], a0, bf[0][0], bf[0][1], T());
                    mma16816(acc[b][1], a0, bf[0][2], bf[0][3], T());
                    mma16816(acc[b][0], a1, bf[1][0], bf[1][1], T());
                    mma16816(acc[b][1], a1, bf[1][2], bf[1][3], T());
                    // the m-tile is in registers: clear it for the next step
                    sts128_zero(za + b * 1024);
                    sts128_zero(za + b * 1024 + 512);
                }
            }

            // ------------------------------------------------------------ finished rows leave
            {
                const int gid = lane >> 2, tig = lane & 3;
                const int my0 = cur.oy + band_row0, mx = cur.ox + band_col0 + gid;
                float *p0 = sc.gv_img + (ptrdiff_t)my0 * row_stride + (ptrdiff_t)mx * C + 2 * tig;
                const bool ok0 = (unsigned)mx < (unsigned)q.W, ok1 = (unsigned)(mx + 8) < (unsigned)q.W;
                if (s + 1 < kSteps) {
                    flush_rows<kPatchH>(acc, p0, my0, q.H, row_stride, C, ok0, ok1);
#pragma unroll
                    for (int b = 0; b < kMTiles; ++b)
#pragma unroll
                        for (int n = 0; n < 2; ++n)
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                if (b + kPatchH < kMTiles) acc[b][n][e] = acc[b + kPatchH < kMTiles ? b + kPatchH : 0][n][e];
                                else acc[b][n][e] = 0.f;
                            }
                } else {
                    flush_rows<kMTiles>(acc, p0, my0, q.H, row_stride, C, ok0, ok1);
                }
            }
        }
        if (!has_next) break;
        cur = nxt;
        t = t_next;
    }
}

template <typename T>
static bool launch_typed(const void *value, const void *offset, const void *mask, const void *grad_out,
                         float *gv_acc, void *grad_offset, void *grad_mask, const Geom &q, int dtype,
                         cudaStream_t stream, cudaError_t *err) {
    if (q.gc != kCh || q.kh != 3 || q.kw != 3 || q.sh != 1 || q.sw != 1 || q.dh != 1 || q.dw != 1) return false;
    if (!(q.sigma >= 0.5f && q.sigma <= 1.25f)) return false;   // band = taps +- 3 px: keep >= 2.4 sigma of slack
    if (((uintptr_t)value | (uintptr_t)grad_out | (uintptr_t)gv_acc) % 16) return false;
    if (((uintptr_t)offset | (uintptr_t)grad_offset | (uintptr_t)mask) % 4) return false;
    if ((uintptr_t)grad_mask % 2) return false;
    const int C = q.G * q.gc;
    Params pp;
    // nominal taps of a pixel x along an axis: x + a + i*sigma, i = 0..2, a = (1 - pad) - sigma
    const float a_w = (float)(1 - q.pw) - q.sigma, a_h = (float)(1 - q.ph) - q.sigma;
    const float span_w = (kTileW - 1) + 2 * q.sigma, span_h = (kTileH - 1) + 2 * q.sigma;
    pp.ox_rel = (int)std::floor(a_w + 0.5f * span_w - 0.5f * (kWinW - 2));
    pp.oy_rel = (int)std::floor(a_h + 0.5f * span_h - 0.5f * (kWinH - 2));
    // band of strip 0 / step 0: centred on the patch's taps
    const int bx_rel = (int)std::floor(a_w + q.sigma + 0.5f * (kStripW - 1) + 0.5f - 0.5f * kBandW);
    const int by_rel = (int)std::floor(a_h + q.sigma + 0.5f * (kPatchH - 1) + 0.5f - 0.5f * kBandH);
    pp.bxw = bx_rel - pp.ox_rel;
    pp.byw = by_rel - pp.oy_rel;
    if (pp.bxw < 0 || pp.bxw + (kWarps - 1) * kStripW + kBandW > kWinW) return false;
    if (pp.byw < 0 || pp.byw + (kSteps - 1) * kPatchH + kBandH > kWinH) return false;
    pp.tiles_x = (q.Wo + kTileW - 1) / kTileW;
    const int tiles_y = (q.Ho + kTileH - 1) / kTileH;
    const long long total = (long long)pp.tiles_x * tiles_y * q.G * q.N;
    if (total >= (1LL << 31)) return false;
    pp.tiles_xy = pp.tiles_x * tiles_y;
    pp.total_tiles = (int)total;
    pp.mask_bytes = (unsigned long long)q.N * q.Ho * q.Wo * q.G * kP * 2ull;
    if ((long long)(3 * q.Wo + 8) * q.G * kP + kP >= (1LL << 24)) return false;   // staging index table packing
    CUtensorMap tmap;
    if (!make_nhwc_tensor_map(&tmap, value, dtype, q.N, q.H, q.W, C, kCh, kWinW, kWinH)) return false;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int ctas = (int)std::min<long long>(total, 4LL * num_sms);
    cudaFuncSetAttribute(bwd_strip<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    bwd_strip<T><<<ctas, kThreads, kSmemBytes, stream>>>(
        tmap, static_cast<const T *>(value), static_cast<const T *>(offset), static_cast<const T *>(mask),
        static_cast<const T *>(grad_out), gv_acc, static_cast<T *>(grad_offset), static_cast<T *>(grad_mask), q, pp);
    *err = cudaGetLastError();
    return true;
}


// ================================================================================================
// Value-only form: second half of the SPLIT backward (grad_offset / grad_mask come from
// dcnv3_backward_dots.cu).  grad_value needs no value window and no channel sums: a step is
// coordinates -> A build -> acc += A x grad_out -> finished rows leave as reductions.  Without the
// 21 KB window a two-warp CTA needs 35 KB of shared memory, so 12 warps fit an SM instead of 8, and
// the per-point work drops to the coordinate arithmetic and four 16-bit read-modify-writes.
constexpr int kVSmemBytes = kWarps * (kABytes + kStageBytes + kSpillBytes) + kIoTblBytes;

struct VParams {
    int bx_rel, by_rel;      // band origin relative to the first pixel of a warp's patch (map coordinates)
    int tiles_x, tiles_xy, total_tiles;
    int steps;               // 4-row steps per tile: without a window the strip can be as tall as the map
    unsigned long long mask_bytes;
};

// One sampling point, general path (band overflow of the spill list, rare): coefficients into the
// A tile or straight to the global accumulator.
template <typename T>
__device__ __noinline__ void slow_vpoint(const SlowCtx<T> *cp, int band_y0, int band_x0, float loc_h, float loc_w, float m) {
    const SlowCtx<T> &c = *cp;
    const bool inside = loc_h > -1.f && loc_w > -1.f && loc_h < (float)c.H && loc_w < (float)c.W;
    if (!inside) return;
    const float fh = floorf(loc_h), fw = floorf(loc_w);
    const float lh = loc_h - fh, lw = loc_w - fw, hh = 1.f - lh, hw = 1.f - lw;
    const int h0 = (int)fh, w0 = (int)fw;
    const int br = h0 - band_y0, bc = w0 - band_x0;
    const float cf[4] = {hh * hw * m, hh * lw * m, lh * hw * m, lh * lw * m};
    if ((unsigned)br < (unsigned)(kBandH - 1) && (unsigned)bc < (unsigned)(kBandW - 1)) {
        const int cb = br * kBandW + bc;
        T *e0 = reinterpret_cast<T *>(c.abuf + a_elem_off(cb, c.lane));
        T *e1 = reinterpret_cast<T *>(c.abuf + a_elem_off(cb + 1, c.lane));
        T *e2 = reinterpret_cast<T *>(c.abuf + a_elem_off(cb + kBandW, c.lane));
        T *e3 = reinterpret_cast<T *>(c.abuf + a_elem_off(cb + kBandW + 1, c.lane));
        const float a0 = to_f32(*e0), a1 = to_f32(*e1), a2 = to_f32(*e2), a3 = to_f32(*e3);
        *e0 = from_f32<T>(a0 + cf[0]);
        *e1 = from_f32<T>(a1 + cf[1]);
        *e2 = from_f32<T>(a2 + cf[2]);
        *e3 = from_f32<T>(a3 + cf[3]);
    } else {
        const uint4 gq_a = *reinterpret_cast<const uint4 *>(c.s_gout + c.lane * kSliceBytes + c.half * 16);
        const uint4 gq_b = *reinterpret_cast<const uint4 *>(c.s_gout + c.lane * kSliceBytes + (c.half ^ 1) * 16);
        const bool top = h0 >= 0, bot = h0 + 1 < c.H, lef = w0 >= 0, rig = w0 + 1 < c.W;
        const int at[4] = {h0 * c.row_stride + w0 * c.C, h0 * c.row_stride + (w0 + 1) * c.C,
                           (h0 + 1) * c.row_stride + w0 * c.C, (h0 + 1) * c.row_stride + (w0 + 1) * c.C};
        const float cg[4] = {top && lef ? cf[0] : 0.f, top && rig ? cf[1] : 0.f, bot && lef ? cf[2] : 0.f, bot && rig ? cf[3] : 0.f};
        direct_scatter<T>(c, gq_a, gq_b, at, cg);
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 6)
bwd_vstrip(const T *__restrict__ offset, const T *__restrict__ mask, const T *__restrict__ grad_out,
           float *__restrict__ gv_acc, const Geom q, const VParams pp) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char *abuf = smem + warp * kABytes;
    unsigned char *stage = smem + kWarps * kABytes + warp * kStageBytes;
    unsigned char *spill = smem + kWarps * (kABytes + kStageBytes) + warp * kSpillBytes;
    uint32_t *io_tbl = reinterpret_cast<uint32_t *>(smem + kWarps * (kABytes + kStageBytes + kSpillBytes));
    const uint32_t stage_addr = smem_u32(stage), a_base = smem_u32(abuf), spill_addr = smem_u32(spill);

    const int C = q.G * q.gc, row_stride = q.W * C;
    const int half = lane & 1;
    const int px_x = lane & 7, px_y = lane >> 3;

    int t = blockIdx.x;
    if (t >= pp.total_tiles) return;
    auto decode = [&](int tt, int &n, int &g, int &wo0, int &ho0) {
        const int txy = tt % pp.tiles_xy, r = tt / pp.tiles_xy;
        g = r % q.G; n = r / q.G;
        wo0 = (txy % pp.tiles_x) * kTileW; ho0 = (txy / pp.tiles_x) * (pp.steps * kPatchH);
    };
    int n, g, wo0, ho0;
    decode(t, n, g, wo0, ho0);
    auto patch_pix = [&](int nn, int w0, int h0, int s) -> size_t {
        return ((size_t)nn * q.Ho + (h0 + s * kPatchH)) * q.Wo + (w0 + warp * kStripW);
    };
    __shared__ __align__(16) IoCtx<T> io;
    if (tid == 0) {
        io.offset = offset; io.mask = mask; io.grad_out = grad_out; io.grad_offset = nullptr; io.grad_mask = nullptr;
        io.mask_end = reinterpret_cast<const unsigned char *>(mask) + pp.mask_bytes;
        io.Wo = q.Wo; io.Ho = q.Ho; io.G = q.G; io.C = C;
    }
    build_io_table(io_tbl, q.Wo, q.G * kP, tid, kThreads);
    __syncthreads();
    for (int i = lane; i < kABytes / 16; i += 32) sts128_zero(a_base + i * 16);
    if (lane == 0) *reinterpret_cast<unsigned *>(spill) = 0u;
    stage_io<T>(&io, io_tbl, stage_addr, 0, 0, 0, 0, 0, patch_pix(n, wo0, ho0, 0), g, wo0 + warp * kStripW, ho0, 1);

    SlowCtx<T> sc;
    sc.H = q.H; sc.W = q.W; sc.C = C; sc.row_stride = row_stride;
    sc.lane = lane; sc.j = lane & 7; sc.half = half; sc.win_addr = 0;
    sc.abuf = abuf; sc.spill = spill; sc.s_gout = stage + kStageGout; sc.img = nullptr; sc.oy = 0; sc.ox = 0;

    float acc[kMTiles][2][4];
    for (;;) {
        const int t_next = t + gridDim.x;
        const bool has_next = t_next < pp.total_tiles;
        int n2 = n, g2 = g, wo2 = wo0, ho2 = ho0;
        if (has_next) decode(t_next, n2, g2, wo2, ho2);
        float *gv_img = gv_acc + (size_t)n * q.H * row_stride + g * q.gc;
        sc.gv_img = gv_img;
#pragma unroll
        for (int b = 0; b < kMTiles; ++b)
#pragma unroll
            for (int nn = 0; nn < 2; ++nn)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[b][nn][e] = 0.f;

#pragma unroll 1
        for (int s = 0; s < pp.steps; ++s) {
            const int wb = wo0 + warp * kStripW, hb = ho0 + s * kPatchH;
            const int wo = wb + px_x, ho = hb + px_y;
            const bool live = wo < q.Wo && ho < q.Ho;
            const int band_x0 = wb + pp.bx_rel, band_y0 = hb + pp.by_rel;   // band origin, map coordinates
            const size_t pix = patch_pix(n, wo0, ho0, s);
            cp_async_wait_all();
            __syncwarp();

            // ------------------------------------------------------------ A build (no gather here)
            if (live) {
                const float base_w = axis_base(wo, 3, 1, q.pw, 1, q.sigma), base_h = axis_base(ho, 3, 1, q.ph, 1, q.sigma);
                const float bw = base_w - (float)band_x0, bh = base_h - (float)band_y0;   // band-relative anchors
                const unsigned sh = (unsigned)(((pix + (size_t)(px_y * q.Wo + px_x)) * q.G + g) * kP) & 1u;
                const uint32_t s_off_lane = stage_addr + kStageOff + lane * (kP * 4);
                const uint32_t s_msk_lane = stage_addr + kStageMsk + lane * (kMskWords * 4) + sh * 2;
                const uint32_t a_lane = a_base + (lane & 7) * 2, chunk = (uint32_t)lane >> 3;
                unsigned ovf = 0;
                // one point at a time (no per-point state is kept: the accumulators need the registers)
#pragma unroll
                for (int p = 0; p < kP; ++p) {
                    const float2 d = unpack2(lds32(s_off_lane + p * 4), T());
                    const float m = f32_of((uint16_t)lds16(s_msk_lane + p * 2), T());
                    const float ub = bw + ((float)(p / 3) + d.x) * q.sigma;
                    const float vb = bh + ((float)(p % 3) + d.y) * q.sigma;
                    const float fw = floorf(ub), fh = floorf(vb);
                    const float lw = ub - fw, lh = vb - fh;
                    const float hm = (1.f - lh) * m, lm = lh * m, hw = 1.f - lw;
                    // 0 <= x < limit on the float's bit pattern: negative values and NaN compare as large unsigned
                    if (__float_as_uint(ub) < __float_as_uint((float)(kBandW - 1)) &&
                        __float_as_uint(vb) < __float_as_uint((float)(kBandH - 1))) {
                        const uint32_t c0 = (uint32_t)((int)fh * kBandW + (int)fw), c1 = c0 + 1u;
                        const uint32_t e0 = a_lane + c0 * 64u + ((chunk ^ ((c0 >> 1) & 3u)) << 4);
                        const uint32_t e1 = a_lane + c1 * 64u + ((chunk ^ ((c1 >> 1) & 3u)) << 4);
                        const float a0 = f32_of((uint16_t)lds16(e0), T()), a1 = f32_of((uint16_t)lds16(e1), T());
                        const float a2 = f32_of((uint16_t)lds16(e0 + kBandW * 64), T()), a3 = f32_of((uint16_t)lds16(e1 + kBandW * 64), T());
                        sts16(e0, bits16(a0 + hm * hw, T()));
                        sts16(e1, bits16(a1 + hm * lw, T()));
                        sts16(e0 + kBandW * 64, bits16(a2 + lm * hw, T()));
                        sts16(e1 + kBandW * 64, bits16(a3 + lm * lw, T()));
                    } else {
                        // beyond the band: the reference's range test decides whether the point counts at all
                        const float lw_abs = ub + (float)band_x0, lh_abs = vb + (float)band_y0;
                        if (lh_abs > -1.f && lw_abs > -1.f && lh_abs < (float)q.H && lw_abs < (float)q.W) {
                            uint32_t pos;
                            asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(pos) : "r"(spill_addr), "r"(1u) : "memory");
                            if (pos < (uint32_t)kSpillCap) {
                                const uint32_t e = spill_addr + 16 + pos * 32;
                                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(e), "r"((int)fh + band_y0), "r"((int)fw + band_x0), "r"(lane), "r"(0) : "memory");
                                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(e + 16), "f"(hm * hw), "f"(hm * lw), "f"(lm * hw), "f"(lm * lw) : "memory");
                            } else {
                                ovf |= 1u << p;   // list full: redone through the general path below
                            }
                        }
                    }
                }
                if (ovf) {   // rare: re-derive those points from the (untouched) staging data
#pragma unroll 1
                    for (int p = 0; p < kP; ++p)
                        if ((ovf >> p) & 1u) {
                            const float2 d = unpack2(lds32(s_off_lane + p * 4), T());
                            const float m = f32_of((uint16_t)lds16(s_msk_lane + p * 2), T());
                            const int i = (p * 11) >> 5, jj = p - 3 * i;
                            slow_vpoint<T>(&sc, band_y0, band_x0, base_h + ((float)jj + d.y) * q.sigma,
                                           base_w + ((float)i + d.x) * q.sigma, m);
                        }
                }
            }
            __syncwarp();

            // ---- spilled coefficients (points beyond the band): cooperative reductions
            {
                const unsigned cnt = min(*reinterpret_cast<volatile unsigned *>(spill), (unsigned)kSpillCap);
                if (cnt) {
                    const SpillEntry *se = reinterpret_cast<const SpillEntry *>(spill + 16);
                    const int corner = lane >> 3, chp = (lane & 7) * 2;
                    for (unsigned e = 0; e < cnt; ++e) {
                        const int hh = se[e].h0 + (corner >> 1), ww = se[e].w0 + (corner & 1);
                        if ((unsigned)hh < (unsigned)q.H && (unsigned)ww < (unsigned)q.W) {
                            const float cf = se[e].c[corner];
                            if (cf != 0.f) {
                                const float2 gf = unpack2(lds32(stage_addr + kStageGout + (se[e].lane & 31) * kSliceBytes + chp * 2), T());
                                red_add2(gv_img + (ptrdiff_t)hh * row_stride + (ptrdiff_t)ww * C + chp, cf * gf.x, cf * gf.y);
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) *reinterpret_cast<volatile unsigned *>(spill) = 0u;
                }
            }

            // ---- B fragments (grad_out of the 32 pixels), then request the next step's inputs
            uint32_t bf[2][4];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const int pxl = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                ldmatrix_x4_trans(bf[ks], stage_addr + kStageGout + pxl * kSliceBytes + (lane >> 4) * 16);
            }
            {
                const bool in_tile = s + 1 < pp.steps;
                const int nn = in_tile ? n : n2, gg = in_tile ? g : g2, w0n = in_tile ? wo0 : wo2, h0n = in_tile ? ho0 : ho2;
                const int ns = in_tile ? s + 1 : 0;
                stage_io<T>(&io, io_tbl, stage_addr, 0, 0, 0, 0, 0, patch_pix(nn, w0n, h0n, ns), gg, w0n + warp * kStripW,
                            h0n + ns * kPatchH, in_tile || has_next);
            }

            // ------------------------------------------------------------ acc += A x grad_out
            {
                const int r_in = (lane & 7) + ((lane >> 3) & 1) * 8;
                const uint32_t kc_in = (uint32_t)lane >> 4;
                const uint32_t sw = ((uint32_t)r_in >> 1) & 3u;
                const uint32_t ra0 = a_base + (uint32_t)r_in * 64u + ((kc_in ^ sw) << 4);
                const uint32_t ra1 = a_base + (uint32_t)r_in * 64u + (((2u + kc_in) ^ sw) << 4);
                const uint32_t za = a_base + lane * 16;
#pragma unroll
                for (int b = 0; b < kMTiles; ++b) {
                    uint32_t a0[4], a1[4];
                    ldmatrix_x4(a0, ra0 + b * 1024);
                    ldmatrix_x4(a1, ra1 + b * 1024);
                    mma16816(acc[b][0], a0, bf[0][0], bf[0][1], T());
                    mma16816(acc[b][1], a0, bf[0][2], bf[0][3], T());
                    mma16816(acc[b][0], a1, bf[1][0], bf[1][1], T());
                    mma16816(acc[b][1], a1, bf[1][2], bf[1][3], T());
                    sts128_zero(za + b * 1024);
                    sts128_zero(za + b * 1024 + 512);
                }
            }

            // ------------------------------------------------------------ finished rows leave
            {
                const int gid = lane >> 2, tig = lane & 3;
                const int mx = band_x0 + gid;
                float *p0 = gv_img + (ptrdiff_t)band_y0 * row_stride + (ptrdiff_t)mx * C + 2 * tig;
                const bool ok0 = (unsigned)mx < (unsigned)q.W, ok1 = (unsigned)(mx + 8) < (unsigned)q.W;
                if (s + 1 < pp.steps) {
                    flush_rows<kPatchH>(acc, p0, band_y0, q.H, row_stride, C, ok0, ok1);
#pragma unroll
                    for (int b = 0; b < kMTiles; ++b)
#pragma unroll
                        for (int nn = 0; nn < 2; ++nn)
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                if (b + kPatchH < kMTiles) acc[b][nn][e] = acc[b + kPatchH < kMTiles ? b + kPatchH : 0][nn][e];
                                else acc[b][nn][e] = 0.f;
                            }
                } else {
                    flush_rows<kMTiles>(acc, p0, band_y0, q.H, row_stride, C, ok0, ok1);
                }
            }
        }
        if (!has_next) break;
        n = n2; g = g2; wo0 = wo2; ho0 = ho2;
        t = t_next;
    }
}

template <typename T>
static bool launch_value_typed(const void *offset, const void *mask, const void *grad_out, float *gv_acc, const Geom &q,
                               cudaStream_t stream, cudaError_t *err) {
    if (q.gc != kCh || q.kh != 3 || q.kw != 3 || q.sh != 1 || q.sw != 1 || q.dh != 1 || q.dw != 1) return false;
    if (!(q.sigma >= 0.5f && q.sigma <= 1.25f)) return false;
    if (((uintptr_t)grad_out | (uintptr_t)gv_acc) % 16 || ((uintptr_t)offset | (uintptr_t)mask) % 4) return false;
    VParams pp;
    const float a_w = (float)(1 - q.pw) - q.sigma, a_h = (float)(1 - q.ph) - q.sigma;
    pp.bx_rel = (int)std::floor(a_w + q.sigma + 0.5f * (kStripW - 1) + 0.5f - 0.5f * kBandW);
    pp.by_rel = (int)std::floor(a_h + q.sigma + 0.5f * (kPatchH - 1) + 0.5f - 0.5f * kBandH);
    pp.tiles_x = (q.Wo + kTileW - 1) / kTileW;
    // tall tiles: the rows a strip flushes are (4 steps + 8) per (4 steps) rows -- 1.5x at 4 steps, 1.2x at 10
    pp.steps = std::max(1, std::min((q.Ho + kPatchH - 1) / kPatchH, 10));
    if (const char *e = std::getenv("DCNV3_VSTEPS")) pp.steps = std::max(1, std::min(64, atoi(e)));
    const int tile_h = pp.steps * kPatchH;
    const int tiles_y = (q.Ho + tile_h - 1) / tile_h;
    const long long total = (long long)pp.tiles_x * tiles_y * q.G * q.N;
    if (total >= (1LL << 31)) return false;
    pp.tiles_xy = pp.tiles_x * tiles_y;
    pp.total_tiles = (int)total;
    pp.mask_bytes = (unsigned long long)q.N * q.Ho * q.Wo * q.G * kP * 2ull;
    if ((long long)(3 * q.Wo + 8) * q.G * kP + kP >= (1LL << 24)) return false;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int ctas = (int)std::min<long long>(total, 6LL * num_sms);
    cudaFuncSetAttribute(bwd_vstrip<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVSmemBytes);
    bwd_vstrip<T><<<ctas, kThreads, kVSmemBytes, stream>>>(static_cast<const T *>(offset), static_cast<const T *>(mask),
                                                          static_cast<const T *>(grad_out), gv_acc, q, pp);
    *err = cudaGetLastError();
    return true;
}

}  // namespace strip

// gv_acc: zero-initialised fp32 accumulator with the shape of value.  Returns false if the shape
// is not eligible (caller tries the tiled tensor-core kernel next).
bool try_launch_backward_strip(const void *value, const void *offset, const void *mask,
                               const void *grad_out, float *gv_acc, void *grad_offset, void *grad_mask,
                               const Geom &q, int dtype, cudaStream_t stream, cudaError_t *err) {
    const char *e = std::getenv("DCNV3_BWD");   // development knob: any value selects an older kernel
    if (e && e[0] && !(e[0] == 's' && e[1] == 't')) return false;
    if ((long long)q.N * q.Ho * q.Wo == 0) return false;
    if (dtype == 1) return strip::launch_typed<__half>(value, offset, mask, grad_out, gv_acc, grad_offset, grad_mask, q, dtype, stream, err);
    if (dtype == 2) return strip::launch_typed<__nv_bfloat16>(value, offset, mask, grad_out, gv_acc, grad_offset, grad_mask, q, dtype, stream, err);
    return false;
}

// grad_value only (accumulated into the zeroed fp32 plane gv_acc); the split backward's second half.
bool try_launch_backward_vstrip(const void *offset, const void *mask, const void *grad_out, float *gv_acc,
                                const Geom &q, int dtype, cudaStream_t stream, cudaError_t *err) {
    if ((long long)q.N * q.Ho * q.Wo == 0) return false;
    if (dtype == 1) return strip::launch_value_typed<__half>(offset, mask, grad_out, gv_acc, q, stream, err);
    if (dtype == 2) return strip::launch_value_typed<__nv_bfloat16>(offset, mask, grad_out, gv_acc, q, stream, err);
    return false;
}

}  // namespace dcnv3
