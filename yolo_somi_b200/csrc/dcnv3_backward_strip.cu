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

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace dcnv3 {
namespace strip {

constexpr int kStripW = 8, kPatchH = 4, kSteps = 4;
constexpr int kWarps = 2;
constexpr int kTileW = kStripW * kWarps, kTileH = kPatchH * kSteps;   // 16 x 16 output pixels
constexpr int kThreads = kWarps * 32;
constexpr int kWinW = 26, kWinH = 26;            // value window (2 mod 4: conflict-free corners)
constexpr int kBandW = 16, kBandH = 14;          // cells one step of one warp can reach
constexpr int kMTiles = kBandH;                  // one m-tile (16 cells) per band row
constexpr int kCh = 16, kSliceBytes = 32;
constexpr int kP = 9;
constexpr int kSpillCap = 8;
static_assert(kWinW % 4 == 2, "window width must be 2 mod 4");

constexpr int kWinBytes = kWinW * kWinH * kSliceBytes;          // 21632
constexpr int kABytes = kBandH * kBandW * 64;                   // 14336 per warp
constexpr int kMskWords = 7;                                    // per pixel: 5 words used, odd stride
constexpr int kStageOff = 0;                                    // [32][9] u32
constexpr int kStageMsk = kStageOff + 32 * kP * 4;              // [32][7] u32
constexpr int kStageGout = kStageMsk + 32 * kMskWords * 4;      // [32][32 B]
constexpr int kStageBytes = kStageGout + 32 * kSliceBytes;      // 3072 per warp
constexpr int kSpillBytes = 16 + kSpillCap * 32;                // 272 per warp
constexpr int kSmemBytes = kWinBytes + kWarps * (kABytes + kStageBytes + kSpillBytes);

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

__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128_zero(uint32_t a) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(a), "r"(0u) : "memory");
}
template <typename V> __device__ __forceinline__ void rotate4(V (&x)[4], int r) {  // out[t] = in[(t+r)&3]
    if (r & 1) { const V t = x[0]; x[0] = x[1]; x[1] = x[2]; x[2] = x[3]; x[3] = t; }
    if (r & 2) { V t = x[0]; x[0] = x[2]; x[2] = t; t = x[1]; x[1] = x[3]; x[3] = t; }
}
__device__ __forceinline__ void red_add2(float *p, float a, float b) { atomicAdd(reinterpret_cast<float2 *>(p), make_float2(a, b)); }
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

// cp.async with zero fill: copies `src_bytes` (<= size) and zero-fills the rest
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// byte offset of A[cell][pixel] inside a warp's A tile: 64-byte rows, 16-byte chunks swizzled
__device__ __forceinline__ uint32_t a_elem_off(int cell, int pixel) {
    return (uint32_t)cell * 64u + ((((uint32_t)pixel >> 3) ^ (((uint32_t)cell >> 1) & 3u)) << 4) + (((uint32_t)pixel & 7u) << 1);
}

template <typename T> struct Ctx {
    int H, W, C, row_stride;
    int oy, ox;                      // window origin in the map
    int band_row0, band_col0;        // band origin of this warp / step in window coordinates
    int j, half, lane;
    uint32_t win_addr;               // shared address of the window (+ half * 16)
    unsigned char *abuf;             // this warp's A tile
    unsigned char *spill;            // this warp's spill list
    const T *img;                    // value + (n, 0, 0, ch0)
    float *gv_img;                   // fp32 accumulator + (n, 0, 0, ch0)
    uint4 gq_a, gq_b;                // upstream gradient of the pixel: chunk `half` / the other chunk
};

// coefficient x grad_out of one lane's pixel straight to the global accumulator (rare paths)
template <typename T>
__device__ __forceinline__ void direct_scatter(const Ctx<T> &c, const int (&at)[4], const float (&cf)[4]) {
    constexpr int E = 8;
    float ga[E], gb[E];
    unpack<T>(c.gq_a, ga);
    unpack<T>(c.gq_b, gb);
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

// One sampling point: channel sums (grad_mask, grad_offset / sigma) and the point's four
// coefficients into the A tile / the spill list / the global accumulator.
template <typename T>
__device__ __forceinline__ void process_point(const Ctx<T> &c, float loc_h, float loc_w, float m,
                                              float &gm, float &gx, float &gy) {
    gm = gx = gy = 0.f;
    // range test of the reference (dcnv3_im2col_cuda.cuh:262-263); also rejects NaN
    const bool inside = loc_h > -1.f && loc_w > -1.f && loc_h < (float)c.H && loc_w < (float)c.W;
    if (!inside) return;
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
            dr[t] = dot<T>(c.gq_a, qa, 0.f) + dot<T>(c.gq_b, qb, 0.f);
        }
        rotate4(dr, (4 - rho) & 3);   // back to corner order TL, TR, BL, BR
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
        gm = w1 * dr[0] + w2 * dr[1] + w3 * dr[2] + w4 * dr[3];
        gx = m * (hh * (dr[1] - dr[0]) + lh * (dr[3] - dr[2]));
        gy = m * (hw * (dr[2] - dr[0]) + lw * (dr[3] - dr[1]));
        const int br = hwin - c.band_row0, bc = wwin - c.band_col0;
        if ((unsigned)br < (unsigned)(kBandH - 1) && (unsigned)bc < (unsigned)(kBandW - 1)) {
            // this pixel's column of A (thread-exclusive; the four corner cells are distinct:
            // read all four, then write all four)
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
            // beyond the band: hand the coefficients to the warp (spill list), or reduce directly
            unsigned *cnt = reinterpret_cast<unsigned *>(c.spill);
            const unsigned pos = atomicAdd(cnt, 1u);
            const float cf[4] = {w1 * m, w2 * m, w3 * m, w4 * m};
            if (pos < (unsigned)kSpillCap) {
                SpillEntry *e = reinterpret_cast<SpillEntry *>(c.spill + 16) + pos;
                e->h0 = h0; e->w0 = w0; e->lane = c.lane; e->pad = 0;
                e->c[0] = cf[0]; e->c[1] = cf[1]; e->c[2] = cf[2]; e->c[3] = cf[3];
            } else {
                const bool top = h0 >= 0, bot = h0 + 1 < c.H, lef = w0 >= 0, rig = w0 + 1 < c.W;
                const int at[4] = {h0 * c.row_stride + w0 * c.C, h0 * c.row_stride + (w0 + 1) * c.C,
                                   (h0 + 1) * c.row_stride + w0 * c.C, (h0 + 1) * c.row_stride + (w0 + 1) * c.C};
                const float cg[4] = {top && lef ? cf[0] : 0.f, top && rig ? cf[1] : 0.f,
                                     bot && lef ? cf[2] : 0.f, bot && rig ? cf[3] : 0.f};
                direct_scatter<T>(c, at, cg);
            }
        }
    } else {
        // ---- fallback: clamped global reads, direct reductions
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
            dk[t] = dot<T>(c.gq_a, qa, 0.f) + dot<T>(c.gq_b, qb, 0.f);
        }
        const float fy_lo = ct.hh * ct.top, fy_hi = ct.lh * ct.bot;
        const float fx_lo = ct.hw * ct.lef, fx_hi = ct.lw * ct.rig;
        const float wk[4] = {fy_lo * fx_lo, fy_lo * fx_hi, fy_hi * fx_lo, fy_hi * fx_hi};
        gm = wk[0] * dk[0] + wk[1] * dk[1] + wk[2] * dk[2] + wk[3] * dk[3];
        gx = m * (fy_lo * (ct.rig * dk[1] - ct.lef * dk[0]) + fy_hi * (ct.rig * dk[3] - ct.lef * dk[2]));
        gy = m * (fx_lo * (ct.bot * dk[2] - ct.top * dk[0]) + fx_hi * (ct.bot * dk[3] - ct.top * dk[1]));
        const float cf[4] = {wk[0] * m, wk[1] * m, wk[2] * m, wk[3] * m};
        direct_scatter<T>(c, at, cf);
    }
}

// Reductions of NR finished band rows, straight from the accumulator fragments:
// lane (gid, tig) holds cells gid / gid+8 of every row and channels 2tig,2tig+1 (+8).
template <int NR>
__device__ __forceinline__ void flush_rows(const float (&acc)[kMTiles][2][4], float *gv_img, int my0, int mx0,
                                           int H, int W, int row_stride, int C, int lane) {
    const int gid = lane >> 2, tig = lane & 3;
#pragma unroll
    for (int b = 0; b < NR; ++b) {
        const int my = my0 + b;
        if ((unsigned)my < (unsigned)H) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int mx = mx0 + gid + 8 * hf;
                if ((unsigned)mx < (unsigned)W) {
                    float *p = gv_img + (ptrdiff_t)my * row_stride + (ptrdiff_t)mx * C + 2 * tig;
                    red_add2(p, acc[b][0][2 * hf], acc[b][0][2 * hf + 1]);
                    red_add2(p + 8, acc[b][1][2 * hf], acc[b][1][2 * hf + 1]);
                }
            }
        }
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
    uint32_t *s_off = reinterpret_cast<uint32_t *>(stage + kStageOff);
    uint16_t *s_msk = reinterpret_cast<uint16_t *>(stage + kStageMsk);
    unsigned char *s_gout = stage + kStageGout;
    const uint32_t stage_addr = smem_u32(stage);
    const uint32_t a_base = smem_u32(abuf);

    const int C = q.G * q.gc, row_stride = q.W * C;
    const int j = lane & 7, half = j & 1;
    const int px_x = lane & 7, px_y = lane >> 3;      // this lane's pixel inside the 8x4 patch

    int t = blockIdx.x;
    if (t >= pp.total_tiles) return;
    TileAt cur = decode_tile(t, q, pp);

    // cp.async staging of one step's offsets / masks / grad_out for this warp (no registers held)
    auto prefetch = [&](const TileAt &a, int s) {
        const size_t img_pix = (size_t)a.n * q.Ho * q.Wo;
        const int wb = a.wo0 + warp * kStripW, hb = a.ho0 + s * kPatchH;
        const uint32_t *obase = reinterpret_cast<const uint32_t *>(offset);
        const unsigned char *mbase = reinterpret_cast<const unsigned char *>(mask);
#pragma unroll
        for (int it = 0; it < kP; ++it) {
            const int idx = lane + it * 32, px = idx / kP, p = idx - px * kP;
            const int w = wb + (px & 7), h = hb + (px >> 3);
            const bool ok = w < q.Wo && h < q.Ho;
            const size_t e0 = ok ? ((img_pix + (size_t)h * q.Wo + w) * q.G + a.g) * kP + p : 0;
            cp_async4(stage_addr + kStageOff + idx * 4, obase + e0, ok ? 4 : 0);
        }
#pragma unroll
        for (int it = 0; it < 5; ++it) {
            const int idx = lane + it * 32, px = idx / 5, wd = idx - px * 5;
            const int w = wb + (px & 7), h = hb + (px >> 3);
            const bool ok = w < q.Wo && h < q.Ho;
            const unsigned long long byte0 = ok ? (((img_pix + (size_t)h * q.Wo + w) * q.G + a.g) * kP * 2ull & ~3ull) + wd * 4 : 0ull;
            const long long rem = (long long)(pp.mask_bytes - byte0);
            const int nb = ok ? (int)(rem < 4 ? (rem < 0 ? 0 : rem) : 4) : 0;
            cp_async4(stage_addr + kStageMsk + (px * kMskWords + wd) * 4, mbase + (nb ? byte0 : 0ull), nb);
        }
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int idx = lane + it * 32, px = idx >> 1, ck = idx & 1;
            const int w = wb + (px & 7), h = hb + (px >> 3);
            const bool ok = w < q.Wo && h < q.Ho;
            const size_t e0 = ok ? (img_pix + (size_t)h * q.Wo + w) * C + a.g * q.gc + ck * 8 : 0;
            cp_async16(stage_addr + kStageGout + px * kSliceBytes + ck * 16, grad_out + e0, ok ? 16 : 0);
        }
        cp_async_commit();
    };

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    {   // prologue: zero this warp's A tile and spill counter, request the first staging data
        for (int i = lane; i < kABytes / 16; i += 32) sts128_zero(a_base + i * 16);
        if (lane == 0) *reinterpret_cast<unsigned *>(spill) = 0u;
        prefetch(cur, 0);
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, kWinBytes);
        tma_load_4d(win, &tmap, &bar, cur.g * q.gc, cur.ox, cur.oy, cur.n);
    }

    float acc[kMTiles][2][4];

    for (int it = 0;; ++it) {
        const int t_next = t + gridDim.x;
        const bool has_next = t_next < pp.total_tiles;
        TileAt nxt = cur;
        if (has_next) nxt = decode_tile(t_next, q, pp);
        const size_t img_base = (size_t)cur.n * q.H * row_stride + cur.g * q.gc;
        const size_t img_pix = (size_t)cur.n * q.Ho * q.Wo;
        float *gv_img = gv_acc + img_base;

#pragma unroll
        for (int b = 0; b < kMTiles; ++b)
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[b][n][e] = 0.f;

        mbar_wait(&bar, it & 1);   // value window of this tile has landed

#pragma unroll 1
        for (int s = 0; s < kSteps; ++s) {
            const int wo = cur.wo0 + warp * kStripW + px_x, ho = cur.ho0 + s * kPatchH + px_y;
            const bool live = wo < q.Wo && ho < q.Ho;
            cp_async_wait_all();
            __syncwarp();

            // ------------------------------------------------------------ gather + A build
            if (live) {
                Ctx<T> c;
                c.H = q.H; c.W = q.W; c.C = C; c.row_stride = row_stride; c.oy = cur.oy; c.ox = cur.ox;
                c.band_row0 = pp.byw + s * kPatchH; c.band_col0 = pp.bxw + warp * kStripW;
                c.j = j; c.half = half; c.lane = lane;
                c.win_addr = smem_u32(win) + half * 16;
                c.abuf = abuf; c.spill = spill;
                c.img = value + img_base; c.gv_img = gv_img;
                c.gq_a = *reinterpret_cast<const uint4 *>(s_gout + lane * kSliceBytes + half * 16);
                c.gq_b = *reinterpret_cast<const uint4 *>(s_gout + lane * kSliceBytes + (half ^ 1) * 16);
                const float base_w = axis_base(wo, 3, 1, q.pw, 1, q.sigma);
                const float base_h = axis_base(ho, 3, 1, q.ph, 1, q.sigma);
                // mask run of this pixel starts at element e0 of the tensor; staged from the
                // enclosing 4-byte words, so it sits at a 0- or 1-element shift
                const unsigned sh = (unsigned)((((img_pix + (size_t)ho * q.Wo + wo) * q.G + cur.g) * kP) & 1);
                uint16_t *mrow = s_msk + lane * (kMskWords * 2) + sh;
#pragma unroll 1
                for (int i = 0; i < 3; ++i) {
#pragma unroll
                    for (int jj = 0; jj < 3; ++jj) {
                        const int p = i * 3 + jj;
                        const float2 d = unpack2(s_off[lane * kP + p], T());
                        const float m = f32_of(mrow[p], T());
                        const float loc_w = base_w + ((float)i + d.x) * q.sigma;
                        const float loc_h = base_h + ((float)jj + d.y) * q.sigma;
                        float gm, gx, gy;
                        process_point<T>(c, loc_h, loc_w, m, gm, gx, gy);
                        s_off[lane * kP + p] = pack2(q.sigma * gx, q.sigma * gy, T());
                        mrow[p] = bits16(gm, T());
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
                            const uint32_t g2 = *reinterpret_cast<const uint32_t *>(s_gout + se[e].lane * kSliceBytes + chp * 2);
                            const float2 gf = unpack2(g2, T());
                            if (cf != 0.f) red_add2(gv_img + (ptrdiff_t)hh * row_stride + (ptrdiff_t)ww * C + chp, cf * gf.x, cf * gf.y);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) *reinterpret_cast<volatile unsigned *>(spill) = 0u;
                }
            }

            // ---- grad_offset / grad_mask of this step: coalesced write-out of the staged results
            {
                const int wb = cur.wo0 + warp * kStripW, hb = cur.ho0 + s * kPatchH;
                uint32_t *ob = reinterpret_cast<uint32_t *>(grad_offset);
                uint16_t *mb = reinterpret_cast<uint16_t *>(grad_mask);
#pragma unroll
                for (int k = 0; k < kP; ++k) {
                    const int idx = lane + k * 32, px = idx / kP, p = idx - px * kP;
                    const int w = wb + (px & 7), h = hb + (px >> 3);
                    if (w < q.Wo && h < q.Ho) {
                        const size_t e0 = ((img_pix + (size_t)h * q.Wo + w) * q.G + cur.g) * kP;
                        ob[e0 + p] = s_off[idx];
                        mb[e0 + p] = s_msk[px * (kMskWords * 2) + (unsigned)(e0 & 1) + p];
                    }
                }
            }

            // ---- B fragments (grad_out of the 32 pixels), then the staging buffer is free
            uint32_t bf[2][4];   // [k-step][{n0:k0-7, n0:k8-15, n1:k0-7, n1:k8-15}]
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const int px = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                ldmatrix_x4_trans(bf[ks], smem_u32(s_gout) + px * kSliceBytes + (lane >> 4) * 16);
            }
            __syncwarp();
            if (s + 1 < kSteps) prefetch(cur, s + 1);
            else if (has_next) prefetch(nxt, 0);

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
                const int kc_in = lane >> 4;
#pragma unroll
                for (int b = 0; b < kMTiles; ++b) {
                    const int row = b * 16 + r_in;
                    const uint32_t sw = ((uint32_t)row >> 1) & 3u;
                    uint32_t a0[4], a1[4];
                    ldmatrix_x4(a0, a_base + (uint32_t)row * 64u + (((uint32_t)kc_in ^ sw) << 4));
                    ldmatrix_x4(a1, a_base + (uint32_t)row * 64u + (((uint32_t)(2 + kc_in) ^ sw) << 4));
                    mma16816(acc[b][0], a0, bf[0][0], bf[0][1], T());
                    mma16816(acc[b][1], a0, bf[0][2], bf[0][3], T());
                    mma16816(acc[b][0], a1, bf[1][0], bf[1][1], T());
                    mma16816(acc[b][1], a1, bf[1][2], bf[1][3], T());
                    // the m-tile is in registers: clear it for the next step
                    sts128_zero(a_base + b * 1024 + lane * 16);
                    sts128_zero(a_base + b * 1024 + 512 + lane * 16);
                }
            }

            // ------------------------------------------------------------ finished rows leave
            const int my0 = cur.oy + pp.byw + s * kPatchH;
            const int mx0 = cur.ox + pp.bxw + warp * kStripW;
            if (s + 1 < kSteps) {
                flush_rows<kPatchH>(acc, gv_img, my0, mx0, q.H, q.W, row_stride, C, lane);
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
                flush_rows<kMTiles>(acc, gv_img, my0, mx0, q.H, q.W, row_stride, C, lane);
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

}  // namespace dcnv3
