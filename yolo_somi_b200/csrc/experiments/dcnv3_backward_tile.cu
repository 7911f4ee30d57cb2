// dcnv3_backward_tile.cu -- DCNv3 core backward, shared-memory tiled variant (the fast path).
//
// Why: the direct kernel (dcnv3_backward.cu) sends every corner contribution to L2 as a 16-byte
// vector reduction: 229.5 M RED ops at N=16, 80x80, C=256, G=16, which the L1/L2 reduction path
// retires at ~215 G op/s -> ~1.07 ms (profiles/README.md).  Here contributions are first summed in
// shared memory and each CTA sends only its accumulated window to L2 (~25x fewer RED ops).
//
// A CTA (4 warps) owns an 8x16 (w x h) tile of output pixels of one (image, group); a warp owns
// 4 rows of it.  Per kernel column (kh sampling points at a time):
//   G phase, thread <-> pixel.  The value window (18x26 pixels, one TMA box, zero-filled outside the
//     map) is gathered exactly like the forward (conflict-free rotated LDS.128, see
//     dcnv3_forward_tile.cu); the four per-corner dot products with the upstream gradient use
//     exact FHFMA products.  grad_offset / grad_mask are staged in shared memory (written out
//     coalesced at the end); the point leaves a record {top-left cell, 4 coefficients w_k*m}.
//   S phase, half-warp <-> record, lane <-> channel.  Each warp adds the records of ITS OWN 32
//     pixels into a private fp32 band of the accumulator (15 window rows x 18 columns x 16
//     channels), so warps never touch the same shared-memory word and plain LDS/FFMA/STS
//     read-modify-writes replace atomics (shared-memory float atomics are CAS loops on this
//     architecture).  A record's 16 channels are one contiguous 64-byte run, so the two records
//     of a step cost one or two wavefronts per access -- this is what makes the scatter
//     shared-memory-efficient (a lane-per-record mapping was 4x slower: random 16-byte accesses).
//     The two half-warps only collide when their records' 2x2 blocks overlap; that step is
//     then run as two passes.  The accumulation order inside a CTA is fixed.
// Points whose 2x2 corner block leaves the window (or the warp's band) fall back to direct global
// reductions.  Finally the bands are summed and added to the fp32 grad_value accumulator in
// global memory with 128-bit reductions (cells outside the map and all-zero pieces are skipped).
#include "../dcnv3_common.cuh"
#include "../dcnv3_launch.h"
#include "../dcnv3_tma.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace dcnv3 {

constexpr int kBTileW = 8, kBTileH = 16;           // output pixels per tile
constexpr int kBThreads = kBTileW * kBTileH;       // 128: one thread per pixel, 4 warps
constexpr int kBWarps = kBThreads / 32;
constexpr int kBRowsPerWarp = kBTileH / kBWarps;   // 4 tile rows per warp
constexpr int kBWinW = 18, kBWinH = 26;            // window (value pixels == accumulator cells)
constexpr int kBCells = kBWinW * kBWinH;
constexpr int kBBandH = kBRowsPerWarp + (kBWinH - kBTileH) + 1;   // 15 window rows reachable by a warp
constexpr int kBBandCells = kBBandH * kBWinW;
constexpr int kBSliceBytes = 32;                   // value bytes per pixel per CTA
static_assert(kBWinW % 4 == 2, "window width must be 2 mod 4 (conflict-free corner layout)");
static_assert(kBTileW == 8, "a quarter-warp must be one tile row (rotation scheme)");

struct BwdTileParams {
    int ox_rel, oy_rel;      // window origin relative to (wo0*stride_w, ho0*stride_h)
    int tiles_x;
    int slices_per_group;
    int n0;
};

template <typename T> struct BwdPairOf { using type = uint32_t; };
template <> struct BwdPairOf<float> { using type = float2; };
__device__ __forceinline__ float2 bpair_to_f32(uint32_t w, __half) { return unpack2(w, __half()); }
__device__ __forceinline__ float2 bpair_to_f32(uint32_t w, __nv_bfloat16) { return unpack2(w, __nv_bfloat16()); }
__device__ __forceinline__ float2 bpair_to_f32(float2 w, float) { return w; }
__device__ __forceinline__ uint32_t bpair_from_f32(float a, float b, __half) { return pack2(a, b, __half()); }
__device__ __forceinline__ uint32_t bpair_from_f32(float a, float b, __nv_bfloat16) { return pack2(a, b, __nv_bfloat16()); }
__device__ __forceinline__ float2 bpair_from_f32(float a, float b, float) { return make_float2(a, b); }

// shared-memory carve-up (bytes); every region is 16-byte aligned
template <typename T> struct BwdTileLayout {
    static constexpr int SLICE = kBSliceBytes / sizeof(T);      // channels per CTA
    static constexpr size_t win = 0;                            // [kBWinH][kBWinW][32 B]
    static constexpr size_t acc = win + (size_t)kBCells * kBSliceBytes;          // [warp][band cell][SLICE] fp32
    static constexpr size_t gout = acc + (size_t)kBWarps * kBBandCells * SLICE * 4;   // [128][32 B]
    static constexpr size_t coef = gout + (size_t)kBThreads * kBSliceBytes;      // [kh][128] float4
    __host__ __device__ static size_t cell(int kh) { return coef + (size_t)kh * kBThreads * 16; }    // [kh][128] int
    __host__ __device__ static size_t off(int kh) { return cell(kh) + (size_t)kh * kBThreads * 4; }  // [128][P] pairs
    __host__ __device__ static size_t msk(int kh, int P) { return off(kh) + (size_t)kBThreads * P * 2 * sizeof(T); }
    __host__ __device__ static size_t total(int kh, int P) { return (msk(kh, P) + (size_t)kBThreads * P * sizeof(T) + 15) & ~(size_t)15; }
};

__device__ __forceinline__ uint4 blds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
template <typename V> __device__ __forceinline__ void brotate4(V (&x)[4], int r) {  // out[t] = in[(t+r)&3]
    if (r & 1) { const V t = x[0]; x[0] = x[1]; x[1] = x[2]; x[2] = x[3]; x[3] = t; }
    if (r & 2) { V t = x[0]; x[0] = x[2]; x[2] = t; t = x[1]; x[1] = x[3]; x[3] = t; }
}

__device__ __forceinline__ void red_add(float *p, float4 v) { atomicAdd(reinterpret_cast<float4 *>(p), v); }
__device__ __forceinline__ void red_add(float *p, float2 v) { atomicAdd(reinterpret_cast<float2 *>(p), v); }

template <typename T, int KH, int KW>
__global__ void __launch_bounds__(kBThreads)
bwd_tile(const __grid_constant__ CUtensorMap tmap, const T *__restrict__ value,
         const T *__restrict__ offset, const T *__restrict__ mask, const T *__restrict__ grad_out,
         float *__restrict__ gv_acc, T *__restrict__ grad_offset, T *__restrict__ grad_mask,
         const Geom q, const BwdTileParams tp) {
    using L = BwdTileLayout<T>;
    using Pair = typename BwdPairOf<T>::type;
    constexpr int E = Chunk<T>::kElems;
    constexpr int SLICE = L::SLICE;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;

    const int kh = KH ? KH : q.kh, kw = KW ? KW : q.kw;
    const int P = kh * kw;
    unsigned char *win = smem + L::win;
    float *s_acc = reinterpret_cast<float *>(smem + L::acc);
    unsigned char *s_gout = smem + L::gout;
    float4 *s_coef = reinterpret_cast<float4 *>(smem + L::coef);
    int *s_cell = reinterpret_cast<int *>(smem + L::cell(kh));
    Pair *s_off = reinterpret_cast<Pair *>(smem + L::off(kh));
    T *s_msk = reinterpret_cast<T *>(smem + L::msk(kh, P));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_x = blockIdx.x % tp.tiles_x, tile_y = blockIdx.x / tp.tiles_x;
    const int g = blockIdx.y / tp.slices_per_group, sub = blockIdx.y % tp.slices_per_group;
    const int n = tp.n0 + blockIdx.z;
    const int wo0 = tile_x * kBTileW, ho0 = tile_y * kBTileH;
    const int wo = wo0 + (tid % kBTileW), ho = ho0 + (tid / kBTileW);
    const bool live = wo < q.Wo && ho < q.Ho;
    const int ox = wo0 * q.sw + tp.ox_rel, oy = ho0 * q.sh + tp.oy_rel;
    const int C = q.G * q.gc;
    const int row_stride = q.W * C;
    const int ch0 = g * q.gc + sub * SLICE;

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, kBCells * kBSliceBytes);
        tma_load_4d(win, &tmap, &bar, ch0, ox, oy, n);
    }

    // ---- while the box is in flight: zero the accumulator, stage offsets / masks / grad_out
    {
        float4 *z = reinterpret_cast<float4 *>(s_acc);
        for (int i = tid; i < kBWarps * kBBandCells * SLICE / 4; i += kBThreads) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        const size_t img_pix = (size_t)n * q.Ho * q.Wo;
        for (int idx = tid; idx < kBThreads * P; idx += kBThreads) {
            const int px = idx / P, p = idx - px * P;
            const int w = wo0 + (px % kBTileW), h = ho0 + (px / kBTileW);
            if (w < q.Wo && h < q.Ho) {
                const size_t pgi = (img_pix + (size_t)h * q.Wo + w) * q.G + g;
                s_off[idx] = __ldg(reinterpret_cast<const Pair *>(offset) + pgi * P + p);
                s_msk[idx] = __ldg(mask + pgi * P + p);
            }
        }
        for (int idx = tid; idx < kBThreads * 2; idx += kBThreads) {   // 2 x 16-byte chunks per pixel
            const int px = idx >> 1, c = idx & 1;
            const int w = wo0 + (px % kBTileW), h = ho0 + (px / kBTileW);
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (w < q.Wo && h < q.Ho)
                v = __ldg(reinterpret_cast<const uint4 *>(grad_out + (img_pix + (size_t)h * q.Wo + w) * C + ch0 + c * E));
            *reinterpret_cast<uint4 *>(s_gout + px * kBSliceBytes + c * 16) = v;
        }
    }

    const int j = tid & 7;                         // lane within the quarter-warp
    const int half = j & 1;                        // 16-byte chunk read first
    const float base_w = axis_base(wo, kw, q.sw, q.pw, q.dw, q.sigma);
    const float base_h = axis_base(ho, kh, q.sh, q.ph, q.dh, q.sigma);
    const uint32_t win_addr = smem_u32(win) + half * 16;
    const size_t img_base = (size_t)n * q.H * row_stride + ch0;   // element index of (n, 0, 0, ch0)
    const T *img = value + img_base;

    __syncthreads();
    mbar_wait(&bar, 0);

    // upstream gradient of this thread's pixel: chunk `half` and the other chunk, packed
    const uint4 gq_a = *reinterpret_cast<const uint4 *>(s_gout + tid * kBSliceBytes + half * 16);
    const uint4 gq_b = *reinterpret_cast<const uint4 *>(s_gout + tid * kBSliceBytes + (half ^ 1) * 16);

    for (int i = 0; i < kw; ++i) {
        // ------------------------------------------------------------------ G phase (column i)
#pragma unroll
        for (int jj = 0; jj < (KH ? KH : 8); ++jj) {
            if (jj >= kh) break;
            const int p = i * kh + jj;
            int rec_cell = -1;
            float4 rec_coef = make_float4(0.f, 0.f, 0.f, 0.f);
            float gm = 0.f, gx = 0.f, gy = 0.f;
            if (live) {
                const float2 d = bpair_to_f32(s_off[tid * P + p], T());
                const float m = to_f32(s_msk[tid * P + p]);
                const float loc_w = base_w + ((float)(i * q.dw) + d.x) * q.sigma;
                const float loc_h = base_h + ((float)(jj * q.dh) + d.y) * q.sigma;
                // range test of the reference (dcnv3_im2col_cuda.cuh:262-263); also rejects NaN
                const bool inside = loc_h > -1.f && loc_w > -1.f && loc_h < (float)q.H && loc_w < (float)q.W;
                const float fh = floorf(loc_h), fw = floorf(loc_w);
                const float lh = loc_h - fh, lw = loc_w - fw, hh = 1.f - lh, hw = 1.f - lw;
                const int hwin = (int)fh - oy, wwin = (int)fw - ox;
                if (!inside) {
                    // contributes nothing (all three gradients are zero)
                } else if ((unsigned)hwin < (unsigned)(kBWinH - 1) && (unsigned)wwin < (unsigned)(kBWinW - 1)) {
                    // ---- window path; out-of-map corners read zeros and their cells are dropped
                    //      at flush time, so no range test is needed here
                    int o[4] = {0, kBSliceBytes, kBWinW * kBSliceBytes, kBWinW * kBSliceBytes + kBSliceBytes};
                    const int rho = ((j >> 1) - (wwin + 2 * hwin)) & 3;
                    brotate4(o, rho);
                    const uint32_t tl = win_addr + (uint32_t)(hwin * kBWinW + wwin) * kBSliceBytes;
                    float dr[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const uint32_t a = tl + o[t];
                        const uint4 qa = blds128(a), qb = blds128(a ^ 16u);
                        dr[t] = dot<T>(gq_a, qa, 0.f) + dot<T>(gq_b, qb, 0.f);
                    }
                    brotate4(dr, (4 - rho) & 3);   // back to corner order TL, TR, BL, BR
                    gm = hh * hw * dr[0] + hh * lw * dr[1] + lh * hw * dr[2] + lh * lw * dr[3];
                    gx = m * (hh * (dr[1] - dr[0]) + lh * (dr[3] - dr[2]));
                    gy = m * (hw * (dr[2] - dr[0]) + lw * (dr[3] - dr[1]));
                    rec_cell = hwin * kBWinW + wwin;
                    rec_coef = make_float4(hh * hw * m, hh * lw * m, lh * hw * m, lh * lw * m);
                } else {
                    const ClampedTap ct = make_clamped_tap(loc_h, loc_w, q.H, q.W);
                    {
                        // ---- fallback: clamped global reads, direct reductions
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
                        const float wk[4] = {fy_lo * fx_lo, fy_lo * fx_hi, fy_hi * fx_lo, fy_hi * fx_hi};
                        gm = wk[0] * dk[0] + wk[1] * dk[1] + wk[2] * dk[2] + wk[3] * dk[3];
                        gx = m * (fy_lo * (ct.rig * dk[1] - ct.lef * dk[0]) + fy_hi * (ct.rig * dk[3] - ct.lef * dk[2]));
                        gy = m * (fx_lo * (ct.bot * dk[2] - ct.top * dk[0]) + fx_hi * (ct.bot * dk[3] - ct.top * dk[1]));
                        float ga[E], gb[E];
                        unpack<T>(gq_a, ga);
                        unpack<T>(gq_b, gb);
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const float c = wk[t] * m;
                            if (c != 0.f) {
                                float *dst = gv_acc + img_base + at[t];
#pragma unroll
                                for (int e = 0; e < E; e += 4) {
                                    red_add(dst + ea + e, make_float4(c * ga[e], c * ga[e + 1], c * ga[e + 2], c * ga[e + 3]));
                                    red_add(dst + eb + e, make_float4(c * gb[e], c * gb[e + 1], c * gb[e + 2], c * gb[e + 3]));
                                }
                            }
                        }
                    }
                }
                s_off[tid * P + p] = bpair_from_f32(q.sigma * gx, q.sigma * gy, T());
                s_msk[tid * P + p] = from_f32<T>(gm);
            }
            s_cell[jj * kBThreads + tid] = rec_cell;
            s_coef[jj * kBThreads + tid] = rec_coef;
        }
        __syncthreads();
        // ------------------------------------------------------------------ S phase (column i)
        // half-warp <-> record, lane <-> channel; a warp scatters the records of its own 32 pixels
        // into its private band.  Records r and r+16 (two tile rows apart) share a step.
        {
            const int hl = lane >> 4, ch = lane & 15;
            const bool ch_ok = ch < SLICE;
            float *band = s_acc + (size_t)warp * kBBandCells * SLICE;
            const int band_cell0 = warp * kBRowsPerWarp * kBWinW;   // first window cell of the band
            const int rbase = warp * 32 + 16 * hl;
            const int n_steps = kh * 16;                            // (jj, step) flattened
            // software pipeline: the next step's record is fetched while this one is accumulated
            int cell = s_cell[rbase];
            float4 cf = s_coef[rbase];
            float gch = ch_ok ? to_f32(reinterpret_cast<const T *>(s_gout + rbase * kBSliceBytes)[ch]) : 0.f;
            for (int it = 0; it < n_steps; ++it) {
                const int cur_cell = cell;
                const float4 cur_cf = cf;
                const float cur_g = gch;
                if (it + 1 < n_steps) {
                    const int nx = it + 1, r = rbase + (nx & 15);
                    cell = s_cell[(nx >> 4) * kBThreads + r];
                    cf = s_coef[(nx >> 4) * kBThreads + r];
                    gch = ch_ok ? to_f32(reinterpret_cast<const T *>(s_gout + r * kBSliceBytes)[ch]) : 0.f;
                }
                const int cb = cur_cell - band_cell0;
                const bool act = cur_cell >= 0;
                const bool in_band = act && cb >= 0 && cb < (kBBandH - 1) * kBWinW - 1;
                // the two records of a step may be accumulated together unless their 2x2 blocks
                // overlap: |delta cell| in {0, 1, W-1, W, W+1}
                const int other = __shfl_xor_sync(0xffffffffu, in_band ? cur_cell : -100000, 16);
                const int ad = abs(other - cur_cell);
                const bool clash = in_band && (ad <= 1 || (ad >= kBWinW - 1 && ad <= kBWinW + 1));
                const bool any_clash = __any_sync(0xffffffffu, clash);
                float *c0 = band + (size_t)(in_band ? cb : 0) * SLICE + ch;
                const bool mine = in_band && ch_ok;
                for (int pass = 0; pass < 2; ++pass) {
                    const bool go = mine && (any_clash ? (pass == hl) : (pass == 0));
                    if (go) {
                        const float a0 = c0[0], a1 = c0[SLICE], a2 = c0[kBWinW * SLICE], a3 = c0[(kBWinW + 1) * SLICE];
                        c0[0] = a0 + cur_cf.x * cur_g;
                        c0[SLICE] = a1 + cur_cf.y * cur_g;
                        c0[kBWinW * SLICE] = a2 + cur_cf.z * cur_g;
                        c0[(kBWinW + 1) * SLICE] = a3 + cur_cf.w * cur_g;
                    }
                    __syncwarp();
                    if (!any_clash) break;
                }
                if (act && !in_band && ch_ok) {
                    // in the window but outside this warp's band (large vertical offset): rare
                    const int y = oy + cur_cell / kBWinW, x = ox + cur_cell % kBWinW;
                    const float cfs[4] = {cur_cf.x, cur_cf.y, cur_cf.z, cur_cf.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int yy = y + (k >> 1), xx = x + (k & 1);
                        if ((unsigned)yy < (unsigned)q.H && (unsigned)xx < (unsigned)q.W && cfs[k] != 0.f)
                            atomicAdd(gv_acc + img_base + (size_t)yy * row_stride + (size_t)xx * C + ch, cfs[k] * cur_g);
                    }
                }
            }
        }
        __syncthreads();
    }

    // ---- grad_offset / grad_mask: coalesced write-out of the staged values
    {
        const size_t img_pix = (size_t)n * q.Ho * q.Wo;
        for (int idx = tid; idx < kBThreads * P; idx += kBThreads) {
            const int px = idx / P, p = idx - px * P;
            const int w = wo0 + (px % kBTileW), h = ho0 + (px / kBTileW);
            if (w < q.Wo && h < q.Ho) {
                const size_t pgi = (img_pix + (size_t)h * q.Wo + w) * q.G + g;
                reinterpret_cast<Pair *>(grad_offset)[pgi * P + p] = s_off[idx];
                grad_mask[pgi * P + p] = s_msk[idx];
            }
        }
    }
    // ---- flush: sum the warps' bands per window cell and add to the global accumulator;
    //      lane <-> (cell, 16-byte piece), piece fastest => a cell's pieces are contiguous in global
    constexpr int PIECES = SLICE / 4;
    for (int idx = tid; idx < kBCells * PIECES; idx += kBThreads) {
        const int cell = idx / PIECES, piece = idx % PIECES;
        const int wy = cell / kBWinW, wx = cell % kBWinW;
        const int y = oy + wy, x = ox + wx;
        if ((unsigned)y < (unsigned)q.H && (unsigned)x < (unsigned)q.W) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int w = 0; w < kBWarps; ++w) {
                const int by = wy - w * kBRowsPerWarp;
                if (by >= 0 && by < kBBandH) {
                    const float4 t = *reinterpret_cast<const float4 *>(
                        s_acc + ((size_t)w * kBBandCells + by * kBWinW + wx) * SLICE + piece * 4);
                    v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
                }
            }
            if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
                red_add(gv_acc + img_base + (size_t)y * row_stride + (size_t)x * C + piece * 4, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
template <typename T>
static bool launch_bwd_tile_typed(const void *value, const void *offset, const void *mask,
                                  const void *grad_out, float *gv_acc, void *grad_offset,
                                  void *grad_mask, const Geom &q, int dtype, cudaStream_t stream,
                                  cudaError_t *err) {
    using L = BwdTileLayout<T>;
    constexpr int SLICE = L::SLICE;
    // a group must be exactly one slice: with several slices per group each CTA would see only part
    // of the channel sum of grad_offset / grad_mask
    if (q.gc != SLICE) return false;
    if (q.kh > 8) return false;
    if (((uintptr_t)value | (uintptr_t)grad_out | (uintptr_t)gv_acc) % 16) return false;
    const float span_w = (kBTileW - 1) * q.sw + (q.kw - 1) * q.dw * q.sigma;
    const float span_h = (kBTileH - 1) * q.sh + (q.kh - 1) * q.dh * q.sigma;
    if (!(q.sigma > 0.f) || span_w + 4 > kBWinW - 2 || span_h + 4 > kBWinH - 2) return false;
    const int C = q.G * q.gc;
    const int P = q.kh * q.kw;
    const size_t smem = L::total(q.kh, P);
    if (smem > 200 * 1024) return false;
    CUtensorMap tmap;
    if (!make_nhwc_tensor_map(&tmap, value, dtype, q.N, q.H, q.W, C, SLICE, kBWinW, kBWinH)) return false;

    BwdTileParams tp;
    const int cw = (q.dw * (q.kw - 1)) >> 1, chh = (q.dh * (q.kh - 1)) >> 1;
    const float a_w = (float)(cw - q.pw) - cw * q.sigma, a_h = (float)(chh - q.ph) - chh * q.sigma;
    tp.ox_rel = (int)std::floor(a_w + 0.5f * span_w - 0.5f * (kBWinW - 2));
    tp.oy_rel = (int)std::floor(a_h + 0.5f * span_h - 0.5f * (kBWinH - 2));
    tp.tiles_x = (q.Wo + kBTileW - 1) / kBTileW;
    tp.slices_per_group = 1;
    const int tiles_y = (q.Ho + kBTileH - 1) / kBTileH;
    if (q.G > 65535) return false;
    const T *v = static_cast<const T *>(value), *o = static_cast<const T *>(offset),
            *m = static_cast<const T *>(mask), *go = static_cast<const T *>(grad_out);
    T *goff = static_cast<T *>(grad_offset), *gmsk = static_cast<T *>(grad_mask);
    const bool k33 = q.kh == 3 && q.kw == 3;
    if (k33) cudaFuncSetAttribute(bwd_tile<T, 3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else cudaFuncSetAttribute(bwd_tile<T, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int n0 = 0; n0 < q.N; n0 += 65535) {
        tp.n0 = n0;
        const dim3 grid((unsigned)(tp.tiles_x * tiles_y), (unsigned)q.G, (unsigned)std::min(65535, q.N - n0));
        if (k33) bwd_tile<T, 3, 3><<<grid, kBThreads, smem, stream>>>(tmap, v, o, m, go, gv_acc, goff, gmsk, q, tp);
        else bwd_tile<T, 0, 0><<<grid, kBThreads, smem, stream>>>(tmap, v, o, m, go, gv_acc, goff, gmsk, q, tp);
    }
    *err = cudaGetLastError();
    return true;
}

// gv_acc: zero-initialised fp32 accumulator with the shape of value.  Returns false when the
// shape is not eligible (caller uses the direct kernel).
bool try_launch_backward_tile(const void *value, const void *offset, const void *mask,
                              const void *grad_out, float *gv_acc, void *grad_offset,
                              void *grad_mask, const Geom &q, int dtype, cudaStream_t stream,
                              cudaError_t *err) {
    // Opt-in (DCNV3_BWD=tile): correct and parity-tested, but on B200 it is not yet faster than the
    // direct kernel -- the scatter phase issues ~45 warp-instructions per sampling point at 8
    // resident warps per SM (profiles/README.md, r1 bwd_tile).  Kept as the base for the next step.
    const char *e = std::getenv("DCNV3_BWD");
    if (!(e && e[0] == 't')) return false;
    if ((long long)q.N * q.Ho * q.Wo == 0) return false;
    switch (dtype) {
    case 0: return launch_bwd_tile_typed<float>(value, offset, mask, grad_out, gv_acc, grad_offset, grad_mask, q, dtype, stream, err);
    case 1: return launch_bwd_tile_typed<__half>(value, offset, mask, grad_out, gv_acc, grad_offset, grad_mask, q, dtype, stream, err);
    default: return launch_bwd_tile_typed<__nv_bfloat16>(value, offset, mask, grad_out, gv_acc, grad_offset, grad_mask, q, dtype, stream, err);
    }
}

}  // namespace dcnv3
