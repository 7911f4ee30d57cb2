// dcnv3_forward_tile.cu -- DCNv3 core forward, shared-memory tiled variant (the fast path).
//
// Why: the direct kernel (dcnv3_forward.cu) is bound by the L1 global-load path, which retires
// about one 32-byte sector per cycle per SM; every bilinear corner of every sampling point is one
// sector, so N=16, 80x80, C=256, G=16 costs ~59 M sector-cycles / 148 SMs ~ 200 us (measured
// 172-198 us, profiles/README.md).  Shared memory delivers 128 B/cycle when bank conflicts are
// avoided -- four corners per cycle.
//
// How: a CTA owns a 16x16 tile of output pixels of one (image, 32-byte channel slice of a group).
//   * one thread issues ONE TMA box load (cp.async.bulk.tensor.4d) of the slice's 26x26-pixel value
//     window around the tile into shared memory, dense [h][w][32 B]; out-of-map pixels are
//     zero-filled by the TMA unit, which implements the op's zero padding for free;
//   * while the box is in flight the CTA stages the tile's (dx,dy) offsets and mask weights in
//     shared memory with coalesced loads (a pixel's pairs are contiguous, pixels are G*P apart);
//   * each thread then walks its sampling points.  A point whose 2x2 corner block lies inside
//     the window (offsets within about +-4 px of the kernel tap) is served from shared memory with
//     eight conflict-free LDS.128; any other point falls back to clamped global loads.
//
// Conflict-free gather.  The window is 26 pixels wide (26 = 2 mod 4), so pixel (h,w) sits in
// 32-byte bank slot (w + 2h) mod 4 and the four corners of ANY 2x2 block occupy the four slots
// exactly once.  In each LDS.128 the eight lanes of a quarter-warp must hit eight different
// 16-byte slots: lane j visits its corners in the rotated order c_t = (rho + t) mod 4 with
// rho = (j/2 - slot(top-left)) mod 4 and reads half (j mod 2) first, the other half second, so in
// step t lane j touches slot 2*((j/2 + t) mod 4) + (j mod 2) -- a permutation of 0..7.  The
// bilinear weights are rotated the same way (two conditional-swap stages).
//
// 16-bit data is never unpacked (FHFMA, see dcnv3_common.cuh); accumulation is fp32.
#include "dcnv3_common.cuh"
#include "dcnv3_launch.h"
#include "dcnv3_stage.cuh"
#include "dcnv3_tma.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace dcnv3 {

constexpr int kTile = 16;               // output pixels per tile side
constexpr int kWin = 26;                // value window side (must be 2 mod 4)
constexpr int kSliceBytes = 32;         // channel bytes of one pixel held by a CTA
constexpr int kTileThreads = kTile * kTile;
static_assert(kWin % 4 == 2, "window width must be 2 mod 4 for the conflict-free corner layout");

struct TileParams {
    int ox_rel, oy_rel;      // window origin relative to (wo0*stride_w, ho0*stride_h)
    int tiles_x;             // tiles per output row
    int slices_per_group;    // 32-byte channel slices per group
    int n0;                  // first image of this launch (gridDim.z limit)
};

constexpr int kWinBytes = kWin * kWin * kSliceBytes;

template <typename T> struct PairOf { using type = uint32_t; };   // two 16-bit values
template <> struct PairOf<float> { using type = float2; };
__device__ __forceinline__ float2 pair_to_f32(uint32_t w, __half) { return unpack2(w, __half()); }
__device__ __forceinline__ float2 pair_to_f32(uint32_t w, __nv_bfloat16) { return unpack2(w, __nv_bfloat16()); }
__device__ __forceinline__ float2 pair_to_f32(float2 w, float) { return w; }

template <typename T> static size_t tile_smem_bytes(int P) {
    return kWinBytes + (size_t)kTileThreads * P * (2 * sizeof(T)) + (size_t)kTileThreads * P * sizeof(T);
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// rotate a 4-vector left by r: out[t] = in[(t + r) & 3]
template <typename V> __device__ __forceinline__ void rotate4(V (&x)[4], int r) {
    if (r & 1) { const V t = x[0]; x[0] = x[1]; x[1] = x[2]; x[2] = x[3]; x[3] = t; }
    if (r & 2) { V t = x[0]; x[0] = x[2]; x[2] = t; t = x[1]; x[1] = x[3]; x[3] = t; }
}

template <typename T, bool FAST, int KH, int KW>
__global__ void __launch_bounds__(kTileThreads)
fwd_tile(const __grid_constant__ CUtensorMap tmap, const T *__restrict__ value,
         const T *__restrict__ offset, const T *__restrict__ mask, T *__restrict__ out,
         const Geom q, const TileParams tp) {
    constexpr int E = Chunk<T>::kElems;          // channels per 16-byte chunk
    constexpr int SLICE = kSliceBytes / sizeof(T);  // channels per CTA slice (2 chunks)
    using Pair = typename PairOf<T>::type;        // one (dx,dy)
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *win = smem;                                             // [kWin][kWin][32 B]
    Pair *s_off = reinterpret_cast<Pair *>(smem + kWinBytes);              // [256][P]
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const int tile_x = blockIdx.x % tp.tiles_x, tile_y = blockIdx.x / tp.tiles_x;
    const int g = blockIdx.y / tp.slices_per_group, sub = blockIdx.y % tp.slices_per_group;
    const int n = tp.n0 + blockIdx.z;
    const int wo0 = tile_x * kTile, ho0 = tile_y * kTile;
    const int wo = wo0 + (tid % kTile), ho = ho0 + (tid / kTile);
    const bool live = wo < q.Wo && ho < q.Ho;
    const int ox = wo0 * q.sw + tp.ox_rel, oy = ho0 * q.sh + tp.oy_rel;
    const int ch0 = g * q.gc + sub * SLICE;

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, kWinBytes);
        tma_load_4d(win, &tmap, &bar, ch0, ox, oy, n);
    }

    const int kh = KH ? KH : q.kh, kw = KW ? KW : q.kw;
    const int P = kh * kw;
    const int C = q.G * q.gc;
    const int row_stride = q.W * C;
    T *s_msk = reinterpret_cast<T *>(s_off + kTileThreads * P);             // [256][P]

    // While the box is in flight: stage the tile's offsets and masks.  A pixel's P (dx,dy) pairs
    // are contiguous but pixels are G*P pairs apart, so consecutive threads take consecutive
    // pairs of the same pixel (coalesced) instead of each thread walking its own pixel.
    stage_offsets_masks<T, KH * KW, kTileThreads, kTile>(offset, mask, s_off, s_msk, P, tid, wo0, ho0, q.Wo, q.Ho,
                                                         q.G, g, (size_t)n * q.Ho * q.Wo);
    const size_t pg = (((size_t)n * q.Ho + (live ? ho : 0)) * q.Wo + (live ? wo : 0)) * q.G + g;
    const int j = tid & 7;                        // lane within the quarter-warp
    const int half = j & 1;                       // which 16-byte chunk this lane reads FIRST
    const T *img = value + (size_t)n * q.H * row_stride + ch0;

    const float base_w = axis_base(wo, kw, q.sw, q.pw, q.dw, q.sigma);
    const float base_h = axis_base(ho, kh, q.sh, q.ph, q.dh, q.sigma);
    const uint32_t win_addr = smem_u32(win) + half * 16;

    float acc_a[E], acc_b[E];   // acc_a: channels of chunk `half`, acc_b: the other chunk
#pragma unroll
    for (int v = 0; v < E; ++v) acc_a[v] = acc_b[v] = 0.f;

    __syncthreads();            // offsets / masks staged
    mbar_wait(&bar, 0);

    if (live) {
#pragma unroll
        for (int i = 0; i < kw; ++i) {
#pragma unroll
            for (int jj = 0; jj < kh; ++jj) {
                const int p = i * kh + jj;
                const float2 d = pair_to_f32(s_off[tid * P + p], T());
                const float m = to_f32(s_msk[tid * P + p]);
                const float loc_w = base_w + ((float)(i * q.dw) + d.x) * q.sigma;
                const float loc_h = base_h + ((float)(jj * q.dh) + d.y) * q.sigma;
                const bool inside = loc_h > -1.f && loc_w > -1.f && loc_h < (float)q.H && loc_w < (float)q.W;
                if (!inside) continue;
                const float fh = floorf(loc_h), fw = floorf(loc_w);
                const int h0 = (int)fh, w0 = (int)fw;
                const float lh = loc_h - fh, lw = loc_w - fw, hh = 1.f - lh, hw = 1.f - lw;
                const int hwin = h0 - oy, wwin = w0 - ox;   // top-left corner in window coordinates
                if ((unsigned)hwin < (unsigned)(kWin - 1) && (unsigned)wwin < (unsigned)(kWin - 1)) {
                    // ---- shared-memory path: zero padding already materialised by the TMA fill
                    float w[4] = {hh * hw * m, hh * lw * m, lh * hw * m, lh * lw * m};
                    int o[4] = {0, kSliceBytes, kWin * kSliceBytes, kWin * kSliceBytes + kSliceBytes};
                    const int rho = ((j >> 1) - (wwin + 2 * hwin)) & 3;
                    rotate4(w, rho);
                    rotate4(o, rho);
                    const uint32_t tl = win_addr + (uint32_t)(hwin * kWin + wwin) * kSliceBytes;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const uint32_t a = tl + o[t];
                        const uint4 qa = lds128(a), qb = lds128(a ^ 16u);
                        const Weight<T, FAST> wt(w[t]);
                        axpy<T, FAST>(acc_a, qa, wt);
                        axpy<T, FAST>(acc_b, qb, wt);
                    }
                } else {
                    // ---- fallback: the 2x2 block leaves the window; clamped global reads
                    const ClampedTap ct = make_clamped_tap(loc_h, loc_w, q.H, q.W);
                    const T *r_lo = img + ct.row_lo * row_stride, *r_hi = img + ct.row_hi * row_stride;
                    const int c_lo = ct.col_lo * C, c_hi = ct.col_hi * C;
                    const int ea = half * E, eb = (half ^ 1) * E;
                    const float fy_lo = ct.hh * ct.top * m, fy_hi = ct.lh * ct.bot * m;
                    const float fx_lo = ct.hw * ct.lef, fx_hi = ct.lw * ct.rig;
                    const T *corner[4] = {r_lo + c_lo, r_lo + c_hi, r_hi + c_lo, r_hi + c_hi};
                    const float wc[4] = {fy_lo * fx_lo, fy_lo * fx_hi, fy_hi * fx_lo, fy_hi * fx_hi};
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const uint4 qa = __ldg(reinterpret_cast<const uint4 *>(corner[t] + ea));
                        const uint4 qb = __ldg(reinterpret_cast<const uint4 *>(corner[t] + eb));
                        const Weight<T, FAST> wt(wc[t]);
                        axpy<T, FAST>(acc_a, qa, wt);
                        axpy<T, FAST>(acc_b, qb, wt);
                    }
                }
            }
        }
        T *dst = out + pg * q.gc + sub * SLICE;
        *reinterpret_cast<uint4 *>(dst + half * E) = pack<T>(acc_a);
        *reinterpret_cast<uint4 *>(dst + (half ^ 1) * E) = pack<T>(acc_b);
    }
}

// ---------------------------------------------------------------------------------------------
template <typename T>
static bool launch_tile_typed(const void *value, const void *offset, const void *mask, void *out,
                              const Geom &q, int dtype, bool fast, cudaStream_t stream, cudaError_t *err) {
    constexpr int SLICE = kSliceBytes / sizeof(T);
    if (q.gc % SLICE != 0) return false;
    if ((uintptr_t)value % 16 || (uintptr_t)out % 16) return false;
    // the tile's nominal tap span must leave at least 2 pixels of offset slack on each side
    const float span_w = (kTile - 1) * q.sw + (q.kw - 1) * q.dw * q.sigma;
    const float span_h = (kTile - 1) * q.sh + (q.kh - 1) * q.dh * q.sigma;
    if (!(q.sigma > 0.f) || span_w + 4 > kWin - 2 || span_h + 4 > kWin - 2) return false;
    const int C = q.G * q.gc;
    CUtensorMap tmap;
    if (!make_nhwc_tensor_map(&tmap, value, dtype, q.N, q.H, q.W, C, SLICE, kWin, kWin)) return false;

    TileParams tp;
    const int cw = (q.dw * (q.kw - 1)) >> 1, chh = (q.dh * (q.kh - 1)) >> 1;
    const float a_w = (float)(cw - q.pw) - cw * q.sigma, a_h = (float)(chh - q.ph) - chh * q.sigma;
    tp.ox_rel = (int)std::floor(a_w + 0.5f * span_w - 0.5f * (kWin - 2));
    tp.oy_rel = (int)std::floor(a_h + 0.5f * span_h - 0.5f * (kWin - 2));
    tp.tiles_x = (q.Wo + kTile - 1) / kTile;
    tp.slices_per_group = q.gc / SLICE;
    const int tiles_y = (q.Ho + kTile - 1) / kTile;
    const long long slices = (long long)q.G * tp.slices_per_group;
    if (slices > 65535) return false;
    const T *v = static_cast<const T *>(value), *o = static_cast<const T *>(offset),
            *m = static_cast<const T *>(mask);
    T *y = static_cast<T *>(out);
    const bool k33 = q.kh == 3 && q.kw == 3;
    const size_t smem = tile_smem_bytes<T>(q.kh * q.kw);
    if (smem > 200 * 1024) return false;
    auto set_smem = [&](auto kernel) {
        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    };
    for (int n0 = 0; n0 < q.N; n0 += 65535) {
        tp.n0 = n0;
        const dim3 grid((unsigned)(tp.tiles_x * tiles_y), (unsigned)slices, (unsigned)std::min(65535, q.N - n0));
        if (fast) {
            if constexpr (sizeof(T) == 2) {
                if (k33) { set_smem(fwd_tile<T, true, 3, 3>); fwd_tile<T, true, 3, 3><<<grid, kTileThreads, smem, stream>>>(tmap, v, o, m, y, q, tp); }
                else { set_smem(fwd_tile<T, true, 0, 0>); fwd_tile<T, true, 0, 0><<<grid, kTileThreads, smem, stream>>>(tmap, v, o, m, y, q, tp); }
            }
        } else {
            if (k33) { set_smem(fwd_tile<T, false, 3, 3>); fwd_tile<T, false, 3, 3><<<grid, kTileThreads, smem, stream>>>(tmap, v, o, m, y, q, tp); }
            else { set_smem(fwd_tile<T, false, 0, 0>); fwd_tile<T, false, 0, 0><<<grid, kTileThreads, smem, stream>>>(tmap, v, o, m, y, q, tp); }
        }
    }
    *err = cudaGetLastError();
    return true;
}

// Returns true if the tiled kernel took the call (result in *err), false if the shape is not
// eligible and the caller should use the direct kernel.
bool try_launch_forward_tile(const void *value, const void *offset, const void *mask, void *out,
                             const Geom &q, int dtype, bool fast, cudaStream_t stream, cudaError_t *err) {
    const char *e = std::getenv("DCNV3_FWD");   // development knob: DCNV3_FWD=gather|mma select other kernels
    if (e && (e[0] == 'g' || e[0] == 'm')) return false;
    if ((long long)q.N * q.Ho * q.Wo == 0) return false;
    switch (dtype) {
    case 0: return launch_tile_typed<float>(value, offset, mask, out, q, dtype, false, stream, err);
    case 1: return launch_tile_typed<__half>(value, offset, mask, out, q, dtype, fast, stream, err);
    default: return launch_tile_typed<__nv_bfloat16>(value, offset, mask, out, q, dtype, fast, stream, err);
    }
}

}  // namespace dcnv3
