// dcnv3_forward_gs.cu -- DCNv3 core forward for 16-bit I/O, group_channels == 16, 3x3 / stride 1 /
// dilation 1, group count a multiple of 8: the "group-slice" kernel (default forward for these).
// group_channels == 32 (G % 4 == 0) runs the same kernel with a group as TWO 16-channel slices (template
// parameter GSH = 1): lanes 2j and 2j+1 share the group's offsets / masks; its 72-byte mask runs are not
// legal TMA boxes, so the mask box starts on the 16-byte boundary below the run (m_shift).
// group_channels == 8 (G % 8 == 0; BASELINE configs[4], group 32 at C = 256) runs it with 16-byte slices (template
// parameter CH = 8): a cell of eight groups is 128 bytes = one full bank row, a quarter-warp's eight 16-byte reads
// cover it exactly once whatever cell each lane samples, the window shrinks to 41 KB and three CTAs share an SM.
//
// Why (profiles/README.md, r1_v3 fwd_tile): the tiled forward was issue-bound at 201 instructions
// per sampled point, of which only 64 are the FHFMAs and 8 the gather loads; ~40 went into staging
// offsets / masks with per-element index arithmetic and ~25 into rotating the corner order of
// every point so that a quarter-warp of eight PIXELS reads eight different 16-byte bank slots.
//
// How: a CTA (512 threads) owns an 8x8 tile of output pixels for EIGHT groups, and the eight lanes
// of a quarter-warp are the eight GROUPS of one pixel.
//   * the value window is staged as [18][18][8 groups x 32 B]: a cell is 256 bytes = two full bank
//     rows, so the bank slot of a lane's 16-byte read is (2 g + half) mod 8 whatever cell it
//     samples -- group g reads half (g>>2)&1 first and the quarter-warp covers the eight slots
//     exactly once.  Conflict-free for ANY offsets, with no rotation of corners or weights;
//   * a pixel's offsets / masks for eight groups are 288 / 144 contiguous bytes, so the tile's
//     offsets, masks and the value window are THREE TMA box loads issued by one thread
//     (out-of-map elements zero-filled = the op's zero padding; a zero mask outside the output map);
//   * their shared-memory layout is conflict-free for the per-point 32-bit / 16-bit reads
//     ((8 px + 9 g + p) mod 32 is a bijection of the 32 lanes);
//   * the output of a warp is four runs of 256 contiguous bytes.
// Default for group_channels == 16 since the end of round 1 (template parameter KG = 4): the same kernel with FOUR
// groups per 256-thread CTA and four CTAs per SM -- 128-byte cells, a quarter-warp is two pixels x four groups and
// the odd pixel reads its second 16-byte chunk first (all eight slots once, for any offsets); the 72-byte mask
// runs are staged like those of gc == 32.  Same occupancy, but while one CTA waits for its window three others
// compute instead of one: 79 -> 75 us on cfg2 (DCNV3_GS_KG=8 selects the 512-thread form).
// The loop over the nine points is branch-free: a point whose corner block leaves the window
// (offsets beyond about +-4 px of the kernel tap) contributes zero there and is redone afterwards
// from global memory, so the compiler is free to keep several points' loads in flight.
//
// Coordinates are window-relative (the window origin is folded into the anchors; 16-bit I/O only,
// where the one-ulp difference to the reference's association is far below the I/O rounding); the
// reference's range test (dcnv3_im2col_cuda.cuh:262-263) is implied by the zero fill: a point that
// fails it has all four corners outside the map.
#include "dcnv3_common.cuh"
#include "dcnv3_launch.h"
#include "dcnv3_tma.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace dcnv3 {
namespace gs {

constexpr int kTile = 8;                       // output pixels per tile side
constexpr int kWin = 18;                       // value window side
constexpr int kGroups = 8;                     // groups per CTA
constexpr int kCh = 16;                        // channels per slice (32 bytes of 16-bit data); 8 for gc == 8
constexpr int kPix = kTile * kTile;            // 64
constexpr int kThreads = kPix * kGroups;       // 512 (256 when a CTA takes four slices, template parameter KG = 4)
constexpr int kP = 9;
__host__ __device__ constexpr int cell_bytes(int ch, int kg = kGroups) { return kg * ch * 2; }                  // 256 / 128
__host__ __device__ constexpr int win_bytes(int ch, int kg = kGroups) { return kWin * kWin * cell_bytes(ch, kg); } // 82944 / 41472
// staged offset / mask bytes per pixel for `groups` groups: 36 / 18 bytes each; a mask run that is not a multiple
// of 16 bytes (4 groups: 72) is staged from the 16-byte boundary below it (up to 8 bytes of shift): 80-byte pitch
__host__ __device__ constexpr int off_pitch(int groups) { return groups * kP * 4; }
__host__ __device__ constexpr int msk_pitch(int groups) { return (groups * kP * 2) % 16 ? 80 : groups * kP * 2; }
__host__ __device__ constexpr int smem_bytes(int ch, int kg = kGroups, int gsh = 0) {
    return win_bytes(ch, kg) + kPix * (off_pitch(kg >> gsh) + msk_pitch(kg >> gsh));
}

struct Params {
    int ox_rel, oy_rel;      // window origin relative to the tile origin
    int tiles_x;
    int gblocks;             // 16-channel slices / 8
    int n0;
    // group_channels == 32: a group is TWO 16-channel slices (lanes 2j, 2j+1 share the group's offsets / masks);
    // a CTA's 8 slices are then 4 groups
    int gsh;                 // log2(slices per group): 0 or 1
    int o_pitch, m_pitch;    // staged offset / mask bytes per pixel
};

__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

// 2-D / 4-D tensor maps over the offset and mask tensors: dims (G*K elements, Wo, Ho, N)
static bool make_rows_tensor_map(CUtensorMap *map, const void *base, int dtype, int N, int Ho, int Wo,
                                 int row_elems, int box_elems) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t es = 2;
    const CUtensorMapDataType dt = dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const cuuint64_t dims[4] = {(cuuint64_t)row_elems, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)row_elems * es, (cuuint64_t)Wo * row_elems * es,
                                   (cuuint64_t)Ho * Wo * row_elems * es};
    const cuuint32_t box[4] = {(cuuint32_t)box_elems, (cuuint32_t)kTile, (cuuint32_t)kTile, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    return fn(map, dt, 4, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Diagnostic builds only (scripts/gpu_gsdiag.sh; results are WRONG, the timing shows what a part costs):
// -DGS_DIAG=1 the gather's shared-memory loads without the arithmetic, 2 the arithmetic without the gather
#ifndef GS_DIAG
#define GS_DIAG 0
#endif

template <typename T, bool FAST, int GSH, int CH, int KG = kGroups>
__global__ void __launch_bounds__(kPix * KG, KG == 4 ? 4 : CH == 8 ? 3 : 2)
fwd_gs(const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o,
       const __grid_constant__ CUtensorMap tmap_m, const T *__restrict__ value, const T *__restrict__ offset,
       const T *__restrict__ mask, T *__restrict__ out, const Geom q, const Params tp) {
    constexpr int E = 8;
    constexpr int kCellBytes = cell_bytes(CH, KG), kWinBytes = win_bytes(CH, KG);
    constexpr int kOPitch = off_pitch(KG >> GSH), kMPitch = msk_pitch(KG >> GSH), kOffBytes = kPix * kOPitch;
    constexpr bool PADM = ((KG >> GSH) * kP * 2) % 16 != 0;
    constexpr bool TWO = CH == 16;              // a lane owns two 16-byte chunks of a cell (one when gc == 8)
    static_assert(CH == 16 || (CH == 8 && GSH == 0), "slices of 16 or 8 channels");
    static_assert(KG == 8 || (KG == 4 && GSH == 0 && CH == 16), "eight slices per CTA, or four 16-channel groups");
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    unsigned char *win = smem;
    const uint32_t s_off = smem_u32(smem + kWinBytes);               // [64 px][8 g][9] (dx, dy) pairs
    const uint32_t s_msk = smem_u32(smem + kWinBytes + kOffBytes);   // [64 px][8 g][9]

    const int tid = threadIdx.x;
    const int g = tid & (KG - 1), pix = tid / KG, px = pix & 7, py = pix >> 3;
    const int tile_x = blockIdx.x % tp.tiles_x, tile_y = blockIdx.x / tp.tiles_x;
    const int g0 = blockIdx.y * KG;                            // first 16-channel slice of the CTA
    const int gr = g >> GSH;                                   // this lane's group inside the CTA's block
    const int G0 = blockIdx.y * (KG >> GSH);                   // first group of the CTA
    // the mask box starts on the 16-byte boundary below the block's run (72-byte runs when gc == 32)
    const int m_shift = PADM ? (G0 * kP * 2) & 15 : 0;
    const int n = tp.n0 + blockIdx.z;
    const int wo0 = tile_x * kTile, ho0 = tile_y * kTile;
    const int ox = wo0 + tp.ox_rel, oy = ho0 + tp.oy_rel;

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, kWinBytes + kPix * (kOPitch + kMPitch));
        tma_load_4d(smem + kWinBytes, &tmap_o, &bar, G0 * kP * 2, wo0, ho0, n);
        tma_load_4d(smem + kWinBytes + kOffBytes, &tmap_m, &bar, (G0 * kP * 2 - m_shift) >> 1, wo0, ho0, n);
        tma_load_4d(win, &tmap_v, &bar, g0 * CH, ox, oy, n);
    }

    const int wo = wo0 + px, ho = ho0 + py;
    const bool live = wo < q.Wo && ho < q.Ho;
    // 16-byte chunk this lane reads FIRST (KG == 4: a quarter-warp is two pixels x four groups of 128-byte cells)
    const int half = TWO ? (KG == 8 ? (g >> 2) & 1 : pix & 1) : 0;
    const float base_w = axis_base(wo, 3, 1, q.pw, 1, q.sigma);
    const float base_h = axis_base(ho, 3, 1, q.ph, 1, q.sigma);
    const float bw = base_w - (float)ox, bh = base_h - (float)oy;     // window-relative anchors
    const uint32_t win_addr = smem_u32(win) + g * (CH * 2) + half * 16;
    const uint32_t my_off = s_off + pix * kOPitch + gr * (kP * 4);
    const uint32_t my_msk = s_msk + pix * kMPitch + m_shift + gr * (kP * 2);

    float acc_a[E], acc_b[E];   // acc_a: channels of chunk `half`, acc_b: the other chunk
#pragma unroll
    for (int v = 0; v < E; ++v) acc_a[v] = acc_b[v] = 0.f;

    mbar_wait(&bar, 0);
    if (!live) return;

    unsigned miss = 0;
#pragma unroll
    for (int p = 0; p < kP; ++p) {
        const int i = p / 3, jj = p % 3;
        uint32_t o2;
        uint16_t m16;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(o2) : "r"(my_off + p * 4));
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(m16) : "r"(my_msk + p * 2));
        const float2 d = unpack2(o2, T());
        float m = f32_of(m16, T());
        float u = bw + ((float)i + d.x) * q.sigma;
        float v = bh + ((float)jj + d.y) * q.sigma;
        // 0 <= x < kWin-1 on the bit pattern: negative values and NaN compare as large unsigned
        const bool hit = __float_as_uint(u) < __float_as_uint((float)(kWin - 1)) &&
                         __float_as_uint(v) < __float_as_uint((float)(kWin - 1));
        if (!hit) {
            miss |= 1u << p;
            u = 0.f; v = 0.f; m = 0.f;
        }
        const float fu = floorf(u), fv = floorf(v);
        const float lw = u - fu, lh = v - fv, hw = 1.f - lw, hh = 1.f - lh;
        const uint32_t tl = win_addr + (uint32_t)((int)fv * kWin + (int)fu) * kCellBytes;
        const float hm = hh * m, lm = lh * m;
        const float w[4] = {hm * hw, hm * lw, lm * hw, lm * lw};
        const uint32_t o[4] = {0u, (uint32_t)kCellBytes, (uint32_t)(kWin * kCellBytes), (uint32_t)((kWin + 1) * kCellBytes)};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const Weight<T, FAST> wt(w[t]);
#if GS_DIAG == 1   // gather only: the eight 128-bit shared-memory loads of a point, consumed by three XORs each (WRONG results)
            const uint4 qa = lds128(tl + o[t]);
            acc_a[t] = __uint_as_float(((__float_as_uint(acc_a[t]) ^ qa.x ^ qa.y ^ qa.z ^ qa.w) & 0x007fffffu) | 0x3f800000u);   // (never NaN)
            if (TWO) {
                const uint4 qb = lds128((tl ^ 16u) + o[t]);
                acc_b[t] = __uint_as_float(((__float_as_uint(acc_b[t]) ^ qb.x ^ qb.y ^ qb.z ^ qb.w) & 0x007fffffu) | 0x3f800000u);
            }
#elif GS_DIAG == 2   // arithmetic only: every point reads the window's first cell (no gather traffic; WRONG results)
            axpy<T, FAST>(acc_a, lds128(win_addr), wt);
            if (TWO) axpy<T, FAST>(acc_b, lds128(win_addr ^ 16u), wt);
            if (t == 3) acc_a[0] += __uint_as_float(tl);   // (keeps the address arithmetic alive)
#else
            axpy<T, FAST>(acc_a, lds128(tl + o[t]), wt);
            if (TWO) axpy<T, FAST>(acc_b, lds128((tl ^ 16u) + o[t]), wt);
#endif
        }
    }

    if (miss) {
        // ---- points whose corner block leaves the window: clamped global reads
        const int C = q.G * q.gc, row_stride = q.W * C;
        const T *img = value + (size_t)n * q.H * row_stride + (g0 + g) * CH;
        const size_t e0 = ((((size_t)n * q.Ho + ho) * q.Wo + wo) * q.G + (G0 + gr)) * kP;
        for (int p = 0; p < kP; ++p) {
            if (!((miss >> p) & 1u)) continue;
            const int i = p / 3, jj = p % 3;
            const float2 d = load_pair(offset + (e0 + p) * 2);
            const float m = to_f32(__ldg(mask + e0 + p));
            const float loc_w = base_w + ((float)i + d.x) * q.sigma;
            const float loc_h = base_h + ((float)jj + d.y) * q.sigma;
            const ClampedTap ct = make_clamped_tap(loc_h, loc_w, q.H, q.W);
            if (!ct.inside) continue;
            const T *r_lo = img + ct.row_lo * row_stride, *r_hi = img + ct.row_hi * row_stride;
            const int c_lo = ct.col_lo * C, c_hi = ct.col_hi * C;
            const int ea = half * E, eb = (half ^ 1) * E;
            const float fy_lo = ct.hh * ct.top * m, fy_hi = ct.lh * ct.bot * m;
            const float fx_lo = ct.hw * ct.lef, fx_hi = ct.lw * ct.rig;
            const T *corner[4] = {r_lo + c_lo, r_lo + c_hi, r_hi + c_lo, r_hi + c_hi};
            const float wc[4] = {fy_lo * fx_lo, fy_lo * fx_hi, fy_hi * fx_lo, fy_hi * fx_hi};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const Weight<T, FAST> wt(wc[t]);
                axpy<T, FAST>(acc_a, __ldg(reinterpret_cast<const uint4 *>(corner[t] + ea)), wt);
                if (TWO) axpy<T, FAST>(acc_b, __ldg(reinterpret_cast<const uint4 *>(corner[t] + eb)), wt);
            }
        }
    }

    T *dst = out + (((size_t)n * q.Ho + ho) * q.Wo + wo) * (size_t)(q.G * q.gc) + (g0 + g) * CH;
    *reinterpret_cast<uint4 *>(dst + half * E) = pack<T>(acc_a);
    if (TWO) *reinterpret_cast<uint4 *>(dst + (half ^ 1) * E) = pack<T>(acc_b);
}

template <typename T>
static bool launch_typed(const void *value, const void *offset, const void *mask, void *out, const Geom &q,
                         int dtype, bool fast, cudaStream_t stream, cudaError_t *err) {
    if (!(((q.gc == kCh || q.gc == 8) && q.G % kGroups == 0) || (q.gc == 2 * kCh && q.G % (kGroups / 2) == 0)) || q.kh != 3 ||
        q.kw != 3 || q.sh != 1 || q.sw != 1 || q.dh != 1 || q.dw != 1)
        return false;
    if (((uintptr_t)value | (uintptr_t)out | (uintptr_t)offset | (uintptr_t)mask) % 16) return false;
    const int gsh = q.gc == 2 * kCh ? 1 : 0;
    const int ch = q.gc == 8 ? 8 : kCh;                                // channels per slice
    // gc == 16: four groups per 256-thread CTA, four CTAs per SM (DCNV3_GS_KG=8 selects the 512-thread form)
    const char *ekg = std::getenv("DCNV3_GS_KG");
    const int kg = (q.gc == kCh && !(ekg && ekg[0] == '8')) ? 4 : kGroups;
    const int grp = kg >> gsh;                                         // groups per CTA
    // the tile's nominal tap span must leave at least 2 pixels of offset slack on each side
    const float span = (kTile - 1) + 2 * q.sigma;
    if (!(q.sigma > 0.f) || span + 4 > kWin - 2) return false;
    const int C = q.G * q.gc;
    CUtensorMap tmap_v, tmap_o, tmap_m;
    if (!make_nhwc_tensor_map(&tmap_v, value, dtype, q.N, q.H, q.W, C, kg * ch, kWin, kWin)) return false;
    Params tp;
    tp.gsh = gsh;
    tp.o_pitch = off_pitch(grp);                                // 288 / 144 bytes
    tp.m_pitch = msk_pitch(grp);                                // 144 bytes, or the 72-byte run + up to 8 bytes of shift
    if (!make_rows_tensor_map(&tmap_o, offset, dtype, q.N, q.Ho, q.Wo, q.G * kP * 2, tp.o_pitch / 2)) return false;
    if (!make_rows_tensor_map(&tmap_m, mask, dtype, q.N, q.Ho, q.Wo, q.G * kP, tp.m_pitch / 2)) return false;
    const float a_w = (float)(1 - q.pw) - q.sigma, a_h = (float)(1 - q.ph) - q.sigma;
    tp.ox_rel = (int)std::floor(a_w + 0.5f * span - 0.5f * (kWin - 2));
    tp.oy_rel = (int)std::floor(a_h + 0.5f * span - 0.5f * (kWin - 2));
    tp.tiles_x = (q.Wo + kTile - 1) / kTile;
    tp.gblocks = q.G / grp;
    const int tiles_y = (q.Ho + kTile - 1) / kTile;
    if (tp.gblocks > 65535) return false;
    const T *v = static_cast<const T *>(value), *o = static_cast<const T *>(offset), *m = static_cast<const T *>(mask);
    T *y = static_cast<T *>(out);
    auto kern = kg == 4 ? (fast ? fwd_gs<T, true, 0, 16, 4> : fwd_gs<T, false, 0, 16, 4>)
                : ch == 8 ? (fast ? fwd_gs<T, true, 0, 8> : fwd_gs<T, false, 0, 8>)
                : gsh   ? (fast ? fwd_gs<T, true, 1, 16> : fwd_gs<T, false, 1, 16>)
                        : (fast ? fwd_gs<T, true, 0, 16> : fwd_gs<T, false, 0, 16>);
    const int kSmemBytes = smem_bytes(ch, kg, gsh);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    for (int n0 = 0; n0 < q.N; n0 += 65535) {
        tp.n0 = n0;
        const dim3 grid((unsigned)(tp.tiles_x * tiles_y), (unsigned)tp.gblocks, (unsigned)std::min(65535, q.N - n0));
        kern<<<grid, kPix * kg, kSmemBytes, stream>>>(tmap_v, tmap_o, tmap_m, v, o, m, y, q, tp);
    }
    *err = cudaGetLastError();
    return true;
}

}  // namespace gs

// Returns true if the group-slice kernel took the call (result in *err), false if the shape is not
// eligible and the caller should try the tiled kernel.
bool try_launch_forward_gs(const void *value, const void *offset, const void *mask, void *out, const Geom &q,
                           int dtype, bool fast, cudaStream_t stream, cudaError_t *err) {
    const char *e = std::getenv("DCNV3_FWD");   // development knob: any value selects an older kernel
    if (e && e[0] && !(e[0] == 'g' && e[1] == 's')) return false;
    if ((long long)q.N * q.Ho * q.Wo == 0) return false;
    if (dtype == 1) return gs::launch_typed<__half>(value, offset, mask, out, q, dtype, fast, stream, err);
    if (dtype == 2) return gs::launch_typed<__nv_bfloat16>(value, offset, mask, out, q, dtype, fast, stream, err);
    return false;
}

}  // namespace dcnv3
