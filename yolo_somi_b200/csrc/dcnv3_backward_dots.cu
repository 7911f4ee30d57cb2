// dcnv3_backward_dots.cu -- first half of the split DCNv3 backward for 16-bit I/O, group_channels == 16
// (G % 8 == 0) or 32 (G % 4 == 0: a group is two 16-channel slices, template parameter GSH = 1 -- the two lanes
// of a group share its offsets / masks and add their partial sums with one shuffle; the 72-byte mask runs are
// not legal TMA boxes, so the mask box starts on the 16-byte boundary below the run and grad_mask leaves by
// plain 32-bit stores) or 8 (G % 8 == 0: 16-byte slices, template parameter CH = 8 -- a 128-byte cell, three CTAs
// per SM), 3x3 / stride 1 / dilation 1: grad_offset and grad_mask (the per-point channel sums
// of dcnv3_im2col_cuda.cuh:106-146, 278-370) with the forward's group-slice layout.
//
// Why split (profiles/README.md, r1_v4): the fused strip backward holds a 12 KB coefficient tile
// per warp, so only 8 warps fit an SM and every warp issues one instruction per ~4.4 cycles; its
// gather also needs the per-point corner rotation.  The channel sums need no coefficient tile at
// all, so they run here exactly like the forward (dcnv3_forward_gs.cu): a CTA of 512 threads owns
// 8x8 pixels x 8 groups, a quarter-warp is the 8 groups of one pixel (bank slot fixed by the
// group: conflict-free gather without rotation), window / offsets / masks arrive as three TMA
// boxes, 32 warps per SM.  Per point: four corner dot products grad_out . value (exact FHFMA
// products, fp32 sums) -> grad_mask, grad_offset.  Results overwrite the staged offsets / masks in
// shared memory and leave as two TMA stores (a pixel's 8 groups are 288 / 144 contiguous bytes).
// grad_value is produced by dcnv3_backward_vstrip.cu.
#include "dcnv3_common.cuh"
#include "dcnv3_launch.h"
#include "dcnv3_tma.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace dcnv3 {
namespace bdots {

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
    // group_channels == 32: a group is TWO 16-channel slices (lanes 2j, 2j+1 share the group's offsets / masks
    // and add their partial channel sums); a CTA's 8 slices are then 4 groups
    int gsh;                 // log2(slices per group): 0 or 1
    int o_pitch, m_pitch;    // staged offset / mask bytes per pixel
    // the fp32 plane the value kernel reduces into is zeroed HERE, a slice per CTA, while the CTA waits for its
    // window (no separate memset, no side stream): 16-byte units per CTA / in all
    unsigned zero_per_cta;
    unsigned long long zero_total;
};

__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, const void *src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(src)) : "memory");
}

// 4-D tensor maps over the offset / mask (and their gradient) tensors: dims (G*K elements, Wo, Ho, N)
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

template <typename T, int GSH, int CH, int KG = kGroups>
__global__ void __launch_bounds__(kPix * KG, KG == 4 ? 4 : CH == 8 ? 3 : 2)
bwd_dots(const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o,
         const __grid_constant__ CUtensorMap tmap_m, const __grid_constant__ CUtensorMap tmap_go,
         const __grid_constant__ CUtensorMap tmap_gm, const T *__restrict__ value, const T *__restrict__ offset,
         const T *__restrict__ mask, const T *__restrict__ grad_out, T *__restrict__ grad_mask_out,
         float4 *__restrict__ zero_dst, const Geom q, const Params tp) {
    constexpr int E = 8;
    constexpr int kCellBytes = cell_bytes(CH, KG), kWinBytes = win_bytes(CH, KG);
    constexpr int kOPitch = off_pitch(KG >> GSH), kMPitch = msk_pitch(KG >> GSH), kOffBytes = kPix * kOPitch;
    constexpr bool PADM = ((KG >> GSH) * kP * 2) % 16 != 0;      // 72-byte mask runs: no legal TMA box
    constexpr int kThreadsK = kPix * KG;
    constexpr bool TWO = CH == 16;              // a lane owns two 16-byte chunks of a cell (one when gc == 8)
    static_assert(CH == 16 || (CH == 8 && GSH == 0), "slices of 16 or 8 channels");
    static_assert(KG == 8 || (KG == 4 && GSH == 0 && CH == 16), "eight slices per CTA, or four 16-channel groups");
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    unsigned char *win = smem;
    unsigned char *off_tile = smem + kWinBytes;                      // [64 px][8 g][9] (dx, dy) pairs -> results
    unsigned char *msk_tile = smem + kWinBytes + kOffBytes;          // [64 px][8 g][9] masks -> results
    const uint32_t s_off = smem_u32(off_tile), s_msk = smem_u32(msk_tile);

    const int tid = threadIdx.x;
    const int g = tid & (KG - 1), pix = tid / KG, px = pix & 7, py = pix >> 3;
    const int tile_x = blockIdx.x % tp.tiles_x, tile_y = blockIdx.x / tp.tiles_x;
    const int g0 = blockIdx.y * KG;                            // first 16-channel slice of the CTA
    const int gr = g >> GSH;                                   // this lane's group inside the CTA's block
    const int G0 = blockIdx.y * (KG >> GSH);                   // first group of the CTA
    // the mask box starts on the 16-byte boundary below the block's run (72-byte runs when gc == 32)
    const int m_shift = PADM ? (G0 * kP * 2) & 15 : 0;
    const int C = q.G * q.gc;
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
        tma_load_4d(off_tile, &tmap_o, &bar, G0 * kP * 2, wo0, ho0, n);
        tma_load_4d(msk_tile, &tmap_m, &bar, (G0 * kP * 2 - m_shift) >> 1, wo0, ho0, n);
        tma_load_4d(win, &tmap_v, &bar, g0 * CH, ox, oy, n);
    }

    const int wo = wo0 + px, ho = ho0 + py;
    const bool live = wo < q.Wo && ho < q.Ho;
    // 16-byte chunk this lane reads FIRST (KG == 4: a quarter-warp is two pixels x four groups of 128-byte cells)
    const int half = TWO ? (KG == 8 ? (g >> 2) & 1 : pix & 1) : 0;
    const float base_w = axis_base(wo, 3, 1, q.pw, 1, q.sigma);
    const float base_h = axis_base(ho, 3, 1, q.ph, 1, q.sigma);
    const float bw = base_w - (float)ox, bh = base_h - (float)oy;     // window-relative anchors
    // The reference's range test (dcnv3_im2col_cuda.cuh:334: loc > -1 && loc < extent) is implied by the window's zero
    // fill everywhere EXCEPT at loc == -1 exactly (row / column -1 reads zeros, so grad_mask is 0, but the derivative
    // across that row / column is not): window-relative -1 on either axis zeroes the point's offset gradient.
    const float u_m1 = -1.f - (float)ox, v_m1 = -1.f - (float)oy;
    const uint32_t win_addr = smem_u32(win) + g * (CH * 2) + half * 16;
    const uint32_t my_off = s_off + pix * kOPitch + gr * (kP * 4);
    const uint32_t my_msk = s_msk + pix * kMPitch + m_shift + gr * (kP * 2);

    // upstream gradient of this (pixel, group): chunk `half` and the other chunk (a warp reads four
    // runs of 256 contiguous bytes)
    uint4 gq_a = make_uint4(0u, 0u, 0u, 0u), gq_b = gq_a;
    if (live) {
        const T *gp = grad_out + (((size_t)n * q.Ho + ho) * q.Wo + wo) * (size_t)C + (g0 + g) * CH;
        gq_a = __ldg(reinterpret_cast<const uint4 *>(gp + half * E));
        if (TWO) gq_b = __ldg(reinterpret_cast<const uint4 *>(gp + (half ^ 1) * E));
    }

    // the value kernel (launched with programmatic stream serialisation) may start its prologue and input loads as
    // soon as every CTA of this grid has started; it waits for this grid's completion before it touches the plane
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (zero_dst) {   // this CTA's slice of the fp32 plane (independent of everything else the kernel does)
        const unsigned long long cta = blockIdx.x + (unsigned long long)gridDim.x * (blockIdx.y + (unsigned long long)gridDim.y * blockIdx.z);
        const unsigned long long b = cta * tp.zero_per_cta;
        const unsigned long long e = b + tp.zero_per_cta < tp.zero_total ? b + tp.zero_per_cta : tp.zero_total;
        for (unsigned long long i = b + tid; i < e; i += kThreadsK) zero_dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    mbar_wait(&bar, 0);

    if (live) {
        unsigned miss = 0;
#pragma unroll
        for (int p = 0; p < kP; ++p) {
            const int i = p / 3, jj = p % 3;
            uint32_t o2;
            uint16_t m16;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(o2) : "r"(my_off + p * 4));
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(m16) : "r"(my_msk + p * 2));
            const float2 d = unpack2(o2, T());
            const float m = f32_of(m16, T());
            float u = bw + ((float)i + d.x) * q.sigma;
            float v = bh + ((float)jj + d.y) * q.sigma;
            // 0 <= x < kWin-1 on the bit pattern: negative values and NaN compare as large unsigned
            const bool hit = __float_as_uint(u) < __float_as_uint((float)(kWin - 1)) &&
                             __float_as_uint(v) < __float_as_uint((float)(kWin - 1));
            if (!hit) {
                miss |= 1u << p;
                u = 0.f; v = 0.f;
            }
            const float fu = floorf(u), fv = floorf(v);
            const float lw = u - fu, lh = v - fv, hw = 1.f - lw, hh = 1.f - lh;
            const uint32_t tl = win_addr + (uint32_t)((int)fv * kWin + (int)fu) * kCellBytes;
            const uint32_t o[4] = {0u, (uint32_t)kCellBytes, (uint32_t)(kWin * kCellBytes), (uint32_t)((kWin + 1) * kCellBytes)};
            float dr[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                dr[t] = dot<T>(gq_a, lds128(tl + o[t]), 0.f);
                if (TWO) dr[t] += dot<T>(gq_b, lds128((tl ^ 16u) + o[t]), 0.f);
            }
            if (GSH) {   // the group's other 16 channels are the neighbouring lane's
                const unsigned am = __activemask();
#pragma unroll
                for (int t = 0; t < 4; ++t) dr[t] += __shfl_xor_sync(am, dr[t], 1);
            }
            const float gm = hh * (hw * dr[0] + lw * dr[1]) + lh * (hw * dr[2] + lw * dr[3]);
            const float mg = (u == u_m1 || v == v_m1) ? 0.f : m;
            const float gx = mg * (hh * (dr[1] - dr[0]) + lh * (dr[3] - dr[2]));
            const float gy = mg * (hw * (dr[2] - dr[0]) + lw * (dr[3] - dr[1]));
            // results overwrite the staged inputs of this point (a miss is redone below)
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(my_off + p * 4), "r"(pack2(q.sigma * gx, q.sigma * gy, T())) : "memory");
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(my_msk + p * 2), "h"(bits16(gm, T())) : "memory");
        }

        if (miss) {
            // ---- points whose corner block leaves the window: clamped global reads
            const int row_stride = q.W * C;
            const T *img = value + (size_t)n * q.H * row_stride + (g0 + g) * CH;
            // (the staged inputs of a missed point were overwritten above: re-read them from the tensors)
            for (int p = 0; p < kP; ++p) {
                if (!((miss >> p) & 1u)) continue;
                const int i = p / 3, jj = p % 3;
                const size_t e0 = ((((size_t)n * q.Ho + ho) * q.Wo + wo) * q.G + (G0 + gr)) * kP + p;
                const float2 d = load_pair(offset + e0 * 2);
                const float m = to_f32(__ldg(mask + e0));
                const float loc_w = base_w + ((float)i + d.x) * q.sigma;
                const float loc_h = base_h + ((float)jj + d.y) * q.sigma;
                const ClampedTap ct = make_clamped_tap(loc_h, loc_w, q.H, q.W);
                float gm = 0.f, gx = 0.f, gy = 0.f;
                if (ct.inside) {
                    const int r_lo = ct.row_lo * row_stride, r_hi = ct.row_hi * row_stride;
                    const int c_lo = ct.col_lo * C, c_hi = ct.col_hi * C;
                    const int at[4] = {r_lo + c_lo, r_lo + c_hi, r_hi + c_lo, r_hi + c_hi};
                    const int ea = half * E, eb = (half ^ 1) * E;
                    float dk[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        dk[t] = dot<T>(gq_a, __ldg(reinterpret_cast<const uint4 *>(img + at[t] + ea)), 0.f);
                        if (TWO) dk[t] += dot<T>(gq_b, __ldg(reinterpret_cast<const uint4 *>(img + at[t] + eb)), 0.f);
                    }
                    if (GSH) {   // (both lanes of a group take the same path: same coordinates)
                        const unsigned am = __activemask();
#pragma unroll
                        for (int t = 0; t < 4; ++t) dk[t] += __shfl_xor_sync(am, dk[t], 1);
                    }
                    const float fy_lo = ct.hh * ct.top, fy_hi = ct.lh * ct.bot;
                    const float fx_lo = ct.hw * ct.lef, fx_hi = ct.lw * ct.rig;
                    gm = fy_lo * (fx_lo * dk[0] + fx_hi * dk[1]) + fy_hi * (fx_lo * dk[2] + fx_hi * dk[3]);
                    gx = m * (fy_lo * (ct.rig * dk[1] - ct.lef * dk[0]) + fy_hi * (ct.rig * dk[3] - ct.lef * dk[2]));
                    gy = m * (fx_lo * (ct.bot * dk[2] - ct.top * dk[0]) + fx_hi * (ct.bot * dk[3] - ct.top * dk[1]));
                }
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(my_off + p * 4), "r"(pack2(q.sigma * gx, q.sigma * gy, T())) : "memory");
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(my_msk + p * 2), "h"(bits16(gm, T())) : "memory");
            }
        }
    }
    // ---- the tile's results leave as two TMA stores (pixels beyond the map are clipped)
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        tma_store_4d(&tmap_go, off_tile, G0 * kP * 2, wo0, ho0, n);
        if (!PADM) tma_store_4d(&tmap_gm, msk_tile, G0 * kP, wo0, ho0, n);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (PADM) {
        // four groups per CTA (gc == 32, or KG == 4): the block's mask run is 72 bytes at a 72-byte pitch -- not a
        // legal TMA box; plain 64-bit stores (a run starts on an 8-byte boundary: 72 G0 bytes into a row of 18 G)
        uint2 *gmw = reinterpret_cast<uint2 *>(grad_mask_out);
#pragma unroll
        for (int it = 0; it < (kPix * 9 + kThreadsK - 1) / kThreadsK; ++it) {
            const int idx = tid + it * kThreadsK;
            const int pxl = idx / 9, w = idx - pxl * 9;
            const int xo = wo0 + (pxl & 7), yo = ho0 + (pxl >> 3);
            if (idx < kPix * 9 && xo < q.Wo && yo < q.Ho) {
                uint2 vword;
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(vword.x), "=r"(vword.y) : "r"(s_msk + pxl * kMPitch + m_shift + w * 8));
                gmw[(((((size_t)n * q.Ho + yo) * q.Wo + xo) * q.G + G0) * kP) / 4 + w] = vword;
            }
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory outlives the reads
}

template <typename T>
static bool launch_typed(const void *value, const void *offset, const void *mask, const void *grad_out,
                         void *grad_offset, void *grad_mask, const Geom &q, int dtype, cudaStream_t stream,
                         cudaError_t *err, float *zero_plane, size_t zero_bytes) {
    if (!(((q.gc == kCh || q.gc == 8) && q.G % kGroups == 0) || (q.gc == 2 * kCh && q.G % (kGroups / 2) == 0)) || q.kh != 3 ||
        q.kw != 3 || q.sh != 1 || q.sw != 1 || q.dh != 1 || q.dw != 1)
        return false;
    const int gsh = q.gc == 2 * kCh ? 1 : 0;
    const int ch = q.gc == 8 ? 8 : kCh;                                // channels per slice
    // gc == 16: four groups per 256-thread CTA, four CTAs per SM, as in the forward (DCNV3_GS_KG=8 selects the
    // 512-thread form): 99 -> 95 us on cfg2 once the grad_mask copy-out became an unrolled loop of 64-bit stores
    const char *ekg = std::getenv("DCNV3_GS_KG");
    const int kg = (q.gc == kCh && !(ekg && ekg[0] == '8')) ? 4 : kGroups;
    const int grp = kg >> gsh;                                         // groups per CTA
    if (((uintptr_t)value | (uintptr_t)grad_out | (uintptr_t)offset | (uintptr_t)mask | (uintptr_t)grad_offset |
         (uintptr_t)grad_mask) % 16)
        return false;
    const float span = (kTile - 1) + 2 * q.sigma;
    if (!(q.sigma > 0.f) || span + 4 > kWin - 2) return false;
    const int C = q.G * q.gc;
    CUtensorMap tv, to, tm, tgo, tgm;
    if (!make_nhwc_tensor_map(&tv, value, dtype, q.N, q.H, q.W, C, kg * ch, kWin, kWin)) return false;
    Params tp;
    tp.gsh = gsh;
    tp.o_pitch = off_pitch(grp);                                // 288 / 144 bytes
    tp.m_pitch = msk_pitch(grp);                                // 144 bytes, or the 72-byte run + up to 8 bytes of shift
    if (!make_rows_tensor_map(&to, offset, dtype, q.N, q.Ho, q.Wo, q.G * kP * 2, tp.o_pitch / 2)) return false;
    if (!make_rows_tensor_map(&tm, mask, dtype, q.N, q.Ho, q.Wo, q.G * kP, tp.m_pitch / 2)) return false;
    if (!make_rows_tensor_map(&tgo, grad_offset, dtype, q.N, q.Ho, q.Wo, q.G * kP * 2, tp.o_pitch / 2)) return false;
    if (!make_rows_tensor_map(&tgm, grad_mask, dtype, q.N, q.Ho, q.Wo, q.G * kP, tp.m_pitch / 2)) return false;
    const float a_w = (float)(1 - q.pw) - q.sigma, a_h = (float)(1 - q.ph) - q.sigma;
    tp.ox_rel = (int)std::floor(a_w + 0.5f * span - 0.5f * (kWin - 2));
    tp.oy_rel = (int)std::floor(a_h + 0.5f * span - 0.5f * (kWin - 2));
    tp.tiles_x = (q.Wo + kTile - 1) / kTile;
    tp.gblocks = q.G / grp;
    const int tiles_y = (q.Ho + kTile - 1) / kTile;
    if (tp.gblocks > 65535) return false;
    auto kern = kg == 4 ? bwd_dots<T, 0, 16, 4> : ch == 8 ? bwd_dots<T, 0, 8> : gsh ? bwd_dots<T, 1, 16> : bwd_dots<T, 0, 16>;
    const int kSmemBytes = smem_bytes(ch, kg, gsh);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    for (int n0 = 0; n0 < q.N; n0 += 65535) {
        tp.n0 = n0;
        const dim3 grid((unsigned)(tp.tiles_x * tiles_y), (unsigned)tp.gblocks, (unsigned)std::min(65535, q.N - n0));
        // the first launch zeroes the whole plane, a slice per CTA
        const bool zero = n0 == 0 && zero_plane && zero_bytes;
        const unsigned long long ctas = (unsigned long long)grid.x * grid.y * grid.z;
        tp.zero_total = zero ? zero_bytes / 16 : 0;
        tp.zero_per_cta = zero ? (unsigned)((tp.zero_total + ctas - 1) / ctas) : 0;
        kern<<<grid, kPix * kg, kSmemBytes, stream>>>(tv, to, tm, tgo, tgm, static_cast<const T *>(value),
                                                     static_cast<const T *>(offset), static_cast<const T *>(mask),
                                                     static_cast<const T *>(grad_out), static_cast<T *>(grad_mask),
                                                     zero ? reinterpret_cast<float4 *>(zero_plane) : nullptr, q, tp);
    }
    *err = cudaGetLastError();
    return true;
}

}  // namespace bdots

// grad_offset / grad_mask only.  Returns false if the shape is not eligible.
bool try_launch_backward_dots(const void *value, const void *offset, const void *mask, const void *grad_out,
                              void *grad_offset, void *grad_mask, const Geom &q, int dtype, cudaStream_t stream,
                              cudaError_t *err, float *zero_plane, size_t zero_bytes) {
    if ((long long)q.N * q.Ho * q.Wo == 0) return false;
    if (zero_bytes % 16 || (uintptr_t)zero_plane % 16) return false;
    if (dtype == 1) return bdots::launch_typed<__half>(value, offset, mask, grad_out, grad_offset, grad_mask, q, dtype, stream, err, zero_plane, zero_bytes);
    if (dtype == 2) return bdots::launch_typed<__nv_bfloat16>(value, offset, mask, grad_out, grad_offset, grad_mask, q, dtype, stream, err, zero_plane, zero_bytes);
    return false;
}

}  // namespace dcnv3
