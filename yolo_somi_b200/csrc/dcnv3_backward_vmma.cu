// dcnv3_backward_vmma.cu -- grad_value of the DCNv3 core backward for 16-bit I/O, group_channels 8, 16 or 32,
// 3x3 / stride 1 / dilation 1, as a tcgen05 product with the accumulators in TENSOR MEMORY.
// (group_channels == 8: N = 8 is not a legal UMMA shape at M = 128, so the B operand is the 16-channel run of the
// group PAIR (g & ~1, g | 1) -- the same two TMA boxes as for 16 channels -- and the drain keeps the group's own
// eight accumulator columns; the product is not what bounds the kernel.)
// Second half of the split backward (grad_offset / grad_mask: dcnv3_backward_dots.cu).
//
// What it computes (reference dcnv3_im2col_cuda.cuh:82-147, col2im bilinear):
//     grad_value[cell, c] += sum over (pixel, point, corner) hitting `cell` of  w_corner * m * grad_out[pixel, c]
// i.e. for one group the sparse product  D[cells x 16 ch] = A[cells x pixels] . G[pixels x 16 ch],  A holding
// the 36 bilinear-times-mask coefficients of every pixel.
//
// Why tensor memory (profiles/README.md, r1_v4): the register-accumulator strip kernel keeps 96
// accumulator registers per thread (252 in all, 8 warps per SM), moves the A tile through ldmatrix and
// shifts the accumulators down the strip with 64 moves per step.  Here
//   * a CTA of four warps walks ONE 8-pixel-wide strip down the map, eight rows per step, four CTAs per SM;
//     a thread of warps 0-1 is one pixel of the step's 8 x 8 patch and writes that pixel's 36 coefficients
//     into ITS column of the A tile (16-bit read-modify-writes, thread-exclusive) -- 256 band cells
//     (16 x 16: taps +- 3 px) x 64 pixels, K-major with the 128-byte swizzle the UMMA descriptors expect;
//     grad_out of the patch is the B operand exactly as two TMA boxes deliver it ([8-channel half][64 px]
//     [16 B] = the MN-major canonical layout without swizzle), nothing is transposed;
//   * one thread issues tcgen05.mma.cta_group::1.kind::f16 (M 128 cells, N 16 channels, K 16 pixels):
//     the band's upper 8 rows accumulate into the TMEM block that was the LOWER block of the previous
//     step, the lower 8 rows start a fresh block (accumulate flag off) -- "sliding the accumulator down
//     the strip" is a swap of two TMEM column offsets, nothing is moved or zeroed;
//   * the finished upper block leaves through tcgen05.ld by all four warps (thread <-> cell, 16 fp32
//     channels) as four 128-bit vector reductions per thread into the fp32 plane (lane pairs swap halves:
//     whole sectors); the A tile is re-zeroed by a 32 KB bulk copy from an L2-resident zero page (async
//     proxy, no LSU instructions) that lands while the block is drained and the next inputs are read.
// One CTA barrier and one mbarrier wait per step; offsets / masks / grad_out of the next step arrive as
// four TMA boxes (8 x 8 pixels x this group's 36 / 18 / 2 x 16-byte run; the offset / mask boxes start on
// the 16-byte boundary below the run and are 48 / 32 bytes wide, the run's start inside them is a
// warp-uniform shift) into the other of two stages while the current step is built; pixels beyond the
// output map read zeros (TMA fill), so ragged tiles need no guards.
//
// Measured on cfg2 bf16 (profiles/README.md, r1_v5): 158 us against 230 us for the HMMA value-only strip
// kernel; variants tried on the way: two strips per CTA / 2 CTAs per SM with cp.async staging 207 us, with
// TMA staging 173 us, two threads per pixel (row-parity ownership) 160 us.
//
// A point whose corner block leaves the band (|offset| beyond ~3 px) sends its four coefficient x
// grad_out rows straight to the plane (rare for trained offsets; correct for any).
#include "dcnv3_common.cuh"
#include "dcnv3_launch.h"
#include "dcnv3_tma.cuh"
#include "dcnv3_strip_io.cuh"
#include "dcnv3_tc.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>

// Diagnostic builds only (scripts/gpu_vdiag.sh; results are WRONG, the timing shows what a phase costs):
// -DVMMA_DIAG=1 no scatter into the A tile, 2 no zero refill, 3 no reductions in the drain, 4 no tcgen05.mma
#ifndef VMMA_DIAG
#define VMMA_DIAG 0
#endif

namespace dcnv3 {
namespace vmma {

using namespace strip;   // staging buffer layout, stage_io, PTX wrappers
using namespace tc;      // tcgen05 wrappers, drain helpers

constexpr int kStrips = 1;                         // strips per CTA
constexpr int kCtasPerSm = 4;
constexpr int kWarpsV = 4, kThreadsV = 128;
constexpr int kRows = 8;                           // output rows per step (two 8 x 4 warp patches)
constexpr int kBandW = 16, kBandH = 16;            // band of one strip and step: 2 blocks of 8 rows x 16 columns
constexpr int kATileBytes = kBandW * kBandH * 128; // 256 cells x 64 pixels x 2 B = 32768 per strip
constexpr int kBlockBytes = kATileBytes / 2;       // 128 cells (one UMMA M block)
// staging of one step's inputs, filled by TMA (three boxes per strip: 8 x 8 pixels x this group's run)
constexpr int kOffRow = 48, kMskRow = 32;          // bytes per pixel: 36 / 18 used, padded to 16-byte multiples
constexpr int kStOff = 0, kStMsk = 64 * kOffRow, kStGout = kStMsk + 64 * kMskRow;
// grad_out of the patch follows: NCH / 8 boxes of [64 px][8 channels] (NCH = channels per group: 16 or 32)
__host__ __device__ constexpr int st_bytes(int nch) { return kStGout + 64 * nch * 2; }   // 7168 / 9216 per stage
constexpr int kStages = 2;
__host__ __device__ constexpr int smem_bytes(int nch) { return 1024 + kStrips * kATileBytes + kStages * st_bytes(nch); }

// zeros for the bulk re-fill of the A tiles (L2-resident)
__device__ __align__(128) unsigned char g_zero_tile[kStrips * kATileBytes];
__device__ __forceinline__ void bulk_fill(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// re-zeroing of the A tile without the L2 round trip: st.bulk (sm_100: UMEMSETS) by the issuing thread, then a plain
// arrive on the refill barrier (DCNV3_VMMA_REFILL=st)
__device__ __forceinline__ void st_bulk_zero(uint32_t dst, uint32_t bytes) {
    asm volatile("st.bulk.weak.shared::cta [%0], %1, 0;" ::"r"(dst), "l"((uint64_t)bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}

struct VParams {
    // run only if *cond > thr (the far-point count of dcnv3_backward_vres.cu; nullptr = always)
    const unsigned long long *cond;
    unsigned long long thr;
    int refill_st;
    int bx_rel, by_rel;      // band origin relative to the first pixel of a strip's 8 x 8 patch
    int tiles_x, tiles_xy, total_tiles;
    int steps;               // 8-row steps per work item
};

#define VMMA_TMEM_LD_8(taddr, r)                                                                 \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"         \
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) \
                 : "r"(taddr))
// `gsel` (NCH == 8 only): which half of the pair's 16 accumulator columns is this group's
template <int NCH>
__device__ __forceinline__ void drain_block(uint32_t tmem_base, int slot, int warp, int lane, float *gv_img, int y0,
                                            int x0, int H, int W, int row_stride, int C, int gsel) {
    const int y = y0 + 2 * warp + (lane >> 4);
    if constexpr (NCH == 8) {
        // a cell's eight fp32 channels are one 32-byte sector: every lane reduces its own cell
        const int x = x0 + (lane & 15);
        float r[8];
        VMMA_TMEM_LD_8(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(slot * 16 + gsel * 8), r);
        tmem_ld_wait();
        bool nz = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) nz |= r[j] != 0.f;
        if (nz && (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W) {
            float *p = gv_img + (ptrdiff_t)y * row_stride + (ptrdiff_t)x * C;
            red_add4(p, make_float4(r[0], r[1], r[2], r[3]));
            red_add4(p + 4, make_float4(r[4], r[5], r[6], r[7]));
        }
        return;
    }
    const int xe = x0 + (lane & 14);            // column of the pair's even cell
    const bool oky = (unsigned)y < (unsigned)H;
    float *p = gv_img + (ptrdiff_t)y * row_stride + (ptrdiff_t)xe * C;
#pragma unroll
    for (int h = 0; h < NCH / 16; ++h) {
        float r0[16];
        VMMA_TMEM_LD_16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(slot * NCH + h * 16), r0);
        tmem_ld_wait();
        drain_cells(r0, lane, p + h * 16, oky && (unsigned)xe < (unsigned)W, oky && (unsigned)(xe + 1) < (unsigned)W, C);
    }
}

// AMN: layout of the coefficient tile.  false = K-major rows of 64 pixels with the 128-byte swizzle, 16-bit
// read-modify-writes (r1); true = MN-major core matrices [8 cells of one band row][8 pixels] -- a pixel's eight cells
// are 16 contiguous bytes whose bank group is fixed by the PIXEL (k & 7), a horizontal corner pair that starts on an
// even column is one 32-bit word: three word read-modify-writes per point on average instead of four 16-bit ones, and
// only the word inside the lane's own bank group depends on the data.
template <typename T, int NCH, bool AMN>
__global__ void __launch_bounds__(kThreadsV, kCtasPerSm)
bwd_vmma(const __grid_constant__ CUtensorMap tmap_off, const __grid_constant__ CUtensorMap tmap_msk,
         const __grid_constant__ CUtensorMap tmap_gout, float *__restrict__ gv_acc, const Geom q, const VParams pp) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mma_bar, mma_bar1, zero_bar, full_bar[kStages];
    __shared__ uint32_t tmem_base_s;

    if (pp.cond) {   // the fall-back of the resident-accumulator kernel: only when it left too many points to the far path
        asm volatile("griddepcontrol.wait;" ::: "memory");
        unsigned long long count;                        // a volatile load: must not move above the wait
        asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(count) : "l"(pp.cond) : "memory");
        if (count <= pp.thr) return;
    }
    // (the warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the descriptors of the
    // elected issuing lane in uniform registers; with `tid == 0` every tcgen05.mma sat in an elect / R2UR loop,
    // ~80 cycles per product -- measured in dcnv3_backward_vres.cu)
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    // warps 0-1 (par == 0): builders, thread <-> pixel of the 8 x 8 patch; warps 2-3 join for the drain
    // (a tcgen05.ld reaches the 32 TMEM lanes of the warp's quarter, so 128 cells need four warps)
    const int strip_id = 0, hw = warp & 1, par = warp >> 1;
    const int k = hw * 32 + lane;                       // the pixel = K index inside the strip
    const int px_x = lane & 7, px_y = hw * kPatchH + (lane >> 3);

    unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t a_addr0 = smem_u32(base);                                   // A tiles: [strip][256 rows][128 B]
    unsigned char *stages = base + kStrips * kATileBytes;                      // [stage][strip][off | msk | gout]
    constexpr int NB = NCH == 8 ? 16 : NCH;     // channels of the B operand = accumulator columns per block
    constexpr int kStBytes = st_bytes(NB), kStStrip = kStBytes, kTmemCols = 2 * NB;   // (one strip per CTA)
    const uint32_t st_thr = smem_u32(stages) + strip_id * kStStrip;            // + stage * kStBytes
    const uint32_t a_strip = a_addr0 + strip_id * kATileBytes;
    const uint32_t kc = (uint32_t)k >> 3, kl = ((uint32_t)k & 7u) * 2u;
    const uint32_t a_thr = AMN ? a_strip + kc * 128u + ((uint32_t)k & 7u) * 16u : a_strip + kl;

    const int C = q.G * q.gc, row_stride = q.W * C;

    int t = blockIdx.x;
    if (t >= pp.total_tiles) return;
    auto decode = [&](int tt, int &n, int &g, int &wo0, int &ho0) {
        const int txy = tt % pp.tiles_xy, r = tt / pp.tiles_xy;
        g = r % q.G; n = r / q.G;
        wo0 = (txy % pp.tiles_x) * (kStrips * kStripW); ho0 = (txy / pp.tiles_x) * (pp.steps * kRows);
    };
    int n, g, wo0, ho0;
    decode(t, n, g, wo0, ho0);
    // one step's inputs of both strips: six TMA boxes (offsets, masks, grad_out of the 8 x 8 patches) into a stage
    auto request = [&](unsigned stage, int nn, int gg, int w0, int h0) {
        uint64_t *bar = &full_bar[stage];
        unsigned char *dst = stages + stage * kStBytes;
        mbar_expect_tx(bar, kStBytes);
#pragma unroll
        for (int st = 0; st < kStrips; ++st) {
            // a box starts on a 16-byte boundary of the row: the group's run begins 0..3 words / 0..7 elements in
            tma_load_4d(dst + st * kStStrip + kStOff, &tmap_off, bar, (gg * kP * 4 & ~15) >> 1, w0 + st * kStripW, h0, nn);
            tma_load_4d(dst + st * kStStrip + kStMsk, &tmap_msk, bar, (gg * kP * 2 & ~15) >> 1, w0 + st * kStripW, h0, nn);
            // grad_out as two boxes of 8 channels: [half][64 px][16 B] is the MMA's B operand as it lands (MN-major)
#pragma unroll
            for (int c8 = 0; c8 < NB / 8; ++c8)   // (NCH == 8: the 16-channel run of the group pair)
                tma_load_4d(dst + st * kStStrip + kStGout + c8 * 1024, &tmap_gout, bar,
                            (NCH == 8 ? (gg >> 1) * 16 : gg * NCH) + c8 * 8, w0 + st * kStripW, h0, nn);
        }
    };

    if (tid == 0) {
        mbar_init(&mma_bar, 1);
        mbar_init(&mma_bar1, 1);
        mbar_init(&zero_bar, 1);
        for (int i = 0; i < kStages; ++i) mbar_init(&full_bar[i], 1);
        fence_barrier_init();
        prefetch_tensormap(&tmap_off);
        prefetch_tensormap(&tmap_msk);
        prefetch_tensormap(&tmap_gout);
    }
    unsigned fills = 0, gstep = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {   // the A tiles start (and after every product restart) as zeros: bulk copy, async proxy
        if (pp.refill_st) {
            st_bulk_zero(a_addr0, kStrips * kATileBytes);
            mbar_arrive(&zero_bar);
        } else {
            mbar_expect_tx(&zero_bar, kStrips * kATileBytes);
            bulk_fill(a_addr0, g_zero_tile, kStrips * kATileBytes, &zero_bar);
        }
    }
    const uint32_t idesc = umma_idesc(std::is_same<T, __nv_bfloat16>::value ? 1 : 0, 128, NB) | (AMN ? 1u << 15 : 0u);

    if (tid == 0) request(0, n, g, wo0, ho0);

    unsigned commits = 0;
    bool dep_waited = false;   // programmatic dependent launch: the plane (zeroed by the channel-sum kernel) is first
                               // touched by this CTA's first A build (far points) / drain -- wait for that grid there
    for (;;) {
        const int t_next = t + gridDim.x;
        const bool has_next = t_next < pp.total_tiles;
        int n2 = n, g2 = g, wo2 = wo0, ho2 = ho0;
        if (has_next) decode(t_next, n2, g2, wo2, ho2);
        float *gv_img = gv_acc + (size_t)n * q.H * row_stride + g * q.gc;
        const int band_x0 = wo0 + strip_id * kStripW + pp.bx_rel;      // this thread's strip

#pragma unroll 1
        for (int s = 0; s < pp.steps; ++s) {
            const int hb = ho0 + s * kRows;
            const int wo = wo0 + strip_id * kStripW + px_x, ho = hb + px_y;
            const bool live = wo < q.Wo && ho < q.Ho;
            const int band_y0 = hb + pp.by_rel;
            // ---- this step's inputs have landed; the other stage (read one step ago by every thread, all of
            // which have passed that step's barrier) takes the next step's
            const unsigned stage = gstep & 1u;
            mbar_wait(&full_bar[stage], (gstep >> 1) & 1u);
            ++gstep;
            if (tid == 0) {
                const bool in_item = s + 1 < pp.steps;
                if (in_item) request(stage ^ 1u, n, g, wo0, hb + kRows);
                else if (has_next) request(stage ^ 1u, n2, g2, wo2, ho2);
            }
            uint32_t off[kP] = {}, mw[5] = {};
            const uint32_t sa = st_thr + stage * kStBytes;
            if (par == 0) {   // builders only
                const uint4 o0 = lds128(sa + kStOff + k * kOffRow), o1 = lds128(sa + kStOff + k * kOffRow + 16),
                            o2 = lds128(sa + kStOff + k * kOffRow + 32);
                const uint32_t w[12] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w, o2.x, o2.y, o2.z, o2.w};
                const unsigned osh = ((unsigned)(g * kP * 4) & 15u) >> 2;         // warp-uniform: 0..3 words
#pragma unroll
                for (int p = 0; p < kP; ++p) off[p] = osh == 0 ? w[p] : osh == 1 ? w[p + 1] : osh == 2 ? w[p + 2] : w[p + 3];
                const uint4 m0 = lds128(sa + kStMsk + k * kMskRow), m1 = lds128(sa + kStMsk + k * kMskRow + 16);
                const uint32_t v[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
                const unsigned esh = ((unsigned)(g * kP * 2) & 15u) >> 1;         // warp-uniform: 0..7 elements
                uint32_t u[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) u[j] = (esh & 1u) ? __funnelshift_r(v[j], j + 1 < 8 ? v[j + 1] : 0u, 16) : v[j];
                const unsigned ws = esh >> 1;
#pragma unroll
                for (int j = 0; j < 5; ++j)
                    mw[j] = ws == 0 ? u[j] : ws == 1 ? u[j + 1] : ws == 2 ? u[j + 2] : (j + 3 < 8 ? u[j + 3] : 0u);
            }

            // ---- A build: the pixel's 36 coefficients into its column of the strip's tile
            mbar_wait(&zero_bar, fills & 1u);
            ++fills;
            if (!dep_waited) {
                asm volatile("griddepcontrol.wait;" ::: "memory");
                dep_waited = true;
            }
            if (live && par == 0) {
                const float bw = axis_base(wo, 3, 1, q.pw, 1, q.sigma) - (float)band_x0;
                const float bh = axis_base(ho, 3, 1, q.ph, 1, q.sigma) - (float)band_y0;
#pragma unroll
                for (int p = 0; p < kP; ++p) {
                    const float2 d = unpack2(off[p], T());
                    const float m = f32_of((uint16_t)(mw[p >> 1] >> (16 * (p & 1))), T());
                    const float ub = bw + ((float)(p / 3) + d.x) * q.sigma;
                    const float vb = bh + ((float)(p % 3) + d.y) * q.sigma;
                    const float fw = floorf(ub), fh = floorf(vb);
                    const float lw = ub - fw, lh = vb - fh;
                    // 0 <= x < limit on the float's bit pattern: negative values and NaN compare as large unsigned
                    if (__float_as_uint(ub) < __float_as_uint((float)(kBandW - 1)) &&
                        __float_as_uint(vb) < __float_as_uint((float)(kBandH - 1))) {
                        if (VMMA_DIAG == 1) continue;
                        const uint32_t cx = (uint32_t)(int)fw, ry = (uint32_t)(int)fh;
                        const float hm = (1.f - lh) * m, lm = lh * m, hwt = 1.f - lw;
                        if constexpr (AMN) {
                            // word of (ry, cx): core-matrix group 2 ry + (cx >> 3) (1024 B each), word (cx & 7) >> 1
                            const uint32_t cx1 = cx + 1u;
                            const uint32_t e0 = a_thr + (ry * 2u + (cx >> 3)) * 1024u + ((cx & 6u) << 1);
                            const uint32_t e1 = a_thr + (ry * 2u + (cx1 >> 3)) * 1024u + ((cx1 & 6u) << 1);
                            const bool oddc = cx & 1u;
                            float2 f0 = unpack2(lds32(e0), T()), f2 = unpack2(lds32(e0 + 2048u), T());
                            if (oddc) { f0.y += hm * hwt; f2.y += lm * hwt; }
                            else { f0.x += hm * hwt; f0.y += hm * lw; f2.x += lm * hwt; f2.y += lm * lw; }
                            sts32(e0, pack2(f0.x, f0.y, T()));
                            sts32(e0 + 2048u, pack2(f2.x, f2.y, T()));
                            if (oddc) {
                                float2 f1 = unpack2(lds32(e1), T()), f3 = unpack2(lds32(e1 + 2048u), T());
                                f1.x += hm * lw; f3.x += lm * lw;
                                sts32(e1, pack2(f1.x, f1.y, T()));
                                sts32(e1 + 2048u, pack2(f3.x, f3.y, T()));
                            }
                            continue;
                        }
                        const uint32_t e0 = a_thr + (ry * kBandW + cx) * 128u + ((kc ^ (cx & 7u)) << 4);
                        const uint32_t e1 = a_thr + (ry * kBandW + cx + 1u) * 128u + ((kc ^ ((cx + 1u) & 7u)) << 4);
                        const float a0 = f32_of((uint16_t)lds16(e0), T()), a1 = f32_of((uint16_t)lds16(e1), T());
                        const float a2 = f32_of((uint16_t)lds16(e0 + kBandW * 128), T()), a3 = f32_of((uint16_t)lds16(e1 + kBandW * 128), T());
                        sts16(e0, bits16(a0 + hm * hwt, T()));
                        sts16(e1, bits16(a1 + hm * lw, T()));
                        sts16(e0 + kBandW * 128, bits16(a2 + lm * hwt, T()));
                        sts16(e1 + kBandW * 128, bits16(a3 + lm * lw, T()));
                    } else {
                        // beyond the band: the reference's range test decides whether the point counts at all
                        const float lw_abs = ub + (float)band_x0, lh_abs = vb + (float)band_y0;
                        if (lh_abs > -1.f && lw_abs > -1.f && lh_abs < (float)q.H && lw_abs < (float)q.W) {
                            const float hm = (1.f - lh) * m, lm = lh * m, hwt = 1.f - lw;
                            if constexpr (NCH == 8) {
                                const uint4 g8 = lds128(sa + kStGout + (g & 1) * 1024 + k * 16);
                                far_point<T, 8>(gv_img, q.H, q.W, row_stride, C, (int)fh + band_y0, (int)fw + band_x0,
                                                hm * hwt, hm * lw, lm * hwt, lm * lw, g8, g8);
                            } else {
#pragma unroll 1
                                for (int h = 0; h < NCH / 16; ++h)
                                    far_point<T>(gv_img + h * 16, q.H, q.W, row_stride, C, (int)fh + band_y0, (int)fw + band_x0,
                                                 hm * hwt, hm * lw, lm * hwt, lm * lw, lds128(sa + kStGout + (2 * h) * 1024 + k * 16),
                                                 lds128(sa + kStGout + (2 * h + 1) * 1024 + k * 16));
                            }
                        }
                    }
                }
            }

            // ---- the tiles are complete: hand them to the tensor core
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (warp == 0 && elect_one_sync()) {
                tc_fence_after();
#pragma unroll
                for (int st = 0; st < kStrips; ++st)
#pragma unroll
                    for (int blk = 0; blk < 2; ++blk) {
                        const uint32_t d = tmem_base + (uint32_t)((st * 2 + ((s + blk) & 1)) * NB);
                        const uint32_t aa = a_addr0 + st * kATileBytes + blk * kBlockBytes;
                        const uint32_t bb = smem_u32(stages) + stage * kStBytes + st * kStStrip + kStGout;
#pragma unroll
                        for (int j = 0; j < (VMMA_DIAG == 4 ? 0 : 4); ++j)   // K step = 16 pixels: 32 B of an A row, 256 B of the pixel-major B
                            tc_mma(d, AMN ? umma_desc_mn_plain(aa + j * 256, 128, 1024) : umma_desc_k_sw128(aa + j * 32),
                                   umma_desc_mn_plain(bb + j * 256, 128, 1024), idesc, (uint32_t)(j > 0 || (blk == 0 && s > 0)));
                        // the UPPER block (the one drained below) completes first: its own commit, so that the drain
                        // starts while the lower block's products still run
                        tc_commit(blk == 0 ? &mma_bar : &mma_bar1);
                    }
            }
            mbar_wait(&mma_bar, commits & 1u);
            tc_fence_after();

            if (tid == 0) {   // the tensor core is done with the tiles (both blocks): refill them with zeros
                mbar_wait(&mma_bar1, commits & 1u);
                if (VMMA_DIAG == 2) {
                    mbar_expect_tx(&zero_bar, 0);
                } else if (pp.refill_st) {
                    st_bulk_zero(a_addr0, kStrips * kATileBytes);
                    mbar_arrive(&zero_bar);
                } else {
                    mbar_expect_tx(&zero_bar, kStrips * kATileBytes);
                    bulk_fill(a_addr0, g_zero_tile, kStrips * kATileBytes, &zero_bar);
                }
            }
            ++commits;
            // ---- the band's upper block is final: reductions
            drain_block<NCH>(tmem_base, s & 1, warp, lane, gv_img, band_y0, wo0 + pp.bx_rel, q.H, q.W, row_stride, C, g & 1);
        }
        // ---- the last step's lower block (its products have their own barrier, which only thread 0 has waited on)
        mbar_wait(&mma_bar1, (commits - 1u) & 1u);
        tc_fence_after();
        drain_block<NCH>(tmem_base, pp.steps & 1, warp, lane, gv_img, ho0 + (pp.steps - 1) * kRows + pp.by_rel + 8,
                    wo0 + pp.bx_rel, q.H, q.W, row_stride, C, g & 1);
        if (!has_next) break;
        n = n2; g = g2; wo0 = wo2; ho0 = ho2;
        t = t_next;
    }

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the narrowing pass may be scheduled
    mbar_wait(&zero_bar, fills & 1u);   // the last refill must land before the CTA's memory is released
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
}

template <typename T>
static bool launch_typed(const void *offset, const void *mask, const void *grad_out, float *gv_acc, const Geom &q,
                         cudaStream_t stream, cudaError_t *err, const unsigned long long *cond, unsigned long long thr) {
    if (!backward_vmma_eligible(offset, mask, grad_out, gv_acc, q)) return false;
    VParams pp;
    static const bool refill_st = [] { const char *e = std::getenv("DCNV3_VMMA_REFILL"); return e && e[0] == 's'; }();
    pp.refill_st = refill_st;
    pp.cond = cond;
    pp.thr = thr;
    // nominal taps of a pixel x along an axis: x + a + i*sigma, i = 0..2, a = (1 - pad) - sigma; band centred on them
    const float a_w = (float)(1 - q.pw) - q.sigma, a_h = (float)(1 - q.ph) - q.sigma;
    pp.bx_rel = (int)std::floor(a_w + q.sigma + 0.5f * (kStripW - 1) + 0.5f - 0.5f * kBandW);
    pp.by_rel = (int)std::floor(a_h + q.sigma + 0.5f * (kRows - 1) + 0.5f - 0.5f * kBandH);
    pp.tiles_x = (q.Wo + kStrips * kStripW - 1) / (kStrips * kStripW);
    // tall items: an item drains (steps + 1) blocks for `steps` patches
    const int max_steps = (q.Ho + kRows - 1) / kRows;
    pp.steps = std::max(1, std::min(max_steps, 5));
    if (const char *e = std::getenv("DCNV3_VSTEPS")) pp.steps = std::max(1, std::min(max_steps, atoi(e)));
    const int tile_h = pp.steps * kRows;
    const int tiles_y = (q.Ho + tile_h - 1) / tile_h;
    const long long total = (long long)pp.tiles_x * tiles_y * q.G * q.N;
    if (total >= (1LL << 31)) return false;
    pp.tiles_xy = pp.tiles_x * tiles_y;
    pp.total_tiles = (int)total;
    CUtensorMap to, tm, tg;
    const int dtype = std::is_same<T, __half>::value ? 1 : 2;
    if (!make_run_tensor_map(&to, offset, dtype, q.N, q.Ho, q.Wo, q.G * kP * 2, kOffRow / 2)) return false;
    if (!make_run_tensor_map(&tm, mask, dtype, q.N, q.Ho, q.Wo, q.G * kP, kMskRow / 2)) return false;
    if (!make_run_tensor_map(&tg, grad_out, dtype, q.N, q.Ho, q.Wo, q.G * q.gc, 8)) return false;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    // (as the conditional fall-back of dcnv3_backward_vres.cu the kernel almost always exits at once: one CTA per SM
    // keeps that exit cheap; the work loop strides by the grid size)
    const int ctas = (int)std::min<long long>(total, (long long)(cond ? 1 : kCtasPerSm) * num_sms);
    // coefficient-tile layout: K-major swizzled tile (default); DCNV3_VMMA_A=m = MN-major with paired 32-bit
    // read-modify-writes (measured: 10 % fewer shared-memory wavefronts, but +8 us -- the kernel is bound by its
    // serial phases, not by the scatter's bank conflicts)
    static const bool a_mn = [] { const char *e = std::getenv("DCNV3_VMMA_A"); return e && (e[0] == 'm' || e[0] == 'M'); }();
    auto go = [&](auto kern, int nb) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(nb));
        *err = pdl_launch(pdl_for(q) || cond, kern, dim3(ctas), dim3(kThreadsV), smem_bytes(nb), stream, to, tm, tg, gv_acc, q, pp);
    };
    if (q.gc == 8) { if (a_mn) go(bwd_vmma<T, 8, true>, 16); else go(bwd_vmma<T, 8, false>, 16); }
    else if (q.gc == 16) { if (a_mn) go(bwd_vmma<T, 16, true>, 16); else go(bwd_vmma<T, 16, false>, 16); }
    else { if (a_mn) go(bwd_vmma<T, 32, true>, 32); else go(bwd_vmma<T, 32, false>, 32); }
    if (*err == cudaSuccess) *err = cudaGetLastError();
    return true;
}

}  // namespace vmma

bool backward_vmma_eligible(const void *offset, const void *mask, const void *grad_out, const float *gv_acc, const Geom &q) {
    using namespace strip;
    if ((q.gc != 8 && q.gc != 16 && q.gc != 32) || q.kh != 3 || q.kw != 3 || q.sh != 1 || q.sw != 1 || q.dh != 1 || q.dw != 1) return false;
    if (!(q.sigma >= 0.5f && q.sigma <= 1.25f)) return false;   // band = taps +- 3 px
    // TMA staging: 16-byte aligned bases and row strides (G * 18 B for the masks: G % 8 == 0)
    if (((uintptr_t)grad_out | (uintptr_t)gv_acc | (uintptr_t)offset | (uintptr_t)mask) % 16 || q.G % 8) return false;
    if ((long long)q.N * q.Ho * q.Wo == 0) return false;
    if ((long long)((q.Wo + 7) / 8) * ((q.Ho + 7) / 8) * q.G * q.N >= (1LL << 31)) return false;
    return true;
}

// grad_value only (accumulated into the zeroed fp32 plane gv_acc); tcgen05 / TMEM form.
bool try_launch_backward_vmma(const void *offset, const void *mask, const void *grad_out, float *gv_acc,
                              const Geom &q, int dtype, cudaStream_t stream, cudaError_t *err,
                              const unsigned long long *cond, unsigned long long thr) {
    if ((long long)q.N * q.Ho * q.Wo == 0) return false;
    if (dtype == 1) return vmma::launch_typed<__half>(offset, mask, grad_out, gv_acc, q, stream, err, cond, thr);
    if (dtype == 2) return vmma::launch_typed<__nv_bfloat16>(offset, mask, grad_out, gv_acc, q, stream, err, cond, thr);
    return false;
}

}  // namespace dcnv3
