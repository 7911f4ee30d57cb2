// dcnv3_backward_vband.cu -- grad_value of the DCNv3 core backward for 16-bit I/O, group_channels == 16,
// 3x3 / stride 1 / dilation 1, with the tcgen05 accumulator as a CIRCULAR band in tensor memory and SMALL coefficient
// tiles (one product = 16 pixels x 128 band cells).
//
// What it computes (reference dcnv3_im2col_cuda.cuh:82-147, col2im bilinear), per group:
//     grad_value[cell, c] += sum over (pixel, point, corner) hitting `cell` of  w_corner * m * grad_out[pixel, c]
// i.e. D[cells x 16 ch] = A[cells x pixels] . G[pixels x 16 ch] with A holding every pixel's 36 coefficients.
//
// Why a second form next to dcnv3_backward_vmma.cu (profiles/README.md, r1_v7 -> r2): that kernel scatters the
// coefficients into a 256-cell x 64-pixel tile with 16-bit shared-memory read-modify-writes whose bank is data
// dependent (72 % of its shared-memory wavefronts were conflicts), refills the 93 %-empty tile with zeros and makes
// the tensor core read it back: 1.9 wavefronts of shared memory per sampled point, phases strictly serial.  Here
//   * one product covers 16 pixels (8 wide x 2 rows) x a band of 128 cells (16 columns x 8 rows: taps +-3 px in x,
//     +-2 px in y) = ONE tcgen05.mma (M 128, N 16, K 16); the tensor core reads 4 KB per 144 points instead of 32 KB
//     per 576.  A is MN-major (8 cells of one pixel = 16 bytes), so a builder thread (= pixel) owns a column whose
//     16-byte pieces sit in a bank group fixed by the PIXEL, whatever cell is hit.  Two builders:
//       DENSE  the whole 128-cell column in registers as a sum of nine separable rank-1 terms (m hat_y) (x) hat_x
//              -- 64 packed HFMA2 per point -- written with sixteen conflict-free 128-bit stores; no
//              read-modify-write, no zero fill;
//       SPARSE the thread zeroes its column (sixteen 128-bit stores) and adds each point's 2 x 2 block with 32-bit
//              packed read-modify-writes (a horizontal corner pair that starts on an even column is ONE word);
//   * the band slides down the strip two rows per product.  TMEM lane = (band row mod 8) * 16 + column, so the
//     accumulator is a ring: after product s the two oldest rows (one 32-lane quarter = one warp's slice of tensor
//     memory) are final -- that warp drains them (tcgen05.ld -> whole-sector vector reductions into the fp32 plane),
//     overwrites them with zeros (tcgen05.st) and issues product s+1 itself.  Nothing is drained twice inside an item;
//   * warp-specialised CTA: warps 0-3 build (warp = one of FOUR groups, lane = pixel of an 8 x 4 half-patch = two
//     products), warps 4-7 own the four TMEM lane quarters, pass the issue token round through named barriers (a
//     waiting warp executes nothing) and also request the inputs: offsets / masks / grad_out of four groups arrive as
//     TMA boxes three half-patches ahead (4 stages); setmaxnreg moves registers from the drain warps to the builders.
// A point whose corner block leaves the band goes straight to the plane, the builder warp together.
#include "../dcnv3_common.cuh"
#include "../dcnv3_launch.h"
#include "../dcnv3_tma.cuh"
#include "../dcnv3_strip_io.cuh"
#include "../dcnv3_tc.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>

namespace dcnv3 {
namespace vband {

using namespace strip;
using namespace tc;

constexpr int kGroupsV = 4;                         // groups per CTA (lock step)
constexpr int kHpRows = 4, kHpPix = 32;             // half-patch: 8 x 4 pixels = two products
constexpr int kBandW = 16, kBandH = 8;              // cells of one product: 128 = UMMA M
constexpr int kTileBytes = kBandW * kBandH * 16 * 2;            // 128 cells x 16 pixels x 2 B = 4096
constexpr int kABufBytes = kGroupsV * 2 * kTileBytes;           // one half-patch of four groups: 32768
constexpr int kOffPitch = kGroupsV * kP * 4, kMskPitch = 80;    // staged bytes per pixel: 144 / 72 (+ up to 8 of shift)
constexpr int kStOff = 0, kStMsk = kHpPix * kOffPitch, kStGout = kStMsk + kHpPix * kMskPitch;
constexpr int kGoutC8 = kHpPix * 16;                            // one 8-channel block of the half-patch: 512
constexpr int kStBytes = kStGout + 2 * kGroupsV * kGoutC8;      // 11264
constexpr int kStages = 4;
constexpr int kAhead = 3;                           // half-patches between a request and its use
constexpr int kSmemBytes = 1024 + 2 * kABufBytes + kStages * kStBytes;   // 111616: two CTAs per SM
constexpr int kThreadsB = 256;
constexpr int kTmemCols = 2 * kGroupsV * 16;        // two accumulator sets (items alternate) x 4 groups x 16 channels

// development only (DCNV3_VBAND_DIAG & 64): cycle stamps of CTA 0 -- drain side [link][7], builder side [hp][5]
__device__ long long g_vband_dbg[2][256][8];

struct BParams {
    int bx_rel, by_rel;      // band origin relative to the first pixel of a product's 8 x 2 patch
    float c_w, c_h;          // band-relative anchor of pixel (0, 0): (1 - pad) - sigma - b?_rel
    int tiles_x, segs, seg_hp, gblocks, total_items;
    int diag;                // development only (DCNV3_VBAND_DIAG): leave one phase out, results are WRONG
};
struct Item { int n, g0, x0, y0, hps; };

__device__ __forceinline__ Item decode_item(int t, const BParams &pp, int Ho) {
    Item it;
    const int seg = t % pp.segs; int r = t / pp.segs;
    const int sx = r % pp.tiles_x; r /= pp.tiles_x;
    it.g0 = (r % pp.gblocks) * kGroupsV; it.n = r / pp.gblocks;
    it.x0 = sx * kStripW; it.y0 = seg * pp.seg_hp * kHpRows;
    const int left = (Ho - it.y0 + kHpRows - 1) / kHpRows;
    it.hps = left < pp.seg_hp ? left : pp.seg_hp;
    return it;
}
// position in the CTA's sequence of half-patches (items blockIdx.x, + gridDim.x, ...)
struct Cursor {
    int t, h;
    Item it;
    __device__ __forceinline__ void init(const BParams &pp, int Ho) { t = blockIdx.x; h = 0; if (t < pp.total_items) it = decode_item(t, pp, Ho); }
    __device__ __forceinline__ bool live(const BParams &pp) const { return t < pp.total_items; }
    __device__ __forceinline__ void next(const BParams &pp, int Ho) {
        if (++h == it.hps) { h = 0; t += gridDim.x; if (t < pp.total_items) it = decode_item(t, pp, Ho); }
    }
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t *bar, unsigned parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Waiting without stealing issue slots.  Polling an mbarrier (try_wait loop, with or without a suspend-time hint, or
// test_wait + nanosleep(40)) came back every few tens of cycles: the four drain warps of a CTA wait most of the time and
// executed a quarter to a half of the kernel's instructions, on the ALU pipe the builders need.  So: a warp whose turn
// is far away BLOCKS on a named barrier (no instructions at all) until the warp before it in the chain passes the
// token, and only then polls the tensor core's completion barrier (for about one product's latency).
__device__ __forceinline__ void mbar_wait_idle(uint64_t *bar, unsigned parity, unsigned ns) {
    while (!mbar_test(bar, parity)) __nanosleep(ns);
}
__device__ __forceinline__ void token_pass(int q) { asm volatile("bar.arrive %0, 64;" ::"r"(2 + q) : "memory"); }
__device__ __forceinline__ void token_take(int q) { asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// packed 16-bit pair arithmetic in the I/O dtype (the A operand's format)
template <typename T> struct Pk;
template <> struct Pk<__nv_bfloat16> {
    static __device__ __forceinline__ uint32_t fma2(uint32_t a, uint32_t b, uint32_t c) {
        uint32_t d; asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
    }
    static __device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) {
        uint32_t d; asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
    }
};
template <> struct Pk<__half> {
    static __device__ __forceinline__ uint32_t fma2(uint32_t a, uint32_t b, uint32_t c) {
        uint32_t d; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
    }
    static __device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) {
        uint32_t d; asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
    }
};

// instruction descriptor: fp32 accumulate, A and B both MN-major (bits 15 / 16), N >> 3, M >> 4
__host__ __device__ constexpr uint32_t band_idesc(int fmt) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(16 >> 3) << 17) |
           ((uint32_t)(128 >> 4) << 24);
}
#define VBAND_TMEM_ST_ZERO_16(taddr)                                                                          \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" \
                 ::"r"(taddr), "r"(0u) : "memory")
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 4-D tensor map over a [N, Ho, Wo, row_elems] tensor of 16-bit elements, box (box_elems, 8, 4, 1)
static bool make_hp_tensor_map(CUtensorMap *map, const void *base, int dtype, int N, int Ho, int Wo, int row_elems,
                               int box_elems) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t es = 2;
    const CUtensorMapDataType dt = dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const cuuint64_t dims[4] = {(cuuint64_t)row_elems, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)row_elems * es, (cuuint64_t)Wo * row_elems * es,
                                   (cuuint64_t)Ho * Wo * row_elems * es};
    const cuuint32_t box[4] = {(cuuint32_t)box_elems, (cuuint32_t)kStripW, (cuuint32_t)kHpRows, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    return fn(map, dt, 4, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Points beyond the band (about 1 % for N(0, 1)-pixel offsets), processed by the whole builder warp, out of line.
// `far` = this lane's 9-bit mask of such points.  Per pass up to EIGHT points, four lanes each: lane = corner reads the
// owner's staged offset / mask / grad_out, recomputes the tap and sends its coefficient x the 16 channels (four whole
// 16-byte reductions = two sectors).  The reference's range test decides whether such a point counts at all.
template <typename T>
__device__ __noinline__ void far_points(unsigned far, int lane, uint32_t stage_addr, int w, uint32_t m_shift, float c_w,
                                        float c_h, float sigma, int H, int W, float *gv_img, int band_x0, int band_y0_hp,
                                        int row_stride, int C) {
    const int slot = lane >> 2, t = lane & 3;
    unsigned any;
#pragma unroll 1
    while ((any = __ballot_sync(0xffffffffu, far != 0)) != 0) {
        // the slot-th lane that still holds a point (eight slots per pass)
        unsigned src = 0xffffffffu, rest = any;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const unsigned pos = rest ? (unsigned)__ffs(rest) - 1u : 0xffffffffu;
            if (j == slot) src = pos;
            rest &= rest - 1u;
        }
        const unsigned fsrc = __shfl_sync(0xffffffffu, far, src & 31u);
        if (far && __popc(any & ((1u << lane) - 1u)) < 8) far &= far - 1;   // the owners served in this pass
        if (src != 0xffffffffu) {
            const int p = __ffs(fsrc) - 1, pi = (p * 11) >> 5;
            const float2 d = unpack2(lds32(stage_addr + kStOff + src * kOffPitch + w * (kP * 4) + p * 4), T());
            const float m = f32_of((uint16_t)lds16(stage_addr + kStMsk + src * kMskPitch + w * (kP * 2) + m_shift + p * 2), T());
            const int prow = (int)src >> 3;
            const float ub = c_w + (float)(src & 7u) + ((float)pi + d.x) * sigma;
            const float vb = c_h + (float)(prow & 1) + ((float)(p - 3 * pi) + d.y) * sigma;
            const int band_y0 = band_y0_hp + 2 * (prow >> 1);
            const float fw = floorf(ub), fh = floorf(vb);
            const float lw = ub - fw, lh = vb - fh;
            const float lw_abs = ub + (float)band_x0, lh_abs = vb + (float)band_y0;
            const int hh = (int)fh + band_y0 + (t >> 1), ww = (int)fw + band_x0 + (t & 1);
            const float cf = ((t >> 1) ? lh : 1.f - lh) * ((t & 1) ? lw : 1.f - lw) * m;
            if (lh_abs > -1.f && lw_abs > -1.f && lh_abs < (float)H && lw_abs < (float)W &&
                (unsigned)hh < (unsigned)H && (unsigned)ww < (unsigned)W && cf != 0.f) {
                const uint32_t ga = stage_addr + kStGout + (2 * w) * kGoutC8 + src * 16;
                float g[16];
                unpack<T>(lds128(ga), g);
                unpack<T>(lds128(ga + kGoutC8), g + 8);
                float *dst = gv_img + (ptrdiff_t)hh * row_stride + (ptrdiff_t)ww * C;
#pragma unroll
                for (int e = 0; e < 16; e += 4)
                    red_add4(dst + e, make_float4(cf * g[e], cf * g[e + 1], cf * g[e + 2], cf * g[e + 3]));
            }
        }
    }
}

template <typename T, bool SPARSE>
__global__ void __launch_bounds__(kThreadsB, 2)
bwd_vband(const __grid_constant__ CUtensorMap tmap_off, const __grid_constant__ CUtensorMap tmap_msk,
          const __grid_constant__ CUtensorMap tmap_gout, float *__restrict__ gv_acc, const Geom q, const BParams pp) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kStages], a_full[2], a_free[2], commit_bar[4];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t a_addr = smem_u32(base);                          // A tiles: [buf 2][group 4][product 2][4096]
    unsigned char *stages = base + 2 * kABufBytes;                   // [stage][off | msk | gout]
    const uint32_t st_addr = smem_u32(stages);
    const int C = q.G * q.gc, row_stride = q.W * C;
    const int total = pp.total_items;

    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) mbar_init(&full_bar[i], 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_free[i], 1); }
        for (int i = 0; i < 4; ++i) mbar_init(&commit_bar[i], 1);
        fence_barrier_init();
        prefetch_tensormap(&tmap_off);
        prefetch_tensormap(&tmap_msk);
        prefetch_tensormap(&tmap_gout);
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < 4) {
        // ======================================================================================== builders
        asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        const int w = warp;                                     // this warp's group inside the block of four
        const int px_x = lane & 7, prow = lane >> 3, sub = prow >> 1;
        const float bw = pp.c_w + (float)px_x, bh = pp.c_h + (float)(prow & 1);
        const uint32_t my_tile = a_addr + (w * 2 + sub) * kTileBytes + (lane & 15) * 16;
        const uint32_t my_off = st_addr + kStOff + lane * kOffPitch + w * (kP * 4);
        const uint32_t my_msk = st_addr + kStMsk + lane * kMskPitch + w * (kP * 2);

        unsigned gh = 0;
        bool dep_waited = false;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            const Item it = decode_item(t, pp, q.Ho);
            const uint32_t m_shift = (uint32_t)(it.g0 * kP * 2) & 15u;
            float *gv_img = gv_acc + (size_t)it.n * q.H * row_stride + (it.g0 + w) * 16;
            const int band_x0 = it.x0 + pp.bx_rel;
#pragma unroll 1
            for (int h = 0; h < it.hps; ++h, ++gh) {
                const unsigned stage = gh & (kStages - 1), buf = gh & 1u;
                const bool dbg = (pp.diag & 64) && blockIdx.x == 0 && tid == 0 && gh < 256;
                if (dbg) g_vband_dbg[1][gh][0] = clock64();
                // A[buf] was read by the products of half-patch gh - 2
                if (gh >= 2) mbar_wait_idle(&a_free[buf], ((gh >> 1) - 1u) & 1u, 32);
                if (dbg) g_vband_dbg[1][gh][1] = clock64();
                mbar_wait_idle(&full_bar[stage], (gh >> 2) & 1u, 32);
                const uint32_t sb = stage * kStBytes;
                if (dbg) g_vband_dbg[1][gh][2] = clock64();
                const uint32_t tile = my_tile + buf * kABufBytes;
                const int rot = (4 * h + 2 * sub) & 7;          // band row c lives in ring row (c + 2 s) mod 8
                unsigned far = 0;
                if (pp.diag & 1) {
                    fence_proxy_async();
                    mbar_arrive(&a_full[buf]);
                    continue;
                }
                if constexpr (SPARSE) {
                    // ---- zero the pixel's column, then add each point's 2 x 2 block with packed 32-bit read-modify-writes
#pragma unroll
                    for (int c = 0; c < 2 * kBandH; ++c) sts128(tile + c * 256, 0u, 0u, 0u, 0u);
                    auto prep = [&](int idx, uint32_t (&wd)[4], uint32_t (&ad)[4], bool &in, bool &odd) {
                        const uint32_t o2 = lds32(my_off + sb + idx * 4), m16 = lds16(my_msk + sb + m_shift + idx * 2);
                        const float2 d = unpack2(o2, T());
                        const float m = f32_of((uint16_t)m16, T());
                        const int pi = (idx * 11) >> 5;                    // idx / 3 for idx < 9
                        const float ub = bw + ((float)pi + d.x) * q.sigma;
                        const float vb = bh + ((float)(idx - 3 * pi) + d.y) * q.sigma;
                        in = __float_as_uint(ub) < __float_as_uint((float)(kBandW - 1)) &&
                             __float_as_uint(vb) < __float_as_uint((float)(kBandH - 1));
                        if (!in) far |= 1u << idx;
                        const float fw = floorf(ub), fh = floorf(vb);
                        const float lw = ub - fw, lh = vb - fh, hwt = 1.f - lw;
                        const int ifw = (int)fw, ifh = (int)fh;
                        odd = ifw & 1;
                        const float hm = (1.f - lh) * m, lm = lh * m;
                        const float c00 = hm * hwt, c01 = hm * lw, c10 = lm * hwt, c11 = lm * lw;
                        // word A = the column pair that holds fw, word B = the next pair (only an odd fw reaches it)
                        wd[0] = pack2(odd ? 0.f : c00, odd ? c00 : c01, T());
                        wd[1] = pack2(odd ? 0.f : c10, odd ? c10 : c11, T());
                        wd[2] = pack2(c01, 0.f, T());
                        wd[3] = pack2(c11, 0.f, T());
                        const uint32_t r0 = tile + (uint32_t)((ifh + rot) & 7) * 512u, r1 = tile + (uint32_t)((ifh + 1 + rot) & 7) * 512u;
                        const uint32_t ca = (uint32_t)(ifw >> 3) * 256u + (uint32_t)(ifw & 6) * 2u;
                        const uint32_t xb = (uint32_t)(ifw & ~1) + 2u, cb = (xb >> 3) * 256u + (xb & 6u) * 2u;
                        ad[0] = r0 + ca; ad[1] = r1 + ca; ad[2] = r0 + cb; ad[3] = r1 + cb;
                    };
                    auto rmw = [&](const uint32_t (&wd)[4], const uint32_t (&ad)[4], bool in, bool odd) {
                        if (in) {
                            const uint32_t v0 = lds32(ad[0]), v1 = lds32(ad[1]);
                            uint32_t v2 = 0u, v3 = 0u;
                            if (odd) { v2 = lds32(ad[2]); v3 = lds32(ad[3]); }
                            sts32(ad[0], Pk<T>::add2(v0, wd[0]));
                            sts32(ad[1], Pk<T>::add2(v1, wd[1]));
                            if (odd) { sts32(ad[2], Pk<T>::add2(v2, wd[2])); sts32(ad[3], Pk<T>::add2(v3, wd[3])); }
                        }
                    };
                    uint32_t wa[4], aa[4], wb[4], ab[4];
                    bool ia, oa, ib, ob;
                    prep(0, wa, aa, ia, oa);
#pragma unroll 1
                    for (int p = 0; p < kP - 1; p += 2) {
                        prep(p + 1, wb, ab, ib, ob);
                        rmw(wa, aa, ia, oa);
                        prep(p + 2, wa, aa, ia, oa);
                        rmw(wb, ab, ib, ob);
                    }
                    rmw(wa, aa, ia, oa);
                    if (dbg) g_vband_dbg[1][gh][3] = clock64();
                } else {
                    uint32_t acc[kBandH][8];
#pragma unroll
                    for (int c = 0; c < kBandH; ++c)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[c][j] = 0u;
                    // the nine points, two per iteration of a ROLLED loop (a point is ~140 instructions = 2.2 KB: the body
                    // stays in the 6 KB L0 instruction cache; fully unrolled, 19 % of the stall samples were instruction
                    // fetches).  Software pipelined: the operands of point p + 1 are prepared (ALU pipe) while the 64
                    // HFMA2 of point p run (FMA pipe, one per two cycles).
                    auto prep = [&](int idx, uint32_t (&hx)[8], uint32_t (&hy)[kBandH]) {
                        const uint32_t o2 = lds32(my_off + sb + idx * 4), m16 = lds16(my_msk + sb + m_shift + idx * 2);
                        const float2 d = unpack2(o2, T());
                        const float m = f32_of((uint16_t)m16, T());
                        const int pi = (idx * 11) >> 5;                    // idx / 3 for idx < 9
                        const float ub = bw + ((float)pi + d.x) * q.sigma;
                        const float vb = bh + ((float)(idx - 3 * pi) + d.y) * q.sigma;
                        // 0 <= x < limit on the float's bit pattern: negative values and NaN compare as large unsigned
                        const bool inb = __float_as_uint(ub) < __float_as_uint((float)(kBandW - 1)) &&
                                         __float_as_uint(vb) < __float_as_uint((float)(kBandH - 1));
                        const float fw = floorf(ub), fh = floorf(vb);
                        const float lw = ub - fw, lh = vb - fh;
                        int ifw = (int)fw, ifh = (int)fh;
                        if (!inb) { far |= 1u << idx; ifw = -64; ifh = -64; }
                        // hat_x: (1 - lw) at column fw, lw at fw + 1, as packed pairs of columns (2j, 2j + 1)
                        const uint32_t P = pack2(1.f - lw, lw, T());
                        const int e = ifw >> 1;
                        const bool odd = ifw & 1;
                        const uint32_t W0 = odd ? (P << 16) : P, W1 = odd ? (P >> 16) : 0u;
                        // m * hat_y: (1 - lh) m at row fh, lh m at fh + 1, broadcast to both halves
                        const float hm = (1.f - lh) * m, lm = lh * m;
                        const uint32_t Q0 = pack2(hm, hm, T()), Q1 = pack2(lm, lm, T());
#pragma unroll
                        for (int j = 0; j < 8; ++j) hx[j] = j == e ? W0 : (j == e + 1 ? W1 : 0u);
#pragma unroll
                        for (int c = 0; c < kBandH; ++c) hy[c] = c == ifh ? Q0 : (c == ifh + 1 ? Q1 : 0u);
                    };
                    auto rank1 = [&](const uint32_t (&hx)[8], const uint32_t (&hy)[kBandH]) {
#pragma unroll
                        for (int c = 0; c < kBandH; ++c)
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[c][j] = Pk<T>::fma2(hy[c], hx[j], acc[c][j]);
                    };
                    uint32_t ax[8], ay[kBandH], bx[8], by[kBandH];
                    prep(0, ax, ay);
#pragma unroll 1
                    for (int p = 0; p < kP - 1; p += 2) {
                        prep(p + 1, bx, by);
                        rank1(ax, ay);
                        prep(p + 2, ax, ay);
                        rank1(bx, by);
                    }
                    rank1(ax, ay);
                    if (dbg) g_vband_dbg[1][gh][3] = clock64();
#pragma unroll
                    for (int c = 0; c < kBandH; ++c) {
                        const uint32_t ra = tile + (uint32_t)((c + rot) & 7) * 512u;
                        sts128(ra, acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
                        sts128(ra + 256u, acc[c][4], acc[c][5], acc[c][6], acc[c][7]);
                    }
                }

                if (__ballot_sync(0xffffffffu, far != 0) && !(pp.diag & 2)) {
                    if (!dep_waited) { asm volatile("griddepcontrol.wait;" ::: "memory"); dep_waited = true; }
                    far_points<T>(far, lane, st_addr + sb, w, m_shift, pp.c_w, pp.c_h, q.sigma, q.H, q.W, gv_img, band_x0,
                                  it.y0 + h * kHpRows + pp.by_rel, row_stride, C);
                }
                fence_proxy_async();          // the tile is read by the tensor core (async proxy)
                mbar_arrive(&a_full[buf]);
                if (dbg) g_vband_dbg[1][gh][4] = clock64();
            }
        }
    } else {
        // ============================================================================ drain / issue / request warps
        asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
        const int qd = warp - 4;                                // this warp's TMEM lane quarter = ring rows 2 qd, 2 qd + 1
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(qd * 32) << 16);
        const uint32_t idesc = band_idesc(std::is_same<T, __nv_bfloat16>::value ? 1 : 0);
        const uint64_t adesc0 = umma_desc_mn_plain(a_addr, 128, 256), bdesc0 = umma_desc_mn_plain(st_addr + kStGout, 128, kGoutC8);
        bool dep_waited = false;

        // inputs of one half-patch: ten TMA boxes into stage ghp & 3 (one lane)
        auto request = [&](const Cursor &c, unsigned ghp) {
            const unsigned stage = ghp & (kStages - 1);
            uint64_t *bar = &full_bar[stage];
            unsigned char *dst = stages + stage * kStBytes;
            const int y = c.it.y0 + c.h * kHpRows;
            if (pp.diag & 32) { mbar_expect_tx(bar, 0); return; }
            mbar_expect_tx(bar, kStBytes);
            tma_load_4d(dst + kStOff, &tmap_off, bar, c.it.g0 * kP * 2, c.it.x0, y, c.it.n);
            // the four groups' 72-byte mask run is staged from the 16-byte boundary below it
            tma_load_4d(dst + kStMsk, &tmap_msk, bar, (c.it.g0 * kP * 2 & ~15) >> 1, c.it.x0, y, c.it.n);
#pragma unroll
            for (int c8 = 0; c8 < 2 * kGroupsV; ++c8)   // grad_out: [8-channel block][32 px][16 B] = the MMA's B operand as it lands
                tma_load_4d(dst + kStGout + c8 * kGoutC8, &tmap_gout, bar, c.it.g0 * 16 + c8 * 8, c.it.x0, y, c.it.n);
        };
        // the four products of one link (one per group): A = the builders' tiles, B = grad_out as the TMA delivered it
        auto issue = [&](unsigned ghn, int sub, unsigned ell, uint32_t acc_col, uint32_t accumulate) {
            const unsigned buf = ghn & 1u, stage = ghn & (kStages - 1);
            mbar_wait_idle(&a_full[buf], (ghn >> 1) & 1u, 100);   // (already complete for the second product of a half-patch)
            tc_fence_after();
            if (lane == 0) {
                const uint64_t aa = adesc0 + (uint64_t)((buf * kABufBytes + sub * kTileBytes) >> 4);
                const uint64_t bb = bdesc0 + (uint64_t)((stage * kStBytes + sub * 256) >> 4);
#pragma unroll
                for (int g = 0; g < kGroupsV; ++g)
                    if (!(pp.diag & 16))
                        tc_mma(tmem_base + acc_col + g * 16, aa + (uint64_t)((g * 2 * kTileBytes) >> 4), bb + (uint64_t)((2 * g * kGoutC8) >> 4),
                               idesc, accumulate);
                tc_commit(&commit_bar[ell & 3u]);
                if (sub == 1) tc_commit(&a_free[buf]);      // both products that read A[buf] / the stage are done
            }
            __syncwarp();
        };
        // one quarter (32 cells x 4 groups x 16 channels) -> reductions
        auto reduce = [&](const float (&r)[kGroupsV][16], float *gv_img, int y, int x0) {
            if (pp.diag & 4) return;
            if (!dep_waited) { asm volatile("griddepcontrol.wait;" ::: "memory"); dep_waited = true; }
            const int xe = x0 + (lane & 14);
            const bool oky = (unsigned)y < (unsigned)q.H;
            float *p = gv_img + (ptrdiff_t)y * row_stride + (ptrdiff_t)xe * C;
#pragma unroll
            for (int g = 0; g < kGroupsV; ++g)
                drain_cells(r[g], lane, p + g * 16, oky && (unsigned)xe < (unsigned)q.W, oky && (unsigned)(xe + 1) < (unsigned)q.W, C);
        };

        // the request cursor runs kAhead half-patches in front of the chain
        Cursor rc;
        rc.init(pp, q.Ho);
        for (int i = 0; i < kAhead; ++i) {
            if (qd == 0 && lane == 0 && rc.live(pp)) request(rc, (unsigned)i);
            if (rc.live(pp)) rc.next(pp, q.Ho);
        }
        unsigned ell = 0, gh_base = 0, item_idx = 0;
        if (qd == 0) issue(0, 0, 0, 0, 0);
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++item_idx) {
            const Item it = decode_item(t, pp, q.Ho);
            const int S = 2 * it.hps;
            const uint32_t acc_col = (item_idx & 1u) * (kGroupsV * 16);
            float *gv_img = gv_acc + (size_t)it.n * q.H * row_stride + it.g0 * 16;
            const int x0 = it.x0 + pp.bx_rel, band_y0 = it.y0 + pp.by_rel;
            const bool more = t + (int)gridDim.x < total;
#pragma unroll 1
            for (int s = 0; s < S; ++s, ++ell) {
                const bool last = s == S - 1, mine = (s & 3) == qd;
                const bool req = !(s & 1) && rc.live(pp);       // this link opens half-patch gh: request gh + kAhead
                if (!mine && !last) {
                    if (req) rc.next(pp, q.Ho);
                    continue;
                }
                const bool dbg = (pp.diag & 64) && blockIdx.x == 0 && lane == 0 && ell < 256 && mine;
                if (dbg) g_vband_dbg[0][ell][0] = clock64();
                if (mine) {
                    if (ell > 0) token_take(qd);                       // blocked until the previous link's handler has issued this one
                    if (dbg) g_vband_dbg[0][ell][1] = clock64();
                    mbar_wait(&commit_bar[ell & 3u], (ell >> 2) & 1u);  // its products are in flight: short poll
                    if (dbg) g_vband_dbg[0][ell][2] = clock64();
                } else {
                    mbar_wait_idle(&commit_bar[ell & 3u], (ell >> 2) & 1u, 400);   // (item end: once per item)
                }
                tc_fence_after();
                float r[kGroupsV][16];
                if (!(pp.diag & 8)) {
#pragma unroll
                    for (int g = 0; g < kGroupsV; ++g) VMMA_TMEM_LD_16(lane_taddr + acc_col + g * 16, r[g]);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int g = 0; g < kGroupsV; ++g)
#pragma unroll
                        for (int j = 0; j < 16; ++j) r[g][j] = 0.f;
                }
                if (dbg) g_vband_dbg[0][ell][3] = clock64();
                if (mine) {
                    if (!last && !(pp.diag & 8)) {   // the quarter becomes the band's two NEW rows
#pragma unroll
                        for (int g = 0; g < kGroupsV; ++g) VBAND_TMEM_ST_ZERO_16(lane_taddr + acc_col + g * 16);
                        tmem_st_wait();
                    }
                    if (dbg) g_vband_dbg[0][ell][4] = clock64();
                    tc_fence_before();
                    __syncwarp();
                    // pass the token on: the next link of this item, or the first link of the CTA's next item
                    if (!last) { issue(gh_base + ((s + 1) >> 1), (s + 1) & 1, ell + 1, acc_col, 1); token_pass((s + 1) & 3); }
                    else if (more) { issue(gh_base + it.hps, 0, ell + 1, acc_col ^ (kGroupsV * 16), 0); token_pass(0); }
                }
                if (dbg) g_vband_dbg[0][ell][5] = clock64();
                // the stage of half-patch gh - 1 is free (its products completed before this link was issued, the
                // builders left it before they signalled): it takes half-patch gh + kAhead
                if (req) {
                    if (mine && lane == 0) request(rc, gh_base + (unsigned)(s >> 1) + kAhead);
                    rc.next(pp, q.Ho);
                }
                // ring rows 2 qd, 2 qd + 1 hold band rows (2 qd - 2 s) mod 8 (+1) of product s (0, 1 for the handler)
                reduce(r, gv_img, band_y0 + 2 * s + ((2 * qd - 2 * s) & 7) + (lane >> 4), x0);
                if (dbg) g_vband_dbg[0][ell][6] = clock64();
                if (last) {   // every quarter of this accumulator set has been read: the set may be reused
                    tc_fence_before();
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    tc_fence_after();
                }
            }
            gh_base += it.hps;
        }
    }

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the narrowing pass may be scheduled
    tc_fence_before();
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
}

template <typename T>
static bool launch_typed(const void *offset, const void *mask, const void *grad_out, float *gv_acc, const Geom &q,
                         cudaStream_t stream, cudaError_t *err) {
    BParams pp;
    pp.bx_rel = -3 - q.pw;                       // 16 columns centred on the taps of 8 pixels: +-3 px of offset
    pp.by_rel = -2 - q.ph;                       // 8 rows centred on the taps of 2 pixel rows: +-2 px of offset
    pp.c_w = (float)(1 - q.pw) - q.sigma - (float)pp.bx_rel;
    pp.c_h = (float)(1 - q.ph) - q.sigma - (float)pp.by_rel;
    pp.tiles_x = (q.Wo + kStripW - 1) / kStripW;
    pp.gblocks = q.G / kGroupsV;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const long long slots = 2LL * num_sms;
    // half-patches per item: an item drains (2 seg + 3) quarters for 2 seg products; fewest waves x quarters wins
    const int hp_all = (q.Ho + kHpRows - 1) / kHpRows;
    const long long per_seg = (long long)pp.tiles_x * pp.gblocks * q.N;
    long long best = -1;
    pp.seg_hp = hp_all;
    for (int seg = std::min(hp_all, 2); seg <= hp_all; ++seg) {
        const long long items = per_seg * ((hp_all + seg - 1) / seg);
        const long long cost = ((items + slots - 1) / slots) * (2 * seg + 3);
        if (best < 0 || cost <= best) { best = cost; pp.seg_hp = seg; }
    }
    pp.diag = 0;
    if (const char *e = std::getenv("DCNV3_VBAND_DIAG")) pp.diag = atoi(e);
    if (const char *e = std::getenv("DCNV3_VBAND_SEG")) pp.seg_hp = std::max(1, std::min(hp_all, atoi(e)));
    pp.segs = (hp_all + pp.seg_hp - 1) / pp.seg_hp;
    const long long total = per_seg * pp.segs;
    if (total >= (1LL << 31)) return false;
    pp.total_items = (int)total;
    CUtensorMap to, tm, tg;
    const int dtype = std::is_same<T, __half>::value ? 1 : 2;
    if (!make_hp_tensor_map(&to, offset, dtype, q.N, q.Ho, q.Wo, q.G * kP * 2, kOffPitch / 2)) return false;
    if (!make_hp_tensor_map(&tm, mask, dtype, q.N, q.Ho, q.Wo, q.G * kP, kMskPitch / 2)) return false;
    if (!make_hp_tensor_map(&tg, grad_out, dtype, q.N, q.Ho, q.Wo, q.G * q.gc, 8)) return false;
    const int ctas = (int)std::min<long long>(total, slots);
    const char *eb = std::getenv("DCNV3_VBAND_BUILD");
    const bool dense = eb && eb[0] == 'd';
    auto kern = dense ? bwd_vband<T, false> : bwd_vband<T, true>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    *err = pdl_launch(pdl_for(q), kern, dim3(ctas), dim3(kThreadsB), kSmemBytes, stream, to, tm, tg, gv_acc, q, pp);
    if (*err == cudaSuccess) *err = cudaGetLastError();
    return true;
}

}  // namespace vband

bool backward_vband_eligible(const void *offset, const void *mask, const void *grad_out, const float *gv_acc, const Geom &q) {
    if (q.gc != 16 || q.kh != 3 || q.kw != 3 || q.sh != 1 || q.sw != 1 || q.dh != 1 || q.dw != 1) return false;
    if (!(q.sigma >= 0.5f && q.sigma <= 1.25f)) return false;   // band = taps +- 2 px (rows) / +- 3 px (columns)
    // TMA staging: 16-byte aligned bases and row strides (G * 18 B for the masks: G % 8 == 0)
    if (((uintptr_t)grad_out | (uintptr_t)gv_acc | (uintptr_t)offset | (uintptr_t)mask) % 16 || q.G % 8) return false;
    if ((long long)q.N * q.Ho * q.Wo == 0) return false;
    return true;
}

// grad_value only (accumulated into the zeroed fp32 plane gv_acc); circular-band tcgen05 form.
bool try_launch_backward_vband(const void *offset, const void *mask, const void *grad_out, float *gv_acc,
                               const Geom &q, int dtype, cudaStream_t stream, cudaError_t *err) {
    if (!backward_vband_eligible(offset, mask, grad_out, gv_acc, q)) return false;
    if (dtype == 1) return vband::launch_typed<__half>(offset, mask, grad_out, gv_acc, q, stream, err);
    if (dtype == 2) return vband::launch_typed<__nv_bfloat16>(offset, mask, grad_out, gv_acc, q, stream, err);
    return false;
}

}  // namespace dcnv3

extern "C" __attribute__((visibility("default"))) int dcnv3_vband_debug_read(long long *dst) {
    return (int)cudaMemcpyFromSymbol(dst, dcnv3::vband::g_vband_dbg, sizeof(dcnv3::vband::g_vband_dbg));
}
