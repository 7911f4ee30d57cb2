// dcnv3_backward_vres.cu -- grad_value of the DCNv3 core backward for 16-bit I/O, group_channels == 16,
// 3x3 / stride 1 / dilation 1, with the accumulator of a whole band of the value map RESIDENT in tensor memory:
// no fp32 plane in HBM, no reductions, no zeroing pass, no narrowing pass.
//
// What it computes (reference dcnv3_im2col_cuda.cuh:82-147, col2im bilinear), per (image, group):
//     grad_value[cell, c] = sum over (pixel, point, corner) hitting `cell` of  w_corner * m * grad_out[pixel, c]
// i.e. the sparse product  D[cells x 16 ch] = A[cells x pixels] . G[pixels x 16 ch]  (as dcnv3_backward_vmma.cu).
//
// Why a third form (profiles/README.md, r2): in dcnv3_backward_vmma.cu the band of a strip is 16 cells wide for 8
// pixels, neighbouring strips' bands overlap and the sum of the overlapping blocks is formed in HBM -- the fp32
// plane is zeroed (105 MB at cfg2), receives 2.4x its size in vector reductions, is read back and narrowed: 480 MB
// of DRAM traffic and three passes for a 52 MB result -- and its four warps run inputs -> build -> product -> drain
// strictly one after the other (45 % of its stall samples are the CTA barrier and the mbarrier waits).  Here
//   * tensor memory holds the accumulator as 8 x 8-cell BLOCKS: tcgen05.mma with M = 64 writes row i of D to lane
//     (i / 16) * 32 + i % 16, and the D address may carry a lane offset of 16 (scripts/probes/m64_probe.cu), so one
//     16-column tile holds two independent blocks and the 512 columns hold 64 blocks = 4096 cells x 16 channels.
//     A patch of 8 x 8 pixels touches exactly the 2 x 2 blocks around it (the block grid is the patch grid shifted
//     by the band origin, -4 cells for pad 1), so the blocks of neighbouring patches coincide instead of
//     overlapping: their sums form in tensor memory.  Blocks are kept for the full width of the map and a ring of
//     R block rows (R = 5 at W = 80); a block row is final once the patch row below it is done -- it leaves through
//     tcgen05.ld as plain 32-byte stores of the 16-bit result, once;
//   * one persistent CTA per SM owns a contiguous run of the (group, image, patch row) list -- group-major, with the
//     same cuts for every group, so all groups walk the images in lock step and a pixel's offset / mask rows (all
//     groups side by side) come from DRAM once: 148 MB instead of 586 MB of DRAM reads; the block row at the head of a
//     run also needs the patch row above, which the CTA builds a second time (one patch row in ~18: 5.8 % extra
//     builds) instead of exchanging partial sums with its neighbour: every cell has ONE writer;
//   * warp-specialised, nobody executes a CTA-wide barrier inside the loop:
//       - 16 builder warps = 4 groups x 128 threads: two threads per pixel of the patch, one owns the even band rows,
//         one the odd (so each of a point's two rows has one writer and the 16-bit read-modify-writes of the pixel's
//         column of the K-major swizzled coefficient tile never race); two passes per patch -- all nine points'
//         addresses / coefficients first (independent chains), then the read-modify-writes in order;
//       - one warp requests the inputs by TMA (offsets / masks into four buffers, free again as soon as a group has
//         them in registers; grad_out = the B operand next to its coefficient tile) and refills a tile with zeros
//         (32 KB bulk copy) once its products are done;
//       - one warp issues the 16 products of a patch (M 64, N 16, K 16; the whole warp walks the schedule so that the
//         descriptors stay in uniform registers, one elected lane issues) and the commits;
//       - four drain warps, one per tensor-memory lane quarter.
//     Six slots (coefficient tile + grad_out) rotate between them through mbarriers.
// A point whose corner block leaves the 16 x 16 band of its patch (|offset| beyond ~3 px; rare for trained offsets)
// is only LISTED (a builder warp reserves its places with one atomic on the call's far-point counter); `far_points`
// adds those afterwards, one warp per point, with 16-bit reductions -- one extra rounding of the stored result for
// the cells they touch, and grad_value is reproducible only up to the order of those additions.  If more than 1 / 32
// of all points are far, the caller's fall-back (the plane form) recomputes grad_value; both decide on the device
// from the counter, no host round trip.
//
// Measured on cfg2 bf16 (profiles/README.md, r2): 152 us (ncu) against 149 + 23 (narrow) + ~8 (plane zeroing inside
// the channel-sum kernel) for the plane form; backward pass 264.4 -> 259.8 us, its DRAM traffic 816 -> 440 MB.
// What bounds it: the builders (per patch ~3000 cycles of build in a group, four groups in flight; ~850 shared-memory
// wavefronts per patch = 63 % of the LSU's cycles, the 16-bit scatter conflicting ~3.4-fold for N(0, 1) offsets in
// any operand layout).  Variants measured on the way (same file, -DVRES_TPP / -DVRES_GROUPS): one thread per pixel
// with four / five / six groups 273 / 264 / 271 us; five slots 261.7 us; refill by two dedicated warps (+7 us), by the
// builders themselves (+19 us), by st.bulk (no change); products issued by `lane == 0` instead of an elected lane of
// a warp-uniform loop: 1300 instead of 550 cycles per patch (every tcgen05.mma sat in an elect / R2UR loop).
#include "dcnv3_common.cuh"
#include "dcnv3_launch.h"
#include "dcnv3_tma.cuh"
#include "dcnv3_strip_io.cuh"
#include "dcnv3_tc.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>

namespace dcnv3 {
namespace vres {

using namespace strip;
using namespace tc;

#ifndef VRES_TPP
#define VRES_TPP 2
#endif
#ifndef VRES_GROUPS
#define VRES_GROUPS (VRES_TPP == 1 ? 5 : 4)
#endif
constexpr int kTpp = VRES_TPP;                     // builder threads per pixel (2: one owns the even band rows, one the odd)
constexpr int kSlots = 6;                          // coefficient tiles (+ the patch's grad_out, the B operand) in rotation
constexpr int kGroups = VRES_GROUPS;               // builder groups (64 kTpp threads = the pixels of a patch)
#ifndef VRES_AMN
#define VRES_AMN 0
#endif
// Layout of the coefficient tile.  false: K-major rows of 64 pixels, 128-byte swizzle, 16-bit read-modify-writes (two per
// point row, ~3.4 shared-memory wavefronts each for N(0, 1) offsets).  true: MN-major core matrices [8 cells of a band
// row][8 pixels] with the pixel's 16-byte place inside its group of eight ROTATED by twice the patch row,
// k' = (k & ~7) | ((k & 7) + 2 (k >> 3)) & 7: a horizontal corner pair that starts on an even column is one 32-bit
// word (1.5 accesses per point row), the bank group of a lane is fixed by its pixel, and the four lanes that share a
// bank group are pixels two columns apart in four different patch rows, so for small offsets they hit four different
// words (simulated for N(0, 1) offsets: 4.25 wavefronts per point row against 6.74).  The rotation is a permutation of
// K: grad_out (the B operand) is written to the same permuted places by the builders.
// MEASURED (-DVRES_AMN=1, parity-green): 18.4 M shared wavefronts instead of 23.3 M, 10.1 M conflicts instead of 16.5 M --
// and 172 us instead of 155 us (backward 263.6 against 250.9 us): 11 % more instructions (unpack / pack of word pairs, the
// second word under a divergent branch, the grad_out copy) at 66 % issue utilisation.  The builders are bound by their
// instruction stream and its latencies, not by shared-memory bandwidth; the K-major tile stays the default.
constexpr bool kAmn = VRES_AMN;
// -DVRES_ZOOB=1: the zero refill as ONE tensor-map load of a box that lies wholly outside its tensor (the hardware
// zero-fills and fetches nothing) instead of a 32 KB bulk copy from an L2-resident page of zeros.  MEASURED, parity-green:
// backward 349.6 us against 251.3 us -- the tensor path generates out-of-bounds fill far slower than it copies.
#ifndef VRES_ZOOB
#define VRES_ZOOB 0
#endif
constexpr bool kZOob = VRES_ZOOB;
// -DVRES_LOADER2=1: the loader thread keeps TWO cursors over the patch list -- input boxes and zero refills -- and serves
// whichever is ready (non-blocking polls) instead of both in program order per patch (a refill then does not wait
// behind an input buffer the builders have not released yet).  MEASURED, parity-green: backward 252.0 us against 251.2 us --
// as with the two-warp loader before it, issuing the refill earlier does not shorten the step.
#ifndef VRES_LOADER2
#define VRES_LOADER2 0
#endif
constexpr bool kLoader2 = VRES_LOADER2;
// -DVRES_TMADIAG=n (WRONG results, timing only): 1 = the grad_out boxes (128 rows of 16 bytes per patch) are loaded for
// the first patches only, 2 = the offset / mask boxes (64 rows of 48 / 32 bytes) as well: what the small-row requests cost.
#ifndef VRES_TMADIAG
#define VRES_TMADIAG 0
#endif
constexpr int kTmaDiag = VRES_TMADIAG;
// -DVRES_PIPE=1: the builders' read-modify-writes software-pipelined (the loads of point pt + 1 issued before the stores
// of point pt unless the two share a cell).  MEASURED, parity-green: backward 274.9 us against 250.8 us -- the address
// comparisons and the extra live registers cost more than the shortened chains save; off by default.
#ifndef VRES_PIPE
#define VRES_PIPE 0
#endif
constexpr bool kPipe = VRES_PIPE;
static_assert(!kAmn || VRES_TPP == 2, "the MN-major tile is implemented for two threads per pixel");
constexpr int kOmStages = kAmn ? 3 : 4;            // staged inputs in rotation: free again after the group's decode
constexpr int kGroupWarps = 2 * kTpp;
constexpr int kBuilderWarps = kGroupWarps * kGroups;
// four drain warps, warp index = 0 mod 4 first (a warp reads the tensor-memory lane quarter warp & 3); the two single-lane
// roles sit before them when the builders end on a half quartet, behind them otherwise
constexpr bool kRolesFirst = kBuilderWarps % 4 == 2;
constexpr int kLoadWarp = kRolesFirst ? kBuilderWarps : kBuilderWarps + 4, kMmaWarp = kLoadWarp + 1;
constexpr int kDrainWarp0 = kRolesFirst ? kBuilderWarps + 2 : kBuilderWarps;
constexpr int kThreads = (kBuilderWarps + 6) * 32; // 512 (kTpp 1, five groups) / 704 (kTpp 2, four groups)
static_assert(kDrainWarp0 % 4 == 0, "drain warp w must own tensor-memory lane quarter w & 3");
constexpr int kBand = 16;                          // band of a patch: 16 x 16 cells = 2 x 2 blocks
constexpr int kSubBytes = 64 * 128;                // one block's coefficients: 64 cells x 64 pixels x 2 B
constexpr int kATileBytes = 4 * kSubBytes;         // 32768
constexpr int kOffRow = 48, kMskRow = 32;          // staged bytes per pixel (36 / 18 used, 16-byte multiples)
constexpr int kStOff = 0, kStMsk = 64 * kOffRow;
constexpr int kGoutBytes = 2 * 1024;               // grad_out of a patch as [8-channel half][64 px][16 B]
constexpr int kStGout = kStMsk + 64 * kMskRow;     // (MN-major tile: grad_out is staged with the offsets / masks)
constexpr int kOmBytes = kStGout + (kAmn ? kGoutBytes : 0);   // 5120 / 7168
constexpr int kSmemBytes = 1024 + kSlots * (kATileBytes + kGoutBytes) + kOmStages * kOmBytes;   // 230400 / 231424
constexpr int kMaxRing = 5;
constexpr int kTmemCols = 512;

// Zeros for the refill of a coefficient tile: a 32 KB bulk copy from this L2-resident page (async proxy, no LSU
// instructions).  Measured alternatives: two dedicated warps storing zeros (slot turn-around 3000 -> 2000 cycles, but
// the builders slow down by as much: backward +7 us); st.bulk / UMEMSETS (+31 us in dcnv3_backward_vmma.cu); a shared ->
// shared bulk copy (illegal instruction without a cluster launch).
__device__ __align__(128) unsigned char g_zero_tile[kATileBytes];
__device__ __forceinline__ void bulk_fill(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// development only (-DVRES_HANG_DEBUG): a waiter that has polled ~2^20 times reports (role, barrier tag, patch) into host-mapped memory set by
// dcnv3_vres_debug_hang(ptr) -- a hung pipeline can then be read from the host while the kernel still spins
__device__ unsigned long long *g_vres_hang;
__device__ __noinline__ void report_hang(int tag, unsigned p, unsigned parity) {
    unsigned long long *h = g_vres_hang;
    if (!h || blockIdx.x >= 2) return;
    const unsigned slot = atomicAdd((unsigned *)h, 1u);
    if (slot < 60) h[1 + slot] = ((unsigned long long)blockIdx.x << 48) | ((unsigned long long)(threadIdx.x >> 5) << 40) |
                                 ((unsigned long long)tag << 32) | ((unsigned long long)parity << 31) | p;
    __threadfence_system();
}

// a wait that yields its issue slots while it polls (roles that wait for long: drain, refill, builders out of slots)
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, unsigned parity, unsigned ns, int tag = 0, unsigned p = 0) {
    for (unsigned it = 0;; ++it) {
#ifdef VRES_HANG_DEBUG
        if (it == (1u << 20)) report_hang(tag, p, parity);
#endif
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        if (ns) __nanosleep(ns);
    }
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// development only (DCNV3_VRES_DIAG & 1): cycle stamps of CTA 0, per patch -- builder [0..2], loader [3..4], products [5..7]
__device__ long long g_vres_dbg[256][8];

struct RParams {
    int diag;
    int bx_rel, by_rel;      // origin of the block grid relative to the patch grid (cells; -4 for pad 1, sigma 1)
    int S, PR;               // strips per patch row, patch rows per image
    int nbxp, ring;          // blocks per block row (even), block rows resident
    int total_rows;          // N * G * PR
    int cpg;                 // CTAs per group (0: one flat split of the list)
    unsigned long long far_cap;   // entries the far-point list can take (= the fall-back threshold)
};

// The CTA's list of patch rows.  Every role walks the same list, so event numbers (= owned block rows in the order
// of their first product) agree without communication.
struct Row {
    int n, g, i;
    bool dup;                // the patch row above the CTA's run: built again for the block row below it only
    bool up_own, down_own;   // does this CTA own block row i / i + 1
    bool up_fresh;           // block row i gets its first product from this patch row
    int ev_up, ev_down;
};
// The list is GROUP-major, f = (g * N + n) * PR + i: every group's share of the list is cut into CTA runs the same way, so at
// any moment all groups work on the same few images and the 576-byte offset / 288-byte mask rows of a pixel (all groups
// side by side) are fetched from DRAM once, not once per group (image-major order: 598 MB of DRAM reads for 264 MB).
struct Sched {
    int cur, hi, PR, N, ev, last_down;
    bool pending_dup;
    __device__ Sched(int lo, int hi_, int PR_, int N_) : cur(lo), hi(hi_), PR(PR_), N(N_), ev(0), last_down(-1),
                                                          pending_dup(lo < hi_ && lo % PR_ != 0) {}
    __device__ bool next(Row &r) {
        if (pending_dup) {
            pending_dup = false;
            const int f = cur - 1, ng = f / PR;
            r.i = f - ng * PR; r.n = ng % N; r.g = ng / N;
            r.dup = true; r.up_own = false; r.down_own = true; r.up_fresh = false;
            r.ev_up = -1; r.ev_down = ev++; last_down = r.ev_down;
            return true;
        }
        if (cur >= hi) return false;
        const int f = cur++, ng = f / PR;
        r.i = f - ng * PR; r.n = ng % N; r.g = ng / N;
        r.dup = false; r.up_own = true; r.up_fresh = r.i == 0;
        r.down_own = r.i == PR - 1 || cur < hi;
        r.ev_up = r.up_fresh ? ev++ : last_down;
        if (r.down_own) { r.ev_down = ev++; last_down = r.ev_down; } else r.ev_down = -1;
        return true;
    }
};

template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
bwd_vres(const __grid_constant__ CUtensorMap tmap_off, const __grid_constant__ CUtensorMap tmap_msk,
         const __grid_constant__ CUtensorMap tmap_gout, const __grid_constant__ CUtensorMap tmap_zero, T *__restrict__ grad_value,
         uint32_t *__restrict__ far_list, unsigned long long *__restrict__ far_count, const Geom q, const RParams pp) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t a_ready[kSlots], a_full[kSlots], a_done[kSlots], om_full[kOmStages], om_free[kOmStages],
        row_done[kMaxRing], acc_free[kMaxRing];
    __shared__ uint32_t tmem_base_s;
    // which patch a staged offsets / masks buffer currently belongs to.  There are fewer buffers than builder groups, and
    // a group may run more than kOmStages patches ahead of another: a parity wait alone would then be satisfied by the
    // phase before last (measured: a dead-locked pipeline).  A builder first waits for its patch number to appear here.
    __shared__ volatile uint32_t om_seq[kOmStages];

    // (the warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the product warp's
    // descriptors in uniform registers -- with `tid >> 5` every tcgen05.mma was wrapped in an elect / R2UR loop,
    // ~80 cycles per product)
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t slot0 = smem_u32(base), gout0 = slot0 + kSlots * kATileBytes, om0 = gout0 + kSlots * kGoutBytes;

    // this CTA's run of the flattened patch-row list
    // (pp.cpg CTAs per group, every group's rows cut the same way: all groups walk the images in lock step)
    int lo, hi;
    if (pp.cpg > 0) {
        const int g = blockIdx.x % q.G, c = blockIdx.x / q.G, np = q.N * pp.PR;
        lo = g * np + (int)((long long)np * c / pp.cpg);
        hi = g * np + (int)((long long)np * (c + 1) / pp.cpg);
    } else {
        lo = (int)((long long)pp.total_rows * blockIdx.x / gridDim.x);
        hi = (int)((long long)pp.total_rows * (blockIdx.x + 1) / gridDim.x);
    }

    if (tid == 0) {
        for (int i = 0; i < kSlots; ++i) { mbar_init(&a_ready[i], 1); mbar_init(&a_full[i], 64 * kTpp); mbar_init(&a_done[i], 1); }
        for (int i = 0; i < kOmStages; ++i) { mbar_init(&om_full[i], 1); mbar_init(&om_free[i], kGroupWarps); om_seq[i] = 0xffffffffu; }
        for (int i = 0; i < kMaxRing; ++i) { mbar_init(&row_done[i], 1); mbar_init(&acc_free[i], 4); }
        fence_barrier_init();
        prefetch_tensormap(&tmap_off);
        prefetch_tensormap(&tmap_msk);
        prefetch_tensormap(&tmap_gout);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    // Programmatic dependent launch: the prologue above runs under the channel-sum kernel's last wave; the far-point
    // counter is zeroed by that kernel, so everything below waits for it.  (Waiting only at the end -- the pipeline
    // overlapping the channel sums' tail -- was measured: no difference, a CTA of this kernel needs a whole SM.)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int C = q.G * q.gc;

    if (warp < kBuilderWarps) {
        // ================================================================== builders: thread <-> pixel of the patch
        // (kTpp == 2: two threads per pixel; `par` owns the band rows of that parity, so each of a point's two rows has
        // exactly one writer and the 16-bit read-modify-writes of a column never race)
        const int grp = warp / kGroupWarps, wg = warp % kGroupWarps, hw = wg & 1, par = wg >> 1;
        const int k = hw * 32 + lane;
        const int px_x = lane & 7, px_y = hw * kPatchH + (lane >> 3);
        const uint32_t kc = (uint32_t)k >> 3, kl = ((uint32_t)k & 7u) * 2u;
        Sched sch(lo, hi, pp.PR, q.N);
        Row row;
        unsigned p = 0;
        while (sch.next(row)) {
            const int ho = row.i * 8 + px_y;
            const float bh = axis_base(ho, 3, 1, q.ph, 1, q.sigma) - (float)(row.i * 8 + pp.by_rel);
            for (int j = 0; j < pp.S; ++j, ++p) {
                if ((int)(p % kGroups) != grp) continue;
                const unsigned slot = p % kSlots, stage = p % kOmStages;
                // (MN-major: the pixel's 16-byte place -- eight cells of a band row -- inside its group of eight pixels)
                const uint32_t kp = ((uint32_t)k & ~7u) | ((((uint32_t)k & 7u) + 2u * kc) & 7u);
                const uint32_t a_thr = kAmn ? slot0 + slot * kATileBytes + kp * 16u : slot0 + slot * kATileBytes + kl;
                const uint32_t sa = om0 + stage * kOmBytes;
                const bool dbg = (pp.diag & 1) && blockIdx.x == 0 && wg == 0 && lane == 0 && p < 256u;
                if (dbg) g_vres_dbg[p][0] = clock64();
                if (kOmStages < kGroups) while (om_seq[stage] != p) __nanosleep(32);
                mbar_wait_sleep(&om_full[stage], (p / kOmStages) & 1u, (pp.diag & 32) ? 0 : 32, 1, p);
                if (dbg) g_vres_dbg[p][3] = clock64();
                const int wo = j * 8 + px_x;
                const bool live = wo < q.Wo && ho < q.Ho;
                unsigned far = 0;
                uint32_t off[kP] = {}, mw[5] = {};
                if (live) {
                    const uint4 o0 = lds128(sa + kStOff + k * kOffRow), o1 = lds128(sa + kStOff + k * kOffRow + 16),
                                o2 = lds128(sa + kStOff + k * kOffRow + 32);
                    const uint32_t w[12] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w, o2.x, o2.y, o2.z, o2.w};
                    const unsigned osh = ((unsigned)(row.g * kP * 4) & 15u) >> 2;     // warp-uniform: 0..3 words
#pragma unroll
                    for (int pt = 0; pt < kP; ++pt) off[pt] = osh == 0 ? w[pt] : osh == 1 ? w[pt + 1] : osh == 2 ? w[pt + 2] : w[pt + 3];
                    const uint4 m0 = lds128(sa + kStMsk + k * kMskRow), m1 = lds128(sa + kStMsk + k * kMskRow + 16);
                    const uint32_t v[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
                    const unsigned esh = ((unsigned)(row.g * kP * 2) & 15u) >> 1;     // warp-uniform: 0..7 elements
                    uint32_t u[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) u[t] = (esh & 1u) ? __funnelshift_r(v[t], t + 1 < 8 ? v[t + 1] : 0u, 16) : v[t];
                    const unsigned ws = esh >> 1;
#pragma unroll
                    for (int t = 0; t < 5; ++t)
                        mw[t] = ws == 0 ? u[t] : ws == 1 ? u[t + 1] : ws == 2 ? u[t + 2] : (t + 3 < 8 ? u[t + 3] : 0u);
                }
                uint4 gq0 = make_uint4(0u, 0u, 0u, 0u), gq1 = gq0;
                if (kAmn && par == 0) {   // this pixel's grad_out: to the B operand's permuted place once the slot is free
                    gq0 = lds128(sa + kStGout + k * 16);
                    gq1 = lds128(sa + kStGout + 1024 + k * 16);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&om_free[stage]);                 // the staged inputs are in registers
                mbar_wait_sleep(&a_ready[slot], (p / kSlots) & 1u, (pp.diag & 32) ? 0 : 32, 2, p);      // the tile is zero again
                if (dbg) g_vres_dbg[p][1] = clock64();
                if (kAmn && par == 0) {
                    sts128(gout0 + slot * kGoutBytes + kp * 16u, gq0);
                    sts128(gout0 + slot * kGoutBytes + 1024u + kp * 16u, gq1);
                }
                if (live) {
                    const float bw = axis_base(wo, 3, 1, q.pw, 1, q.sigma) - (float)(j * 8 + pp.bx_rel);
                    if constexpr (kTpp == 2) {
                        // Two passes.  First every point's two word addresses and coefficients, branch-free (nine
                        // independent chains the scheduler can interleave); then the read-modify-writes, which must stay
                        // in order (two points of a pixel may hit the same cell).
                        const uint32_t kcs = kc << 4;
                        uint32_t ea[kP], eb[kP];
                        float wa[kP], wb[kP];
#pragma unroll
                        for (int pt = 0; pt < kP; ++pt) {
                            const float2 d = unpack2(off[pt], T());
                            const float m = f32_of((uint16_t)(mw[pt >> 1] >> (16 * (pt & 1))), T());
                            const float ub = bw + ((float)(pt / 3) + d.x) * q.sigma;
                            const float vb = bh + ((float)(pt % 3) + d.y) * q.sigma;
                            const float fw = floorf(ub), fh = floorf(vb);
                            const float lw = ub - fw, lh = vb - fh;
                            // 0 <= x < limit on the float's bit pattern: negative values and NaN compare as large unsigned
                            const bool inb = __float_as_uint(ub) < __float_as_uint((float)(kBand - 1)) &&
                                             __float_as_uint(vb) < __float_as_uint((float)(kBand - 1));
                            far |= inb ? 0u : 1u << pt;
                            const uint32_t cx = (uint32_t)(int)fw, ry = (uint32_t)(int)fh, cx1 = cx + 1u;
                            // this thread's row of the point: the one whose parity it owns
                            const bool top = (ry & 1u) == (uint32_t)par;
                            const uint32_t rr = top ? ry : ry + 1u;
                            // a block row this CTA does not own is never multiplied
                            const bool act = inb && (rr < 8u ? row.up_own : row.down_own);
                            const float vm = (top ? 1.f - lh : lh) * m;
                            // cell (rr, cx) -> sub-tile (rr >> 3, cx >> 3), row (rr & 7) * 8 + (cx & 7), chunk kc ^ (cx & 7):
                            // (cx & 7) * 0x90 = row * 128 | (cx & 7) * 16, the chunk's xor touches bits 4..6 only
                            const uint32_t r0 = a_thr + ((rr & 8u) << 11) + ((rr & 7u) << 10);
                            if constexpr (kAmn) {
                                // word of cell c in the row: sub-tile (c & 8) << 10, 32-bit word (c & 6) << 1 of the pixel's place;
                                // an even column: both cells in ONE word (eb = 0), an odd one: high half, then the next word's low half
                                ea[pt] = act ? r0 + ((cx & 8u) << 10) + ((cx & 6u) << 1) : 0u;
                                eb[pt] = (cx & 1u) ? r0 + ((cx1 & 8u) << 10) + ((cx1 & 6u) << 1) : 0u;
                            } else {
                                ea[pt] = act ? r0 + ((((cx & 7u) * 0x90u) ^ kcs) | ((cx & 8u) << 10)) : 0u;
                                eb[pt] = r0 + ((((cx1 & 7u) * 0x90u) ^ kcs) | ((cx1 & 8u) << 10));
                            }
                            wa[pt] = vm * (1.f - lw);
                            wb[pt] = vm * lw;
                        }
                        if constexpr (!kAmn && kPipe) {
                            // The read-modify-writes, software-pipelined: the loads of point pt + 1 are issued BEFORE the stores
                            // of point pt unless the two points share a cell (then they wait for the stores) -- a column's
                            // nine read-modify-write pairs are otherwise one serial chain of shared-memory round trips.
                            uint32_t la = 0, lb = 0;
                            if (ea[0]) { la = lds16(ea[0]); lb = lds16(eb[0]); }
#pragma unroll
                            for (int pt = 0; pt < kP; ++pt) {
                                uint32_t na = 0, nb = 0;
                                bool late = false;
                                if (pt + 1 < kP && ea[pt + 1]) {
                                    late = ea[pt] && (ea[pt + 1] == ea[pt] || ea[pt + 1] == eb[pt] || eb[pt + 1] == ea[pt] || eb[pt + 1] == eb[pt]);
                                    if (!late) { na = lds16(ea[pt + 1]); nb = lds16(eb[pt + 1]); }
                                }
                                if (ea[pt]) {
                                    sts16(ea[pt], bits16(f32_of((uint16_t)la, T()) + wa[pt], T()));
                                    sts16(eb[pt], bits16(f32_of((uint16_t)lb, T()) + wb[pt], T()));
                                }
                                if (late) { na = lds16(ea[pt + 1]); nb = lds16(eb[pt + 1]); }
                                la = na; lb = nb;
                            }
                        } else {
#pragma unroll
                        for (int pt = 0; pt < kP; ++pt)
                            if (ea[pt]) {
                                if constexpr (kAmn) {
                                    const bool odd = eb[pt] != 0u;
                                    float2 f = unpack2(lds32(ea[pt]), T());
                                    f.x += odd ? 0.f : wa[pt];
                                    f.y += odd ? wa[pt] : wb[pt];
                                    sts32(ea[pt], pack2(f.x, f.y, T()));
                                    if (odd) {   // the pair straddles two words: cell cx + 1 is the next word's low half
                                        float2 h = unpack2(lds32(eb[pt]), T());
                                        h.x += wb[pt];
                                        sts32(eb[pt], pack2(h.x, h.y, T()));
                                    }
                                } else {
                                    const float a0 = f32_of((uint16_t)lds16(ea[pt]), T()), a1 = f32_of((uint16_t)lds16(eb[pt]), T());
                                    sts16(ea[pt], bits16(a0 + wa[pt], T()));
                                    sts16(eb[pt], bits16(a1 + wb[pt], T()));
                                }
                            }
                        }
                    } else {
                        // one thread per pixel: batches of three points -- first their four word addresses and
                        // coefficients (independent chains), then the read-modify-writes in order
                        const uint32_t kcs = kc << 4;
#pragma unroll
                        for (int b3 = 0; b3 < kP; b3 += 3) {
                            uint32_t e0[3], e1[3], e2[3], e3[3];
                            float w0[3], w1[3], w2[3], w3[3];
#pragma unroll
                            for (int u = 0; u < 3; ++u) {
                                const int pt = b3 + u;
                                const float2 d = unpack2(off[pt], T());
                                const float m = f32_of((uint16_t)(mw[pt >> 1] >> (16 * (pt & 1))), T());
                                const float ub = bw + ((float)(pt / 3) + d.x) * q.sigma;
                                const float vb = bh + ((float)(pt % 3) + d.y) * q.sigma;
                                const float fw = floorf(ub), fh = floorf(vb);
                                const float lw = ub - fw, lh = vb - fh;
                                const bool inb = __float_as_uint(ub) < __float_as_uint((float)(kBand - 1)) &&
                                                 __float_as_uint(vb) < __float_as_uint((float)(kBand - 1));
                                far |= inb ? 0u : 1u << pt;
                                const uint32_t cx = (uint32_t)(int)fw, ry = (uint32_t)(int)fh, cx1 = cx + 1u, ry1 = ry + 1u;
                                // a block row this CTA does not own is never multiplied: skip points entirely inside it
                                const bool act = inb && !((!row.up_own && ry < 7u) || (!row.down_own && ry >= 8u));
                                const float hm = (1.f - lh) * m, lm = lh * m, hwt = 1.f - lw;
                                const uint32_t r0 = a_thr + ((ry & 8u) << 11) + ((ry & 7u) << 10);
                                const uint32_t r1 = a_thr + ((ry1 & 8u) << 11) + ((ry1 & 7u) << 10);
                                const uint32_t c0 = (((cx & 7u) * 0x90u) ^ kcs) | ((cx & 8u) << 10);
                                const uint32_t c1 = (((cx1 & 7u) * 0x90u) ^ kcs) | ((cx1 & 8u) << 10);
                                e0[u] = act ? r0 + c0 : 0u; e1[u] = r0 + c1; e2[u] = r1 + c0; e3[u] = r1 + c1;
                                w0[u] = hm * hwt; w1[u] = hm * lw; w2[u] = lm * hwt; w3[u] = lm * lw;
                            }
#pragma unroll
                            for (int u = 0; u < 3; ++u)
                                if (e0[u]) {
                                    const float a0 = f32_of((uint16_t)lds16(e0[u]), T()), a1 = f32_of((uint16_t)lds16(e1[u]), T());
                                    const float a2 = f32_of((uint16_t)lds16(e2[u]), T()), a3 = f32_of((uint16_t)lds16(e3[u]), T());
                                    sts16(e0[u], bits16(a0 + w0[u], T()));
                                    sts16(e1[u], bits16(a1 + w1[u], T()));
                                    sts16(e2[u], bits16(a2 + w2[u], T()));
                                    sts16(e3[u], bits16(a3 + w3[u], T()));
                                }
                        }
                    }
                }
                if (dbg) g_vres_dbg[p][2] = clock64();
                fence_proxy_async();
                mbar_arrive(&a_full[slot]);
                if (!row.dup && par == 0 && __ballot_sync(0xffffffffu, far != 0u)) {
                    // rare: the warp appends its far points to the list (one atomic per warp for the positions)
                    const int nf = __popc(far);
                    int incl = nf;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += t;
                    }
                    const int total = __shfl_sync(0xffffffffu, incl, 31);
                    unsigned long long pos = 0;
                    if (lane == 0) pos = atomicAdd(far_count, (unsigned long long)total);
                    pos = __shfl_sync(0xffffffffu, pos, 0) + (unsigned long long)(incl - nf);
                    const uint32_t pg = (uint32_t)((((size_t)row.n * q.Ho + ho) * q.Wo + wo) * q.G + row.g);
                    for (unsigned f = far; f; f &= f - 1u, ++pos)
                        if (pos < pp.far_cap) far_list[pos] = (pg << 4) | (uint32_t)(__ffs(f) - 1);
                }
            }
        }
    } else if (warp >= kDrainWarp0 && warp < kDrainWarp0 + 4) {
        // ================================================================== drain: one tensor-memory lane quarter each
        const int wq = warp & 3;
        Sched sch(lo, hi, pp.PR, q.N);
        Row row;
        auto drain = [&](int ev, int r, int n, int g) {
            const int sr = ev % pp.ring;
            mbar_wait_sleep(&row_done[sr], (unsigned)(ev / pp.ring) & 1u, 256, 5, (unsigned)ev);
            tc_fence_after();
            const int y = r * 8 + pp.by_rel + 2 * wq + ((lane >> 3) & 1);
            const bool oky = (unsigned)y < (unsigned)q.H;
            T *rowp = grad_value + ((size_t)n * q.H + (oky ? y : 0)) * q.W * C + g * kCh;
            const int half = lane >> 4, cx = lane & 7;
            const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(sr * (pp.nbxp >> 1) * 16);
            for (int t = 0; t < (pp.nbxp >> 1); ++t) {
                float v[16];
                VMMA_TMEM_LD_16(taddr + (uint32_t)(t * 16), v);
                tmem_ld_wait();
                const int x = (2 * t + half) * 8 + pp.bx_rel + cx;
                if (oky && (unsigned)x < (unsigned)q.W) {
                    uint4 *dst = reinterpret_cast<uint4 *>(rowp + (size_t)x * C);
                    dst[0] = pack<T>(v);
                    dst[1] = pack<T>(v + 8);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free[sr]);
        };
        while (sch.next(row)) {
            if (row.up_own) drain(row.ev_up, row.i, row.n, row.g);
            if (row.i == pp.PR - 1 && row.down_own) drain(row.ev_down, pp.PR, row.n, row.g);
        }
    } else if (warp == kLoadWarp) {
        // ================================================================== refill + inputs of the next patches
        if (kLoader2 && lane == 0) {
            auto try_wait = [](uint64_t *bar, unsigned parity) {
                uint32_t ok;
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                             : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
                return ok != 0;
            };
            Sched sa(lo, hi, pp.PR, q.N), sb(lo, hi, pp.PR, q.N);
            Row ra, rb;
            bool live_a = sa.next(ra), live_b = sb.next(rb);
            int ja = 0, jb = 0;
            unsigned pa = 0, pb = 0;
            while (live_a || live_b) {
                bool progress = false;
                if (live_a) {      // offsets / masks of patch pa: the buffer is free once patch pa - kOmStages has been decoded
                    const unsigned stage = pa % kOmStages;
                    if (pa < (unsigned)kOmStages || try_wait(&om_free[stage], (pa / kOmStages - 1u) & 1u)) {
                        unsigned char *st = base + kSlots * (kATileBytes + kGoutBytes) + stage * kOmBytes;
                        uint64_t *bar = &om_full[stage];
                        om_seq[stage] = pa;
                        mbar_expect_tx(bar, kOmBytes);
                        tma_load_4d(st + kStOff, &tmap_off, bar, (ra.g * kP * 4 & ~15) >> 1, ja * 8, ra.i * 8, ra.n);
                        tma_load_4d(st + kStMsk, &tmap_msk, bar, (ra.g * kP * 2 & ~15) >> 1, ja * 8, ra.i * 8, ra.n);
                        ++pa;
                        if (++ja >= pp.S) { ja = 0; live_a = sa.next(ra); }
                        progress = true;
                    }
                }
                if (live_b) {      // zeros + grad_out of patch pb: its slot is free once patch pb - kSlots is through its products
                    const unsigned slot = pb % kSlots;
                    if (pb < (unsigned)kSlots || try_wait(&a_done[slot], (pb / kSlots - 1u) & 1u)) {
                        uint64_t *bar = &a_ready[slot];
                        mbar_expect_tx(bar, kATileBytes + kGoutBytes);
                        bulk_fill(slot0 + slot * kATileBytes, g_zero_tile, kATileBytes, bar);
                        unsigned char *gs = base + kSlots * kATileBytes + slot * kGoutBytes;
                        tma_load_4d(gs, &tmap_gout, bar, rb.g * kCh, jb * 8, rb.i * 8, rb.n);
                        tma_load_4d(gs + 1024, &tmap_gout, bar, rb.g * kCh + 8, jb * 8, rb.i * 8, rb.n);
                        ++pb;
                        if (++jb >= pp.S) { jb = 0; live_b = sb.next(rb); }
                        progress = true;
                    }
                }
                if (!progress) __nanosleep(32);
            }
        } else if (lane == 0) {
            Sched sch(lo, hi, pp.PR, q.N);
            Row row;
            unsigned p = 0;
            while (sch.next(row)) {
                for (int j = 0; j < pp.S; ++j, ++p) {
                    const unsigned slot = p % kSlots, stage = p % kOmStages;
                    const bool dbg = (pp.diag & 1) && blockIdx.x == 0 && p < 256u;
                    // offsets / masks: the buffer is free once the group of patch p - kOmStages has decoded them
                    if (p >= (unsigned)kOmStages) mbar_wait_sleep(&om_free[stage], (p / kOmStages - 1u) & 1u, (pp.diag & 64) ? 0 : 32, 3, p);
                    {
                        unsigned char *st = base + kSlots * (kATileBytes + kGoutBytes) + stage * kOmBytes;
                        uint64_t *bar = &om_full[stage];
                        om_seq[stage] = p;
                        const bool skip_om = kTmaDiag >= 2 && p >= (unsigned)kOmStages;
                        mbar_expect_tx(bar, skip_om ? 0u : kOmBytes);
                        // a box starts on a 16-byte boundary of the row: the group's run begins 0..3 words / 0..7 elements in
                        if (!skip_om) {
                        tma_load_4d(st + kStOff, &tmap_off, bar, (row.g * kP * 4 & ~15) >> 1, j * 8, row.i * 8, row.n);
                        tma_load_4d(st + kStMsk, &tmap_msk, bar, (row.g * kP * 2 & ~15) >> 1, j * 8, row.i * 8, row.n);
                        }
                        if (kAmn) {
                            tma_load_4d(st + kStGout, &tmap_gout, bar, row.g * kCh, j * 8, row.i * 8, row.n);
                            tma_load_4d(st + kStGout + 1024, &tmap_gout, bar, row.g * kCh + 8, j * 8, row.i * 8, row.n);
                        }
                    }
                    // patch p - kSlots is through its products: its tile is zeroed for patch p and its grad_out replaced
                    if (p >= (unsigned)kSlots) mbar_wait_sleep(&a_done[slot], (p / kSlots - 1u) & 1u, (pp.diag & 64) ? 0 : 32, 4, p);
                    if (dbg) g_vres_dbg[p][4] = clock64();
                    uint64_t *bar = &a_ready[slot];
                    const bool skip_g = kTmaDiag >= 1 && p >= (unsigned)kSlots;
                    mbar_expect_tx(bar, kATileBytes + ((kAmn || skip_g) ? 0 : kGoutBytes));
                    if (kZOob) {
                        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                     ::"r"(slot0 + slot * kATileBytes), "l"(&tmap_zero), "r"(0), "r"(1 << 20), "r"(smem_u32(bar)) : "memory");
                    } else {
                        bulk_fill(slot0 + slot * kATileBytes, g_zero_tile, kATileBytes, bar);
                    }
                    if (!kAmn && !skip_g) {
                        unsigned char *gs = base + kSlots * kATileBytes + slot * kGoutBytes;
                        tma_load_4d(gs, &tmap_gout, bar, row.g * kCh, j * 8, row.i * 8, row.n);
                        tma_load_4d(gs + 1024, &tmap_gout, bar, row.g * kCh + 8, j * 8, row.i * 8, row.n);
                    }
                }
            }
        }
    } else {
        // ================================================================== products (the whole warp walks the list;
        // one elected lane issues)
        {
            const uint32_t idesc = umma_idesc(std::is_same<T, __nv_bfloat16>::value ? 1 : 0, 64, kCh) | (kAmn ? 1u << 15 : 0u);
            // descriptors of slot 0 / block 0 / K step 0; the others differ in the 14-bit address field only
            const uint64_t adesc0 = kAmn ? umma_desc_mn_plain(slot0, 128, 1024) : umma_desc_k_sw128(slot0), bdesc0 = umma_desc_mn_plain(gout0, 128, 1024);
            Sched sch(lo, hi, pp.PR, q.N);
            Row row;
            unsigned p = 0;
            while (sch.next(row)) {
                // a block row that starts here takes over a ring place: its previous tenant must have been drained
                if (row.up_own && row.up_fresh && row.ev_up >= pp.ring)
                    mbar_wait(&acc_free[row.ev_up % pp.ring], (unsigned)(row.ev_up / pp.ring - 1) & 1u);
                if (row.down_own && row.ev_down >= pp.ring)
                    mbar_wait(&acc_free[row.ev_down % pp.ring], (unsigned)(row.ev_down / pp.ring - 1) & 1u);
                tc_fence_after();
                const int base_up = row.up_own ? (row.ev_up % pp.ring) * pp.nbxp : 0;
                const int base_down = row.down_own ? (row.ev_down % pp.ring) * pp.nbxp : 0;
                for (int j = 0; j < pp.S; ++j, ++p) {
                    const unsigned slot = p % kSlots;
                    const uint64_t ad = adesc0 + (uint64_t)((slot * kATileBytes) >> 4), bd = bdesc0 + (uint64_t)((slot * kGoutBytes) >> 4);
                    const bool dbg = (pp.diag & 1) && blockIdx.x == 0 && p < 256u;
                    if (dbg && lane == 0) g_vres_dbg[p][5] = clock64();
                    mbar_wait(&a_full[slot], (p / kSlots) & 1u);
                    if (dbg && lane == 0) g_vres_dbg[p][6] = clock64();
                    tc_fence_after();
                    // K-step-major: consecutive products go to DIFFERENT accumulator blocks.  Products into the same
                    // block form a dependent chain that advances one product per pipeline latency (~80 cycles measured
                    // for these N = 16 shapes); the four blocks' chains interleave.
                    uint32_t dd[4];
                    bool first[4], own[4];
#pragma unroll
                    for (int sub = 0; sub < 4; ++sub) {
                        const bool down = sub >> 1;
                        own[sub] = down ? row.down_own : row.up_own;
                        const int blk = (down ? base_down : base_up) + j + (sub & 1);
                        dd[sub] = tmem_base + ((uint32_t)(blk & 1) << 20) + ((uint32_t)(blk >> 1) << 4);
                        first[sub] = (down || row.up_fresh) && ((sub & 1) || j == 0);
                    }
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)   // K step = 16 pixels: 32 B of an A row, 256 B of the pixel-major B
#pragma unroll
                            for (int sub = 0; sub < 4; ++sub)
                                if (own[sub])
                                    tc_mma(dd[sub], ad + (uint64_t)((sub * kSubBytes + ks * (kAmn ? 256 : 32)) >> 4), bd + (uint64_t)((ks * 256) >> 4), idesc,
                                           (uint32_t)(ks > 0 || !first[sub]));
                        tc_commit(&a_done[slot]);
                        if (j == pp.S - 1) {
                            if (row.up_own) tc_commit(&row_done[row.ev_up % pp.ring]);
                            if (row.i == pp.PR - 1 && row.down_own) tc_commit(&row_done[row.ev_down % pp.ring]);
                        }
                    }
                    __syncwarp();
                    if (dbg && lane == 0) g_vres_dbg[p][7] = clock64();
                }
            }
        }
    }

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    tc_fence_before();
    __syncthreads();   // the drain warps have seen the last commit: every product and refill has completed
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// Points that left the band of their patch: thread <-> (pixel, group); almost every thread reads a zero mask and exits.
// 32-bit reduction of two 16-bit values
__device__ __forceinline__ void red_add2(__nv_bfloat16 *p, uint32_t v) { asm volatile("red.global.add.noftz.bf16x2 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_add2(__half *p, uint32_t v) { asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// After the product kernel: one WARP per listed far point -- lane = (corner, channel pair): every lane computes the
// point's location (broadcast loads of its offset / mask), takes its two grad_out channels and adds its corner's
// share to the stored 16-bit result (a corner = one 32-byte sector of eight 32-bit reductions).
// If the call produced more far points than `thr`, nothing is added here: the caller's fall-back (the plane form,
// launched behind this kernel on the same condition) recomputes grad_value, and this kernel zeroes its fp32 plane.
// (A first form scanned one 16-bit mask per pixel and group -- 3.3 MB -- and let single threads walk their points:
// 19 us for ~3500 points, most of it the serial tail of the few threads that had any.)
template <typename T>
__global__ void __launch_bounds__(256)
far_points(const uint32_t *far_list, const unsigned long long *far_count, unsigned long long thr,
           const T *__restrict__ offset, const T *__restrict__ mask, const T *__restrict__ grad_out, T *__restrict__ grad_value,
           float4 *__restrict__ plane, size_t plane_vec, const Geom q) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // (volatile asm loads: with plain loads through `const __restrict__` pointers the compiler hoisted the counter's
    // LDG.CONSTANT ABOVE the grid-dependency wait -- the kernel then saw the count before the product kernel had
    // finished adding to it, and took the sparse path while the caller's fall-back kernels saw the final count)
    unsigned long long count;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(count) : "l"(far_count) : "memory");
    const size_t tid = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (count > thr) {
        for (size_t i = tid; i < plane_vec; i += (size_t)gridDim.x * 256) plane[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    // Two phases per batch of 32 entries.  Phase 1, lane <-> entry: the entry, its offset pair and its mask -- a chain of
    // dependent loads that mostly miss L2 -- are fetched for 32 points at once.  Phase 2, lane <-> (corner, channel pair):
    // the points' locations are broadcast by shuffles, four points at a time (their grad_out loads are issued together,
    // then their reductions).  (One warp per point, its loads one after the other: 125 us for the ~330 k far points of
    // N(0, 2)-pixel offsets at cfg2; this form 50 us (ncu): backward 491 -> 366 us there, 440 -> 322 us for U(-4, 4).  Issuing ALL grad_out loads of a batch
    // before its reductions instead of four at a time: 106 registers, no gain there and +3 us on the usual short list.)
    const int lane = threadIdx.x & 31, corner = lane >> 3, cp = lane & 7;
    const int C = q.G * q.gc;
    const size_t n_warps = ((size_t)gridDim.x * 256) >> 5;
    // entries per warp and batch: as few as spread the list over all warps (a few thousand far points -- the usual case --
    // are one point per warp, all in flight at once: 6 us; with 32 per warp the same list took 19 us on 110 warps)
    const unsigned bsz = (unsigned)((count + n_warps - 1) / n_warps < 32 ? (count + n_warps - 1) / n_warps : 32);
    for (size_t base = (tid >> 5) * bsz; base < count; base += n_warps * bsz) {
        const size_t i = base + lane;
        float lw_abs = 0.f, lh_abs = 0.f, m = 0.f;
        uint32_t pix32 = 0;
        int g = 0;
        bool ok = (unsigned)lane < bsz && i < count;
        if (ok) {
            uint32_t e;
            asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(e) : "l"(far_list + i) : "memory");
            const int pt = (int)(e & 15u);
            const size_t pg = e >> 4;
            g = (int)(pg % q.G);
            const size_t pix = pg / q.G;                   // pixel (n, ho, wo); < 2^28 (plan())
            pix32 = (uint32_t)pix;
            const int wo = (int)(pix % q.Wo), ho = (int)((pix / q.Wo) % q.Ho);
            const float2 d = load_pair(offset + (pix * q.G + g) * (kP * 2) + 2 * pt);
            m = to_f32(mask[(pix * q.G + g) * kP + pt]);
            lw_abs = axis_base(wo, 3, 1, q.pw, 1, q.sigma) + ((float)(pt / 3) + d.x) * q.sigma;
            lh_abs = axis_base(ho, 3, 1, q.ph, 1, q.sigma) + ((float)(pt % 3) + d.y) * q.sigma;
            ok = lh_abs > -1.f && lw_abs > -1.f && lh_abs < (float)q.H && lw_abs < (float)q.W;   // (:262-263)
        }
        unsigned todo = __ballot_sync(0xffffffffu, ok);
        while (todo) {
            const T *gsrc[4];
            T *gdst[4];
            float cfs[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int src = todo ? __ffs(todo) - 1 : 0;
                const bool have = todo != 0u;
                todo &= todo - 1u;
                const float bw = __shfl_sync(0xffffffffu, lw_abs, src), bh = __shfl_sync(0xffffffffu, lh_abs, src);
                const float bm = __shfl_sync(0xffffffffu, m, src);
                const uint32_t bpix = __shfl_sync(0xffffffffu, pix32, src);
                const int bg = __shfl_sync(0xffffffffu, g, src);
                const float fw = floorf(bw), fh = floorf(bh);
                const float lw = bw - fw, lh = bh - fh;
                const float cf = ((corner & 2) ? lh : 1.f - lh) * bm * ((corner & 1) ? lw : 1.f - lw);
                const int hh = (int)fh + (corner >> 1), ww = (int)fw + (corner & 1);
                const size_t n = bpix / ((size_t)q.Wo * q.Ho);
                const bool use = have && (unsigned)hh < (unsigned)q.H && (unsigned)ww < (unsigned)q.W && cf != 0.f;
                cfs[u] = use ? cf : 0.f;
                gsrc[u] = grad_out + (size_t)bpix * C + bg * kCh + 2 * cp;
                gdst[u] = use ? grad_value + ((n * q.H + hh) * q.W + ww) * C + bg * kCh + 2 * cp : nullptr;
            }
            uint32_t gw[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) gw[u] = gdst[u] ? __ldg(reinterpret_cast<const uint32_t *>(gsrc[u])) : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (gdst[u]) {
                    const float2 gq = unpack2(gw[u], T());
                    red_add2(gdst[u], pack2(cfs[u] * gq.x, cfs[u] * gq.y, T()));
                }
        }
    }
}

static bool plan(const Geom &q, RParams *pp) {
    if (q.gc != kCh || q.kh != 3 || q.kw != 3 || q.sh != 1 || q.sw != 1 || q.dh != 1 || q.dw != 1 || q.G % 8) return false;
    if (!(q.sigma >= 0.5f && q.sigma <= 1.25f)) return false;   // band = taps +- 3 px
    if ((long long)q.N * q.Ho * q.Wo == 0) return false;
    // nominal taps of a pixel x along an axis: x + a + i*sigma, i = 0..2, a = (1 - pad) - sigma; band centred on them
    const float a_w = (float)(1 - q.pw) - q.sigma, a_h = (float)(1 - q.ph) - q.sigma;
    pp->bx_rel = (int)std::floor(a_w + q.sigma + 0.5f * 7 + 0.5f - 0.5f * kBand);
    pp->by_rel = (int)std::floor(a_h + q.sigma + 0.5f * 7 + 0.5f - 0.5f * kBand);
    pp->S = (q.Wo + 7) / 8;
    pp->PR = (q.Ho + 7) / 8;
    // every cell of the map must lie in a block somebody drains: blocks 0..S x 0..PR from the grid origin
    if (pp->bx_rel > 0 || pp->by_rel > 0 || (pp->S + 1) * 8 + pp->bx_rel < q.W || (pp->PR + 1) * 8 + pp->by_rel < q.H) return false;
    pp->nbxp = (pp->S + 2) & ~1;                        // S + 1 blocks per block row, padded to a whole tile
    pp->ring = std::min(kMaxRing, 64 / pp->nbxp);
    // (three rows: two in work + one leaving.  With two -- maps 157 .. 240 px wide -- the first product of a patch row waits
    // until the block row above the previous one has been drained; the builders run ahead into their slots meanwhile)
    static const int min_ring = [] { const char *e = std::getenv("DCNV3_VRES_MINRING"); return e ? atoi(e) : 2; }();
    if (pp->ring < min_ring) return false;
    if ((long long)q.N * q.Ho * q.Wo * q.G >= (1LL << 28)) return false;   // a far-list entry = (pixel, group) << 4 | point
    const long long rows = (long long)q.N * q.G * pp->PR;
    if (rows >= (1LL << 30)) return false;

    pp->total_rows = (int)rows;
    return true;
}

template <typename T>
static bool launch_typed(const void *offset, const void *mask, const void *grad_out, void *grad_value, void *scratch,
                         size_t scratch_bytes, unsigned long long *counter, unsigned long long *thr_out, const Geom &q,
                         cudaStream_t stream, cudaError_t *err) {
    RParams pp;
    if (!plan(q, &pp)) return false;
    static const int diag = [] { const char *e = std::getenv("DCNV3_VRES_DIAG"); return e ? atoi(e) : 0; }();
    pp.diag = diag;
    CUtensorMap to, tm, tg;
    const int dtype = std::is_same<T, __half>::value ? 1 : 2;
    if (!make_run_tensor_map(&to, offset, dtype, q.N, q.Ho, q.Wo, q.G * kP * 2, kOffRow / 2)) return false;
    if (!make_run_tensor_map(&tm, mask, dtype, q.N, q.Ho, q.Wo, q.G * kP, kMskRow / 2)) return false;
    if (!make_run_tensor_map(&tg, grad_out, dtype, q.N, q.Ho, q.Wo, q.G * q.gc, 8)) return false;
    CUtensorMap tz = tg;
    if (kZOob) {
        // [64][256] 16-bit elements = one coefficient tile, box = the whole map; loaded at row 2^20: all zero fill
        EncodeTiledFn fn = encode_tiled_fn();
        void *zp = nullptr;
        if (!fn || cudaGetSymbolAddress(&zp, g_zero_tile) != cudaSuccess) return false;
        const cuuint64_t dims[2] = {256, 64}, strides[1] = {512};
        const cuuint32_t box[2] = {256, 64}, estr[2] = {1, 1};
        if (fn(&tz, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, zp, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    }
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    int ctas = std::min(pp.total_rows, num_sms);
    pp.cpg = 0;
    if (q.G <= num_sms && q.N * pp.PR >= num_sms / q.G) {
        pp.cpg = num_sms / q.G;
        ctas = pp.cpg * q.G;
    }
    if (const char *e = std::getenv("DCNV3_VRES_FLAT")) if (e[0] == '1') { pp.cpg = 0; ctas = std::min(pp.total_rows, num_sms); }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(bwd_vres<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        attr_set = true;
    }
    // beyond 1 / 32 of all points on the far path the caller's fall-back takes over; the list holds that many entries
    const unsigned long long thr = (unsigned long long)q.N * q.Ho * q.Wo * q.G * kP / 32;
    if (thr * sizeof(uint32_t) > scratch_bytes) return false;
    uint32_t *far = static_cast<uint32_t *>(scratch);
    pp.far_cap = thr;
    *thr_out = thr;
    *err = pdl_launch(true, bwd_vres<T>, dim3(ctas), dim3(kThreads), kSmemBytes, stream, to, tm, tg, tz,
                      static_cast<T *>(grad_value), far, counter, q, pp);
    if (*err == cudaSuccess) *err = cudaGetLastError();
    if (*err != cudaSuccess) return true;
    *err = pdl_launch(true, far_points<T>, dim3(num_sms * 4), dim3(256), 0, stream,
                      static_cast<const uint32_t *>(far), static_cast<const unsigned long long *>(counter), thr,
                      static_cast<const T *>(offset), static_cast<const T *>(mask), static_cast<const T *>(grad_out),
                      static_cast<T *>(grad_value), static_cast<float4 *>(scratch), scratch_bytes / 16, q);
    if (*err == cudaSuccess) *err = cudaGetLastError();
    return true;
}

}  // namespace vres

bool backward_vres_eligible(const void *offset, const void *mask, const void *grad_out, const void *grad_value, const Geom &q) {
    vres::RParams pp;
    if (((uintptr_t)grad_out | (uintptr_t)grad_value | (uintptr_t)offset | (uintptr_t)mask) % 16) return false;
    return vres::plan(q, &pp);
}

// Is the resident-accumulator kernel the FASTER choice for this shape?  (plan() says whether it is possible.)  A CTA builds
// the patch row above its run a second time and every patch row ends with a drain, so short runs and short rows lose to
// the plane form (scripts/vres_vs_plane.py, backward in us, this kernel / plane form): 80 x 80, G = 16: N = 1 51 / 44, N = 4
// 86 / 79, N = 6 112 / 110, N = 8 138 / 141, N = 16 249 / 264; 40 x 40: N = 16 78 / 78, N = 64 246 / 261; 20 x 20, G = 32 (three
// patches per row): N = 16 60 / 52, N = 128 356 / 322; 192 x 192 (ring of two): N = 1 122 / 109, N = 4 361 / 377.
// Hence: at least four patches per row and at least `min_rows_per_sm` (8; DCNV3_VRES_MIN_ROWS, 0 = always) patch rows per SM.
bool backward_vres_preferred(const Geom &q) {
    static const int min_rows_per_sm = [] { const char *e = std::getenv("DCNV3_VRES_MIN_ROWS"); return e ? atoi(e) : 8; }();
    if (min_rows_per_sm <= 0) return true;
    static const int num_sms = [] {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        return n;
    }();
    const long long rows = (long long)q.N * q.G * ((q.Ho + 7) / 8);
    return (q.Wo + 7) / 8 >= 4 && rows >= (long long)min_rows_per_sm * num_sms;
}

size_t backward_vres_scratch_bytes(const Geom &q) {
    vres::RParams pp;
    if (!vres::plan(q, &pp)) return 0;
    return (size_t)q.N * q.Ho * q.Wo * q.G * vres::kP / 32 * sizeof(uint32_t);
}

// grad_value written directly (16-bit); `scratch` holds the far-point masks (backward_vres_scratch_bytes)
bool try_launch_backward_vres(const void *offset, const void *mask, const void *grad_out, void *grad_value, void *scratch,
                              size_t scratch_bytes, unsigned long long *counter, unsigned long long *thr,
                              const Geom &q, int dtype, cudaStream_t stream, cudaError_t *err) {
    if (!backward_vres_eligible(offset, mask, grad_out, grad_value, q) || ((uintptr_t)scratch | scratch_bytes) % 16) return false;
    if (dtype == 1) return vres::launch_typed<__half>(offset, mask, grad_out, grad_value, scratch, scratch_bytes, counter, thr, q, stream, err);
    if (dtype == 2) return vres::launch_typed<__nv_bfloat16>(offset, mask, grad_out, grad_value, scratch, scratch_bytes, counter, thr, q, stream, err);
    return false;
}

}  // namespace dcnv3

// development only: the cycle stamps of the last launch (scripts/vres_timeline.py)
extern "C" __attribute__((visibility("default"))) int dcnv3_vres_debug_hang(void *host_mapped) {
    unsigned long long *p = static_cast<unsigned long long *>(host_mapped);
    return (int)cudaMemcpyToSymbol(dcnv3::vres::g_vres_hang, &p, sizeof(p));
}
extern "C" __attribute__((visibility("default"))) int dcnv3_vres_debug_read(void *dst) {
    return (int)cudaMemcpyFromSymbol(dst, dcnv3::vres::g_vres_dbg, sizeof(dcnv3::vres::g_vres_dbg));
}
