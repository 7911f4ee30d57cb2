// dcnv3_tc.cuh -- tcgen05 / tensor-memory wrappers and the drain helpers shared by the grad_value kernels
// (dcnv3_backward_vmma.cu, dcnv3_backward_vband.cu): UMMA descriptors, TMEM loads, whole-sector reductions of a
// finished accumulator block into the fp32 plane, and the direct path of a point that leaves the band.
#pragma once
#include "dcnv3_common.cuh"
#include "dcnv3_tma.cuh"
#include "dcnv3_strip_io.cuh"

#ifndef VMMA_DIAG
#define VMMA_DIAG 0
#endif

namespace dcnv3 {
namespace tc {

using namespace strip;

// ------------------------------------------------------------------------------------ tcgen05 wrappers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], both K-major
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major operand tile with 128-byte swizzle (rows of 64 16-bit elements, 8-row groups of 1024 B):
// start address >> 4 at [0,14), stride byte offset >> 4 at [32,46), version 1 at [46,48),
// SWIZZLE_128B = 2 at [61,64)  (same encoding as dcnv3_proj.cu)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major operand without swizzle (cute/atom/mma_traits_sm100.hpp, Major::MN / INTERLEAVE): core matrices of
// 8 K-rows x 16 bytes (8 MN elements) stored contiguously (128 B); `lbo` = byte distance between core matrices
// along K (leading byte offset, [16,30)), `sbo` = along MN (stride byte offset, [32,46)); layout type 0.
__device__ __forceinline__ uint64_t umma_desc_mn_plain(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: fp32 accumulate, a/b format (F16 = 0, BF16 = 1), K-major A (bit 15 = 0), MN-major B
// (bit 16 = 1: grad_out stays pixel-major, as the TMA delivers it), N >> 3, M >> 4
__host__ __device__ constexpr uint32_t umma_idesc(int fmt, int M, int N) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
#define VMMA_TMEM_LD_16(taddr, r)                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),    \
                   "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]) \
                 : "r"(taddr))
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// A point beyond the band: its four coefficient x grad_out rows go straight to the fp32 plane.
template <typename T, int NC = 16>
__device__ __noinline__ void far_point(float *gv_img, int H, int W, int row_stride, int C, int h0, int w0, float c0,
                                       float c1, float c2, float c3, uint4 ga, uint4 gb) {
    float g[16];
    unpack<T>(ga, g);
    unpack<T>(gb, g + 8);   // (NC == 8: unused)
    const float cf[4] = {c0, c1, c2, c3};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int hh = h0 + (t >> 1), ww = w0 + (t & 1);
        if ((unsigned)hh < (unsigned)H && (unsigned)ww < (unsigned)W && cf[t] != 0.f) {
            float *dst = gv_img + (ptrdiff_t)hh * row_stride + (ptrdiff_t)ww * C;
#pragma unroll
            for (int e = 0; e < NC; e += 4)
                red_add4(dst + e, make_float4(cf[t] * g[e], cf[t] * g[e + 1], cf[t] * g[e + 2], cf[t] * g[e + 3]));
        }
    }
}

// One finished block (8 band rows x 16 columns) of both strips: TMEM -> registers -> reductions.
// Thread (warp, lane) holds TMEM lane 32 warp + lane = cell (row 2 warp + (lane >> 4), column lane & 15),
// 16 fp32 channels = two 32-byte sectors.  A lane pair (two adjacent cells) swaps halves so that the two
// lanes of a pair write the two halves of ONE sector in the same instruction (whole-sector reductions).
__device__ __forceinline__ void drain_cells(const float (&r)[16], int lane, float *p_even, bool ok_even, bool ok_odd, int C) {
    const bool odd = lane & 1;
    // a cell nobody sampled (the band's outer columns / rows mostly) holds exact zeros: nothing to reduce.
    // (0 x Inf / NaN in grad_out gives NaN != 0, so a poisoned neighbourhood is still written.)
    bool nz = false;
#pragma unroll
    for (int j = 0; j < 16; ++j) nz |= r[j] != 0.f;
    const bool nz_other = __shfl_xor_sync(0xffffffffu, (int)nz, 1) != 0;
    ok_even = ok_even && (odd ? nz_other : nz);
    ok_odd = ok_odd && (odd ? nz : nz_other);
    float rv[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        rv[j] = __shfl_xor_sync(0xffffffffu, odd ? r[j] : r[4 + j], 1);
        rv[4 + j] = __shfl_xor_sync(0xffffffffu, odd ? r[8 + j] : r[12 + j], 1);
    }
    // even lane: E[0:4] own, O[0:4] recv, E[8:12] own, O[8:12] recv; odd lane: E[4:8] recv, O[4:8] own, E[12:16] recv, O[12:16] own
    float *pe = p_even + (odd ? 4 : 0), *po = p_even + C + (odd ? 4 : 0);
    if (VMMA_DIAG == 3) { ok_even = ok_even && rv[0] == 12345.678f; ok_odd = ok_odd && rv[1] == 12345.678f; }
    if (ok_even) {
        red_add4(pe, odd ? make_float4(rv[0], rv[1], rv[2], rv[3]) : make_float4(r[0], r[1], r[2], r[3]));
        red_add4(pe + 8, odd ? make_float4(rv[4], rv[5], rv[6], rv[7]) : make_float4(r[8], r[9], r[10], r[11]));
    }
    if (ok_odd) {
        red_add4(po, odd ? make_float4(r[4], r[5], r[6], r[7]) : make_float4(rv[0], rv[1], rv[2], rv[3]));
        red_add4(po + 8, odd ? make_float4(r[12], r[13], r[14], r[15]) : make_float4(rv[4], rv[5], rv[6], rv[7]));
    }
}
}  // namespace tc
}  // namespace dcnv3
