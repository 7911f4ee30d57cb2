// dcnv3_strip_io.cuh -- pieces shared by the strip-shaped backward kernels (dcnv3_backward_strip.cu,
// dcnv3_backward_vmma.cu): the per-warp staging buffer of an 8 x 4 pixel patch (offsets, masks, grad_out)
// with its sector-minimal cp.async / write-back order, and the small PTX wrappers they use.
#pragma once
#include "dcnv3_common.cuh"

namespace dcnv3 {
namespace strip {

constexpr int kStripW = 8, kPatchH = 4;          // a warp's patch: lane <-> pixel
constexpr int kCh = 16, kSliceBytes = 32;
constexpr int kP = 9;
constexpr int kMskWords = 7;                                    // per pixel: 5 words used, odd stride
constexpr int kStageOff = 0;                                    // [32][9] u32
constexpr int kStageMsk = kStageOff + 32 * kP * 4;              // [32][7] u32
constexpr int kStageGout = kStageMsk + 32 * kMskWords * 4;      // [32][32 B]
constexpr int kStageBytes = kStageGout + 32 * kSliceBytes;      // 3072 per warp
constexpr int kIoTblBytes = (kP + 5) * 32 * 4;                  // staging I/O index table, per CTA

__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128_zero(uint32_t a) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(a), "r"(0u) : "memory");
}
__device__ __forceinline__ void red_add2(float *p, float a, float b) { atomicAdd(reinterpret_cast<float2 *>(p), make_float2(a, b)); }
__device__ __forceinline__ void red_add4(float *p, float4 v) { atomicAdd(reinterpret_cast<float4 *>(p), v); }

// cp.async with zero fill: copies `src_bytes` (<= size) and zero-fills the rest
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory"); }

// ------------------------------------------------------------------------------------------------
// Staging I/O of one warp, out of line (one copy of the code, called once per step): write the
// finished step's grad_offset / grad_mask out of the staging buffer, then request the next step's
// offsets / masks / grad_out with cp.async (no registers held while they are in flight).
template <typename T> struct IoCtx {   // kernel-constant, one copy per CTA in shared memory
    const T *offset, *mask, *grad_out;
    T *grad_offset, *grad_mask;
    const unsigned char *mask_end;
    int Wo, Ho, G, C;
};

// Index table (shared memory, built once per CTA): element k = lane + 32 it of a patch's 32 x 9
// staged values belongs to pixel k / 9, point k % 9 -- consecutive lanes walk a pixel's contiguous
// 36-byte (offsets) / 18-byte (masks) run, so a warp access touches the minimal number of 32-byte
// sectors (one LSU wavefront each).  Entry = element offset from the patch's first run | pixel << 27;
// rows 9..13: the same for the mask runs' 5 enclosing words (k / 5, k % 5 << 24).
__device__ __forceinline__ void build_io_table(uint32_t *tbl, int Wo, int G9, int tid, int nthreads) {
    for (int e = tid; e < (kP + 5) * 32; e += nthreads) {
        const int it = e >> 5, lane = e & 31;
        uint32_t v;
        if (it < kP) {
            const int k = lane + 32 * it, px = k / kP, p = k - px * kP;
            v = (uint32_t)(((px >> 3) * Wo + (px & 7)) * G9 + p) | ((uint32_t)px << 27);
        } else {
            const int k = lane + 32 * (it - kP), px = k / 5, wd = k - px * 5;
            v = (uint32_t)(((px >> 3) * Wo + (px & 7)) * G9) | ((uint32_t)wd << 24) | ((uint32_t)px << 27);
        }
        tbl[e] = v;
    }
}

template <typename T>
__device__ __noinline__ void stage_io(const IoCtx<T> *io_s, const uint32_t *tbl, uint32_t sa /* staging buffer */,
                                      size_t pix_w, int g_w, int wb_w, int hb_w, int do_w, size_t pix_p, int g_p,
                                      int wb_p, int hb_p, int do_p) {
    const IoCtx<T> io = *io_s;     // registers from here on (the asm statements below clobber memory)
    const int lane = threadIdx.x & 31;
    const int Wo = io.Wo, Ho = io.Ho;
    uint32_t ent[kP];
#pragma unroll
    for (int it = 0; it < kP; ++it) ent[it] = tbl[it * 32 + lane];
    // Patches that lie fully inside the output map (the common case) take warp-uniform branches with
    // unconditional accesses; a per-element guard costs a divergence barrier around every access.
    if (do_w) {
        const size_t e0 = (pix_w * io.G + g_w) * kP;             // first element of the patch's first run
        uint32_t *ob = reinterpret_cast<uint32_t *>(io.grad_offset) + e0;
        uint16_t *mb = reinterpret_cast<uint16_t *>(io.grad_mask) + e0;
        const unsigned par0 = (unsigned)e0 & 1u;
        const bool full = wb_w + kStripW <= Wo && hb_w + kPatchH <= Ho;
        uint32_t vo[kP], vm[kP];
        // all reads first (independent), then the stores
#pragma unroll
        for (int it = 0; it < kP; ++it) {
            const int px = (int)(ent[it] >> 27), rel = (int)(ent[it] & 0x7ffffffu);
            const int p = lane + 32 * it - px * kP;
            const unsigned shp = (par0 + (unsigned)(rel - p)) & 1u;   // misalignment of the staged mask run
            vo[it] = lds32(sa + kStageOff + (lane + 32 * it) * 4);
            vm[it] = lds16(sa + kStageMsk + px * (kMskWords * 4) + (shp + p) * 2);
        }
        if (full) {
#pragma unroll
            for (int it = 0; it < kP; ++it) {
                const int rel = (int)(ent[it] & 0x7ffffffu);
                ob[rel] = vo[it];
                mb[rel] = (uint16_t)vm[it];
            }
        } else {
#pragma unroll 1
            for (int it = 0; it < kP; ++it) {
                const int px = (int)(ent[it] >> 27), rel = (int)(ent[it] & 0x7ffffffu);
                if (wb_w + (px & 7) < Wo && hb_w + (px >> 3) < Ho) {
                    ob[rel] = vo[it];
                    mb[rel] = (uint16_t)vm[it];
                }
            }
        }
    }
    __syncwarp();   // every lane is done reading the staging buffer (gather results, B fragments)
    if (do_p) {
        const size_t e0 = (pix_p * io.G + g_p) * kP;
        const uint32_t *os = reinterpret_cast<const uint32_t *>(io.offset) + e0;
        const unsigned char *ms = reinterpret_cast<const unsigned char *>(io.mask) + e0 * 2;
        const unsigned mlow = (unsigned)(uintptr_t)ms & 3u;
        const bool full = wb_p + kStripW <= Wo && hb_p + kPatchH <= Ho;
        const T *g_first = io.grad_out + pix_p * io.C + g_p * kCh;
        if (full) {
#pragma unroll
            for (int it = 0; it < kP; ++it)
                cp_async4(sa + kStageOff + (lane + 32 * it) * 4, os + (ent[it] & 0x7ffffffu), 4);
            // the 18-byte mask run of a pixel is staged from its 5 enclosing 4-byte words; only the
            // tensor's very last word can be half outside (warp-uniform test for the whole patch)
            const bool tail = ms + ((size_t)((kPatchH - 1) * Wo + kStripW) * (io.G * kP)) * 2 + 4 > io.mask_end;
#pragma unroll
            for (int it = 0; it < 5; ++it) {
                const uint32_t en = tbl[(kP + it) * 32 + lane];
                const int px = (int)(en >> 27), wd = (int)((en >> 24) & 7u);
                const unsigned relb = (en & 0xffffffu) * 2u;                      // byte offset of the run
                const unsigned low = (mlow + relb) & 3u;                          // its misalignment (0 or 2)
                const unsigned char *src = ms + ((ptrdiff_t)relb - (ptrdiff_t)low + wd * 4);
                if (!tail) cp_async4(sa + kStageMsk + (px * kMskWords + wd) * 4, src, 4);
                else cp_async4(sa + kStageMsk + (px * kMskWords + wd) * 4, src, src + 4 <= io.mask_end ? 4 : 2);
            }
#pragma unroll
            for (int r2 = 0; r2 < 2; ++r2) {
                const int px = r2 * 16 + (lane >> 1), row = px >> 3, c8 = px & 7, ck = lane & 1;
                cp_async16(sa + kStageGout + px * kSliceBytes + ck * 16, g_first + (size_t)(row * Wo + c8) * io.C + ck * 8, 16);
            }
        } else {
            // ragged patch: copies of pixels outside the map are skipped by a zero source size
            // (offsets / masks of such pixels are never read; their grad_out must read as zero:
            // the A columns are zero, but the product must not see NaN bits)
#pragma unroll 1
            for (int it = 0; it < kP; ++it) {
                const int px = (int)(ent[it] >> 27);
                const bool ok = wb_p + (px & 7) < Wo && hb_p + (px >> 3) < Ho;
                cp_async4(sa + kStageOff + (lane + 32 * it) * 4, os + (ok ? (ent[it] & 0x7ffffffu) : 0u), ok ? 4 : 0);
            }
#pragma unroll 1
            for (int it = 0; it < 5; ++it) {
                const uint32_t en = tbl[(kP + it) * 32 + lane];
                const int px = (int)(en >> 27), wd = (int)((en >> 24) & 7u);
                const bool ok = wb_p + (px & 7) < Wo && hb_p + (px >> 3) < Ho;
                const unsigned relb = (en & 0xffffffu) * 2u;
                const unsigned low = (mlow + relb) & 3u;
                const unsigned char *src = ok ? ms + ((ptrdiff_t)relb - (ptrdiff_t)low + wd * 4) : reinterpret_cast<const unsigned char *>(io.mask);
                cp_async4(sa + kStageMsk + (px * kMskWords + wd) * 4, src, ok ? (src + 4 <= io.mask_end ? 4 : 2) : 0);
            }
#pragma unroll
            for (int r2 = 0; r2 < 2; ++r2) {
                const int px = r2 * 16 + (lane >> 1), row = px >> 3, c8 = px & 7, ck = lane & 1;
                const bool ok = wb_p + c8 < Wo && hb_p + row < Ho;
                cp_async16(sa + kStageGout + px * kSliceBytes + ck * 16,
                           ok ? g_first + (size_t)(row * Wo + c8) * io.C + ck * 8 : io.grad_out, ok ? 16 : 0);
            }
        }
    }
    cp_async_commit();
}

}  // namespace strip
}  // namespace dcnv3
