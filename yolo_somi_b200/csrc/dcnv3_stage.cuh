// dcnv3_stage.cuh -- cooperative staging of a tile's (dx,dy) offsets and mask weights in shared
// memory, shared by the tiled kernels.
//
// A pixel's P pairs are contiguous in global memory but pixels are G*P pairs apart, so consecutive
// threads take consecutive pairs of the same pixel (coalesced) rather than each thread walking its
// own pixel (which costs one L1 sector lookup per 4-byte load).  With a compile-time point count all
// loads of a thread are issued BEFORE the first store, so a thread waits for DRAM once, not P times
// (the rolled loop was the top stall of the first tiled kernels: long_scoreboard 4.2 per issue).
#pragma once

#include "dcnv3_common.cuh"

namespace dcnv3 {

template <typename T> struct StagePair { using type = uint32_t; };   // two 16-bit values
template <> struct StagePair<float> { using type = float2; };

// s_off / s_msk are indexed [pixel-in-tile * P + p], pixel-in-tile = ty * TILE_W + tx.
template <typename T, int PCT /* compile-time P or 0 */, int NTHREADS, int TILE_W>
__device__ __forceinline__ void stage_offsets_masks(const T *__restrict__ offset, const T *__restrict__ mask,
                                                    typename StagePair<T>::type *s_off, T *s_msk, int P,
                                                    int tid, int wo0, int ho0, int Wo, int Ho, int G, int g,
                                                    size_t img_pix) {
    using Pair = typename StagePair<T>::type;
    // 64-bit base once per thread; everything per element is 32-bit (per-image extents < 2^31)
    const size_t base = (img_pix * G + g) * (size_t)(PCT > 0 ? PCT : P);
    const Pair *obase = reinterpret_cast<const Pair *>(offset) + base;
    const T *mbase = mask + base;
    if constexpr (PCT > 0) {
        Pair po[PCT];
        T pm[PCT];
        bool ok[PCT];
#pragma unroll
        for (int it = 0; it < PCT; ++it) {
            const unsigned idx = tid + it * NTHREADS;
            const unsigned px = idx / PCT, p = idx - px * PCT;
            const unsigned w = wo0 + (px % TILE_W), h = ho0 + (px / TILE_W);
            ok[it] = w < (unsigned)Wo && h < (unsigned)Ho;
            if (ok[it]) {
                const unsigned rel = (h * Wo + w) * (unsigned)(G * PCT) + p;
                po[it] = __ldg(obase + rel);
                pm[it] = __ldg(mbase + rel);
            }
        }
#pragma unroll
        for (int it = 0; it < PCT; ++it) {
            const int idx = tid + it * NTHREADS;
            if (ok[it]) {
                s_off[idx] = po[it];
                s_msk[idx] = pm[it];
            }
        }
    } else {
        for (unsigned idx = tid; idx < (unsigned)(NTHREADS * P); idx += NTHREADS) {
            const unsigned px = idx / P, p = idx - px * P;
            const unsigned w = wo0 + (px % TILE_W), h = ho0 + (px / TILE_W);
            if (w < (unsigned)Wo && h < (unsigned)Ho) {
                const unsigned rel = (h * Wo + w) * (unsigned)(G * P) + p;
                s_off[idx] = __ldg(obase + rel);
                s_msk[idx] = __ldg(mbase + rel);
            }
        }
    }
}

}  // namespace dcnv3
