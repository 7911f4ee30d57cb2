// dcnv3_tma.cuh -- TMA (cp.async.bulk.tensor) + mbarrier helpers for the tiled DCNv3 kernels.
//
// The value tensor [N,H,W,C] is described to the TMA unit as a 4-D tiled tensor map
// (C, W, H, N); a CTA fetches the box (slice channels, WIN_W, WIN_H, 1) around its output tile
// with one instruction.  Box coordinates may be negative or run past the map: the hardware
// zero-fills those elements, which is exactly the op's zero padding
// (reference: dcnv3_im2col_cuda.cuh:56-75 reads 0 for out-of-map corners).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace dcnv3 {

// ------------------------------------------------------------------------------------ host side
using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                   const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                   const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn driver_encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// cuTensorMapEncodeTiled behind a small per-thread cache keyed by ALL of its arguments: a training loop calls the
// same layer with the same buffers (caching allocator) and shapes step after step, and a backward pass needs up to
// twelve maps -- on the small feature maps of a real model the encodes were a visible part of the launch cost.
// (A map encodes nothing but its arguments, so an equal key is an equal map.)
inline CUresult cached_encode_tiled(CUtensorMap *map, CUtensorMapDataType dt, cuuint32_t rank, void *base, const cuuint64_t *dims,
                                    const cuuint64_t *strides, const cuuint32_t *box, const cuuint32_t *estr,
                                    CUtensorMapInterleave il, CUtensorMapSwizzle sw, CUtensorMapL2promotion l2,
                                    CUtensorMapFloatOOBfill oob) {
    struct Key {
        void *base;
        cuuint64_t dims[5], strides[4];
        cuuint32_t box[5], estr[5], rank;
        int dt, il, sw, l2, oob;
    };
    struct Entry { Key key; CUtensorMap map; bool used; };
    constexpr int kEntries = 64;
    thread_local Entry cache[kEntries];
    thread_local int next = 0;
    EncodeTiledFn fn = driver_encode_tiled_fn();
    if (!fn || rank > 5) return fn ? fn(map, dt, rank, base, dims, strides, box, estr, il, sw, l2, oob) : CUDA_ERROR_NOT_SUPPORTED;
    Key k;
    memset(&k, 0, sizeof(k));
    k.base = base; k.rank = rank; k.dt = (int)dt; k.il = (int)il; k.sw = (int)sw; k.l2 = (int)l2; k.oob = (int)oob;
    for (cuuint32_t i = 0; i < rank; ++i) { k.dims[i] = dims[i]; k.box[i] = box[i]; k.estr[i] = estr[i]; if (i + 1 < rank) k.strides[i] = strides[i]; }
    for (int i = 0; i < kEntries; ++i)
        if (cache[i].used && memcmp(&cache[i].key, &k, sizeof(k)) == 0) { *map = cache[i].map; return CUDA_SUCCESS; }
    const CUresult r = fn(map, dt, rank, base, dims, strides, box, estr, il, sw, l2, oob);
    if (r == CUDA_SUCCESS) { cache[next].key = k; cache[next].map = *map; cache[next].used = true; next = (next + 1) % kEntries; }
    return r;
}

inline EncodeTiledFn encode_tiled_fn() { return driver_encode_tiled_fn() ? &cached_encode_tiled : nullptr; }

// Tensor map over a dense channels-last tensor [N,H,W,C] of `elem_bytes`-wide elements with a
// box of (box_c, box_w, box_h, 1).  Returns false if the driver refuses (caller falls back to the
// non-TMA kernel).
inline bool make_nhwc_tensor_map(CUtensorMap *map, const void *base, int dtype /*0 f32,1 f16,2 bf16*/,
                                 int N, int H, int W, int C, int box_c, int box_w, int box_h) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t es = dtype == 0 ? 4 : 2;
    const CUtensorMapDataType dt = dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                              : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
    const cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    return fn(map, dt, 4, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 4-D tensor map over a [N, Ho, Wo, row_elems] tensor of 16-bit elements, box (box_elems, 8, 8, 1).  The box
// may start at any element (a group's run) and run past the row / the map: the hardware zero-fills.
inline bool make_run_tensor_map(CUtensorMap *map, const void *base, int dtype, int N, int Ho, int Wo, int row_elems,
                                int box_elems) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t es = 2;
    const CUtensorMapDataType dt = dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const cuuint64_t dims[4] = {(cuuint64_t)row_elems, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)row_elems * es, (cuuint64_t)Wo * row_elems * es,
                                   (cuuint64_t)Ho * Wo * row_elems * es};
    const cuuint32_t box[4] = {(cuuint32_t)box_elems, 8u, 8u, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    return fn(map, dt, 4, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---------------------------------------------------------------------------------- device side
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// 4-D box load, coordinates (c, w, h, n), completion on `bar`
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c, int w,
                                            int h, int n) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(c), "r"(w), "r"(h), "r"(n), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

}  // namespace dcnv3
