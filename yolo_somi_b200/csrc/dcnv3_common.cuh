// dcnv3_common.cuh -- shared device helpers for the sm_100a DCNv3 kernels.
//
// Semantics implemented by every kernel in this directory (reference paths are relative to the
// reference tree, models/ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh):
//   sampling location   :232-260   loc_w = (c_w - pad_w + wo*stride_w) - c_w*s + (i*dil_w + dx)*s
//   range test          :262-263   loc > -1 && loc < extent on both axes
//   bilinear, 0 border  :32-80
//   gradients           :82-147
// All arithmetic is fp32 whatever the I/O dtype (reference: opmath_t, :30).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dcnv3 {

struct Geom {
    int N, H, W, Ho, Wo, G, gc;
    int kh, kw, sh, sw, ph, pw, dh, dw;
    float sigma;  // offset_scale
};

// ---------------------------------------------------------------------------------------------
// scalar conversions
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
    return __float2bfloat16_rn(v);
}

// two packed 16-bit values <-> float2
__device__ __forceinline__ float2 unpack2(uint32_t w, __half) {
    return __half22float2(*reinterpret_cast<const __half2 *>(&w));
}
__device__ __forceinline__ float2 unpack2(uint32_t w, __nv_bfloat16) {
    // bf16 -> fp32 is a 16-bit shift: low half << 16, high half masked
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ uint32_t pack2(float a, float b, __half) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ uint32_t pack2(float a, float b, __nv_bfloat16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// ---------------------------------------------------------------------------------------------
// VEC contiguous channels of type T <-> float[VEC].  VEC*sizeof(T) is 16 bytes on the vector
// paths (one 128-bit access) and sizeof(T) on the scalar path.
template <typename T, int VEC> struct ChanVec;

template <typename T> struct ChanVec<T, 1> {
    static __device__ __forceinline__ void load(const T *p, bool pred, float (&f)[1]) {
        f[0] = pred ? to_f32(__ldg(p)) : 0.f;
    }
    static __device__ __forceinline__ void store(T *p, const float (&f)[1]) { p[0] = from_f32<T>(f[0]); }
};

template <> struct ChanVec<float, 4> {
    static __device__ __forceinline__ void load(const float *p, bool pred, float (&f)[4]) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pred) v = __ldg(reinterpret_cast<const float4 *>(p));
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    }
    static __device__ __forceinline__ void store(float *p, const float (&f)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(f[0], f[1], f[2], f[3]);
    }
};

template <typename T> struct ChanVec<T, 8> {  // T = __half | __nv_bfloat16
    static __device__ __forceinline__ void load(const T *p, bool pred, float (&f)[8]) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (pred) v = __ldg(reinterpret_cast<const uint4 *>(p));
        float2 a = unpack2(v.x, T()), b = unpack2(v.y, T()), c = unpack2(v.z, T()), d = unpack2(v.w, T());
        f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
    }
    static __device__ __forceinline__ void store(T *p, const float (&f)[8]) {
        uint4 v;
        v.x = pack2(f[0], f[1], T()); v.y = pack2(f[2], f[3], T());
        v.z = pack2(f[4], f[5], T()); v.w = pack2(f[6], f[7], T());
        *reinterpret_cast<uint4 *>(p) = v;
    }
};

// (dx, dy) of one sampling point: adjacent elements, naturally aligned to 2*sizeof(T).
__device__ __forceinline__ float2 load_pair(const float *p) {
    return __ldg(reinterpret_cast<const float2 *>(p));
}
__device__ __forceinline__ float2 load_pair(const __half *p) {
    return unpack2(__ldg(reinterpret_cast<const uint32_t *>(p)), __half());
}
__device__ __forceinline__ float2 load_pair(const __nv_bfloat16 *p) {
    return unpack2(__ldg(reinterpret_cast<const uint32_t *>(p)), __nv_bfloat16());
}
__device__ __forceinline__ void store_pair(float *p, float a, float b) {
    *reinterpret_cast<float2 *>(p) = make_float2(a, b);
}
__device__ __forceinline__ void store_pair(__half *p, float a, float b) {
    *reinterpret_cast<uint32_t *>(p) = pack2(a, b, __half());
}
__device__ __forceinline__ void store_pair(__nv_bfloat16 *p, float a, float b) {
    *reinterpret_cast<uint32_t *>(p) = pack2(a, b, __nv_bfloat16());
}

// ---------------------------------------------------------------------------------------------
// One bilinear tap: location -> corner predicates and weights.
struct Tap {
    bool inside;         // range test (:262-263); nothing is read or written when false
    bool tl, tr, bl, br; // corner lies inside the map (zero padding otherwise)
    int h0, w0;
    float lh, lw, hh, hw;
};

__device__ __forceinline__ Tap make_tap(float loc_h, float loc_w, int H, int W) {
    Tap t;
    t.inside = loc_h > -1.f && loc_w > -1.f && loc_h < (float)H && loc_w < (float)W;
    const float fh = floorf(loc_h), fw = floorf(loc_w);
    t.h0 = (int)fh;
    t.w0 = (int)fw;
    t.lh = loc_h - fh;
    t.lw = loc_w - fw;
    t.hh = 1.f - t.lh;
    t.hw = 1.f - t.lw;
    const bool top = t.h0 >= 0, bot = t.h0 + 1 <= H - 1;
    const bool lef = t.w0 >= 0, rig = t.w0 + 1 <= W - 1;
    t.tl = t.inside && top && lef;
    t.tr = t.inside && top && rig;
    t.bl = t.inside && bot && lef;
    t.br = t.inside && bot && rig;
    return t;
}

// Anchor of output column/row `o` along one axis before the learned offset is applied:
// (c - pad + o*stride) - c*sigma, with c = (dil*(k-1))>>1.   (:232-236,249-252)
__device__ __forceinline__ float axis_base(int o, int k, int stride, int pad, int dil, float sigma) {
    const int c = (dil * (k - 1)) >> 1;
    return (float)(c - pad + o * stride) - (float)c * sigma;
}

}  // namespace dcnv3
