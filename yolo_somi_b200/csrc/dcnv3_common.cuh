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
// Packed 16-byte channel chunks (uint4 = 4 fp32 or 8 bf16/fp16) and the math on them.
//
// 16-bit types never get unpacked: sm_100a has a mixed-precision FMA (PTX fma.rn.f32.bf16 /
// fma.rn.f32.f16 -> SASS FHFMA) that multiplies two 16-bit operands -- either half of a 32-bit
// register, selected in the instruction -- exactly and accumulates in fp32 with one rounding.
//   * products of two DATA values (grad_out * value in the backward) are therefore exact;
//   * products weight * value need the fp32 weight as 16-bit operands: `Weight` splits it into
//     hi + lo (w = hi + lo up to 2^-17 relative), two FHFMAs per element.  With kFastWeights the
//     lo term is dropped (weight rounded to the I/O dtype, 2^-9 relative for bf16).
template <typename T> struct Chunk { static constexpr int kElems = 16 / sizeof(T); };

__device__ __forceinline__ uint4 ldg16(const void *p, bool pred) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (pred) v = __ldg(reinterpret_cast<const uint4 *>(p));
    return v;
}

__device__ __forceinline__ float mix_fma(uint16_t a, uint16_t b, float c, __nv_bfloat16) {
    float d;
    asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float mix_fma(uint16_t a, uint16_t b, float c, __half) {
    float d;
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
    return d;
}
__device__ __forceinline__ uint16_t lo16(uint32_t w) { return (uint16_t)(w & 0xffffu); }
__device__ __forceinline__ uint16_t hi16(uint32_t w) { return (uint16_t)(w >> 16); }

__device__ __forceinline__ uint16_t bits16(float v, __nv_bfloat16) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }
__device__ __forceinline__ uint16_t bits16(float v, __half) { return __half_as_ushort(__float2half_rn(v)); }
__device__ __forceinline__ float f32_of(uint16_t b, __nv_bfloat16) { return __uint_as_float((uint32_t)b << 16); }
__device__ __forceinline__ float f32_of(uint16_t b, __half) { return __half2float(__ushort_as_half(b)); }

// fp32 weight prepared for `axpy`
template <typename T, bool FAST> struct Weight {
    uint16_t hi, lo;
    __device__ __forceinline__ explicit Weight(float w) {
        hi = bits16(w, T());
        lo = FAST ? (uint16_t)0 : bits16(w - f32_of(hi, T()), T());
    }
};
template <bool FAST> struct Weight<float, FAST> {
    float w;
    __device__ __forceinline__ explicit Weight(float w_) : w(w_) {}
};

// acc[0..kElems) += w * chunk
template <typename T, bool FAST>
__device__ __forceinline__ void axpy(float *acc, const uint4 &q, const Weight<T, FAST> &w) {
    if constexpr (sizeof(T) == 4) {
        acc[0] += w.w * __uint_as_float(q.x);
        acc[1] += w.w * __uint_as_float(q.y);
        acc[2] += w.w * __uint_as_float(q.z);
        acc[3] += w.w * __uint_as_float(q.w);
    } else {
        const uint32_t r[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc[2 * i] = mix_fma(lo16(r[i]), w.hi, acc[2 * i], T());
            acc[2 * i + 1] = mix_fma(hi16(r[i]), w.hi, acc[2 * i + 1], T());
            if (!FAST) {
                acc[2 * i] = mix_fma(lo16(r[i]), w.lo, acc[2 * i], T());
                acc[2 * i + 1] = mix_fma(hi16(r[i]), w.lo, acc[2 * i + 1], T());
            }
        }
    }
}

// sum_i a_i * b_i + init, both chunks of the same type (exact products for 16-bit types)
template <typename T> __device__ __forceinline__ float dot(const uint4 &a, const uint4 &b, float init) {
    const uint32_t x[4] = {a.x, a.y, a.z, a.w}, y[4] = {b.x, b.y, b.z, b.w};
    float s0 = init, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s0 = mix_fma(lo16(x[i]), lo16(y[i]), s0, T());
        s1 = mix_fma(hi16(x[i]), hi16(y[i]), s1, T());
    }
    return s0 + s1;
}
template <> __device__ __forceinline__ float dot<float>(const uint4 &a, const uint4 &b, float init) {
    float s = init;
    s += __uint_as_float(a.x) * __uint_as_float(b.x);
    s += __uint_as_float(a.y) * __uint_as_float(b.y);
    s += __uint_as_float(a.z) * __uint_as_float(b.z);
    s += __uint_as_float(a.w) * __uint_as_float(b.w);
    return s;
}

// chunk -> fp32 values
template <typename T> __device__ __forceinline__ void unpack(const uint4 &q, float *f) {
    const float2 a = unpack2(q.x, T()), b = unpack2(q.y, T()), c = unpack2(q.z, T()), d = unpack2(q.w, T());
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
template <> __device__ __forceinline__ void unpack<float>(const uint4 &q, float *f) {
    f[0] = __uint_as_float(q.x); f[1] = __uint_as_float(q.y);
    f[2] = __uint_as_float(q.z); f[3] = __uint_as_float(q.w);
}
// fp32 values -> chunk
template <typename T> __device__ __forceinline__ uint4 pack(const float *f) {
    return make_uint4(pack2(f[0], f[1], T()), pack2(f[2], f[3], T()), pack2(f[4], f[5], T()), pack2(f[6], f[7], T()));
}
template <> __device__ __forceinline__ uint4 pack<float>(const float *f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
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

// Branch-free form used by the vector kernels (valid only when `inside`): corner rows/columns are
// clamped into the map so that every corner read is unconditional, and the zero padding of the
// reference (:56-75) is applied to the 1-D interpolation factors instead: a factor is 0 when its
// row / column lies outside.  A clamped read always lands on the tap's OTHER row / column, whose
// factor is non-zero, so NaN/Inf propagate exactly as they do in the reference.
struct ClampedTap {
    bool inside;
    int row_lo, row_hi, col_lo, col_hi;  // clamped corner coordinates
    float top, bot, lef, rig;            // 1.0 if that row / column is inside the map, else 0.0
    float hh, lh, hw, lw;                // raw factors (hh = 1 - lh, hw = 1 - lw)
};

__device__ __forceinline__ ClampedTap make_clamped_tap(float loc_h, float loc_w, int H, int W) {
    ClampedTap t;
    t.inside = loc_h > -1.f && loc_w > -1.f && loc_h < (float)H && loc_w < (float)W;
    const float fh = floorf(loc_h), fw = floorf(loc_w);
    const int h0 = (int)fh, w0 = (int)fw;
    t.lh = loc_h - fh;
    t.lw = loc_w - fw;
    t.hh = 1.f - t.lh;
    t.hw = 1.f - t.lw;
    t.top = h0 >= 0 ? 1.f : 0.f;
    t.bot = h0 + 1 <= H - 1 ? 1.f : 0.f;
    t.lef = w0 >= 0 ? 1.f : 0.f;
    t.rig = w0 + 1 <= W - 1 ? 1.f : 0.f;
    t.row_lo = max(h0, 0);
    t.row_hi = min(h0 + 1, H - 1);
    t.col_lo = max(w0, 0);
    t.col_hi = min(w0 + 1, W - 1);
    return t;
}

// Anchor of output column/row `o` along one axis before the learned offset is applied:
// (c - pad + o*stride) - c*sigma, with c = (dil*(k-1))>>1.   (:232-236,249-252)
__device__ __forceinline__ float axis_base(int o, int k, int stride, int pad, int dil, float sigma) {
    const int c = (dil * (k - 1)) >> 1;
    return (float)(c - pad + o * stride) - (float)c * sigma;
}

}  // namespace dcnv3
