// dcnv3_capi.cu -- the C ABI declared in include/dcnv3_sm100.h: argument validation + dispatch.
// Mirrors the host-side checks of the reference launchers
// (models/ops_dcnv3/src/cuda/dcnv3_cuda.cu:29-53,40-45) without any torch type in the signature.
#include "dcnv3_sm100.h"

#include "dcnv3_launch.h"

namespace {

int conv_out(int in, int pad, int dil, int k, int stride) {
    return (in + 2 * pad - (dil * (k - 1) + 1)) / stride + 1;  // dcnv3_cuda.cu:40-45
}

size_t elem_size(int dtype) { return dtype == DCNV3_F64 ? 8 : dtype == DCNV3_F32 ? 4 : 2; }

int check_geometry(int N, int H, int W, int Ho, int Wo, int G, int gc, int kh, int kw, int sh,
                   int sw, int ph, int pw, int dh, int dw, int dtype, dcnv3::Geom *q) {
    if (dtype != DCNV3_F32 && dtype != DCNV3_F16 && dtype != DCNV3_BF16 && dtype != DCNV3_F64) return DCNV3_E_DTYPE;
    if (N < 0 || H <= 0 || W <= 0 || G <= 0 || gc <= 0 || kh <= 0 || kw <= 0 || sh <= 0 ||
        sw <= 0 || dh <= 0 || dw <= 0 || ph < 0 || pw < 0)
        return DCNV3_E_SHAPE;
    if (Ho != conv_out(H, ph, dh, kh, sh) || Wo != conv_out(W, pw, dw, kw, sw) || Ho < 0 || Wo < 0)
        return DCNV3_E_SHAPE;
    // per-image index arithmetic is 32-bit inside the kernels (as in the reference), the batch
    // dimension is 64-bit
    const long long per_image_in = (long long)H * W * G * gc;
    const long long per_image_pts = (long long)Ho * Wo * G * kh * kw * 2;
    if (per_image_in >= (1LL << 31) || per_image_pts >= (1LL << 31)) return DCNV3_E_TOO_LARGE;
    *q = dcnv3::Geom{N, H, W, Ho, Wo, G, gc, kh, kw, sh, sw, ph, pw, dh, dw, 0.f};
    return DCNV3_OK;
}

}  // namespace

extern "C" {

int dcnv3_sm100_abi_version(void) { return DCNV3_SM100_ABI_VERSION; }

const char *dcnv3_sm100_strerror(int code) {
    switch (code) {
    case DCNV3_OK: return "ok";
    case DCNV3_E_DTYPE: return "dcnv3: unsupported dtype (fp32, fp16, bf16, fp64; no deterministic backward for fp64)";
    case DCNV3_E_SHAPE: return "dcnv3: invalid geometry (extent/kernel/stride/dilation/pad or Ho/Wo mismatch)";
    case DCNV3_E_NULL: return "dcnv3: null pointer";
    case DCNV3_E_WORKSPACE: return "dcnv3: workspace too small";
    case DCNV3_E_TOO_LARGE: return "dcnv3: per-image tensor exceeds 2^31 elements";
    case DCNV3_E_ALIGN: return "dcnv3: offset / grad_offset must be aligned to 2*sizeof(dtype), workspace to 16 bytes";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "dcnv3: unknown error";
    }
}

int dcnv3_forward_sm100(const void *value, const void *offset, const void *mask, void *out, int N,
                        int H, int W, int Ho, int Wo, int G, int gc, int kernel_h, int kernel_w,
                        int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                        float offset_scale, int dtype, void *stream) {
    dcnv3::Geom q;
    if (int rc = check_geometry(N, H, W, Ho, Wo, G, gc, kernel_h, kernel_w, stride_h, stride_w,
                                pad_h, pad_w, dil_h, dil_w, dtype, &q))
        return rc;
    q.sigma = offset_scale;
    if ((long long)N * Ho * Wo == 0) return DCNV3_OK;
    if (!value || !offset || !mask || !out) return DCNV3_E_NULL;
    if ((uintptr_t)offset % (2 * elem_size(dtype))) return DCNV3_E_ALIGN;
    if (dtype == DCNV3_F64) return (int)dcnv3::launch_forward_f64(value, offset, mask, out, q, (cudaStream_t)stream);
    return (int)dcnv3::launch_forward(value, offset, mask, out, q, dtype, (cudaStream_t)stream);
}

size_t dcnv3_backward_workspace_bytes(int N, int H, int W, int G, int gc, int dtype, unsigned flags) {
    if (N <= 0 || H <= 0 || W <= 0 || G <= 0 || gc <= 0) return 0;
    dcnv3::Geom q{};
    q.N = N; q.H = H; q.W = W; q.G = G; q.gc = gc;
    if (dtype == DCNV3_F64) return 0;   // fp64 atomics straight into grad_value
    return dcnv3::backward_workspace_bytes(q, dtype, flags);
}

int dcnv3_backward_sm100(const void *value, const void *offset, const void *mask,
                         const void *grad_out, void *grad_value, void *grad_offset, void *grad_mask,
                         void *workspace, size_t workspace_bytes, int N, int H, int W, int Ho,
                         int Wo, int G, int gc, int kernel_h, int kernel_w, int stride_h,
                         int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                         float offset_scale, int dtype, unsigned flags, void *stream) {
    dcnv3::Geom q;
    if (int rc = check_geometry(N, H, W, Ho, Wo, G, gc, kernel_h, kernel_w, stride_h, stride_w,
                                pad_h, pad_w, dil_h, dil_w, dtype, &q))
        return rc;
    q.sigma = offset_scale;
    if (N == 0) return DCNV3_OK;
    if (!value || !grad_value) return DCNV3_E_NULL;
    if ((long long)Ho * Wo != 0 && (!offset || !mask || !grad_out || !grad_offset || !grad_mask))
        return DCNV3_E_NULL;
    if (dtype == DCNV3_F64) {
        if (flags & DCNV3_BWD_DETERMINISTIC) return DCNV3_E_DTYPE;   // its grad_value is a sum of fp64 atomics
        if (((uintptr_t)offset | (uintptr_t)grad_offset) % 8) return DCNV3_E_ALIGN;
        return (int)dcnv3::launch_backward_f64(value, offset, mask, grad_out, grad_value, grad_offset, grad_mask, q,
                                               (cudaStream_t)stream);
    }
    const size_t need = dcnv3::backward_workspace_bytes(q, dtype, flags);
    if (need) {
        if (!workspace) return DCNV3_E_NULL;
        if (workspace_bytes < need) return DCNV3_E_WORKSPACE;
        if ((uintptr_t)workspace % 16) return DCNV3_E_ALIGN;
    }
    if (((uintptr_t)offset | (uintptr_t)grad_offset) % (2 * elem_size(dtype))) return DCNV3_E_ALIGN;
    return (int)dcnv3::launch_backward(value, offset, mask, grad_out, grad_value, grad_offset,
                                       grad_mask, workspace, q, dtype, flags, (cudaStream_t)stream);
}

}  // extern "C"
