/*
 * dcnv3_sm100.h -- C ABI of libdcnv3_sm100.so: the DCNv3 core (grouped deformable bilinear
 * sampling x modulation mask, forward + backward) as hand-written sm_100a CUDA kernels.
 *
 * This is the drop-in seam for the native extension the reference calls "DCNv3"
 * (paths relative to the reference tree):
 *
 *   dcnv3_forward_sm100   replaces  dcnv3_forward   models/ops_dcnv3/src/dcnv3.h:20-38
 *                                   -> dcnv3_cuda_forward   src/cuda/dcnv3_cuda.cu:21-85
 *                                   -> dcnv3_im2col_cuda    src/cuda/dcnv3_im2col_cuda.cuh:841-868
 *   dcnv3_backward_sm100  replaces  dcnv3_backward  models/ops_dcnv3/src/dcnv3.h:40-59
 *                                   -> dcnv3_cuda_backward  src/cuda/dcnv3_cuda.cu:87-174
 *                                   -> dcnv3_col2im_cuda    src/cuda/dcnv3_im2col_cuda.cuh:870-1045
 *
 * The reference binds those two through pybind11 (src/vision.cpp:14-17) with at::Tensor
 * arguments.  Here the boundary is plain C: device pointers, sizes, a dtype tag and a CUDA
 * stream.  The Python module `DCNv3` (repo root) rebuilds the reference's tensor-level
 * signatures on top of it, see INTEGRATION.md.
 *
 * Contract (same as the reference unless noted)
 *   - every tensor is dense, channels-last:  value [N,H,W,G*gc], offset [N,Ho,Wo,G*K*2]
 *     ((dx,dy) per point, point index p = i_w*kernel_h + j_h), mask [N,Ho,Wo,G*K],
 *     out / grad_out [N,Ho,Wo,G*gc];  Ho = (H + 2*pad_h - (dil_h*(kernel_h-1)+1))/stride_h + 1
 *     (src/cuda/dcnv3_cuda.cu:40-45), K = kernel_h*kernel_w;
 *   - all tensors of one call share one dtype (fp32 / fp16 / bf16 / fp64; the reference has no
 *     bf16);  arithmetic is fp32 for the first three and fp64 for fp64 I/O -- a plain correctness
 *     path, the one the reference's own test script drives first (models/ops_dcnv3/test.py:33-57)
 *     (reference: opmath_t, dcnv3_im2col_cuda.cuh:30);  gradients are produced in the I/O dtype
 *     (reference: fp32 accumulate then cast, dcnv3_cuda.cu:126-133,168-173);
 *   - inputs are borrowed, outputs are caller-allocated and fully overwritten: no pre-zeroing
 *     is needed (the reference zero-fills everything, dcnv3_cuda.cu:55-57,131-133);
 *   - work is enqueued on `stream`; the call does not synchronise and does not allocate;
 *   - re-entrant, no global state, no device guard (caller selects the device).
 *
 * Return value: 0 on success; a positive value is a cudaError_t from a launch (the reference
 * only printf()s launch errors, dcnv3_im2col_cuda.cuh:864-867,1041-1044); negative values are the
 * DCNV3_E_* argument errors below.  dcnv3_sm100_strerror() describes either.
 */
#ifndef DCNV3_SM100_H_
#define DCNV3_SM100_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCNV3_SM100_ABI_VERSION 1

#if defined(__GNUC__)
#define DCNV3_API __attribute__((visibility("default")))
#else
#define DCNV3_API
#endif

enum dcnv3_dtype { DCNV3_F32 = 0, DCNV3_F16 = 1, DCNV3_BF16 = 2, DCNV3_F64 = 3 };

enum dcnv3_error {
    DCNV3_OK = 0,
    DCNV3_E_DTYPE = -1,     /* dtype tag not one of dcnv3_dtype (or DETERMINISTIC with fp64) */
    DCNV3_E_SHAPE = -2,     /* non-positive extent / kernel / stride / dilation, or Ho/Wo  */
                            /* inconsistent with the formula above                         */
    DCNV3_E_NULL = -3,      /* a required pointer is NULL                                  */
    DCNV3_E_WORKSPACE = -4, /* workspace smaller than dcnv3_backward_workspace_bytes()     */
    DCNV3_E_TOO_LARGE = -5, /* a per-image tensor has 2^31 or more elements                */
    DCNV3_E_ALIGN = -6      /* offset/grad_offset not aligned to 2*sizeof(dtype), or       */
                            /* workspace not 16-byte aligned                               */
};

/* flags for dcnv3_backward_sm100 */
#define DCNV3_BWD_DETERMINISTIC 1u /* bit-reproducible grad_value (fixed-point accumulation) */

DCNV3_API int dcnv3_sm100_abi_version(void);
DCNV3_API const char *dcnv3_sm100_strerror(int code);

/* out[n,ho,wo,g,c] = sum_p mask_p * bilinear(value[n,:,:,g,c], loc_p)
 * (reference kernel: dcnv3_im2col_gpu_kernel, dcnv3_im2col_cuda.cuh:216-275). */
DCNV3_API int dcnv3_forward_sm100(const void *value, const void *offset, const void *mask, void *out,
                        int N, int H, int W, int Ho, int Wo, int G, int gc,
                        int kernel_h, int kernel_w, int stride_h, int stride_w,
                        int pad_h, int pad_w, int dil_h, int dil_w,
                        float offset_scale, int dtype, void *stream /* cudaStream_t */);

/* Scratch the backward needs (fp32 / fixed-point accumulator for grad_value when the I/O dtype
 * is 16-bit or the deterministic flag is set); 0 when none is needed. */
DCNV3_API size_t dcnv3_backward_workspace_bytes(int N, int H, int W, int G, int gc, int dtype,
                                      unsigned flags);

/* grad_value, grad_offset, grad_mask of the forward above for upstream gradient grad_out
 * (reference kernels: dcnv3_col2im_*, dcnv3_im2col_cuda.cuh:82-147,278-839).
 * `workspace` must be device memory of at least dcnv3_backward_workspace_bytes() bytes
 * (may be NULL when that is 0); its contents need not be initialised. */
DCNV3_API int dcnv3_backward_sm100(const void *value, const void *offset, const void *mask,
                         const void *grad_out, void *grad_value, void *grad_offset,
                         void *grad_mask, void *workspace, size_t workspace_bytes,
                         int N, int H, int W, int Ho, int Wo, int G, int gc,
                         int kernel_h, int kernel_w, int stride_h, int stride_w,
                         int pad_h, int pad_w, int dil_h, int dil_w,
                         float offset_scale, int dtype, unsigned flags,
                         void *stream /* cudaStream_t */);

/* ------------------------------------------------------------------------------------------------
 * Host-buffer pipeline: forward + backward of the core with every tensor in HOST memory.
 *
 * In the reference the host<->device traffic of a step is torch's: `imgs.to(device,
 * non_blocking=True)` (train.py:249) before the model, `.cpu()` / `.item()` reads after it
 * (train.py:279-282), with the op called in between (functions/dcnv3_func.py:39-43,53-58), all on
 * one stream, one after the other.  This entry point is what a host-side caller of the DCNv3
 * core binds instead: the batch is cut into chunks of images (images are independent in forward
 * and backward, dcnv3_im2col_cuda.cuh:238,247,315-317) and three streams overlap the H2D copy of
 * chunk c+1, the kernels of chunk c and the D2H copy of chunk c-1 (PCIe is full duplex), across
 * calls as well: _run() only enqueues, _sync() waits.
 *
 *   - host buffers should be page-locked (cudaHostAlloc / torch pin_memory); pageable memory works
 *     but serialises the copies;
 *   - the pipeline owns its device buffers (two slots of `chunk_images` images incl. the backward
 *     scratch), its streams and events; one pipeline per host thread and shape;
 *   - outputs are complete after _sync() returns 0.
 */
typedef struct dcnv3_host_pipeline dcnv3_host_pipeline;

DCNV3_API int dcnv3_host_pipeline_create(dcnv3_host_pipeline **out, int chunk_images,
                        int H, int W, int G, int gc, int kernel_h, int kernel_w,
                        int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                        float offset_scale, int dtype, unsigned flags);

/* forward + backward of N images: h_out, h_grad_value, h_grad_offset, h_grad_mask are written */
DCNV3_API int dcnv3_host_pipeline_run(dcnv3_host_pipeline *p, const void *h_value, const void *h_offset,
                        const void *h_mask, const void *h_grad_out, void *h_out, void *h_grad_value,
                        void *h_grad_offset, void *h_grad_mask, int N);

DCNV3_API int dcnv3_host_pipeline_sync(dcnv3_host_pipeline *p);
DCNV3_API void dcnv3_host_pipeline_destroy(dcnv3_host_pipeline *p);

/* ------------------------------------------------------------------------------------------------
 * The step in front of the sampler: the layer's `offset` and `mask` linears plus the softmax over
 * the K*K points of each group (models/ops_dcnv3/modules/dcnv3.py:330-334: two nn.Linear, reshape,
 * F.softmax, .type(dtype)) as one tcgen05 GEMM with bias, softmax and cast in its epilogue.
 *
 *   x        [M, C]            16-bit activations (M = N*H*W rows), dense
 *   w_cat    [padded_cols, C]  rows 0..2GP-1 = offset.weight, 2GP..3GP-1 = mask.weight, rest zero
 *   bias_cat [padded_cols]     fp32, same order
 *   offset   [M, 2GP], mask [M, GP] (soft-maxed over P within each group), dtype of x
 *
 * Eligible shapes: P == 9, G % 8 == 0, 3GP <= 512, C % 64 == 0, fp16 / bf16; DCNV3_E_SHAPE
 * otherwise (the caller keeps the two linears + softmax for those).
 */
DCNV3_API int dcnv3_offset_mask_proj_padded_cols(int G, int P);
DCNV3_API int dcnv3_offset_mask_proj_sm100(const void *x, const void *w_cat, const float *bias_cat,
                        void *offset, void *mask, long long M, int C, int G, int P, int dtype,
                        void *stream /* cudaStream_t */);

/* ------------------------------------------------------------------------------------------------
 * The producer of the layer's x1: depthwise k x k convolution (stride 1, zero padding (k-1)/2) +
 * LayerNorm over the channels + exact-erf GELU, channels-last in and out, one pass
 * (models/ops_dcnv3/modules/dcnv3.py:276-289,328-329; build_norm_layer :41-62).
 *
 *   x [N,H,W,C] 16-bit;  w_dw [k*k][C] in the dtype of x (= conv.weight[c,0,j,i] at [(j*k+i)*C + c]);
 *   b_dw, gamma, beta [C] fp32;  out [N,H,W,C] in the dtype of x;  conv_out: NULL, or [N,H,W,C]
 *   receiving the convolution's output before the LayerNorm (what a backward pass needs).
 * Eligible: C in {64, 128, 256}, odd k <= 7, fp16 / bf16; DCNV3_E_SHAPE otherwise.
 */
DCNV3_API int dcnv3_dwconv_ln_gelu_sm100(const void *x, const void *w_dw, const float *b_dw,
                        const float *gamma, const float *beta, void *out, void *conv_out,
                        int N, int H, int W, int C, int k, float eps, int dtype,
                        void *stream /* cudaStream_t */);

/* Backward of that producer (k == 3), two passes over the activation tensor instead of the reference's autograd chain
 * (permute / conv backward / permute / LayerNorm backward / GELU backward, modules/dcnv3.py:276-289,328-329):
 *   x, conv_out (as written by the forward), grad_out [N,H,W,C] 16-bit; w_dw [9][C] as above; gamma, beta [C] fp32;
 *   du_scratch [N,H,W,C] 16-bit (the LayerNorm's input gradient, consumed by the second pass); grad_x [N,H,W,C];
 *   grad_params: 12 C floats, written as [grad_w_dw [9][C] | grad_b_dw [C] | grad_gamma [C] | grad_beta [C]] (fp32,
 *   zeroed here, summed with reductions: not bit-reproducible across runs, as the reference's cuDNN wgrad).
 */
DCNV3_API int dcnv3_dwconv_ln_gelu_backward_sm100(const void *x, const void *conv_out, const void *grad_out,
                        const void *w_dw, const float *gamma, const float *beta, void *du_scratch, void *grad_x,
                        float *grad_params, int N, int H, int W, int C, int k, float eps, int dtype,
                        void *stream /* cudaStream_t */);

/* Gradient of the mask logits from the gradient of the soft-maxed masks (modules/dcnv3.py:331-334):
 * grad_logit[r, p] = mask[r, p] * (grad_mask[r, p] - sum_q grad_mask[r, q] mask[r, q]),  r = (pixel, group), fp32 math. */
DCNV3_API int dcnv3_mask_softmax_backward_sm100(const void *grad_mask, const void *mask, void *grad_logit,
                        long long rows, int points, int dtype, void *stream /* cudaStream_t */);

/* The elementwise tail of a hosted Conv block in its inference form, act(conv(x) + bias) with the BatchNorm folded
 * (reference: Conv.forward_fuse, models/common.py:55-66, after models/yolo.py fuse() / utils/torch_utils.py:202-222):
 *   y[r, c] = act(x[r, c] + bias[c]);  x, y [rows, C] channels-last 16-bit, y may alias x; bias [C] fp32;
 *   act: 0 = identity, 1 = SiLU; C a multiple of 8; pointers 16-byte aligned. */
DCNV3_API int dcnv3_bias_act_sm100(const void *x, const float *bias, void *y, long long rows, int C, int act,
                        int dtype, void *stream /* cudaStream_t */);

#ifdef __cplusplus
}
#endif
#endif /* DCNV3_SM100_H_ */
