// dcnv3_launch.h -- internal host-side entry points shared by the C ABI (dcnv3_capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include <stddef.h>

#include "dcnv3_common.cuh"

namespace dcnv3 {

constexpr size_t kWorkspaceHeader = 256;  // bytes reserved at the start of the backward scratch

cudaError_t launch_forward(const void *value, const void *offset, const void *mask, void *out,
                           const Geom &q, int dtype, cudaStream_t stream);

// shared-memory tiled forward (dcnv3_forward_tile.cu); false = shape not eligible, use the direct kernel
bool try_launch_forward_tile(const void *value, const void *offset, const void *mask, void *out,
                             const Geom &q, int dtype, bool fast, cudaStream_t stream, cudaError_t *err);
bool fast_weights_requested();
// group-slice forward (dcnv3_forward_gs.cu): 16-bit I/O, group_channels == 16, G % 8 == 0, 3x3 / stride 1
bool try_launch_forward_gs(const void *value, const void *offset, const void *mask, void *out,
                           const Geom &q, int dtype, bool fast, cudaStream_t stream, cudaError_t *err);
// Kernels under csrc/experiments/ are measured alternatives that lost to the defaults (profiles/README.md); they are
// compiled only with DCNV3_BUILD_EXPERIMENTS=1 (-DDCNV3_EXPERIMENTS) and otherwise decline every shape.
// tensor-core forward for 16-bit I/O, group_channels == 16 (experiments/dcnv3_forward_mma.cu)
#ifdef DCNV3_EXPERIMENTS
bool try_launch_forward_mma(const void *value, const void *offset, const void *mask, void *out,
                            const Geom &q, int dtype, cudaStream_t stream, cudaError_t *err);
#else
inline bool try_launch_forward_mma(const void *value, const void *offset, const void *mask, void *out,
                            const Geom &q, int dtype, cudaStream_t stream, cudaError_t *err) { return false; }
#endif

// shared-memory tiled backward (dcnv3_backward_tile.cu); gv_acc = zeroed fp32 accumulator
#ifdef DCNV3_EXPERIMENTS
bool try_launch_backward_tile(const void *value, const void *offset, const void *mask,
                              const void *grad_out, float *gv_acc, void *grad_offset,
                              void *grad_mask, const Geom &q, int dtype, cudaStream_t stream,
                              cudaError_t *err);
#else
inline bool try_launch_backward_tile(const void *value, const void *offset, const void *mask,
                              const void *grad_out, float *gv_acc, void *grad_offset,
                              void *grad_mask, const Geom &q, int dtype, cudaStream_t stream,
                              cudaError_t *err) { return false; }
#endif

// tensor-core backward for 16-bit I/O, group_channels == 16 (dcnv3_backward_mma.cu)
bool try_launch_backward_mma(const void *value, const void *offset, const void *mask,
                             const void *grad_out, float *gv_acc, void *grad_offset, void *grad_mask,
                             const Geom &q, int dtype, cudaStream_t stream, cudaError_t *err);

// tensor-core backward with register-resident accumulators (dcnv3_backward_strip.cu): 16-bit I/O,
// group_channels == 16, 3x3 / stride 1 / dilation 1
bool try_launch_backward_strip(const void *value, const void *offset, const void *mask,
                               const void *grad_out, float *gv_acc, void *grad_offset, void *grad_mask,
                               const Geom &q, int dtype, cudaStream_t stream, cudaError_t *err);

// split backward: grad_offset / grad_mask (dcnv3_backward_dots.cu) + grad_value (value-only strip kernel)
// (`zero_plane` / `zero_bytes`: if given, the kernel also zeroes that buffer -- the value kernel's fp32 plane --
// a slice per CTA while the CTA waits for its inputs)
bool try_launch_backward_dots(const void *value, const void *offset, const void *mask, const void *grad_out,
                              void *grad_offset, void *grad_mask, const Geom &q, int dtype, cudaStream_t stream,
                              cudaError_t *err, float *zero_plane = nullptr, size_t zero_bytes = 0);
bool try_launch_backward_vstrip(const void *offset, const void *mask, const void *grad_out, float *gv_acc,
                                const Geom &q, int dtype, cudaStream_t stream, cudaError_t *err);
// grad_value as a tcgen05 product with TMEM accumulators (dcnv3_backward_vmma.cu)
bool backward_vmma_eligible(const void *offset, const void *mask, const void *grad_out, const float *gv_acc, const Geom &q);
// (`cond` / `thr`: if given, the kernel runs only when *cond > thr -- read on the device after the preceding kernel)
bool try_launch_backward_vmma(const void *offset, const void *mask, const void *grad_out, float *gv_acc,
                              const Geom &q, int dtype, cudaStream_t stream, cudaError_t *err,
                              const unsigned long long *cond = nullptr, unsigned long long thr = 0);

// grad_value with the accumulator resident in tensor memory, written once in the I/O dtype (dcnv3_backward_vres.cu);
// `scratch`: backward_vres_scratch_bytes(q) bytes (one 16-bit far-point mask per pixel and group)
bool backward_vres_eligible(const void *offset, const void *mask, const void *grad_out, const void *grad_value, const Geom &q);
size_t backward_vres_scratch_bytes(const Geom &q);
// the size heuristic: is it also the faster form for this shape (enough patch rows per SM, rows of at least four patches)
bool backward_vres_preferred(const Geom &q);
// `counter`: 64-bit far-point count, zeroed by the caller's previous kernel; `*thr`: the count above which the result is
// NOT final (too many far points for the 16-bit atomics: the caller then runs the plane form, conditionally, on top)
bool try_launch_backward_vres(const void *offset, const void *mask, const void *grad_out, void *grad_value, void *scratch,
                              size_t scratch_bytes, unsigned long long *counter, unsigned long long *thr,
                              const Geom &q, int dtype, cudaStream_t stream, cudaError_t *err);

// grad_value with the coefficient columns built dense in registers and a circular TMEM band (dcnv3_backward_vband.cu)
#ifdef DCNV3_EXPERIMENTS
bool try_launch_backward_vband(const void *offset, const void *mask, const void *grad_out, float *gv_acc,
                               const Geom &q, int dtype, cudaStream_t stream, cudaError_t *err);
#else
inline bool try_launch_backward_vband(const void *offset, const void *mask, const void *grad_out, float *gv_acc,
                               const Geom &q, int dtype, cudaStream_t stream, cudaError_t *err) { return false; }
#endif

#ifdef DCNV3_EXPERIMENTS
bool try_launch_backward_mma2(const void *value, const void *offset, const void *mask,
                              const void *grad_out, float *gv_acc, void *grad_offset, void *grad_mask,
                              const Geom &q, int dtype, cudaStream_t stream, cudaError_t *err);
#else
inline bool try_launch_backward_mma2(const void *value, const void *offset, const void *mask,
                              const void *grad_out, float *gv_acc, void *grad_offset, void *grad_mask,
                              const Geom &q, int dtype, cudaStream_t stream, cudaError_t *err) { return false; }
#endif

// Programmatic dependent launch (PDL) between the kernels of the split backward: a kernel launched with
// `pdl_launch` may start (prologue, loads of tensors no earlier kernel of the call writes) while the previous
// kernel of the stream drains; it executes griddepcontrol.wait before it touches what that kernel produced.
// DCNV3_PDL=0 launches everything fully serialised.  Measured (profiles/README.md): -3 us per backward pass for
// group_channels 16 / 32, but +25 us for group_channels == 8 (cfg5, G = 32), where it therefore stays off.
inline bool pdl_for(const Geom &q) { return q.gc != 8; }
inline bool pdl_enabled() {
    static const bool on = [] { const char *e = std::getenv("DCNV3_PDL"); return !(e && e[0] == '0'); }();
    return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t pdl_launch(bool allow, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = allow && pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// fp64 I/O, all arithmetic in double (dcnv3_f64.cu): the reference's own test script drives the extension in double
cudaError_t launch_forward_f64(const void *value, const void *offset, const void *mask, void *out, const Geom &q,
                               cudaStream_t stream);
cudaError_t launch_backward_f64(const void *value, const void *offset, const void *mask, const void *grad_out,
                                void *grad_value, void *grad_offset, void *grad_mask, const Geom &q, cudaStream_t stream);

size_t backward_workspace_bytes(const Geom &q, int dtype, unsigned flags);

cudaError_t launch_backward(const void *value, const void *offset, const void *mask,
                            const void *grad_out, void *grad_value, void *grad_offset,
                            void *grad_mask, void *workspace, const Geom &q, int dtype,
                            unsigned flags, cudaStream_t stream);

}  // namespace dcnv3
