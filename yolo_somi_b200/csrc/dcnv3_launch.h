// dcnv3_launch.h -- internal host-side entry points shared by the C ABI (dcnv3_capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "dcnv3_common.cuh"

namespace dcnv3 {

constexpr size_t kWorkspaceHeader = 256;  // bytes reserved at the start of the backward scratch

cudaError_t launch_forward(const void *value, const void *offset, const void *mask, void *out,
                           const Geom &q, int dtype, cudaStream_t stream);

size_t backward_workspace_bytes(const Geom &q, int dtype, unsigned flags);

cudaError_t launch_backward(const void *value, const void *offset, const void *mask,
                            const void *grad_out, void *grad_value, void *grad_offset,
                            void *grad_mask, void *workspace, const Geom &q, int dtype,
                            unsigned flags, cudaStream_t stream);

}  // namespace dcnv3
