// dcnv3_hosting.cu -- the elementwise tail of a hosted block in its INFERENCE form (BASELINE configs[2]).
//
// The reference folds every BatchNorm into the convolution in front of it before val / detect (models/yolo.py fuse() over
// utils/torch_utils.py:202-222) and runs Conv.forward_fuse = act(conv(x)) (models/common.py:55-66).  In eager PyTorch the
// folded bias then costs its own pass: cuDNN's convolution does not take a bias for channels-last 16-bit tensors, ATen adds
// it with a broadcast `elementwise_kernel<128, 4>` (49 us per call on YOLOv5l at batch 32: 29 % of the GPU time of the
// whole forward, scripts/infer_prof.py) and SiLU is a further read + write of the same tensor.  One pass here:
//     y[r, c] = act(x[r, c] + bias[c]),   x, y: [rows, C] channels-last 16-bit (y may be x), bias fp32, act = identity | SiLU
// 128-bit loads / stores, fp32 math, the bias chunk of a lane read once per grid stride when C * 2 divides the stride.
#include "dcnv3_sm100.h"

#include "dcnv3_common.cuh"
#include "dcnv3_launch.h"

#include <algorithm>

namespace dcnv3 {
namespace host {

template <typename T, int ACT>
__global__ void __launch_bounds__(256)
bias_act(const uint4 *__restrict__ x, const float *__restrict__ bias, uint4 *__restrict__ y, long long chunks /* rows * C / 8 */,
         int chunks_per_row /* C / 8 */) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += stride) {
        const int c0 = (int)(i % chunks_per_row) * 8;
        const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bias + c0)), b1 = __ldg(reinterpret_cast<const float4 *>(bias + c0 + 4));
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float v[8];
        unpack<T>(x[i], v);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float t = v[e] + b[e];
            // SiLU t / (1 + exp(-t)): __expf / __fdividef are accurate to a few fp32 ulps, far below the 16-bit output rounding
            v[e] = ACT == 1 ? __fdividef(t, 1.f + __expf(-t)) : t;
        }
        y[i] = pack<T>(v);
    }
}

template <typename T>
static int launch(const void *x, const float *bias, void *y, long long rows, int C, int act, cudaStream_t stream) {
    const long long chunks = rows * (C / 8);
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int blocks = (int)std::min<long long>((chunks + 255) / 256, (long long)num_sms * 8);
    const uint4 *xi = static_cast<const uint4 *>(x);
    uint4 *yo = static_cast<uint4 *>(y);
    if (act == 1) bias_act<T, 1><<<blocks, 256, 0, stream>>>(xi, bias, yo, chunks, C / 8);
    else bias_act<T, 0><<<blocks, 256, 0, stream>>>(xi, bias, yo, chunks, C / 8);
    return (int)cudaGetLastError();
}

}  // namespace host
}  // namespace dcnv3

extern "C" int dcnv3_bias_act_sm100(const void *x, const float *bias, void *y, long long rows, int C, int act, int dtype,
                                    void *stream) {
    if (dtype != DCNV3_F16 && dtype != DCNV3_BF16) return DCNV3_E_DTYPE;
    if (rows < 0 || C <= 0 || C % 8 != 0 || (act != 0 && act != 1)) return DCNV3_E_SHAPE;
    if (rows == 0) return DCNV3_OK;
    if (!x || !bias || !y) return DCNV3_E_NULL;
    if (((uintptr_t)x | (uintptr_t)y | (uintptr_t)bias) % 16) return DCNV3_E_ALIGN;
    if (dtype == DCNV3_F16) return dcnv3::host::launch<__half>(x, bias, y, rows, C, act, (cudaStream_t)stream);
    return dcnv3::host::launch<__nv_bfloat16>(x, bias, y, rows, C, act, (cudaStream_t)stream);
}
