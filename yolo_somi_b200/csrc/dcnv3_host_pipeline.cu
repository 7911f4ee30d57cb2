// dcnv3_host_pipeline.cu -- host-buffer entry points of the C ABI (include/dcnv3_sm100.h):
// forward + backward of the DCNv3 core with all tensors in host memory, chunked over the batch
// and software-pipelined over three streams (H2D | kernels | D2H).
//
// Reference call sites this stands in for: the host->device copy of a step's inputs
// (train.py:249), the op (models/ops_dcnv3/functions/dcnv3_func.py:39-43,53-58) and the
// device->host reads of its results, which the reference runs back to back on one stream.
#include "dcnv3_sm100.h"

#include "dcnv3_launch.h"

#include <new>

namespace {

constexpr int kSlots = 3;   // H2D of chunk c + 1 must not wait for the D2H of chunk c - 1: with two slots every chunk paid the kernels' time as a bubble on the copy engines (4.45 ms per step against 3.96 ms of bare copies)

struct Slot {
    void *value = nullptr, *offset = nullptr, *mask = nullptr, *grad_out = nullptr;
    void *out = nullptr, *grad_value = nullptr, *grad_offset = nullptr, *grad_mask = nullptr;
    void *workspace = nullptr;
    cudaEvent_t loaded = nullptr;     // inputs of the chunk are on the device
    cudaEvent_t computed = nullptr;   // kernels of the chunk are done
    cudaEvent_t drained = nullptr;    // outputs of the chunk are back on the host: slot is free
    bool used = false;
};

}  // namespace

struct dcnv3_host_pipeline {
    dcnv3::Geom q;          // N = chunk capacity
    int dtype;
    unsigned flags;
    size_t es;              // element size
    size_t v_img, o_img, m_img, y_img;   // elements per image: value, offset, mask, out
    size_t ws_bytes;
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    Slot slot[kSlots];
    unsigned long long chunks = 0;       // chunks enqueued so far (slot = chunks % kSlots)
    int device = 0;
};

namespace {

int conv_out(int in, int pad, int dil, int k, int stride) {
    return (in + 2 * pad - (dil * (k - 1) + 1)) / stride + 1;
}

void release(dcnv3_host_pipeline *p) {
    if (!p) return;
    for (Slot &s : p->slot) {
        void *bufs[] = {s.value, s.offset, s.mask, s.grad_out, s.out, s.grad_value, s.grad_offset, s.grad_mask, s.workspace};
        for (void *b : bufs)
            if (b) cudaFree(b);
        if (s.loaded) cudaEventDestroy(s.loaded);
        if (s.computed) cudaEventDestroy(s.computed);
        if (s.drained) cudaEventDestroy(s.drained);
    }
    if (p->s_in) cudaStreamDestroy(p->s_in);
    if (p->s_run) cudaStreamDestroy(p->s_run);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    delete p;
}

#define PIPE_TRY(expr)                        \
    do {                                      \
        cudaError_t e_ = (expr);              \
        if (e_ != cudaSuccess) return (int)e_; \
    } while (0)

}  // namespace

extern "C" {

int dcnv3_host_pipeline_create(dcnv3_host_pipeline **out, int chunk_images, int H, int W, int G, int gc,
                               int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                               int dil_h, int dil_w, float offset_scale, int dtype, unsigned flags) {
    if (!out) return DCNV3_E_NULL;
    *out = nullptr;
    if (dtype != DCNV3_F32 && dtype != DCNV3_F16 && dtype != DCNV3_BF16) return DCNV3_E_DTYPE;
    if (chunk_images <= 0 || H <= 0 || W <= 0 || G <= 0 || gc <= 0 || kernel_h <= 0 || kernel_w <= 0 ||
        stride_h <= 0 || stride_w <= 0 || dil_h <= 0 || dil_w <= 0 || pad_h < 0 || pad_w < 0)
        return DCNV3_E_SHAPE;
    const int Ho = conv_out(H, pad_h, dil_h, kernel_h, stride_h), Wo = conv_out(W, pad_w, dil_w, kernel_w, stride_w);
    if (Ho <= 0 || Wo <= 0) return DCNV3_E_SHAPE;
    if ((long long)H * W * G * gc >= (1LL << 31) || (long long)Ho * Wo * G * kernel_h * kernel_w * 2 >= (1LL << 31))
        return DCNV3_E_TOO_LARGE;
    dcnv3_host_pipeline *p = new (std::nothrow) dcnv3_host_pipeline();
    if (!p) return (int)cudaErrorMemoryAllocation;
    p->q = dcnv3::Geom{chunk_images, H, W, Ho, Wo, G, gc, kernel_h, kernel_w, stride_h, stride_w,
                       pad_h, pad_w, dil_h, dil_w, offset_scale};
    p->dtype = dtype;
    p->flags = flags;
    p->es = dtype == DCNV3_F32 ? 4 : 2;
    const size_t P = (size_t)kernel_h * kernel_w;
    p->v_img = (size_t)H * W * G * gc;
    p->y_img = (size_t)Ho * Wo * G * gc;
    p->o_img = (size_t)Ho * Wo * G * P * 2;
    p->m_img = (size_t)Ho * Wo * G * P;
    p->ws_bytes = dcnv3::backward_workspace_bytes(p->q, dtype, flags);
    cudaError_t e = cudaGetDevice(&p->device);
    auto fail = [&](cudaError_t err) { release(p); return (int)err; };
    if (e != cudaSuccess) return fail(e);
    if ((e = cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking)) != cudaSuccess) return fail(e);
    if ((e = cudaStreamCreateWithFlags(&p->s_run, cudaStreamNonBlocking)) != cudaSuccess) return fail(e);
    if ((e = cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking)) != cudaSuccess) return fail(e);
    const size_t n = (size_t)chunk_images;
    for (Slot &s : p->slot) {
        struct { void **ptr; size_t bytes; } allocs[] = {
            {&s.value, n * p->v_img * p->es}, {&s.offset, n * p->o_img * p->es}, {&s.mask, n * p->m_img * p->es},
            {&s.grad_out, n * p->y_img * p->es}, {&s.out, n * p->y_img * p->es}, {&s.grad_value, n * p->v_img * p->es},
            {&s.grad_offset, n * p->o_img * p->es}, {&s.grad_mask, n * p->m_img * p->es}, {&s.workspace, p->ws_bytes}};
        for (auto &a : allocs)
            if (a.bytes && (e = cudaMalloc(a.ptr, a.bytes)) != cudaSuccess) return fail(e);
        if ((e = cudaEventCreateWithFlags(&s.loaded, cudaEventDisableTiming)) != cudaSuccess) return fail(e);
        if ((e = cudaEventCreateWithFlags(&s.computed, cudaEventDisableTiming)) != cudaSuccess) return fail(e);
        if ((e = cudaEventCreateWithFlags(&s.drained, cudaEventDisableTiming)) != cudaSuccess) return fail(e);
    }
    *out = p;
    return DCNV3_OK;
}

int dcnv3_host_pipeline_run(dcnv3_host_pipeline *p, const void *h_value, const void *h_offset, const void *h_mask,
                            const void *h_grad_out, void *h_out, void *h_grad_value, void *h_grad_offset,
                            void *h_grad_mask, int N) {
    if (!p) return DCNV3_E_NULL;
    if (N < 0) return DCNV3_E_SHAPE;
    if (N == 0) return DCNV3_OK;
    if (!h_value || !h_offset || !h_mask || !h_grad_out || !h_out || !h_grad_value || !h_grad_offset || !h_grad_mask)
        return DCNV3_E_NULL;
    const size_t es = p->es;
    auto at = [es](const void *base, size_t img_elems, int img) {
        return static_cast<const char *>(base) + (size_t)img * img_elems * es;
    };
    auto at_w = [es](void *base, size_t img_elems, int img) {
        return static_cast<char *>(base) + (size_t)img * img_elems * es;
    };
    for (int n0 = 0; n0 < N; n0 += p->q.N) {
        const int n = N - n0 < p->q.N ? N - n0 : p->q.N;
        Slot &s = p->slot[p->chunks % kSlots];
        ++p->chunks;
        // ---- H2D (waits until the previous user of the slot has been drained to the host)
        if (s.used) PIPE_TRY(cudaStreamWaitEvent(p->s_in, s.drained, 0));
        PIPE_TRY(cudaMemcpyAsync(s.value, at(h_value, p->v_img, n0), n * p->v_img * es, cudaMemcpyHostToDevice, p->s_in));
        PIPE_TRY(cudaMemcpyAsync(s.offset, at(h_offset, p->o_img, n0), n * p->o_img * es, cudaMemcpyHostToDevice, p->s_in));
        PIPE_TRY(cudaMemcpyAsync(s.mask, at(h_mask, p->m_img, n0), n * p->m_img * es, cudaMemcpyHostToDevice, p->s_in));
        PIPE_TRY(cudaMemcpyAsync(s.grad_out, at(h_grad_out, p->y_img, n0), n * p->y_img * es, cudaMemcpyHostToDevice, p->s_in));
        PIPE_TRY(cudaEventRecord(s.loaded, p->s_in));
        // ---- kernels
        dcnv3::Geom q = p->q;
        q.N = n;
        PIPE_TRY(cudaStreamWaitEvent(p->s_run, s.loaded, 0));
        PIPE_TRY(dcnv3::launch_forward(s.value, s.offset, s.mask, s.out, q, p->dtype, p->s_run));
        PIPE_TRY(dcnv3::launch_backward(s.value, s.offset, s.mask, s.grad_out, s.grad_value, s.grad_offset,
                                        s.grad_mask, s.workspace, q, p->dtype, p->flags, p->s_run));
        PIPE_TRY(cudaEventRecord(s.computed, p->s_run));
        // ---- D2H
        PIPE_TRY(cudaStreamWaitEvent(p->s_out, s.computed, 0));
        PIPE_TRY(cudaMemcpyAsync(at_w(h_out, p->y_img, n0), s.out, n * p->y_img * es, cudaMemcpyDeviceToHost, p->s_out));
        PIPE_TRY(cudaMemcpyAsync(at_w(h_grad_value, p->v_img, n0), s.grad_value, n * p->v_img * es, cudaMemcpyDeviceToHost, p->s_out));
        PIPE_TRY(cudaMemcpyAsync(at_w(h_grad_offset, p->o_img, n0), s.grad_offset, n * p->o_img * es, cudaMemcpyDeviceToHost, p->s_out));
        PIPE_TRY(cudaMemcpyAsync(at_w(h_grad_mask, p->m_img, n0), s.grad_mask, n * p->m_img * es, cudaMemcpyDeviceToHost, p->s_out));
        PIPE_TRY(cudaEventRecord(s.drained, p->s_out));
        s.used = true;
    }
    return DCNV3_OK;
}

int dcnv3_host_pipeline_sync(dcnv3_host_pipeline *p) {
    if (!p) return DCNV3_E_NULL;
    PIPE_TRY(cudaStreamSynchronize(p->s_in));
    PIPE_TRY(cudaStreamSynchronize(p->s_run));
    PIPE_TRY(cudaStreamSynchronize(p->s_out));
    return DCNV3_OK;
}

void dcnv3_host_pipeline_destroy(dcnv3_host_pipeline *p) {
    if (!p) return;
    cudaStreamSynchronize(p->s_in);
    cudaStreamSynchronize(p->s_run);
    cudaStreamSynchronize(p->s_out);
    release(p);
}

}  // extern "C"
