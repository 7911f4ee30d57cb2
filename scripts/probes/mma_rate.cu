// Probe (development aid): issue-to-completion rate of small tcgen05.mma shapes (SS mode, K-major SW128 A, MN-major B,
// N = 16, K = 16): cycles per product for M = 64 / 128, same or rotating accumulator blocks, same or rotating A tiles.
#include <cstdio>
#include <cuda_bf16.h>
#include "dcnv3_tc.cuh"
using namespace dcnv3;
using namespace dcnv3::tc;

__global__ void rate(long long *out, int M, int n_acc, int n_a, int iters, int N) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    unsigned char *base = sm + ((1024u - (smem_u32(sm) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t *)base)[i] = 0;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = umma_idesc(1, M, N);
        const uint32_t a0 = smem_u32(base), b0 = a0 + 128 * 1024;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t d = tb + (uint32_t)((it % n_acc) * N);
            const uint32_t a = a0 + (uint32_t)(it % n_a) * 16384u + (uint32_t)((it / n_a) & 3) * 32u;
            tc_mma(d, umma_desc_k_sw128(a), umma_desc_mn_plain(b0 + ((it & 3) * 256), 128, 1024), idesc, 1);
        }
        const long long t1 = clock64();
        tc_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "n"(512) : "memory");
}

int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
    const int iters = 2048;
    for (int N : {16, 32, 64})
    for (int M : {64, 128})
        for (int n_acc : {1, 4, 16})
            for (int n_a : {1, 8}) {
                if (n_acc * N > 512) continue;
                rate<<<1, 128, 170 * 1024>>>(d, M, n_acc, n_a, iters, N);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("N %3d M %3d accumulators %2d A tiles %d: issue %.1f cycles / product, complete %.1f cycles / product\n", N, M, n_acc, n_a,
                       (double)h[0] / iters, (double)h[1] / iters);
            }
    return 0;
}
