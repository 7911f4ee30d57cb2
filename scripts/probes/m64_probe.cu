// Probe (development aid, not part of the library): where does tcgen05.mma.cta_group::1 with M = 64 put its 64
// accumulator rows in tensor memory, and may the D address carry a lane offset of 16 (two M = 64 blocks in one
// 16-column tile)?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -I yolo_somi_b200/csrc -o scripts/probes/_bin/m64_probe scripts/probes/m64_probe.cu
#include <cstdio>
#include <cuda_bf16.h>
#include "dcnv3_tc.cuh"
using namespace dcnv3;
using namespace dcnv3::tc;

__global__ void probe(float *out, int lane_off, int mode) {
    __shared__ __align__(1024) unsigned char a_sm[128 * 128];     // K-major SW128: 128 rows x 128 B
    __shared__ __align__(1024) unsigned char z_sm[128 * 128];
    __shared__ __align__(128) unsigned char b_sm[2][512];         // [set][c8 (2)][16 px][16 B]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * 128 / 4; i += 128) { ((uint32_t *)a_sm)[i] = 0; ((uint32_t *)z_sm)[i] = 0; }
    for (int i = tid; i < 256; i += 128) ((uint32_t *)b_sm)[i] = 0;
    __syncthreads();
    if (tid < 128) {   // A[r][0] = r + 1
        const int r = tid;
        *(__nv_bfloat16 *)(a_sm + r * 128 + ((0 ^ (r & 7)) << 4)) = __float2bfloat16((float)(r + 1));
    }
    if (tid < 16) {
        const int n = tid;
        *(__nv_bfloat16 *)(b_sm[0] + (n >> 3) * 256 + 0 * 16 + (n & 7) * 2) = __float2bfloat16((float)(n + 1));
        *(__nv_bfloat16 *)(b_sm[1] + (n >> 3) * 256 + 0 * 16 + (n & 7) * 2) = __float2bfloat16(100.f * (float)(n + 1));
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base_s;
    if (tid == 0) {
        const uint32_t id128 = umma_idesc(1, 128, 16), id64 = umma_idesc(1, 64, 16);
        // clear columns 0..31 of all 128 lanes
        tc_mma(tb, umma_desc_k_sw128(smem_u32(z_sm)), umma_desc_mn_plain(smem_u32(b_sm[0]), 128, 256), id128, 0);
        tc_mma(tb + 16, umma_desc_k_sw128(smem_u32(z_sm)), umma_desc_mn_plain(smem_u32(b_sm[0]), 128, 256), id128, 0);
        // M = 64 block at lane 0, columns 0..15
        tc_mma(tb, umma_desc_k_sw128(smem_u32(a_sm)), umma_desc_mn_plain(smem_u32(b_sm[0]), 128, 256), id64, 1);
        if (mode >= 1)   // second M = 64 block (rows 64..127 of A, x100) at a lane offset, columns 0..15 as well
            tc_mma(tb + ((uint32_t)lane_off << 16), umma_desc_k_sw128(smem_u32(a_sm) + 64 * 128),
                   umma_desc_mn_plain(smem_u32(b_sm[1]), 128, 256), id64, 1);
        if (mode >= 2)   // and one in columns 16..31 at the lane offset only
            tc_mma(tb + ((uint32_t)lane_off << 16) + 16, umma_desc_k_sw128(smem_u32(a_sm)),
                   umma_desc_mn_plain(smem_u32(b_sm[1]), 128, 256), id64, 1);
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    float r[16];
    for (int cb = 0; cb < 2; ++cb) {
        VMMA_TMEM_LD_16(tb + ((uint32_t)(warp * 32) << 16) + cb * 16, r);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[(tid * 2 + cb) * 16 + j] = r[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "n"(64) : "memory");
}

int main() {
    float *d, h[128 * 32];
    cudaMalloc(&d, sizeof(h));
    for (int mode = 0; mode < 3; ++mode)
        for (int lo : {16}) {
            if (mode == 0 && lo != 16) continue;
            cudaMemset(d, 0, sizeof(h));
            probe<<<1, 128>>>(d, lo, mode);
            cudaError_t e = cudaDeviceSynchronize();
            printf("mode %d lane_off %d: %s\n", mode, lo, cudaGetErrorString(e));
            if (e != cudaSuccess) return 1;
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            for (int cb = 0; cb < 2; ++cb) {
                printf(" cols %d..: ", cb * 16);
                for (int l = 0; l < 128; ++l) {
                    const float v0 = h[(l * 2 + cb) * 16 + 0], v15 = h[(l * 2 + cb) * 16 + 15];
                    if (v0 != 0.f || v15 != 0.f) printf("L%d=%g(%g) ", l, v0, v15 / 16.f);
                }
                printf("\n");
            }
        }
    return 0;
}
