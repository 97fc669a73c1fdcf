// Development aid: cycles per tcgen05.mma (kind::f16, M = 128, K = 16, both operands from shared memory, no swizzle)
// as a function of N, of how many warps issue, and of the A-operand access pattern.  Answers "what does one MMA of a
// layer program cost" for the cost model in DESIGN.md 4.1.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I nanocaller_b200/csrc -o /tmp/umma_rate tools/umma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "nc_cnn_tc.cuh"

using namespace nc;

// mode bit0: 4 issuing warps (else 1); bit1: A start address walks (tap shifts) instead of staying put;
// bit2: the two K groups of A are 6656 B apart (another plane) instead of adjacent-ish
__global__ void __launch_bounds__(128, 1) rate_kernel(int M, int N, int iters, int mode, long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    const int nissue = (mode & 1) ? 4 : 1;
    if (t == 0) { mbar_init(&bar, nissue); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t a16 = smem_u32(smem) >> 4, b16 = smem_u32(smem + 96 * 1024) >> 4;
    const uint32_t idesc = make_idesc_f16(M, N);
    const uint32_t a_lbo = (mode & 4) ? 416u : 105u;
    long long t0 = clock64();
    if (warp < nissue && elect_one()) {
        const uint32_t d = tmem + warp * 128;
        const uint32_t blo = b16 | ((uint32_t)N << 16);
        const int n = iters / nissue;
        for (int i = 0; i < n; i += 8) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t sh = (mode & 2) ? (uint32_t)(j * 45 + warp * 7) : 0u;
                umma_f16(d, sdesc16((a16 + sh) | (a_lbo << 16)), sdesc16(blo + j * 2 * N), idesc, 1u);
            }
        }
        umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    tc_fence_after();
    if (t == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 4096;
  for (int M : {128, 64}) {
    printf("M=%d K=16 kind::f16 SS no-swizzle; cycles per MMA (148 CTAs, one per SM)\n", M);
    printf("%5s %10s %10s %10s %10s %10s\n", "N", "1w/fixed", "4w/fixed", "1w/walk", "4w/walk", "4w/walk/far");
    for (int N : {8, 16, 32, 48, 64, 96, 128}) {
        printf("%5d", N);
        for (int mode : {0, 1, 2, 3, 7}) {
            long long h = 0;
            for (int rep = 0; rep < 2; rep++) {
                rate_kernel<<<148, 128, 200 * 1024>>>(M, N, iters, mode, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf(" err %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            }
            printf(" %10.1f", (double)h / iters);
        }
        printf("\n");
    }
  }
    return 0;
}
