// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, M = 128, K = 16 per instruction) as a function of N, with the
// A operand in shared memory (SS) or in tensor memory (TS), 128-byte and 64-byte swizzled K-major tiles.  One CTA per SM,
// operands resident in shared memory (no TMA in the loop): this is the tensor pipe's own pace for the model's tile shapes
// (conv2 N = 96, conv3 N = 160, everything else N = 256).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../kiri-ocr_b200/csrc -I../include mma_rate.cu -o bin/mma_rate
#include "common.cuh"

#include <cstdio>

namespace kiri { void set_last_error(const char*, ...) {} }
using namespace kiri;

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

// mode 0: SS, SW128 (KC = 64: 4 MMAs per k-block); 1: SS, SW64 (KC = 32: 2 MMAs per k-block); 2: TS (A in TMEM), B SW128
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int mode, int kblocks, int per_commit, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 16384);
    const int kc = mode == 1 ? 32 : 64;
    const uint64_t layout = mode == 1 ? UMMA_LAYOUT_SW64 : UMMA_LAYOUT_SW128;
    const uint32_t sbo = 8 * kc * 2;
    uint32_t phase = 0;
    const long long t0 = clock64();
    int since = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
      if (elect_one()) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          if (h * 16 < kc) {
            const uint64_t bd = umma_desc_kmajor(b_addr + h * 32, sbo, layout);
            if (mode == 2) umma_bf16_ts(tmem, tmem + 256 + h * 8, bd, idesc, (kb | h) != 0);
            else umma_bf16(tmem, umma_desc_kmajor(a_addr + h * 32, sbo, layout), bd, idesc, (kb | h) != 0);
          }
        }
      }
      __syncwarp();
      if (++since == per_commit || kb + 1 == kblocks) {
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, phase);                    // like a pipeline stage that is released by the MMAs' completion
        phase ^= 1;
        since = 0;
      }
    }
    const long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int kblocks = 4096;
  const char* names[3] = {"SS  A,B SW128 (4 MMAs per k-block)", "SS  A,B SW64  (2 MMAs per k-block)", "TS  A in TMEM, B SW128          "};
  for (int mode = 0; mode < 3; ++mode)
    for (int per_commit : {1, 8}) {
      for (int N : {32, 64, 96, 128, 160, 192, 256}) {
        long long h = 0;
        for (int rep = 0; rep < 2; ++rep) {
          mma_rate_kernel<<<148, 128, 50 * 1024>>>(N, mode, kblocks, per_commit, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
        const int mmas = kblocks * (mode == 1 ? 2 : 4);
        printf("%s  N=%3d  commit+wait every %d k-block(s): %7.1f cycles per MMA (ideal 128*N/256 = %3d)  %6.0f MAC/clk/SM\n", names[mode], N,
               per_commit, (double)h / mmas, 128 * N / 256, 128.0 * N * 16 * mmas / h);
      }
    }
  return 0;
}
