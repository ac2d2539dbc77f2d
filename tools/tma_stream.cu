// Micro-benchmark: how fast can ONE thread per SM stream L2-resident operand tiles into shared memory
// with TMA?  (Decides the stage geometry of the fused FFN kernel: per-SM bytes/clk for 128-byte vs
// 64-byte rows, strided vs contiguous boxes, ring depth.)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../kiri-ocr_b200/csrc -I../include \
//        tma_stream.cu -o tma_stream -lcuda && ./tma_stream
#include "common.cuh"

#include <cstdio>
#include <vector>

namespace kiri { void set_last_error(const char*, ...) {} }
using namespace kiri;

struct Bars { uint64_t full[8]; uint64_t empty[8]; };

// mode 0: tensor-map boxes; mode 1: plain bulk copies of contiguous `unit_bytes` blocks
__global__ void __launch_bounds__(160, 1)
stream_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t* __restrict__ flat, int mode, int stages,
              int boxes_per_stage, int box_bytes, int box_rows, int n_row_blocks, int n_chunks, int units,
              int distinct, int producers, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ Bars bars;
  const int stage_bytes = boxes_per_stage * box_bytes;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.empty[s], 1); }
    fence_mbar_init();
    tma_prefetch_desc(&tm);
  }
  __syncthreads();
  const long long t0 = clock64();
  if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < producers) {
    const int pw = threadIdx.x >> 5;
    // division-free walk: this producer owns units pw, pw + P, ...; stage/phase/coordinates advance incrementally
    int stage = pw % stages; uint32_t phase = (pw / stages) & 1;
    int rb = (pw * boxes_per_stage) % n_row_blocks, ck = ((pw * boxes_per_stage) / n_row_blocks) % n_chunks;
    const int rb_step = ((producers - 1) * boxes_per_stage) % n_row_blocks;
    const int ck_step = ((producers - 1) * boxes_per_stage) / n_row_blocks;
    const uint32_t full0 = smem_u32(&bars.full[0]);
    for (int it = pw; it < units; it += producers) {
      mbar_wait(&bars.empty[stage], phase ^ 1);
      mbar_arrive_expect_tx(&bars.full[stage], stage_bytes);
      uint8_t* dst = smem + stage * stage_bytes;
      for (int b = 0; b < boxes_per_stage; ++b) {
        if (mode == 0) {
          tma_load_3d(dst, &tm, &bars.full[stage], 0, ck, rb * box_rows);
        } else {
          const uint8_t* src = flat + static_cast<size_t>(ck * n_row_blocks + rb) * box_bytes;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(dst)), "l"(src), "r"(box_bytes), "r"(full0 + stage * 8) : "memory");
        }
        dst += box_bytes;
        if (++rb == n_row_blocks) { rb = 0; if (++ck == n_chunks) ck = 0; }
      }
      rb += rb_step; ck += ck_step;
      if (rb >= n_row_blocks) { rb -= n_row_blocks; ++ck; }
      while (ck >= n_chunks) ck -= n_chunks;
      stage += producers;
      while (stage >= stages) { stage -= stages; phase ^= 1; }
    }
  } else if (threadIdx.x == 128) {
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < units; ++it) {
      mbar_wait(&bars.full[stage], phase);
      mbar_arrive(&bars.empty[stage]);
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(p);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const size_t buf_bytes = 8u << 20;
  uint8_t* buf; cudaMalloc(&buf, buf_bytes); cudaMemset(buf, 1, buf_bytes);
  long long* cyc; cudaMalloc(&cyc, 8);
  struct Cfg { const char* name; int mode, K, KC, rows, box_rows, boxes, stages, distinct, producers; };
  const Cfg cfgs[] = {
      {"SW128 K=1024 box 256x64 (32KB) x1, 4 stages, 1 producer", 0, 1024, 64, 1024, 256, 1, 4, 0, 1},
      {"SW128 K=1024 box 256x64 (32KB) x1, 4 stages, 2 producers", 0, 1024, 64, 1024, 256, 1, 4, 0, 2},
      {"SW128 K=1024 box 256x64 (32KB) x1, 4 stages, 4 producers", 0, 1024, 64, 1024, 256, 1, 4, 0, 4},
      {"SW128 K=1024 box 256x64 (32KB) x1, 6 stages, 3 producers", 0, 1024, 64, 1024, 256, 1, 6, 0, 3},
      {"SW128 K=256  box 128x64 (16KB) x1, 6 stages, 1 producer", 0, 256, 64, 1024, 128, 1, 6, 0, 1},
      {"SW128 K=256  box 128x64 (16KB) x1, 6 stages, 2 producers", 0, 256, 64, 1024, 128, 1, 6, 0, 2},
      {"SW128 K=256  box 128x64 (16KB) x1, 6 stages, 3 producers", 0, 256, 64, 1024, 128, 1, 6, 0, 3},
      {"SW128 K=256  box 128x64 (16KB) x1, 8 stages, 4 producers", 0, 256, 64, 1024, 128, 1, 8, 0, 4},
      {"SW128 K=256  box 128x64 (16KB) x2, 4 stages, 2 producers", 0, 256, 64, 1024, 128, 2, 4, 0, 2},
      {"SW64  K=1024 box 256x32 (16KB) x2, 4 stages, 2 producers", 0, 1024, 32, 1024, 256, 2, 4, 0, 2},
      {"SW64  K=1024 box 256x32 (16KB) x2, 4 stages, 4 producers", 0, 1024, 32, 1024, 256, 2, 4, 0, 4},
      {"bulk copy 32KB contiguous x1, 4 stages, 2 producers", 1, 64, 64, 16384, 256, 1, 4, 0, 2},
      {"bulk copy 32KB contiguous x1, 4 stages, 4 producers", 1, 64, 64, 16384, 256, 1, 4, 0, 4},
  };

  for (const Cfg& c : cfgs) {
    CUtensorMap tm;
    const int chunks = c.K / c.KC;
    cuuint64_t dims[3] = {(cuuint64_t)c.KC, (cuuint64_t)chunks, (cuuint64_t)c.rows};
    cuuint64_t str[2] = {(cuuint64_t)c.KC * 2, (cuuint64_t)c.K * 2};
    cuuint32_t box[3] = {(cuuint32_t)c.KC, 1, (cuuint32_t)c.box_rows};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     c.KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
    const int box_bytes = c.box_rows * c.KC * 2;
    const int units = 2048 / c.boxes;
    const int smem = c.stages * c.boxes * box_bytes + 1024;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(a);
      stream_kernel<<<sms, 160, smem>>>(tm, buf, c.mode, c.stages, c.boxes, box_bytes, c.box_rows, c.rows / c.box_rows, chunks,
                                       units, c.distinct, c.producers, cyc);
      cudaEventRecord(b);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (ms < best) best = ms;
    }
    long long hc; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
    const double bytes_per_cta = (double)units * c.boxes * box_bytes;
    printf("%-66s %7.3f ms  %6.2f TB/s chip  %5.1f B/clk/SM (CTA0 clock64)  %5.0f cycles per 32 KB\n", c.name, best,
           bytes_per_cta * sms / (best * 1e-3) / 1e12, bytes_per_cta / (double)hc, (double)hc / (bytes_per_cta / 32768.0));
  }
  return 0;
}
