// K2: first stem layer on CUDA cores — Conv(1->48, 3x3, s1, p1, no bias) + BN(eval) + SiLU.
//
// Replaces  ConvStem.net[0:3]   kiri_ocr/model.py:215-217 (cuDNN conv + ATen BN + SiLU).
// K = 9 is not tensor-core work.  Input is the uint8 plane written by K1; the reference's
// normalisation ((v/255 - 0.5)/0.5, model.py:337-338) is applied here in fp32 from a 256-entry
// table, so the model sees exactly the reference's fp32 pixel values (not a bf16 rounding of
// them).  BN is folded into the 48x9 weights + bias on the host; they travel as a kernel
// parameter, i.e. in the constant bank, so every FFMA reads its weight operand for free.
// Output is NHWC bf16 with the 48 channels padded to 64 (zeros) = two 32-channel TMA chunks.
#include "common.cuh"
#include "kiri_b200.h"

namespace kiri {

static constexpr int kC1 = 48;
static constexpr int kC1Pad = 64;
static constexpr int kConv1Threads = 128;

struct Conv1Params {
  float w[kC1 * 9];   // [cout][ky*3+kx], BN folded
  float b[kC1];
};

__global__ void __launch_bounds__(kConv1Threads)
conv1_bn_silu_kernel(const uint8_t* __restrict__ planes, __nv_bfloat16* __restrict__ out, int H,
                     int W, const __grid_constant__ Conv1Params p) {
  __shared__ float s_norm[256];
  __shared__ __align__(16) uint8_t s_out[kConv1Threads * kC1Pad * 2];   // 16 KiB staging tile
  const int tid = threadIdx.x;
  for (int v = tid; v < 256; v += kConv1Threads)
    s_norm[v] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v), 255.0f), 0.5f), 0.5f);
  __syncthreads();
  pdl_trigger();
  pdl_wait();                                       // the planes come from the previous kernel

  const int tiles_per_row = W / kConv1Threads;
  const int tile = blockIdx.x;
  const int xt = tile % tiles_per_row;
  const int by = tile / tiles_per_row;            // b * H + y
  const int y = by % H;
  const int b = by / H;
  const int x = xt * kConv1Threads + tid;
  const uint8_t* img = planes + static_cast<size_t>(b) * H * W;

  float in[9];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = y + ky - 1;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = x + kx - 1;
      const bool ok = (yy >= 0) && (yy < H) && (xx >= 0) && (xx < W);
      in[ky * 3 + kx] = ok ? s_norm[img[yy * W + xx]] : 0.0f;      // conv zero padding
    }
  }
  uint32_t packed[kC1Pad / 2];
#pragma unroll
  for (int c = 0; c < kC1; c += 2) {
    float a0 = p.b[c], a1 = p.b[c + 1];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      a0 = fmaf(in[k], p.w[c * 9 + k], a0);
      a1 = fmaf(in[k], p.w[(c + 1) * 9 + k], a1);
    }
    packed[c / 2] = pack_bf16x2(silu_fast(a0), silu_fast(a1));
  }
#pragma unroll
  for (int c = kC1 / 2; c < kC1Pad / 2; ++c) packed[c] = 0u;

  // stage [pixel][128 B] with a 16-byte-chunk XOR swizzle, then write the tile out coalesced
  uint4* so = reinterpret_cast<uint4*>(s_out);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    so[tid * 8 + (j ^ (tid & 7))] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(by) * W + xt * kConv1Threads) * kC1Pad);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int q = i * kConv1Threads + tid;
    const int px = q >> 3, j = q & 7;
    dst[q] = so[px * 8 + (j ^ (px & 7))];
  }
}

}  // namespace kiri

using namespace kiri;

extern "C" int kiri_conv1(const uint8_t* planes_u8, const float* w_host, const float* b_host, int n_lines,
                          int H, int W, void* out_bf16_nhwc64, cudaStream_t stream) {
  KIRI_REQUIRE(planes_u8 && w_host && b_host && out_bf16_nhwc64, "kiri_conv1: null pointer");
  KIRI_REQUIRE(W % kConv1Threads == 0, "kiri_conv1: width %d must be a multiple of %d", W, kConv1Threads);
  if (n_lines == 0) return 0;
  Conv1Params p;
  for (int i = 0; i < kC1 * 9; ++i) p.w[i] = w_host[i];
  for (int i = 0; i < kC1; ++i) p.b[i] = b_host[i];
  const long long tiles = static_cast<long long>(n_lines) * H * (W / kConv1Threads);
  KIRI_REQUIRE(tiles < 0x7fffffffll, "kiri_conv1: grid too large");
  KIRI_CHECK_CUDA(launch_pdl(conv1_bn_silu_kernel, dim3(static_cast<unsigned>(tiles)), dim3(kConv1Threads), 0, stream,
                             planes_u8, reinterpret_cast<__nv_bfloat16*>(out_bf16_nhwc64), H, W, p));
  return 0;
}
