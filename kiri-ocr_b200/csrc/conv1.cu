// K2: first stem layer on CUDA cores — Conv(1->48, 3x3, s1, p1, no bias) + BN(eval) + SiLU.
//
// Replaces  ConvStem.net[0:3]   kiri_ocr/model.py:215-217 (cuDNN conv + ATen BN + SiLU).
// K = 9 is not tensor-core work.  Input is the uint8 plane written by K1; the reference's
// normalisation ((v/255 - 0.5)/0.5, model.py:337-338) is applied here in fp32 from a 256-entry
// table, so the model sees exactly the reference's fp32 pixel values (not a bf16 rounding of
// them).  BN is folded into the 48x9 weights + bias on the host; they travel as a kernel
// parameter, i.e. in the constant bank.
// Output is DENSE NHWC bf16 with 48 channels (96 B per pixel).  conv2's TMA tensor map declares an inner extent
// of 48 with a 64-element box, so the 16 channels that pad K to a whole 128-byte swizzle row are zero-filled by
// the TMA unit in shared memory and never exist in HBM (round 1 stored 64 channels: 25 % of this layer's
// writes and of conv2's reads were zeros).
//
// Packed fp32 math: Blackwell's FFMA2 does two IEEE fp32 FMAs per instruction, but its weight operand must sit
// in registers instead of the constant bank.  A thread therefore computes TWO vertically adjacent pixels for a
// PAIR of output channels at a time: one weight-pair load feeds two FFMA2s, the 432 FMAs per pixel become
// 216 FFMA2 + 108 loads, and the SiLU runs on pairs too (mul2, 2 x tanh, fma2, mul2).  Same fp32 operations
// in the same order as the scalar form (round 1 measured the scalar kernel at 0.186 ms and a warp-level
// tensor-core form on exact bf16 operands at +11 %; both were removed from the library in round 2).
#include "common.cuh"
#include "kiri_b200.h"

#include <cstring>

namespace kiri {

static constexpr int kC1 = 48;
static constexpr int kC1Chunks = kC1 * 2 / 16;        // 16-byte chunks per pixel
static constexpr int kConv1Threads = 128;

// Width groups of one batch: group i owns tiles [tile_begin[i], tile_begin[i+1]) (one tile = 2 rows x 128 pixels).
struct Conv1Groups {
  int n;
  int tile_begin[9];
  const uint8_t* planes[8];
  __nv_bfloat16* out[8];
  int W[8];
};

struct Conv1PairParams {
  float2 w[(kC1 / 2) * 9];     // [channel pair][tap] = {w[2c][k], w[2c+1][k]}, BN folded
  float2 b[kC1 / 2];
};

// staging slot of 16-byte chunk j of pixel px: a rotation by one for every other group of four pixels makes the
// 8 lanes of a 128-bit store phase (pixel pitch 96 B = 6 chunks) hit 8 different 16-byte bank groups
__device__ __forceinline__ int c1_slot(int px, int j) {
  int r = j + ((px >> 2) & 1);
  if (r >= kC1Chunks) r -= kC1Chunks;
  return px * kC1Chunks + r;
}

// tile = 2 image rows x 128 columns; thread = one column, both rows
__global__ void __launch_bounds__(kConv1Threads)
conv1_pair_kernel(const __grid_constant__ Conv1Groups G, int H, const __grid_constant__ Conv1PairParams p) {
  __shared__ float s_norm[256];
  __shared__ __align__(16) uint8_t s_out[2 * kConv1Threads * kC1 * 2];   // 2 x 12 KiB staging tiles
  const int tid = threadIdx.x;
  for (int v = tid; v < 256; v += kConv1Threads)
    s_norm[v] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v), 255.0f), 0.5f), 0.5f);
  __syncthreads();
  pdl_trigger();
  pdl_wait();                                       // the planes come from the previous kernel

  int gi = 0;
#pragma unroll
  for (int i = 1; i < 8; ++i)
    if (i < G.n && static_cast<int>(blockIdx.x) >= G.tile_begin[i]) gi = i;
  const int W = G.W[gi];
  const uint8_t* __restrict__ planes = G.planes[gi];
  __nv_bfloat16* __restrict__ out = G.out[gi];
  const int tiles_per_row = W / kConv1Threads;
  const int tile = static_cast<int>(blockIdx.x) - G.tile_begin[gi];
  const int xt = tile % tiles_per_row;
  const int bp = tile / tiles_per_row;            // b * (H/2) + y/2
  const int y = (bp % (H / 2)) * 2;
  const int b = bp / (H / 2);
  const int x = xt * kConv1Threads + tid;
  const uint8_t* img = planes + static_cast<size_t>(b) * H * W;

  float in[4][3];                                  // rows y-1 .. y+2, columns x-1 .. x+1
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int yy = y + r - 1;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = x + kx - 1;
      const bool ok = (yy >= 0) && (yy < H) && (xx >= 0) && (xx < W);
      in[r][kx] = ok ? s_norm[img[yy * W + xx]] : 0.0f;            // conv zero padding
    }
  }
  uint4* so = reinterpret_cast<uint4*>(s_out);
#pragma unroll
  for (int j = 0; j < kC1Chunks; ++j) {            // 8 channels (four pairs) per 16-byte chunk
    uint32_t pk0[4], pk1[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int cp = j * 4 + q;
      float2 a0 = p.b[cp], a1 = a0;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float2 wv = p.w[cp * 9 + ky * 3 + kx];
          a0 = ffma2(make_float2(in[ky][kx], in[ky][kx]), wv, a0);
          a1 = ffma2(make_float2(in[ky + 1][kx], in[ky + 1][kx]), wv, a1);
        }
      const float2 s0 = silu_fast2(a0), s1 = silu_fast2(a1);
      pk0[q] = pack_bf16x2(s0.x, s0.y);
      pk1[q] = pack_bf16x2(s1.x, s1.y);
    }
    so[c1_slot(tid, j)] = make_uint4(pk0[0], pk0[1], pk0[2], pk0[3]);
    so[kConv1Threads * kC1Chunks + c1_slot(tid, j)] = make_uint4(pk1[0], pk1[1], pk1[2], pk1[3]);
  }
  __syncthreads();
  // the tile's two rows are 2 x 12 KiB of contiguous global memory: coalesced 16-byte stores
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    uint4* dst = reinterpret_cast<uint4*>(out + ((static_cast<size_t>(b) * H + y + r) * W + xt * kConv1Threads) * kC1);
#pragma unroll
    for (int i = 0; i < kC1Chunks; ++i) {
      const int q = i * kConv1Threads + tid;
      const int px = q / kC1Chunks, j = q - px * kC1Chunks;
      dst[q] = so[r * kConv1Threads * kC1Chunks + c1_slot(px, j)];
    }
  }
}

}  // namespace kiri

using namespace kiri;

extern "C" int kiri_conv1_multi(const uint8_t* const* planes_u8, void* const* out_bf16_nhwc48, const int* group_lines,
                                const int* group_W, int n_groups, const float* w_host, const float* b_host, int H,
                                cudaStream_t stream) {
  KIRI_REQUIRE(planes_u8 && out_bf16_nhwc48 && group_lines && group_W && w_host && b_host, "kiri_conv1: null pointer");
  KIRI_REQUIRE(n_groups >= 0 && n_groups <= 8, "kiri_conv1_multi: at most 8 groups");
  KIRI_REQUIRE(H > 0 && H % 2 == 0, "kiri_conv1: plane height %d must be even", H);
  Conv1Groups G;
  memset(&G, 0, sizeof(G));
  long long tiles = 0;
  for (int g = 0; g < n_groups; ++g) {
    if (group_lines[g] <= 0) continue;
    KIRI_REQUIRE(planes_u8[g] && out_bf16_nhwc48[g], "kiri_conv1_multi: null pointer in group %d", g);
    KIRI_REQUIRE(group_W[g] % kConv1Threads == 0, "kiri_conv1: width %d must be a multiple of %d", group_W[g], kConv1Threads);
    G.tile_begin[G.n] = static_cast<int>(tiles);
    G.planes[G.n] = planes_u8[g];
    G.out[G.n] = reinterpret_cast<__nv_bfloat16*>(out_bf16_nhwc48[g]);
    G.W[G.n] = group_W[g];
    tiles += static_cast<long long>(group_lines[g]) * (H / 2) * (group_W[g] / kConv1Threads);
    KIRI_REQUIRE(tiles < 0x7fffffffll, "kiri_conv1: grid too large");
    ++G.n;
  }
  for (int i = G.n; i < 9; ++i) G.tile_begin[i] = static_cast<int>(tiles);
  if (tiles == 0) return 0;
  Conv1PairParams pp;
  for (int c = 0; c < kC1 / 2; ++c) {
    for (int k = 0; k < 9; ++k) pp.w[c * 9 + k] = make_float2(w_host[(2 * c) * 9 + k], w_host[(2 * c + 1) * 9 + k]);
    pp.b[c] = make_float2(b_host[2 * c], b_host[2 * c + 1]);
  }
  KIRI_CHECK_CUDA(launch_pdl(conv1_pair_kernel, dim3(static_cast<unsigned>(tiles)), dim3(kConv1Threads), 0, stream, G, H, pp));
  return 0;
}

extern "C" int kiri_conv1(const uint8_t* planes_u8, const float* w_host, const float* b_host, int n_lines,
                          int H, int W, void* out_bf16_nhwc48, cudaStream_t stream) {
  KIRI_REQUIRE(planes_u8 && w_host && b_host && out_bf16_nhwc48, "kiri_conv1: null pointer");
  return kiri_conv1_multi(&planes_u8, &out_bf16_nhwc48, &n_lines, &W, 1, w_host, b_host, H, stream);
}
