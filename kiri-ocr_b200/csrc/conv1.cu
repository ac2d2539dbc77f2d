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

#include <cstring>

namespace kiri {

static constexpr int kC1 = 48;
static constexpr int kC1Pad = 64;
static constexpr int kConv1Threads = 128;

struct Conv1Params {
  float w[kC1 * 9];   // [cout][ky*3+kx], BN folded
  float b[kC1];
};

// Width groups of one batch: group i owns tiles [tile_begin[i], tile_begin[i+1]) (one tile = 128 pixels of a row).
struct Conv1Groups {
  int n;
  int tile_begin[9];
  const uint8_t* planes[8];
  __nv_bfloat16* out[8];
  int W[8];
};

__global__ void __launch_bounds__(kConv1Threads)
conv1_bn_silu_kernel(const __grid_constant__ Conv1Groups G, int H, const __grid_constant__ Conv1Params p) {
  __shared__ float s_norm[256];
  __shared__ __align__(16) uint8_t s_out[kConv1Threads * kC1Pad * 2];   // 16 KiB staging tile
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


// ------------------------------------------------------------------------------------------------
// Packed-fp32 form (default): Blackwell's FFMA2 does two IEEE fp32 FMAs per instruction, but its weight
// operand must sit in (uniform) registers instead of the constant bank.  A thread therefore computes TWO
// vertically adjacent pixels for a PAIR of output channels at a time: one weight-pair load feeds two
// FFMA2s, the 432 FMAs per pixel become 216 FFMA2 + 108 loads, and the SiLU runs on pairs too
// (mul2, 2 x tanh, fma2, mul2).  Results are bit-identical to the scalar kernel (same fp32 operations).
struct Conv1PairParams {
  float2 w[(kC1 / 2) * 9];     // [channel pair][tap] = {w[2c][k], w[2c+1][k]}, BN folded
  float2 b[kC1 / 2];
};

// tile = 2 image rows x 128 columns; thread = one column, both rows
__global__ void __launch_bounds__(kConv1Threads)
conv1_pair_kernel(const __grid_constant__ Conv1Groups G, int H, const __grid_constant__ Conv1PairParams p) {
  __shared__ float s_norm[256];
  __shared__ __align__(16) uint8_t s_out[2 * kConv1Threads * kC1Pad * 2];   // 2 x 16 KiB staging tiles
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
  for (int j = 0; j < 6; ++j) {                    // 8 channels (four pairs) per 16-byte chunk
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
    so[tid * 8 + (j ^ (tid & 7))] = make_uint4(pk0[0], pk0[1], pk0[2], pk0[3]);
    so[kConv1Threads * 8 + tid * 8 + (j ^ (tid & 7))] = make_uint4(pk1[0], pk1[1], pk1[2], pk1[3]);
  }
#pragma unroll
  for (int j = 6; j < 8; ++j) {                    // the 16 zero channels
    so[tid * 8 + (j ^ (tid & 7))] = make_uint4(0u, 0u, 0u, 0u);
    so[kConv1Threads * 8 + tid * 8 + (j ^ (tid & 7))] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    uint4* dst = reinterpret_cast<uint4*>(out + ((static_cast<size_t>(b) * H + y + r) * W + xt * kConv1Threads) * kC1Pad);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int q = i * kConv1Threads + tid;
      const int px = q >> 3, j = q & 7;
      dst[q] = so[r * kConv1Threads * 8 + px * 8 + (j ^ (px & 7))];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core form (opt-in, KIRI_CONV1_TC=1; measured 11 % SLOWER than the FFMA kernel above, see kiri_conv1).
// The 9-tap products run on warp-level mma.sync m16n8k16; what remains is the SiLU epilogue and the
// 128 B/pixel store, which also bound the FFMA form.  Exactness is kept without fp32
// operands: with u = v - 128 for in-image taps and u = -0.5 for padded taps (both exact in bf16),
//     (v/255 - 0.5)/0.5 = (u + 0.5)/127.5      and the conv padding value 0 <-> u = -0.5,
// so  out[c] = b[c] + 0.5*sum_k w'[c][k] + sum_k w'[c][k]*u_k,   w' = w/127.5 = w'_hi + w'_lo (two bf16).
// A = [16 pixels x 16 taps (9 used)], B = [taps x 8 channels]; two k-steps (hi, lo weights) per n-tile.
// Products of two bf16 are exact in the fp32 accumulator, the weight split leaves 2^-17 relative error.
struct Conv1TcParams {
  uint32_t bfrag[6][2][32][2];   // B fragments in mma register order: [n-tile][hi|lo][lane][reg]
  float bias[kC1];               // b[c] + 0.5 * sum_k w'[c][k]
};

__device__ __forceinline__ void mma16816_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                              uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

static constexpr int kPatchPitch = 136;     // 128 + 2 halo columns, padded

__global__ void __launch_bounds__(kConv1Threads, 8)
conv1_tc_kernel(const uint8_t* __restrict__ planes, __nv_bfloat16* __restrict__ out, int H, int W, int total_tiles,
                const __grid_constant__ Conv1TcParams p) {
  __shared__ __align__(16) unsigned short s_patch[3 * kPatchPitch];          // bf16 bits of u
  __shared__ __align__(16) uint8_t s_out[kConv1Threads * kC1Pad * 2];        // 16 KiB staging tile [pixel][128 B]
  __shared__ __align__(8) uint32_t s_bfrag[6 * 2 * 32 * 2];                  // B fragments, lane-contiguous
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  for (int i = tid; i < 6 * 2 * 32 * 2; i += kConv1Threads) s_bfrag[i] = (&p.bfrag[0][0][0][0])[i];
  float bs[6][2];                                   // biases of this thread's channels 8j + 2t, 8j + 2t + 1
#pragma unroll
  for (int j = 0; j < 6; ++j) { bs[j][0] = p.bias[8 * j + 2 * t]; bs[j][1] = p.bias[8 * j + 2 * t + 1]; }
  // the 16 zero channels (chunks 6, 7 of every pixel row) are written once: no other store touches them
  for (int i = tid; i < kConv1Threads * 2; i += kConv1Threads) {
    const int px = i >> 1, j = 6 + (i & 1);
    reinterpret_cast<uint4*>(s_out)[px * 8 + (j ^ (px & 7))] = make_uint4(0u, 0u, 0u, 0u);
  }
  // taps 2t, 2t+1 of the 3x3 window as patch offsets (tap k = ky*3 + kx); tap 8 belongs to t == 0
  const int k0 = 2 * t, k1 = 2 * t + 1;
  const int o0 = (k0 / 3) * kPatchPitch + (k0 % 3), o1 = (k1 / 3) * kPatchPitch + (k1 % 3);
  const int o8 = 2 * kPatchPitch + 2;
  pdl_trigger();
  pdl_wait();                                       // the planes come from the previous kernel
  const int tiles_per_row = W / kConv1Threads;
  const unsigned short kPad = 0xBF00;               // bf16(-0.5): the reference's zero padding
  // the 3 x 130 patch of a tile is fetched into registers one tile ahead (4 pixels per thread)
  unsigned short nxt[4];
  auto fetch = [&](int tile) {
    const int xt = tile % tiles_per_row, by = tile / tiles_per_row;
    const int y = by % H, b = by / H, x0 = xt * kConv1Threads;
    const uint8_t* img = planes + static_cast<size_t>(b) * H * W;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = tid + r * kConv1Threads;
      const int ky = i / 130, c = i - ky * 130;
      const int yy = y + ky - 1, xx = x0 - 1 + c;
      const bool ok = (i < 3 * 130) && (yy >= 0) && (yy < H) && (xx >= 0) && (xx < W);
      nxt[r] = ok ? __bfloat16_as_ushort(__float2bfloat16(static_cast<float>(__ldg(img + yy * W + xx)) - 128.0f)) : kPad;   // exact
    }
  };
  if (static_cast<int>(blockIdx.x) < total_tiles) fetch(blockIdx.x);
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int xt = tile % tiles_per_row;
    const int by = tile / tiles_per_row;            // b * H + y
    const int x0 = xt * kConv1Threads;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = tid + r * kConv1Threads;
      if (i < 3 * 130) { const int ky = i / 130; s_patch[ky * kPatchPitch + (i - ky * 130)] = nxt[r]; }
    }
    __syncthreads();                                // patch complete; the previous tile's copy-out has read s_out
    if (tile + static_cast<int>(gridDim.x) < total_tiles) fetch(tile + gridDim.x);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int pa = warp * 32 + mt * 16 + g, pb = pa + 8;      // this thread's two pixels of the m-tile
      const uint32_t a0 = static_cast<uint32_t>(s_patch[pa + o0]) | (static_cast<uint32_t>(s_patch[pa + o1]) << 16);
      const uint32_t a1 = static_cast<uint32_t>(s_patch[pb + o0]) | (static_cast<uint32_t>(s_patch[pb + o1]) << 16);
      const uint32_t a2 = t == 0 ? static_cast<uint32_t>(s_patch[pa + o8]) : 0u;
      const uint32_t a3 = t == 0 ? static_cast<uint32_t>(s_patch[pb + o8]) : 0u;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const uint2 bh = *reinterpret_cast<const uint2*>(&s_bfrag[((j * 2 + 0) * 32 + lane) * 2]);
        const uint2 bl = *reinterpret_cast<const uint2*>(&s_bfrag[((j * 2 + 1) * 32 + lane) * 2]);
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816_bf16(c, a0, a1, a2, a3, bl.x, bl.y);           // low weight halves first
        mma16816_bf16(c, a0, a1, a2, a3, bh.x, bh.y);
        const uint32_t va = pack_bf16x2(silu_fast(c[0] + bs[j][0]), silu_fast(c[1] + bs[j][1]));
        const uint32_t vb = pack_bf16x2(silu_fast(c[2] + bs[j][0]), silu_fast(c[3] + bs[j][1]));
        *reinterpret_cast<uint32_t*>(s_out + pa * 128 + ((j ^ (pa & 7)) << 4) + t * 4) = va;
        *reinterpret_cast<uint32_t*>(s_out + pb * 128 + ((j ^ (pb & 7)) << 4) + t * 4) = vb;
      }
    }
    __syncthreads();                                // staging tile complete; everyone is done with s_patch
    const uint4* so = reinterpret_cast<const uint4*>(s_out);
    uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(by) * W + x0) * kC1Pad);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int q = i * kConv1Threads + tid;
      const int px = q >> 3, j = q & 7;
      dst[q] = so[px * 8 + (j ^ (px & 7))];
    }
  }
}

static uint16_t host_bf16_rn(float f) {            // round-to-nearest-even fp32 -> bf16 bits
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
static float host_bf16_to_f(uint16_t h) {
  const uint32_t u = static_cast<uint32_t>(h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

}  // namespace kiri

using namespace kiri;

extern "C" int kiri_conv1_multi(const uint8_t* const* planes_u8, void* const* out_bf16_nhwc64, const int* group_lines,
                                const int* group_W, int n_groups, const float* w_host, const float* b_host, int H,
                                cudaStream_t stream) {
  KIRI_REQUIRE(planes_u8 && out_bf16_nhwc64 && group_lines && group_W && w_host && b_host, "kiri_conv1: null pointer");
  KIRI_REQUIRE(n_groups >= 0 && n_groups <= 8, "kiri_conv1_multi: at most 8 groups");
  Conv1Groups G;
  memset(&G, 0, sizeof(G));
  long long tiles = 0;
  for (int g = 0; g < n_groups; ++g) {
    if (group_lines[g] <= 0) continue;
    KIRI_REQUIRE(planes_u8[g] && out_bf16_nhwc64[g], "kiri_conv1_multi: null pointer in group %d", g);
    KIRI_REQUIRE(group_W[g] % kConv1Threads == 0, "kiri_conv1: width %d must be a multiple of %d", group_W[g], kConv1Threads);
    G.tile_begin[G.n] = static_cast<int>(tiles);
    G.planes[G.n] = planes_u8[g];
    G.out[G.n] = reinterpret_cast<__nv_bfloat16*>(out_bf16_nhwc64[g]);
    G.W[G.n] = group_W[g];
    tiles += static_cast<long long>(group_lines[g]) * H * (group_W[g] / kConv1Threads);
    KIRI_REQUIRE(tiles < 0x7fffffffll, "kiri_conv1: grid too large");
    ++G.n;
  }
  for (int i = G.n; i < 9; ++i) G.tile_begin[i] = static_cast<int>(tiles);
  if (tiles == 0) return 0;
  static const bool scalar_form = getenv("KIRI_CONV1_SCALAR") != nullptr;
  if (!scalar_form && H % 2 == 0) {
    // tiles of the pair kernel cover two rows: half as many, same per-group order
    Conv1PairParams pp;
    for (int c = 0; c < kC1 / 2; ++c) {
      for (int k = 0; k < 9; ++k) pp.w[c * 9 + k] = make_float2(w_host[(2 * c) * 9 + k], w_host[(2 * c + 1) * 9 + k]);
      pp.b[c] = make_float2(b_host[2 * c], b_host[2 * c + 1]);
    }
    for (int i = 0; i <= 8; ++i) G.tile_begin[i] /= 2;
    KIRI_CHECK_CUDA(launch_pdl(conv1_pair_kernel, dim3(static_cast<unsigned>(tiles / 2)), dim3(kConv1Threads), 0, stream, G, H, pp));
    return 0;
  }
  Conv1Params p;
  for (int i = 0; i < kC1 * 9; ++i) p.w[i] = w_host[i];
  for (int i = 0; i < kC1; ++i) p.b[i] = b_host[i];
  KIRI_CHECK_CUDA(launch_pdl(conv1_bn_silu_kernel, dim3(static_cast<unsigned>(tiles)), dim3(kConv1Threads), 0, stream, G, H, p));
  return 0;
}

extern "C" int kiri_conv1_ffma(const uint8_t* planes_u8, const float* w_host, const float* b_host, int n_lines,
                               int H, int W, void* out_bf16_nhwc64, cudaStream_t stream) {
  KIRI_REQUIRE(planes_u8 && w_host && b_host && out_bf16_nhwc64, "kiri_conv1: null pointer");
  return kiri_conv1_multi(&planes_u8, &out_bf16_nhwc64, &n_lines, &W, 1, w_host, b_host, H, stream);
}

extern "C" int kiri_conv1_tc(const uint8_t* planes_u8, const float* w_host, const float* b_host, int n_lines,
                             int H, int W, void* out_bf16_nhwc64, cudaStream_t stream);

extern "C" int kiri_conv1(const uint8_t* planes_u8, const float* w_host, const float* b_host, int n_lines,
                          int H, int W, void* out_bf16_nhwc64, cudaStream_t stream) {
  // Measured on the B200 (256 bucketed lines): FFMA form 0.211 ms, tensor-core form 0.235 ms — the layer is bound
  // by the 48 SiLUs + the 128-byte store per pixel, not by the 432 FMAs, and the FFMA kernel keeps more warps
  // resident (56 vs 63 registers, no per-tile fragment traffic).  The tensor-core form stays as an opt-in.
  static const bool use_tc = getenv("KIRI_CONV1_TC") != nullptr;
  if (!use_tc) return kiri_conv1_ffma(planes_u8, w_host, b_host, n_lines, H, W, out_bf16_nhwc64, stream);
  return kiri_conv1_tc(planes_u8, w_host, b_host, n_lines, H, W, out_bf16_nhwc64, stream);
}

extern "C" int kiri_conv1_tc(const uint8_t* planes_u8, const float* w_host, const float* b_host, int n_lines,
                             int H, int W, void* out_bf16_nhwc64, cudaStream_t stream) {
  KIRI_REQUIRE(planes_u8 && w_host && b_host && out_bf16_nhwc64, "kiri_conv1: null pointer");
  KIRI_REQUIRE(W % kConv1Threads == 0, "kiri_conv1: width %d must be a multiple of %d", W, kConv1Threads);
  if (n_lines == 0) return 0;
  // fragment table: a few thousand flops on the host, cached for the (single) weight set of a process
  static Conv1TcParams cached;
  static const float* cached_w = nullptr;
  static float cached_w0 = 0.f, cached_b0 = 0.f;
  if (cached_w != w_host || cached_w0 != w_host[0] || cached_b0 != b_host[0]) {
    uint16_t hi[kC1][16], lo[kC1][16];
    for (int c = 0; c < kC1; ++c) {
      double half_sum = 0.0;
      for (int k = 0; k < 16; ++k) {
        const float wp = k < 9 ? w_host[c * 9 + k] / 127.5f : 0.f;
        hi[c][k] = host_bf16_rn(wp);
        lo[c][k] = host_bf16_rn(wp - host_bf16_to_f(hi[c][k]));
        half_sum += 0.5 * (static_cast<double>(host_bf16_to_f(hi[c][k])) + static_cast<double>(host_bf16_to_f(lo[c][k])));
      }
      cached.bias[c] = static_cast<float>(static_cast<double>(b_host[c]) + half_sum);
    }
    for (int j = 0; j < 6; ++j)
      for (int h = 0; h < 2; ++h)
        for (int lane = 0; lane < 32; ++lane) {
          const int g = lane >> 2, t = lane & 3, c = 8 * j + g;
          const uint16_t(*w)[16] = h == 0 ? hi : lo;
          cached.bfrag[j][h][lane][0] = static_cast<uint32_t>(w[c][2 * t]) | (static_cast<uint32_t>(w[c][2 * t + 1]) << 16);
          cached.bfrag[j][h][lane][1] = static_cast<uint32_t>(w[c][2 * t + 8]) | (static_cast<uint32_t>(w[c][2 * t + 9]) << 16);
        }
    cached_w = w_host; cached_w0 = w_host[0]; cached_b0 = b_host[0];
  }
  const long long tiles = static_cast<long long>(n_lines) * H * (W / kConv1Threads);
  KIRI_REQUIRE(tiles < 0x7fffffffll, "kiri_conv1: grid too large");
  static int sms = 0;
  if (sms == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  static bool carveout_set = false;
  if (!carveout_set) {
    cudaFuncSetAttribute(conv1_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    carveout_set = true;
  }
  const long long cap = static_cast<long long>(sms) * 8;           // 8 resident CTAs per SM, each walks its tiles
  const unsigned grid = static_cast<unsigned>(tiles < cap ? tiles : cap);
  KIRI_CHECK_CUDA(launch_pdl(conv1_tc_kernel, dim3(grid), dim3(kConv1Threads), 0, stream, planes_u8,
                             reinterpret_cast<__nv_bfloat16*>(out_bf16_nhwc64), H, W, static_cast<int>(tiles), cached));
  return 0;
}
