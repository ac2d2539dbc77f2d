// LayerNorm helpers shared by norm.cu and decoder.cu: one warp per 256-wide row, 8 channels per lane.
#pragma once
#include "common.cuh"

namespace kiri {

static constexpr int kD = 256;
static constexpr float kLnEps = 1e-5f;

__device__ __forceinline__ void ln8(float (&v)[8], const float* __restrict__ g,
                                    const float* __restrict__ b, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / kD);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q += d * d; }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / kD) + kLnEps);
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(g) + lane * 2);
  const float4 g1 = __ldg(reinterpret_cast<const float4*>(g) + lane * 2 + 1);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(b) + lane * 2);
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(b) + lane * 2 + 1);
  const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * gg[i] + bb[i];
}

__device__ __forceinline__ void st_f32x8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st_bf16x8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 pk;
  pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]);
  pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = pk;
}

}  // namespace kiri
