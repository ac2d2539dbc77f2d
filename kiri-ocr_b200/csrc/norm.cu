// K6/K7 + every LayerNorm of the encoder/decoder: one warp per token, D = 256 (8 channels per
// lane, 16-byte accesses), fp32 statistics, eps 1e-5.
//
//   pool_pos_ln : mean over the RH stem rows + constant 2-D positional table + enc_ln_in
//                 (+ optionally the first layer's norm1)      model.py:194-208, 302-304
//   ln_chain    : LN (-> fp32 and/or bf16) [-> second LN -> bf16]
//                 used for norm1/norm2 of each layer, and for enc_ln followed by ctc_head.0
//                 (model.py:306, 264-268).
#include "common.cuh"
#include "kiri_b200.h"
#include "ln_utils.cuh"

#include <cstring>

namespace kiri {

static constexpr int kLnThreads = 256;    // 8 tokens per CTA

// Width groups of one token stream: group i owns tokens [tok_begin[i], tok_begin[i+1]) and its own stem output.
struct PoolGroups {
  int n;
  int tok_begin[9];
  const __nv_bfloat16* act[8];
  int T[8];
};

// act: bf16 [B, RH, T, 256] per group (NHWC stem output); pos: fp32 [T, 256]
__global__ void __launch_bounds__(kLnThreads)
pool_pos_ln_kernel(const __grid_constant__ PoolGroups G, const float* __restrict__ pos, int n_tok,
                   int RH, const float* g0, const float* b0, const float* g1, const float* b1,
                   float* __restrict__ x_f32, __nv_bfloat16* __restrict__ a_bf16) {
  const int lane = threadIdx.x & 31;
  const int tok = blockIdx.x * (kLnThreads / 32) + (threadIdx.x >> 5);
  pdl_trigger();
  pdl_wait();                                       // the stem output comes from the previous kernel
  if (tok >= n_tok) return;
  int gi = 0;
#pragma unroll
  for (int i = 1; i < 8; ++i)
    if (i < G.n && tok >= G.tok_begin[i]) gi = i;
  const int T = G.T[gi];
  const __nv_bfloat16* __restrict__ act = G.act[gi];
  const int ltok = tok - G.tok_begin[gi];
  const int b = ltok / T, t = ltok - b * T;
  float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const uint4* row = reinterpret_cast<const uint4*>(act + ((static_cast<size_t>(b) * RH) * T + t) * kD + lane * 8);
  const size_t row_pitch = static_cast<size_t>(T) * kD / 8;          // in uint4
  auto add = [&](const uint4 pk) {
    v[0] += bf16_lo(pk.x); v[1] += bf16_hi(pk.x); v[2] += bf16_lo(pk.y); v[3] += bf16_hi(pk.y);
    v[4] += bf16_lo(pk.z); v[5] += bf16_hi(pk.z); v[6] += bf16_lo(pk.w); v[7] += bf16_hi(pk.w);
  };
  if (RH == 6) {                                    // line height 48: all six row loads in flight at once (same sum order)
    uint4 pk[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) pk[r] = __ldg(row + r * row_pitch);
#pragma unroll
    for (int r = 0; r < 6; ++r) add(pk[r]);
  } else {
    for (int r = 0; r < RH; ++r) add(__ldg(row + r * row_pitch));
  }
  const float inv = 1.0f / static_cast<float>(RH);
  const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos + static_cast<size_t>(t) * kD) + lane * 2);
  const float4 p1 = __ldg(reinterpret_cast<const float4*>(pos + static_cast<size_t>(t) * kD) + lane * 2 + 1);
  const float pp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = v[i] * inv + pp[i];
  ln8(v, g0, b0, lane);
  st_f32x8(x_f32 + static_cast<size_t>(tok) * kD + lane * 8, v);
  if (a_bf16) {
    ln8(v, g1, b1, lane);
    st_bf16x8(a_bf16 + static_cast<size_t>(tok) * kD + lane * 8, v);
  }
}

__global__ void __launch_bounds__(kLnThreads)
ln_chain_kernel(const float* __restrict__ x, int n_tok, const float* g0, const float* b0,
                float* __restrict__ y_f32, __nv_bfloat16* __restrict__ y_bf16, const float* g1,
                const float* b1, __nv_bfloat16* __restrict__ z_bf16) {
  const int lane = threadIdx.x & 31;
  const int tok = blockIdx.x * (kLnThreads / 32) + (threadIdx.x >> 5);
  pdl_trigger();
  pdl_wait();
  if (tok >= n_tok) return;
  const float4 x0 = *reinterpret_cast<const float4*>(x + static_cast<size_t>(tok) * kD + lane * 8);
  const float4 x1 = *reinterpret_cast<const float4*>(x + static_cast<size_t>(tok) * kD + lane * 8 + 4);
  float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
  ln8(v, g0, b0, lane);
  if (y_f32) st_f32x8(y_f32 + static_cast<size_t>(tok) * kD + lane * 8, v);
  if (y_bf16) st_bf16x8(y_bf16 + static_cast<size_t>(tok) * kD + lane * 8, v);
  if (z_bf16) {
    ln8(v, g1, b1, lane);
    st_bf16x8(z_bf16 + static_cast<size_t>(tok) * kD + lane * 8, v);
  }
}

}  // namespace kiri

using namespace kiri;

extern "C" int kiri_pool_pos_ln_multi(const void* const* act_bf16, const int* group_lines, const int* group_T, int n_groups,
                                      const float* pos_table, int RH, int D, const float* g0, const float* b0, const float* g1,
                                      const float* b1, float* x_f32, void* a_bf16, cudaStream_t stream) {
  KIRI_REQUIRE(D == kD, "kiri_pool_pos_ln: model width %d unsupported (kernels are built for 256)", D);
  KIRI_REQUIRE(act_bf16 && group_lines && group_T && pos_table && g0 && b0 && x_f32, "kiri_pool_pos_ln: null pointer");
  KIRI_REQUIRE(!a_bf16 || (g1 && b1), "kiri_pool_pos_ln: second LayerNorm needs its affine");
  KIRI_REQUIRE(n_groups >= 0 && n_groups <= 8, "kiri_pool_pos_ln_multi: at most 8 groups");
  PoolGroups G;
  memset(&G, 0, sizeof(G));
  long long tok = 0;
  for (int g = 0; g < n_groups; ++g) {
    if (group_lines[g] <= 0) continue;
    KIRI_REQUIRE(act_bf16[g] && group_T[g] > 0, "kiri_pool_pos_ln_multi: bad group %d", g);
    G.tok_begin[G.n] = static_cast<int>(tok);
    G.act[G.n] = reinterpret_cast<const __nv_bfloat16*>(act_bf16[g]);
    G.T[G.n] = group_T[g];
    tok += static_cast<long long>(group_lines[g]) * group_T[g];
    KIRI_REQUIRE(tok < 0x7fffffffll, "kiri_pool_pos_ln_multi: too many tokens");
    ++G.n;
  }
  for (int i = G.n; i < 9; ++i) G.tok_begin[i] = static_cast<int>(tok);
  const int n_tok = static_cast<int>(tok);
  if (n_tok == 0) return 0;
  const int per = kLnThreads / 32;
  KIRI_CHECK_CUDA(launch_pdl(pool_pos_ln_kernel, dim3((n_tok + per - 1) / per), dim3(kLnThreads), 0, stream, G, pos_table, n_tok,
                             RH, g0, b0, g1, b1, x_f32, reinterpret_cast<__nv_bfloat16*>(a_bf16)));
  return 0;
}

extern "C" int kiri_pool_pos_ln(const void* act_bf16, const float* pos_table, int n_lines, int RH, int T,
                                int D, const float* g0, const float* b0, const float* g1, const float* b1,
                                float* x_f32, void* a_bf16, cudaStream_t stream) {
  return kiri_pool_pos_ln_multi(&act_bf16, &n_lines, &T, 1, pos_table, RH, D, g0, b0, g1, b1, x_f32, a_bf16, stream);
}

extern "C" int kiri_layernorm(const float* x, int n_tok, int D, const float* g0, const float* b0, float* y_f32,
                              void* y_bf16, const float* g1, const float* b1, void* z_bf16,
                              cudaStream_t stream) {
  KIRI_REQUIRE(D == kD, "kiri_layernorm: model width %d unsupported (kernels are built for 256)", D);
  KIRI_REQUIRE(x && g0 && b0, "kiri_layernorm: null pointer");
  KIRI_REQUIRE(!z_bf16 || (g1 && b1), "kiri_layernorm: second LayerNorm needs its affine");
  if (n_tok == 0) return 0;
  const int per = kLnThreads / 32;
  KIRI_CHECK_CUDA(launch_pdl(ln_chain_kernel, dim3((n_tok + per - 1) / per), dim3(kLnThreads), 0, stream,
                             x, n_tok, g0, b0, y_f32, reinterpret_cast<__nv_bfloat16*>(y_bf16), g1, b1,
                             reinterpret_cast<__nv_bfloat16*>(z_bf16)));
  return 0;
}
