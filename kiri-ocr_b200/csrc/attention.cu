// Encoder self-attention (part of K8): softmax(Q K^T / sqrt(32)) V for one (line, head) per CTA.
//
// Replaces the scaled-dot-product inside nn.TransformerEncoderLayer (kiri_ocr/model.py:246-261;
// 8 heads, head_dim 32, no mask of any kind in the reference).  T <= 160 keys fit in shared
// memory at once, so this is a single-tile flash-style kernel: S stays in registers, the
// soft-max is done on the accumulator fragments, and P is re-used in place as the A operand
// of P*V.  It is 3.5 % of the encoder FLOPs with K = 32 inner dimensions, so it runs on
// warp-level mma.sync (m16n8k16, bf16 -> fp32); the GEMMs around it are the tcgen05 kernels.
// An optional per-line key length masks padded columns (bucketed mode (c)).
#include "common.cuh"
#include "kiri_b200.h"

namespace kiri {

static constexpr int kHd = 32;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// shared tile [rows][32 bf16] = 4 x 16-byte chunks per row, chunk index XOR-swizzled by row
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  return static_cast<uint32_t>(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}

// Two CTAs per (line, head), each owning half of the query rows (T/32 warps of 16 rows): with one
// 10-warp CTA the 106 registers per thread allowed a single resident CTA per SM, so the global-load
// phase of one (line, head) never overlapped the MMA phase of another; 5-warp CTAs fit three per SM.
template <int T>
__global__ void __launch_bounds__(T)
encoder_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int D,
                         const int* __restrict__ kv_len) {
  constexpr int NB = T / 8;          // key blocks of 8
  __shared__ __align__(128) uint8_t s_q[T * 64];
  __shared__ __align__(128) uint8_t s_k[T * 64];
  __shared__ __align__(128) uint8_t s_v[T * 64];
  const int head = blockIdx.x, line = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_trigger();
  pdl_wait();                                       // qkv comes from the previous kernel
  const size_t ld = static_cast<size_t>(3) * D;
  const __nv_bfloat16* base = qkv + static_cast<size_t>(line) * T * ld + head * kHd;

  const int q0 = static_cast<int>(blockIdx.z) * (T / 2);       // first query row of this CTA
  for (int idx = tid; idx < 2 * T * 4 + (T / 2) * 4; idx += T) {
    int which, r, c;
    if (idx < 2 * T * 4) { which = 1 + idx / (T * 4); const int rem = idx % (T * 4); r = rem >> 2; c = rem & 3; }
    else { which = 0; const int rem = idx - 2 * T * 4; r = q0 + (rem >> 2); c = rem & 3; }
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + r * ld + which * D + c * 8));
    uint8_t* dst = which == 0 ? s_q : (which == 1 ? s_k : s_v);
    *reinterpret_cast<uint4*>(dst + tile_off(r, c)) = v;
  }
  __syncthreads();

  const int m0 = q0 + warp * 16;
  const uint32_t q_base = smem_u32(s_q), k_base = smem_u32(s_k), v_base = smem_u32(s_v);

  // Q fragments for the two k16 steps of head_dim 32
  uint32_t qa[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
    ldsm_x4(q_base + tile_off(m0 + (lane & 15), ks * 2 + (lane >> 4)), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);

  float s[NB][4];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) { s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f; }
#pragma unroll
  for (int np = 0; np < NB / 2; ++np) {        // 16 keys per iteration
    const int mi = lane >> 3;                  // which 8x8 matrix this lane addresses
    const int krow = np * 16 + (mi >> 1) * 8 + (lane & 7);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(k_base + tile_off(krow, ks * 2 + (mi & 1)), b0, b1, b2, b3);
      mma16816(s[2 * np], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b0, b1);
      mma16816(s[2 * np + 1], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b2, b3);
    }
  }

  // soft-max over keys for rows g = lane/4 (regs 0,1) and g+8 (regs 2,3)
  const float sl2 = 0.17677669529663687f * 1.4426950408889634f;   // 1/sqrt(32) * log2(e)
  const int klen = kv_len ? kv_len[line] : T;
  const int t2 = (lane & 3) * 2;
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    if (kv_len) {
      const int key = nb * 8 + t2;
      if (key >= klen) { s[nb][0] = -INFINITY; s[nb][2] = -INFINITY; }
      if (key + 1 >= klen) { s[nb][1] = -INFINITY; s[nb][3] = -INFINITY; }
    }
    mx0 = fmaxf(mx0, fmaxf(s[nb][0], s[nb][1]));
    mx1 = fmaxf(mx1, fmaxf(s[nb][2], s[nb][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  const float o0 = mx0 * sl2, o1 = mx1 * sl2;
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    s[nb][0] = exp2f(fmaf(s[nb][0], sl2, -o0));
    s[nb][1] = exp2f(fmaf(s[nb][1], sl2, -o0));
    s[nb][2] = exp2f(fmaf(s[nb][2], sl2, -o1));
    s[nb][3] = exp2f(fmaf(s[nb][3], sl2, -o1));
    sum0 += s[nb][0] + s[nb][1];
    sum1 += s[nb][2] + s[nb][3];
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

  // O = P V
  float o[4][4];
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) { o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.f; }
#pragma unroll
  for (int kp = 0; kp < NB / 2; ++kp) {        // 16 keys per iteration
    const uint32_t a0 = pack_bf16x2(s[2 * kp][0], s[2 * kp][1]);
    const uint32_t a1 = pack_bf16x2(s[2 * kp][2], s[2 * kp][3]);
    const uint32_t a2 = pack_bf16x2(s[2 * kp + 1][0], s[2 * kp + 1][1]);
    const uint32_t a3 = pack_bf16x2(s[2 * kp + 1][2], s[2 * kp + 1][3]);
    const int mi = lane >> 3;
    const int vrow = kp * 16 + (mi & 1) * 8 + (lane & 7);
#pragma unroll
    for (int nh = 0; nh < 2; ++nh) {           // head-dim halves of 16
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(v_base + tile_off(vrow, nh * 2 + (mi >> 1)), b0, b1, b2, b3);
      mma16816(o[2 * nh], a0, a1, a2, a3, b0, b1);
      mma16816(o[2 * nh + 1], a0, a1, a2, a3, b2, b3);
    }
  }
  const float r0 = 1.0f / sum0, r1 = 1.0f / sum1;
  const int g = lane >> 2;
  __nv_bfloat16* orow0 = out + (static_cast<size_t>(line) * T + m0 + g) * D + head * kHd + t2;
  __nv_bfloat16* orow1 = orow0 + static_cast<size_t>(8) * D;
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) {
    *reinterpret_cast<uint32_t*>(orow0 + nb * 8) = pack_bf16x2(o[nb][0] * r0, o[nb][1] * r0);
    *reinterpret_cast<uint32_t*>(orow1 + nb * 8) = pack_bf16x2(o[nb][2] * r1, o[nb][3] * r1);
  }
}

}  // namespace kiri

using namespace kiri;

extern "C" int kiri_encoder_attention(const void* qkv_bf16, void* out_bf16, int n_lines, int T, int heads,
                                      int D, const int* kv_len, cudaStream_t stream) {
  KIRI_REQUIRE(qkv_bf16 && out_bf16, "kiri_encoder_attention: null pointer");
  KIRI_REQUIRE(D == heads * kHd, "kiri_encoder_attention: head_dim must be 32 (D=%d, heads=%d)", D, heads);
  if (n_lines == 0) return 0;
  dim3 grid(heads, n_lines, 2);
  static bool carveout_set = false;
  if (!carveout_set) {
    // several CTAs per SM need the large shared-memory carve-out (the driver's default for a
    // 30 KB static kernel was a 32 KB configuration = one resident CTA)
    cudaFuncSetAttribute(encoder_attention_kernel<32>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(encoder_attention_kernel<64>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(encoder_attention_kernel<96>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(encoder_attention_kernel<128>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(encoder_attention_kernel<160>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    carveout_set = true;
  }
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(qkv_bf16);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  switch (T) {
    case 32:  KIRI_CHECK_CUDA(launch_pdl(encoder_attention_kernel<32>, grid, dim3(32), 0, stream, q, o, D, kv_len)); break;
    case 64:  KIRI_CHECK_CUDA(launch_pdl(encoder_attention_kernel<64>, grid, dim3(64), 0, stream, q, o, D, kv_len)); break;
    case 96:  KIRI_CHECK_CUDA(launch_pdl(encoder_attention_kernel<96>, grid, dim3(96), 0, stream, q, o, D, kv_len)); break;
    case 128: KIRI_CHECK_CUDA(launch_pdl(encoder_attention_kernel<128>, grid, dim3(128), 0, stream, q, o, D, kv_len)); break;
    case 160: KIRI_CHECK_CUDA(launch_pdl(encoder_attention_kernel<160>, grid, dim3(160), 0, stream, q, o, D, kv_len)); break;
    default: KIRI_REQUIRE(false, "kiri_encoder_attention: T=%d not in {32,64,96,128,160}", T);
  }
  KIRI_CHECK_CUDA(cudaGetLastError());
  return 0;
}
