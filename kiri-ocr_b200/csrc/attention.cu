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

#include <cstring>

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

// 2^x in one MUFU (arguments are <= 0 here; results below 2^-126 flush to zero, which a soft-max wants anyway)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// shared tile [rows][32 bf16] = 4 x 16-byte chunks per row, chunk index XOR-swizzled by row
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  return static_cast<uint32_t>(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}

// One launch covers every width group of the batch (the token stream is the concatenation of the
// groups): a CTA finds its group in a small grid-constant table and then its
// (line, head, half).  Lines of up to 96 tokens are handled by one CTA per (line, head) (T/16 warps of
// 16 query rows); longer lines by two CTAs that each own half of the query rows, because with one
// 10-warp CTA the 106 registers per thread allowed a single resident CTA per SM and the global-load
// phase of one (line, head) never overlapped the MMA phase of another.  Groups are ordered longest
// first so the short CTAs fill the tail.
static constexpr int kAttnThreads = 192;
static constexpr int kAttnMaxGroups = 8;
struct AttnGroups {
  int n;
  int cta_begin[kAttnMaxGroups + 1];
  int row0[kAttnMaxGroups];      // first token row of the group
  int line0[kAttnMaxGroups];     // first line of the group (kv_len index)
  int T[kAttnMaxGroups];
};

template <int T, int QR>
__device__ __forceinline__ void attn_tile(__nv_bfloat16* __restrict__ obase, const int q0, const int D, const int klen,
                                          const bool masked, const uint8_t* s_q, const uint8_t* s_k, const uint8_t* s_v) {
  constexpr int NB = T / 8;          // key blocks of 8
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp >= QR / 16) return;

  const int ml = warp * 16;                       // first query row of this warp inside s_q
  const uint32_t q_base = smem_u32(s_q), k_base = smem_u32(s_k), v_base = smem_u32(s_v);

  // Q fragments for the two k16 steps of head_dim 32
  uint32_t qa[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
    ldsm_x4(q_base + tile_off(ml + (lane & 15), ks * 2 + (lane >> 4)), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);

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
  const int t2 = (lane & 3) * 2;
  float mx0 = -INFINITY, mx1 = -INFINITY;
  if (masked) {                                   // one branch for the whole row, not one per key block
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const int key = nb * 8 + t2;
      if (key >= klen) { s[nb][0] = -INFINITY; s[nb][2] = -INFINITY; }
      if (key + 1 >= klen) { s[nb][1] = -INFINITY; s[nb][3] = -INFINITY; }
    }
  }
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    mx0 = fmaxf(mx0, fmaxf(s[nb][0], s[nb][1]));
    mx1 = fmaxf(mx1, fmaxf(s[nb][2], s[nb][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  const float o0 = mx0 * sl2, o1 = mx1 * sl2;
  float sum0, sum1;
  {
    // packed fp32 math: one FFMA2 scales and shifts two scores, one FADD2 accumulates two probabilities
    const float2 sc = make_float2(sl2, sl2), n0 = make_float2(-o0, -o0), n1 = make_float2(-o1, -o1);
    float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const float2 a = ffma2(make_float2(s[nb][0], s[nb][1]), sc, n0);
      const float2 b = ffma2(make_float2(s[nb][2], s[nb][3]), sc, n1);
      s[nb][0] = ex2_approx(a.x); s[nb][1] = ex2_approx(a.y);
      s[nb][2] = ex2_approx(b.x); s[nb][3] = ex2_approx(b.y);
      acc0 = fadd2(acc0, make_float2(s[nb][0], s[nb][1]));
      acc1 = fadd2(acc1, make_float2(s[nb][2], s[nb][3]));
    }
    sum0 = acc0.x + acc0.y;
    sum1 = acc1.x + acc1.y;
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
  __nv_bfloat16* orow0 = obase + static_cast<size_t>(q0 + ml + g) * D + t2;
  __nv_bfloat16* orow1 = orow0 + static_cast<size_t>(8) * D;
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) {
    *reinterpret_cast<uint32_t*>(orow0 + nb * 8) = pack_bf16x2(o[nb][0] * r0, o[nb][1] * r0);
    *reinterpret_cast<uint32_t*>(orow1 + nb * 8) = pack_bf16x2(o[nb][2] * r1, o[nb][3] * r1);
  }
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

static constexpr int kAttnBufBytes = (96 + 2 * 160) * 64;     // Q (<= 96 rows) | K | V of one work item

struct AttnItem { int T, q0, qr, klen; const __nv_bfloat16* base; __nv_bfloat16* obase; };

// Persistent CTAs: the Q/K/V slices of the NEXT (line, head, half) stream into the other shared-memory
// buffer with cp.async while the current one is computed (the one-shot form spent 48 % of its issue slots
// in long-scoreboard stalls on these loads at 25 % occupancy, profiles/r01_ncu_full_summary_v2.json).
__global__ void __launch_bounds__(kAttnThreads, 3)
encoder_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, const int D, const int heads,
                         const int* __restrict__ kv_len, const int total_items, const __grid_constant__ AttnGroups G) {
  extern __shared__ __align__(128) uint8_t s_attn[];
  const int tid = threadIdx.x;
  const bool masked = kv_len != nullptr;
  auto decode = [&](int item) -> AttnItem {
    int gi = 0;
#pragma unroll
    for (int i = 1; i < kAttnMaxGroups; ++i)
      if (i < G.n && item >= G.cta_begin[i]) gi = i;
    AttnItem w;
    w.T = G.T[gi];
    const int hs = w.T > 96 ? 1 : 0;                // two CTAs (query halves) per (line, head) above 96 tokens
    const int halves = 1 << hs;
    int local = item - G.cta_begin[gi];
    const int half = local & (halves - 1);
    local >>= hs;
    const int head = local % heads;
    const int line = local / heads;
    w.qr = w.T / halves;
    w.q0 = half * w.qr;
    const size_t row0 = static_cast<size_t>(G.row0[gi]) + static_cast<size_t>(line) * w.T;
    w.base = qkv + row0 * 3 * D + head * kHd;
    w.obase = out + row0 * D + head * kHd;
    w.klen = masked ? kv_len[G.line0[gi] + line] : w.T;
    return w;
  };
  // A thread owns one 16-byte chunk column c and the rows r0, r0 + 48, r0 + 96, ...: (row >> 1) & 3 is the same for all
  // of them (48 / 2 is a multiple of 4), so the swizzled chunk is fixed and a copy costs two adds, not an index decode.
  static_assert(kAttnThreads == 192, "issue_load: 48 rows per pass");
  auto issue_load = [&](const AttnItem& w, uint8_t* bufp) {
    const uint32_t sq = smem_u32(bufp), sk = sq + 96 * 64, sv = sk + 160 * 64;
    const size_t ld = static_cast<size_t>(3) * D;
    const int c = tid & 3, r0 = tid >> 2;
    const uint32_t off0 = tile_off(r0, c);
    const __nv_bfloat16* src0 = w.base + r0 * ld + c * 8;
    {
      const __nv_bfloat16* src = src0 + D;          // K; V is D further on
      uint32_t off = off0;
#pragma unroll 1
      for (int r = r0; r < w.T; r += 48, off += 48 * 64, src += 48 * ld) {
        cp_async16(sk + off, src);
        cp_async16(sv + off, src + D);
      }
    }
    {
      const __nv_bfloat16* src = src0 + w.q0 * ld;
      uint32_t off = off0;
#pragma unroll 1
      for (int r = r0; r < w.qr; r += 48, off += 48 * 64, src += 48 * ld) cp_async16(sq + off, src);
    }
  };
  pdl_trigger();
  pdl_wait();                                       // qkv comes from the previous kernel
  int item = blockIdx.x;
  int b = 0;
  AttnItem cur = {};
  if (item < total_items) { cur = decode(item); issue_load(cur, s_attn); }
  cp_async_commit();
  while (item < total_items) {
    const int nxt_item = item + gridDim.x;
    AttnItem nxt = {};
    if (nxt_item < total_items) { nxt = decode(nxt_item); issue_load(nxt, s_attn + (b ^ 1) * kAttnBufBytes); }
    cp_async_commit();
    cp_async_wait<1>();                             // everything but the newest group has landed
    __syncthreads();
    const uint8_t* s_q = s_attn + b * kAttnBufBytes;
    const uint8_t* s_k = s_q + 96 * 64;
    const uint8_t* s_v = s_k + 160 * 64;
    switch (cur.T) {
      case 32:  attn_tile<32, 32>(cur.obase, cur.q0, D, cur.klen, masked, s_q, s_k, s_v); break;
      case 64:  attn_tile<64, 64>(cur.obase, cur.q0, D, cur.klen, masked, s_q, s_k, s_v); break;
      case 96:  attn_tile<96, 96>(cur.obase, cur.q0, D, cur.klen, masked, s_q, s_k, s_v); break;
      case 128: attn_tile<128, 64>(cur.obase, cur.q0, D, cur.klen, masked, s_q, s_k, s_v); break;
      default:  attn_tile<160, 80>(cur.obase, cur.q0, D, cur.klen, masked, s_q, s_k, s_v); break;
    }
    __syncthreads();                                // buffer b is refilled by the next iteration's loads
    cur = nxt;
    item = nxt_item;
    b ^= 1;
  }
}

}  // namespace kiri

using namespace kiri;

extern "C" int kiri_encoder_attention_multi(const void* qkv_bf16, void* out_bf16, const int* group_lines, const int* group_T,
                                            int n_groups, int heads, int D, const int* kv_len, cudaStream_t stream) {
  KIRI_REQUIRE(qkv_bf16 && out_bf16 && group_lines && group_T, "kiri_encoder_attention_multi: null pointer");
  KIRI_REQUIRE(D == heads * kHd, "kiri_encoder_attention: head_dim must be 32 (D=%d, heads=%d)", D, heads);
  KIRI_REQUIRE(n_groups >= 0 && n_groups <= kAttnMaxGroups, "kiri_encoder_attention_multi: at most %d groups", kAttnMaxGroups);
  // token rows / line indices follow the caller's group order; CTAs are dealt longest group first
  int row0[kAttnMaxGroups], line0[kAttnMaxGroups], order[kAttnMaxGroups];
  long long r = 0;
  int l = 0, n_used = 0;
  for (int g = 0; g < n_groups; ++g) {
    const int T = group_T[g];
    KIRI_REQUIRE(T == 32 || T == 64 || T == 96 || T == 128 || T == 160, "kiri_encoder_attention: T=%d not in {32,64,96,128,160}", T);
    KIRI_REQUIRE(group_lines[g] >= 0 && r < (1ll << 31), "kiri_encoder_attention_multi: bad group %d", g);
    row0[g] = static_cast<int>(r);
    line0[g] = l;
    r += static_cast<long long>(group_lines[g]) * T;
    l += group_lines[g];
    if (group_lines[g] > 0) order[n_used++] = g;
  }
  if (n_used == 0) return 0;
  for (int i = 1; i < n_used; ++i)                  // insertion sort by T, descending
    for (int j = i; j > 0 && group_T[order[j]] > group_T[order[j - 1]]; --j) { const int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t; }
  AttnGroups G;
  memset(&G, 0, sizeof(G));
  G.n = n_used;
  long long ctas = 0;
  for (int i = 0; i < n_used; ++i) {
    const int g = order[i];
    G.cta_begin[i] = static_cast<int>(ctas);
    G.row0[i] = row0[g];
    G.line0[i] = line0[g];
    G.T[i] = group_T[g];
    ctas += static_cast<long long>(group_lines[g]) * heads * (group_T[g] > 96 ? 2 : 1);
    KIRI_REQUIRE(ctas < 0x7fffffffll, "kiri_encoder_attention_multi: grid too large");
  }
  for (int i = n_used; i <= kAttnMaxGroups; ++i) G.cta_begin[i] = static_cast<int>(ctas);
  static int sms_dev[kMaxDevices] = {0};              // function attributes are per device
  int& sms = sms_dev[kiri_cur_device_slot()];
  if (sms == 0) {
    int dev = 0, n = 0;
    KIRI_CHECK_CUDA(cudaGetDevice(&dev));
    KIRI_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    KIRI_CHECK_CUDA(cudaFuncSetAttribute(encoder_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kAttnBufBytes));
    cudaFuncSetAttribute(encoder_attention_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    sms = n;
  }
  const long long cap = static_cast<long long>(sms) * 3;          // three resident CTAs per SM walk the work items
  const unsigned grid = static_cast<unsigned>(ctas < cap ? ctas : cap);
  KIRI_CHECK_CUDA(launch_pdl(encoder_attention_kernel, dim3(grid), dim3(kAttnThreads), 2 * kAttnBufBytes, stream,
                             reinterpret_cast<const __nv_bfloat16*>(qkv_bf16), reinterpret_cast<__nv_bfloat16*>(out_bf16), D, heads,
                             kv_len, static_cast<int>(ctas), G));
  KIRI_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int kiri_encoder_attention(const void* qkv_bf16, void* out_bf16, int n_lines, int T, int heads,
                                      int D, const int* kv_len, cudaStream_t stream) {
  return kiri_encoder_attention_multi(qkv_bf16, out_bf16, &n_lines, &T, 1, heads, D, kv_len, stream);
}
