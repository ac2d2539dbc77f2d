// K12 (fused): the whole greedy attention decode of a batch in ONE persistent cluster kernel.
//
// Replaces  beam_decode_one_batched at BEAM=1   kiri_ocr/model.py:390-600 (core.py:560-568)
//           greedy_decode_streaming (token rule) kiri_ocr/model.py:779-946
//
// The step-per-launch decoder (decoder.cu) spends ~430 us per step in 27 dependent small kernels
// (profiles/r01_launches_acc.csv).  The decode step is a latency / weight-streaming problem
// (M = lines, 6.3 MB of bf16 weights and ~0.5 MB of cross K/V per line per step), so here:
//   * a thread-block CLUSTER of CS CTAs owns 16 lines (the M of mma.m16n8k16) for the whole decode:
//     no launches, no host polling, per-cluster early exit when its 16 lines are done;
//   * the cluster is tensor-parallel: CTA r owns heads [r*8/CS, ...) of both attentions and a 1/CS
//     column slice of every projection, so each SM streams only 1/CS of the weights per step;
//     slices are exchanged with DSMEM stores (st.shared::cluster) + barrier.cluster, 19 per step;
//   * weights are pre-packed in mma B-fragment order (pack_frag_kernel), so a warp streams its
//     n-tile with fully coalesced 512-byte LDG.128 straight from L2 into registers - no staging;
//   * residual stream, LayerNorms, LM fusion, the four cumulative repeat penalties, arg-max and
//     the stop rule are replicated in every CTA of the cluster (identical inputs -> identical
//     tokens), so no token broadcast is needed; rank 0 writes the global outputs.
// Numerics follow decoder.cu: bf16 operands, fp32 accumulate, fp32 residual stream.
#include <cooperative_groups.h>

#include <cstdlib>
#include <cstring>

#include "internal.cuh"
#include "ln_utils.cuh"

namespace cg = cooperative_groups;

namespace kiri {

static constexpr int kFL = 16;          // lines per cluster (MMA M)
static constexpr int kFWarps = 16;
static constexpr int kFThreads = kFWarps * 32;
static constexpr int kHd = 32;
static constexpr int kHeads = 8;
static constexpr int kTokBOS = 1, kTokEOS = 2;
static constexpr int kPad = 8;          // bf16 elements of row padding (bank spread for A fragments)

struct FusedLayer {
  const uint4 *wqkv, *wo, *wcq, *wco, *w1, *w2;      // fragment-packed bf16
  const float *bqkv, *bo, *bcq, *bco, *b1, *b2;
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *ln3_g, *ln3_b;
};
struct FusedArgs {
  FusedLayer layer[KIRI_MAX_LAYERS];
  const uint4* wheads; const float* bheads;
  const float *dec_ln_g, *dec_ln_b, *emb, *pe;
  int layers, ff, Vd, Vp, has_pos;
  const __nv_bfloat16* crosskv; int crosskv_ld;       // [rows, layers*2*D], or head-major per line (kv_hm)
  int kv_hm;                                          // 1: line block = [layer][K|V][head][t][32] (crosskv_headmajor)
  const int* mem_row0; const int* mem_len;            // per line (nullable -> b*T, T)
  int T;
  __nv_bfloat16 *self_k, *self_v;                     // [layers][B][Lmax][D]
  const int* len_est; const int* forced;
  const int* line_perm;                               // nullable: decode slot -> line (longest lines first; -1 = empty slot)
  int B, Lmax;
  KiriDecodeParams p;
  int *ids, *n_out; float *sum_logp, *step_logp, *step_prob;
  int* steps_max;                                     // device int: max steps run by any cluster
  // beam search (model.py:536-559): `beam` hypotheses per line share a cluster (16 / beam lines each)
  int beam;                                           // hypotheses per line (1 when greedy)
  int bmode;                                          // 1: beam bookkeeping + bm_* outputs, 0: greedy outputs
  double lenp;                                        // cfg.BEAM_LENP
  int* seqbuf; float* lpbuf;                          // [2][n_slots][Lmax] ping-pong hypothesis records (rank 0)
  double* bm_score; int* bm_len; int* bm_state;       // [B, beam]
  int* bm_ids; float* bm_logp;                        // [B, beam, Lmax]
  int timing;                                         // 1: accumulate phase cycles into g_dec_prof
  // live streaming: outputs may live in mapped host memory; every step is published with
  // system-scope fences so that a host thread polling `progress` sees the step's ids / log-probs
  int prefetch;                                       // 1: L2 prefetch of the cross K/V blocks a layer ahead of their use
  int publish;                                        // 1: fence before the progress words
  int* progress;                                      // nullable [B]: steps available | (1 << 30) once the line is done
  int stream_rule;                                    // beam: 1 = beam_decode_streaming's rule (prune by score / L^lenp, stop
                                                      //       when the best hypothesis ended, model.py:1112-1150)
  int* bm_trace;                                      // nullable [B, Lmax, beam, 3]: per step and kept hypothesis (rank
                                                      //       order): parent rank | appended token (-1 carried) | logp bits
};
static constexpr int kDoneBit = 1 << 30;

// ---------------------------------------------------------------- weight packing
// dst word ((nt*K/32 + kt)*32 + lane)*4 + wd  =  W[nt*8 + lane/4][kt*32 + wd*8 + (lane%4)*2 .. +1]
__global__ void pack_frag_kernel(const __nv_bfloat16* __restrict__ w, int N, int K, uint32_t* __restrict__ dst) {
  const size_t total = static_cast<size_t>(N) * K / 2;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int wd = i & 3, lane = (i >> 2) & 31;
    const size_t t = i >> 7;
    const int kt = t % (K / 32), nt = t / (K / 32);
    const int n = nt * 8 + (lane >> 2), k = kt * 32 + wd * 8 + (lane & 3) * 2;
    dst[i] = *reinterpret_cast<const uint32_t*>(w + static_cast<size_t>(n) * K + k);
  }
}

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// c += A[16 x (k32 tiles kb..ke)] * W_ntile^T ; wp points at the n-tile's packed stream
__device__ __forceinline__ void mma_ntile(float (&c)[4], const uint4* __restrict__ wp, const __nv_bfloat16* A, int lda,
                                          int kb, int ke, int lane) {
  const int r = lane >> 2, cq = (lane & 3) * 2;
  const __nv_bfloat16* a_lo = A + r * lda + cq;
  const __nv_bfloat16* a_hi = a_lo + 8 * lda;
#pragma unroll 8
  for (int kt = kb; kt < ke; ++kt) {
    const uint4 w = ldg_stream(wp + kt * 32 + lane);
    const int k = kt * 32;
    const uint32_t a0 = *reinterpret_cast<const uint32_t*>(a_lo + k), a1 = *reinterpret_cast<const uint32_t*>(a_hi + k);
    const uint32_t a2 = *reinterpret_cast<const uint32_t*>(a_lo + k + 8), a3 = *reinterpret_cast<const uint32_t*>(a_hi + k + 8);
    mma_bf16_16816(c, a0, a1, a2, a3, w.x, w.y);
    const uint32_t e0 = *reinterpret_cast<const uint32_t*>(a_lo + k + 16), e1 = *reinterpret_cast<const uint32_t*>(a_hi + k + 16);
    const uint32_t e2 = *reinterpret_cast<const uint32_t*>(a_lo + k + 24), e3 = *reinterpret_cast<const uint32_t*>(a_hi + k + 24);
    mma_bf16_16816(c, e0, e1, e2, e3, w.z, w.w);
  }
}

// Y[16 x (8*nt_count)] = A[16 x K] * W^T for the n-tiles nt_of(0..nt_count-1) of a packed matrix.
// epi(row, global_col (even), v0, v1) is called exactly once per column pair.  Ends with all
// epilogues issued (caller synchronises).  `part` is shared scratch of >= 16*512 bytes.
template <class NtOf, class Epi>
__device__ __forceinline__ void cta_gemm(const uint4* __restrict__ wp, int K, const __nv_bfloat16* A, int lda,
                                         int nt_count, NtOf nt_of, float* part, Epi epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k32 = K / 32;
  int KS = 1;
  while (nt_count * KS * 2 <= kFWarps && (k32 % (KS * 2)) == 0 && KS < 8) KS *= 2;
  const int items = nt_count * KS;
  const int kper = k32 / KS;
  for (int item = warp; item < items; item += kFWarps) {
    const int ni = item / KS, ks = item - ni * KS;
    const int nt = nt_of(ni);
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    mma_ntile(c, wp + static_cast<size_t>(nt) * k32 * 32, A, lda, ks * kper, (ks + 1) * kper, lane);
    if (KS == 1) {
      const int r = lane >> 2, col = nt * 8 + (lane & 3) * 2;
      epi(r, col, c[0], c[1]);
      epi(r + 8, col, c[2], c[3]);
    } else {
      *reinterpret_cast<float4*>(part + (item * 32 + lane) * 4) = make_float4(c[0], c[1], c[2], c[3]);
    }
  }
  if (KS > 1) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < nt_count * 64; idx += kFThreads) {
      const int ni = idx >> 6, e = idx & 63, l = e >> 1, hi = e & 1;
      float v0 = 0.f, v1 = 0.f;
      for (int ks = 0; ks < KS; ++ks) {
        const float2 pv = *reinterpret_cast<const float2*>(part + ((ni * KS + ks) * 32 + l) * 4 + hi * 2);
        v0 += pv.x; v1 += pv.y;
      }
      epi((l >> 2) + hi * 8, nt_of(ni) * 8 + (l & 3) * 2, v0, v1);
    }
  }
}

// per-line decode state, replicated in every CTA of the cluster
struct LineState {
  int hist[kFL][8];      // sequence so far (ring is unnecessary: only the last 6 are needed) - kept as last-8 window
  int n_tok[kFL];
  int finished[kFL];
  int max_steps[kFL];
  int target[kFL];
  int cur_tok[kFL];
  int valid[kFL];
  int row0[kFL];
  int mlen[kFL];
  int line[kFL];         // global line index of each slot (outputs, forced ids)
};

// beam bookkeeping, replicated in every CTA of the cluster
struct BeamState {
  double score[kFL];     // sum of chosen log-probs (Python float in the reference)
  int state[kFL];        // 0 empty, 1 alive, 2 done
  float cand_v[kFL][8];  // top-`beam` penalised log-probs of every alive slot
  int cand_i[kFL][8];
  int line_done[kFL];    // per line of the cluster
  int src[kFL];          // scratch of the selection: old slot each new slot descends from,
  int new_tok[kFL];      //   appended token (-1: a finished hypothesis carried over),
  float new_v[kFL];      //   its penalised log-prob,
  int old_len[kFL];      //   tokens after BOS before this step
};
struct FusedSmem {                       // byte offsets into dynamic shared memory
  int x, gath, a, obuf, hbuf, qloc, logits, part, vstage, state, params, beam, anc, total;
};
// Biases and LayerNorm affines of every layer live in shared memory for the whole decode: every
// cluster barrier invalidates L1, so a parameter read from global costs an L2 round trip per phase.
// Per layer (floats): bqkv 768 | bo 256 | bcq 256 | bco 256 | b1 ff | b2 256 | ln1 g,b | ln2 g,b | ln3 g,b
__host__ __device__ inline int fused_layer_floats(int ff) { return 768 + 4 * 256 + ff + 6 * 256; }
__host__ __device__ inline FusedSmem fused_smem_plan(int ff, int Vp, int cs, int layers, int beam = 0, int Lmax = 0) {
  FusedSmem s;
  int off = 0;
  auto take = [&](int bytes) { const int o = off; off += (bytes + 127) & ~127; return o; };
  s.x = take(kFL * 256 * 4);
  s.gath = take(kFL * 256 * 4);
  s.a = take(kFL * (256 + kPad) * 2);
  s.obuf = take(kFL * (256 + kPad) * 2);
  s.hbuf = take(kFL * (ff + kPad) * 2);
  s.qloc = take(kFL * (256 / cs) * 4);
  s.logits = take(kFL * 2 * Vp * 4);
  s.part = take(kFWarps * 512);
  s.vstage = take(0);                                  // (round 1: V staging tiles of the attention; no longer used)
  s.state = take(static_cast<int>(sizeof(LineState)));
  s.params = take((layers * fused_layer_floats(ff) + 2 * Vp + 512) * 4);
  s.beam = take(beam >= 1 ? static_cast<int>(sizeof(BeamState)) : 0);       // beam = 0: greedy, no bookkeeping
  s.anc = take(beam >= 1 ? 2 * kFL * ((Lmax + 15) & ~15) : 0);      // [2][slot][pos] -> physical K/V slot
  s.total = off;
  return s;
}

// Single-query attention of one warp over n keys; K/V rows are 32 bf16 (64 B) at base + row_off(j) elements.
// q: 32 fp32 in shared memory.  Returns o[lane] (output dimension = lane).
//
// Lane l OWNS keys l, l+32, ... of a group of kAttC*32 keys: it computes their scores with a 32-wide dot product on its
// own K rows and accumulates its own V rows into a private 32-dim partial sum; one 31-step butterfly at the very end
// moves dimension d to lane d.  (Round 1 gave every lane one output DIMENSION instead: per 32-key chunk the V rows went
// through a shared-memory staging tile and every key cost a shuffle + a shared load + a convert + an FMA in a serial
// chain - ~2000 warp instructions per (line, head) at T = 160, measured 11.4 us per layer; this form needs ~800 and its
// 16 resident warps hide the row loads.)  Groups are merged with the running-max rule of online soft-max.
static constexpr int kAttC = 5;          // 32-key chunks per group: one group covers T <= 160 (every cross-attention)

// MULTI = false: n <= kAttC*32 (one group; the output partial sums are not live during the score pass);
// MULTI = true: any n, groups merged with the running-max rule of online soft-max.
// Row j of the K / V operand lives at base + j*stride (+ (slot0 + anc[j]) * slot_stride when `anc` is given: beam
// hypotheses read position j from the physical cache slot their ancestor wrote, see the ancestor table below).
// One row per lane is in flight at a time: with two or three (tried as explicit batches, inlined and as separate
// functions) ptxas front-batches the loads of the whole unrolled loop and spills 1-5 KB at the kernel's 128-register cap.
template <bool READONLY, bool MULTI>
__device__ __forceinline__ float attend_warp(const float* q, const __nv_bfloat16* kbase, const __nv_bfloat16* vbase,
                                          size_t stride, const uint8_t* anc, size_t slot_stride, int slot0, int n, int lane) {
  auto row_off = [&](int j) -> size_t {
    return static_cast<size_t>(j) * stride + (anc ? static_cast<size_t>(slot0 + anc[j]) * slot_stride : 0);
  };
  const float4* q4 = reinterpret_cast<const float4*>(q);       // broadcast LDS.128 per use: q stays out of the registers
  const float scale = 0.17677669529663687f;           // 1/sqrt(32)
  auto ld_row = [&](const __nv_bfloat16* base, int j, uint4 (&r)[4]) {
    if (j < n) {
      const uint4* p = reinterpret_cast<const uint4*>(base + row_off(j));
#pragma unroll
      for (int i = 0; i < 4; ++i) r[i] = READONLY ? __ldg(p + i) : p[i];
    }
  };
  auto dot_row = [&](const uint4 (&r)[4]) -> float {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 u = r[i];
      const float4 qa = q4[2 * i], qb = q4[2 * i + 1];
      acc = fmaf(qa.x, bf16_lo(u.x), acc); acc = fmaf(qa.y, bf16_hi(u.x), acc);
      acc = fmaf(qa.z, bf16_lo(u.y), acc); acc = fmaf(qa.w, bf16_hi(u.y), acc);
      acc = fmaf(qb.x, bf16_lo(u.z), acc); acc = fmaf(qb.y, bf16_hi(u.z), acc);
      acc = fmaf(qb.z, bf16_lo(u.w), acc); acc = fmaf(qb.w, bf16_hi(u.w), acc);
    }
    return acc;
  };
  constexpr int RBK = MULTI ? 1 : 1;                   // K rows in flight per lane
  constexpr int RBV = 1;                               // V rows in flight per lane (32 partial sums are live)
  float m_run = -INFINITY, l_run = 0.f;
  float o[kHd];
  if (MULTI) {
#pragma unroll
    for (int d = 0; d < kHd; ++d) o[d] = 0.f;
  }
  const int n_grp = MULTI ? n : 1;
  for (int g0 = 0; g0 < n_grp; g0 += kAttC * 32) {
    // ---- scores of this lane's keys (rows of a batch are loaded together: one memory latency per batch)
    float sc[kAttC];
#pragma unroll
    for (int c0 = 0; c0 < kAttC; c0 += RBK) {
      uint4 rr[RBK][4];
#pragma unroll
      for (int r = 0; r < RBK; ++r)
        if (c0 + r < kAttC) ld_row(kbase, g0 + lane + 32 * (c0 + r), rr[r]);
#pragma unroll
      for (int r = 0; r < RBK; ++r)
        if (c0 + r < kAttC) {
          const int j = g0 + lane + 32 * (c0 + r);
          sc[c0 + r] = (j < n) ? dot_row(rr[r]) * scale : -INFINITY;
        }
    }
    // ---- soft-max bookkeeping (running max over groups)
    float gm = sc[0];
#pragma unroll
    for (int c = 1; c < kAttC; ++c) gm = fmaxf(gm, sc[c]);
    const float m_new = fmaxf(m_run, warp_max(gm));
    const float corr = __expf(m_run - m_new);          // 0 for the first group (m_run = -inf)
    float ps = 0.f;
#pragma unroll
    for (int c = 0; c < kAttC; ++c) { sc[c] = __expf(sc[c] - m_new); ps += sc[c]; }     // exp(-inf) = 0 past the end
    l_run = l_run * corr + warp_sum(ps);
    if (MULTI) {
      if (g0 > 0) {
#pragma unroll
        for (int d = 0; d < kHd; ++d) o[d] *= corr;
      }
    } else {
#pragma unroll
      for (int d = 0; d < kHd; ++d) o[d] = 0.f;
    }
    m_run = m_new;
    // ---- this lane's V rows into its private partial sum
#pragma unroll
    for (int c0 = 0; c0 < kAttC; c0 += RBV) {
      uint4 rr[RBV][4];
#pragma unroll
      for (int r = 0; r < RBV; ++r)
        if (c0 + r < kAttC) ld_row(vbase, g0 + lane + 32 * (c0 + r), rr[r]);
#pragma unroll
      for (int r = 0; r < RBV; ++r)
        if (c0 + r < kAttC && g0 + lane + 32 * (c0 + r) < n) {
          const float pj = sc[c0 + r];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 u = rr[r][i];
            o[8 * i + 0] = fmaf(pj, bf16_lo(u.x), o[8 * i + 0]); o[8 * i + 1] = fmaf(pj, bf16_hi(u.x), o[8 * i + 1]);
            o[8 * i + 2] = fmaf(pj, bf16_lo(u.y), o[8 * i + 2]); o[8 * i + 3] = fmaf(pj, bf16_hi(u.y), o[8 * i + 3]);
            o[8 * i + 4] = fmaf(pj, bf16_lo(u.z), o[8 * i + 4]); o[8 * i + 5] = fmaf(pj, bf16_hi(u.z), o[8 * i + 5]);
            o[8 * i + 6] = fmaf(pj, bf16_lo(u.w), o[8 * i + 6]); o[8 * i + 7] = fmaf(pj, bf16_hi(u.w), o[8 * i + 7]);
          }
        }
    }
  }
  // ---- butterfly: sum the 32 partial vectors over the lanes, dimension d ends up in lane d
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool up = (lane & w) != 0;                   // this lane keeps the upper half of what it still holds
#pragma unroll
    for (int k = 0; k < w; ++k) {
      const float keep = up ? o[k + w] : o[k];
      const float send = up ? o[k] : o[k + w];
      o[k] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  return o[0] / l_run;
}

// ln8 (ln_utils.cuh) with the affine in shared memory (plain loads)
__device__ __forceinline__ void ln8s(float (&v)[8], const float* g, const float* b, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / kD);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q += d * d; }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / kD) + kLnEps);
  const float4 g0 = *reinterpret_cast<const float4*>(g + lane * 8), g1 = *reinterpret_cast<const float4*>(g + lane * 8 + 4);
  const float4 b0 = *reinterpret_cast<const float4*>(b + lane * 8), b1 = *reinterpret_cast<const float4*>(b + lane * 8 + 4);
  const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * gg[i] + bb[i];
}

__device__ __forceinline__ float warp_lse_s(const float* x, int n, int lane) {
  float m = -INFINITY;
  for (int v = lane; v < n; v += 32) m = fmaxf(m, x[v]);
  m = warp_max(m);
  float s = 0.f;
  for (int v = lane; v < n; v += 32) s += expf(x[v] - m);
  s = warp_sum(s);
  return m + logf(s);
}

// Phase timing of cluster 0 / rank 0 (clock64 deltas), read back with kiri_debug_decode_timing().
__device__ long long g_dec_prof[32];
__device__ long long g_dec_clu[64 * 4];     // per cluster: start ns, end ns, steps, smid of rank 0
#define DEC_TICK(slot)                                                   \
  do {                                                                   \
    if (A.timing && blockIdx.x == 0 && threadIdx.x == 0) {               \
      const long long t_now = clock64();                                 \
      g_dec_prof[slot] += t_now - t_last;                                \
      t_last = t_now;                                                    \
    }                                                                    \
  } while (0)

template <int CS>
__global__ void __launch_bounds__(kFThreads, 1) dec_fused_kernel(const __grid_constant__ FusedArgs A) {
  extern __shared__ __align__(128) uint8_t sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (CS == 1) ? 0 : static_cast<int>(cluster.block_rank());
  const int cid = blockIdx.x / CS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int HPC = kHeads / CS;                 // heads per CTA
  constexpr int DC = 256 / CS;                     // output columns of a D-wide projection per CTA
  const FusedSmem L = fused_smem_plan(A.ff, A.Vp, CS, A.layers, A.bmode ? A.beam : 0, A.Lmax);
  float* x = reinterpret_cast<float*>(sm + L.x);
  float* gath = reinterpret_cast<float*>(sm + L.gath);
  __nv_bfloat16* a = reinterpret_cast<__nv_bfloat16*>(sm + L.a);
  __nv_bfloat16* obuf = reinterpret_cast<__nv_bfloat16*>(sm + L.obuf);
  __nv_bfloat16* hbuf = reinterpret_cast<__nv_bfloat16*>(sm + L.hbuf);
  float* qloc = reinterpret_cast<float*>(sm + L.qloc);
  float* logits = reinterpret_cast<float*>(sm + L.logits);
  float* part = reinterpret_cast<float*>(sm + L.part);
  LineState* st = reinterpret_cast<LineState*>(sm + L.state);
  float* prm = reinterpret_cast<float*>(sm + L.params);
  const int lfl = fused_layer_floats(A.ff);
  float* p_heads = prm + A.layers * lfl;            // 2*Vp head biases, then dec_ln g | b
  float* p_decln = p_heads + 2 * A.Vp;
  for (int l = 0; l < A.layers; ++l) {
    const FusedLayer& W = A.layer[l];
    float* d = prm + l * lfl;
    for (int t = threadIdx.x; t < 768; t += kFThreads) d[t] = __ldg(W.bqkv + t);
    for (int t = threadIdx.x; t < 256; t += kFThreads) {
      d[768 + t] = __ldg(W.bo + t); d[1024 + t] = __ldg(W.bcq + t); d[1280 + t] = __ldg(W.bco + t);
      d[1536 + A.ff + t] = __ldg(W.b2 + t);
      float* n = d + 1792 + A.ff;
      n[t] = __ldg(W.ln1_g + t); n[256 + t] = __ldg(W.ln1_b + t); n[512 + t] = __ldg(W.ln2_g + t);
      n[768 + t] = __ldg(W.ln2_b + t); n[1024 + t] = __ldg(W.ln3_g + t); n[1280 + t] = __ldg(W.ln3_b + t);
    }
    for (int t = threadIdx.x; t < A.ff; t += kFThreads) d[1536 + t] = __ldg(W.b1 + t);
  }
  for (int t = threadIdx.x; t < 2 * A.Vp; t += kFThreads) p_heads[t] = __ldg(A.bheads + t);
  for (int t = threadIdx.x; t < 256; t += kFThreads) { p_decln[t] = __ldg(A.dec_ln_g + t); p_decln[256 + t] = __ldg(A.dec_ln_b + t); }
  constexpr int lda = 256 + kPad;
  const int ldh = A.ff + kPad;
  const int b0 = cid * kFL;                          // first physical slot (K/V cache, hypothesis records)
  const int D = 256;
  const int BEAM = A.beam;
  const bool BM = A.bmode != 0;                      // beam bookkeeping (even at width 1) vs greedy
  const int LPC = kFL / BEAM;                        // lines per cluster (16 when greedy)
  BeamState* bs = reinterpret_cast<BeamState*>(sm + L.beam);
  const int anc_ld = (A.Lmax + 15) & ~15;
  uint8_t* anc = sm + L.anc;                         // [2][slot][pos], beam mode only
  int pp = 0;                                        // ping-pong index of anc / seqbuf / lpbuf

  // remote views of the exchange buffers
  __nv_bfloat16* r_obuf[CS]; float* r_gath[CS]; __nv_bfloat16* r_hbuf[CS]; float* r_logits[CS];
#pragma unroll
  for (int d = 0; d < CS; ++d) {
    if (CS == 1) { r_obuf[d] = obuf; r_gath[d] = gath; r_hbuf[d] = hbuf; r_logits[d] = logits; }
    else {
      r_obuf[d] = cluster.map_shared_rank(obuf, d); r_gath[d] = cluster.map_shared_rank(gath, d);
      r_hbuf[d] = cluster.map_shared_rank(hbuf, d); r_logits[d] = cluster.map_shared_rank(logits, d);
    }
  }
  auto csync = [&]() { if (CS == 1) __syncthreads(); else cluster.sync(); };

  // ---- init (model.py:416-425: Python float arithmetic, int() truncation)
  if (threadIdx.x < kFL) {
    const int i = threadIdx.x;
    const int li = i / BEAM, bi = i - li * BEAM;     // line of the cluster, hypothesis of the line
    const int dslot = cid * LPC + li;                // decode order index of the line
    const int pb = (li < LPC && dslot < A.B) ? (A.line_perm ? A.line_perm[dslot] : dslot) : -1;   // -1: an empty slot
    const bool ok = pb >= 0;
    const int b = ok ? pb : 0;
    st->line[i] = b;
    int ms = 0, tl = 0, Tm = A.T, r0 = 0;
    if (ok) {
      tl = A.len_est[b];
      Tm = A.mem_len ? A.mem_len[b] : A.T;
      r0 = A.mem_row0 ? A.mem_row0[b] : b * A.T;
      if (tl > 0) ms = __double2int_rz(__dmul_rn(static_cast<double>(tl), A.p.len_ratio)) + A.p.len_pad;
      else ms = __double2int_rz(__dmul_rn(static_cast<double>(Tm), A.p.mem_ratio)) + A.p.len_pad;
      if (ms > A.p.max_dec_len) ms = A.p.max_dec_len;
      if (ms > A.Lmax) ms = A.Lmax;
    }
    st->valid[i] = ok; st->row0[i] = r0; st->mlen[i] = Tm;
    st->max_steps[i] = ms; st->target[i] = tl;
    st->finished[i] = (!ok || ms <= 0 || bi != 0) ? 1 : 0;       // a line starts with ONE hypothesis [BOS]
    st->n_tok[i] = 1; st->cur_tok[i] = kTokBOS;
#pragma unroll
    for (int k = 0; k < 8; ++k) st->hist[i][k] = -1 - k;
    st->hist[i][0] = kTokBOS;
    if (ok && rank == 0 && bi == 0 && A.progress) A.progress[b] = ms <= 0 ? kDoneBit : 0;
    if (!BM) {
      if (ok && rank == 0) { A.n_out[b] = 0; A.sum_logp[b] = 0.f; }
    } else {
      bs->score[i] = 0.0;
      bs->state[i] = st->finished[i] ? 0 : 1;
      if (bi == 0) bs->line_done[li] = (!ok || ms <= 0) ? 1 : 0;
      if (ok && rank == 0) {
        // a line whose step budget is 0 keeps its initial hypothesis [BOS] (model.py:431-443)
        A.bm_state[b * BEAM + bi] = (bi == 0) ? 1 : 0;
        A.bm_score[b * BEAM + bi] = 0.0;
        A.bm_len[b * BEAM + bi] = 0;
      }
    }
  }
  __syncthreads();
  csync();                                          // every CTA of the cluster is resident and initialised

  long long t_last = clock64();
  if (A.timing && rank == 0 && threadIdx.x == 0 && cid < 64) {
    unsigned long long ns; unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_dec_clu[cid * 4 + 0] = static_cast<long long>(ns);
    g_dec_clu[cid * 4 + 3] = smid;
  }
  int step = 0;
  for (; step < A.Lmax; ++step) {
    {
      int alive = 0;
      if (!BM) {
#pragma unroll
        for (int i = 0; i < kFL; ++i) alive += st->finished[i] ? 0 : 1;
      } else {
        for (int li = 0; li < LPC; ++li) alive += bs->line_done[li] ? 0 : 1;
      }
      if (alive == 0) break;
    }
    // ---- S0: x = emb[tok] + pe[step]; a = LN1_0(x)   (one warp per line)
    {
      const int i = warp;
      if (BM && lane == 0) anc[(pp * kFL + i) * anc_ld + step] = static_cast<uint8_t>(i);
      const int tokid = st->finished[i] ? 0 : st->cur_tok[i];   // finished lines feed the pad token
      const float4* e = reinterpret_cast<const float4*>(A.emb + static_cast<size_t>(tokid) * D) + lane * 2;
      const float4 e0 = __ldg(e), e1 = __ldg(e + 1);
      float v[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
      if (A.has_pos) {
        const float4* pp = reinterpret_cast<const float4*>(A.pe + static_cast<size_t>(step) * D) + lane * 2;
        const float4 p0 = __ldg(pp), p1 = __ldg(pp + 1);
        v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w;
        v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
      }
      st_f32x8(x + i * D + lane * 8, v);
      ln8s(v, prm + 1792 + A.ff, prm + 1792 + A.ff + 256, lane);
      st_bf16x8(a + i * lda + lane * 8, v);
    }
    __syncthreads();
    DEC_TICK(0);

    for (int l = 0; l < A.layers; ++l) {
      const FusedLayer& W = A.layer[l];
      const float* PB = prm + l * lfl;                 // this layer's biases
      // The cross K/V blocks this CTA will read in phase F do not depend on the tokens: pull them from HBM into L2
      // NOW (prefetch.global.L2 needs no registers), ~10 us of other phases ahead of their use.  The per-lane loads
      // of phase F then pay an L2 hit instead of an HBM miss per 32-key chunk (the cross K/V of a 256-line batch,
      // 80-126 MB, do not stay L2-resident from one step to the next).
      if (A.kv_hm && A.prefetch) {
        const int i = warp;
        if (st->valid[i] && !st->finished[i]) {
          const int Tm = st->mlen[i];
          const int n128 = (Tm * kHd * 2 + 127) >> 7;              // 128-byte lines of one (head, K|V) block
#pragma unroll
          for (int hl = 0; hl < HPC; ++hl) {
            const int head = rank * HPC + hl;
            const char* kb = reinterpret_cast<const char*>(A.crosskv + static_cast<size_t>(st->row0[i]) * A.crosskv_ld +
                                                           (static_cast<size_t>(l) * 2 * kHeads + head) * Tm * kHd);
            const char* vb = kb + static_cast<size_t>(kHeads) * Tm * kHd * 2;
            for (int c = lane; c < n128; c += 32) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(kb + (c << 7)));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(vb + (c << 7)));
            }
          }
        }
      }
      const float* PN = PB + 1792 + A.ff;              // ln1 g,b | ln2 g,b | ln3 g,b
      // the cache is indexed by PHYSICAL slot: B slots when greedy, 16 per cluster in beam mode
      const size_t cache_slots = !BM ? static_cast<size_t>(A.B) : static_cast<size_t>(gridDim.x / CS) * kFL;
      __nv_bfloat16* kc = A.self_k + static_cast<size_t>(l) * cache_slots * A.Lmax * D;
      __nv_bfloat16* vc = A.self_v + static_cast<size_t>(l) * cache_slots * A.Lmax * D;
      // ---- A: q,k,v of my heads.  q -> qloc (fp32), k/v -> global cache row `step` (bf16)
      cta_gemm(W.wqkv, D, a, lda, 12 * HPC,
               [&](int i) { const int sec = i / (4 * HPC), rem = i - sec * 4 * HPC; return sec * 32 + rank * HPC * 4 + rem; },
               part,
               [&](int row, int col, float v0, float v1) {
                 v0 += PB[col]; v1 += PB[col + 1];
                 const int sec = col >> 8, c = col & 255;
                 if (sec == 0) {
                   *reinterpret_cast<float2*>(qloc + row * DC + (c - rank * DC)) = make_float2(v0, v1);
                 } else if (st->valid[row]) {
                   __nv_bfloat16* dst = (sec == 1 ? kc : vc) + (static_cast<size_t>(b0 + row) * A.Lmax + step) * D + c;
                   *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(v0, v1);
                 }
               });
      __syncthreads();
      DEC_TICK(1);
      // ---- B: self-attention of my heads over keys 0..step -> o slices broadcast to every CTA
      for (int pidx = warp; pidx < kFL * HPC; pidx += kFWarps) {
        const int i = pidx % kFL, hl = pidx / kFL, head = rank * HPC + hl;
        float o = 0.f;
        if (st->valid[i] && !st->finished[i]) {
          if (!BM) {
            const size_t base = (static_cast<size_t>(b0 + i) * A.Lmax) * D + head * kHd;
            if (step < kAttC * 32)
              o = attend_warp<false, false>(qloc + i * DC + hl * kHd, kc + base, vc + base, D, nullptr, 0, 0, step + 1, lane);
            else
              o = attend_warp<false, true>(qloc + i * DC + hl * kHd, kc + base, vc + base, D, nullptr, 0, 0, step + 1, lane);
          } else {
            // hypothesis i reads position j from the physical slot its ancestor wrote it to
            const uint8_t* ar = anc + (pp * kFL + i) * anc_ld;
            const size_t base = static_cast<size_t>(head) * kHd;
            o = attend_warp<false, true>(qloc + i * DC + hl * kHd, kc + base, vc + base, D, ar,
                                         static_cast<size_t>(A.Lmax) * D, b0, step + 1, lane);
          }
        }
        const float on = __shfl_down_sync(0xffffffffu, o, 1);
        if ((lane & 1) == 0) {
          const uint32_t pk = pack_bf16x2(o, on);
#pragma unroll
          for (int d = 0; d < CS; ++d) *reinterpret_cast<uint32_t*>(r_obuf[d] + i * lda + head * kHd + lane) = pk;
        }
      }
      DEC_TICK(2);
      csync();
      DEC_TICK(3);
      // ---- C: out-proj slice -> gath (all CTAs)
      cta_gemm(W.wo, D, obuf, lda, DC / 8, [&](int i) { return rank * (DC / 8) + i; }, part,
               [&](int row, int col, float v0, float v1) {
                 const float2 v = make_float2(v0 + PB[768 + col], v1 + PB[768 + col + 1]);
#pragma unroll
                 for (int d = 0; d < CS; ++d) *reinterpret_cast<float2*>(r_gath[d] + row * D + col) = v;
               });
      DEC_TICK(4);
      csync();
      DEC_TICK(5);
      // ---- D: x += gath; a = LN2(x)
      {
        const int i = warp;
        float v[8];
        const float4 x0 = *reinterpret_cast<const float4*>(x + i * D + lane * 8), x1 = *reinterpret_cast<const float4*>(x + i * D + lane * 8 + 4);
        const float4 g0 = *reinterpret_cast<const float4*>(gath + i * D + lane * 8), g1 = *reinterpret_cast<const float4*>(gath + i * D + lane * 8 + 4);
        v[0] = x0.x + g0.x; v[1] = x0.y + g0.y; v[2] = x0.z + g0.z; v[3] = x0.w + g0.w;
        v[4] = x1.x + g1.x; v[5] = x1.y + g1.y; v[6] = x1.z + g1.z; v[7] = x1.w + g1.w;
        st_f32x8(x + i * D + lane * 8, v);
        ln8s(v, PN + 512, PN + 768, lane);
        st_bf16x8(a + i * lda + lane * 8, v);
      }
      __syncthreads();
      DEC_TICK(6);
      // ---- E: cross-attention query of my heads
      cta_gemm(W.wcq, D, a, lda, DC / 8, [&](int i) { return rank * (DC / 8) + i; }, part,
               [&](int row, int col, float v0, float v1) {
                 *reinterpret_cast<float2*>(qloc + row * DC + (col - rank * DC)) =
                     make_float2(v0 + PB[1024 + col], v1 + PB[1024 + col + 1]);
               });
      __syncthreads();
      DEC_TICK(7);
      // ---- F: cross-attention over the line's memory (K | V of layer l inside the crosskv row)
      for (int pidx = warp; pidx < kFL * HPC; pidx += kFWarps) {
        const int i = pidx % kFL, hl = pidx / kFL, head = rank * HPC + hl;
        float o = 0.f;
        if (st->valid[i] && !st->finished[i]) {
          const int Tm = st->mlen[i];
          if (A.kv_hm) {
            // head-major block of the line: K of (layer, head) is Tm contiguous 64-byte rows -> a warp's
            // 32 key loads are one contiguous 2 KB run instead of 32 rows 3 KB apart
            const __nv_bfloat16* kb = A.crosskv + static_cast<size_t>(st->row0[i]) * A.crosskv_ld +
                                      (static_cast<size_t>(l) * 2 * kHeads + head) * Tm * kHd;
            if (Tm <= kAttC * 32)
              o = attend_warp<true, false>(qloc + i * DC + hl * kHd, kb, kb + static_cast<size_t>(kHeads) * Tm * kHd, kHd, nullptr, 0, 0,
                                           Tm, lane);
            else
              o = attend_warp<true, true>(qloc + i * DC + hl * kHd, kb, kb + static_cast<size_t>(kHeads) * Tm * kHd, kHd, nullptr, 0, 0,
                                          Tm, lane);
          } else {
            const __nv_bfloat16* kb = A.crosskv + static_cast<size_t>(st->row0[i]) * A.crosskv_ld + l * 2 * D + head * kHd;
            const size_t ldk = A.crosskv_ld;
            o = attend_warp<true, true>(qloc + i * DC + hl * kHd, kb, kb + D, ldk, nullptr, 0, 0, Tm, lane);
          }
        }
        const float on = __shfl_down_sync(0xffffffffu, o, 1);
        if ((lane & 1) == 0) {
          const uint32_t pk = pack_bf16x2(o, on);
#pragma unroll
          for (int d = 0; d < CS; ++d) *reinterpret_cast<uint32_t*>(r_obuf[d] + i * lda + head * kHd + lane) = pk;
        }
      }
      DEC_TICK(8);
      csync();
      DEC_TICK(9);
      // ---- G: cross out-proj slice -> gath
      cta_gemm(W.wco, D, obuf, lda, DC / 8, [&](int i) { return rank * (DC / 8) + i; }, part,
               [&](int row, int col, float v0, float v1) {
                 const float2 v = make_float2(v0 + PB[1280 + col], v1 + PB[1280 + col + 1]);
#pragma unroll
                 for (int d = 0; d < CS; ++d) *reinterpret_cast<float2*>(r_gath[d] + row * D + col) = v;
               });
      DEC_TICK(10);
      csync();
      DEC_TICK(11);
      // ---- H: x += gath; a = LN3(x)
      {
        const int i = warp;
        float v[8];
        const float4 x0 = *reinterpret_cast<const float4*>(x + i * D + lane * 8), x1 = *reinterpret_cast<const float4*>(x + i * D + lane * 8 + 4);
        const float4 g0 = *reinterpret_cast<const float4*>(gath + i * D + lane * 8), g1 = *reinterpret_cast<const float4*>(gath + i * D + lane * 8 + 4);
        v[0] = x0.x + g0.x; v[1] = x0.y + g0.y; v[2] = x0.z + g0.z; v[3] = x0.w + g0.w;
        v[4] = x1.x + g1.x; v[5] = x1.y + g1.y; v[6] = x1.z + g1.z; v[7] = x1.w + g1.w;
        st_f32x8(x + i * D + lane * 8, v);
        ln8s(v, PN + 1024, PN + 1280, lane);
        st_bf16x8(a + i * lda + lane * 8, v);
      }
      __syncthreads();
      DEC_TICK(12);
      // ---- I: FFN first linear + GELU(erf) slice -> hbuf (all CTAs)
      {
        const int ntc = A.ff / 8 / CS;
        cta_gemm(W.w1, D, a, lda, ntc, [&](int i) { return rank * ntc + i; }, part,
                 [&](int row, int col, float v0, float v1) {
                   const uint32_t pk = pack_bf16x2(gelu_erf(v0 + PB[1536 + col]), gelu_erf(v1 + PB[1536 + col + 1]));
#pragma unroll
                   for (int d = 0; d < CS; ++d) *reinterpret_cast<uint32_t*>(r_hbuf[d] + row * ldh + col) = pk;
                 });
      }
      DEC_TICK(13);
      csync();
      DEC_TICK(14);
      // ---- J: FFN second linear slice -> gath
      cta_gemm(W.w2, A.ff, hbuf, ldh, DC / 8, [&](int i) { return rank * (DC / 8) + i; }, part,
               [&](int row, int col, float v0, float v1) {
                 const float2 v = make_float2(v0 + PB[1536 + A.ff + col], v1 + PB[1536 + A.ff + col + 1]);
#pragma unroll
                 for (int d = 0; d < CS; ++d) *reinterpret_cast<float2*>(r_gath[d] + row * D + col) = v;
               });
      DEC_TICK(15);
      csync();
      DEC_TICK(16);
      // ---- K: x += gath; a = LN1 of the next layer (or dec_ln)
      {
        const float* ng = (l + 1 < A.layers) ? PN + lfl : p_decln;            // next layer's ln1, or dec_ln
        const float* nb = ng + 256;
        const int i = warp;
        float v[8];
        const float4 x0 = *reinterpret_cast<const float4*>(x + i * D + lane * 8), x1 = *reinterpret_cast<const float4*>(x + i * D + lane * 8 + 4);
        const float4 g0 = *reinterpret_cast<const float4*>(gath + i * D + lane * 8), g1 = *reinterpret_cast<const float4*>(gath + i * D + lane * 8 + 4);
        v[0] = x0.x + g0.x; v[1] = x0.y + g0.y; v[2] = x0.z + g0.z; v[3] = x0.w + g0.w;
        v[4] = x1.x + g1.x; v[5] = x1.y + g1.y; v[6] = x1.z + g1.z; v[7] = x1.w + g1.w;
        st_f32x8(x + i * D + lane * 8, v);
        ln8s(v, ng, nb, lane);
        st_bf16x8(a + i * lda + lane * 8, v);
      }
      __syncthreads();
      DEC_TICK(17);
    }
    // ---- heads: dec_head | lm_head slices -> logits (all CTAs)
    {
      const int nth = 2 * A.Vp / 8;
      const int t0 = rank * nth / CS, t1 = (rank + 1) * nth / CS;
      cta_gemm(A.wheads, D, a, lda, t1 - t0, [&](int i) { return t0 + i; }, part,
               [&](int row, int col, float v0, float v1) {
                 const float2 v = make_float2(v0 + p_heads[col], v1 + p_heads[col + 1]);
#pragma unroll
                 for (int d = 0; d < CS; ++d) *reinterpret_cast<float2*>(r_logits[d] + row * 2 * A.Vp + col) = v;
               });
    }
    DEC_TICK(18);
    csync();
    DEC_TICK(19);
    // ---- token selection, one warp per line, replicated in every CTA (model.py:480-537)
    {
      const int i = warp, b = st->line[i];
      if (!st->finished[i]) {
        const KiriDecodeParams& p = A.p;
        const float* dec = logits + i * 2 * A.Vp;
        const float* lm = dec + A.Vp;
        const int Vd = A.Vd;
        const float lse_d = warp_lse_s(dec, Vd, lane);
        const float lse_l = p.lm_alpha != 0.f ? warp_lse_s(lm, Vd, lane) : 0.f;
        const int n = st->n_tok[i];                       // ids so far, BOS included (== step + 1)
        int pid[8]; float pam[8]; int np = 0;
        const int cur_len = n - 1, tl = st->target[i];
        if (tl > 0) {
          int half = __double2int_rz(__dmul_rn(static_cast<double>(tl), 0.5));
          if (half < 1) half = 1;
          const int min_len = p.eos_bias_until_len < half ? p.eos_bias_until_len : half;
          if (cur_len < min_len) { pid[np] = kTokEOS; pam[np++] = p.eos_bias; }
          else if (cur_len >= tl) { pid[np] = kTokEOS; pam[np++] = -p.eos_boost; }
        } else if (cur_len < p.eos_bias_until_len) { pid[np] = kTokEOS; pam[np++] = p.eos_bias; }
        // hist[k] = k-th most recent id (k = 0 newest); entries beyond the sequence hold distinct negatives
        const int s1 = st->hist[i][0], s2 = st->hist[i][1], s3 = st->hist[i][2];
        const int s4 = st->hist[i][3], s5 = st->hist[i][4], s6 = st->hist[i][5];
        if (n >= 4 && s1 == s2 && s2 == s3) { pid[np] = s1; pam[np++] = p.rep_last; }
        if (n >= 4 && s2 == s4 && s1 == s3) {
          pid[np] = s1; pam[np++] = p.rep_bigram;
          pid[np] = s2; pam[np++] = p.rep_bigram;
        }
        if (n >= 4 && s1 == s3 && s2 == s4) { pid[np] = s1; pam[np++] = p.rep_bigram; }
        if (n >= 6 && s3 == s6 && s2 == s5 && s1 == s4) {
          pid[np] = s1; pam[np++] = p.rep_trigram;
          pid[np] = s2; pam[np++] = p.rep_trigram;
          pid[np] = s3; pam[np++] = p.rep_trigram;
        }
        auto fused = [&](int v) -> float {
          float lp = dec[v] - lse_d;
          if (p.lm_alpha != 0.f) lp += p.lm_alpha * (lm[v] - lse_l);
          for (int k = 0; k < np; ++k)
            if (pid[k] == v) lp -= pam[k];
          if (v == p.unk_id) lp -= p.unk_penalty;
          return lp;
        };
        if (!BM) {
        float best = -INFINITY;
        int bid = 0x7fffffff;
        for (int v = lane; v < Vd; v += 32) {
          const float val = p.select_raw ? dec[v] : fused(v);
          if (val > best) { best = val; bid = v; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bid, o);
          if (ob > best || (ob == best && oi < bid)) { best = ob; bid = oi; }
        }
        if (A.forced) bid = A.forced[static_cast<size_t>(b) * A.Lmax + step];
        __syncwarp();
        if (lane == 0) {
          const float lp = fused(bid);
#pragma unroll
          for (int k = 7; k > 0; --k) st->hist[i][k] = st->hist[i][k - 1];
          st->hist[i][0] = bid;
          st->n_tok[i] = n + 1;
          st->cur_tok[i] = bid;
          const bool done = (bid == kTokEOS) || (step + 1 >= st->max_steps[i]);
          if (done) st->finished[i] = 1;
          if (rank == 0) {
            A.ids[static_cast<size_t>(b) * A.Lmax + step] = bid;
            A.sum_logp[b] += lp;
            if (A.step_logp) A.step_logp[static_cast<size_t>(b) * A.Lmax + step] = lp;
            if (A.step_prob) A.step_prob[static_cast<size_t>(b) * A.Lmax + step] = expf(dec[bid] - lse_d);
            if (A.publish) __threadfence_system();                 // the step's records first, then the counters
            A.n_out[b] = step + 1;
            if (A.progress) A.progress[b] = (step + 1) | (done ? kDoneBit : 0);
          }
        }
        } else {
          // top-BEAM of the penalised log-probs (torch.topk, model.py:537): BEAM warp arg-max rounds
          int chosen[8];
          for (int k = 0; k < BEAM; ++k) {
            float best = -INFINITY;
            int bid = 0x7fffffff;
            for (int v = lane; v < Vd; v += 32) {
              bool taken = false;
              for (int t = 0; t < k; ++t) taken |= (chosen[t] == v);
              if (!taken) {
                const float val = fused(v);
                if (val > best) { best = val; bid = v; }
              }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float ob = __shfl_xor_sync(0xffffffffu, best, o);
              const int oi = __shfl_xor_sync(0xffffffffu, bid, o);
              if (ob > best || (ob == best && oi < bid)) { best = ob; bid = oi; }
            }
            chosen[k] = bid;
            if (lane == 0) { bs->cand_v[i][k] = best; bs->cand_i[i][k] = bid; }
          }
        }
      }
    }
    if (BM) {
      __syncthreads();
      // ---- beam bookkeeping, one warp per line (model.py:539-559): candidates = finished hypotheses
      // (in order) then every alive hypothesis' top-BEAM (in order); stable sort by the length-normalised
      // score, keep BEAM.  Lane = candidate.
      const int li = warp;
      if (li < LPC && !bs->line_done[li]) {
        const int s0 = li * BEAM;
        int n_done = 0, n_alive = 0, done_slot[8], alive_slot[8];
        for (int r = 0; r < BEAM; ++r) {
          const int stt = bs->state[s0 + r];
          if (stt == 2) done_slot[n_done++] = s0 + r;
          else if (stt == 1) alive_slot[n_alive++] = s0 + r;
        }
        const int n_cand = n_done + n_alive * BEAM;             // <= 30 for BEAM <= 5
        const bool has = lane < n_cand;
        int src = s0, tok = -1, ntok_new = 1;
        float cv = 0.f;
        double sc = 0.0;
        bool fin = true;
        if (has) {
          if (lane < n_done) {
            for (int r = 0; r < n_done; ++r) if (r == lane) src = done_slot[r];
            sc = bs->score[src]; ntok_new = st->n_tok[src];
          } else {
            const int a = (lane - n_done) / BEAM, k = (lane - n_done) - a * BEAM;
            for (int r = 0; r < n_alive; ++r) if (r == a) src = alive_slot[r];
            tok = bs->cand_i[src][k]; cv = bs->cand_v[src][k];
            sc = bs->score[src] + static_cast<double>(cv);        // Python: base_score + float(v)
            ntok_new = st->n_tok[src] + 1;
            fin = (tok == kTokEOS);
          }
        }
        const int Ln = ntok_new - 1 > 1 ? ntok_new - 1 : 1;
        const double pen = A.stream_rule ? pow(static_cast<double>(Ln), A.lenp)                      // model.py:1112-1115
                                         : pow(5.0 + static_cast<double>(Ln), A.lenp) / pow(6.0, A.lenp);   // model.py:550-555
        const double normed = has ? sc / pen : -1e300;
        int rk = 0;
        for (int j = 0; j < n_cand; ++j) {
          const double oj = __shfl_sync(0xffffffffu, normed, j);
          if (j != lane && (oj > normed || (oj == normed && j < lane))) ++rk;
        }
        const int nb = n_cand < BEAM ? n_cand : BEAM;
        const bool win = has && rk < nb;
        // read everything the new hypothesis inherits BEFORE any slot is overwritten
        int h_old[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) h_old[k] = st->hist[src][k];
        const int cur_old = st->cur_tok[src], len_old = st->n_tok[src] - 1;
        __syncwarp();
        if (win) {
          const int ns = s0 + rk;
          bs->score[ns] = sc;
          bs->state[ns] = fin ? 2 : 1;
          st->finished[ns] = fin ? 1 : 0;
          st->n_tok[ns] = ntok_new;
          if (tok >= 0) {
            st->cur_tok[ns] = tok;
            st->hist[ns][0] = tok;
#pragma unroll
            for (int k = 1; k < 8; ++k) st->hist[ns][k] = h_old[k - 1];
          } else {
            st->cur_tok[ns] = cur_old;
#pragma unroll
            for (int k = 0; k < 8; ++k) st->hist[ns][k] = h_old[k];
          }
          bs->src[ns] = src; bs->new_tok[ns] = tok; bs->new_v[ns] = cv; bs->old_len[ns] = len_old;
        }
        if (lane >= nb && lane < BEAM) { bs->state[s0 + lane] = 0; st->finished[s0 + lane] = 1; }
        __syncwarp();
        // inherit the ancestors' K/V slots and (rank 0) the hypothesis records, ping-pong buffers
        bool all_done = true;
        for (int r = 0; r < nb; ++r) {
          const int ns = s0 + r, sr = bs->src[ns], ol = bs->old_len[ns], tk = bs->new_tok[ns];
          const uint8_t* a_src = anc + (pp * kFL + sr) * anc_ld;
          uint8_t* a_dst = anc + ((pp ^ 1) * kFL + ns) * anc_ld;
          for (int t = lane; t <= step; t += 32) a_dst[t] = a_src[t];
          if (rank == 0) {
            const size_t nsl = static_cast<size_t>(gridDim.x / CS) * kFL;      // physical slots of the launch
            const int* q_src = A.seqbuf + (static_cast<size_t>(pp) * nsl + b0 + sr) * A.Lmax;
            int* q_dst = A.seqbuf + (static_cast<size_t>(pp ^ 1) * nsl + b0 + ns) * A.Lmax;
            const float* l_src = A.lpbuf + (static_cast<size_t>(pp) * nsl + b0 + sr) * A.Lmax;
            float* l_dst = A.lpbuf + (static_cast<size_t>(pp ^ 1) * nsl + b0 + ns) * A.Lmax;
            for (int t = lane; t < ol; t += 32) { q_dst[t] = q_src[t]; l_dst[t] = l_src[t]; }
            if (lane == 0 && tk >= 0) { q_dst[ol] = tk; l_dst[ol] = bs->new_v[ns]; }
          }
          all_done &= (bs->state[ns] == 2);
          if (rank == 0 && A.bm_trace && lane == 0) {
            int* tr = A.bm_trace + ((static_cast<size_t>(st->line[s0]) * A.Lmax + step) * BEAM + r) * 3;
            tr[0] = sr - s0; tr[1] = tk; tr[2] = __float_as_int(bs->new_v[ns]);
          }
        }
        if (rank == 0 && A.bm_trace && lane == 0)
          for (int r = nb; r < BEAM; ++r) A.bm_trace[((static_cast<size_t>(st->line[s0]) * A.Lmax + step) * BEAM + r) * 3] = -1;
        // batch beam search runs until every hypothesis ended (model.py:444-452); the streaming form stops as soon as
        // the BEST one ended (model.py:1148-1150)
        const bool best_done = A.stream_rule && nb > 0 && bs->state[s0] == 2;
        const bool ldone = all_done || best_done || (step + 1 >= st->max_steps[s0]);
        if (rank == 0 && A.progress && lane == 0) {
          if (A.publish) __threadfence_system();
          A.progress[st->line[s0]] = (step + 1) | (ldone ? kDoneBit : 0);
        }
        __syncwarp();
        if (ldone) {
          if (lane == 0) bs->line_done[li] = 1;
          for (int r = 0; r < BEAM; ++r) st->finished[s0 + r] = 1;
          if (rank == 0) {
            // the line is complete: publish its hypotheses (score, length, ids, log-probs)
            const int b = st->line[s0];
            const size_t nsl = static_cast<size_t>(gridDim.x / CS) * kFL;
            for (int r = 0; r < BEAM; ++r) {
              const int ns = s0 + r;
              const int len = r < nb ? st->n_tok[ns] - 1 : 0;
              if (lane == 0) {
                A.bm_state[b * BEAM + r] = r < nb ? bs->state[ns] : 0;
                A.bm_score[b * BEAM + r] = r < nb ? bs->score[ns] : 0.0;
                A.bm_len[b * BEAM + r] = len;
              }
              const int* q_src = A.seqbuf + (static_cast<size_t>(pp ^ 1) * nsl + b0 + ns) * A.Lmax;
              const float* l_src = A.lpbuf + (static_cast<size_t>(pp ^ 1) * nsl + b0 + ns) * A.Lmax;
              for (int t = lane; t < len; t += 32) {
                A.bm_ids[(static_cast<size_t>(b) * BEAM + r) * A.Lmax + t] = q_src[t];
                A.bm_logp[(static_cast<size_t>(b) * BEAM + r) * A.Lmax + t] = l_src[t];
              }
            }
          }
        }
      }
      pp ^= 1;
    }
    // the next heads-GEMM writes `logits` remotely only after 18 more cluster barriers, and
    // `st` is CTA-local: a block barrier is enough here
    __syncthreads();
    DEC_TICK(20);
  }
  if (rank == 0 && threadIdx.x == 0 && A.steps_max) atomicMax(A.steps_max, step);
  if (A.timing && rank == 0 && threadIdx.x == 0 && cid < 64) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    g_dec_clu[cid * 4 + 1] = static_cast<long long>(ns);
    g_dec_clu[cid * 4 + 2] = step;
  }
  csync();                                          // no CTA may exit while peers can still write its smem
}

// ---------------------------------------------------------------- cross K/V relayout
// src: the cross-K/V GEMM output [M, ld = layers*2*256] (token-major).  dst: per line the block
// [layer][K|V][head][t][32]; one CTA moves one token row (ld/8 16-byte chunks).
__global__ void __launch_bounds__(256)
crosskv_headmajor_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int ld, const int* __restrict__ mem_row0,
                         const int* __restrict__ mem_len, int T_uniform) {
  const int b = blockIdx.y, t = blockIdx.x;
  const int T = mem_len ? mem_len[b] : T_uniform;
  if (t >= T) return;
  const size_t row0 = mem_row0 ? static_cast<size_t>(mem_row0[b]) : static_cast<size_t>(b) * T_uniform;
  const int chunks = ld / 8;                         // 16-byte chunks per row
  const uint4* s_row = src + (row0 + t) * chunks;
  uint4* d_line = dst + row0 * chunks;
  for (int c = threadIdx.x; c < chunks; c += blockDim.x) {
    const int col = c * 8, sec = col >> 8, head = (col & 255) >> 5, d0 = col & 31;
    d_line[((static_cast<size_t>(sec) * kHeads + head) * T + t) * 4 + (d0 >> 3)] = __ldg(s_row + c);
  }
}

int crosskv_headmajor(const __nv_bfloat16* src, __nv_bfloat16* dst, int ld, const int* mem_row0, const int* mem_len,
                      int T_uniform, int max_T, int n_lines, cudaStream_t stream) {
  KIRI_REQUIRE(ld % 256 == 0, "crosskv_headmajor: row length %d must be a multiple of 256", ld);
  if (n_lines == 0) return 0;
  crosskv_headmajor_kernel<<<dim3(max_T, n_lines), 192, 0, stream>>>(reinterpret_cast<const uint4*>(src),
                                                                    reinterpret_cast<uint4*>(dst), ld, mem_row0, mem_len, T_uniform);
  KIRI_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- host side
struct FusedPacked {
  void* blob = nullptr;
  FusedArgs args;
};

static size_t packed_words(int N, int K) { return static_cast<size_t>(N) * K / 2; }

int fused_decoder_build(KiriHandle* h) {
  const KiriDims& d = h->d;
  const KiriWeights& w = h->w;
  const int D = d.dec_dim, FF = d.dec_ff, L = d.dec_layers;
  const int Vp = (d.dec_vocab + 15) / 16 * 16;
  KIRI_REQUIRE(D == 256 && d.dec_heads == kHeads, "fused decoder: DEC_DIM must be 256 with 8 heads");
  KIRI_REQUIRE(FF % 256 == 0 && FF <= 2048, "fused decoder: DEC_FF=%d must be a multiple of 256 (<= 2048)", FF);
  {
    // the decode kernel keeps the logits of 16 lines and every bias / LayerNorm affine in shared memory: say so at load
    // time when a vocabulary does not fit, instead of failing at the first "accurate" call (greedy plan, cluster of 8)
    int dev = 0, optin = 0;
    KIRI_CHECK_CUDA(cudaGetDevice(&dev));
    KIRI_CHECK_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int need = fused_smem_plan(FF, Vp, 8, L, 0, 0).total;
    KIRI_REQUIRE(need <= optin, "fused decoder: decoder vocabulary %d (FF %d, %d layers) needs %d bytes of shared memory per CTA, "
                 "%d available: the attention decoder of this checkpoint cannot run on this build", d.dec_vocab, FF, L, need, optin);
  }
  size_t words = 0;
  const size_t per_layer = packed_words(3 * D, D) + 3 * packed_words(D, D) + 2 * packed_words(FF, D);
  words = per_layer * L + packed_words(2 * Vp, D);
  FusedPacked* fp = new FusedPacked;
  KIRI_CHECK_CUDA(cudaMalloc(&fp->blob, words * 4));
  uint32_t* cur = reinterpret_cast<uint32_t*>(fp->blob);
  auto pack = [&](const void* src, int N, int K) -> const uint4* {
    uint32_t* dst = cur;
    cur += packed_words(N, K);
    pack_frag_kernel<<<148, 256>>>(reinterpret_cast<const __nv_bfloat16*>(src), N, K, dst);
    return reinterpret_cast<const uint4*>(dst);
  };
  FusedArgs& a = fp->args;
  memset(&a, 0, sizeof(a));
  for (int l = 0; l < L; ++l) {
    const KiriDecLayerWeights& s = w.dec[l];
    FusedLayer& t = a.layer[l];
    t.wqkv = pack(s.wqkv, 3 * D, D); t.wo = pack(s.wo, D, D); t.wcq = pack(s.wcq, D, D);
    t.wco = pack(s.wco, D, D); t.w1 = pack(s.w1, FF, D); t.w2 = pack(s.w2, D, FF);
    t.bqkv = s.bqkv; t.bo = s.bo; t.bcq = s.bcq; t.bco = s.bco; t.b1 = s.b1; t.b2 = s.b2;
    t.ln1_g = s.ln1_g; t.ln1_b = s.ln1_b; t.ln2_g = s.ln2_g; t.ln2_b = s.ln2_b; t.ln3_g = s.ln3_g; t.ln3_b = s.ln3_b;
  }
  a.wheads = pack(w.heads_w, 2 * Vp, D);
  a.bheads = w.heads_b;
  a.dec_ln_g = w.dec_ln_g; a.dec_ln_b = w.dec_ln_b; a.emb = w.dec_emb; a.pe = w.dec_pe;
  a.layers = L; a.ff = FF; a.Vd = d.dec_vocab; a.Vp = Vp; a.has_pos = d.has_dec_pos;
  KIRI_CHECK_CUDA(cudaGetLastError());
  KIRI_CHECK_CUDA(cudaDeviceSynchronize());
  h->fused = fp;
  return 0;
}

void fused_decoder_free(KiriHandle* h) {
  FusedPacked* fp = reinterpret_cast<FusedPacked*>(h->fused);
  if (!fp) return;
  cudaFree(fp->blob);
  delete fp;
  h->fused = nullptr;
}

template <int CS>
static int launch_fused(const FusedArgs& a, int n_clusters, cudaStream_t stream) {
  const FusedSmem L = fused_smem_plan(a.ff, a.Vp, CS, a.layers, a.bmode ? a.beam : 0, a.Lmax);
  auto kern = dec_fused_kernel<CS>;
  static int configured[kMaxDevices] = {0};           // per device and cluster size
  const int dslot = kiri_cur_device_slot();
  if (configured[dslot] < L.total) {
    KIRI_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    configured[dslot] = L.total;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(n_clusters * CS);
  cfg.blockDim = dim3(kFThreads);
  cfg.dynamicSmemBytes = L.total;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  KIRI_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  return 0;
}

// Runs the whole greedy decode; crosskv must already hold the cross K/V rows.
int fused_decoder_run(KiriHandle* h, const __nv_bfloat16* crosskv, int crosskv_ld, const int* mem_row0, const int* mem_len,
                      int T, __nv_bfloat16* self_k, __nv_bfloat16* self_v, const int* len_est, const int* forced,
                      const int* line_perm, int B, int Lmax, const KiriDecodeParams* p, int* ids, int* n_out,
                      float* sum_logp, float* step_logp, float* step_prob, int* steps_max_dev, int cluster_size,
                      cudaStream_t stream, const FusedBeam* beam, int kv_headmajor, const FusedLive* live) {
  FusedPacked* fp = reinterpret_cast<FusedPacked*>(h->fused);
  KIRI_REQUIRE(fp, "fused decoder: handle was created without decoder weights");
  FusedArgs a = fp->args;
  a.crosskv = crosskv; a.crosskv_ld = crosskv_ld; a.kv_hm = kv_headmajor; a.mem_row0 = mem_row0; a.mem_len = mem_len; a.T = T;
  a.self_k = self_k; a.self_v = self_v; a.len_est = len_est; a.forced = forced; a.line_perm = line_perm; a.B = B; a.Lmax = Lmax; a.p = *p;
  a.ids = ids; a.n_out = n_out; a.sum_logp = sum_logp; a.step_logp = step_logp; a.step_prob = step_prob;
  a.steps_max = steps_max_dev;
  a.timing = getenv("KIRI_DEC_TIMING") != nullptr;
  a.prefetch = getenv("KIRI_DEC_PREFETCH") != nullptr;       // measured: no net gain (profiles/README.md), off by default
  a.beam = 1; a.bmode = 0; a.lenp = 0.0;
  a.publish = 0; a.progress = nullptr; a.stream_rule = 0; a.bm_trace = nullptr;
  if (live) { a.publish = live->publish; a.progress = live->progress; a.stream_rule = live->stream_rule; a.bm_trace = live->bm_trace; }
  if (beam) {
    a.bmode = 1;
    KIRI_REQUIRE(beam->beam >= 1 && beam->beam <= 5, "fused decoder: beam width %d not in 1..5", beam->beam);
    KIRI_REQUIRE(cluster_size > 1 || Lmax <= 160, "fused decoder: beam search needs a cluster size > 1 for Lmax=%d", Lmax);
    a.beam = beam->beam; a.lenp = beam->lenp; a.seqbuf = beam->seqbuf; a.lpbuf = beam->lpbuf;
    a.bm_score = beam->score; a.bm_len = beam->len; a.bm_state = beam->state; a.bm_ids = beam->ids; a.bm_logp = beam->logp;
  }
  const int lpc = kFL / a.beam;
  const int n_clusters = (B + lpc - 1) / lpc;
  switch (cluster_size) {
    case 1: return launch_fused<1>(a, n_clusters, stream);
    case 2: return launch_fused<2>(a, n_clusters, stream);
    case 4: return launch_fused<4>(a, n_clusters, stream);
    case 8: return launch_fused<8>(a, n_clusters, stream);
    default: KIRI_REQUIRE(false, "fused decoder: cluster size %d not in {1,2,4,8}", cluster_size);
  }
}

}  // namespace kiri

// Debug: cycles per decode phase of cluster 0 / CTA 0 accumulated since the last call (KIRI_DEC_TIMING=1).
extern "C" int kiri_debug_decode_timing(long long* out_host, int n) {
  long long buf[32];
  if (cudaDeviceSynchronize() != cudaSuccess) return -2;
  if (cudaMemcpyFromSymbol(buf, kiri::g_dec_prof, sizeof(buf)) != cudaSuccess) return -2;
  for (int i = 0; i < n && i < 32; ++i) out_host[i] = buf[i];
  long long zero[32] = {0};
  if (cudaMemcpyToSymbol(kiri::g_dec_prof, zero, sizeof(zero)) != cudaSuccess) return -2;
  return 0;
}
extern "C" int kiri_debug_decode_clusters(long long* out_host, int n) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -2;
  if (n > 256) n = 256;
  if (cudaMemcpyFromSymbol(out_host, kiri::g_dec_clu, sizeof(long long) * n) != cudaSuccess) return -2;
  return 0;
}
