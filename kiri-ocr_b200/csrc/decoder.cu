// K11-K12: batched greedy attention decoder with KV cache.
//
// Replaces  beam_decode_one_batched at BEAM=1   kiri_ocr/model.py:390-600 (core.py:560-568)
//           greedy_decode_streaming (token rule) kiri_ocr/model.py:779-946
// The reference re-runs the whole decoder over the full prefix for every new token, re-projects
// the memory into K/V at every step and syncs to the host per step.  Here:
//   * cross-attention K/V for all layers come from ONE GEMM over the encoder memory
//     ((W_k;W_v)_l @ W_memproj is folded on the host), once per batch;
//   * every step feeds ONE token per line: embedding + position + LN, then per layer
//     qkv GEMM -> cache append + single-query self-attention -> out-proj GEMM (+residual+LN) ->
//     cross-q GEMM -> single-query cross-attention -> out GEMM (+residual+LN) -> FFN GEMMs;
//   * both heads are one GEMM; log-softmax, LM fusion, the four cumulative repeat penalties,
//     UNK/EOS adjustments, arg-max and the stop rule run in one warp per line on the device.
// All B lines advance in lock-step; finished lines are frozen.  M = B is small, so the step is
// latency-bound (weights stay L2-resident); see DESIGN.md.
#include <cstdlib>

#include "internal.cuh"
#include "ln_utils.cuh"

namespace kiri {

static constexpr int kHdDec = 32;
static constexpr int kBOS = 1, kEOS = 2;

struct DecodeWs {
  size_t crosskv, self_k, self_v, x, a, qkv, o, qc, h, logits, seq, n_tok, finished, max_steps, target, alive, total;
};
static inline size_t al256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

static DecodeWs plan_decode(const KiriDims& d, int B, int T, int Lmax) {
  DecodeWs w;
  size_t off = 0;
  const size_t L = d.dec_layers, D = d.dec_dim;
  const size_t Vp = (d.dec_vocab + 15) / 16 * 16;
  w.crosskv = off; off += al256(static_cast<size_t>(B) * T * L * 2 * D * 2);
  w.self_k = off;  off += al256(L * B * Lmax * D * 2);
  w.self_v = off;  off += al256(L * B * Lmax * D * 2);
  w.x = off;       off += al256(static_cast<size_t>(B) * D * 4);
  w.a = off;       off += al256(static_cast<size_t>(B) * D * 2);
  w.qkv = off;     off += al256(static_cast<size_t>(B) * 3 * D * 2);
  w.o = off;       off += al256(static_cast<size_t>(B) * D * 2);
  w.qc = off;      off += al256(static_cast<size_t>(B) * D * 2);
  w.h = off;       off += al256(static_cast<size_t>(B) * d.dec_ff * 2);
  w.logits = off;  off += al256(static_cast<size_t>(B) * 2 * Vp * 4);
  w.seq = off;     off += al256(static_cast<size_t>(B) * (Lmax + 1) * 4);
  w.n_tok = off;   off += al256(static_cast<size_t>(B) * 4);
  w.finished = off; off += al256(static_cast<size_t>(B) * 4);
  w.max_steps = off; off += al256(static_cast<size_t>(B) * 4);
  w.target = off;  off += al256(static_cast<size_t>(B) * 4);
  w.alive = off;   off += 256;
  w.total = off;
  return w;
}

// ------------------------------------------------------------------ init
__global__ void dec_init_kernel(const int* __restrict__ len_est, int B, int T, int Lmax, KiriDecodeParams p,
                                int* seq, int* n_tok, int* finished, int* max_steps, int* target, int* alive,
                                int* n_out, float* sum_logp) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) *alive = B;
  if (b >= B) return;
  const int tl = len_est[b];
  int ms;
  // model.py:416-425 — Python float (double) arithmetic, int() truncation
  if (tl > 0) ms = __double2int_rz(__dmul_rn(static_cast<double>(tl), p.len_ratio)) + p.len_pad;
  else ms = __double2int_rz(__dmul_rn(static_cast<double>(T), p.mem_ratio)) + p.len_pad;
  if (ms > p.max_dec_len) ms = p.max_dec_len;
  if (ms > Lmax) ms = Lmax;
  seq[static_cast<size_t>(b) * (Lmax + 1)] = kBOS;
  n_tok[b] = 1;
  finished[b] = ms <= 0 ? 1 : 0;
  if (ms <= 0) atomicSub(alive, 1);
  max_steps[b] = ms;
  target[b] = tl;
  n_out[b] = 0;
  sum_logp[b] = 0.f;
}

// ------------------------------------------------------------------ embedding + position + LN
__global__ void __launch_bounds__(256)
dec_embed_ln_kernel(const int* __restrict__ seq, const int* __restrict__ finished, int B, int Lmax,
                    const int* __restrict__ step_dev, const float* __restrict__ emb,
                    const float* __restrict__ pe, int has_pos, const float* g, const float* bta,
                    float* __restrict__ x, __nv_bfloat16* __restrict__ a) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;
  const int step = *step_dev;
  // finished lines are frozen: their seq row is not extended, so feed the pad token
  const int tokid = finished[b] ? 0 : seq[static_cast<size_t>(b) * (Lmax + 1) + step];
  const float4* e = reinterpret_cast<const float4*>(emb + static_cast<size_t>(tokid) * kD) + lane * 2;
  float4 e0 = __ldg(e), e1 = __ldg(e + 1);
  float v[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
  if (has_pos) {
    const float4* pp = reinterpret_cast<const float4*>(pe + static_cast<size_t>(step) * kD) + lane * 2;
    const float4 p0 = __ldg(pp), p1 = __ldg(pp + 1);
    v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w;
    v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
  }
  st_f32x8(x + static_cast<size_t>(b) * kD + lane * 8, v);
  ln8(v, g, bta, lane);
  st_bf16x8(a + static_cast<size_t>(b) * kD + lane * 8, v);
}

// ------------------------------------------------------------------ single-query attention
// One warp per (line, head).  APPEND: first store this step's K/V rows into the cache.
// q: bf16 [B, q_ld] (+head*32); K/V row r of line b at kbase + (b*kv_rows + r)*kv_ld + head*32.
template <bool APPEND, int MAXCH>
__global__ void __launch_bounds__(128)
dec_attention_kernel(const __nv_bfloat16* __restrict__ q, int q_ld, const __nv_bfloat16* __restrict__ knew,
                     const __nv_bfloat16* __restrict__ vnew, __nv_bfloat16* __restrict__ kbase,
                     __nv_bfloat16* __restrict__ vbase, int kv_rows, int kv_ld, int n_keys_static,
                     const int* __restrict__ step_dev, const int* kv_len, int B, int heads,
                     __nv_bfloat16* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (wid >= B * heads) return;
  const int b = wid / heads, head = wid - b * heads;
  __nv_bfloat16* kb = kbase + static_cast<size_t>(b) * kv_rows * kv_ld + head * kHdDec;
  __nv_bfloat16* vb = vbase + static_cast<size_t>(b) * kv_rows * kv_ld + head * kHdDec;
  const int n_keys = APPEND ? (*step_dev + 1) : n_keys_static;   // self-attention: keys 0..step
  int len = n_keys;
  if (kv_len) len = min(len, kv_len[b]);
  if (APPEND) {
    const int pos = n_keys - 1;
    kb[static_cast<size_t>(pos) * kv_ld + lane] = knew[static_cast<size_t>(b) * q_ld + head * kHdDec + lane];
    vb[static_cast<size_t>(pos) * kv_ld + lane] = vnew[static_cast<size_t>(b) * q_ld + head * kHdDec + lane];
    __syncwarp();
  }
  // q in registers (all lanes hold the whole 32-vector)
  float qf[32];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(q + static_cast<size_t>(b) * q_ld + head * kHdDec);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 u = qp[i];
      qf[8 * i + 0] = bf16_lo(u.x); qf[8 * i + 1] = bf16_hi(u.x); qf[8 * i + 2] = bf16_lo(u.y); qf[8 * i + 3] = bf16_hi(u.y);
      qf[8 * i + 4] = bf16_lo(u.z); qf[8 * i + 5] = bf16_hi(u.z); qf[8 * i + 6] = bf16_lo(u.w); qf[8 * i + 7] = bf16_hi(u.w);
    }
  }
  const float scale = 0.17677669529663687f;           // 1/sqrt(32)
  float s[MAXCH];
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < MAXCH; ++c) {
    const int j = c * 32 + lane;
    float acc = -INFINITY;
    if (j < len) {
      const uint4* kp = reinterpret_cast<const uint4*>(kb + static_cast<size_t>(j) * kv_ld);
      acc = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 u = kp[i];
        acc = fmaf(qf[8 * i + 0], bf16_lo(u.x), acc); acc = fmaf(qf[8 * i + 1], bf16_hi(u.x), acc);
        acc = fmaf(qf[8 * i + 2], bf16_lo(u.y), acc); acc = fmaf(qf[8 * i + 3], bf16_hi(u.y), acc);
        acc = fmaf(qf[8 * i + 4], bf16_lo(u.z), acc); acc = fmaf(qf[8 * i + 5], bf16_hi(u.z), acc);
        acc = fmaf(qf[8 * i + 6], bf16_lo(u.w), acc); acc = fmaf(qf[8 * i + 7], bf16_hi(u.w), acc);
      }
      acc *= scale;
    }
    s[c] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < MAXCH; ++c) {
    s[c] = (c * 32 + lane < len) ? __expf(s[c] - mx) : 0.f;
    sum += s[c];
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  // o[lane] = sum_j p_j V[j][lane]
  float o = 0.f;
#pragma unroll
  for (int c = 0; c < MAXCH; ++c) {
    const int n_here = min(32, len - c * 32);
    for (int jj = 0; jj < n_here; ++jj) {
      const float pj = __shfl_sync(0xffffffffu, s[c], jj);
      o = fmaf(pj, __bfloat162float(vb[static_cast<size_t>(c * 32 + jj) * kv_ld + lane]), o);
    }
  }
  out[static_cast<size_t>(b) * (heads * kHdDec) + head * kHdDec + lane] = __float2bfloat16_rn(o * inv);
}

// ------------------------------------------------------------------ token selection
__device__ __forceinline__ float warp_lse(const float* __restrict__ x, int n, int lane) {
  float m = -INFINITY;
  for (int v = lane; v < n; v += 32) m = fmaxf(m, x[v]);
  m = warp_max(m);
  float s = 0.f;
  for (int v = lane; v < n; v += 32) s += expf(x[v] - m);
  s = warp_sum(s);
  return m + logf(s);
}

__global__ void __launch_bounds__(128)
dec_select_kernel(const float* __restrict__ logits, int B, int Vd, int Vp, int Lmax,
                  const int* __restrict__ step_dev, KiriDecodeParams p,
                  int* seq, int* n_tok, int* finished, const int* __restrict__ max_steps,
                  const int* __restrict__ target, int* alive, const int* __restrict__ forced, int* ids_out,
                  int* n_out, float* sum_logp, float* step_logp, float* step_prob) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= B) return;
  if (finished[b]) return;
  const int step = *step_dev;
  const float* dec = logits + static_cast<size_t>(b) * 2 * Vp;
  const float* lm = dec + Vp;
  const float lse_d = warp_lse(dec, Vd, lane);
  const float lse_l = p.lm_alpha != 0.f ? warp_lse(lm, Vd, lane) : 0.f;
  int* sq = seq + static_cast<size_t>(b) * (Lmax + 1);
  const int n = n_tok[b];                       // ids so far, BOS included (== step + 1)
  // up to 8 (id, amount) adjustments, applied cumulatively (model.py:490-534)
  int pid[8];
  float pam[8];
  int np = 0;
  const int cur_len = n - 1;
  const int tl = target[b];
  if (tl > 0) {
    int half = __double2int_rz(__dmul_rn(static_cast<double>(tl), 0.5));
    if (half < 1) half = 1;
    const int min_len = p.eos_bias_until_len < half ? p.eos_bias_until_len : half;
    if (cur_len < min_len) { pid[np] = kEOS; pam[np++] = p.eos_bias; }
    else if (cur_len >= tl) { pid[np] = kEOS; pam[np++] = -p.eos_boost; }
  } else if (cur_len < p.eos_bias_until_len) { pid[np] = kEOS; pam[np++] = p.eos_bias; }
  const int s1 = n >= 1 ? sq[n - 1] : -1, s2 = n >= 2 ? sq[n - 2] : -2, s3 = n >= 3 ? sq[n - 3] : -3;
  const int s4 = n >= 4 ? sq[n - 4] : -4, s5 = n >= 5 ? sq[n - 5] : -5, s6 = n >= 6 ? sq[n - 6] : -6;
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
  // at most 1 + 1 + 2 + 1 + 3 = 8 entries; UNK handled separately
  auto fused = [&](int v) -> float {
    float lp = dec[v] - lse_d;
    if (p.lm_alpha != 0.f) lp += p.lm_alpha * (lm[v] - lse_l);
    for (int i = 0; i < np; ++i)
      if (pid[i] == v) lp -= pam[i];
    if (v == p.unk_id) lp -= p.unk_penalty;
    return lp;
  };
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
  if (forced) bid = forced[static_cast<size_t>(b) * Lmax + step];
  if (lane == 0) {
    const float lp = fused(bid);
    sq[n] = bid;
    n_tok[b] = n + 1;
    ids_out[static_cast<size_t>(b) * Lmax + step] = bid;
    n_out[b] = step + 1;
    sum_logp[b] += lp;
    if (step_logp) step_logp[static_cast<size_t>(b) * Lmax + step] = lp;
    if (step_prob) step_prob[static_cast<size_t>(b) * Lmax + step] = expf(dec[bid] - lse_d);
    if (bid == kEOS || step + 1 >= max_steps[b]) {
      finished[b] = 1;
      atomicSub(alive, 1);
    }
  }
}

__global__ void dec_advance_kernel(int* step_dev) { *step_dev += 1; }

// MAXCH buckets of the self-attention kernel (keys = 32 * MAXCH)
static const int kAttBuckets[6] = {1, 2, 3, 5, 8, 17};
static int att_bucket(int n_keys) {
  const int ch = (n_keys + 31) / 32;
  for (int i = 0; i < 6; ++i)
    if (ch <= kAttBuckets[i]) return i;
  return -1;
}

template <bool APPEND>
static int launch_attention(const __nv_bfloat16* q, int q_ld, const __nv_bfloat16* knew, const __nv_bfloat16* vnew,
                            __nv_bfloat16* kb, __nv_bfloat16* vb, int kv_rows, int kv_ld, int n_keys, int bucket,
                            const int* step_dev, const int* kv_len, int B, int heads, __nv_bfloat16* out,
                            cudaStream_t stream) {
  const int grid = (B * heads + 3) / 4;
#define KIRI_ATT(N) dec_attention_kernel<APPEND, N><<<grid, 128, 0, stream>>>(q, q_ld, knew, vnew, kb, vb, kv_rows, kv_ld, n_keys, step_dev, kv_len, B, heads, out)
  switch (bucket) {
    case 0: KIRI_ATT(1); break;
    case 1: KIRI_ATT(2); break;
    case 2: KIRI_ATT(3); break;
    case 3: KIRI_ATT(5); break;
    case 4: KIRI_ATT(8); break;
    case 5: KIRI_ATT(17); break;
    default: KIRI_REQUIRE(false, "decoder attention over %d keys unsupported (max 544)", n_keys);
  }
#undef KIRI_ATT
  KIRI_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace kiri

using namespace kiri;

extern "C" size_t kiri_decode_workspace_bytes(const KiriHandle* h, int B, int T, int Lmax) {
  if (!h || B <= 0 || T <= 0 || Lmax <= 0) return 0;
  return plan_decode(h->d, B, T, Lmax).total;
}

extern "C" int kiri_decode_greedy(KiriHandle* h, const void* mem_bf16, const int* len_est, int B, int T, int Lmax,
                                  const KiriDecodeParams* p, void* workspace, size_t workspace_bytes, int* ids,
                                  int* n_out, float* sum_logp, float* step_logp, float* step_prob,
                                  const int* forced_ids, int* steps_run_host, int poll_every,
                                  cudaStream_t stream) {
  KIRI_REQUIRE(h && mem_bf16 && len_est && p && workspace && ids && n_out && sum_logp, "kiri_decode_greedy: null pointer");
  KIRI_REQUIRE(B > 0 && T > 0 && Lmax > 0 && Lmax <= h->d.max_pos && Lmax <= 544, "kiri_decode_greedy: bad sizes B=%d T=%d Lmax=%d", B, T, Lmax);
  const KiriDims& d = h->d;
  const KiriWeights& w = h->w;
  const DecodeWs ws = plan_decode(d, B, T, Lmax);
  KIRI_REQUIRE(workspace_bytes >= ws.total, "kiri_decode_greedy: workspace of %zu bytes given, %zu needed", workspace_bytes, ws.total);
  uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
  const int D = d.dec_dim, L = d.dec_layers, heads = d.dec_heads;
  const int Vd = d.dec_vocab, Vp = (Vd + 15) / 16 * 16;
  __nv_bfloat16* crosskv = reinterpret_cast<__nv_bfloat16*>(base + ws.crosskv);
  __nv_bfloat16* self_k = reinterpret_cast<__nv_bfloat16*>(base + ws.self_k);
  __nv_bfloat16* self_v = reinterpret_cast<__nv_bfloat16*>(base + ws.self_v);
  float* x = reinterpret_cast<float*>(base + ws.x);
  __nv_bfloat16* a = reinterpret_cast<__nv_bfloat16*>(base + ws.a);
  __nv_bfloat16* qkv = reinterpret_cast<__nv_bfloat16*>(base + ws.qkv);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(base + ws.o);
  __nv_bfloat16* qc = reinterpret_cast<__nv_bfloat16*>(base + ws.qc);
  __nv_bfloat16* hb = reinterpret_cast<__nv_bfloat16*>(base + ws.h);
  float* logits = reinterpret_cast<float*>(base + ws.logits);
  int* seq = reinterpret_cast<int*>(base + ws.seq);
  int* n_tok = reinterpret_cast<int*>(base + ws.n_tok);
  int* finished = reinterpret_cast<int*>(base + ws.finished);
  int* max_steps = reinterpret_cast<int*>(base + ws.max_steps);
  int* target = reinterpret_cast<int*>(base + ws.target);
  int* alive = reinterpret_cast<int*>(base + ws.alive);

  // cross K/V of every layer straight from the encoder memory: [B*T, L*2*D]
  { ProfScope ps(PS_DEC_CROSSKV, stream);
    KIRI_TRY(gemm_call(mem_bf16, w.crosskv_w, w.crosskv_b, B * T, L * 2 * D, d.enc_dim, EPI_BIAS_BF16, crosskv, nullptr,
                       nullptr, nullptr, nullptr, stream)); }
  int* step_dev = alive + 1;                     // lives next to the alive counter
  KIRI_CHECK_CUDA(cudaMemsetAsync(step_dev, 0, sizeof(int), stream));

  static int* alive_host = nullptr;
  if (!alive_host) KIRI_CHECK_CUDA(cudaMallocHost(&alive_host, sizeof(int)));

  // Default: the whole decode in ONE persistent cluster kernel (decoder_fused.cu).  The
  // step-per-launch path below is kept for A/B runs (KIRI_DEC_LEGACY=1).
  if (getenv("KIRI_DEC_LEGACY") == nullptr) {
    int cs = 8;
    if (const char* e = getenv("KIRI_DEC_CLUSTER")) cs = atoi(e);
    { ProfScope ps_step(PS_DEC_STEP, stream);
      KIRI_TRY(fused_decoder_run(h, crosskv, L * 2 * D, nullptr, nullptr, T, self_k, self_v, len_est, forced_ids, nullptr, B, Lmax, p,
                                 ids, n_out, sum_logp, step_logp, step_prob, step_dev, cs, stream)); }
    if (steps_run_host) {
      KIRI_CHECK_CUDA(cudaMemcpyAsync(alive_host, step_dev, sizeof(int), cudaMemcpyDeviceToHost, stream));
      KIRI_CHECK_CUDA(cudaStreamSynchronize(stream));
      *steps_run_host = *alive_host;
    }
    return 0;
  }
  dec_init_kernel<<<(B + 127) / 128, 128, 0, stream>>>(len_est, B, T, Lmax, *p, seq, n_tok, finished, max_steps, target,
                                                       alive, n_out, sum_logp);
  KIRI_CHECK_CUDA(cudaGetLastError());


  // one decode step = 3 + 11 * layers launches reading `step` from device memory, so a step can be
  // captured once per self-attention bucket and replayed as a CUDA graph (the step is launch-bound)
  const int cross_bucket = att_bucket(T);
  KIRI_REQUIRE(cross_bucket >= 0, "kiri_decode_greedy: memory length %d unsupported", T);
  auto enqueue_step = [&](int self_bucket) -> int {
    dec_embed_ln_kernel<<<(B + 7) / 8, 256, 0, stream>>>(seq, finished, B, Lmax, step_dev, w.dec_emb, w.dec_pe,
                                                         d.has_dec_pos, w.dec[0].ln1_g, w.dec[0].ln1_b, x, a);
    KIRI_CHECK_CUDA(cudaGetLastError());
    for (int l = 0; l < L; ++l) {
      const KiriDecLayerWeights& lw = w.dec[l];
      KIRI_TRY(gemm_call(a, lw.wqkv, lw.bqkv, B, 3 * D, D, EPI_BIAS_BF16, qkv, nullptr, nullptr, nullptr, nullptr, stream));
      __nv_bfloat16* kc = self_k + static_cast<size_t>(l) * B * Lmax * D;
      __nv_bfloat16* vc = self_v + static_cast<size_t>(l) * B * Lmax * D;
      KIRI_TRY(launch_attention<true>(qkv, 3 * D, qkv + D, qkv + 2 * D, kc, vc, Lmax, D, 0, self_bucket, step_dev,
                                      nullptr, B, heads, o, stream));
      KIRI_TRY(gemm_call(o, lw.wo, lw.bo, B, D, D, EPI_BIAS_RESID_LN, x, x, lw.ln2_g, lw.ln2_b, a, stream));
      KIRI_TRY(gemm_call(a, lw.wcq, lw.bcq, B, D, D, EPI_BIAS_BF16, qc, nullptr, nullptr, nullptr, nullptr, stream));
      // cross K/V of layer l: columns [l*2D, l*2D + D) are K, the next D are V; rows b*T + t
      KIRI_TRY(launch_attention<false>(qc, D, nullptr, nullptr, crosskv + static_cast<size_t>(l) * 2 * D,
                                       crosskv + static_cast<size_t>(l) * 2 * D + D, T, L * 2 * D, T, cross_bucket,
                                       step_dev, nullptr, B, heads, o, stream));
      KIRI_TRY(gemm_call(o, lw.wco, lw.bco, B, D, D, EPI_BIAS_RESID_LN, x, x, lw.ln3_g, lw.ln3_b, a, stream));
      KIRI_TRY(gemm_call(a, lw.w1, lw.b1, B, d.dec_ff, D, EPI_BIAS_GELU_BF16, hb, nullptr, nullptr, nullptr, nullptr, stream));
      const float* ng = (l + 1 < L) ? w.dec[l + 1].ln1_g : w.dec_ln_g;
      const float* nb = (l + 1 < L) ? w.dec[l + 1].ln1_b : w.dec_ln_b;
      KIRI_TRY(gemm_call(hb, lw.w2, lw.b2, B, D, d.dec_ff, EPI_BIAS_RESID_LN, x, x, ng, nb, a, stream));
    }
    KIRI_TRY(gemm_call(a, w.heads_w, w.heads_b, B, 2 * Vp, D, EPI_BIAS_F32, logits, nullptr, nullptr, nullptr, nullptr, stream));
    dec_select_kernel<<<(B + 3) / 4, 128, 0, stream>>>(logits, B, Vd, Vp, Lmax, step_dev, *p, seq, n_tok, finished,
                                                       max_steps, target, alive, forced_ids, ids, n_out, sum_logp,
                                                       step_logp, step_prob);
    KIRI_CHECK_CUDA(cudaGetLastError());
    dec_advance_kernel<<<1, 1, 0, stream>>>(step_dev);
    KIRI_CHECK_CUDA(cudaGetLastError());
    return 0;
  };

  if (poll_every <= 0) poll_every = 8;
  cudaGraphExec_t graphs[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // the legacy default stream cannot be captured: fall back to plain launches there
  const bool use_graphs = getenv("KIRI_NO_GRAPH") == nullptr && stream != nullptr && stream != cudaStreamLegacy;
  int rc = 0;
  int step = 0;
  for (; step < Lmax; ++step) {
    ProfScope ps_step(PS_DEC_STEP, stream);
    const int bucket = att_bucket(step + 1);
    if (step == 0 || !use_graphs) {
      // the first step runs eagerly (it also sets the kernels' one-time attributes)
      if ((rc = enqueue_step(bucket)) != 0) break;
    } else {
      if (!graphs[bucket]) {
        cudaGraph_t g = nullptr;
        if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) { rc = -2; set_last_error("kiri_decode_greedy: stream capture failed to begin"); break; }
        const int erc = enqueue_step(bucket);
        const cudaError_t ce = cudaStreamEndCapture(stream, &g);
        if (erc != 0 || ce != cudaSuccess || !g) { rc = erc ? erc : -2; if (!erc) set_last_error("kiri_decode_greedy: stream capture failed: %s", cudaGetErrorString(ce)); break; }
        const cudaError_t ie = cudaGraphInstantiate(&graphs[bucket], g, 0);
        cudaGraphDestroy(g);
        if (ie != cudaSuccess) { rc = -2; set_last_error("kiri_decode_greedy: graph instantiate failed: %s", cudaGetErrorString(ie)); break; }
      }
      if (cudaGraphLaunch(graphs[bucket], stream) != cudaSuccess) { rc = -2; set_last_error("kiri_decode_greedy: graph launch failed"); break; }
    }
    if ((step + 1) % poll_every == 0 || step + 1 == Lmax) {
      if (cudaMemcpyAsync(alive_host, alive, sizeof(int), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
          cudaStreamSynchronize(stream) != cudaSuccess) { rc = -2; set_last_error("kiri_decode_greedy: poll failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
      if (*alive_host <= 0) { ++step; break; }
    }
  }
  if (rc == 0) cudaStreamSynchronize(stream);      // graphs must be idle before they are destroyed
  for (int i = 0; i < 6; ++i)
    if (graphs[i]) cudaGraphExecDestroy(graphs[i]);
  if (rc != 0) return rc;
  if (steps_run_host) *steps_run_host = step;
  return 0;
}

// ------------------------------------------------------------------ multi-group (concatenated token stream)
extern "C" size_t kiri_decode_multi_workspace_bytes(const KiriHandle* h, int B, long long M_total, int Lmax) {
  if (!h || B <= 0 || M_total <= 0 || Lmax <= 0) return 0;
  const size_t L = h->d.dec_layers, D = h->d.dec_dim;
  return 2 * al256(static_cast<size_t>(M_total) * L * 2 * D * 2) + 2 * al256(L * B * Lmax * D * 2) + 256;
}

extern "C" int kiri_decode_greedy_multi(KiriHandle* h, const void* mem_bf16, long long M_total, const int* mem_row0,
                                        const int* mem_len, int max_T, const int* len_est, const int* line_perm, int B, int Lmax,
                                        const KiriDecodeParams* p, void* workspace, size_t workspace_bytes, int* ids,
                                        int* n_out, float* sum_logp, float* step_logp, float* step_prob,
                                        const int* forced_ids, int* steps_run_host, cudaStream_t stream) {
  KIRI_REQUIRE(h && mem_bf16 && mem_row0 && mem_len && len_est && p && workspace && ids && n_out && sum_logp,
               "kiri_decode_greedy_multi: null pointer");
  KIRI_REQUIRE(B > 0 && M_total > 0 && M_total < (1ll << 31) && Lmax > 0 && Lmax <= h->d.max_pos && Lmax <= 544,
               "kiri_decode_greedy_multi: bad sizes B=%d M=%lld Lmax=%d", B, M_total, Lmax);
  const KiriDims& d = h->d;
  const KiriWeights& w = h->w;
  const size_t L = d.dec_layers, D = d.dec_dim;
  KIRI_REQUIRE(workspace_bytes >= kiri_decode_multi_workspace_bytes(h, B, M_total, Lmax),
               "kiri_decode_greedy_multi: workspace too small");
  uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
  __nv_bfloat16* crosskv = reinterpret_cast<__nv_bfloat16*>(base);
  size_t off = al256(static_cast<size_t>(M_total) * L * 2 * D * 2);
  __nv_bfloat16* crosskv_hm = reinterpret_cast<__nv_bfloat16*>(base + off); off += al256(static_cast<size_t>(M_total) * L * 2 * D * 2);
  __nv_bfloat16* self_k = reinterpret_cast<__nv_bfloat16*>(base + off); off += al256(L * B * Lmax * D * 2);
  __nv_bfloat16* self_v = reinterpret_cast<__nv_bfloat16*>(base + off); off += al256(L * B * Lmax * D * 2);
  int* steps_dev = reinterpret_cast<int*>(base + off);
  { ProfScope ps(PS_DEC_CROSSKV, stream);
    KIRI_TRY(gemm_call(mem_bf16, w.crosskv_w, w.crosskv_b, static_cast<int>(M_total), static_cast<int>(L * 2 * D), d.enc_dim,
                       EPI_BIAS_BF16, crosskv, nullptr, nullptr, nullptr, nullptr, stream));
    KIRI_TRY(crosskv_headmajor(crosskv, crosskv_hm, static_cast<int>(L * 2 * D), mem_row0, mem_len, 0, max_T, B, stream)); }
  KIRI_CHECK_CUDA(cudaMemsetAsync(steps_dev, 0, sizeof(int), stream));
  int cs = 8;
  if (const char* e = getenv("KIRI_DEC_CLUSTER")) cs = atoi(e);
  { ProfScope ps_step(PS_DEC_STEP, stream);
    KIRI_TRY(fused_decoder_run(h, crosskv_hm, static_cast<int>(L * 2 * D), mem_row0, mem_len, 0, self_k, self_v, len_est,
                               forced_ids, line_perm, B, Lmax, p, ids, n_out, sum_logp, step_logp, step_prob, steps_dev, cs, stream,
                               nullptr, 1)); }
  if (steps_run_host) {
    static int* steps_host = nullptr;
    if (!steps_host) KIRI_CHECK_CUDA(cudaMallocHost(&steps_host, sizeof(int)));
    KIRI_CHECK_CUDA(cudaMemcpyAsync(steps_host, steps_dev, sizeof(int), cudaMemcpyDeviceToHost, stream));
    KIRI_CHECK_CUDA(cudaStreamSynchronize(stream));
    *steps_run_host = *steps_host;
  }
  return 0;
}

// ------------------------------------------------------------------ beam search (config 4)
extern "C" size_t kiri_decode_beam_workspace_bytes(const KiriHandle* h, int B, long long M_total, int Lmax, int beam) {
  if (!h || B <= 0 || M_total <= 0 || Lmax <= 0 || beam < 1 || beam > 5) return 0;
  const size_t L = h->d.dec_layers, D = h->d.dec_dim, ns = fused_decoder_slots(B, beam);
  return 2 * al256(static_cast<size_t>(M_total) * L * 2 * D * 2) + 2 * al256(L * ns * Lmax * D * 2) +
         2 * al256(2 * ns * static_cast<size_t>(Lmax) * 4) + 256;
}

extern "C" int kiri_decode_beam_multi(KiriHandle* h, const void* mem_bf16, long long M_total, const int* mem_row0,
                                      const int* mem_len, int max_T, const int* len_est, const int* line_perm, int B, int Lmax,
                                      int beam, double lenp, const KiriDecodeParams* p, void* workspace,
                                      size_t workspace_bytes, double* bm_score, int* bm_len, int* bm_state,
                                      int* bm_ids, float* bm_logp, cudaStream_t stream) {
  KIRI_REQUIRE(h && mem_bf16 && mem_row0 && mem_len && len_est && p && workspace && bm_score && bm_len && bm_state &&
               bm_ids && bm_logp, "kiri_decode_beam_multi: null pointer");
  KIRI_REQUIRE(B > 0 && M_total > 0 && M_total < (1ll << 31) && Lmax > 0 && Lmax <= h->d.max_pos && Lmax <= 544 &&
               beam >= 1 && beam <= 5, "kiri_decode_beam_multi: bad sizes B=%d M=%lld Lmax=%d beam=%d", B, M_total, Lmax, beam);
  KIRI_REQUIRE(p->select_raw == 0, "kiri_decode_beam_multi: the raw-logit selection rule is greedy-only");
  const KiriDims& d = h->d;
  const KiriWeights& w = h->w;
  const size_t L = d.dec_layers, D = d.dec_dim, ns = fused_decoder_slots(B, beam);
  KIRI_REQUIRE(workspace_bytes >= kiri_decode_beam_workspace_bytes(h, B, M_total, Lmax, beam),
               "kiri_decode_beam_multi: workspace too small");
  uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
  size_t off = 0;
  __nv_bfloat16* crosskv = reinterpret_cast<__nv_bfloat16*>(base); off += al256(static_cast<size_t>(M_total) * L * 2 * D * 2);
  __nv_bfloat16* crosskv_hm = reinterpret_cast<__nv_bfloat16*>(base + off); off += al256(static_cast<size_t>(M_total) * L * 2 * D * 2);
  __nv_bfloat16* self_k = reinterpret_cast<__nv_bfloat16*>(base + off); off += al256(L * ns * Lmax * D * 2);
  __nv_bfloat16* self_v = reinterpret_cast<__nv_bfloat16*>(base + off); off += al256(L * ns * Lmax * D * 2);
  FusedBeam fb;
  fb.beam = beam; fb.lenp = lenp;
  fb.seqbuf = reinterpret_cast<int*>(base + off); off += al256(2 * ns * static_cast<size_t>(Lmax) * 4);
  fb.lpbuf = reinterpret_cast<float*>(base + off); off += al256(2 * ns * static_cast<size_t>(Lmax) * 4);
  int* steps_dev = reinterpret_cast<int*>(base + off);
  fb.score = bm_score; fb.len = bm_len; fb.state = bm_state; fb.ids = bm_ids; fb.logp = bm_logp;
  { ProfScope ps(PS_DEC_CROSSKV, stream);
    KIRI_TRY(gemm_call(mem_bf16, w.crosskv_w, w.crosskv_b, static_cast<int>(M_total), static_cast<int>(L * 2 * D), d.enc_dim,
                       EPI_BIAS_BF16, crosskv, nullptr, nullptr, nullptr, nullptr, stream));
    KIRI_TRY(crosskv_headmajor(crosskv, crosskv_hm, static_cast<int>(L * 2 * D), mem_row0, mem_len, 0, max_T, B, stream)); }
  KIRI_CHECK_CUDA(cudaMemsetAsync(steps_dev, 0, sizeof(int), stream));
  int cs = 8;
  if (const char* e = getenv("KIRI_DEC_CLUSTER")) cs = atoi(e);
  ProfScope ps_step(PS_DEC_STEP, stream);
  // the self-attention cache is indexed by physical slot: B of the run function = decode slots in use
  return fused_decoder_run(h, crosskv_hm, static_cast<int>(L * 2 * D), mem_row0, mem_len, 0, self_k, self_v, len_est, nullptr,
                           line_perm, B, Lmax, p, nullptr, nullptr, nullptr, nullptr, nullptr, steps_dev, cs, stream, &fb, 1);
}
