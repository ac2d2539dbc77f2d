// K11-K12: host entry points of the batched attention decoder (greedy "accurate", beam, live streaming).
//
// Replaces  beam_decode_one_batched               kiri_ocr/model.py:390-600 (BEAM = 1 via core.py:560-568, BEAM > 1)
//           greedy_decode_streaming (token rule)   kiri_ocr/model.py:779-946
//           beam_decode_streaming (pruning rule)   kiri_ocr/model.py:949-1152
// The reference re-runs the whole decoder over the full prefix for every new token, re-projects the memory into K/V
// at every step and syncs to the host per step.  Here the cross-attention K/V of all layers come from ONE GEMM over
// the encoder memory ((W_k;W_v)_l @ W_memproj is folded on the host), re-laid head-major once per batch, and the
// whole decode of the batch runs in ONE persistent thread-block-cluster kernel (decoder_fused.cu): embedding,
// KV-cached self-attention, cross-attention, FFN, both heads, log-softmax, LM fusion, the four cumulative repeat
// penalties, UNK/EOS adjustments, arg-max / top-k + beam bookkeeping and the stop rules all stay on the device.
// (Round 1 also carried a step-per-launch form of the decoder for A/B runs - 27 dependent launches per step, 431 us
// per step against 119 us; it was removed in round 2, profiles/r01_launches_accurate_legacy_steps.csv keeps its trace.)
#include <cstdlib>

#include "internal.cuh"

namespace kiri {
static inline size_t al256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

static int cluster_size() {
  int cs = 8;
  if (const char* e = getenv("KIRI_DEC_CLUSTER")) cs = atoi(e);
  return cs;
}
}  // namespace kiri

using namespace kiri;

// ------------------------------------------------------------------ one width group (uniform T)
extern "C" size_t kiri_decode_workspace_bytes(const KiriHandle* h, int B, int T, int Lmax) {
  if (!h || B <= 0 || T <= 0 || Lmax <= 0) return 0;
  const size_t L = h->d.dec_layers, D = h->d.dec_dim;
  return al256(static_cast<size_t>(B) * T * L * 2 * D * 2) + 2 * al256(L * B * Lmax * D * 2) + 256;
}

extern "C" int kiri_decode_greedy(KiriHandle* h, const void* mem_bf16, const int* len_est, int B, int T, int Lmax,
                                  const KiriDecodeParams* p, void* workspace, size_t workspace_bytes, int* ids,
                                  int* n_out, float* sum_logp, float* step_logp, float* step_prob,
                                  const int* forced_ids, int* steps_run_host, int poll_every,
                                  cudaStream_t stream) {
  (void)poll_every;                                   // (the persistent kernel needs no host polling)
  KIRI_REQUIRE(h && mem_bf16 && len_est && p && workspace && ids && n_out && sum_logp, "kiri_decode_greedy: null pointer");
  KIRI_REQUIRE(B > 0 && T > 0 && Lmax > 0 && Lmax <= h->d.max_pos && Lmax <= 544, "kiri_decode_greedy: bad sizes B=%d T=%d Lmax=%d", B, T, Lmax);
  const KiriDims& d = h->d;
  const KiriWeights& w = h->w;
  const size_t L = d.dec_layers, D = d.dec_dim;
  KIRI_REQUIRE(workspace_bytes >= kiri_decode_workspace_bytes(h, B, T, Lmax), "kiri_decode_greedy: workspace too small");
  uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
  size_t off = 0;
  __nv_bfloat16* crosskv = reinterpret_cast<__nv_bfloat16*>(base); off += al256(static_cast<size_t>(B) * T * L * 2 * D * 2);
  __nv_bfloat16* self_k = reinterpret_cast<__nv_bfloat16*>(base + off); off += al256(L * B * Lmax * D * 2);
  __nv_bfloat16* self_v = reinterpret_cast<__nv_bfloat16*>(base + off); off += al256(L * B * Lmax * D * 2);
  int* steps_dev = reinterpret_cast<int*>(base + off);
  // cross K/V of every layer straight from the encoder memory: [B*T, L*2*D]
  { ProfScope ps(PS_DEC_CROSSKV, stream);
    KIRI_TRY(gemm_call(mem_bf16, w.crosskv_w, w.crosskv_b, B * T, static_cast<int>(L * 2 * D), d.enc_dim, EPI_BIAS_BF16, crosskv,
                       nullptr, nullptr, nullptr, nullptr, stream)); }
  KIRI_CHECK_CUDA(cudaMemsetAsync(steps_dev, 0, sizeof(int), stream));
  { ProfScope ps_step(PS_DEC_STEP, stream);
    KIRI_TRY(fused_decoder_run(h, crosskv, static_cast<int>(L * 2 * D), nullptr, nullptr, T, self_k, self_v, len_est, forced_ids,
                               nullptr, B, Lmax, p, ids, n_out, sum_logp, step_logp, step_prob, steps_dev, cluster_size(), stream)); }
  if (steps_run_host) {
    static int* steps_host = nullptr;
    if (!steps_host) KIRI_CHECK_CUDA(cudaMallocHost(&steps_host, sizeof(int)));
    KIRI_CHECK_CUDA(cudaMemcpyAsync(steps_host, steps_dev, sizeof(int), cudaMemcpyDeviceToHost, stream));
    KIRI_CHECK_CUDA(cudaStreamSynchronize(stream));
    *steps_run_host = *steps_host;
  }
  return 0;
}

// ------------------------------------------------------------------ multi-group (concatenated token stream)
extern "C" size_t kiri_decode_multi_workspace_bytes(const KiriHandle* h, int B, long long M_total, int Lmax) {
  if (!h || B <= 0 || M_total <= 0 || Lmax <= 0) return 0;
  const size_t L = h->d.dec_layers, D = h->d.dec_dim;
  return 2 * al256(static_cast<size_t>(M_total) * L * 2 * D * 2) + 2 * al256(L * B * Lmax * D * 2) + 256;
}

extern "C" int kiri_decode_greedy_multi(KiriHandle* h, const void* mem_bf16, long long M_total, const int* mem_row0,
                                        const int* mem_len, int max_T, const int* len_est, const int* line_perm, int n_slots,
                                        int B, int Lmax, const KiriDecodeParams* p, void* workspace, size_t workspace_bytes,
                                        int* ids, int* n_out, float* sum_logp, float* step_logp, float* step_prob,
                                        const int* forced_ids, int* steps_run_host, int* progress, int publish,
                                        cudaStream_t stream) {
  KIRI_REQUIRE(h && mem_bf16 && mem_row0 && mem_len && len_est && p && workspace && ids && n_out && sum_logp,
               "kiri_decode_greedy_multi: null pointer");
  KIRI_REQUIRE(B > 0 && M_total > 0 && M_total < (1ll << 31) && Lmax > 0 && Lmax <= h->d.max_pos && Lmax <= 544,
               "kiri_decode_greedy_multi: bad sizes B=%d M=%lld Lmax=%d", B, M_total, Lmax);
  const KiriDims& d = h->d;
  const KiriWeights& w = h->w;
  const size_t L = d.dec_layers, D = d.dec_dim;
  if (n_slots <= 0 || !line_perm) n_slots = B;
  KIRI_REQUIRE(n_slots >= B && n_slots <= 16 * B + 16, "kiri_decode_greedy_multi: %d decode slots for %d lines", n_slots, B);
  KIRI_REQUIRE(workspace_bytes >= kiri_decode_multi_workspace_bytes(h, n_slots, M_total, Lmax),
               "kiri_decode_greedy_multi: workspace too small");
  uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
  __nv_bfloat16* crosskv = reinterpret_cast<__nv_bfloat16*>(base);
  size_t off = al256(static_cast<size_t>(M_total) * L * 2 * D * 2);
  __nv_bfloat16* crosskv_hm = reinterpret_cast<__nv_bfloat16*>(base + off); off += al256(static_cast<size_t>(M_total) * L * 2 * D * 2);
  __nv_bfloat16* self_k = reinterpret_cast<__nv_bfloat16*>(base + off); off += al256(L * n_slots * Lmax * D * 2);
  __nv_bfloat16* self_v = reinterpret_cast<__nv_bfloat16*>(base + off); off += al256(L * n_slots * Lmax * D * 2);
  int* steps_dev = reinterpret_cast<int*>(base + off);
  { ProfScope ps(PS_DEC_CROSSKV, stream);
    KIRI_TRY(gemm_call(mem_bf16, w.crosskv_w, w.crosskv_b, static_cast<int>(M_total), static_cast<int>(L * 2 * D), d.enc_dim,
                       EPI_BIAS_BF16, crosskv, nullptr, nullptr, nullptr, nullptr, stream));
    KIRI_TRY(crosskv_headmajor(crosskv, crosskv_hm, static_cast<int>(L * 2 * D), mem_row0, mem_len, 0, max_T, B, stream)); }
  KIRI_CHECK_CUDA(cudaMemsetAsync(steps_dev, 0, sizeof(int), stream));
  FusedLive live = {publish, progress, 0, nullptr};
  { ProfScope ps_step(PS_DEC_STEP, stream);
    KIRI_TRY(fused_decoder_run(h, crosskv_hm, static_cast<int>(L * 2 * D), mem_row0, mem_len, 0, self_k, self_v, len_est,
                               forced_ids, line_perm, n_slots, Lmax, p, ids, n_out, sum_logp, step_logp, step_prob, steps_dev,
                               cluster_size(), stream, nullptr, 1, &live)); }
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
                                      int* bm_ids, float* bm_logp, int stream_rule, int* bm_trace, int* progress, int publish,
                                      cudaStream_t stream) {
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
  FusedLive live = {publish, progress, stream_rule, bm_trace};
  ProfScope ps_step(PS_DEC_STEP, stream);
  // the self-attention cache is indexed by physical slot: B of the run function = decode slots in use
  return fused_decoder_run(h, crosskv_hm, static_cast<int>(L * 2 * D), mem_row0, mem_len, 0, self_k, self_v, len_est, nullptr,
                           line_perm, B, Lmax, p, nullptr, nullptr, nullptr, nullptr, nullptr, steps_dev, cluster_size(), stream,
                           &fb, 1, &live);
}
