// K10: fused CTC greedy decode — per frame soft-max statistics + arg-max, then blank/repeat
// collapse, mean max-probability confidence and the length estimate, in one pass over the logits.
//
// Replaces  compute_ctc_confidence   kiri_ocr/model.py:343-373  (softmax, argmax, max, mean,
//                                    two D2H syncs and a Python loop per line)
//           CharTokenizer.decode_ctc kiri_ocr/model.py:109-119  (the id-level part: skip
//                                    idx == prev, then skip idx < 2)
//
// One CTA per line, one warp per frame (round-robin): lanes stride the class axis with
// coalesced loads, reduce (max, first arg-max) and sum(exp) with shuffles.  Collapse is a
// ballot/popcount compaction over the T frame ids.  HBM-bound: T*C*sizeof(logit) bytes in,
// ~4*T bytes out per line.
#include "common.cuh"
#include "kiri_b200.h"

namespace kiri {

static constexpr int kCtcThreads = 256;
static constexpr int kCtcMaxT = 1024;

template <typename T> __device__ __forceinline__ float ld_logit(const T* p);
template <> __device__ __forceinline__ float ld_logit<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_logit<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}

template <typename T>
__global__ void __launch_bounds__(kCtcThreads)
ctc_greedy_kernel(const T* __restrict__ logits, int Tn, int C, int ld, int* __restrict__ ids,
                  int* __restrict__ n_ids, float* __restrict__ conf, int* __restrict__ frame_ids,
                  float* __restrict__ frame_prob) {
  __shared__ int s_id[kCtcMaxT];
  __shared__ float s_psum[kCtcThreads / 32];
  __shared__ int s_cnt[kCtcThreads / 32];
  const int line = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = kCtcThreads >> 5;
  const T* base = logits + static_cast<size_t>(line) * Tn * ld;

  float psum = 0.f;
  for (int t = warp; t < Tn; t += nwarps) {
    const T* row = base + static_cast<size_t>(t) * ld;
    // pass 1: max and FIRST arg-max (torch.argmax tie rule)
    float m = -INFINITY;
    int am = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
      const float v = ld_logit<T>(row + c);
      if (v > m) { m = v; am = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, o);
      const int oa = __shfl_xor_sync(0xffffffffu, am, o);
      if (om > m || (om == m && oa < am)) { m = om; am = oa; }
    }
    // pass 2 (row is L1-resident): sum exp(x - max); max prob = 1 / sum
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += __expf(ld_logit<T>(row + c) - m);
    s = warp_sum(s);
    const float p = 1.0f / s;
    if (lane == 0) {
      s_id[t] = am;
      psum += p;
      if (frame_ids) frame_ids[static_cast<size_t>(line) * Tn + t] = am;
      if (frame_prob) frame_prob[static_cast<size_t>(line) * Tn + t] = p;
    }
  }
  if (lane == 0) s_psum[warp] = psum;
  __syncthreads();

  // collapse: keep frame t iff id[t] != id[t-1] and id[t] >= 2 (blank = 0, pad = 1)
  int base_out = 0;
  for (int t0 = 0; t0 < Tn; t0 += kCtcThreads) {
    const int t = t0 + threadIdx.x;
    bool keep = false;
    int id = 0;
    if (t < Tn) {
      id = s_id[t];
      keep = (id >= 2) && (t == 0 || id != s_id[t - 1]);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_cnt[warp] = __popc(bal);
    __syncthreads();
    int woff = 0, total = 0;
    for (int wi = 0; wi < nwarps; ++wi) {
      if (wi < warp) woff += s_cnt[wi];
      total += s_cnt[wi];
    }
    if (keep) ids[static_cast<size_t>(line) * Tn + base_out + woff + __popc(bal & ((1u << lane) - 1))] = id;
    base_out += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int wi = 0; wi < nwarps; ++wi) s += s_psum[wi];
    conf[line] = s / static_cast<float>(Tn);
    n_ids[line] = base_out;
  }
}

}  // namespace kiri

using namespace kiri;

extern "C" int kiri_ctc_greedy(const void* logits, int logits_dtype, int n_lines, int T, int C, int ld,
                               int* ids, int* n_ids, float* conf, int* frame_ids, float* frame_prob,
                               cudaStream_t stream) {
  KIRI_REQUIRE(logits && ids && n_ids && conf, "kiri_ctc_greedy: null pointer");
  KIRI_REQUIRE(T > 0 && T <= kCtcMaxT && C > 0 && ld >= C, "kiri_ctc_greedy: bad shape T=%d C=%d ld=%d", T, C, ld);
  if (n_lines == 0) return 0;
  if (logits_dtype == KIRI_DTYPE_F32)
    ctc_greedy_kernel<float><<<n_lines, kCtcThreads, 0, stream>>>(
        reinterpret_cast<const float*>(logits), T, C, ld, ids, n_ids, conf, frame_ids, frame_prob);
  else if (logits_dtype == KIRI_DTYPE_BF16)
    ctc_greedy_kernel<__nv_bfloat16><<<n_lines, kCtcThreads, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(logits), T, C, ld, ids, n_ids, conf, frame_ids, frame_prob);
  else
    KIRI_REQUIRE(false, "kiri_ctc_greedy: unknown dtype %d", logits_dtype);
  KIRI_CHECK_CUDA(cudaGetLastError());
  return 0;
}
