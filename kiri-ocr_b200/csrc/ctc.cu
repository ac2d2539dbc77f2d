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
#include "internal.cuh"

namespace kiri {

static constexpr int kCtcThreads = 256;
static constexpr int kCtcMaxT = 1024;

template <typename T> __device__ __forceinline__ float ld_logit(const T* p);
template <> __device__ __forceinline__ float ld_logit<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_logit<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}

// one frame row held by a warp: lane l owns float4 chunks l and l+32 (C <= 256 classes), or scalar
// strided elements for other element types / unaligned rows
// 16-byte loads, the whole row in a warp's registers: ONE pass over the logits (the scalar version
// read every row twice and issued 4-byte loads)
__device__ __forceinline__ void ctc_load_row(const float* __restrict__ row, int C, int lane, float (&v)[8]) {
  const float4* r4 = reinterpret_cast<const float4*>(row);
  const int nch = C >> 2;                                      // full float4 chunks
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int ch = lane + 32 * h;
    float4 t = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (ch < nch) t = __ldg(r4 + ch);
    else if (ch == nch && (C & 3)) {                           // ragged tail (C not a multiple of 4)
      const float* rs = row + 4 * ch;
      const int nv = C & 3;
      t.x = __ldg(rs); if (nv > 1) t.y = __ldg(rs + 1); if (nv > 2) t.z = __ldg(rs + 2);
    }
    v[4 * h] = t.x; v[4 * h + 1] = t.y; v[4 * h + 2] = t.z; v[4 * h + 3] = t.w;
  }
}

template <typename T>
__device__ __forceinline__ void ctc_frame(const T* __restrict__ row, int C, bool vec, int lane, int& am_out, float& p_out,
                                          const float* v) {
  if (vec) {
    float m = -INFINITY;
    int am = 0x7fffffff;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float x = v[4 * h + k];
        if (x > m) { m = x; am = 4 * (lane + 32 * h) + k; }     // ascending class index per lane: first max wins
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, o);
      const int oa = __shfl_xor_sync(0xffffffffu, am, o);
      if (om > m || (om == m && oa < am)) { m = om; am = oa; }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += __expf(v[k] - m);          // exp(-inf) = 0 for the padding
    s = warp_sum(s);
    am_out = am; p_out = 1.0f / s;
    return;
  }
  // scalar path: max and FIRST arg-max (torch.argmax tie rule), then sum exp(x - max)
  float m = -INFINITY;
  int am = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    const float x = ld_logit<T>(row + c);
    if (x > m) { m = x; am = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > m || (om == m && oa < am)) { m = om; am = oa; }
  }
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += __expf(ld_logit<T>(row + c) - m);
  s = warp_sum(s);
  am_out = am; p_out = 1.0f / s;
}

#ifndef KIRI_CTC_IT
#define KIRI_CTC_IT 1          // warp passes (4 frames each) whose loads are issued together
#endif
#ifndef KIRI_CTC_MINB
#define KIRI_CTC_MINB 4        // resident CTAs per SM the register budget is held to
#endif
template <typename T>
__global__ void __launch_bounds__(kCtcThreads, KIRI_CTC_MINB)
ctc_greedy_kernel(const T* __restrict__ logits, int Tn, int C, int ld, int* __restrict__ ids,
                  int* __restrict__ n_ids, float* __restrict__ conf, int* __restrict__ frame_ids,
                  float* __restrict__ frame_prob, const int* __restrict__ row0, const int* __restrict__ lens) {
  __shared__ int s_id[kCtcMaxT];
  __shared__ float s_psum[kCtcThreads / 32];
  __shared__ int s_cnt[kCtcThreads / 32];
  const int line = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = kCtcThreads >> 5;
  pdl_trigger();
  pdl_wait();                                       // the logits come from the previous kernel
  // fixed-shape batch: line i owns rows [i*Tn, (i+1)*Tn); token-stream form: rows [row0[i], +lens[i])
  const size_t r0 = row0 ? static_cast<size_t>(row0[line]) : static_cast<size_t>(line) * Tn;
  if (lens) Tn = lens[line];
  const T* base = logits + r0 * ld;
  const bool vec = (sizeof(T) == 4) && (C <= 256) && ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);

  float psum = 0.f;
  if (vec) {
    // EIGHT lanes per frame, four frames per warp pass: a row of C <= 256 fp32 logits is at most 64 float4 chunks, lane
    // `sub` of a frame's group owns chunks sub, sub + 8, ...  The reductions are 3 shuffle steps that serve four frames
    // at once and the loads of two passes (8 frames per warp) are issued before any arithmetic.  (One frame per warp
    // needed 5-step reductions for a single frame: 208 warp instructions per frame, issue-bound at 54 % of the HBM
    // peak at 8 192 lines; this form needs ~45.)
    const int sub = lane & 7, slot = lane >> 3;
    const int nch = C >> 2, tail = C & 3;
    constexpr int kIt = KIRI_CTC_IT;
    for (int tb = warp * 4; tb < Tn; tb += nwarps * 4 * kIt) {
      float4 v[kIt][8];
#pragma unroll
      for (int it = 0; it < kIt; ++it) {
        const int tf = tb + it * nwarps * 4 + slot;
        const float* row = reinterpret_cast<const float*>(base + static_cast<size_t>(tf) * ld);
        const float4* r4 = reinterpret_cast<const float4*>(row);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int ch = sub + 8 * k;
          float4 t4 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
          if (tf < Tn) {
            if (ch < nch) t4 = __ldg(r4 + ch);
            else if (ch == nch && tail) {                        // ragged tail (C not a multiple of 4)
              const float* rs = row + 4 * ch;
              t4.x = __ldg(rs); if (tail > 1) t4.y = __ldg(rs + 1); if (tail > 2) t4.z = __ldg(rs + 2);
            }
          }
          v[it][k] = t4;
        }
      }
#pragma unroll
      for (int it = 0; it < kIt; ++it) {
        const int tf = tb + it * nwarps * 4 + slot;
        float m = -INFINITY;
        int am = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < 8; ++k) {                            // ascending class index per lane: the first maximum wins
          const int c0 = 4 * (sub + 8 * k);
          if (v[it][k].x > m) { m = v[it][k].x; am = c0; }
          if (v[it][k].y > m) { m = v[it][k].y; am = c0 + 1; }
          if (v[it][k].z > m) { m = v[it][k].z; am = c0 + 2; }
          if (v[it][k].w > m) { m = v[it][k].w; am = c0 + 3; }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {                        // inside the frame's 8-lane group
          const float om = __shfl_xor_sync(0xffffffffu, m, o);
          const int oa = __shfl_xor_sync(0xffffffffu, am, o);
          if (om > m || (om == m && oa < am)) { m = om; am = oa; }
        }
        float sx = 0.f;
        if (tf < Tn) {
#pragma unroll
          for (int k = 0; k < 8; ++k)                             // exp(-inf) = 0 for the padding
            sx += (__expf(v[it][k].x - m) + __expf(v[it][k].y - m)) + (__expf(v[it][k].z - m) + __expf(v[it][k].w - m));
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) sx += __shfl_xor_sync(0xffffffffu, sx, o);
        if (sub == 0 && tf < Tn) {
          const float pr = 1.0f / sx;
          s_id[tf] = am;
          psum += pr;
          if (frame_ids) frame_ids[r0 + tf] = am;
          if (frame_prob) frame_prob[r0 + tf] = pr;
        }
      }
    }
    // the four frame slots of the warp, in a fixed order
    const float p1 = __shfl_sync(0xffffffffu, psum, 8), p2 = __shfl_sync(0xffffffffu, psum, 16), p3 = __shfl_sync(0xffffffffu, psum, 24);
    psum = ((psum + p1) + p2) + p3;                              // (meaningful in lane 0)
  } else {
    // scalar path (other element types, unaligned rows, C > 256): one frame per warp
    float dummy[8];
    for (int t = warp; t < Tn; t += nwarps) {
      int am = 0;
      float pr = 0.f;
      ctc_frame<T>(base + static_cast<size_t>(t) * ld, C, false, lane, am, pr, dummy);
      if (lane == 0) {
        s_id[t] = am;
        psum += pr;
        if (frame_ids) frame_ids[r0 + t] = am;
        if (frame_prob) frame_prob[r0 + t] = pr;
      }
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
    if (keep) ids[r0 + base_out + woff + __popc(bal & ((1u << lane) - 1))] = id;
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

// Collapse stage alone, for frame decisions that were taken in the CTC head's GEMM epilogue (EPI_CTC_STATS: the logits
// never reach HBM in fast mode): per line keep frame t iff id[t] != id[t-1] and id[t] >= 2, confidence = mean of the
// frames' arg-max probabilities in a fixed summation order.
__global__ void __launch_bounds__(128)
ctc_collapse_kernel(const int* __restrict__ frame_ids, const float* __restrict__ frame_prob, const int* __restrict__ row0,
                    const int* __restrict__ lens, int* __restrict__ ids, int* __restrict__ n_ids, float* __restrict__ conf) {
  __shared__ float s_psum[4];
  __shared__ int s_cnt[4];
  const int line = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  pdl_wait();
  const size_t r0 = static_cast<size_t>(row0[line]);
  const int Tn = lens[line];
  float psum = 0.f;
  int base_out = 0;
  for (int t0 = 0; t0 < Tn; t0 += 128) {
    const int t = t0 + threadIdx.x;
    bool keep = false;
    int id = 0;
    if (t < Tn) {
      id = __ldg(frame_ids + r0 + t);
      psum += __ldg(frame_prob + r0 + t);
      keep = (id >= 2) && (t == 0 || id != __ldg(frame_ids + r0 + t - 1));
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_cnt[warp] = __popc(bal);
    __syncthreads();
    int woff = 0, total = 0;
#pragma unroll
    for (int wi = 0; wi < 4; ++wi) {
      if (wi < warp) woff += s_cnt[wi];
      total += s_cnt[wi];
    }
    if (keep) ids[r0 + base_out + woff + __popc(bal & ((1u << lane) - 1))] = id;
    base_out += total;
    __syncthreads();
  }
  psum = warp_sum(psum);
  if (lane == 0) s_psum[warp] = psum;
  __syncthreads();
  if (threadIdx.x == 0) {
    conf[line] = ((s_psum[0] + s_psum[1]) + (s_psum[2] + s_psum[3])) / static_cast<float>(Tn);
    n_ids[line] = base_out;
  }
}

}  // namespace kiri

using namespace kiri;

extern "C" int kiri_ctc_collapse_multi(const int* frame_ids, const float* frame_prob, int n_lines, const int* row0, const int* len,
                                       int* ids, int* n_ids, float* conf, cudaStream_t stream) {
  KIRI_REQUIRE(frame_ids && frame_prob && row0 && len && ids && n_ids && conf, "kiri_ctc_collapse_multi: null pointer");
  if (n_lines <= 0) return 0;
  ProfScope ps(PS_CTC_GREEDY, stream);
  KIRI_CHECK_CUDA(launch_pdl(ctc_collapse_kernel, dim3(n_lines), dim3(128), 0, stream, frame_ids, frame_prob, row0, len, ids, n_ids, conf));
  return 0;
}

extern "C" int kiri_ctc_greedy(const void* logits, int logits_dtype, int n_lines, int T, int C, int ld,
                               int* ids, int* n_ids, float* conf, int* frame_ids, float* frame_prob,
                               cudaStream_t stream) {
  KIRI_REQUIRE(logits && ids && n_ids && conf, "kiri_ctc_greedy: null pointer");
  KIRI_REQUIRE(T > 0 && T <= kCtcMaxT && C > 0 && ld >= C, "kiri_ctc_greedy: bad shape T=%d C=%d ld=%d", T, C, ld);
  if (n_lines == 0) return 0;
  if (logits_dtype == KIRI_DTYPE_F32)
    ctc_greedy_kernel<float><<<n_lines, kCtcThreads, 0, stream>>>(
        reinterpret_cast<const float*>(logits), T, C, ld, ids, n_ids, conf, frame_ids, frame_prob, nullptr, nullptr);
  else if (logits_dtype == KIRI_DTYPE_BF16)
    ctc_greedy_kernel<__nv_bfloat16><<<n_lines, kCtcThreads, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(logits), T, C, ld, ids, n_ids, conf, frame_ids, frame_prob, nullptr, nullptr);
  else
    KIRI_REQUIRE(false, "kiri_ctc_greedy: unknown dtype %d", logits_dtype);
  KIRI_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int kiri_ctc_greedy_multi(const void* logits, int logits_dtype, int n_lines, const int* row0, const int* len,
                                     int max_T, int C, int ld, int* ids, int* n_ids, float* conf, int* frame_ids,
                                     float* frame_prob, cudaStream_t stream) {
  KIRI_REQUIRE(logits && row0 && len && ids && n_ids && conf, "kiri_ctc_greedy_multi: null pointer");
  KIRI_REQUIRE(max_T > 0 && max_T <= kCtcMaxT && C > 0 && ld >= C, "kiri_ctc_greedy_multi: bad shape T=%d C=%d ld=%d", max_T, C, ld);
  KIRI_REQUIRE(logits_dtype == KIRI_DTYPE_F32, "kiri_ctc_greedy_multi: fp32 logits only");
  if (n_lines == 0) return 0;
  ProfScope ps(PS_CTC_GREEDY, stream);
  KIRI_CHECK_CUDA(launch_pdl(ctc_greedy_kernel<float>, dim3(n_lines), dim3(kCtcThreads), 0, stream,
                             reinterpret_cast<const float*>(logits), max_T, C, ld, ids, n_ids, conf, frame_ids, frame_prob,
                             row0, len));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// K13: CTC forward-algorithm score of decoder hypotheses (beam rescoring).
//
// Replaces  compute_ctc_alignment_score   kiri_ocr/model.py:603-668  (a Python double loop of
//           T x S torch.logsumexp calls, ~0.5 s per hypothesis on the CPU)
// One CTA per line: all warps first reduce the per-frame log-softmax statistics, then warp r runs
// the alpha recursion of hypothesis r over the extended label sequence [b, l0, b, l1, ..., b] in
// fp32 log space, with the same 1/2/3-term logsumexp as the reference.
namespace kiri {

__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == -INFINITY) return -INFINITY;
  return m + logf(expf(a - m) + expf(b - m));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == -INFINITY) return -INFINITY;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

__global__ void __launch_bounds__(256)
ctc_align_kernel(const float* __restrict__ logits, int ld, int C, const int* __restrict__ mem_row0,
                 const int* __restrict__ mem_len, int beam, int Lmax, const int* __restrict__ bm_ids,
                 const int* __restrict__ bm_len, const int* __restrict__ bm_state, int vocab_size, int unk_ctc,
                 float* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  const int b = blockIdx.x;
  const int T = mem_len[b];
  const float* lg = logits + static_cast<size_t>(mem_row0[b]) * ld;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* fmax_s = reinterpret_cast<float*>(sm_raw);              // [T] row max
  float* flog_s = fmax_s + T;                                    // [T] log(sum(exp(x - max)))
  const int Smax = 2 * Lmax + 1;
  int* ext_all = reinterpret_cast<int*>(flog_s + T);
  // ---- per-frame log-softmax statistics (torch: x - max - log(sum(exp(x - max))))
  for (int t = warp; t < T; t += nwarps) {
    const float* row = lg + static_cast<size_t>(t) * ld;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(row[c] - m);
    s = warp_sum(s);
    if (lane == 0) { fmax_s[t] = m; flog_s[t] = logf(s); }
  }
  __syncthreads();
  if (warp >= beam) return;
  const int r = warp;
  if (bm_state[b * beam + r] == 0) { if (lane == 0) out[b * beam + r] = 0.f; return; }
  int* ext = ext_all + r * 3 * Smax;
  float* a0 = reinterpret_cast<float*>(ext + Smax);
  float* a1 = a0 + Smax;
  // ---- labels: ids up to EOS, pad/bos skipped, mapped to CTC ids (model.py:611-620, 137-144)
  const int* seq = bm_ids + (static_cast<size_t>(b) * beam + r) * Lmax;
  const int n_tok = bm_len[b * beam + r];
  int n_lab = 0;
  if (lane == 0) {
    for (int i = 0; i < n_tok; ++i) {
      const int x = seq[i];
      if (x == 2) break;
      if (x == 0 || x == 1) continue;
      const int raw = x - 3;
      ext[2 * n_lab + 1] = (raw >= 0 && raw < vocab_size) ? raw + 2 : unk_ctc;
      ++n_lab;
    }
    for (int i = 0; i <= n_lab; ++i) ext[2 * i] = 0;
  }
  n_lab = __shfl_sync(0xffffffffu, n_lab, 0);
  __syncwarp();
  auto lp = [&](int t, int c) { return (lg[static_cast<size_t>(t) * ld + c] - fmax_s[t]) - flog_s[t]; };
  if (n_lab == 0) {
    float s = 0.f;
    for (int t = lane; t < T; t += 32) s += lp(t, 0);
    s = warp_sum(s);
    if (lane == 0) out[b * beam + r] = s / static_cast<float>(T > 1 ? T : 1);
    return;
  }
  const int S = 2 * n_lab + 1;
  for (int s = lane; s < S; s += 32) a0[s] = -INFINITY;
  __syncwarp();
  if (lane == 0) { a0[0] = lp(0, 0); a0[1] = lp(0, ext[1]); }
  __syncwarp();
  float* cur = a0;
  float* nxt = a1;
  for (int t = 1; t < T; ++t) {
    for (int s = lane; s < S; s += 32) {
      const int e = ext[s];
      float v;
      if (s == 0) v = cur[0];
      else if (s > 1 && e != 0 && e != ext[s - 2]) v = lse3(cur[s], cur[s - 1], cur[s - 2]);
      else v = lse2(cur[s], cur[s - 1]);
      nxt[s] = v + lp(t, e);
    }
    __syncwarp();
    float* tmp = cur; cur = nxt; nxt = tmp;
  }
  if (lane == 0) {
    const float total = lse2(cur[S - 1], cur[S - 2]);
    out[b * beam + r] = total / static_cast<float>(n_lab);
  }
}

}  // namespace kiri

extern "C" int kiri_ctc_align_score(const float* logits, int ld, int C, const int* mem_row0, const int* mem_len,
                                    int n_lines, int beam, int Lmax, const int* bm_ids, const int* bm_len,
                                    const int* bm_state, int vocab_size, int unk_ctc_id, int max_T, float* out,
                                    cudaStream_t stream) {
  KIRI_REQUIRE(logits && mem_row0 && mem_len && bm_ids && bm_len && bm_state && out, "kiri_ctc_align_score: null pointer");
  KIRI_REQUIRE(beam >= 1 && beam <= 8 && Lmax > 0 && max_T > 0, "kiri_ctc_align_score: bad sizes");
  if (n_lines == 0) return 0;
  const int smem = max_T * 8 + beam * 3 * (2 * Lmax + 1) * 4;
  static int configured[kiri::kMaxDevices] = {0};
  const int dslot = kiri::kiri_cur_device_slot();
  if (smem > 48 * 1024 && configured[dslot] < smem) {
    KIRI_CHECK_CUDA(cudaFuncSetAttribute(kiri::ctc_align_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[dslot] = smem;
  }
  kiri::ctc_align_kernel<<<n_lines, 256, smem, stream>>>(logits, ld, C, mem_row0, mem_len, beam, Lmax, bm_ids, bm_len,
                                                         bm_state, vocab_size, unk_ctc_id, out);
  KIRI_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// The exchange step's payload: fixed-stride int32 records
// {n_ids, confidence bits, ids[T]} per line from the token-major CTC output, so that one all-gather moves
// the results of every rank (ids beyond n_ids are zero).
namespace kiri {
__global__ void __launch_bounds__(128)
pack_records_kernel(const int* __restrict__ ids, const int* __restrict__ n_ids, const float* __restrict__ conf,
                    const int* __restrict__ row0, int T, int* __restrict__ rec) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  const int n = n_ids[b];
  int* r = rec + static_cast<size_t>(b) * (2 + T);
  if (threadIdx.x == 0) { r[0] = n; r[1] = __float_as_int(conf[b]); }
  const int* src = ids + row0[b];
  for (int t = threadIdx.x; t < T; t += blockDim.x) r[2 + t] = t < n ? src[t] : 0;
}
}  // namespace kiri

extern "C" int kiri_pack_records(const int* ids, const int* n_ids, const float* conf, const int* mem_row0, int n_lines, int T,
                                 int* records, cudaStream_t stream) {
  KIRI_REQUIRE(ids && n_ids && conf && mem_row0 && records, "kiri_pack_records: null pointer");
  KIRI_REQUIRE(T > 0, "kiri_pack_records: T must be positive");
  if (n_lines == 0) return 0;
  KIRI_CHECK_CUDA(kiri::launch_pdl(kiri::pack_records_kernel, dim3(n_lines), dim3(128), 0, stream, ids, n_ids, conf, mem_row0, T,
                                   records));
  return 0;
}
