// C ABI glue: error text, kernel-level GEMM/conv entry points, the model handle and the
// stem + encoder + CTC-head pipeline (kiri_encode).  See include/kiri_b200.h.
#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "internal.cuh"

namespace kiri {

static thread_local char g_err[1024] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- per-stage event profiler
struct ProfRec { int stage; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_ev_pool;
static cudaEvent_t g_open[PS_COUNT];
static cudaEvent_t ev_get() {
  if (!g_ev_pool.empty()) { cudaEvent_t e = g_ev_pool.back(); g_ev_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
void prof_begin(int stage, cudaStream_t s) {
  if (!g_prof_on) return;
  g_open[stage] = ev_get();
  cudaEventRecord(g_open[stage], s);
}
void prof_end(int stage, cudaStream_t s) {
  if (!g_prof_on) return;
  cudaEvent_t e = ev_get();
  cudaEventRecord(e, s);
  g_prof.push_back({stage, g_open[stage], e});
}

// CUDA-core cross-check GEMM: one thread per output element, fp32 accumulate.
__global__ void gemm_ref_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ w,
                                int M, int N, int K, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k)
    acc = fmaf(__bfloat162float(a[(size_t)m * K + k]), __bfloat162float(w[(size_t)n * K + k]), acc);
  out[(size_t)m * N + n] = acc;
}

}  // namespace kiri

using namespace kiri;

extern "C" const char* kiri_last_error(void) { return g_err; }
extern "C" int kiri_version(void) { return 200; }
// ABI handshake for the ctypes binding: sizes of the structs that cross the boundary
extern "C" int kiri_abi_sizes(int* out, int n) {
  const int v[6] = {(int)sizeof(KiriCropDesc), (int)sizeof(KiriDims), (int)sizeof(KiriWeights), (int)sizeof(KiriGroup),
                    (int)sizeof(KiriDecodeParams), (int)sizeof(KiriEncLayerWeights)};
  for (int i = 0; i < n && i < 6; ++i) out[i] = v[i];
  return 6;
}
extern "C" int kiri_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}

extern "C" int kiri_profile_begin(void) {
  for (auto& r : g_prof) { g_ev_pool.push_back(r.a); g_ev_pool.push_back(r.b); }
  g_prof.clear();
  g_prof_on = true;
  return PS_COUNT;
}
extern "C" int kiri_profile_end(double* ms_by_stage, int* count_by_stage, int n) {
  g_prof_on = false;
  KIRI_REQUIRE(ms_by_stage && count_by_stage && n >= PS_COUNT, "kiri_profile_end: need arrays of %d entries", (int)PS_COUNT);
  for (int i = 0; i < n; ++i) { ms_by_stage[i] = 0.0; count_by_stage[i] = 0; }
  KIRI_CHECK_CUDA(cudaDeviceSynchronize());
  for (auto& r : g_prof) {
    float ms = 0.f;
    KIRI_CHECK_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    ms_by_stage[r.stage] += ms;
    count_by_stage[r.stage] += 1;
    g_ev_pool.push_back(r.a); g_ev_pool.push_back(r.b);
  }
  g_prof.clear();
  return 0;
}

extern "C" int kiri_gemm_ref(const void* a, const void* w, int M, int N, int K, float* out_f32,
                             cudaStream_t stream) {
  KIRI_REQUIRE(a && w && out_f32, "kiri_gemm_ref: null pointer");
  if (M == 0 || N == 0) return 0;
  dim3 grid((N + 127) / 128, M);
  gemm_ref_kernel<<<grid, 128, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(a),
                                            reinterpret_cast<const __nv_bfloat16*>(w), M, N, K, out_f32);
  KIRI_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int kiri::gemm_call(const void* a, const void* w, const float* bias, int M, int N, int K, int epi, void* out,
                     const float* resid, const float* ln_g, const float* ln_b, void* out2,
                     cudaStream_t stream) {
  if (M == 0) return 0;
  GemmLaunch L;
  memset(&L, 0, sizeof(L));
  L.a = a; L.w = w;
  L.NB = 1; L.IH = 1; L.IW = M; L.Cin = K;
  L.OH = 1; L.OW = M;
  L.sw = 1; L.sh = 1; L.pad = 0; L.kw = 1; L.kh = 1;
  L.N = N; L.epi = epi;
  L.e.out = out; L.e.bias = bias; L.e.resid = resid; L.e.ldc = N; L.e.n_valid = N;
  L.e.ln_g = ln_g; L.e.ln_b = ln_b; L.e.out2 = out2;
  return launch_gemm_tc(L, stream);
}

extern "C" int kiri_gemm_bf16(const void* a, const void* w, const float* bias, int M, int N, int K, int epi,
                              void* out, const float* resid, const float* ln_g, const float* ln_b,
                              void* out2, cudaStream_t stream) {
  KIRI_REQUIRE(a && w && bias && out, "kiri_gemm_bf16: null pointer");
  KIRI_REQUIRE(epi >= 0 && epi <= 5, "kiri_gemm_bf16: unknown epilogue %d", epi);
  KIRI_REQUIRE((epi != KIRI_EPI_BIAS_RESID_F32 && epi != KIRI_EPI_BIAS_RESID_LN) || resid,
               "kiri_gemm_bf16: residual epilogue without resid");
  return gemm_call(a, w, bias, M, N, K, epi, out, resid, ln_g, ln_b, out2, stream);
}

static GemmLaunch conv_launch(const void* in, const void* w, const float* bias, int n, int IH, int IW, int Cin, int N,
                              int sh, int sw, void* out, int cin_mem = 0) {
  GemmLaunch L;
  memset(&L, 0, sizeof(L));
  L.a = in; L.w = w;
  L.NB = n; L.IH = IH; L.IW = IW; L.Cin = Cin; L.Cin_mem = cin_mem;
  L.OH = (IH + 2 - 3) / sh + 1; L.OW = (IW + 2 - 3) / sw + 1;
  L.sw = sw; L.sh = sh; L.pad = 1; L.kw = 3; L.kh = 3;
  L.N = N; L.epi = EPI_BIAS_SILU_BF16;
  L.e.out = out; L.e.bias = bias; L.e.resid = nullptr; L.e.ldc = N; L.e.n_valid = N;
  return L;
}
extern "C" int kiri_conv3x3_bf16(const void* in_nhwc, const void* w, const float* bias, int n, int IH, int IW,
                                 int Cin, int N, int sh, int sw, void* out_nhwc, int cin_mem, cudaStream_t stream) {
  KIRI_REQUIRE(in_nhwc && w && bias && out_nhwc, "kiri_conv3x3_bf16: null pointer");
  KIRI_REQUIRE(cin_mem >= 0 && cin_mem <= Cin, "kiri_conv3x3_bf16: cin_mem=%d must be in [0, Cin=%d]", cin_mem, Cin);
  if (n == 0) return 0;
  const GemmLaunch L = conv_launch(in_nhwc, w, bias, n, IH, IW, Cin, N, sh, sw, out_nhwc, cin_mem == Cin ? 0 : cin_mem);
  if (conv2_swap_supported(&L, 1)) return launch_conv2_swap(&L, 1, stream);
  return launch_gemm_tc(L, stream);
}

// ------------------------------------------------------------------------------------------------
// model handle
// ------------------------------------------------------------------------------------------------
extern "C" int kiri_create(const KiriDims* dims, const KiriWeights* weights, KiriHandle** out) {
  KIRI_REQUIRE(dims && weights && out, "kiri_create: null pointer");
  KIRI_REQUIRE(kiri_device_ok(), "kiri_create: the current CUDA device is not compute capability 10.x (B200)");
  KIRI_REQUIRE(dims->enc_dim == 256 && dims->dec_dim == 256, "kiri_create: kernels are built for ENC_DIM = DEC_DIM = 256");
  KIRI_REQUIRE(dims->enc_heads * 32 == dims->enc_dim && dims->dec_heads * 32 == dims->dec_dim,
               "kiri_create: head_dim must be 32");
  KIRI_REQUIRE(dims->enc_layers <= KIRI_MAX_LAYERS && dims->dec_layers <= KIRI_MAX_LAYERS, "kiri_create: too many layers");
  KIRI_REQUIRE(dims->enc_ff % 64 == 0 && dims->dec_ff % 64 == 0, "kiri_create: FF width must be a multiple of 64");
  KIRI_REQUIRE(dims->img_h % 8 == 0, "kiri_create: IMG_H must be a multiple of 8");
  KiriHandle* h = new (std::nothrow) KiriHandle;
  KIRI_REQUIRE(h, "kiri_create: out of host memory");
  h->d = *dims;
  h->w = *weights;
  memcpy(h->conv1_w, weights->conv1_w_host, sizeof(h->conv1_w));
  memcpy(h->conv1_b, weights->conv1_b_host, sizeof(h->conv1_b));
  h->w.conv1_w_host = h->conv1_w;
  h->w.conv1_b_host = h->conv1_b;
  h->fused = nullptr;
  h->enc_consts = nullptr;
  h->conv1_tc_b = nullptr;
  if (conv1_tc_build(h->conv1_w, h->conv1_b, &h->conv1_tc_b) != 0) { delete h; return -1; }
  if (dims->enc_ff % 256 == 0 && dims->enc_ff <= 1024 && dims->enc_layers > 0 && weights->enc[0].wo) {
    h->enc_consts = new (std::nothrow) EbConst[dims->enc_layers];
    KIRI_REQUIRE(h->enc_consts, "kiri_create: out of host memory");
    for (int l = 0; l < dims->enc_layers; ++l) {
      const KiriEncLayerWeights& lw = weights->enc[l];
      const bool last = l + 1 == dims->enc_layers;
      const int rc = encoder_block_consts(&h->enc_consts[l], lw.bo, lw.b1, lw.b2, lw.ln2_g, lw.ln2_b,
                                          last ? nullptr : weights->enc[l + 1].ln1_g, last ? nullptr : weights->enc[l + 1].ln1_b,
                                          dims->enc_ff);
      if (rc != 0) { delete[] h->enc_consts; delete h; return rc; }
    }
  }
  if (dims->dec_layers > 0 && weights->heads_w && weights->dec[0].wqkv) {
    const int rc = fused_decoder_build(h);
    if (rc != 0) { delete h; return rc; }
  }
  *out = h;
  return 0;
}
extern "C" void kiri_destroy(KiriHandle* h) {
  if (!h) return;
  fused_decoder_free(h);
  if (h->conv1_tc_b) cudaFree(h->conv1_tc_b);
  delete[] h->enc_consts;
  delete h;
}

namespace {
struct EncodeWs {          // byte offsets into the caller's workspace
  size_t act1, act2, act3, act4, x, a, qkv, o, hbuf, total;
  size_t g1[8], g2[8], g3[8], g4[8];   // per-group offsets inside act1..act4 (sub-batch buffers / whole-group act4)
  int sc[8];                           // lines per stem sub-batch of each group
  long long m_total;       // tokens of all groups
};
inline size_t al(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

// Lines per stem sub-batch of a width group.  `stem_chunk` is a budget in 640-px lines (the stem's
// activations scale with the width), and the sub-batches of a group are balanced: 66 lines at 512 px
// run as one launch chain instead of 64 + 2 (a 2-line chain is all fixed cost).
inline int stem_sub_batch(int B, int Wb, int stem_chunk) {
  if (stem_chunk <= 0) return B;
  long long cap = static_cast<long long>(stem_chunk) * 640 / Wb;
  if (cap < stem_chunk) cap = stem_chunk;
  if (cap >= B) return B;
  const int n_chunks = static_cast<int>((B + cap - 1) / cap);
  return (B + n_chunks - 1) / n_chunks;
}

// Workspace of a multi-group encode: every group has its own stem buffers (one sub-batch each — the
// conv layers of all groups share launches), the token-stream buffers hold the concatenation of all groups.
EncodeWs plan_encode(const KiriDims& d, const KiriGroup* groups, int n_groups, int stem_chunk) {
  const int H = d.img_h;
  size_t a1 = 0, a2 = 0, a3 = 0, a4 = 0;
  long long M = 0;
  EncodeWs w;
  memset(&w, 0, sizeof(w));
  for (int g = 0; g < n_groups && g < 8; ++g) {
    const int B = groups[g].n_lines, Wb = groups[g].Wb;
    const int sc = stem_sub_batch(B, Wb, stem_chunk);
    const size_t T = Wb / 4;
    w.sc[g] = sc;
    w.g1[g] = a1; a1 += al(static_cast<size_t>(sc) * H * Wb * 48 * 2);          // conv1 output: dense 48 channels
    w.g2[g] = a2; a2 += al(static_cast<size_t>(sc) * (H / 2) * (Wb / 2) * 96 * 2);
    w.g3[g] = a3; a3 += al(static_cast<size_t>(sc) * (H / 4) * (Wb / 4) * 160 * 2);
    w.g4[g] = a4; a4 += al(static_cast<size_t>(B) * (H / 8) * T * 256 * 2);
    M += static_cast<long long>(B) * T;
  }
  size_t off = 0;
  w.act1 = off; off += al(a1);
  w.act2 = off; off += al(a2);
  w.act3 = off; off += al(a3);
  w.act4 = off; off += al(a4);
  w.x = off;    off += al(static_cast<size_t>(M) * 256 * 4);
  w.a = off;    off += al(static_cast<size_t>(M) * 256 * 2);
  w.qkv = off;  off += al(static_cast<size_t>(M) * 768 * 2);
  w.o = off;    off += al(static_cast<size_t>(M) * 256 * 2);
  w.hbuf = off; off += al(static_cast<size_t>(M) * static_cast<size_t>(d.enc_ff) * 2);
  w.total = off;
  w.m_total = M;
  return w;
}
}  // namespace

extern "C" size_t kiri_encode_workspace_bytes(const KiriHandle* h, int B, int Wb, int stem_chunk) {
  if (!h || B <= 0 || Wb <= 0) return 0;
  KiriGroup g = {nullptr, B, Wb};
  return plan_encode(h->d, &g, 1, stem_chunk).total;
}
extern "C" size_t kiri_encode_multi_workspace_bytes(const KiriHandle* h, const KiriGroup* groups, int n_groups,
                                                    int stem_chunk) {
  if (!h || !groups || n_groups <= 0) return 0;
  return plan_encode(h->d, groups, n_groups, stem_chunk).total;
}

extern "C" int kiri_encode(KiriHandle* h, const uint8_t* planes_u8, int B, int Wb, int stem_chunk,
                           void* workspace, size_t workspace_bytes, float* mem_f32, void* mem_bf16,
                           float* logits, float* tok_f32, const int* kv_len, cudaStream_t stream) {
  KIRI_REQUIRE(h && planes_u8 && workspace, "kiri_encode: null pointer");
  KiriGroup g = {planes_u8, B, Wb};
  return kiri_encode_multi(h, &g, 1, stem_chunk, workspace, workspace_bytes, mem_f32, mem_bf16, logits, tok_f32, kv_len,
                           nullptr, nullptr, stream);
}

extern "C" int kiri_encode_multi(KiriHandle* h, const KiriGroup* groups, int n_groups, int stem_chunk,
                                 void* workspace, size_t workspace_bytes, float* mem_f32, void* mem_bf16,
                                 float* logits, float* tok_f32, const int* kv_len, int* frame_ids, float* frame_prob,
                                 cudaStream_t stream) {
  KIRI_REQUIRE(h && groups && workspace && n_groups > 0, "kiri_encode_multi: null pointer");
  KIRI_REQUIRE((frame_ids == nullptr) == (frame_prob == nullptr), "kiri_encode_multi: frame_ids and frame_prob go together");
  const KiriDims& d = h->d;
  const KiriWeights& w = h->w;
  for (int g = 0; g < n_groups; ++g) {
    KIRI_REQUIRE(groups[g].planes && groups[g].n_lines > 0, "kiri_encode_multi: group %d is empty", g);
    KIRI_REQUIRE(groups[g].Wb > 0 && groups[g].Wb % 128 == 0 && groups[g].Wb / 4 <= d.max_t,
                 "kiri_encode: batch width %d must be a positive multiple of 128 and <= %d", groups[g].Wb, d.max_t * 4);
  }
  const EncodeWs ws = plan_encode(d, groups, n_groups, stem_chunk);
  KIRI_REQUIRE(workspace_bytes >= ws.total, "kiri_encode: workspace of %zu bytes given, %zu needed", workspace_bytes, ws.total);
  KIRI_REQUIRE(ws.m_total < (1ll << 31), "kiri_encode: too many tokens");
  uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
  const int H = d.img_h, D = d.enc_dim;
  const int M = static_cast<int>(ws.m_total);
  float* x = reinterpret_cast<float*>(base + ws.x);
  uint8_t* a = base + ws.a;

  KIRI_REQUIRE(n_groups <= 8, "kiri_encode_multi: at most 8 width groups");
  // ---- stem: the sub-batches of all groups advance together, and every conv layer runs as ONE launch over
  // the groups (per-group launches paid ~20 us each of prologue, tail and wave quantisation: 15 launches)
  int rounds = 0;
  for (int g = 0; g < n_groups; ++g) rounds = std::max(rounds, (groups[g].n_lines + ws.sc[g] - 1) / ws.sc[g]);
  for (int r = 0; r < rounds; ++r) {
    GemmLaunch L2[8], L3[8], L4[8];
    const uint8_t* c1_in[8];
    void* c1_out[8];
    int c1_lines[8], c1_W[8];
    int np = 0;
    for (int g = 0; g < n_groups; ++g) {
      const int B = groups[g].n_lines, Wb = groups[g].Wb, T = Wb / 4;
      const int b0 = r * ws.sc[g];
      if (b0 >= B) continue;
      const int nb = std::min(ws.sc[g], B - b0);
      uint8_t* a1 = base + ws.act1 + ws.g1[g];
      uint8_t* a2 = base + ws.act2 + ws.g2[g];
      uint8_t* a3 = base + ws.act3 + ws.g3[g];
      uint8_t* a4 = base + ws.act4 + ws.g4[g] + static_cast<size_t>(b0) * (H / 8) * T * 256 * 2;
      const uint8_t* planes = groups[g].planes + static_cast<size_t>(b0) * H * Wb;
      c1_in[np] = planes; c1_out[np] = a1; c1_lines[np] = nb; c1_W[np] = Wb;
      L2[np] = conv_launch(a1, w.conv2_w, w.conv2_b, nb, H, Wb, 64, 96, 2, 2, a2, 48);     // K = 64 per tap, 48 stored
      L3[np] = conv_launch(a2, w.conv3_w, w.conv3_b, nb, H / 2, Wb / 2, 96, 160, 2, 2, a3);
      L4[np] = conv_launch(a3, w.conv4_w, w.conv4_b, nb, H / 4, Wb / 4, 160, 256, 2, 1, a4);
      ++np;
    }
    if (np == 0) continue;
    { ProfScope ps(PS_CONV1, stream);
      // tensor-pipe form (conv1_tc.cu); KIRI_CONV1_FFMA=1: the CUDA-core form (conv1.cu)
      static const bool conv1_ffma = getenv("KIRI_CONV1_FFMA") != nullptr;
      if (h->conv1_tc_b && !conv1_ffma) KIRI_TRY(conv1_tc_launch(c1_in, c1_out, c1_lines, c1_W, np, h->conv1_tc_b, H, stream));
      else KIRI_TRY(kiri_conv1_multi(c1_in, c1_out, c1_lines, c1_W, np, w.conv1_w_host, w.conv1_b_host, H, stream)); }
    { ProfScope ps(PS_CONV2, stream);
      // channels as M, 256 pixels as N (conv2_swap.cu) when every group has the shape it is built for
      if (conv2_swap_supported(L2, np)) KIRI_TRY(launch_conv2_swap(L2, np, stream));
      else KIRI_TRY(launch_gemm_tc_multi(L2, np, stream)); }
    { ProfScope ps(PS_CONV3, stream); KIRI_TRY(launch_gemm_tc_multi(L3, np, stream)); }
    { ProfScope ps(PS_CONV4, stream); KIRI_TRY(launch_gemm_tc_multi(L4, np, stream)); }
  }
  // ---- pool + positional table + enc_ln_in (+ norm1 of layer 0) of every group, written as the concatenated
  // token stream (one launch)
  {
    const void* p_act[8];
    int p_lines[8], p_T[8];
    for (int g = 0; g < n_groups; ++g) {
      p_act[g] = base + ws.act4 + ws.g4[g];
      p_lines[g] = groups[g].n_lines;
      p_T[g] = groups[g].Wb / 4;
    }
    ProfScope ps(PS_POOL_LN, stream);
    KIRI_TRY(kiri_pool_pos_ln_multi(p_act, p_lines, p_T, n_groups, w.pos_table, H / 8, D, w.enc_ln_in_g, w.enc_ln_in_b,
                                    w.enc[0].ln1_g, w.enc[0].ln1_b, x, a, stream));
  }
  if (tok_f32) KIRI_CHECK_CUDA(cudaMemcpyAsync(tok_f32, x, static_cast<size_t>(M) * D * 4, cudaMemcpyDeviceToDevice, stream));
  // ---- encoder layers over the concatenated token stream (the GEMMs do not see line boundaries)
  for (int l = 0; l < d.enc_layers; ++l) {
    const KiriEncLayerWeights& lw = w.enc[l];
    { ProfScope ps(PS_QKV, stream);
      KIRI_TRY(gemm_call(a, lw.wqkv, lw.bqkv, M, 3 * D, D, EPI_BIAS_BF16, base + ws.qkv, nullptr, nullptr, nullptr, nullptr, stream)); }
    { ProfScope ps(PS_ATTN, stream);
      int g_lines[8], g_T[8];
      KIRI_REQUIRE(n_groups <= 8, "kiri_encode_multi: at most 8 width groups");
      for (int g = 0; g < n_groups; ++g) { g_lines[g] = groups[g].n_lines; g_T[g] = groups[g].Wb / 4; }
      KIRI_TRY(kiri_encoder_attention_multi(base + ws.qkv, base + ws.o, g_lines, g_T, n_groups, d.enc_heads, D, kv_len, stream)); }
    static const bool fused_tail = getenv("KIRI_NO_FUSED_BLOCK") == nullptr;
    if (fused_tail && h->enc_consts && M % 32 == 0) {
      // x += out_proj(o); x += FFN(norm2(x)); a = next layer's norm1(x) — one kernel (encoder_block.cu)
      ProfScope ps(PS_FF2, stream);
      const bool last = l + 1 == d.enc_layers;
      KIRI_TRY(launch_encoder_block(base + ws.o, x, last ? nullptr : a, lw.wo, lw.w1, lw.w2, &h->enc_consts[l], !last, M, d.enc_ff,
                                    stream));
      continue;
    }
    // x += out_proj(o); a = norm2(x)
    { ProfScope ps(PS_OUTPROJ, stream);
      KIRI_TRY(gemm_call(base + ws.o, lw.wo, lw.bo, M, D, D, EPI_BIAS_RESID_LN, x, x, lw.ln2_g, lw.ln2_b, a, stream)); }
    { ProfScope ps(PS_FF1, stream);
      KIRI_TRY(gemm_call(a, lw.w1, lw.b1, M, d.enc_ff, D, EPI_BIAS_GELU_BF16, base + ws.hbuf, nullptr, nullptr, nullptr, nullptr, stream)); }
    ProfScope ps_ff2(PS_FF2, stream);
    // x += linear2(h); a = next layer's norm1(x)  (last layer: enc_ln, handled below)
    if (l + 1 < d.enc_layers) {
      KIRI_TRY(gemm_call(base + ws.hbuf, lw.w2, lw.b2, M, D, d.enc_ff, EPI_BIAS_RESID_LN, x, x, w.enc[l + 1].ln1_g,
                         w.enc[l + 1].ln1_b, a, stream));
    } else {
      KIRI_TRY(gemm_call(base + ws.hbuf, lw.w2, lw.b2, M, D, d.enc_ff, EPI_BIAS_RESID_F32, x, x, nullptr, nullptr, nullptr, stream));
    }
  }
  // ---- mem = enc_ln(x); head input = ctc_head.0(mem)
  void* mem_b = mem_bf16;                              // only the decoder reads it: not written when the caller passes none
  { ProfScope ps(PS_LN_FINAL, stream);
    KIRI_TRY(kiri_layernorm(x, M, D, w.enc_ln_g, w.enc_ln_b, mem_f32, mem_b, w.ctc_ln_g, w.ctc_ln_b, a, stream)); }
  ProfScope ps_head(PS_CTC_HEAD, stream);
  const int Cp = (d.ctc_classes + 15) / 16 * 16;
  if (frame_ids && Cp <= 256) {
    // frame decisions in the GEMM epilogue; the logits are stored only when the caller wants them (beam rescoring)
    GemmLaunch L;
    memset(&L, 0, sizeof(L));
    L.a = a; L.w = w.ctc_w;
    L.NB = 1; L.IH = 1; L.IW = M; L.Cin = D; L.OH = 1; L.OW = M;
    L.sw = 1; L.sh = 1; L.pad = 0; L.kw = 1; L.kh = 1;
    L.N = Cp; L.epi = EPI_CTC_STATS;
    L.e.out = logits ? static_cast<void*>(logits) : static_cast<void*>(base);   // (any aligned base: the tensor map of a store that never happens)
    L.e.bias = w.ctc_b; L.e.ldc = Cp; L.e.n_valid = Cp;
    L.e.stat_id = frame_ids; L.e.stat_p = frame_prob; L.e.n_stat = d.ctc_classes; L.e.store_out = logits ? 1 : 0;
    KIRI_TRY(launch_gemm_tc(L, stream));
    return 0;
  }
  KIRI_REQUIRE(!frame_ids, "kiri_encode_multi: frame decisions in the head epilogue need at most 256 CTC classes (have %d)", Cp);
  if (logits)
    KIRI_TRY(gemm_call(a, w.ctc_w, w.ctc_b, M, Cp, D, EPI_BIAS_F32, logits, nullptr, nullptr, nullptr, nullptr, stream));
  return 0;
}
