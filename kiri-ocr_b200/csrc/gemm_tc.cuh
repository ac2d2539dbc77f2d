// tcgen05 / TMEM / TMA implicit-GEMM for the conv stem (3x3, NHWC bf16) and every dense layer.
//
// One kernel serves K3-K5, K8, K9, K11 of SURVEY.md §2.3:
//   D[128 x bn] (fp32, TMEM) = sum over (tap, 32-channel chunk) A[128 x 32] * B[bn x 32]^T
// * A rows are output pixels.  A tile is NSEG "segments"; one segment is an R x SEG patch of
//   output pixels fetched by ONE 5-D TMA box (c32, W, H, chunk, image) whose W/H traversal
//   strides are the conv strides and whose out-of-bounds taps are zero-filled by TMA -> the
//   3x3 halo, the stride-2 sampling and the zero padding cost no instructions.  A plain
//   [M,K] matrix is the degenerate case (1 tap, R=1, SEG=128, H=1, one image).
// * B is the [N, taps*Cin] weight matrix (K contiguous), fetched by a 3-D box (c32, N, chunk).
// * K is cut into 32-element (64-byte) chunks, SWIZZLE_64B, because the stem's channel counts
//   (64, 96, 160) are multiples of 32 but not of 64.  A pipeline stage holds CPS chunks.
// * Warp roles: warp0 = TMA producer (1 thread), warp1 = MMA issuer (1 thread) + TMEM owner,
//   warps 2..5 = epilogue (tcgen05.ld -> bias/activation/residual -> global).  Two fp32
//   accumulators (2 x 256 TMEM columns) let the epilogue of tile i overlap the MMAs of i+1.
// * Persistent: grid = min(#tiles, #SMs); tiles are taken round-robin.
#pragma once

#include "common.cuh"

namespace kiri {

enum EpiMode : int {
  EPI_BIAS_BF16 = 0,       // out_bf16 = acc + bias
  EPI_BIAS_SILU_BF16 = 1,  // out_bf16 = silu(acc + bias)          (conv + folded BN + SiLU)
  EPI_BIAS_GELU_BF16 = 2,  // out_bf16 = gelu_erf(acc + bias)      (FFN first linear)
  EPI_BIAS_RESID_F32 = 3,  // out_f32  = resid_f32 + acc + bias    (attention out-proj, FFN second)
  EPI_BIAS_F32 = 4,        // out_f32  = acc + bias                (CTC / decoder heads)
  EPI_BIAS_RESID_LN = 5,   // out_f32  = resid + acc + bias; out2_bf16 = LayerNorm(out_f32), N == 256
  EPI_CTC_STATS = 6,       // logits = acc + bias (fp32, stored only if out != null); per row: first arg-max over the
                           // n_stat valid classes and 1 / sum exp(logit - max)  (the CTC head, N <= 256: one n-tile)
};

struct ConvGeom {
  int n_seg_total;     // segments in the whole problem
  int segs_per_img;    // (OH/R) * (OW/SEG)
  int segs_per_row;    // OW / SEG
  int OH, OW;          // output spatial size (GEMM: OH=1, OW=M)
  int R, SEG;          // segment = R output rows x SEG output columns; NSEG*R*SEG == 128
  int sw, sh, pad;     // conv stride (w,h) and padding (GEMM: 1,1,0)
  int kw;              // kernel width (3 or 1)
  int taps;            // kw*kh (9 or 1)
  int chunks_per_tap;  // Cin / 32
  int cgs;             // chunk groups per tap = chunks_per_tap / CPS
  int k16;             // 16-wide K steps of a chunk that hold data (KC/16, fewer when the TMA box zero-fills the chunk's tail)
};

struct EpiParams {
  void* out;           // bf16 or fp32, row-major [rows, ldc]
  const float* bias;   // [N] (never null; pass zeros for bias-free layers)
  const float* resid;  // fp32 [rows, ldc] for EPI_BIAS_RESID_F32 (may alias out)
  int ldc;             // elements per output row
  int n_valid;         // N (columns >= n_valid are neither read from bias nor stored)
  const float* ln_g;   // EPI_BIAS_RESID_LN: LayerNorm affine [256]
  const float* ln_b;
  void* out2;          // EPI_BIAS_RESID_LN: bf16 [rows, 256]
  int timing;          // debug: accumulate role-level cycle counts of CTA 0 (set by the launcher)
  int* stat_id;        // EPI_CTC_STATS: [rows] arg-max class per row
  float* stat_p;       // EPI_CTC_STATS: [rows] soft-max probability of that class
  int n_stat;          // EPI_CTC_STATS: classes that take part (columns >= n_stat are head padding)
  int store_out;       // EPI_CTC_STATS: 1 = also store the logits
};

// Host launcher (gemm_tc.cu).  A is described by an NHWC activation [NB, IH, IW, Cin]
// (Cin % 32 == 0) or, for a plain GEMM, [1, 1, M, K]; W is [N, taps*Cin] bf16.
struct GemmLaunch {
  const void* a;       // bf16 activations
  const void* w;       // bf16 weights [N, taps*Cin]
  int NB, IH, IW, Cin; // input geometry (Cin = channels per tap in the weight matrix)
  int Cin_mem;         // channels stored per pixel (0 = Cin); < Cin: the TMA box zero-fills the rest (one chunk per tap only)
  int OH, OW;          // output geometry
  int sw, sh, pad, kw, kh;
  int N;               // output channels
  int epi;             // EpiMode
  EpiParams e;
};
int launch_gemm_tc(const GemmLaunch& L, cudaStream_t stream);
// Up to kMaxProblems conv problems of ONE layer (the width groups of a batch) in one launch (two when the
// groups need different tile forms): one prologue / tail and one wave-quantisation loss instead of five.
constexpr int kMaxProblems = 8;
int launch_gemm_tc_multi(const GemmLaunch* Ls, int n, cudaStream_t stream);
struct ProblemSet {              // grid-constant kernel parameter
  CUtensorMap tmA[kMaxProblems];
  CUtensorMap tmOut[kMaxProblems];
  ConvGeom g[kMaxProblems];
  int mtile_begin[kMaxProblems + 1];
  int n;
};
int gemm_tc_num_sms();
int gemm_tc_max_smem();

// cached cuTensorMapEncodeTiled (gemm_tc.cu)
int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
               const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapSwizzle swz,
               CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
// 2-D row-major [rows, cols] view moved in 32-row x 128-byte tiles (epilogue stores / residual loads)
int encode_rowtile_map(CUtensorMap* m, const void* base, long long rows, int cols, int ld, bool f32);

// 16-byte chunk j of row `lane` inside a 32 x 128 B tile laid out with the 128-byte swizzle that the
// TMA tensor maps of the epilogue use: conflict-free for "one thread = one row" accesses.
__device__ __forceinline__ uint32_t stg_off(int lane, int j) {
  return static_cast<uint32_t>(lane * 128 + ((j ^ (lane & 7)) << 4));
}

// erf-GELU with erf(z) ~ tanh(z (a + b z^2 + c z^4)), |err| < 4.1e-5 on the clamped range: ONE MUFU
// per element (the FFN epilogue is MUFU/issue-bound: 32 K activations per 128 x 256 tile).
__device__ __forceinline__ float gelu_tanh_erf(float v) {
  const float u = fminf(fmaxf(v * 0.70710678118654752440f, -4.5f), 4.5f);
  const float u2 = u * u;
  float p = fmaf(-0.00181363f, u2, 0.10414107f);
  p = fmaf(p, u2, 1.12812423f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u * p));
  const float h = 0.5f * v;
  return fmaf(h, t, h);
}
// the same on a pair (packed fp32 multiplies / FMAs; min, max and tanh stay scalar)
__device__ __forceinline__ float2 gelu_tanh_erf2(float2 v) {
  float2 u = fmul2(v, make_float2(0.70710678118654752440f, 0.70710678118654752440f));
  u.x = fminf(fmaxf(u.x, -4.5f), 4.5f);
  u.y = fminf(fmaxf(u.y, -4.5f), 4.5f);
  const float2 u2 = fmul2(u, u);
  float2 p = ffma2(make_float2(-0.00181363f, -0.00181363f), u2, make_float2(0.10414107f, 0.10414107f));
  p = ffma2(p, u2, make_float2(1.12812423f, 1.12812423f));
  const float2 a = fmul2(u, p);
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(a.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(a.y));
  const float2 h = fmul2(make_float2(0.5f, 0.5f), v);
  return ffma2(h, t, h);
}


}  // namespace kiri
