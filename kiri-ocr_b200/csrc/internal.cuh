// Internal (non-ABI) declarations shared by api.cu and decoder.cu.
#pragma once
#include "common.cuh"
#include "gemm_tc.cuh"
#include "kiri_b200.h"

namespace kiri { struct EbConst; }
struct KiriHandle {
  KiriDims d;
  KiriWeights w;
  float conv1_w[48 * 9];
  float conv1_b[48];
  void* fused;          // fragment-packed decoder weights of the fused decode kernel (decoder_fused.cu)
  kiri::EbConst* enc_consts;   // [enc_layers] host copies of the encoder-tail constants (encoder_block.cu), or null
  void* conv1_tc_b;            // device: conv1's weights as the B operand of the tensor-pipe form (conv1_tc.cu), or null
};

namespace kiri {
// out[M,N] = epilogue(a[M,K] @ w[N,K]^T + bias) on the tcgen05 kernel (api.cu)
int gemm_call(const void* a, const void* w, const float* bias, int M, int N, int K, int epi, void* out,
              const float* resid, const float* ln_g, const float* ln_b, void* out2, cudaStream_t stream);

// encoder_block.cu: out_proj + LN + FFN + LN of one encoder layer in one kernel
struct EbConst {             // per-layer biases / LayerNorm affines, passed by value in the kernel parameters
  float bo[256], b2[256], ln_mid_g[256], ln_mid_b[256], ln_out_g[256], ln_out_b[256], b1[1024];
  int affine;                // 0: both LayerNorm affines are the identity (folded into the weights by the caller)
};
int encoder_block_consts(EbConst* out, const float* bo, const float* b1, const float* b2, const float* ln_mid_g,
                         const float* ln_mid_b, const float* ln_out_g, const float* ln_out_b, int FF);
int launch_encoder_block(const void* o, float* x, void* a_out, const void* wo, const void* w1, const void* w2,
                         const EbConst* consts_host, bool has_ln_out, int M, int FF, cudaStream_t stream);

// conv1_tc.cu: conv1 as one tcgen05 MMA pair per 128 pixels
int conv1_tc_build(const float* w_host, const float* b_host, void** bmat_dev);
int conv1_tc_launch(const uint8_t* const* planes_u8, void* const* out_bf16_nhwc48, const int* group_lines, const int* group_W,
                    int n_groups, const void* bmat_dev, int H, cudaStream_t stream);

// conv2_swap.cu: conv2 with the channels as M and 256 pixels as N of the tcgen05 instruction
bool conv2_swap_supported(const GemmLaunch* Ls, int n);
int launch_conv2_swap(const GemmLaunch* Ls, int n, cudaStream_t stream);

// decoder_fused.cu: whole-decode persistent cluster kernel
struct FusedBeam {           // beam-search mode of the fused decoder (nullptr = greedy)
  int beam; double lenp;
  int* seqbuf; float* lpbuf; // [2][n_slots][Lmax] scratch, n_slots = fused_decoder_slots(B, beam)
  double* score; int* len; int* state; int* ids; float* logp;   // outputs [B, beam(, Lmax)]
};
struct FusedLive {           // live publication of decode steps / the streaming beam rule (nullptr = none)
  int publish; int* progress; int stream_rule; int* bm_trace;
};
inline int fused_decoder_slots(int B, int beam) { const int lpc = 16 / beam; return (B + lpc - 1) / lpc * 16; }
int fused_decoder_build(KiriHandle* h);
void fused_decoder_free(KiriHandle* h);
int fused_decoder_run(KiriHandle* h, const __nv_bfloat16* crosskv, int crosskv_ld, const int* mem_row0, const int* mem_len,
                      int T, __nv_bfloat16* self_k, __nv_bfloat16* self_v, const int* len_est, const int* forced,
                      const int* line_perm, int B, int Lmax, const KiriDecodeParams* p, int* ids, int* n_out,
                      float* sum_logp, float* step_logp, float* step_prob, int* steps_max_dev, int cluster_size,
                      cudaStream_t stream, const FusedBeam* beam = nullptr, int kv_headmajor = 0,
                      const FusedLive* live = nullptr);
// token-major cross K/V [M, ld] -> per-line head-major blocks [layer][K|V][head][t][32]
int crosskv_headmajor(const __nv_bfloat16* src, __nv_bfloat16* dst, int ld, const int* mem_row0, const int* mem_len,
                      int T_uniform, int max_T, int n_lines, cudaStream_t stream);
}  // namespace kiri

namespace kiri {
// Optional per-stage CUDA-event timing (bench.py's roofline numbers): events are recorded on the
// launching stream around each stage while profiling is on.
enum ProfStage : int {
  PS_CONV1 = 0, PS_CONV2, PS_CONV3, PS_CONV4, PS_POOL_LN, PS_QKV, PS_ATTN, PS_OUTPROJ, PS_FF1, PS_FF2,
  PS_LN_FINAL, PS_CTC_HEAD, PS_DEC_CROSSKV, PS_DEC_STEP, PS_PREPROCESS, PS_CTC_GREEDY, PS_COUNT
};
void prof_begin(int stage, cudaStream_t s);
void prof_end(int stage, cudaStream_t s);
struct ProfScope {
  int st; cudaStream_t s;
  ProfScope(int stage, cudaStream_t stream) : st(stage), s(stream) { prof_begin(st, s); }
  ~ProfScope() { prof_end(st, s); }
};
}  // namespace kiri

#define KIRI_TRY(expr)            \
  do {                            \
    int _rc = (expr);             \
    if (_rc != 0) return _rc;     \
  } while (0)
