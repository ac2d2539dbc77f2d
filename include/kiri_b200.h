/* kiri_b200.h — C ABI of libkiri_b200.so: B200 (sm_100a) kernels for kiri-ocr's batched
 * text-line recognition path.
 *
 * The reference (mrrtmob/kiri-ocr, pure Python/PyTorch) has no FFI of its own; every entry
 * point below replaces a *call site* of the reference's hot path and cites it.  A Python
 * binding (ctypes) is what the reference-side maintainer adds — see INTEGRATION.md.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller unless the name ends in _host;
 *  - every function is asynchronous on `stream`, allocates no device memory, returns 0 on
 *    success and a negative code on failure (kiri_last_error() gives the thread-local text);
 *  - bf16 tensors are passed as void*; "NHWC" activations are [lines, rows, cols, channels];
 *  - handles are not thread-safe; use one per (device, stream).
 *  - buffers read with 32-bit loads (kiri_preprocess_pack `src`) need 4 readable bytes of
 *    slack after their last byte.
 */
#ifndef KIRI_B200_H_
#define KIRI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define KIRI_DTYPE_F32 0
#define KIRI_DTYPE_BF16 1
#define KIRI_MAX_LAYERS 8

/* epilogues of kiri_gemm_bf16 / kiri_conv3x3_bf16 */
#define KIRI_EPI_BIAS_BF16 0      /* out_bf16 = acc + bias                      */
#define KIRI_EPI_BIAS_SILU_BF16 1 /* out_bf16 = silu(acc + bias)                */
#define KIRI_EPI_BIAS_GELU_BF16 2 /* out_bf16 = gelu_erf(acc + bias)            */
#define KIRI_EPI_BIAS_RESID_F32 3 /* out_f32  = resid + acc + bias              */
#define KIRI_EPI_BIAS_F32 4       /* out_f32  = acc + bias                      */
#define KIRI_EPI_BIAS_RESID_LN 5  /* out_f32 = resid + acc + bias; out2_bf16 = LayerNorm(out_f32) (N = 256) */

const char* kiri_last_error(void);
int kiri_version(void);
/* sizeof of {KiriCropDesc, KiriDims, KiriWeights, KiriGroup, KiriDecodeParams, KiriEncLayerWeights} as this library
 * was built: a binding compares them with its own struct layouts before the first call (returns 6). */
int kiri_abi_sizes(int* out, int n);
/* 1 when the current device is compute capability 10.x (the kernels are sm_100a-only). */
int kiri_device_ok(void);

/* Per-stage CUDA-event timing of kiri_encode / kiri_decode_greedy (used by bench.py for the roofline
 * numbers).  kiri_profile_begin() starts recording and returns the number of stages;
 * kiri_profile_end() synchronises the device and returns, per stage, the summed milliseconds and
 * the number of timed intervals.  Stage order: conv1, conv2, conv3, conv4, pool_ln, qkv, attention,
 * out_proj, ff1, ff2, ln_final, ctc_head, dec_crosskv, dec_step, preprocess, ctc_greedy. */
int kiri_profile_begin(void);
int kiri_profile_end(double* ms_by_stage_host, int* count_by_stage_host, int n);

/* ---------------------------------------------------------------- K1: preprocessing
 * Replaces OCR._preprocess_region (kiri_ocr/core.py:489-528) + ResizeKeepRatioPadNoCrop /
 * preprocess_pil (kiri_ocr/model.py:316-339), i.e. numpy slicing + Pillow BILINEAR + pad,
 * done per line on the CPU by the reference.  Bit-exact with Pillow's fixed-point resample. */
typedef struct {
  int64_t src_offset; /* byte offset of the crop's top-left pixel inside `src`            */
  int64_t out_offset; /* ELEMENT offset of the crop's [img_h, Wb] plane inside the outputs */
  int32_t pitch;      /* bytes between source rows                                         */
  int32_t w, h;       /* crop size after the reference's clamp-pad (core.py:510-515)       */
  int32_t nw;         /* max(1, round(w * img_h / h)) — Python round (model.py:321-322)    */
  int32_t Wb;         /* batch width of the crop's group: the plane is cropped/padded to it */
  int32_t strip_w;    /* output columns resampled by one CTA (fits the shared-memory budget)  */
  int32_t flags;      /* KIRI_CROP_NO_INVERT: skip the dark-background inversion test (the input is an already
                         preprocessed plane: OCR.recognize_region never inverts, core.py:530-568)            */
  int32_t reserved;
} KiriCropDesc;
#define KIRI_CROP_NO_INVERT 1

/* shared memory needed by one crop for a given strip width (host helper, no GPU call) */
int kiri_preprocess_smem_bytes(int w, int h, int nw, int img_h, int Wb, int strip_w);

/* planes_u8: uint8 buffer holding every crop's [img_h, Wb] plane at its out_offset (all width
 * groups of a batch go in ONE launch); norm_bf16 (nullable): same offsets, (v/255-0.5)/0.5.
 * One CTA resamples one strip of strip_w output columns of one crop; max_strips >= the largest
 * ceil(min(nw, Wb) / strip_w) over the crops (grid = n_crops x max_strips).  smem_bytes >= the
 * largest kiri_preprocess_smem_bytes() of the crops; anything above lets a CTA stage more source
 * rows per phase (all of them for ordinary lines).  A first small launch sums every crop for the
 * reference's dark-background inversion test (core.py:524). */
int kiri_preprocess_pack(const uint8_t* src, const KiriCropDesc* descs, int n_crops, int img_h,
                         int smem_bytes, int max_strips, uint8_t* planes_u8, void* norm_bf16,
                         unsigned long long* crop_sums_scratch /* device, n_crops entries */, cudaStream_t stream);

/* Page ingest on the device: interleaved BGR uint8 [n_pixels, 3] -> gray uint8 [n_pixels], bit-exact with
 * cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) (kiri_ocr/core.py:762-766; OpenCV's 15-bit fixed-point BT.601 weights:
 * (B*3735 + G*19235 + R*9798 + 16384) >> 15).  Both buffers 4-byte aligned. */
int kiri_bgr_to_gray(const uint8_t* bgr_u8, long long n_pixels, uint8_t* gray_u8, cudaStream_t stream);

/* ---------------------------------------------------------------- K2: stem layer 1
 * Replaces ConvStem.net[0:3] (kiri_ocr/model.py:215-217).  w_host[48*9], b_host[48]: BN-folded
 * fp32 weights in HOST memory (they travel as kernel parameters).  out: DENSE NHWC bf16 with 48 channels
 * (96 bytes per pixel).  W must be a multiple of 128, H even. */
int kiri_conv1(const uint8_t* planes_u8, const float* w_host, const float* b_host, int n_lines, int H,
               int W, void* out_bf16_nhwc48, cudaStream_t stream);
/* The same for several width groups in ONE launch: planes / out / lines / W per group (host arrays). */
int kiri_conv1_multi(const uint8_t* const* planes_u8, void* const* out_bf16_nhwc48, const int* group_lines,
                     const int* group_W, int n_groups, const float* w_host, const float* b_host, int H,
                     cudaStream_t stream);
/* The same layer on the tensor pipe (csrc/conv1_tc.cu): one tcgen05 MMA pair per 128 pixels on EXACT bf16 operands
 * (u = v - 128, weights split into two bf16 terms, bias and normalisation folded into the operand), SiLU + store on the
 * CUDA cores.  Same inputs, outputs and layout as kiri_conv1_multi; group widths are multiples of 128.  kiri_encode_multi
 * uses this form (KIRI_CONV1_FFMA=1 selects the CUDA-core one).  This stand-alone entry packs and uploads the weights on
 * every call and synchronises the stream. */
int kiri_conv1_tc_multi(const uint8_t* const* planes_u8, void* const* out_bf16_nhwc48, const int* group_lines,
                        const int* group_W, int n_groups, const float* w_host, const float* b_host, int H,
                        cudaStream_t stream);

/* ---------------------------------------------------------------- K3-K5, K8, K9, K11: tcgen05 GEMMs
 * 3x3 conv as implicit GEMM (replaces ConvStem.net[3:12], kiri_ocr/model.py:218-226): input NHWC
 * bf16 [n, IH, IW, cin_mem], weights bf16 [N, 9*Cin] ordered (ky, kx, cin) with Cin % 32 == 0, pad 1,
 * stride (sh, sw); output NHWC bf16 [n, OH, OW, N] = silu(conv + bias).  cin_mem <= Cin is the number of channels
 * actually stored per pixel (0 = Cin): the TMA unit zero-fills channels cin_mem..Cin-1 in shared memory
 * (conv2 reads conv1's dense 48-channel output with Cin = 64). */
int kiri_conv3x3_bf16(const void* in_nhwc, const void* w, const float* bias, int n, int IH, int IW,
                      int Cin, int N, int sh, int sw, void* out_nhwc, int cin_mem, cudaStream_t stream);
/* out[M, N] = epilogue(a[M, K] @ w[N, K]^T + bias) — replaces the nn.Linear call sites of the
 * encoder / CTC head / decoder (kiri_ocr/model.py:246-297).  K % 64 == 0.  `resid` (fp32 [M, N])
 * may alias `out`.  For KIRI_EPI_BIAS_RESID_LN: N == 256, out2 = bf16 [M, 256]. */
int kiri_gemm_bf16(const void* a, const void* w, const float* bias, int M, int N, int K, int epi,
                   void* out, const float* resid, const float* ln_g, const float* ln_b, void* out2,
                   cudaStream_t stream);
/* plain CUDA-core reference GEMM (fp32 accumulate) used by the tests to cross-check the
 * tensor-core kernel on the device: out_f32[M,N] = a[M,K] @ w[N,K]^T. */
int kiri_gemm_ref(const void* a, const void* w, int M, int N, int K, float* out_f32, cudaStream_t stream);

/* ---------------------------------------------------------------- K6/K7 + LayerNorms
 * mean over RH stem rows + positional table + LayerNorm (+ optional second LayerNorm):
 * replaces PosEnc2D / adaptive_avg_pool2d / permute / enc_ln_in (kiri_ocr/model.py:194-208, 302-304). */
int kiri_pool_pos_ln(const void* act_bf16, const float* pos_table, int n_lines, int RH, int T, int D,
                     const float* g0, const float* b0, const float* g1, const float* b1, float* x_f32,
                     void* a_bf16, cudaStream_t stream);
/* The same for several width groups of one token stream in ONE launch (group order = token order). */
int kiri_pool_pos_ln_multi(const void* const* act_bf16, const int* group_lines, const int* group_T, int n_groups,
                           const float* pos_table, int RH, int D, const float* g0, const float* b0, const float* g1,
                           const float* b1, float* x_f32, void* a_bf16, cudaStream_t stream);
/* y = LN0(x) -> y_f32 / y_bf16 (nullable); z_bf16 = LN1(y) (nullable).  D must be 256. */
int kiri_layernorm(const float* x, int n_tok, int D, const float* g0, const float* b0, float* y_f32,
                   void* y_bf16, const float* g1, const float* b1, void* z_bf16, cudaStream_t stream);

/* ---------------------------------------------------------------- K8: encoder self-attention
 * Replaces the SDPA inside nn.TransformerEncoderLayer (kiri_ocr/model.py:246-261).
 * qkv: bf16 [n_lines*T, 3*D] rows = [Q | K | V]; out: bf16 [n_lines*T, D]; head_dim 32;
 * T in {32,64,96,128,160}; kv_len (nullable): per-line number of valid keys. */
int kiri_encoder_attention(const void* qkv_bf16, void* out_bf16, int n_lines, int T, int heads, int D,
                           const int* kv_len, cudaStream_t stream);
/* The same for a concatenated token stream of several width groups in ONE launch: group g owns
 * group_lines[g] lines of group_T[g] tokens, rows and kv_len entries in group order (host arrays,
 * at most 8 groups). */
int kiri_encoder_attention_multi(const void* qkv_bf16, void* out_bf16, const int* group_lines, const int* group_T,
                                 int n_groups, int heads, int D, const int* kv_len, cudaStream_t stream);

/* ---------------------------------------------------------------- K8b: fused encoder-layer tail
 * Replaces out_proj + residual, norm2, linear1 + GELU, linear2 + residual and the next layer's norm1
 * of nn.TransformerEncoderLayer (kiri_ocr/model.py:246-261, norm_first=True, activation="gelu") in
 * ONE kernel; the 1024-wide hidden activation and norm2's output never reach HBM.
 *   x[M,256] fp32 (in/out) += o[M,256] @ wo[256,256]^T + bo;  a2 = LN(x; ln_mid);
 *   x += gelu(a2 @ w1[FF,256]^T + b1) @ w2[256,FF]^T + b2;    a_out[M,256] bf16 = LN(x; ln_out)
 * ln_out_g/ln_out_b/a_out may be NULL together (last layer).  M % 32 == 0, FF % 256 == 0.
 * When every gain is exactly 1 and every shift exactly 0 the kernel variant without per-column LayerNorm parameters
 * runs (~8 % faster): callers fold the affines into w1/b1 and into the consumer of a_out, as kiri-ocr_b200/weights.py does. */
int kiri_encoder_block(const void* o_bf16, float* x_f32, void* a_out_bf16, const void* wo, const float* bo,
                       const void* w1, const float* b1, const void* w2, const float* b2, const float* ln_mid_g,
                       const float* ln_mid_b, const float* ln_out_g, const float* ln_out_b, int M, int FF,
                       cudaStream_t stream);
/* The same, `iters` launches back to back on x in place (constants fetched once; soak / timing tool). */
int kiri_encoder_block_soak(const void* o_bf16, float* x_f32, void* a_out_bf16, const void* wo, const float* bo,
                            const void* w1, const float* b1, const void* w2, const float* b2, const float* ln_mid_g,
                            const float* ln_mid_b, const float* ln_out_g, const float* ln_out_b, int M, int FF,
                            int iters, cudaStream_t stream);

/* ---------------------------------------------------------------- K10: fused CTC greedy
 * Replaces compute_ctc_confidence + the id-level part of CharTokenizer.decode_ctc
 * (kiri_ocr/model.py:343-373, 109-119).  logits: [n_lines, T, ld] (first C columns valid).
 * ids: [n_lines, T] collapsed ids (repeats removed, ids >= 2), n_ids[n_lines] = their count
 * (= the reference's length estimate), conf[n_lines] = mean over frames of the max soft-max
 * probability.  frame_ids / frame_prob ([n_lines, T], nullable) serve the streaming API. */
int kiri_ctc_greedy(const void* logits, int logits_dtype, int n_lines, int T, int C, int ld, int* ids,
                    int* n_ids, float* conf, int* frame_ids, float* frame_prob, cudaStream_t stream);

/* The same over a concatenated token stream: line b owns logits rows [row0[b], row0[b] + len[b])
 * (device int arrays); ids / frame_ids / frame_prob are flat [M_total] and line b writes at row0[b]. */
int kiri_ctc_greedy_multi(const void* logits, int logits_dtype, int n_lines, const int* row0, const int* len,
                          int max_T, int C, int ld, int* ids, int* n_ids, float* conf, int* frame_ids,
                          float* frame_prob, cudaStream_t stream);

/* The collapse stage alone (kiri_ocr/model.py:109-124, 366-371) for frame decisions taken in the CTC head's GEMM epilogue
 * (kiri_encode_multi with frame_ids / frame_prob: arg-max class and its soft-max probability per token, the logits stay
 * on chip): line b owns tokens [row0[b], row0[b] + len[b]); outputs as kiri_ctc_greedy_multi. */
int kiri_ctc_collapse_multi(const int* frame_ids, const float* frame_prob, int n_lines, const int* row0, const int* len,
                            int* ids, int* n_ids, float* conf, cudaStream_t stream);

/* Multi-GPU exchange payload (the one collective of the path): fixed-stride int32 records
 * {n_ids, confidence bits, ids[T]} per line, built from the token-major output of kiri_ctc_greedy_multi
 * (line b's ids start at ids[mem_row0[b]]; entries beyond n_ids are zero).  records: [n_lines, 2 + T]. */
int kiri_pack_records(const int* ids, const int* n_ids, const float* conf, const int* mem_row0, int n_lines, int T,
                      int* records, cudaStream_t stream);

/* ---------------------------------------------------------------- model-level handle
 * Packed weights (produced once per checkpoint by kiri_ocr_b200/weights.py): BN folded into the conv
 * weights, Linear weights [out, in] in bf16, biases / LayerNorm affines in fp32. */
typedef struct {
  int32_t img_h;        /* 48 */
  int32_t enc_dim;      /* 256 */
  int32_t enc_layers, enc_heads, enc_ff;
  int32_t dec_dim, dec_layers, dec_heads, dec_ff;
  int32_t ctc_classes;  /* V + 2 */
  int32_t dec_vocab;    /* V + 3 */
  int32_t max_pos;      /* rows of dec_pe */
  int32_t max_t;        /* rows of pos_table (IMG_W / 4) */
  int32_t has_dec_pos;  /* 0 for old checkpoints without dec_pos_enc.pe (core.py:255-263) */
} KiriDims;

typedef struct {
  const void* wqkv; const float* bqkv;   /* [3D, D], [3D] */
  const void* wo;   const float* bo;     /* [D, D],  [D]  */
  const void* w1;   const float* b1;     /* [FF, D], [FF] */
  const void* w2;   const float* b2;     /* [D, FF], [D]  */
  const float* ln1_g; const float* ln1_b;
  const float* ln2_g; const float* ln2_b;
} KiriEncLayerWeights;

typedef struct {
  const void* wqkv; const float* bqkv;   /* self-attention in_proj */
  const void* wo;   const float* bo;
  const void* wcq;  const float* bcq;    /* cross-attention query rows of multihead_attn.in_proj */
  const void* wco;  const float* bco;
  const void* w1;   const float* b1;
  const void* w2;   const float* b2;
  const float* ln1_g; const float* ln1_b;
  const float* ln2_g; const float* ln2_b;
  const float* ln3_g; const float* ln3_b;
} KiriDecLayerWeights;

typedef struct {
  const float* conv1_w_host; const float* conv1_b_host;   /* HOST: [48*9], [48] */
  const void* conv2_w; const float* conv2_b;               /* [96, 9*64]  */
  const void* conv3_w; const float* conv3_b;               /* [160, 9*96] */
  const void* conv4_w; const float* conv4_b;               /* [256, 9*160]*/
  const float* pos_table;                                   /* [max_t, 256] */
  const float* enc_ln_in_g; const float* enc_ln_in_b;
  KiriEncLayerWeights enc[KIRI_MAX_LAYERS];
  const float* enc_ln_g; const float* enc_ln_b;
  const float* ctc_ln_g; const float* ctc_ln_b;
  const void* ctc_w; const float* ctc_b;                   /* [Cp, 256], [Cp]: Cp = roundup(C, 16), zero padded */
  /* decoder */
  const void* crosskv_w; const float* crosskv_b;           /* [dec_layers*2*D, 256]: (W_k;W_v)_l @ W_memproj */
  const float* dec_emb;                                     /* fp32 [Vd, D] */
  const float* dec_pe;                                      /* fp32 [max_pos, D] */
  KiriDecLayerWeights dec[KIRI_MAX_LAYERS];
  const float* dec_ln_g; const float* dec_ln_b;
  const void* heads_w; const float* heads_b;               /* [2*Vp, D], [2*Vp]: dec_head rows then lm_head rows, Vp = roundup(Vd, 16) */
} KiriWeights;

typedef struct KiriHandle KiriHandle;
int kiri_create(const KiriDims* dims, const KiriWeights* weights, KiriHandle** out);
void kiri_destroy(KiriHandle* h);

/* scratch bytes kiri_encode needs for a batch of B lines of width Wb, processing the stem in
 * sub-batches of `stem_chunk` lines (so its activations stay L2-resident; 0 = whole batch). */
size_t kiri_encode_workspace_bytes(const KiriHandle* h, int B, int Wb, int stem_chunk);

/* Stem + encoder + CTC head for B preprocessed planes ([B, img_h, Wb] uint8): replaces
 * KiriOCR.encode + ctc_head (kiri_ocr/model.py:299-307, 264-268) as called from
 * OCR.recognize_region (kiri_ocr/core.py:546-552).
 *   mem_f32   (nullable) fp32 [B*T, D]   encoder output ("mem")
 *   mem_bf16  (nullable) bf16 [B*T, D]   same, for kiri_dec_prepare
 *   logits    (nullable) fp32 [B*T, Cp]  CTC logits, Cp = roundup(C, 16); columns >= C are zero
 *   tok_f32   (nullable) fp32 [B*T, D]   encoder input tokens after enc_ln_in (stage parity)
 *   kv_len    (nullable) int  [B]        valid frames per line (masked bucketed mode) */
int kiri_encode(KiriHandle* h, const uint8_t* planes_u8, int B, int Wb, int stem_chunk, void* workspace,
                size_t workspace_bytes, float* mem_f32, void* mem_bf16, float* logits, float* tok_f32,
                const int* kv_len, cudaStream_t stream);

/* The same for several width groups at once (bucketed mode): every group runs
 * its own stem, but the encoder layers and the CTC head run ONCE over the concatenation of all
 * groups' tokens (a GEMM does not see line boundaries), so small groups do not pay a launch and a
 * partial wave per layer.  `groups` is a HOST array.  Outputs are token-major and concatenated in
 * group order: group g owns rows [sum_{i<g} n_lines_i * Wb_i / 4, ...).  kv_len (nullable) has one
 * entry per line in the same order. */
typedef struct {
  const uint8_t* planes; /* device: [n_lines, img_h, Wb] uint8 */
  int32_t n_lines;
  int32_t Wb;
} KiriGroup;
size_t kiri_encode_multi_workspace_bytes(const KiriHandle* h, const KiriGroup* groups_host, int n_groups,
                                         int stem_chunk);
int kiri_encode_multi(KiriHandle* h, const KiriGroup* groups_host, int n_groups, int stem_chunk, void* workspace,
                      size_t workspace_bytes, float* mem_f32, void* mem_bf16, float* logits, float* tok_f32,
                      const int* kv_len, int* frame_ids, float* frame_prob, cudaStream_t stream);
/* frame_ids / frame_prob (nullable together, [M_total]): the CTC head's GEMM epilogue takes the per-token arg-max class
 * (first maximum over the C real classes) and its soft-max probability straight from the accumulator
 * (compute_ctc_confidence, kiri_ocr/model.py:355-363); with logits == NULL the logits are never written to HBM
 * (decode_method "fast" / "accurate" need only these two; "beam" also asks for the logits). */

/* ---------------------------------------------------------------- greedy attention decoder
 * Replaces beam_decode_one_batched at BEAM=1 (kiri_ocr/model.py:390-600 via core.py:560-568)
 * and the token rule of greedy_decode_streaming (model.py:779-946), batched over lines with
 * a per-layer self-attention KV cache and cross K/V computed once per batch. */
typedef struct {
  float lm_alpha;              /* LM_FUSION_ALPHA if fusion is on, else 0 */
  float eos_bias, eos_boost;   /* EOS_LOGP_BIAS / EOS_LOGP_BOOST */
  int32_t eos_bias_until_len;  /* EOS_BIAS_UNTIL_LEN */
  float rep_last, rep_bigram, rep_trigram, unk_penalty;
  int32_t unk_id;              /* decoder id of <unk> */
  double len_ratio;            /* DEC_MAX_LEN_RATIO */
  int32_t len_pad;             /* DEC_MAX_LEN_PAD */
  double mem_ratio;            /* MEM_MAX_LEN_RATIO */
  int32_t max_dec_len;         /* MAX_DEC_LEN */
  int32_t select_raw;          /* 0: arg-max of fused+penalised log-prob (model.py:537);
                                  1: arg-max of raw dec_head soft-max (streaming, model.py:915-917) */
} KiriDecodeParams;

size_t kiri_decode_workspace_bytes(const KiriHandle* h, int B, int T, int Lmax);
/* mem_bf16 [B*T, D]; len_est [B] (from kiri_ctc_greedy n_ids).  Outputs: ids [B, Lmax] chosen
 * tokens (EOS included when emitted), n_out [B], sum_logp [B] (sum of chosen penalised log-probs),
 * step_logp / step_prob [B, Lmax] (nullable).  Runs until every line has emitted EOS or reached
 * its own max_steps; `max_steps_cap` (<= Lmax) bounds the loop.  The call synchronises the
 * stream every `poll_every` steps to read the alive counter. */
int kiri_decode_greedy(KiriHandle* h, const void* mem_bf16, const int* len_est, int B, int T, int Lmax,
                       const KiriDecodeParams* p, void* workspace, size_t workspace_bytes, int* ids,
                       int* n_out, float* sum_logp, float* step_logp, float* step_prob,
                       const int* forced_ids, int* steps_run_host, int poll_every, cudaStream_t stream);
/* The same over the concatenated token stream of kiri_encode_multi: line b attends to the memory
 * rows [mem_row0[b], mem_row0[b] + mem_len[b]) of mem_bf16 [M_total, D] (device int arrays;
 * max_T >= every mem_len).  The cross K/V of a line are re-laid head-major once per batch so the
 * per-step single-query attention reads contiguous [t][32] runs.
 * line_perm (nullable, device int[n_slots]): decode slot -> line, or -1 for an EMPTY slot.  Sixteen consecutive slots
 * share one thread-block cluster, so listing the lines by decreasing len_est lets every cluster stop early and lets the
 * clusters that do not fit in the first wave hide behind the longest ones; a step of a cluster costs a fixed part plus
 * a part per live line, so leaving slots of the clusters that hold the LONGEST lines empty (n_slots > B) shortens the
 * decode's critical path.  n_slots = 0: one slot per line.  kiri_decode_multi_workspace_bytes takes the SLOT count. */
size_t kiri_decode_multi_workspace_bytes(const KiriHandle* h, int B, long long M_total, int Lmax);
int kiri_decode_greedy_multi(KiriHandle* h, const void* mem_bf16, long long M_total, const int* mem_row0,
                             const int* mem_len, int max_T, const int* len_est, const int* line_perm, int n_slots, int B,
                             int Lmax, const KiriDecodeParams* p,
                             void* workspace, size_t workspace_bytes, int* ids, int* n_out, float* sum_logp,
                             float* step_logp, float* step_prob, const int* forced_ids, int* steps_run_host,
                             int* progress, int publish, cudaStream_t stream);
/* Live streaming (kiri_ocr/core.py:887-1026, model.py:779-946): `progress` (nullable, [B]) receives after EVERY decode
 * step of a line   steps_available | (1 << 30 once the line has ended)   ; with publish = 1 the step's records (ids,
 * step_logp, step_prob) are made visible system-wide first (__threadfence_system), so ids / step_* / progress may be
 * MAPPED PINNED HOST memory that a host thread polls while the kernel is still decoding. */

/* ---------------------------------------------------------------- beam search (decode_method="beam")
 * Replaces beam_decode_one_batched at BEAM > 1 (kiri_ocr/model.py:390-600): `beam` (<= 5) hypotheses
 * per line, top-`beam` expansion of every alive hypothesis, stable prune by the length-normalised
 * score ((5+L)/6)^lenp with finished hypotheses listed first (model.py:550-559).  The K/V cache is
 * never copied: every hypothesis keeps a per-position table of the physical slot its ancestor
 * wrote.  Outputs per line and hypothesis: bm_state (0 none, 1 alive at the step limit, 2 ended
 * with EOS), bm_score (sum of the chosen penalised log-probs, double like the reference's Python
 * float), bm_len (tokens after BOS), bm_ids / bm_logp [B, beam, Lmax].  The final CTC-fused
 * ranking (model.py:562-579) is done by the caller with kiri_ctc_align_score. */
size_t kiri_decode_beam_workspace_bytes(const KiriHandle* h, int B, long long M_total, int Lmax, int beam);
int kiri_decode_beam_multi(KiriHandle* h, const void* mem_bf16, long long M_total, const int* mem_row0,
                           const int* mem_len, int max_T, const int* len_est, const int* line_perm, int B, int Lmax, int beam,
                           double lenp, const KiriDecodeParams* p, void* workspace, size_t workspace_bytes,
                           double* bm_score, int* bm_len, int* bm_state, int* bm_ids, float* bm_logp,
                           int stream_rule, int* bm_trace, int* progress, int publish, cudaStream_t stream);
/* stream_rule = 1 selects beam_decode_streaming's variant (kiri_ocr/model.py:949-1152): hypotheses are pruned by
 * score / L^lenp and a line stops as soon as its BEST hypothesis has ended.  bm_trace (nullable, [B, Lmax, beam, 3]
 * int32) records for every step and kept hypothesis, in rank order: {rank of its parent in the previous step, appended
 * token (-1: a finished hypothesis carried over), bits of the token's penalised log-prob} (parent -1 = slot unused), from
 * which the caller rebuilds the best partial hypothesis of every step; progress / publish as for the greedy decoder. */
/* K13: CTC forward-algorithm score of every hypothesis = compute_ctc_alignment_score
 * (kiri_ocr/model.py:603-668).  logits: fp32 [M_total, ld] CTC logits (first C columns valid); line b
 * owns rows [mem_row0[b], +mem_len[b]); max_T >= every mem_len.  out: [n_lines, beam] fp32. */
int kiri_ctc_align_score(const float* logits, int ld, int C, const int* mem_row0, const int* mem_len, int n_lines,
                         int beam, int Lmax, const int* bm_ids, const int* bm_len, const int* bm_state,
                         int vocab_size, int unk_ctc_id, int max_T, float* out, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* KIRI_B200_H_ */
