// tcgen05/TMEM/TMA implicit-GEMM kernel + host launcher.  See gemm_tc.cuh for the design.
#include "gemm_tc.cuh"

#include <mutex>

namespace kiri {

static constexpr int kTileM = 128;
static constexpr int kChunkBytes = 64;                  // 32 bf16 of K
static constexpr int kATileBytes = kTileM * kChunkBytes;  // one A chunk tile: 8 KiB
static constexpr int kAccStride = 256;                  // TMEM columns per accumulator
static constexpr int kTmemCols = 512;
static constexpr int kNumThreads = 192;                 // warp0 TMA, warp1 MMA, warps2-5 epilogue
static constexpr int kMaxStages = 8;

struct __align__(8) PipeBarriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

template <int EPI>
__device__ __forceinline__ void epilogue_store(const float* v, int ncols, int col0, size_t row,
                                               const EpiParams& e) {
  // v[0..ncols) are acc values for columns col0..col0+ncols of output row `row`.
  // ncols is 16 or 32; stores are predicated per 8 (bf16) / 4 (fp32) columns on n_valid.
  if (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_F32) {
    float* out = reinterpret_cast<float*>(e.out) + row * (size_t)e.ldc + col0;
    const float* res = (EPI == EPI_BIAS_RESID_F32) ? e.resid + row * (size_t)e.ldc + col0 : nullptr;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (j < ncols && col0 + j + 4 <= e.n_valid) {
        float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + col0 + j));
        float4 o = make_float4(v[j] + b.x, v[j + 1] + b.y, v[j + 2] + b.z, v[j + 3] + b.w);
        if (EPI == EPI_BIAS_RESID_F32) {
          float4 r = *reinterpret_cast<const float4*>(res + j);
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        *reinterpret_cast<float4*>(out + j) = o;
      }
    }
  } else {
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(e.out) + row * (size_t)e.ldc + col0;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      if (j < ncols && col0 + j + 8 <= e.n_valid) {
        float4 b0 = __ldg(reinterpret_cast<const float4*>(e.bias + col0 + j));
        float4 b1 = __ldg(reinterpret_cast<const float4*>(e.bias + col0 + j + 4));
        float t[8] = {v[j] + b0.x,     v[j + 1] + b0.y, v[j + 2] + b0.z, v[j + 3] + b0.w,
                      v[j + 4] + b1.x, v[j + 5] + b1.y, v[j + 6] + b1.z, v[j + 7] + b1.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (EPI == EPI_BIAS_SILU_BF16) t[q] = silu_fast(t[q]);
          if (EPI == EPI_BIAS_GELU_BF16) t[q] = gelu_erf(t[q]);
        }
        uint4 pk;
        pk.x = pack_bf16x2(t[0], t[1]);
        pk.y = pack_bf16x2(t[2], t[3]);
        pk.z = pack_bf16x2(t[4], t[5]);
        pk.w = pack_bf16x2(t[6], t[7]);
        *reinterpret_cast<uint4*>(out + j) = pk;
      }
    }
  }
}

template <int CPS, int NSEG, int EPI>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const ConvGeom g, const EpiParams e, const int bn, const int num_m_tiles,
               const int num_n_tiles, const int stages) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A: CPS*8K][B: CPS*bn*64] then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t a_stage_bytes = CPS * kATileBytes;
  const uint32_t b_chunk_bytes = bn * kChunkBytes;
  const uint32_t b_stage_bytes = CPS * b_chunk_bytes;
  const uint32_t stage_bytes = a_stage_bytes + b_stage_bytes;
  PipeBarriers* bars = reinterpret_cast<PipeBarriers*>(smem + (size_t)stages * stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = num_m_tiles * num_n_tiles;
  const int num_kb = g.taps * g.cgs;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 1);
      mbar_init(&bars->tmem_empty[a], 4);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      // one TMA box = one 32-channel chunk of one segment: R*SEG rows x 64 B
      const uint32_t box_bytes = g.R * g.SEG * kChunkBytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % num_n_tiles;
        const int m_tile = tile / num_n_tiles;
        // decode the tile's segments once
        int seg_b[NSEG], seg_x[NSEG], seg_y[NSEG];
        int nvalid = 0;
#pragma unroll
        for (int j = 0; j < NSEG; ++j) {
          const int s = m_tile * NSEG + j;
          if (s < g.n_seg_total) {
            const int b = s / g.segs_per_img;
            const int rem = s - b * g.segs_per_img;
            const int yb = rem / g.segs_per_row;
            const int xb = rem - yb * g.segs_per_row;
            seg_b[j] = b;
            seg_y[j] = yb * g.R * g.sh - g.pad;
            seg_x[j] = xb * g.SEG * g.sw - g.pad;
            ++nvalid;
          } else {
            seg_b[j] = -1; seg_x[j] = 0; seg_y[j] = 0;
          }
        }
        const uint32_t tx_bytes = nvalid * CPS * box_bytes + b_stage_bytes;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / g.cgs;
          const int cg = kb - tap * g.cgs;
          const int ky = tap / g.kw;
          const int kx = tap - ky * g.kw;
          mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* a_dst = smem + (size_t)stage * stage_bytes;
          uint8_t* b_dst = a_dst + a_stage_bytes;
          mbar_arrive_expect_tx(&bars->full[stage], tx_bytes);
#pragma unroll
          for (int c = 0; c < CPS; ++c) {
#pragma unroll
            for (int j = 0; j < NSEG; ++j) {
              if (seg_b[j] >= 0)
                tma_load_5d(a_dst + (size_t)c * kATileBytes + (size_t)j * box_bytes, &tmA,
                            &bars->full[stage], 0, cg * CPS + c, seg_x[j] + kx, seg_y[j] + ky,
                            seg_b[j]);
            }
            tma_load_3d(b_dst + (size_t)c * b_chunk_bytes, &tmB, &bars->full[stage], 0,
                        tap * g.chunks_per_tap + cg * CPS + c, n_tile * bn);
          }
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(kTileM, bn);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t b_addr = a_addr + a_stage_bytes;
#pragma unroll
          for (int c = 0; c < CPS; ++c) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint64_t ad = umma_desc_kmajor(a_addr + c * kATileBytes + h * 32, 512, UMMA_LAYOUT_SW64);
              const uint64_t bd = umma_desc_kmajor(b_addr + c * b_chunk_bytes + h * 32, 512, UMMA_LAYOUT_SW64);
              umma_bf16(d_tmem, ad, bd, idesc, (kb | c | h) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&bars->empty[stage]);       // frees the smem stage when the MMAs retire
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&bars->tmem_full[acc]);       // accumulator complete -> epilogue
      }
    }
  } else {
    // ============================ epilogue warps ============================
    const int q = warp & 3;                        // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;                   // tile row
    const int j = m / (g.R * g.SEG);               // segment within tile
    const int within = m - j * (g.R * g.SEG);
    const int jj = within / g.SEG;
    const int ii = within - jj * g.SEG;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int n_tile = tile % num_n_tiles;
      const int m_tile = tile / num_n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      // output row of this thread
      const int s = m_tile * NSEG + j;
      bool valid = s < g.n_seg_total;
      size_t row = 0;
      if (valid) {
        const int b = s / g.segs_per_img;
        const int rem = s - b * g.segs_per_img;
        const int yb = rem / g.segs_per_row;
        const int xb = rem - yb * g.segs_per_row;
        const int oy = yb * g.R + jj;
        const int ox = xb * g.SEG + ii;
        valid = (oy < g.OH) && (ox < g.OW);
        row = ((size_t)b * g.OH + oy) * g.OW + ox;
      }
      mbar_wait(&bars->tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccStride;
      const int col_base = n_tile * bn;
      if (EPI == EPI_BIAS_RESID_LN) {
        // x = resid + acc + bias is written out AND parked back in TMEM, so the row statistics
        // and the normalised bf16 copy need no second trip to global memory.  bn == N == 256.
        float* xo = reinterpret_cast<float*>(e.out) + row * (size_t)e.ldc;
        const float* xr = e.resid + row * (size_t)e.ldc;
        float sum = 0.f;
        for (int c0 = 0; c0 < 256; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int t = 0; t < 32; t += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + c0 + t));
              const float4 q4 = *reinterpret_cast<const float4*>(xr + c0 + t);
              float4 o;
              o.x = __uint_as_float(r[t]) + b.x + q4.x;
              o.y = __uint_as_float(r[t + 1]) + b.y + q4.y;
              o.z = __uint_as_float(r[t + 2]) + b.z + q4.z;
              o.w = __uint_as_float(r[t + 3]) + b.w + q4.w;
              *reinterpret_cast<float4*>(xo + c0 + t) = o;
              sum += (o.x + o.y) + (o.z + o.w);
              r[t] = __float_as_uint(o.x); r[t + 1] = __float_as_uint(o.y);
              r[t + 2] = __float_as_uint(o.z); r[t + 3] = __float_as_uint(o.w);
            }
          }
          tmem_st32(taddr + c0, r);
        }
        tmem_st_wait();
        const float mean = sum * (1.0f / 256.0f);
        float sq = 0.f;
        for (int c0 = 0; c0 < 256; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 32; ++t) { const float d = __uint_as_float(r[t]) - mean; sq = fmaf(d, d, sq); }
        }
        const float rstd = 1.0f / sqrtf(sq * (1.0f / 256.0f) + 1e-5f);
        __nv_bfloat16* ao = reinterpret_cast<__nv_bfloat16*>(e.out2) + row * (size_t)256;
        for (int c0 = 0; c0 < 256; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int t = 0; t < 32; t += 8) {
              const float4 g0 = __ldg(reinterpret_cast<const float4*>(e.ln_g + c0 + t));
              const float4 g1 = __ldg(reinterpret_cast<const float4*>(e.ln_g + c0 + t + 4));
              const float4 h0 = __ldg(reinterpret_cast<const float4*>(e.ln_b + c0 + t));
              const float4 h1 = __ldg(reinterpret_cast<const float4*>(e.ln_b + c0 + t + 4));
              const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
              const float hh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
              float y[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) y[u] = (__uint_as_float(r[t + u]) - mean) * rstd * gg[u] + hh[u];
              uint4 pk;
              pk.x = pack_bf16x2(y[0], y[1]); pk.y = pack_bf16x2(y[2], y[3]);
              pk.z = pack_bf16x2(y[4], y[5]); pk.w = pack_bf16x2(y[6], y[7]);
              *reinterpret_cast<uint4*>(ao + c0 + t) = pk;
            }
          }
        }
      } else {
      int c0 = 0;
      for (; c0 + 32 <= bn; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        tmem_ld_wait();
        if (valid) {
          float v[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) v[t] = __uint_as_float(r[t]);
          epilogue_store<EPI>(v, 32, col_base + c0, row, e);
        }
      }
      if (c0 < bn) {
        uint32_t r[16];
        tmem_ld16(taddr + c0, r);
        tmem_ld_wait();
        if (valid) {
          float v[32];
#pragma unroll
          for (int t = 0; t < 16; ++t) v[t] = __uint_as_float(r[t]);
#pragma unroll
          for (int t = 16; t < 32; ++t) v[t] = 0.f;
          epilogue_store<EPI>(v, 16, col_base + c0, row, e);
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int g_num_sms = 0;
static int g_max_smem = 0;
int gemm_tc_num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&g_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  }
  return g_num_sms;
}

static int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims,
                      const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr) {
  EncodeTiledFn fn = get_encode_fn();
  KIRI_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KIRI_REQUIRE(r == CUDA_SUCCESS,
               "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu box %u,%u,%u", (int)r,
               rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
               (unsigned long long)dims[2], box[0], box[1], box[2]);
  return 0;
}

template <int CPS, int NSEG, int EPI>
static int launch_inst(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvGeom& g,
                       const EpiParams& e, int bn, int num_m_tiles, int num_n_tiles,
                       cudaStream_t stream) {
  const int stage_bytes = CPS * kATileBytes + CPS * bn * kChunkBytes;
  const int overhead = 1024 + (int)sizeof(PipeBarriers);
  int stages = (g_max_smem - overhead) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  KIRI_REQUIRE(stages >= 2, "gemm_tc: stage of %d bytes does not fit twice in shared memory", stage_bytes);
  const int smem = stages * stage_bytes + overhead;
  auto kern = gemm_tc_kernel<CPS, NSEG, EPI>;
  static int configured = 0;
  if (configured < smem) {
    KIRI_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem));
    configured = g_max_smem;
  }
  int grid = num_m_tiles * num_n_tiles;
  if (grid > g_num_sms) grid = g_num_sms;
  kern<<<grid, kNumThreads, smem, stream>>>(tmA, tmB, g, e, bn, num_m_tiles, num_n_tiles, stages);
  KIRI_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_gemm_tc(const GemmLaunch& L, cudaStream_t stream) {
  gemm_tc_num_sms();
  KIRI_REQUIRE(L.Cin % 32 == 0, "gemm_tc: Cin=%d must be a multiple of 32", L.Cin);
  KIRI_REQUIRE(L.e.bias != nullptr && L.e.out != nullptr, "gemm_tc: bias/out must not be null");
  const int chunks = L.Cin / 32;
  const bool is_gemm = (L.kw == 1 && L.kh == 1);
  ConvGeom g;
  int NSEG = 1, CPS = 1;
  if (is_gemm) {
    KIRI_REQUIRE(L.IH == 1 && L.NB == 1 && L.OH == 1 && L.OW == L.IW, "gemm_tc: plain GEMM wants [1,1,M,K]");
    g.R = 1; g.SEG = 128;
    KIRI_REQUIRE(chunks % 2 == 0, "gemm_tc: GEMM K=%d must be a multiple of 64", L.Cin);
    CPS = 2;
  } else {
    if (L.OW % 128 == 0) { g.R = 1; g.SEG = 128; }
    else if (L.OW % 64 == 0 && L.OH % 2 == 0) { g.R = 2; g.SEG = 64; }
    else if (L.OW % 32 == 0 && L.OH % 4 == 0) { g.R = 4; g.SEG = 32; }
    else if (L.OW % 32 == 0) { g.R = 1; g.SEG = 32; NSEG = 4; }
    else { KIRI_REQUIRE(false, "gemm_tc: conv output width %d must be a multiple of 32", L.OW); }
    CPS = (NSEG == 1 && chunks <= 3) ? chunks : 1;
    KIRI_REQUIRE(L.epi == EPI_BIAS_SILU_BF16, "gemm_tc: conv path is built with the SiLU epilogue only");
  }
  g.OH = L.OH; g.OW = L.OW;
  g.sw = L.sw; g.sh = L.sh; g.pad = L.pad; g.kw = L.kw; g.taps = L.kw * L.kh;
  g.chunks_per_tap = chunks; g.cgs = chunks / CPS;
  g.segs_per_row = (L.OW + g.SEG - 1) / g.SEG;
  g.segs_per_img = ((L.OH + g.R - 1) / g.R) * g.segs_per_row;
  g.n_seg_total = L.NB * g.segs_per_img;
  const int num_m_tiles = (g.n_seg_total + NSEG - 1) / NSEG;
  int bn = (L.N + 15) / 16 * 16;
  if (bn > 256) bn = 256;
  const int num_n_tiles = (L.N + bn - 1) / bn;
  KIRI_REQUIRE(g.SEG * g.sw <= 256 && g.R * g.sh <= 256, "gemm_tc: TMA box too large");
  KIRI_REQUIRE(L.e.n_valid == L.N, "gemm_tc: n_valid must equal N");
  KIRI_REQUIRE(L.e.n_valid % ((L.epi == EPI_BIAS_F32 || L.epi == EPI_BIAS_RESID_F32 || L.epi == EPI_BIAS_RESID_LN) ? 4 : 8) == 0,
               "gemm_tc: N=%d not storable with vector stores for epilogue %d", L.N, L.epi);

  // Tensor maps keep global strides ascending; a box covers ONE 32-channel chunk, so a box lands
  // in shared memory as [rows][64 B] — exactly the K-major SWIZZLE_64B operand tile.
  // A: (c32, chunk, W, H, image)
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[5] = {32, (cuuint64_t)chunks, (cuuint64_t)L.IW, (cuuint64_t)L.IH, (cuuint64_t)L.NB};
    cuuint64_t str[4] = {64, (cuuint64_t)L.Cin * 2, (cuuint64_t)L.IW * L.Cin * 2,
                         (cuuint64_t)L.IH * L.IW * L.Cin * 2};
    cuuint32_t box[5] = {32, 1, (cuuint32_t)(g.SEG * g.sw), (cuuint32_t)(g.R * g.sh), 1};
    cuuint32_t es[5] = {1, 1, (cuuint32_t)g.sw, (cuuint32_t)g.sh, 1};
    if (encode_map(&tmA, L.a, 5, dims, str, box, es)) return -1;
  }
  {  // B: (c32, chunk, N)
    const int ktot = g.taps * L.Cin;
    cuuint64_t dims[3] = {32, (cuuint64_t)(ktot / 32), (cuuint64_t)L.N};
    cuuint64_t str[2] = {64, (cuuint64_t)ktot * 2};
    cuuint32_t box[3] = {32, 1, (cuuint32_t)bn};
    cuuint32_t es[3] = {1, 1, 1};
    if (encode_map(&tmB, L.w, 3, dims, str, box, es)) return -1;
  }

#define KIRI_LAUNCH(C, S, E) \
  return launch_inst<C, S, E>(tmA, tmB, g, L.e, bn, num_m_tiles, num_n_tiles, stream)
  if (!is_gemm) {
    if (CPS == 1 && NSEG == 1) KIRI_LAUNCH(1, 1, EPI_BIAS_SILU_BF16);
    if (CPS == 1 && NSEG == 4) KIRI_LAUNCH(1, 4, EPI_BIAS_SILU_BF16);
    if (CPS == 2 && NSEG == 1) KIRI_LAUNCH(2, 1, EPI_BIAS_SILU_BF16);
    if (CPS == 3 && NSEG == 1) KIRI_LAUNCH(3, 1, EPI_BIAS_SILU_BF16);
  } else {
    switch (L.epi) {
      case EPI_BIAS_BF16: KIRI_LAUNCH(2, 1, EPI_BIAS_BF16);
      case EPI_BIAS_SILU_BF16: KIRI_LAUNCH(2, 1, EPI_BIAS_SILU_BF16);
      case EPI_BIAS_GELU_BF16: KIRI_LAUNCH(2, 1, EPI_BIAS_GELU_BF16);
      case EPI_BIAS_RESID_F32: KIRI_LAUNCH(2, 1, EPI_BIAS_RESID_F32);
      case EPI_BIAS_F32: KIRI_LAUNCH(2, 1, EPI_BIAS_F32);
      case EPI_BIAS_RESID_LN:
        if (L.N != 256 || L.e.ldc != 256 || !L.e.ln_g || !L.e.ln_b || !L.e.out2 || !L.e.resid) break;
        KIRI_LAUNCH(2, 1, EPI_BIAS_RESID_LN);
      default: break;
    }
  }
#undef KIRI_LAUNCH
  KIRI_REQUIRE(false, "gemm_tc: no kernel instance for CPS=%d NSEG=%d epi=%d", CPS, NSEG, L.epi);
}

}  // namespace kiri
