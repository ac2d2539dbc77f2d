// K3 with the operands swapped: second stem layer Conv(48->96, 3x3, s2, p1, no bias) + BN(eval) + SiLU with the CHANNELS
// as the M of the tcgen05 instruction and 256 output PIXELS as its N.
//
// Replaces  ConvStem.net[3:6]   kiri_ocr/model.py:218-220   (same contract as the conv2 launch of gemm_tc.cu: conv1's dense
//           48-channel NHWC bf16 activation in, NHWC bf16 [n, H/2, W/2, 96] out, BN folded, weights [96][9 x 64] bf16).
//
// Why: one tcgen05.mma (M = 128, K = 16) costs ~100 cycles that no N amortises (profiles/r02_mma_rate.txt: 168 cycles at
// N = 256, 117 at N = 96; 152 measured inside the pipeline).  With pixels as M and the 96 output channels as N the layer
// issues 27 instructions per 128 pixels = 4.1 k tensor cycles; with the channels as M (96 of 128 rows used) and 256 pixels
// as N it issues 27 per 256 pixels = 4.5 k: 1.8 x fewer tensor cycles per pixel.  The accumulator then holds
// [channel][pixel]; the epilogue transposes it back to NHWC through a 2 KB tile per warp.
//
// Operands: A = weights of one tap, [96 rows + 32 rows of whatever follows][64 K] K-major / 128-byte swizzle, all nine taps
// resident (108 KB); B = the activation tile of one tap, [256 pixels][64 K] fetched by ONE 5-D TMA box whose W / H element
// strides are the conv stride and whose out-of-image taps and channels 48..63 are zero-filled (as in gemm_tc.cu), three
// 32 KB stages.  Only the three K steps that hold channels are issued.  TMEM: two accumulators of 256 columns.
// Tile = 4 output rows x 64 output columns of one image (OH % 4 == 0, OW % 64 == 0: every bucket width of the recogniser).
// Roles (320 threads): warps 0-7 epilogue (lane quarter = 32 channels, warp >> 2 = pixel half), warp 8 TMA, warp 9 MMA.
#include "internal.cuh"

#include <cstdlib>
#include <cstring>

namespace kiri {
namespace {

constexpr int kCo = 96, kTaps = 9, kStages = 3;
constexpr int kTileR = 4, kTileS = 64, kTilePx = kTileR * kTileS;          // 256 pixels
constexpr int kWTapBytes = kCo * 128;                                      // 12 KB per tap
constexpr int kWBytes = kTaps * kWTapBytes + 4096;                         // + the 32 rows the M = 128 instruction reads past tap 8
constexpr int kXBytes = kTilePx * 128;                                     // 32 KB per stage
constexpr int kOutWarpBytes = 32 * 64;                                     // [32 pixels][32 channels] bf16
constexpr int kEpiWarps = 8, kTmaWarp = 8, kMmaWarp = 9, kThreads = 10 * 32;
constexpr int kXOff = kWBytes, kOutOff = kXOff + kStages * kXBytes, kBarOff = kOutOff + kEpiWarps * kOutWarpBytes;

struct SwapBars {
  uint64_t full[kStages], empty[kStages];
  uint64_t tmem_full[2], tmem_empty[2];
  uint64_t w_full;
  uint32_t tmem_base;
  uint32_t pad;
};
constexpr int kSmem = kBarOff + static_cast<int>(sizeof(SwapBars));

struct SwapGroups {                    // width groups of one batch
  int n;
  int tile_begin[9];
  CUtensorMap tmX[8];                  // 5-D activation maps (c, W, H, chunk, image)
  __nv_bfloat16* out[8];
  int OW[8], OH[8];
};

__global__ void __launch_bounds__(kThreads, 1)
conv2_swap_kernel(const __grid_constant__ SwapGroups G, const __grid_constant__ CUtensorMap tmW,
                  const float* __restrict__ bias, int n_tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* sW = smem;
  uint8_t* sX = smem + kXOff;
  uint8_t* sOut = smem + kOutOff;
  SwapBars* bars = reinterpret_cast<SwapBars*>(smem + kBarOff);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == kTmaWarp && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&bars->tmem_full[a], 1); mbar_init(&bars->tmem_empty[a], kEpiWarps); }
    mbar_init(&bars->w_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < G.n; ++i) tma_prefetch_desc(&G.tmX[i]);
  }
  if (warp == kMmaWarp) { tmem_alloc(&bars->tmem_base, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_trigger();
  if (warp != kMmaWarp) pdl_wait();                    // the activation comes from the previous kernel

  // contiguous tile range of this CTA
  const int per = (n_tiles + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int t0 = static_cast<int>(blockIdx.x) * per;
  const int t1 = (t0 + per < n_tiles) ? t0 + per : n_tiles;

  auto locate = [&](int tile, int& gi, int& img, int& oy0, int& ox0) {
    gi = 0;
#pragma unroll
    for (int i = 1; i < 8; ++i)
      if (i < G.n && tile >= G.tile_begin[i]) gi = i;
    const int lt = tile - G.tile_begin[gi];
    const int tpr = G.OW[gi] / kTileS, tpi = tpr * (G.OH[gi] / kTileR);
    img = lt / tpi;
    const int rem = lt - img * tpi;
    const int yb = rem / tpr;
    oy0 = yb * kTileR;
    ox0 = (rem - yb * tpr) * kTileS;
  };

  if (warp == kTmaWarp) {
    // ============================ TMA producer ============================
    if (elect_one()) {                                  // all nine taps of the weights, once
      mbar_arrive_expect_tx(&bars->w_full, kTaps * kWTapBytes);
      for (int t = 0; t < kTaps; ++t) tma_load_3d(sW + t * kWTapBytes, &tmW, &bars->w_full, 0, 0, t);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = t0; tile < t1; ++tile) {
      int gi, img, oy0, ox0;
      locate(tile, gi, img, oy0, ox0);
      const CUtensorMap* tmX = &G.tmX[gi];
      for (int t = 0; t < kTaps; ++t) {
        const int ky = t / 3, kx = t - 3 * ky;
        mbar_wait(&bars->empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars->full[stage], kXBytes);
          tma_load_5d(sX + stage * kXBytes, tmX, &bars->full[stage], 0, ox0 * 2 - 1 + kx, oy0 * 2 - 1 + ky, 0, img);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================ MMA issuer ============================
    const uint32_t idesc = umma_idesc_bf16(128, kTilePx);
    const uint32_t w_addr = smem_u32(sW), x_addr = smem_u32(sX);
    mbar_wait(&bars->w_full, 0);
    tc_fence_after();
    int stage = 0, it = 0;
    uint32_t phase = 0;
    for (int tile = t0; tile < t1; ++tile, ++it) {
      const int acc = it & 1;
      mbar_wait(&bars->tmem_empty[acc], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      for (int t = 0; t < kTaps; ++t) {
        mbar_wait(&bars->full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int h = 0; h < 3; ++h) {                 // K steps that hold channels 0..47 (48..63 are zero fill)
            const uint64_t ad = umma_desc_kmajor(w_addr + t * kWTapBytes + h * 32, 1024, UMMA_LAYOUT_SW128);
            const uint64_t bd = umma_desc_kmajor(x_addr + stage * kXBytes + h * 32, 1024, UMMA_LAYOUT_SW128);
            umma_bf16(d_tmem, ad, bd, idesc, (t | h) != 0 ? 1u : 0u);
          }
          umma_commit(&bars->empty[stage]);
          if (t + 1 == kTaps) umma_commit(&bars->tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ============================ epilogue warps ============================
    // warp w: TMEM lane quarter q = w & 3 = channels [32q, 32q + 32) (q = 3: the padding rows of the M = 128 tile, nothing
    // to store), pixel half ph = w >> 2 = accumulator columns [128 ph, 128 ph + 128) = tile rows 2 ph, 2 ph + 1.
    const int q = warp & 3, ph = warp >> 2;
    const int ch = q * 32 + lane;
    const float b = (ch < kCo) ? __ldg(bias + ch) : 0.f;
    const uint32_t so = smem_u32(sOut) + warp * kOutWarpBytes;
    int it = 0;
    for (int tile = t0; tile < t1; ++tile, ++it) {
      int gi, img, oy0, ox0;
      locate(tile, gi, img, oy0, ox0);
      const int acc = it & 1;
      mbar_wait(&bars->tmem_full[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16) + ph * 128;
      __nv_bfloat16* obase = G.out[gi];
      const int OW = G.OW[gi], OH = G.OH[gi];
#pragma unroll 1
      for (int p = 0; p < 4; ++p) {                     // 32 pixels at a time: half of a tile row
        uint32_t v[32];
        tmem_ld32(taddr + p * 32, v);
        tmem_ld_wait();
        if (p == 3) {                                   // the accumulator has been read: the MMA warp may reuse it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
          if (n_tiles < 0) __trap();                    // (ends the block here: ptxas sinks an arrival below arithmetic otherwise)
        }
        if (q < 3) {
          // channel `ch` of 32 pixels -> [pixel][channel] tile: lane = channel, 64 contiguous bytes per pixel row
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 y = silu_fast2(make_float2(__uint_as_float(v[i]) + b, __uint_as_float(v[i + 1]) + b));
            const uint32_t pk = pack_bf16x2(y.x, y.y);
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(so + i * 64 + lane * 2), "h"(static_cast<unsigned short>(pk & 0xffffu)) : "memory");
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(so + (i + 1) * 64 + lane * 2), "h"(static_cast<unsigned short>(pk >> 16)) : "memory");
          }
          __syncwarp();
          // copy out: a pixel's 32 channels are 64 contiguous bytes of the NHWC row (192 B per pixel)
          const int r = ph * 2 + (p >> 1), x0 = ox0 + (p & 1) * 32;
          const size_t pix0 = (static_cast<size_t>(img) * OH + oy0 + r) * OW + x0;
          uint8_t* gout = reinterpret_cast<uint8_t*>(obase) + pix0 * (kCo * 2) + q * 64;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int e = i * 32 + lane;                // 128 sixteen-byte pieces: pixel e / 4, piece e % 4
            const float4 d = lds128(so + e * 16);
            *reinterpret_cast<float4*>(gout + static_cast<size_t>(e >> 2) * (kCo * 2) + (e & 3) * 16) = d;
          }
          __syncwarp();
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// True when every problem has the shape the kernel is built for (the recogniser's conv2 on bucket widths).
bool conv2_swap_supported(const GemmLaunch* Ls, int n) {
  static const bool off = getenv("KIRI_CONV2_NO_SWAP") != nullptr;
  if (off || n < 1 || n > 8) return false;
  for (int i = 0; i < n; ++i) {
    const GemmLaunch& L = Ls[i];
    if (L.kw != 3 || L.kh != 3 || L.sw != 2 || L.sh != 2 || L.pad != 1 || L.Cin != 64 || L.Cin_mem != 48 || L.N != kCo ||
        L.epi != EPI_BIAS_SILU_BF16 || L.OH % kTileR != 0 || L.OW % kTileS != 0 || L.e.ldc != kCo || L.w != Ls[0].w ||
        L.e.bias != Ls[0].e.bias)
      return false;
  }
  return true;
}

int launch_conv2_swap(const GemmLaunch* Ls, int n, cudaStream_t stream) {
  KIRI_REQUIRE(conv2_swap_supported(Ls, n), "conv2_swap: unsupported problem");
  SwapGroups G;
  memset(&G, 0, sizeof(G));
  long long tiles = 0;
  for (int i = 0; i < n; ++i) {
    const GemmLaunch& L = Ls[i];
    if (L.NB <= 0) continue;
    const int Cm = L.Cin_mem;
    cuuint64_t dims[5] = {(cuuint64_t)Cm, (cuuint64_t)L.IW, (cuuint64_t)L.IH, 1, (cuuint64_t)L.NB};
    cuuint64_t str[4] = {(cuuint64_t)Cm * 2, (cuuint64_t)L.IW * Cm * 2, 128, (cuuint64_t)L.IH * L.IW * Cm * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)(kTileS * 2), (cuuint32_t)(kTileR * 2), 1, 1};
    cuuint32_t es[5] = {1, 2, 2, 1, 1};
    if (encode_map(&G.tmX[G.n], L.a, 5, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16)) return -1;
    G.tile_begin[G.n] = static_cast<int>(tiles);
    G.out[G.n] = reinterpret_cast<__nv_bfloat16*>(L.e.out);
    G.OW[G.n] = L.OW; G.OH[G.n] = L.OH;
    tiles += static_cast<long long>(L.NB) * (L.OH / kTileR) * (L.OW / kTileS);
    KIRI_REQUIRE(tiles < 0x7fffffffll, "conv2_swap: too many tiles");
    ++G.n;
  }
  for (int i = G.n; i < 9; ++i) G.tile_begin[i] = static_cast<int>(tiles);
  if (tiles == 0) return 0;
  CUtensorMap tmW;
  {
    const int ktot = kTaps * 64;
    cuuint64_t dims[3] = {64, (cuuint64_t)kCo, (cuuint64_t)kTaps};
    cuuint64_t str[2] = {(cuuint64_t)ktot * 2, 128};
    cuuint32_t box[3] = {64, (cuuint32_t)kCo, 1};
    cuuint32_t es[3] = {1, 1, 1};
    if (encode_map(&tmW, Ls[0].w, 3, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16)) return -1;
  }
  KIRI_REQUIRE(kSmem <= gemm_tc_max_smem(), "conv2_swap: %d bytes of shared memory needed, %d available", kSmem, gemm_tc_max_smem());
  static bool configured[kMaxDevices] = {false};
  const int dslot = kiri_cur_device_slot();
  if (!configured[dslot]) {
    KIRI_CHECK_CUDA(cudaFuncSetAttribute(conv2_swap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    configured[dslot] = true;
  }
  const int sms = gemm_tc_num_sms();
  const int grid = tiles < sms ? static_cast<int>(tiles) : sms;
  KIRI_CHECK_CUDA(launch_pdl(conv2_swap_kernel, dim3(grid), dim3(kThreads), kSmem, stream, G, tmW, Ls[0].e.bias,
                             static_cast<int>(tiles)));
  return 0;
}

}  // namespace kiri
