// tcgen05/TMEM/TMA implicit-GEMM kernel + host launcher.  See gemm_tc.cuh for the design.
#include "gemm_tc.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#ifndef KIRI_EPI_PIPE
#define KIRI_EPI_PIPE 1
#endif

namespace kiri {

static constexpr int kTileM = 128;
// K is consumed in chunks of KC = 32 or 64 bf16 (64- or 128-byte rows, SWIZZLE_64B / _128B).
// 64-byte rows exist because the stem's channel counts 96 and 160 are multiples of 32 only.
static constexpr int kAccStride = 256;                  // TMEM columns per accumulator
static constexpr int kTmemCols = 512;
static constexpr int kEpiWarps = 8;                     // two per TMEM lane quarter (column halves)
// Epilogue warps take the LOW warp ids: the SM sub-partition arbiter favours the highest warp id, and
// the single-lane TMA / MMA issuers (warps 8, 9) must never be starved by epilogue arithmetic.
// Three TMA producer warps take k-blocks in turn: one thread cannot issue a box faster than every
// ~350-450 cycles (tools/tma_stream.cu: 1 issuer 51 B/clk/SM, 2 issuers 101, 3 issuers 119), and a
// conv4 tile is 45 k-blocks of two boxes.  A k-block's CPS channel chunks travel in ONE box per operand.
static constexpr int kTmaWarps = 3;
static constexpr int kTmaWarp = kEpiWarps, kMmaWarp = kEpiWarps + kTmaWarps;
static constexpr int kNumThreads = (kEpiWarps + kTmaWarps + 1) * 32;
static constexpr int kMaxStages = 8;
static constexpr int kBufBytes = 4096;                  // 32 rows x 128 B epilogue staging tile
static constexpr int kMaxBResident = 128 * 1024;        // weights kept in shared memory when they fit

struct __align__(16) PipeBarriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t b_full, b_empty;          // resident-B handshake (TMA <-> MMA)
  uint64_t res_full[kEpiWarps][4];   // per epilogue warp: residual chunk landed
  uint32_t tmem_base;
  uint32_t pad[3];
  float bias[2][256];                // bias slice of the tile being drained, per accumulator
  float ln_g[256], ln_b[256];
  float xch[2][2][128];              // LayerNorm partial sums: [stat][column half][tile row]
};

template <int EPI>
__device__ __forceinline__ float epi_act(float v) {
  if (EPI == EPI_BIAS_SILU_BF16) return silu_fast(v);
  if (EPI == EPI_BIAS_GELU_BF16) return gelu_tanh_erf(v);
  return v;
}
template <int EPI>
__device__ __forceinline__ float2 epi_act2(float2 v) {       // a pair per instruction (packed fp32 math)
  if (EPI == EPI_BIAS_SILU_BF16) return silu_fast2(v);
  if (EPI == EPI_BIAS_GELU_BF16) return gelu_tanh_erf2(v);
  return v;
}

// Role-level cycle accounting of CTA 0 (KIRI_GEMM_TIMING=1 -> EpiParams::timing), read back with
// kiri_debug_gemm_timing(): [0] TMA wait-empty [1] TMA total [2] MMA wait-full [3] MMA wait-tmem-empty
// [4] MMA total [5] EPI wait-tmem-full [6] EPI total [7] EPI wait-residual [8] EPI wait-store-read [9] tiles
__device__ long long g_gemm_prof[16];
// cycles are accumulated in registers (acc) and flushed once per role: a global read-modify-write
// per sample would stall the single-lane issuers for an L2 round trip and distort the picture
#define GT_BEGIN(var) long long var = 0; if (timing) var = clock64()
#define GT_ACC(acc, var) do { if (timing) { const long long _t = clock64(); acc += _t - var; var = _t; } } while (0)
#define GT_FLUSH(slot, acc) do { if (timing) g_gemm_prof[slot] += acc; } while (0)

template <int KC, int NSEG, int EPI, bool BSTAT>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tc_kernel(const __grid_constant__ ProblemSet P, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmOut2, const EpiParams e,
               const int bn, const int num_m_tiles, const int num_n_tiles, const int stages, const int CPS, const int abox,
               const int bbox) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int kChunkBytes = KC * 2;
  constexpr int kATileBytes = kTileM * kChunkBytes;
  constexpr uint32_t kSBO = 8 * kChunkBytes;                 // 8-row group pitch
  constexpr uint64_t kLayout = (KC == 64) ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW64;
  constexpr bool kF32Out = (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_F32 || EPI == EPI_BIAS_RESID_LN || EPI == EPI_CTC_STATS);
  constexpr bool kResid = (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_RESID_LN);
  constexpr int kNBuf = kResid ? 4 : ((BSTAT || KC == 32) ? 1 : 2);   // staging tiles per epilogue warp (KC = 32: a third 54 KB stage instead)
  // carve: [resident B] | [stages][A: CPS chunk tiles][B: CPS chunk tiles] | staging | barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  // Several problems of ONE layer (the width groups of a batch: same weights, kernel, strides, N) share a
  // launch: the m-tiles of all problems form one range, a tile looks its problem up in P.mtile_begin.
  // Fields that do not depend on the problem are read from problem 0.
  const ConvGeom& g0 = P.g[0];
  const int num_chunks = g0.taps * g0.chunks_per_tap;        // K / KC
  const uint32_t b_chunk_bytes = bn * kChunkBytes;
  const uint32_t b_res_bytes = BSTAT ? ((num_chunks * b_chunk_bytes + 1023u) & ~1023u) : 0u;
  const uint32_t a_stage_bytes = CPS * kATileBytes;
  const uint32_t b_stage_bytes = BSTAT ? 0u : CPS * b_chunk_bytes;
  const uint32_t stage_bytes = (a_stage_bytes + b_stage_bytes + 1023u) & ~1023u;
  uint8_t* bres = smem;
  uint8_t* pipe = smem + b_res_bytes;
  uint8_t* staging = pipe + (size_t)stages * stage_bytes;
  PipeBarriers* bars = reinterpret_cast<PipeBarriers*>(staging + kEpiWarps * kNBuf * kBufBytes);

  // the shuffle makes the warp index (and everything derived from it: role, staging slot, rows) provably
  // warp-uniform for the compiler, so TMA / tcgen05 operands live in uniform registers
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int total_tiles = num_m_tiles * num_n_tiles;
  // contiguous, n-major tile range of this CTA: B changes at most twice per CTA
  const int t_begin = static_cast<int>(static_cast<long long>(blockIdx.x) * total_tiles / gridDim.x);
  const int t_end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * total_tiles / gridDim.x);
  const int num_kb = g0.taps * g0.cgs;
  const bool timing = e.timing != 0 && blockIdx.x == 0;

  if (warp == kTmaWarp && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 1);
      mbar_init(&bars->tmem_empty[a], kEpiWarps);
    }
    mbar_init(&bars->b_full, 1);
    mbar_init(&bars->b_empty, 1);
    for (int wq = 0; wq < kEpiWarps; ++wq)
      for (int r = 0; r < 4; ++r) mbar_init(&bars->res_full[wq][r], 1);
    fence_mbar_init();
    for (int i = 0; i < P.n; ++i) { tma_prefetch_desc(&P.tmA[i]); tma_prefetch_desc(&P.tmOut[i]); }
    tma_prefetch_desc(&tmB);
  }
  if (warp == kMmaWarp) {
    tmem_alloc(&bars->tmem_base, kTmemCols);
    tmem_relinquish();
  }
  if (EPI == EPI_BIAS_RESID_LN && threadIdx.x < 256) {
    const int t = threadIdx.x;
    bars->ln_g[t] = __ldg(e.ln_g + t);
    bars->ln_b[t] = __ldg(e.ln_b + t);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_trigger();                                   // the next kernel may start its prologue
  pdl_wait();                                      // A / residual come from the previous kernel

  if (warp >= kTmaWarp && warp < kTmaWarp + kTmaWarps) {
    // ============================ TMA producers ============================
    // The WHOLE warp walks the loop (uniform control flow, coordinates in uniform registers) and one
    // elected lane issues.  With a single divergent lane running the loop every UTMALDG was wrapped in
    // an ELECT / R2UR.BROADCAST waterfall: ~640 cycles per stage, the bound of every GEMM of the model
    // (tools/tma_stream.cu, profiles/README.md).
    {
      int stage = 0;
      uint32_t phase = 0;
      int cur_n = -1, b_loads = 0;
      const int pw = warp - kTmaWarp;              // this producer owns k-blocks pw, pw + n_prod, ... of the CTA's stream
      // A producer may not run more than one ring revolution ahead of a stage's empty barrier (parity waits
      // alias two phases apart): no more active producers than stages.
      const int n_prod = stages < kTmaWarps ? stages : kTmaWarps;
      int turn = 0;
      GT_BEGIN(tt0);
      long long tt_all = tt0, acc_we = 0, acc_tt = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        const int n_tile = tile / num_m_tiles;
        const int gm_tile = tile - n_tile * num_m_tiles;
        int pi = 0;
#pragma unroll
        for (int i = 1; i < kMaxProblems; ++i)
          if (i < P.n && gm_tile >= P.mtile_begin[i]) pi = i;
        const int m_tile = gm_tile - P.mtile_begin[pi];
        const ConvGeom& g = P.g[pi];
        const CUtensorMap* tmA = &P.tmA[pi];
        // one TMA box = one KC-channel chunk of one segment: R*SEG rows x (KC*2) B
        const uint32_t box_bytes = g.R * g.SEG * kChunkBytes;
        if (BSTAT && pw == 0 && n_tile != cur_n) {
          if (b_loads > 0) mbar_wait(&bars->b_empty, (b_loads - 1) & 1);   // MMAs on the old B retired
          if (elect_one()) {
            mbar_arrive_expect_tx(&bars->b_full, num_chunks * b_chunk_bytes);
            for (int ck = 0; ck < num_chunks; ++ck)
              tma_load_3d(bres + (size_t)ck * b_chunk_bytes, &tmB, &bars->b_full, 0, n_tile * bn, ck);
          }
          __syncwarp();
          cur_n = n_tile;
          ++b_loads;
        }
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
        int ky = 0, kx = 0, cg = 0;                 // k-block = (tap, channel group), walked without divisions
        for (int kb = 0; kb < num_kb; ++kb) {
          if (turn == pw) {
          if (timing) tt0 = clock64();
          mbar_wait(&bars->empty[stage], phase ^ 1);
          GT_ACC(acc_we, tt0);
          if (elect_one()) {
            uint8_t* a_dst = pipe + (size_t)stage * stage_bytes;
            uint8_t* b_dst = a_dst + a_stage_bytes;
            mbar_arrive_expect_tx(&bars->full[stage], tx_bytes);
            // the CPS chunks of the k-block are the slowest box dimension: they land as CPS consecutive
            // K-major tiles (CPS > 1 only with NSEG == 1, i.e. one 128-row segment per tile)
            // (abox / bbox = chunks per A / B box: CPS, or 1 with one box per chunk)
            for (int c = 0; c < CPS; c += abox) {
#pragma unroll
              for (int j = 0; j < NSEG; ++j) {
                if (seg_b[j] >= 0)
                  tma_load_5d(a_dst + (size_t)c * kATileBytes + (size_t)j * box_bytes, tmA, &bars->full[stage], 0, seg_x[j] + kx,
                              seg_y[j] + ky, cg * CPS + c, seg_b[j]);
              }
            }
            if (!BSTAT) {
              for (int c = 0; c < CPS; c += bbox)
                tma_load_3d(b_dst + (size_t)c * b_chunk_bytes, &tmB, &bars->full[stage], 0, n_tile * bn,
                            (ky * g.kw + kx) * g.chunks_per_tap + cg * CPS + c);
            }
          }
          __syncwarp();
          }
          if (++turn == n_prod) turn = 0;
          if (++cg == g.cgs) { cg = 0; if (++kx == g.kw) { kx = 0; ++ky; } }
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
      GT_ACC(acc_tt, tt_all);
      if (lane == 0 && pw == 0) { GT_FLUSH(0, acc_we); GT_FLUSH(1, acc_tt); }
    }
  } else if (warp == kMmaWarp) {
    // ============================ MMA issuer ============================
    // whole warp in the loop, one elected lane issues: descriptors are built in uniform registers
    {
      const uint32_t idesc = umma_idesc_bf16(kTileM, bn);
      const int k16 = g0.k16;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      int cur_n = -1, b_loads = 0;
      const uint32_t bres_addr = smem_u32(bres);
      const uint32_t pipe_addr = smem_u32(pipe);
      GT_BEGIN(tm0);
      long long tm_all = tm0, acc_wf = 0, acc_wt = 0, acc_mt = 0;
      for (int tile = t_begin; tile < t_end; ++tile, ++it) {
        const int n_tile = tile / num_m_tiles;
        if (BSTAT && n_tile != cur_n) {
          mbar_wait(&bars->b_full, b_loads & 1);
          cur_n = n_tile;
          ++b_loads;
        }
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        if (timing) tm0 = clock64();
        mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
        GT_ACC(acc_wt, tm0);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        const bool last_of_b = BSTAT && (tile + 1 == t_end || (tile + 1) / num_m_tiles != n_tile);
        for (int kb = 0; kb < num_kb; ++kb) {
          if (timing) tm0 = clock64();
          mbar_wait(&bars->full[stage], phase);
          GT_ACC(acc_wf, tm0);
          tc_fence_after();
          const uint32_t a_addr = pipe_addr + static_cast<uint32_t>(stage) * stage_bytes;
          const uint32_t b_addr = BSTAT ? bres_addr + (uint32_t)(kb * CPS) * b_chunk_bytes : a_addr + a_stage_bytes;
          if (elect_one()) {
            for (int c = 0; c < CPS; ++c) {
#pragma unroll
              for (int h = 0; h < KC / 16; ++h) {
                if (h >= k16) break;                // zero-filled tail of the chunk (conv2: channels 48..63): nothing to add
                const uint64_t ad = umma_desc_kmajor(a_addr + c * kATileBytes + h * 32, kSBO, kLayout);
                const uint64_t bd = umma_desc_kmajor(b_addr + c * b_chunk_bytes + h * 32, kSBO, kLayout);
                umma_bf16(d_tmem, ad, bd, idesc, (kb | c | h) != 0 ? 1u : 0u);
              }
            }
            umma_commit(&bars->empty[stage]);       // frees the smem stage when the MMAs retire
            if (kb + 1 == num_kb) {
              umma_commit(&bars->tmem_full[acc]);   // accumulator complete -> epilogue
              if (last_of_b) umma_commit(&bars->b_empty);   // last MMAs that read this B
            }
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
      GT_ACC(acc_mt, tm_all);
      if (lane == 0) {
        GT_FLUSH(2, acc_wf); GT_FLUSH(3, acc_wt); GT_FLUSH(4, acc_mt);
        if (timing) g_gemm_prof[9] += t_end - t_begin;
      }
    }
  } else {
    // ============================ epilogue warps ============================
    // Thread = one tile row (TMEM lane) x one column half.  Everything that touches global memory
    // goes through 32-row x 128-byte shared-memory tiles moved by TMA (bulk tensor stores / loads),
    // so the global traffic is whole 128-byte lines although a thread owns a row, not a column range.
    const int q = warp & 3;                        // TMEM lane quarter this warp may read
    const int ew = warp;                           // staging / residual slot of this warp
    const int half = ew >> 2;                      // column half (chunks are dealt round-robin)
    uint8_t* bufs = staging + ew * kNBuf * kBufBytes;
    uint32_t stg_cnt = 0;                          // staging tiles issued (for buffer rotation)
    // first row of this warp inside the tile and its decomposition (per problem)
    const int m0 = q * 32;
    auto tile_rows = [&](int tile, int& row0, int& pi) -> bool {
      const int n_tile = tile / num_m_tiles;
      const int gm_tile = tile - n_tile * num_m_tiles;
      pi = 0;
#pragma unroll
      for (int i = 1; i < kMaxProblems; ++i)
        if (i < P.n && gm_tile >= P.mtile_begin[i]) pi = i;
      const int m_tile = gm_tile - P.mtile_begin[pi];
      const ConvGeom& g = P.g[pi];
      const int j = m0 / (g.R * g.SEG);
      const int within = m0 - j * (g.R * g.SEG);
      const int jj = within / g.SEG;
      const int ii = within - jj * g.SEG;
      const int s = m_tile * NSEG + j;
      row0 = 0;
      if (s >= g.n_seg_total) return false;
      const int b = s / g.segs_per_img;
      const int rem = s - b * g.segs_per_img;
      const int yb = rem / g.segs_per_row;
      const int xb = rem - yb * g.segs_per_row;
      const int oy = yb * g.R + jj;
      const int ox = xb * g.SEG + ii;
      row0 = (b * g.OH + oy) * g.OW + ox;
      return (oy < g.OH) && (ox < g.OW);
    };
    if (kResid && lane == 0 && t_begin < t_end) {
      int row0, pi0;
      if (tile_rows(t_begin, row0, pi0)) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          mbar_arrive_expect_tx(&bars->res_full[ew][c], kBufBytes);
          tma_load_2d(bufs + c * kBufBytes, &tmRes, &bars->res_full[ew][c], half * 128 + c * 32, row0);
        }
      }
    }
    int it = 0;
    uint32_t res_tiles = 0;                        // valid residual tiles consumed (parity)
    const bool etime = timing && warp == 2 && lane == 0;   // epilogue warp 2 (quarter 2, first column half)
    long long te0 = 0, te_all = 0, acc_ef = 0, acc_er = 0, acc_es = 0;
    if (etime) { te0 = clock64(); te_all = te0; }
    for (int tile = t_begin; tile < t_end; ++tile, ++it) {
      const int n_tile = tile / num_m_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int row0, pi;
      const bool valid = tile_rows(tile, row0, pi);
      const CUtensorMap* tmOut = &P.tmOut[pi];
      const int col_base = n_tile * bn;
      // stage this tile's bias slice in shared memory (overlaps the wait for the accumulator)
      float* sbias = bars->bias[acc];
      {
        const int t = ew * 32 + lane;
        sbias[t] = (t < bn && col_base + t < e.n_valid) ? __ldg(e.bias + col_base + t) : 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      if (etime) te0 = clock64();
      mbar_wait(&bars->tmem_full[acc], acc_phase);
      if (etime) acc_ef += clock64() - te0;
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccStride;

      if (kResid) {
        // ---- x = resid + acc + bias (fp32 out); EPI_BIAS_RESID_LN also emits LayerNorm(x) in bf16.
        // The thread's 128 columns live in registers: ONE pass over TMEM, the accumulator is released
        // before any arithmetic, and the row statistics are exchanged with the other column half.
        const int cb = half * 128;
        uint32_t v[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(taddr + cb + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (etime) te0 = clock64();
          if (valid) mbar_wait(&bars->res_full[ew][c], res_tiles & 1);
          if (etime) acc_er += clock64() - te0;
          uint8_t* rb = bufs + c * kBufBytes;
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const float4 b = *reinterpret_cast<const float4*>(sbias + cb + c * 32 + 4 * t);
            float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) q4 = *reinterpret_cast<const float4*>(rb + stg_off(lane, t));
            float4 o;
            o.x = __uint_as_float(v[c * 32 + 4 * t]) + b.x + q4.x;
            o.y = __uint_as_float(v[c * 32 + 4 * t + 1]) + b.y + q4.y;
            o.z = __uint_as_float(v[c * 32 + 4 * t + 2]) + b.z + q4.z;
            o.w = __uint_as_float(v[c * 32 + 4 * t + 3]) + b.w + q4.w;
            *reinterpret_cast<float4*>(rb + stg_off(lane, t)) = o;      // in place: becomes the x tile
            sum += (o.x + o.y) + (o.z + o.w);
            v[c * 32 + 4 * t] = __float_as_uint(o.x); v[c * 32 + 4 * t + 1] = __float_as_uint(o.y);
            v[c * 32 + 4 * t + 2] = __float_as_uint(o.z); v[c * 32 + 4 * t + 3] = __float_as_uint(o.w);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && valid) {
#pragma unroll
          for (int c = 0; c < 4; ++c) tma_store_2d(tmOut, bufs + c * kBufBytes, cb + c * 32, row0);
          bulk_commit_group();
        }
        if (valid) ++res_tiles;
        if (EPI == EPI_BIAS_RESID_LN) {
          const int trow = q * 32 + lane;
          bars->xch[0][half][trow] = sum;
          asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
          const float mean = (sum + bars->xch[0][half ^ 1][trow]) * (1.0f / 256.0f);
          float sq = 0.f;
#pragma unroll
          for (int t = 0; t < 128; ++t) { const float d = __uint_as_float(v[t]) - mean; sq = fmaf(d, d, sq); }
          bars->xch[1][half][trow] = sq;
          asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
          const float rstd = 1.0f / sqrtf((sq + bars->xch[1][half ^ 1][trow]) * (1.0f / 256.0f) + 1e-5f);
          if (etime) te0 = clock64();
          if (lane == 0) bulk_wait_group_read<0>();              // the x stores have read their tiles
          if (etime) acc_es += clock64() - te0;
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 2; ++c) {                          // 64 bf16 columns = 128 B per row
            uint8_t* ob = bufs + c * kBufBytes;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const int col = cb + c * 64 + t * 8;
              const float4 g0 = *reinterpret_cast<const float4*>(bars->ln_g + col);
              const float4 g1 = *reinterpret_cast<const float4*>(bars->ln_g + col + 4);
              const float4 h0 = *reinterpret_cast<const float4*>(bars->ln_b + col);
              const float4 h1 = *reinterpret_cast<const float4*>(bars->ln_b + col + 4);
              const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
              const float hh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
              float y[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) y[u] = (__uint_as_float(v[c * 64 + t * 8 + u]) - mean) * rstd * gg[u] + hh[u];
              uint4 pk;
              pk.x = pack_bf16x2(y[0], y[1]); pk.y = pack_bf16x2(y[2], y[3]);
              pk.z = pack_bf16x2(y[4], y[5]); pk.w = pack_bf16x2(y[6], y[7]);
              *reinterpret_cast<uint4*>(ob + stg_off(lane, t)) = pk;
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && valid) {
            tma_store_2d(&tmOut2, bufs, cb, row0);
            tma_store_2d(&tmOut2, bufs + kBufBytes, cb + 64, row0);
            bulk_commit_group();
          }
        }
        // the residual of the next tile can land as soon as the stores have read the tiles
        if (lane == 0 && tile + 1 < t_end) {
          int nrow0, npi;
          if (tile_rows(tile + 1, nrow0, npi)) {
            bulk_wait_group_read<0>();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              mbar_arrive_expect_tx(&bars->res_full[ew][c], kBufBytes);
              tma_load_2d(bufs + c * kBufBytes, &tmRes, &bars->res_full[ew][c], cb + c * 32, nrow0);
            }
          }
        }
        __syncwarp();
      } else {
        constexpr int kCols = kF32Out ? 32 : 64;                 // columns per 128-byte staging row
        const int nch = (bn + kCols - 1) / kCols;
        if constexpr (!kF32Out && KIRI_EPI_PIPE) {
          // bf16 outputs: a warp has at most two 64-column chunks (bn <= 256).  A chunk is converted in registers
          // BEFORE the wait on the previous bulk store's read of the staging tile, and the second chunk's TMEM read is
          // issued before the first chunk is staged, so both latencies overlap work instead of adding up.
          const int cA = half * kCols, cB = (half + 2) * kCols;
          const bool hasA = half < nch, hasB = half + 2 < nch;
          uint32_t lo[32], hi[32];
          uint4 pk[8];
          auto convert = [&](const int c0) {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const uint32_t* src = (t < 4) ? lo : hi;
              const int o8 = (t & 3) * 8;
              const float4 bb0 = *reinterpret_cast<const float4*>(sbias + c0 + t * 8);
              const float4 bb1 = *reinterpret_cast<const float4*>(sbias + c0 + t * 8 + 4);
              const float bb[8] = {bb0.x, bb0.y, bb0.z, bb0.w, bb1.x, bb1.y, bb1.z, bb1.w};
              float2 y[4];
#pragma unroll
              for (int u = 0; u < 4; ++u)
                y[u] = epi_act2<EPI>(fadd2(make_float2(__uint_as_float(src[o8 + 2 * u]), __uint_as_float(src[o8 + 2 * u + 1])),
                                           make_float2(bb[2 * u], bb[2 * u + 1])));
              pk[t].x = pack_bf16x2(y[0].x, y[0].y); pk[t].y = pack_bf16x2(y[1].x, y[1].y);
              pk[t].z = pack_bf16x2(y[2].x, y[2].y); pk[t].w = pack_bf16x2(y[3].x, y[3].y);
            }
          };
          auto stage_and_store = [&](const int c0) {
            uint8_t* ob = bufs + (stg_cnt % kNBuf) * kBufBytes;
            if (etime) te0 = clock64();
            if (lane == 0) bulk_wait_group_read<kNBuf - 1>();
            if (etime) acc_es += clock64() - te0;
            __syncwarp();
#pragma unroll
            for (int t = 0; t < 8; ++t) *reinterpret_cast<uint4*>(ob + stg_off(lane, t)) = pk[t];
            fence_proxy_async();
            __syncwarp();
            if (lane == 0 && valid) {
              tma_store_2d(tmOut, ob, col_base + c0, row0);
              bulk_commit_group();
            }
            ++stg_cnt;
          };
          auto release = [&]() {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
          };
          if (hasA) {
            tmem_ld32(taddr + cA, lo);                           // may run past bn: the store clips the columns
            tmem_ld32(taddr + cA + 32, hi);
            tmem_ld_wait();
            if (!hasB) release();
            convert(cA);
            if (hasB) { tmem_ld32(taddr + cB, lo); tmem_ld32(taddr + cB + 32, hi); }
            stage_and_store(cA);
            if (hasB) {
              tmem_ld_wait();
              release();
              if (num_m_tiles < 0) __trap();                     // keeps the arrival above the arithmetic
              convert(cB);
              stage_and_store(cB);
            }
          } else {
            release();
          }
          continue;
        }
        // EPI_CTC_STATS: running (max, first arg-max, sum exp) of this thread's row over its column chunks
        float st_m = -INFINITY, st_s = 0.f;
        int st_a = 0x7fffffff;
        for (int ch = half; ch < nch; ch += 2) {
          const int c0 = ch * kCols;
          uint32_t ra[32], rb2[32];
          tmem_ld32(taddr + c0, ra);
          if (!kF32Out) tmem_ld32(taddr + c0 + 32, rb2);         // may run past bn: columns are clipped by the store
          tmem_ld_wait();
          if (ch + 2 >= nch) {                                   // last chunk of this warp: accumulator drained
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
          }
          if (EPI == EPI_CTC_STATS) {
            // logits of this chunk (the same fp32 values the store path writes); the head's padding columns take no part
            float cm = -INFINITY;
            int ca = 0x7fffffff;
            float lv[32];
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              const int col = col_base + c0 + t;
              lv[t] = (col < e.n_stat) ? __uint_as_float(ra[t]) + sbias[c0 + t] : -INFINITY;
              if (lv[t] > cm) { cm = lv[t]; ca = col; }          // ascending columns: the first maximum wins
            }
            const float nm = fmaxf(st_m, cm);
            if (nm > -INFINITY) {
              float ssum = 0.f;
#pragma unroll
              for (int t = 0; t < 32; ++t) ssum += __expf(lv[t] - nm);      // exp(-inf) = 0 in the padding
              st_s = st_s * __expf(st_m - nm) + ssum;
              if (cm > st_m) st_a = ca;                          // strict: an earlier chunk keeps a tie
              st_m = nm;
            }
            if (!e.store_out) { ++stg_cnt; continue; }
          }
          uint8_t* ob = bufs + (stg_cnt % kNBuf) * kBufBytes;
          if (etime) te0 = clock64();
          if (lane == 0) bulk_wait_group_read<kNBuf - 1>();
          if (etime) acc_es += clock64() - te0;
          __syncwarp();
          if (kF32Out) {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const float4 b = *reinterpret_cast<const float4*>(sbias + c0 + 4 * t);
              const float4 o = make_float4(__uint_as_float(ra[4 * t]) + b.x, __uint_as_float(ra[4 * t + 1]) + b.y,
                                           __uint_as_float(ra[4 * t + 2]) + b.z, __uint_as_float(ra[4 * t + 3]) + b.w);
              *reinterpret_cast<float4*>(ob + stg_off(lane, t)) = o;
            }
          } else {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const uint32_t* src = (t < 4) ? ra : rb2;
              const int o8 = (t & 3) * 8;
              const float4 b0 = *reinterpret_cast<const float4*>(sbias + c0 + t * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(sbias + c0 + t * 8 + 4);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              float2 y[4];
#pragma unroll
              for (int u = 0; u < 4; ++u)
                y[u] = epi_act2<EPI>(fadd2(make_float2(__uint_as_float(src[o8 + 2 * u]), __uint_as_float(src[o8 + 2 * u + 1])),
                                           make_float2(bb[2 * u], bb[2 * u + 1])));
              uint4 pk;
              pk.x = pack_bf16x2(y[0].x, y[0].y); pk.y = pack_bf16x2(y[1].x, y[1].y);
              pk.z = pack_bf16x2(y[2].x, y[2].y); pk.w = pack_bf16x2(y[3].x, y[3].y);
              *reinterpret_cast<uint4*>(ob + stg_off(lane, t)) = pk;
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && valid) {
            tma_store_2d(tmOut, ob, col_base + c0, row0);
            bulk_commit_group();
          }
          ++stg_cnt;
        }
        if (half >= nch) {                                       // no chunk for this warp (bn <= 64 / 32)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
        }
        if (EPI == EPI_CTC_STATS) {
          // merge the two warps that share a TMEM lane quarter (they took the 32-column chunks of the row in turn);
          // on equal maxima the lower class index wins (torch.argmax returns the first maximum)
          const int trow = q * 32 + lane;
          bars->xch[0][half][trow] = st_m;
          bars->xch[1][half][trow] = st_s;
          reinterpret_cast<int*>(bars->ln_g)[half * 128 + trow] = st_a;
          asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
          if (half == 0 && valid) {
            const float om = bars->xch[0][1][trow], os = bars->xch[1][1][trow];
            const int oa = reinterpret_cast<const int*>(bars->ln_g)[128 + trow];
            const float m = fmaxf(st_m, om);
            const float stot = st_s * __expf(st_m - m) + os * __expf(om - m);
            e.stat_id[row0 + lane] = (om > st_m || (om == st_m && oa < st_a)) ? oa : st_a;
            e.stat_p[row0 + lane] = 1.0f / stot;
          }
          asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");      // xch may be rewritten by the next tile
        }
      }
    }
    if (etime) {
      g_gemm_prof[6] += clock64() - te_all;
      g_gemm_prof[5] += acc_ef; g_gemm_prof[7] += acc_er; g_gemm_prof[8] += acc_es;
    }
    if (lane == 0) bulk_wait_group<0>();          // all stores of this warp have landed
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
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
int gemm_tc_num_sms();
int gemm_tc_max_smem() { gemm_tc_num_sms(); return g_max_smem; }
int gemm_tc_num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&g_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  }
  return g_num_sms;
}

// cuTensorMapEncodeTiled costs several microseconds on the host; the recogniser re-issues the same
// few dozen (pointer, shape) combinations every step, so encoded maps are cached.
struct MapKey {
  const void* base;
  int rank, swz;
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5], estr[5];
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* p = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return static_cast<size_t>(h);
  }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;
static std::mutex g_map_mutex;

int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims,
               const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr,
               CUtensorMapSwizzle swz, CUtensorMapDataType dtype) {
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.rank = rank; key.swz = static_cast<int>(swz) | (static_cast<int>(dtype) << 8);
  for (int i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; key.estr[i] = estr[i]; }
  for (int i = 0; i < rank - 1; ++i) key.strides[i] = strides[i];
  {
    std::lock_guard<std::mutex> lock(g_map_mutex);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) { *m = it->second; return 0; }
  }
  EncodeTiledFn fn = get_encode_fn();
  KIRI_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  CUresult r = fn(m, dtype, rank, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KIRI_REQUIRE(r == CUDA_SUCCESS,
               "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu box %u,%u,%u", (int)r,
               rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
               (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], box[1], rank > 2 ? box[2] : 0);
  std::lock_guard<std::mutex> lock(g_map_mutex);
  if (g_map_cache.size() > 8192) g_map_cache.clear();
  g_map_cache.emplace(key, *m);
  return 0;
}

// 2-D row-major [rows, cols] view moved in 32-row x 128-byte tiles (epilogue stores / residual loads)
int encode_rowtile_map(CUtensorMap* m, const void* base, long long rows, int cols, int ld, bool f32) {
  const int esz = f32 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t str[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), 32};
  cuuint32_t es[2] = {1, 1};
  return encode_map(m, base, 2, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B,
                    f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
}

template <int KC, int NSEG, int EPI, bool BSTAT>
static int launch_inst(const ProblemSet& P, const CUtensorMap& tmB, const CUtensorMap& tmRes, const CUtensorMap& tmOut2,
                       const EpiParams& e, int bn, int num_m_tiles, int num_n_tiles, int CPS, int abox, int bbox,
                       cudaStream_t stream) {
  constexpr bool kResid = (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_RESID_LN);
  constexpr int kNBuf = kResid ? 4 : ((BSTAT || KC == 32) ? 1 : 2);
  const ConvGeom& g = P.g[0];
  const int num_chunks = g.taps * g.chunks_per_tap;
  const int b_res = BSTAT ? ((num_chunks * bn * KC * 2 + 1023) & ~1023) : 0;
  const int stage_bytes = (CPS * kTileM * KC * 2 + (BSTAT ? 0 : CPS * bn * KC * 2) + 1023) & ~1023;
  const int overhead = 1024 + (int)sizeof(PipeBarriers) + kEpiWarps * kNBuf * kBufBytes + b_res;
  int stages = (g_max_smem - overhead) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  KIRI_REQUIRE(stages >= 2, "gemm_tc: stage of %d bytes does not fit twice in shared memory", stage_bytes);
  const int smem = stages * stage_bytes + overhead;
  auto kern = gemm_tc_kernel<KC, NSEG, EPI, BSTAT>;
  static int configured[kMaxDevices] = {0};          // (one array per template instantiation)
  const int dslot = kiri_cur_device_slot();
  if (configured[dslot] < smem) {
    KIRI_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem));
    configured[dslot] = g_max_smem;
  }
  int grid = num_m_tiles * num_n_tiles;
  if (grid > g_num_sms) grid = g_num_sms;
  KIRI_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(kNumThreads), smem, stream, P, tmB, tmRes, tmOut2, e, bn, num_m_tiles,
                             num_n_tiles, stages, CPS, abox, bbox));
  return 0;
}

// Geometry + tensor maps of one problem.  KC / NSEG / CPS are decided here and must agree between the
// problems of one launch (the caller groups by NSEG).
static int prep_problem(const GemmLaunch& L, bool is_gemm, int KC, bool a_multi, ConvGeom* gp, int* nseg, int* cps,
                        CUtensorMap* tmA, CUtensorMap* tmOut, int* num_m_tiles) {
  ConvGeom& g = *gp;
  const int chunks = L.Cin / KC;
  int NSEG = 1, CPS = 1;
  if (is_gemm) {
    KIRI_REQUIRE(L.IH == 1 && L.NB == 1 && L.OH == 1 && L.OW == L.IW, "gemm_tc: plain GEMM wants [1,1,M,K]");
    g.R = 1; g.SEG = 128;
    KIRI_REQUIRE(L.Cin % 64 == 0, "gemm_tc: GEMM K=%d must be a multiple of 64", L.Cin);
    CPS = 1;
  } else {
    // A tile is 128 output pixels = R rows x SEG columns fetched by ONE TMA box per K chunk.  Partial
    // tiles (zero-filled by TMA, skipped by the epilogue per 32-row warp) are accepted up to 25 %
    // padding because one big box beats four 32-pixel boxes (the NSEG = 4 form, no padding).
    KIRI_REQUIRE(L.OW % 32 == 0, "gemm_tc: conv output width %d must be a multiple of 32", L.OW);
    const int cand[3][2] = {{1, 128}, {2, 64}, {4, 32}};
    double best = 1e9;
    for (int i = 0; i < 3; ++i) {
      const int R = cand[i][0], SEG = cand[i][1];
      const double padded = static_cast<double>((L.OW + SEG - 1) / SEG * SEG) * ((L.OH + R - 1) / R * R);
      const double waste = padded / (static_cast<double>(L.OW) * L.OH);
      if (waste < best - 1e-9) { best = waste; g.R = R; g.SEG = SEG; }
    }
    if (best > 1.25) { g.R = 1; g.SEG = 32; NSEG = 4; }
    if (KC == 64) CPS = 1;
    else CPS = (NSEG == 1 && chunks <= 3) ? chunks : 1;
    KIRI_REQUIRE(L.epi == EPI_BIAS_SILU_BF16, "gemm_tc: conv path is built with the SiLU epilogue only");
  }
  g.OH = L.OH; g.OW = L.OW;
  g.sw = L.sw; g.sh = L.sh; g.pad = L.pad; g.kw = L.kw; g.taps = L.kw * L.kh;
  g.chunks_per_tap = chunks; g.cgs = chunks / CPS;
  g.k16 = (L.Cin_mem > 0 && L.Cin_mem < L.Cin && chunks == 1) ? (L.Cin_mem + 15) / 16 : KC / 16;
  KIRI_REQUIRE(chunks % CPS == 0, "gemm_tc: %d chunks per tap not divisible by %d", chunks, CPS);
  g.segs_per_row = (L.OW + g.SEG - 1) / g.SEG;
  g.segs_per_img = ((L.OH + g.R - 1) / g.R) * g.segs_per_row;
  g.n_seg_total = L.NB * g.segs_per_img;
  *num_m_tiles = (g.n_seg_total + NSEG - 1) / NSEG;
  *nseg = NSEG;
  *cps = CPS;
  KIRI_REQUIRE(g.SEG * g.sw <= 256 && g.R * g.sh <= 256, "gemm_tc: TMA box too large");
  KIRI_REQUIRE(L.e.n_valid == L.N, "gemm_tc: n_valid must equal N");
  const bool f32_out = (L.epi == EPI_BIAS_F32 || L.epi == EPI_BIAS_RESID_F32 || L.epi == EPI_BIAS_RESID_LN || L.epi == EPI_CTC_STATS);
  KIRI_REQUIRE((static_cast<long long>(L.e.ldc) * (f32_out ? 4 : 2)) % 16 == 0,
               "gemm_tc: output row pitch must be a multiple of 16 bytes (ldc=%d)", L.e.ldc);
  KIRI_REQUIRE((reinterpret_cast<uintptr_t>(L.e.out) & 15) == 0, "gemm_tc: output must be 16-byte aligned");
  // Tensor maps: a box covers ONE KC-channel chunk, so a box lands in shared memory as [rows][KC*2 B] —
  // exactly the K-major swizzled operand tile.  A: (cKC, chunk, W, H, image)
  const CUtensorMapSwizzle swz = (KC == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  {
    // (cKC, W, H, chunk, image): the chunk dimension has the SMALLEST stride but sits after W / H, so a box of
    // CPS chunks lands as [chunk][pixel][KC*2 B] = CPS consecutive operand tiles
    // Cm = channels stored per pixel.  Cm < Cin (conv2 on conv1's dense 48 channels, K padded to 64): the inner
    // EXTENT is Cm while the box stays KC wide, so the TMA unit zero-fills channels Cm..KC-1 of every row in shared
    // memory (out-of-bounds elements of a tiled box read as zero) - the padding never exists in HBM
    const int Cm = L.Cin_mem > 0 ? L.Cin_mem : L.Cin;
    KIRI_REQUIRE(Cm == L.Cin || (chunks == 1 && Cm < L.Cin && (Cm * 2) % 16 == 0),
                 "gemm_tc: %d stored channels for Cin=%d needs a single K chunk per tap and a 16-byte pixel pitch", Cm, L.Cin);
    cuuint64_t dims[5] = {(cuuint64_t)(Cm < KC ? Cm : KC), (cuuint64_t)L.IW, (cuuint64_t)L.IH, (cuuint64_t)chunks, (cuuint64_t)L.NB};
    cuuint64_t str[4] = {(cuuint64_t)Cm * 2, (cuuint64_t)L.IW * Cm * 2, (cuuint64_t)KC * 2,
                         (cuuint64_t)L.IH * L.IW * Cm * 2};
    cuuint32_t box[5] = {(cuuint32_t)KC, (cuuint32_t)(g.SEG * g.sw), (cuuint32_t)(g.R * g.sh), (cuuint32_t)(a_multi ? CPS : 1), 1};
    cuuint32_t es[5] = {1, (cuuint32_t)g.sw, (cuuint32_t)g.sh, 1, 1};
    if (encode_map(tmA, L.a, 5, dims, str, box, es, swz)) return -1;
  }
  const long long rows_total = static_cast<long long>(L.NB) * L.OH * L.OW;
  if (encode_rowtile_map(tmOut, L.e.out, rows_total, L.N, L.e.ldc, f32_out)) return -1;
  return 0;
}

int launch_gemm_tc(const GemmLaunch& L, cudaStream_t stream) { return launch_gemm_tc_multi(&L, 1, stream); }

// n problems of one layer (same weights, bias, epilogue, kernel, strides, channel counts; different
// activations / geometry).  Problems that need different tile forms (NSEG) go to separate launches.
int launch_gemm_tc_multi(const GemmLaunch* Ls, int n, cudaStream_t stream) {
  KIRI_REQUIRE(Ls && n >= 1 && n <= kMaxProblems, "gemm_tc: 1..%d problems per call", kMaxProblems);
  GemmLaunch L = Ls[0];
  static const int timing_on = getenv("KIRI_GEMM_TIMING") != nullptr;
  L.e.timing = timing_on;
  gemm_tc_num_sms();
  KIRI_REQUIRE(L.Cin % 32 == 0, "gemm_tc: Cin=%d must be a multiple of 32", L.Cin);
  const bool is_gemm = (L.kw == 1 && L.kh == 1);
  KIRI_REQUIRE(n == 1 || !is_gemm, "gemm_tc: only conv problems can share a launch");
  for (int i = 0; i < n; ++i) {
    const GemmLaunch& Q = Ls[i];
    KIRI_REQUIRE(Q.e.bias != nullptr && Q.e.out != nullptr && Q.a != nullptr, "gemm_tc: a/bias/out must not be null");
    KIRI_REQUIRE(Q.w == L.w && Q.e.bias == L.e.bias && Q.N == L.N && Q.Cin == L.Cin && Q.Cin_mem == L.Cin_mem && Q.epi == L.epi && Q.kw == L.kw &&
                     Q.kh == L.kh && Q.sw == L.sw && Q.sh == L.sh && Q.pad == L.pad && Q.e.ldc == L.e.ldc,
                 "gemm_tc: problem %d is not the same layer as problem 0", i);
  }
  // 128-byte K rows whenever the channels allow; the residual/LayerNorm epilogues keep 128 KB of
  // staging tiles, so their pipeline uses the half-size (64-byte) stages to still be 3 deep
  const bool resid_epi = (L.epi == EPI_BIAS_RESID_F32 || L.epi == EPI_BIAS_RESID_LN);
  const int KC = (L.Cin % 64 == 0 && !resid_epi) ? 64 : 32;
  int bn = (L.N + 15) / 16 * 16;
  if (bn > 256) bn = 256;
  const int num_n_tiles = (L.N + bn - 1) / bn;

  CUtensorMap tmB, tmRes, tmOut2;
  const CUtensorMapSwizzle swz = (KC == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const int taps = L.kw * L.kh;
  // weights stay resident in shared memory for the whole CTA when they fit beside >= 3 A stages
  const int b_total = taps * L.Cin * bn * 2;

  // one box per k-block and operand (all CPS chunks) unless switched off for A/B experiments
  static const bool a_multi = getenv("KIRI_GEMM_ABOX1") == nullptr, b_multi = getenv("KIRI_GEMM_BBOX1") == nullptr;
  // group the problems by tile form
  bool done[kMaxProblems] = {false};
  for (int first = 0; first < n; ++first) {
    if (done[first]) continue;
    ProblemSet P;
    memset(&P, 0, sizeof(P));
    int nseg0 = 0, cps0 = 0, m_total = 0;
    for (int i = first; i < n; ++i) {
      if (done[i]) continue;
      ConvGeom g;
      CUtensorMap ta, to;
      int nseg, cps, mt;
      if (prep_problem(Ls[i], is_gemm, KC, a_multi, &g, &nseg, &cps, &ta, &to, &mt)) return -1;
      if (P.n == 0) { nseg0 = nseg; cps0 = cps; }
      if (nseg != nseg0 || cps != cps0) continue;
      P.g[P.n] = g; P.tmA[P.n] = ta; P.tmOut[P.n] = to;
      P.mtile_begin[P.n] = m_total;
      m_total += mt;
      ++P.n;
      done[i] = true;
    }
    for (int i = P.n; i <= kMaxProblems; ++i) P.mtile_begin[i] = m_total;
    const int NSEG = nseg0, CPS = cps0;
    {  // B: (cKC, N, chunk) — a box holds the CPS chunks of one k-block for bn output channels
      const int ktot = taps * L.Cin;
      cuuint64_t dims[3] = {(cuuint64_t)KC, (cuuint64_t)L.N, (cuuint64_t)(ktot / KC)};
      cuuint64_t str[2] = {(cuuint64_t)ktot * 2, (cuuint64_t)KC * 2};
      cuuint32_t box[3] = {(cuuint32_t)KC, (cuuint32_t)bn, (cuuint32_t)(b_multi ? CPS : 1)};
      cuuint32_t es[3] = {1, 1, 1};
      if (encode_map(&tmB, L.w, 3, dims, str, box, es, swz)) return -1;
    }
    tmRes = P.tmOut[0];
    tmOut2 = P.tmOut[0];
    if (resid_epi) {
      const long long rows_total = static_cast<long long>(L.NB) * L.OH * L.OW;
      KIRI_REQUIRE(L.N == 256 && L.e.ldc == 256 && L.e.resid, "gemm_tc: residual epilogues need N = ldc = 256 and resid");
      if (encode_rowtile_map(&tmRes, L.e.resid, rows_total, 256, 256, true)) return -1;
      if (L.epi == EPI_BIAS_RESID_LN) {
        KIRI_REQUIRE(L.e.ln_g && L.e.ln_b && L.e.out2, "gemm_tc: LayerNorm epilogue needs ln_g, ln_b, out2");
        if (encode_rowtile_map(&tmOut2, L.e.out2, rows_total, 256, 256, false)) return -1;
      }
    }
    const bool bstat = !resid_epi && CPS == 1 && b_total <= kMaxBResident && getenv("KIRI_GEMM_NO_BSTAT") == nullptr &&
                       (g_max_smem - 1024 - (int)sizeof(PipeBarriers) - kEpiWarps * kBufBytes - b_total) >= 3 * CPS * kTileM * KC * 2;
    int rc = -1;
#define KIRI_LAUNCH(K, S, E)                                                                                          \
  rc = bstat ? launch_inst<K, S, E, true>(P, tmB, tmRes, tmOut2, L.e, bn, m_total, num_n_tiles, CPS, a_multi ? CPS : 1, b_multi ? CPS : 1, stream)            \
             : launch_inst<K, S, E, false>(P, tmB, tmRes, tmOut2, L.e, bn, m_total, num_n_tiles, CPS, a_multi ? CPS : 1, b_multi ? CPS : 1, stream)
#define KIRI_LAUNCH_NB(K, S, E) rc = launch_inst<K, S, E, false>(P, tmB, tmRes, tmOut2, L.e, bn, m_total, num_n_tiles, CPS, a_multi ? CPS : 1, b_multi ? CPS : 1, stream)
    if (!is_gemm) {
      if (KC == 64 && NSEG == 1) KIRI_LAUNCH(64, 1, EPI_BIAS_SILU_BF16);
      else if (KC == 64 && NSEG == 4) KIRI_LAUNCH(64, 4, EPI_BIAS_SILU_BF16);
      else if (KC == 32 && NSEG == 1) KIRI_LAUNCH_NB(32, 1, EPI_BIAS_SILU_BF16);
      else if (KC == 32 && NSEG == 4) KIRI_LAUNCH_NB(32, 4, EPI_BIAS_SILU_BF16);
    } else {
      switch (L.epi) {
        case EPI_BIAS_BF16: KIRI_LAUNCH(64, 1, EPI_BIAS_BF16); break;
        case EPI_BIAS_SILU_BF16: KIRI_LAUNCH(64, 1, EPI_BIAS_SILU_BF16); break;
        case EPI_BIAS_GELU_BF16: KIRI_LAUNCH(64, 1, EPI_BIAS_GELU_BF16); break;
        case EPI_BIAS_RESID_F32: KIRI_LAUNCH_NB(32, 1, EPI_BIAS_RESID_F32); break;
        case EPI_BIAS_F32: KIRI_LAUNCH(64, 1, EPI_BIAS_F32); break;
        case EPI_BIAS_RESID_LN: KIRI_LAUNCH_NB(32, 1, EPI_BIAS_RESID_LN); break;
        case EPI_CTC_STATS:
          KIRI_REQUIRE(num_n_tiles == 1 && L.e.stat_id && L.e.stat_p && L.e.n_stat > 0 && L.e.n_stat <= L.N,
                       "gemm_tc: the CTC statistics epilogue needs N <= 256, stat_id, stat_p and 0 < n_stat <= N");
          KIRI_LAUNCH(64, 1, EPI_CTC_STATS); break;
        default: KIRI_REQUIRE(false, "gemm_tc: unknown epilogue %d", L.epi);
      }
    }
#undef KIRI_LAUNCH_NB
#undef KIRI_LAUNCH
    if (rc != 0) return rc;
  }
  return 0;
}


}  // namespace kiri

// Debug: role-level cycles of CTA 0 accumulated since the last call (KIRI_GEMM_TIMING=1).
extern "C" int kiri_debug_gemm_timing(long long* out_host, int n) {
  long long buf[16];
  if (cudaDeviceSynchronize() != cudaSuccess) return -2;
  if (cudaMemcpyFromSymbol(buf, kiri::g_gemm_prof, sizeof(buf)) != cudaSuccess) return -2;
  for (int i = 0; i < n && i < 16; ++i) out_host[i] = buf[i];
  long long zero[16] = {0};
  if (cudaMemcpyToSymbol(kiri::g_gemm_prof, zero, sizeof(zero)) != cudaSuccess) return -2;
  return 0;
}
