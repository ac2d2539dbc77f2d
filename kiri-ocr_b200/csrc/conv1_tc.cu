// K2 on the tensor pipe: first stem layer Conv(1->48, 3x3, s1, p1, no bias) + BN(eval) + SiLU as ONE tcgen05 MMA pair per
// 128 pixels.
//
// Replaces  ConvStem.net[0:3]   kiri_ocr/model.py:215-217   (same contract as conv1.cu: uint8 planes in, dense 48-channel
//           NHWC bf16 out, BN folded, the reference's normalisation (v/255 - 0.5)/0.5 of model.py:337-338 applied inside).
//
// Why: the CUDA-core form (conv1.cu) is issue-bound at 667 thread instructions per pixel (216 FFMA2 + 108 weight fetches
// + 48 SiLUs + staging; ncu: 104 M warp instructions, IPC 2.6 of 4, 0.134 ms for 5.0 M pixels).  The 432 FMAs of a pixel are
// a [1 x 9] x [9 x 48] product: K = 9 is tiny, but as an M = 128 tile it is one tcgen05.mma (K = 16), and what is left for
// the CUDA cores is building the A row (9 byte loads, 9 converts, 4 shared stores) and the SiLU + store of the result.
//
// Exact operands.  x_t = 2 v_t / 255 - 1 for a tap inside the image and 0 for a padded tap.  With u_t = v_t - 128 (an
// integer in [-128, 127]: EXACT in bf16) for inside taps and u_t = -0.5 (exact) for padded ones,
//     sum_t w_t x_t + b  =  sum_t (2 w_t / 255) u_t  +  (b + sum_t w_t / 255)          (both kinds of tap, identically)
// so the A row is [u_0 .. u_8, 1, 0 ...] and the B row of channel n is [w'_n0 .. w'_n8, b'_n, 0 ...] with w' = 2w/255,
// b' = b + sum_t w_t / 255.  The fp32 weights are split into two bf16 terms (hi + lo, 16 mantissa bits: 7.6e-6 relative) that
// sit in the SECOND K step of the same tile, against the same A values: D = A.hi + A.lo, fp32 accumulation.  Error against
// the fp32 convolution ~1e-5 relative, two orders below the bf16 rounding of the output.
//
// Layout: A and B are K-major 128-byte-swizzle tiles ([rows][64 bf16], the layout every other GEMM of the library uses);
// columns 0-15 = first K step (u | hi), 16-31 = second K step (u again | lo); K steps 2 and 3 are never issued.
// CTA = 128 threads = one 128-pixel segment of an image row at a time (thread = pixel = TMEM lane), persistent over a
// contiguous range of segments, software-pipelined over two A tiles and two 64-column accumulators; 50 KB of shared
// memory and 128 TMEM columns: four CTAs per SM.
#include "internal.cuh"

#include <cstdlib>
#include <cstring>

namespace kiri {

static constexpr int kC1 = 48;
static constexpr int kC1Chunks = kC1 * 2 / 16;        // 16-byte chunks per output pixel
static constexpr int kTcThreads = 128;
static constexpr int kTcABytes = 128 * 128, kTcBBytes = 48 * 128, kTcStageBytes = 128 * kC1 * 2;
static constexpr int kTcSmem = 2 * kTcABytes + kTcBBytes + kTcStageBytes + 64;     // 51 264 B: four CTAs per SM

struct Conv1TcGroups {          // width groups of one batch: group i owns segments [seg_begin[i], seg_begin[i+1])
  int n;
  int seg_begin[9];
  const uint8_t* planes[8];
  __nv_bfloat16* out[8];
  int W[8];
};

// staging slot of 16-byte chunk j of pixel px (same rotation as conv1.cu: conflict-free 128-bit stores at a 96-byte pitch)
__device__ __forceinline__ int c1t_slot(int px, int j) {
  int r = j + ((px >> 2) & 1);
  if (r >= kC1Chunks) r -= kC1Chunks;
  return px * kC1Chunks + r;
}

// Position of a 128-pixel segment, advanced without divisions (a CTA walks a contiguous range of segments).
struct SegPos {
  int gi, xt, y, row, spr, W, next_begin;     // group, segment in the row, image row in the plane, global row (b * H + y)
};

// Software pipeline, per CTA:   build A(i+1) | MMA(i+1) issued | epilogue(i): TMEM -> SiLU -> bf16 -> stores
// with two A tiles and two 64-column accumulators: the MMA's round trip and the global loads of the next segment's pixels
// run under the epilogue of the current one.
__global__ void __launch_bounds__(kTcThreads)
conv1_tc_kernel(const __grid_constant__ Conv1TcGroups G, int H, const uint4* __restrict__ bmat, int n_segs) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();         // no static shared memory here: the dynamic window starts aligned
  uint8_t* sB = smem + 2 * kTcABytes;                   // [48 ch][128 B], 128-byte swizzle
  uint4* sOut = reinterpret_cast<uint4*>(smem + 2 * kTcABytes + kTcBBytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * kTcABytes + kTcBBytes + kTcStageBytes);      // [2] MMA done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // weights: the host packed them in the swizzled layout already (conv1_tc_pack_host)
  for (int i = tid; i < kTcBBytes / 16; i += kTcThreads) reinterpret_cast<uint4*>(sB)[i] = __ldg(bmat + i);
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(tmem_slot, 128); tmem_relinquish(); }
  fence_proxy_async();                                  // sB was written through the generic proxy, the MMA reads it through the async one
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();                                           // the planes come from the previous kernel

  // contiguous range of segments of this CTA
  const int per = (n_segs + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int s0 = static_cast<int>(blockIdx.x) * per;
  const int s1 = (s0 + per < n_segs) ? s0 + per : n_segs;
  if (s0 >= s1) {                                       // (uniform: no barrier is skipped)
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
    return;
  }
  const uint32_t idesc = umma_idesc_bf16(128, kC1);
  const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(sB);
  const uint32_t a_row = a_addr + static_cast<uint32_t>(tid) * 128u, swz = static_cast<uint32_t>(tid & 7) << 4;
  const uint32_t mhalf_bf16 = 0xBF00u;                  // -0.5

  // staging / copy-out slots of this thread: the same for every segment
  int stg_slot[kC1Chunks], out_slot[kC1Chunks];
#pragma unroll
  for (int i = 0; i < kC1Chunks; ++i) {
    stg_slot[i] = c1t_slot(tid, i);
    const int e = i * 32 + lane;                        // element of the warp's 192 sixteen-byte chunks
    out_slot[i] = c1t_slot(warp * 32 + e / kC1Chunks, e % kC1Chunks);
  }
  auto locate = [&](int seg) {                          // (divisions: once per CTA)
    SegPos q;
    q.gi = 0;
#pragma unroll
    for (int i = 1; i < 8; ++i)
      if (i < G.n && seg >= G.seg_begin[i]) q.gi = i;
    q.W = G.W[q.gi]; q.spr = q.W >> 7; q.next_begin = G.seg_begin[q.gi + 1];
    const int ls = seg - G.seg_begin[q.gi];
    q.xt = ls % q.spr; q.row = ls / q.spr; q.y = q.row % H;
    return q;
  };
  auto advance = [&](SegPos& q, int seg_next) {
    if (seg_next >= q.next_begin && q.gi + 1 < G.n) {   // first segment of the next width group
      ++q.gi;
      q.W = G.W[q.gi]; q.spr = q.W >> 7; q.next_begin = G.seg_begin[q.gi + 1];
      q.xt = 0; q.row = 0; q.y = 0;
      return;
    }
    if (++q.xt == q.spr) { q.xt = 0; ++q.row; if (++q.y == H) q.y = 0; }
  };
  // the nine source bytes of this thread's pixel of a segment (0x100 marks a tap outside the image)
  auto load_px = [&](const SegPos& q, uint32_t (&px)[9]) {
    const int x = (q.xt << 7) + tid;
    const uint8_t* p = G.planes[q.gi] + static_cast<size_t>(q.row) * q.W + x;      // the pixel itself
    const bool l_ok = x > 0, r_ok = x + 1 < q.W;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const bool y_ok = (ky == 0) ? (q.y > 0) : ((ky == 2) ? (q.y + 1 < H) : true);
      const uint8_t* pr = p + (ky - 1) * q.W;
      px[ky * 3 + 0] = (y_ok && l_ok) ? static_cast<uint32_t>(__ldg(pr - 1)) : 0x100u;
      px[ky * 3 + 1] = y_ok ? static_cast<uint32_t>(__ldg(pr)) : 0x100u;
      px[ky * 3 + 2] = (y_ok && r_ok) ? static_cast<uint32_t>(__ldg(pr + 1)) : 0x100u;
    }
  };
  // A row of the pixel: u_t = v_t - 128 inside the image, -0.5 outside (see the header); written to A tile `buf`
  auto build_a = [&](const uint32_t (&px)[9], int buf) {
    uint32_t ub[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float u = static_cast<float>(static_cast<int>(px[t]) - 128);
      ub[t] = (px[t] & 0x100u) ? mhalf_bf16 : (__float_as_uint(u) >> 16);      // |u| <= 128: the low 16 mantissa bits are zero
    }
    const uint32_t p0 = ub[0] | (ub[1] << 16), p1 = ub[2] | (ub[3] << 16), p2 = ub[4] | (ub[5] << 16), p3 = ub[6] | (ub[7] << 16);
    const uint32_t p4 = ub[8] | (0x3F80u << 16);                               // (u_8, 1.0)
    const uint32_t ar = a_row + static_cast<uint32_t>(buf) * kTcABytes;
    sts128(ar + ((0u << 4) ^ swz), p0, p1, p2, p3);     // K step 0: columns 0-7
    sts128(ar + ((1u << 4) ^ swz), p4, 0u, 0u, 0u);     //           columns 8-15 (u_8, 1, 0 ...)
    sts128(ar + ((2u << 4) ^ swz), p0, p1, p2, p3);     // K step 1: the same values against the low halves of the weights
    sts128(ar + ((3u << 4) ^ swz), p4, 0u, 0u, 0u);
  };
  auto issue_mma = [&](int buf) {                       // warp 0, after the barrier that follows build_a
    tc_fence_after();
    if (elect_one()) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint64_t ad = umma_desc_kmajor(a_addr + buf * kTcABytes + h * 32, 1024, UMMA_LAYOUT_SW128);
        const uint64_t bd = umma_desc_kmajor(b_addr + h * 32, 1024, UMMA_LAYOUT_SW128);
        umma_bf16(tmem_base + buf * 64, ad, bd, idesc, h ? 1u : 0u);
      }
      umma_commit(&bar[buf]);
    }
    __syncwarp();
  };
  // SiLU + bf16 + coalesced stores of the segment whose accumulator is `buf`
  auto epilogue = [&](const SegPos& q, int buf, uint32_t parity) {
    mbar_wait(&bar[buf], parity);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + buf * 64;
    uint32_t r[48];
    tmem_ld32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
    tmem_ld16(taddr + 32, *reinterpret_cast<uint32_t(*)[16]>(&r[32]));
    tmem_ld_wait();
    tc_fence_before();                                  // (ordered before the barrier in front of the MMA that reuses `buf`)
#pragma unroll
    for (int j = 0; j < kC1Chunks; ++j) {
      uint32_t pk[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float2 sv = silu_fast2(make_float2(__uint_as_float(r[8 * j + 2 * c]), __uint_as_float(r[8 * j + 2 * c + 1])));
        pk[c] = pack_bf16x2(sv.x, sv.y);
      }
      sOut[stg_slot[j]] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    __syncwarp();
    // a warp's 32 pixels are 3 KiB of contiguous global memory: staged and copied out by the warp itself (no CTA barrier)
    uint4* dst = reinterpret_cast<uint4*>(G.out[q.gi] + (static_cast<size_t>(q.row) * q.W + (q.xt << 7) + warp * 32) * kC1) + lane;
#pragma unroll
    for (int i = 0; i < kC1Chunks; ++i) dst[i * 32] = sOut[out_slot[i]];
    __syncwarp();                                       // (the warp's staging slots are rewritten by its next epilogue)
  };

  SegPos cur = locate(s0), prev = cur;
  uint32_t px[9];
  load_px(cur, px);
  int it = 0;
  for (int seg = s0; seg < s1; ++seg, ++it) {
    const int buf = it & 1;
    build_a(px, buf);
    SegPos nxt = cur;
    if (seg + 1 < s1) { advance(nxt, seg + 1); load_px(nxt, px); }     // in flight under the barrier, the MMA and the epilogue
    fence_proxy_async();
    __syncthreads();                                    // A(buf) complete (and the accumulator `buf` has been read by everyone);
    if (warp == 0) issue_mma(buf);                      // (an mbarrier only the issuing warp waits on measured 3 % slower)
    if (it > 0) epilogue(prev, buf ^ 1, static_cast<uint32_t>(((it - 1) >> 1) & 1));
    prev = cur;
    cur = nxt;
  }
  epilogue(prev, (it - 1) & 1, static_cast<uint32_t>(((it - 1) >> 1) & 1));

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// B operand of the kernel: [48][64] bf16 in the 128-byte-swizzled K-major layout, from the BN-folded fp32 weights.
static void conv1_tc_pack_host(const float* w, const float* b, uint16_t* bmat /* [48 * 64] */) {
  auto bf16_rn = [](float f) -> uint16_t {
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint32_t lsb = (u >> 16) & 1u;
    u += 0x7FFFu + lsb;                                  // round to nearest even (finite inputs)
    return static_cast<uint16_t>(u >> 16);
  };
  auto bf16_f = [](uint16_t h) -> float {
    const uint32_t u = static_cast<uint32_t>(h) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
  };
  memset(bmat, 0, 48 * 64 * 2);
  for (int n = 0; n < 48; ++n) {
    float v[10];
    double sw = 0.0;
    for (int t = 0; t < 9; ++t) {
      v[t] = static_cast<float>(2.0 * static_cast<double>(w[n * 9 + t]) / 255.0);
      sw += static_cast<double>(w[n * 9 + t]);
    }
    v[9] = static_cast<float>(static_cast<double>(b[n]) + sw / 255.0);
    for (int k = 0; k < 10; ++k) {
      const uint16_t hi = bf16_rn(v[k]);
      const uint16_t lo = bf16_rn(v[k] - bf16_f(hi));
      for (int half = 0; half < 2; ++half) {
        const int col = half * 16 + k;                   // K step `half`, column k
        const int chunk = col >> 3, within = col & 7;
        bmat[n * 64 + ((chunk ^ (n & 7)) << 3) + within] = half ? lo : hi;
      }
    }
  }
}

int conv1_tc_build(const float* w_host, const float* b_host, void** bmat_dev) {
  uint16_t host[48 * 64];
  conv1_tc_pack_host(w_host, b_host, host);
  KIRI_CHECK_CUDA(cudaMalloc(bmat_dev, sizeof(host)));
  KIRI_CHECK_CUDA(cudaMemcpy(*bmat_dev, host, sizeof(host), cudaMemcpyHostToDevice));
  return 0;
}

int conv1_tc_launch(const uint8_t* const* planes_u8, void* const* out_bf16_nhwc48, const int* group_lines, const int* group_W,
                    int n_groups, const void* bmat_dev, int H, cudaStream_t stream) {
  KIRI_REQUIRE(planes_u8 && out_bf16_nhwc48 && group_lines && group_W && bmat_dev, "conv1_tc: null pointer");
  KIRI_REQUIRE(n_groups >= 0 && n_groups <= 8, "conv1_tc: at most 8 groups");
  KIRI_REQUIRE(H > 0, "conv1_tc: bad plane height %d", H);
  Conv1TcGroups G;
  memset(&G, 0, sizeof(G));
  long long segs = 0;
  for (int g = 0; g < n_groups; ++g) {
    if (group_lines[g] <= 0) continue;
    KIRI_REQUIRE(planes_u8[g] && out_bf16_nhwc48[g], "conv1_tc: null pointer in group %d", g);
    KIRI_REQUIRE(group_W[g] % 128 == 0, "conv1_tc: width %d must be a multiple of 128", group_W[g]);
    G.seg_begin[G.n] = static_cast<int>(segs);
    G.planes[G.n] = planes_u8[g];
    G.out[G.n] = reinterpret_cast<__nv_bfloat16*>(out_bf16_nhwc48[g]);
    G.W[G.n] = group_W[g];
    segs += static_cast<long long>(group_lines[g]) * H * (group_W[g] / 128);
    KIRI_REQUIRE(segs < 0x7fffffffll, "conv1_tc: too many segments");
    ++G.n;
  }
  for (int i = G.n; i < 9; ++i) G.seg_begin[i] = static_cast<int>(segs);
  if (segs == 0) return 0;
  static bool configured[kMaxDevices] = {false};
  const int dslot = kiri_cur_device_slot();
  if (!configured[dslot]) {
    KIRI_CHECK_CUDA(cudaFuncSetAttribute(conv1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
    configured[dslot] = true;
  }
  static const int ctas_per_sm = getenv("KIRI_CONV1_TC_CTAS") ? atoi(getenv("KIRI_CONV1_TC_CTAS")) : 4;
  long long grid = static_cast<long long>(gemm_tc_num_sms()) * ctas_per_sm;
  if (grid > segs) grid = segs;
  KIRI_CHECK_CUDA(launch_pdl(conv1_tc_kernel, dim3(static_cast<unsigned>(grid)), dim3(kTcThreads), kTcSmem, stream, G, H,
                             reinterpret_cast<const uint4*>(bmat_dev), static_cast<int>(segs)));
  return 0;
}

}  // namespace kiri

// Stand-alone entry (tests, tools): packs the weights and uploads them on every call (synchronous).
extern "C" int kiri_conv1_tc_multi(const uint8_t* const* planes_u8, void* const* out_bf16_nhwc48, const int* group_lines,
                                   const int* group_W, int n_groups, const float* w_host, const float* b_host, int H,
                                   cudaStream_t stream) {
  KIRI_REQUIRE(w_host && b_host, "kiri_conv1_tc_multi: null pointer");
  void* bmat = nullptr;
  KIRI_TRY(kiri::conv1_tc_build(w_host, b_host, &bmat));
  const int rc = kiri::conv1_tc_launch(planes_u8, out_bf16_nhwc48, group_lines, group_W, n_groups, bmat, H, stream);
  cudaStreamSynchronize(stream);
  cudaFree(bmat);
  return rc;
}
