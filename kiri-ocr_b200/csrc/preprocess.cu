// K1: crop -> (invert if dark) -> Pillow-exact bilinear resize to the line height -> crop/pad
// to the batch width -> uint8 plane (+ optional normalised bf16).
//
// Replaces  OCR._preprocess_region        kiri_ocr/core.py:489-528
//           ResizeKeepRatioPadNoCrop      kiri_ocr/model.py:316-331
//           preprocess_pil                kiri_ocr/model.py:334-339   (per line, CPU, Pillow)
//
// One CTA per (crop, strip of output columns): a 640-wide line is five independent CTAs, so a
// 256-line batch fills the machine instead of leaving it latency-bound on 256 long-running CTAs.
// Pillow's resample is integer fixed point: float64 triangle weights,
// normalised, quantised to int(0.5 + w*2^22); each pass accumulates in integers, adds 2^21,
// shifts by 22 and clips to uint8.  The weights are recomputed on the device in float64 with
// explicitly rounded operations (no FMA contraction), which reproduces the host bit for bit.
// Horizontal pass first (source rows staged through shared memory with 32-bit loads), the
// intermediate lives in shared memory, then the vertical pass writes coalesced rows.
#include "internal.cuh"

#include <type_traits>

namespace kiri {

static constexpr int kPrecisionBits = 22;
static constexpr int kPreThreads = 256;
static constexpr int kRowBlock = 8;       // source rows staged per step

struct Coef {           // double-precision replica of Pillow's precompute_coeffs for one index
  int xmin, n;
};

__device__ __forceinline__ Coef coef_bounds(int xx, double scale, double support, int in_size,
                                            double& center) {
  center = __dmul_rn(static_cast<double>(xx) + 0.5, scale);
  int xmin = __double2int_rz(__dadd_rn(__dsub_rn(center, support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = __double2int_rz(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  Coef c;
  c.xmin = xmin;
  c.n = xmax - xmin;
  return c;
}
__device__ __forceinline__ double tri_weight(int x, int xmin, double center, double ss) {
  double arg = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss);
  arg = fabs(arg);
  return arg < 1.0 ? __dsub_rn(1.0, arg) : 0.0;
}
// Resample geometry of one axis: the two divisions of Pillow's precompute_coeffs, done ONCE per CTA (they were done by
// every coefficient thread: a double-precision division is ~100 instructions on this part).
struct AxisGeom { double scale, fs, ss; };
__device__ __forceinline__ AxisGeom axis_geom(int in_size, int out_size) {
  AxisGeom g;
  g.scale = __ddiv_rn(static_cast<double>(in_size), static_cast<double>(out_size));
  g.fs = g.scale < 1.0 ? 1.0 : g.scale;
  g.ss = __ddiv_rn(1.0, g.fs);
  return g;
}
// int(0.5 + (w / ww) * 2^22) with w / ww correctly rounded, as Pillow computes it - but through ONE reciprocal per output
// index: d' = w * RN(1/ww) is within 2^-51 of RN(w/ww), i.e. 0.5 + d' * 2^22 is within 1e-8 of the exact path, so the
// truncated integer can only differ when the value sits within 1e-8 of an integer; those (and anything within 1e-6) take
// the exact division.  Bit-exact with the division-per-tap form by construction.
__device__ __forceinline__ int quant_weight(double w, double ww, double inv_ww) {
  double v = __dadd_rn(0.5, __dmul_rn(__dmul_rn(w, inv_ww), 4194304.0));
  int q = __double2int_rz(v);
  const double f = v - static_cast<double>(q);
  if (f < 1e-6 || f > 1.0 - 1e-6) {
    v = __dadd_rn(0.5, __dmul_rn(__ddiv_rn(w, ww), 4194304.0));
    q = __double2int_rz(v);
  }
  return q;
}
// writes taps [0, ksize) for output index xx into k[tap * kstride]
__device__ __forceinline__ int fill_coefs(int xx, const AxisGeom& g, int in_size, int ksize, int* k, int kstride) {
  double center;
  const Coef c = coef_bounds(xx, g.scale, g.fs, in_size, center);
  double ww = 0.0;
  double wv[8];
  const bool small = c.n <= 8;                      // (every line-recognition shape: support <= 3.5 source pixels)
  if (small) {
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      wv[x] = 0.0;
      if (x < c.n) { wv[x] = tri_weight(x, c.xmin, center, g.ss); ww = __dadd_rn(ww, wv[x]); }
    }
  } else {
    for (int x = 0; x < c.n; ++x) ww = __dadd_rn(ww, tri_weight(x, c.xmin, center, g.ss));
  }
  const double inv_ww = ww != 0.0 ? __ddiv_rn(1.0, ww) : 0.0;
  if (small) {
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      if (x < ksize) {
        int q = 0;
        if (x < c.n) q = ww != 0.0 ? quant_weight(wv[x], ww, inv_ww) : __double2int_rz(__dadd_rn(0.5, __dmul_rn(wv[x], 4194304.0)));
        k[x * kstride] = q;
      }
    }
    for (int x = 8; x < ksize; ++x) k[x * kstride] = 0;
  } else {
    for (int x = 0; x < ksize; ++x) {
      int q = 0;
      if (x < c.n) {
        const double w = tri_weight(x, c.xmin, center, g.ss);
        q = ww != 0.0 ? quant_weight(w, ww, inv_ww) : __double2int_rz(__dadd_rn(0.5, __dmul_rn(w, 4194304.0)));
      }
      k[x * kstride] = q;
    }
  }
  return c.xmin;
}

__device__ __forceinline__ uint8_t clip8(int acc) {
  acc >>= kPrecisionBits;
  return static_cast<uint8_t>(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
}

// Pass 0 as its own launch: sum of every crop -> invert decision (core.py:524).  grid = (crops, 8):
// block y takes rows y, y+8, ...; one 64-bit atomic per warp.  Each source byte is read once here
// and once by the strip that resamples it (the old per-strip pass re-read the whole crop per strip).
__global__ void __launch_bounds__(kPreThreads)
crop_sum_kernel(const uint8_t* __restrict__ src, const KiriCropDesc* __restrict__ descs,
                unsigned long long* __restrict__ sums) {
  pdl_trigger();
  pdl_wait();
  const KiriCropDesc d = descs[blockIdx.x];
  const uint8_t* crop = src + d.src_offset;
  const int w = d.w, h = d.h;
  unsigned int local = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = kPreThreads >> 5;
  for (int r = blockIdx.y * nwarps + warp; r < h; r += gridDim.y * nwarps) {
    const uint8_t* row = crop + static_cast<size_t>(r) * d.pitch;
    const uintptr_t a = reinterpret_cast<uintptr_t>(row);
    const int head = static_cast<int>(a & 3);          // bytes before `row` in its first word
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(a - head);
    const int nwords = (head + w + 3) >> 2;
    for (int i = lane; i < nwords; i += 32) {
      uint32_t v = __ldg(wp + i);
      const int lo = i * 4 - head;                     // crop column of byte 0 of this word
      uint32_t mask = 0xffffffffu;
      if (lo < 0) mask &= 0xffffffffu << (8 * (-lo));
      if (lo + 4 > w) mask &= 0xffffffffu >> (8 * (lo + 4 - w));
      local = __dp4a(v & mask, 0x01010101u, local);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if (lane == 0 && local) atomicAdd(&sums[blockIdx.x], static_cast<unsigned long long>(local));
}

__global__ void __launch_bounds__(kPreThreads, 4)
preprocess_pack_kernel(const uint8_t* __restrict__ src, const KiriCropDesc* __restrict__ descs,
                       int img_h, uint8_t* __restrict__ planes,
                       __nv_bfloat16* __restrict__ norm_out, int smem_bytes,
                       const unsigned long long* __restrict__ sums) {
  extern __shared__ __align__(16) uint8_t sm[];
  pdl_trigger();
  pdl_wait();                                       // descriptors / source may come from a copy kernel
  const KiriCropDesc d = descs[blockIdx.x];
  const int tid = threadIdx.x;
  const int w = d.w, h = d.h, nw = d.nw, Wb = d.Wb;
  const int Wout = nw < Wb ? nw : Wb;
  const int strip0 = static_cast<int>(blockIdx.y) * d.strip_w;   // first output column of this CTA
  if (strip0 >= Wout) return;                                    // (whole CTA: no barrier is skipped)
  const uint8_t* crop = src + d.src_offset;
  uint8_t* plane = planes + d.out_offset;                       // the crop's [img_h, Wb] plane
  __nv_bfloat16* nplane = norm_out ? norm_out + d.out_offset : nullptr;

  // invert decision (core.py:524) from the crop sum of crop_sum_kernel
  const bool invert = !(d.flags & KIRI_CROP_NO_INVERT) && sums[blockIdx.x] < 127ull * static_cast<unsigned long long>(w) * h;
  const uint32_t inv_mask = invert ? 0xffffffffu : 0u;

  // ---------------- coefficient geometry ----------------
  const bool do_h = (nw != w);
  const bool do_v = (h != img_h);
  __shared__ AxisGeom geom[2];                     // [0] horizontal (w -> nw), [1] vertical (h -> img_h)
  if (tid == 0) geom[0] = axis_geom(w, nw);
  if (tid == 32) geom[1] = axis_geom(h, img_h);
  __syncthreads();
  const int ksize_h = do_h ? (static_cast<int>(ceil(geom[0].fs)) * 2 + 1) : 1;
  const int ksize_v = do_v ? (static_cast<int>(ceil(geom[1].fs)) * 2 + 1) : 1;
  const int Ws = d.strip_w;                               // output columns per strip (host-chosen)

  // shared-memory carve (all 16-byte aligned)
  int off = 0;
  int* kv = reinterpret_cast<int*>(sm + off);      off += ((img_h * ksize_v * 4 + 15) & ~15);
  int* ymin = reinterpret_cast<int*>(sm + off);    off += ((img_h * 4 + 15) & ~15);
  int* kh = reinterpret_cast<int*>(sm + off);      off += ((Ws * ksize_h * 4 + 15) & ~15);
  int* xmin = reinterpret_cast<int*>(sm + off);    off += ((Ws * 4 + 15) & ~15);
  const int Wi = (Ws + 3) & ~3;                           // row pitch of the intermediate (whole 32-bit words)
  uint8_t* inter = sm + off;                       off += ((h * Wi + 15) & ~15);
  uint8_t* srow = sm + off;                        // staged source rows: as many as the budget holds
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int kWarps = kPreThreads / 32;

  if (do_v) {
    const AxisGeom gv = geom[1];
    for (int y = tid; y < img_h; y += kPreThreads) ymin[y] = fill_coefs(y, gv, h, ksize_v, kv + y * ksize_v, 1);
  }

  const int c0 = strip0;
  const int cw = (Wout - c0) < Ws ? (Wout - c0) : Ws;
  // horizontal coefficients of this strip, tap-major so lanes hit consecutive banks
  if (do_h) {
    const AxisGeom gh = geom[0];
    for (int x = tid; x < cw; x += kPreThreads) xmin[x] = fill_coefs(c0 + x, gh, w, ksize_h, kh + x, Ws);
  } else {
    for (int x = tid; x < cw; x += kPreThreads) xmin[x] = c0 + x;
  }
  __syncthreads();
  const int sx0 = xmin[0];                                  // first source column needed
  const int sx1 = do_h ? (xmin[cw - 1] + ksize_h) : (c0 + cw);
  const int span = (sx1 < w ? sx1 : w) - sx0;               // source columns to stage
  // all rows of the strip in ONE staging phase when they fit (typical lines: 3 barriers per CTA
  // instead of 2 per 8 rows, and every load of the strip in flight at once)
  const int srow_cap = (span + 4 + 3 + 15) & ~15;           // bytes per staged row (4-byte alignment head + tail)
  int rows_blk = (smem_bytes - off - 16) / srow_cap;         // 16 bytes of slack: zero-weight taps may read past the last row
  if (rows_blk > h) rows_blk = h;
  if (rows_blk < 1) rows_blk = 1;
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(crop + sx0);
  const uint32_t srow_a = smem_u32(srow), inter_a = smem_u32(inter), kv_a = smem_u32(kv);   // shared-window addresses for the inner loops
  const int head0 = static_cast<int>(a0 & 3), dhead = d.pitch & 3;   // a staged row starts `head` bytes into its first word

  for (int r0 = 0; r0 < h; r0 += rows_blk) {
    const int nr = (h - r0) < rows_blk ? (h - r0) : rows_blk;
    // stage rows r0..r0+nr, columns sx0..sx0+span, inverted if needed: a warp per row, 32-bit loads from the
    // word-aligned address at or below the row's first byte
    for (int rr = warp; rr < nr; rr += kWarps) {
      const int head = (head0 + (r0 + rr) * dhead) & 3;
      const uint32_t* gw = reinterpret_cast<const uint32_t*>(a0 + static_cast<uintptr_t>(r0 + rr) * d.pitch - head);
      const uint32_t sa = srow_a + rr * srow_cap;
      const int nwords = (head + span + 3) >> 2;
      for (int i = lane; i < nwords; i += 32) sts_u32(sa + 4 * i, __ldg(gw + i) ^ inv_mask);
    }
    __syncthreads();
    // horizontal pass -> inter[r][x].  A thread owns ONE output column (x = tid mod 128) and walks
    // the rows: xmin / tap count / up to 8 coefficients live in registers, no per-pixel division.
    {
      const int x = tid & 127, rg = tid >> 7;                // strips are <= 128 columns wide
      if (x < cw) {
        const int xm = xmin[x];
        const int kmax = do_h ? (ksize_h < w - xm ? ksize_h : w - xm) : 1;   // taps beyond the row carry weight 0
        int kc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) kc[k] = (do_h && k < kmax) ? kh[k * Ws + x] : 0;
        const uint32_t sb = srow_a + (xm - sx0);
        const uint32_t ib = inter_a + r0 * Wi + x;
        // the row loop, specialised on the tap count (3 when enlarging, 5 up to 2x reduction, 7 up to 3x): taps past kmax
        // carry coefficient 0 and read bytes inside the staged row's slack, so the unrolled form needs no predicates
        auto rows = [&](auto ks_tag) {
          constexpr int KS = decltype(ks_tag)::value;
          for (int rr = rg; rr < nr; rr += kPreThreads / 128) {
            const uint32_t s = sb + rr * srow_cap + ((head0 + (r0 + rr) * dhead) & 3);
            uint32_t o;
            if (KS == 0) {                                        // no horizontal resampling
              o = lds_u8(s);
            } else if (KS <= 8) {
              int acc = 1 << (kPrecisionBits - 1);
#pragma unroll
              for (int k = 0; k < KS; ++k) acc += static_cast<int>(lds_u8(s + k)) * kc[k];
              o = clip8(acc);
            } else {                                              // generic tap count
              int acc = 1 << (kPrecisionBits - 1);
              for (int k = 0; k < kmax; ++k) acc += static_cast<int>(lds_u8(s + k)) * kh[k * Ws + x];
              o = clip8(acc);
            }
            sts_u8(ib + rr * Wi, o);
          }
        };
        if (!do_h) rows(std::integral_constant<int, 0>{});
        else if (ksize_h == 3) rows(std::integral_constant<int, 3>{});
        else if (ksize_h == 5) rows(std::integral_constant<int, 5>{});
        else if (ksize_h == 7) rows(std::integral_constant<int, 7>{});
        else rows(std::integral_constant<int, 9>{});
      }
    }
    __syncthreads();
  }
  // vertical pass -> plane rows.  A lane owns FOUR adjacent columns (one 32-bit word of the intermediate, one 32-bit store),
  // a warp walks output rows: the row's taps and coefficients are warp-uniform, one shared load feeds four accumulators.
  {
    const int nwq = (cw + 3) >> 2, wq = Wi >> 2;
    const int col = c0 + 4 * lane;                            // first of this lane's columns in the plane
    if (lane < nwq) {
      for (int y = warp; y < img_h; y += kWarps) {
        uint32_t word;
        if (do_v) {
          const int y0 = ymin[y];
          const uint32_t kr = kv_a + 4 * y * ksize_v;           // taps past the last source row carry coefficient 0 (fill_coefs)
          const uint32_t cp = inter_a + 4 * (y0 * wq + lane);   // ... and read rows that lie in the staging area behind `inter`
          int a[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) a[j] = 1 << (kPrecisionBits - 1);
          auto mac = [&](int kk) {
            const uint32_t v = lds_u32(cp + 4 * kk * wq);
            const int c = static_cast<int>(lds_u32(kr + 4 * kk));
            a[0] += static_cast<int>(v & 255u) * c;
            a[1] += static_cast<int>((v >> 8) & 255u) * c;
            a[2] += static_cast<int>((v >> 16) & 255u) * c;
            a[3] += static_cast<int>(v >> 24) * c;
          };
          auto taps = [&](auto ks_tag) {                          // specialised on the tap count like the horizontal pass
            constexpr int KS = decltype(ks_tag)::value;
            if (KS > 0) {
#pragma unroll
              for (int k = 0; k < KS; ++k) mac(k);
            } else {
              const int n = ksize_v < h - y0 ? ksize_v : h - y0;
              for (int k = 0; k < n; ++k) mac(k);
            }
          };
          if (ksize_v == 3) taps(std::integral_constant<int, 3>{});
          else if (ksize_v == 5) taps(std::integral_constant<int, 5>{});
          else if (ksize_v == 7) taps(std::integral_constant<int, 7>{});
          else taps(std::integral_constant<int, 0>{});
          word = static_cast<uint32_t>(clip8(a[0])) | (static_cast<uint32_t>(clip8(a[1])) << 8) |
                 (static_cast<uint32_t>(clip8(a[2])) << 16) | (static_cast<uint32_t>(clip8(a[3])) << 24);
        } else {
          word = lds_u32(inter_a + 4 * (y * wq + lane));
        }
        // columns of the last word past the resized width are padding (gray 128, model.py:329-330)
        if (col + 4 > Wout) {
#pragma unroll
          for (int j = 1; j < 4; ++j)
            if (col + j >= Wout) word = (word & ~(255u << (8 * j))) | (128u << (8 * j));
        }
        *reinterpret_cast<uint32_t*>(plane + y * Wb + col) = word;
        if (nplane) {
          __nv_bfloat162 lo, hi;
          auto nrm = [](uint32_t o) { return __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(o), 255.0f), 0.5f), 0.5f); };
          lo = __floats2bfloat162_rn(nrm(word & 255u), nrm((word >> 8) & 255u));
          hi = __floats2bfloat162_rn(nrm((word >> 16) & 255u), nrm(word >> 24));
          uint2 pkd;
          pkd.x = *reinterpret_cast<uint32_t*>(&lo);
          pkd.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(nplane + y * Wb + col) = pkd;
        }
      }
    }
  }
  // gray-128 padding on the right (model.py:329-330), whole words from the first 4-aligned column past the text
  const int pad0 = (Wout + 3) & ~3;
  if (pad0 < Wb && blockIdx.y == 0) {
    const int padq = (Wb - pad0) >> 2;
    const float f = __fdiv_rn(__fsub_rn(__fdiv_rn(128.0f, 255.0f), 0.5f), 0.5f);
    const __nv_bfloat162 fb2 = __floats2bfloat162_rn(f, f);
    const uint32_t fbw = *reinterpret_cast<const uint32_t*>(&fb2);
    for (int y = warp; y < img_h; y += kWarps) {
      uint32_t* pw = reinterpret_cast<uint32_t*>(plane + y * Wb + pad0);
      for (int i = lane; i < padq; i += 32) pw[i] = 0x80808080u;
      if (nplane) {
        uint2* np = reinterpret_cast<uint2*>(nplane + y * Wb + pad0);
        for (int i = lane; i < padq; i += 32) np[i] = make_uint2(fbw, fbw);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Page ingest: BGR -> gray on the device, bit-exact with cv2.cvtColor(img, COLOR_BGR2GRAY) on uint8 (core.py:762-766).
// OpenCV (imgproc/src/color_rgb.simd.hpp, RGB2Gray<uchar>; opencv-python is a third-party dependency of the reference,
// 4.13.0 here) computes   gray = (B*3735 + G*19235 + R*9798 + (1 << 14)) >> 15   - 15-bit fixed-point BT.601 weights.
// One thread = 4 pixels: three 32-bit loads (12 bytes of BGR), one 32-bit store.  HBM-bound: 3 B in + 1 B out per pixel.
__global__ void __launch_bounds__(256)
bgr_to_gray_kernel(const uint32_t* __restrict__ bgr, uint32_t* __restrict__ gray, long long n_quads, const uint8_t* bgr_tail,
                   uint8_t* gray_tail, int tail_pixels) {
  pdl_trigger();
  pdl_wait();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n_quads) {
    const uint32_t w0 = __ldg(bgr + 3 * i), w1 = __ldg(bgr + 3 * i + 1), w2 = __ldg(bgr + 3 * i + 2);
    // bytes: w0 = B0 G0 R0 B1 | w1 = G1 R1 B2 G2 | w2 = R2 B3 G3 R3  (little endian)
    auto y = [](uint32_t b, uint32_t g, uint32_t r) { return (b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15; };
    const uint32_t p0 = y(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u);
    const uint32_t p1 = y(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u);
    const uint32_t p2 = y((w1 >> 16) & 255u, w1 >> 24, w2 & 255u);
    const uint32_t p3 = y((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24);
    gray[i] = p0 | (p1 << 8) | (p2 << 16) | (p3 << 24);
  }
  if (i == 0) {
    for (int t = 0; t < tail_pixels; ++t)
      gray_tail[t] = static_cast<uint8_t>((bgr_tail[3 * t] * 3735u + bgr_tail[3 * t + 1] * 19235u + bgr_tail[3 * t + 2] * 9798u + 16384u) >> 15);
  }
}

}  // namespace kiri

using namespace kiri;

extern "C" int kiri_bgr_to_gray(const uint8_t* bgr_u8, long long n_pixels, uint8_t* gray_u8, cudaStream_t stream) {
  KIRI_REQUIRE(bgr_u8 && gray_u8, "kiri_bgr_to_gray: null pointer");
  KIRI_REQUIRE(n_pixels >= 0, "kiri_bgr_to_gray: negative size");
  KIRI_REQUIRE((reinterpret_cast<uintptr_t>(bgr_u8) & 3) == 0 && (reinterpret_cast<uintptr_t>(gray_u8) & 3) == 0,
               "kiri_bgr_to_gray: buffers must be 4-byte aligned");
  if (n_pixels == 0) return 0;
  const long long quads = n_pixels / 4;
  const int tail = static_cast<int>(n_pixels - quads * 4);
  const long long blocks = (quads + 255) / 256;
  KIRI_REQUIRE(blocks < 0x7fffffffll, "kiri_bgr_to_gray: image too large");
  KIRI_CHECK_CUDA(launch_pdl(bgr_to_gray_kernel, dim3(static_cast<unsigned>(blocks > 0 ? blocks : 1)), dim3(256), 0, stream,
                             reinterpret_cast<const uint32_t*>(bgr_u8), reinterpret_cast<uint32_t*>(gray_u8), quads,
                             bgr_u8 + quads * 12, gray_u8 + quads * 4, tail));
  return 0;
}

extern "C" int kiri_preprocess_smem_bytes(int w, int h, int nw, int img_h, int Wb, int strip_w) {
  const int Wout = nw < Wb ? nw : Wb;
  const int Ws = strip_w < Wout ? strip_w : Wout;
  const double hs = (double)w / nw, vs = (double)h / img_h;
  const int ksh = (nw != w) ? ((int)ceil(hs < 1.0 ? 1.0 : hs) * 2 + 1) : 1;
  const int ksv = (h != img_h) ? ((int)ceil(vs < 1.0 ? 1.0 : vs) * 2 + 1) : 1;
  long long off = 0;
  off += ((long long)img_h * ksv * 4 + 15) & ~15ll;
  off += ((long long)img_h * 4 + 15) & ~15ll;
  off += ((long long)Ws * ksh * 4 + 15) & ~15ll;
  off += ((long long)Ws * 4 + 15) & ~15ll;
  off += ((long long)h * ((Ws + 3) & ~3) + 15) & ~15ll;
  // staged source rows: span of a strip plus slack for the 4-byte alignment head and taps
  const long long span = (long long)ceil((hs < 1.0 ? 1.0 : hs) * Ws) + 2 * ksh + 8;
  const long long per_row = (span + 4 + 15) & ~15ll;
  off += per_row * kRowBlock + 16 * kRowBlock;
  return off > 0x7fffffff ? 0x7fffffff : (int)off;
}

extern "C" int kiri_preprocess_pack(const uint8_t* src, const KiriCropDesc* descs_dev, int n_crops,
                                    int img_h, int smem_bytes, int max_strips, uint8_t* planes_u8,
                                    void* norm_bf16, unsigned long long* crop_sums, cudaStream_t stream) {
  KIRI_REQUIRE(src && descs_dev && planes_u8 && crop_sums, "kiri_preprocess_pack: null pointer");
  KIRI_REQUIRE(n_crops >= 0 && img_h > 0 && max_strips >= 1 && max_strips <= 65535, "kiri_preprocess_pack: bad sizes");
  if (n_crops == 0) return 0;
  static int max_optin_dev[kMaxDevices] = {0};
  int& max_optin = max_optin_dev[kiri_cur_device_slot()];
  if (!max_optin) {
    int dev = 0;
    KIRI_CHECK_CUDA(cudaGetDevice(&dev));
    int optin = 0;
    KIRI_CHECK_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cudaFuncAttributes fa;
    KIRI_CHECK_CUDA(cudaFuncGetAttributes(&fa, preprocess_pack_kernel));
    optin -= static_cast<int>(fa.sharedSizeBytes);           // static + dynamic must fit the opt-in limit
    KIRI_CHECK_CUDA(cudaFuncSetAttribute(preprocess_pack_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    max_optin = optin;
  }
  KIRI_REQUIRE(smem_bytes <= max_optin, "kiri_preprocess_pack: %d bytes of shared memory requested, %d available",
               smem_bytes, max_optin);
  ProfScope ps(PS_PREPROCESS, stream);
  KIRI_CHECK_CUDA(cudaMemsetAsync(crop_sums, 0, sizeof(unsigned long long) * n_crops, stream));
  KIRI_CHECK_CUDA(launch_pdl(crop_sum_kernel, dim3(n_crops, 8), dim3(kPreThreads), 0, stream, src, descs_dev, crop_sums));
  KIRI_CHECK_CUDA(launch_pdl(preprocess_pack_kernel, dim3(n_crops, max_strips), dim3(kPreThreads), smem_bytes, stream,
                             src, descs_dev, img_h, planes_u8, reinterpret_cast<__nv_bfloat16*>(norm_bf16), smem_bytes,
                             static_cast<const unsigned long long*>(crop_sums)));
  return 0;
}
