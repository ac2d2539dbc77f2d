// Greedy attention decoder (K11-K12) — see include/kiri_b200.h.  Filled in below.
#include "common.cuh"
#include "gemm_tc.cuh"
#include "kiri_b200.h"

using namespace kiri;

extern "C" size_t kiri_decode_workspace_bytes(const KiriHandle*, int, int, int) { return 0; }
extern "C" int kiri_decode_greedy(KiriHandle*, const void*, const int*, int, int, int, const KiriDecodeParams*,
                                  void*, size_t, int*, int*, float*, float*, float*, const int*, int*, int,
                                  cudaStream_t) {
  KIRI_REQUIRE(false, "kiri_decode_greedy: not built yet");
}
