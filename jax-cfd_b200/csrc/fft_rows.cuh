// Helpers shared by the row / line FFT kernels (poisson_2d.cu, poisson_3d.cu).
#pragma once
#include <stdlib.h>

#include <mutex>
#include <type_traits>
#include <unordered_map>

#include "common.cuh"
#include "fft_smem.cuh"

namespace cfd {
namespace {

__host__ __device__ constexpr int row_stride(int M, int rows) {
  // padded line length rounded up to 16 float2, plus a skew so that the transposed access
  // (lane -> row fastest) is bank-conflict free.
  return ((padded_len(M) + 15) & ~15) + (rows >= 16 ? 1 : 16 / rows);
}

// Opt-in to > 48 KB of dynamic shared memory, once per kernel (and per larger size): the attribute
// call is kept off the launch path -- every later launch, including the ones recorded into a CUDA
// graph (plan.cu), is then a plain launch.
template <typename K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    static std::mutex mu;
    static std::unordered_map<const void*, size_t> done[64];  // per device, keyed by kernel
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    size_t& have = done[dev & 63][reinterpret_cast<const void*>(kernel)];
    if (bytes > have) {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
      if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(smem)", e, __FILE__, __LINE__);
      have = bytes;
    }
  }
  return 0;
}

// rows per CTA for the row kernels: as many as fit ~100 KB / 1024 threads, at most 32
constexpr int rows_for(int LM) {
  const int M = 1 << LM, G = (M < 16 ? 1 : M / 16);
  int rows = 32;
  while (rows > 1 && (rows * G > 1024 || rows * (long)row_stride(M, 16) * 8 > 140 * 1024)) rows /= 2;
  return rows;
}
constexpr int lines_for(int LM) {
  const int M = 1 << LM, G = (M < 16 ? 1 : M / 16);
  int lines = 16;
  while (lines > 1 && (lines * G > 256)) lines /= 2;
  return lines;
}

// points per thread (log2) of the x-line kernel.  4: 16 points, one radix-16 butterfly per pass.
// 5: 32 points = two radix-16 butterflies per pass, half the threads per line, which lets TWO
// independent CTAs share an SM for 4096- and 8192-point lines (their exchange / barrier phases then
// overlap each other): 277 -> 221 us at 8192^2, 268 -> 205 us for 4096-point lines.  Shorter lines
// already run several lines per SM and get slower; 16384-point lines stay at one CTA per SM either
// way.  CFD_XLINES_LE=4|5 overrides.
inline int xlines_lemax(int lm) {
  static const int forced = [] {
    const char* e = getenv("CFD_XLINES_LE");
    return e ? atoi(e) : 0;
  }();
  if (forced == 4 || forced == 5) return forced;
  return (lm == 12 || lm == 13) ? 5 : 4;
}

// Balanced radix schedule with one radix-32 pass (FftPlan<LM, 5, 5, true>: 8192 points = 32 x 16 x
// 16, three passes and two exchanges per direction instead of 16 x 16 x 16 x 2 with three) for
// the 8192-point x lines, whose kernel is bound by the shared-memory / L1 data pipe (63 % of its
// peak in round 1, two thirds of it exchange traffic).  CFD_XLINES_BAL=0|1 overrides.
inline bool xlines_balanced(int lm) {
  static const int forced = [] {
    const char* e = getenv("CFD_XLINES_BAL");
    return e ? atoi(e) : -1;
  }();
  if (xlines_lemax(lm) != 5) return false;
  if (forced == 0 || forced == 1) return forced == 1 && lm == 13;
  return lm == 13;
}

// points per thread (log2) of the row kernels: 32 points per thread measured slower there (218 vs
// 190 us, 268 vs 237 us at 8192^2 -- the rows of a CTA already drift apart on their own named
// barriers); CFD_ROWS_LE=4|5 overrides
inline int rows_lemax(int lm) {
  static const int forced = [] {
    const char* e = getenv("CFD_ROWS_LE");
    return e ? atoi(e) : 0;
  }();
  if (forced == 4 || forced == 5) return forced;
  (void)lm;
  return 4;
}

static int rows_shift() {  // tuning knob: CFD_FFT_ROWS_SHIFT=k uses rows_for(LM) >> k rows per CTA
  static const int v = [] {
    const char* e = getenv("CFD_FFT_ROWS_SHIFT");
    return e ? atoi(e) : 0;
  }();
  return v;
}

}  // namespace
}  // namespace cfd

#define CFD_DISPATCH_LM(lm, LO, HI, CALL)                       \
  switch (lm) {                                                 \
    case 4: { constexpr int LM_ = 4; CALL; } break;             \
    case 5: { constexpr int LM_ = 5; CALL; } break;             \
    case 6: { constexpr int LM_ = 6; CALL; } break;             \
    case 7: { constexpr int LM_ = 7; CALL; } break;             \
    case 8: { constexpr int LM_ = 8; CALL; } break;             \
    case 9: { constexpr int LM_ = 9; CALL; } break;             \
    case 10: { constexpr int LM_ = 10; CALL; } break;           \
    case 11: { constexpr int LM_ = 11; CALL; } break;           \
    case 12: { constexpr int LM_ = 12; CALL; } break;           \
    case 13: { constexpr int LM_ = 13; CALL; } break;           \
    case 14: { constexpr int LM_ = 14; CALL; } break;           \
    default: return set_error_msg("unsupported FFT length (need 2^4 .. 2^14 complex points)"); \
  }

