// Host emulation of the shared-memory line FFT (csrc/fft_smem.cuh): the device functions are
// compiled for the CPU and the threads of a line are run pass by pass (compute + store for every
// thread, then the register reload for every thread -- what the barriers order on the GPU).
// Checks every radix schedule the kernels instantiate against a double-precision DFT, and that
// forward * inverse is M * identity.  Built and run by tests/test_host_fft.py (no GPU needed).
#include <cuda_runtime.h>

#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline void __syncthreads() {}
static inline void __syncwarp() {}

#include "../../jax-cfd_b200/csrc/fft_smem.cuh"

using namespace cfd;

static constexpr int ilog2c(int m) { return m <= 1 ? 0 : 1 + ilog2c(m / 2); }

template <class P, int DIR, int PASS>
struct Emu {
  static void run(std::vector<std::vector<float2>>& regs, std::vector<float2>& s, const float2* tw) {
    if constexpr (PASS < P::NP) {
      constexpr int LR = P::lr_fwd(PASS), LNS = P::lns_fwd(PASS);
      for (int t = 0; t < P::G; ++t) {
        float2(&v)[P::E] = *reinterpret_cast<float2(*)[P::E]>(regs[t].data());
        fft_pass_compute<P, LR, LNS, DIR>(v, t, tw + P::tw_off_fwd(PASS));
      }
      if constexpr (PASS + 1 < P::NP) {
        for (int t = 0; t < P::G; ++t) {
          float2(&v)[P::E] = *reinterpret_cast<float2(*)[P::E]>(regs[t].data());
          fft_pass_store<P, LR, LNS>(v, t, s.data());
        }
        for (int t = 0; t < P::G; ++t) {
          float2(&v)[P::E] = *reinterpret_cast<float2(*)[P::E]>(regs[t].data());
          fft_load_regs<P>(v, t, s.data());
        }
        Emu<P, DIR, PASS + 1>::run(regs, s, tw);
      }
    }
  }
};

template <class P>
static double check(const char* name, int lrmax, bool bal) {
  constexpr int M = P::M, G = P::G, E = P::E;
  std::vector<float2> tw = fft_build_twiddles(ilog2c(M), lrmax, bal);
  if ((int)tw.size() != P::tw_len()) {
    printf("%s: twiddle table length %d != plan %d\n", name, (int)tw.size(), P::tw_len());
    return 1.0;
  }
  std::vector<std::complex<double>> x(M);
  srand(1234 + M);
  for (auto& c : x) c = {rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5};
  std::vector<std::vector<float2>> regs(G, std::vector<float2>(E));
  for (int t = 0; t < G; ++t)
    for (int e = 0; e < E; ++e) regs[t][e] = make_float2((float)x[t + G * e].real(), (float)x[t + G * e].imag());
  std::vector<float2> s(P::pad(M) + 64);
  Emu<P, -1, 0>::run(regs, s, tw.data());
  // reference DFT in double
  double num = 0, den = 0;
  std::vector<std::complex<double>> X(M);
  for (int k = 0; k < M; ++k) {
    std::complex<double> acc = 0;
    for (int n = 0; n < M; ++n) {
      const double ang = -2.0 * M_PI * (double)((long long)k * n % M) / M;
      acc += x[n] * std::complex<double>(cos(ang), sin(ang));
    }
    X[k] = acc;
  }
  for (int t = 0; t < G; ++t)
    for (int e = 0; e < E; ++e) {
      const std::complex<double> got(regs[t][e].x, regs[t][e].y);
      num += std::norm(got - X[t + G * e]);
      den += std::norm(X[t + G * e]);
    }
  const double fwd = sqrt(num / den);
  // inverse of the forward result, straight from the registers (as xlines_kernel does)
  Emu<P, +1, 0>::run(regs, s, tw.data());
  num = den = 0;
  for (int t = 0; t < G; ++t)
    for (int e = 0; e < E; ++e) {
      const std::complex<double> got(regs[t][e].x / M, regs[t][e].y / M);
      num += std::norm(got - x[t + G * e]);
      den += std::norm(x[t + G * e]);
    }
  const double rt = sqrt(num / den);
  printf("%-28s M=%5d passes=%d  forward rel-L2 %.2e  round trip %.2e\n", name, M, P::NP, fwd, rt);
  return fwd > rt ? fwd : rt;
}

int main() {
  double worst = 0;
#define CHECK(LM, LE, LR, BAL)                                                          \
  {                                                                                     \
    const double e = check<FftPlan<LM, LE, LR, BAL>>("FftPlan<" #LM "," #LE "," #LR "," #BAL ">", LR, BAL); \
    worst = e > worst ? e : worst;                                                      \
  }
  CHECK(4, 4, 4, false)
  CHECK(5, 4, 4, false)
  CHECK(7, 4, 4, false)
  CHECK(9, 4, 4, false)
  CHECK(12, 4, 4, false)
  CHECK(12, 5, 4, false)
  CHECK(13, 4, 4, false)
  CHECK(13, 5, 4, false)
  CHECK(13, 5, 5, true)
  CHECK(12, 5, 5, true)
  CHECK(10, 5, 5, true)
  CHECK(14, 5, 5, true)
  CHECK(14, 4, 4, false)
  printf("worst %.3e\n", worst);
  return worst < 2e-6 ? 0 : 1;
}
