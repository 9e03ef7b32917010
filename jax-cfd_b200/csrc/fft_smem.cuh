// Shared-memory line FFTs for the pressure solve (replaces jnp.fft.rfftn / irfftn at
// /root/reference/jax_cfd/base/fast_diagonalization.py:223).
//
// A line of M = 2^LM complex points is owned by G = M / E threads (E = points per thread, 16
// unless M < 16).  Between radix passes the points live in REGISTERS; shared memory is used only
// to exchange them (register-staged in-place Stockham: every thread reads its E inputs, the group
// synchronises, every thread writes its E outputs to the auto-sort positions).  Thread t always
// reads logical positions {t + G*e}, so reads are conflict-free; writes are made conflict-free by
// padding one float2 per 16 (PAD()).
//
// Twiddles come from a per-M table built in double precision on the host (plan.cu), laid out
// per pass as tw[(r-1) * Ns + k] = exp(-2 pi i r k / (Ns R)) so that a warp reads consecutive k.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include <vector>

namespace cfd {

__host__ __device__ constexpr int PAD(int i) { return i + (i >> 4); }
__host__ __device__ constexpr int padded_len(int m) { return m + (m >> 4); }

// Optional: complex add / subtract as ONE packed FP32 instruction (FADD2 on sm_100a:
// add.rn.f32x2 works on an aligned register pair, which is what a float2 already is; same IEEE
// rounding as two FADDs).  Measured on B200: 13 % fewer instructions in the line kernels but the
// FADD2 issues at half rate on the FMA pipe, and the 8192^2 step got 1 % SLOWER (1186 vs 1170 us)
// -- the line FFTs are not issue-slot bound.  Off by default; -DCFD_PACKED_FP=1 to retry.
#ifndef CFD_PACKED_FP
#define CFD_PACKED_FP 0
#endif
typedef unsigned long long pk2;
__device__ __forceinline__ pk2 pack2(float2 a) {
  pk2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ float2 unpack2(pk2 r) {
  float2 a;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
  return a;
}
#if CFD_PACKED_FP
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  pk2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack2(a)), "l"(pack2(b)));
  return unpack2(r);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  pk2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack2(a)), "l"(pack2(b)));
  return unpack2(r);
}
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// a * conj(b)
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
// DIR = -1: forward (multiply by -i);  DIR = +1: inverse (multiply by +i)
template <int DIR>
__device__ __forceinline__ float2 mul_dir_i(float2 a) {
  return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}
template <int DIR>
__device__ __forceinline__ float2 twmul(float2 a, float2 w) {  // w is the FORWARD twiddle
  return DIR < 0 ? cmul(a, w) : cmulc(a, w);
}

// ---- in-register DFTs, natural-order in, natural-order out --------------------------------
template <int DIR>
__device__ __forceinline__ void dft2(float2& a0, float2& a1) {
  float2 t = a0;
  a0 = cadd(t, a1);
  a1 = csub(t, a1);
}

template <int DIR>
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 t0 = cadd(a0, a2), t1 = csub(a0, a2);
  float2 t2 = cadd(a1, a3), t3 = mul_dir_i<DIR>(csub(a1, a3));
  a0 = cadd(t0, t2);
  a1 = cadd(t1, t3);
  a2 = csub(t0, t2);
  a3 = csub(t1, t3);
}

template <int DIR>
__device__ __forceinline__ void dft8(float2 (&a)[8]) {
  constexpr float H = 0.70710678118654752440f;
  dft4<DIR>(a[0], a[2], a[4], a[6]);  // E[k] in a[0],a[2],a[4],a[6]
  dft4<DIR>(a[1], a[3], a[5], a[7]);  // O[k] in a[1],a[3],a[5],a[7]
  float2 o1, o2, o3;
  if (DIR < 0) {
    o1 = make_float2((a[3].x + a[3].y) * H, (a[3].y - a[3].x) * H);
    o3 = make_float2((a[7].y - a[7].x) * H, (-a[7].x - a[7].y) * H);
  } else {
    o1 = make_float2((a[3].x - a[3].y) * H, (a[3].x + a[3].y) * H);
    o3 = make_float2((-a[7].x - a[7].y) * H, (a[7].x - a[7].y) * H);
  }
  o2 = mul_dir_i<DIR>(a[5]);
  float2 e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6], o0 = a[1];
  a[0] = cadd(e0, o0);
  a[4] = csub(e0, o0);
  a[1] = cadd(e1, o1);
  a[5] = csub(e1, o1);
  a[2] = cadd(e2, o2);
  a[6] = csub(e2, o2);
  a[3] = cadd(e3, o3);
  a[7] = csub(e3, o3);
}

template <int DIR>
__device__ __forceinline__ void dft16(float2 (&a)[16]) {
  float2 e[8], o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    e[i] = a[2 * i];
    o[i] = a[2 * i + 1];
  }
  dft8<DIR>(e);
  dft8<DIR>(o);
  // forward twiddles exp(-2 pi i k / 16), k = 1..7
  constexpr float C1 = 0.92387953251128675613f, S1 = 0.38268343236508977173f;
  constexpr float H = 0.70710678118654752440f;
  const float2 w1 = make_float2(C1, -S1), w2 = make_float2(H, -H), w3 = make_float2(S1, -C1);
  const float2 w5 = make_float2(-S1, -C1), w6 = make_float2(-H, -H), w7 = make_float2(-C1, -S1);
  o[1] = twmul<DIR>(o[1], w1);
  o[2] = twmul<DIR>(o[2], w2);
  o[3] = twmul<DIR>(o[3], w3);
  o[4] = mul_dir_i<DIR>(o[4]);
  o[5] = twmul<DIR>(o[5], w5);
  o[6] = twmul<DIR>(o[6], w6);
  o[7] = twmul<DIR>(o[7], w7);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    a[k] = cadd(e[k], o[k]);
    a[k + 8] = csub(e[k], o[k]);
  }
}

// 32 points: two 16-point transforms of the even / odd inputs + one twiddle level
template <int DIR>
__device__ __forceinline__ void dft32(float2 (&a)[32]) {
  float2 e[16], o[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    e[i] = a[2 * i];
    o[i] = a[2 * i + 1];
  }
  dft16<DIR>(e);
  dft16<DIR>(o);
  // forward twiddles exp(-2 pi i k / 32), k = 1..15 (k = 8 is -i)
  constexpr float C1 = 0.98078528040323044913f, S1 = 0.19509032201612826785f;
  constexpr float C2 = 0.92387953251128675613f, S2 = 0.38268343236508977173f;
  constexpr float C3 = 0.83146961230254523708f, S3 = 0.55557023301960222474f;
  constexpr float H = 0.70710678118654752440f;
  o[1] = twmul<DIR>(o[1], make_float2(C1, -S1));
  o[2] = twmul<DIR>(o[2], make_float2(C2, -S2));
  o[3] = twmul<DIR>(o[3], make_float2(C3, -S3));
  o[4] = twmul<DIR>(o[4], make_float2(H, -H));
  o[5] = twmul<DIR>(o[5], make_float2(S3, -C3));
  o[6] = twmul<DIR>(o[6], make_float2(S2, -C2));
  o[7] = twmul<DIR>(o[7], make_float2(S1, -C1));
  o[8] = mul_dir_i<DIR>(o[8]);
  o[9] = twmul<DIR>(o[9], make_float2(-S1, -C1));
  o[10] = twmul<DIR>(o[10], make_float2(-S2, -C2));
  o[11] = twmul<DIR>(o[11], make_float2(-S3, -C3));
  o[12] = twmul<DIR>(o[12], make_float2(-H, -H));
  o[13] = twmul<DIR>(o[13], make_float2(-C3, -S3));
  o[14] = twmul<DIR>(o[14], make_float2(-C2, -S2));
  o[15] = twmul<DIR>(o[15], make_float2(-C1, -S1));
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    a[k] = cadd(e[k], o[k]);
    a[k + 16] = csub(e[k], o[k]);
  }
}

template <int R, int DIR>
struct Dft;
template <int DIR>
struct Dft<2, DIR> {
  static __device__ __forceinline__ void run(float2 (&a)[2]) { dft2<DIR>(a[0], a[1]); }
};
template <int DIR>
struct Dft<4, DIR> {
  static __device__ __forceinline__ void run(float2 (&a)[4]) { dft4<DIR>(a[0], a[1], a[2], a[3]); }
};
template <int DIR>
struct Dft<8, DIR> {
  static __device__ __forceinline__ void run(float2 (&a)[8]) { dft8<DIR>(a); }
};
template <int DIR>
struct Dft<16, DIR> {
  static __device__ __forceinline__ void run(float2 (&a)[16]) { dft16<DIR>(a); }
};
template <int DIR>
struct Dft<32, DIR> {
  static __device__ __forceinline__ void run(float2 (&a)[32]) { dft32<DIR>(a); }
};

// ---- radix plan -------------------------------------------------------------------------------
// M = 2^LM points, E = min(2^LEMAX, M) points per thread, radix min(2^LRMAX, E) (defaults: 16
// points, radix-16 passes; <LM, 5, 4>: 32 points = two radix-16 butterflies per thread and pass).
// Passes use the largest radix while possible and one final smaller radix.  Forward and inverse
// share the pass order and the twiddle table (the inverse conjugates).  Every transform starts from
// registers v[e] = x[t + G*e] and ends with v[slot] = X[t + G*slot], so a forward transform can be
// scaled in registers and fed straight into an inverse transform without an exchange.
// Pass schedule of a 2^lm-point transform with radices up to 2^lr (shared by the kernels and by the
// host code that lays out the twiddle table, plan.cu).
//   bal = false: full-radix passes first, one final smaller radix        (13, 4) -> 16 16 16 2
//   bal = true:  the fewest passes, radices as even as possible, larger first (13, 5) -> 32 16 16
__host__ __device__ constexpr int fft_sched_np(int lm, int lr) { return (lm + lr - 1) / lr; }
__host__ __device__ constexpr int fft_sched_lr(int lm, int lr, bool bal, int p) {
  if (!bal) return (p < lm / lr) ? lr : (lm - (lm / lr) * lr);
  const int np = fft_sched_np(lm, lr);
  return lm / np + (p < lm % np ? 1 : 0);
}

// Host side: twiddle table of a 2^lm-point transform with the schedule of FftPlan<lm, ., lrmax, bal>:
// pass p owns (R - 1) * Ns entries, tw[(r - 1) * Ns + k] = exp(-2 pi i r k / (Ns R)), computed in
// double precision and rounded once.
inline std::vector<float2> fft_build_twiddles(int lm, int lrmax = 4, bool bal = false) {
  const int lr = lm < lrmax ? lm : lrmax;
  const int np = fft_sched_np(lm, lr);
  std::vector<float2> tw;
  int lns = 0;
  for (int p = 0; p < np; ++p) {
    const int lrp = fft_sched_lr(lm, lr, bal, p);
    const int R = 1 << lrp, Ns = 1 << lns;
    for (int r = 1; r < R; ++r)
      for (int k = 0; k < Ns; ++k) {
        const double ang = -2.0 * 3.14159265358979323846 * (double)r * (double)k / ((double)Ns * (double)R);
        tw.push_back(make_float2((float)cos(ang), (float)sin(ang)));
      }
    lns += lrp;
  }
  return tw;
}

template <int LM, int LEMAX = 4, int LRMAX = LEMAX, bool BAL = false>
struct FftPlan {
  static constexpr int M = 1 << LM;
  static constexpr int LE = LM < LEMAX ? LM : LEMAX;
  static constexpr int E = 1 << LE;
  static constexpr int G = M / E;  // threads per line
  // radix of the full passes: 2^LR <= E.  LR < LE: a thread runs E / 2^LR independent butterflies
  // per pass (more instruction-level parallelism per thread, half the threads per line)
  static constexpr int LR = LE < LRMAX ? LE : LRMAX;
  static constexpr int NP = fft_sched_np(LM, LR);  // passes
  static constexpr int LPAD = LR < 4 ? 4 : LR;   // one padding slot per 2^LPAD points
  __host__ __device__ static constexpr int pad(int i) { return i + (i >> LPAD); }
  // log2 radix of forward pass p
  __host__ __device__ static constexpr int lr_fwd(int p) { return fft_sched_lr(LM, LR, BAL, p); }
  __host__ __device__ static constexpr int lns_fwd(int p) {  // log2 Ns before pass p
    int s = 0;
    for (int q = 0; q < p; ++q) s += lr_fwd(q);
    return s;
  }
  // twiddle table: pass p owns a block of (R - 1) * Ns entries at offset tw_off_fwd(p).
  __host__ __device__ static constexpr int tw_off_fwd(int p) {
    int o = 0;
    for (int q = 0; q < p; ++q) o += ((1 << lr_fwd(q)) - 1) << lns_fwd(q);
    return o;
  }
  __host__ __device__ static constexpr int tw_len() { return tw_off_fwd(NP); }
};

// The E - NB twiddles thread t needs in a pass (none when Ns = 1).  Kept apart from the butterfly
// so that a caller can issue these loads BEFORE the exchange barrier of the previous pass: their
// L1/L2 latency then overlaps the shared-memory exchange instead of following it.
template <class P, int LR, int LNS>
__device__ __forceinline__ void fft_pass_twiddles(float2 (&w)[P::E], int t,
                                                  const float2* __restrict__ tw) {
  constexpr int R = 1 << LR, NB = P::E / R, NS = 1 << LNS;
  if (LNS > 0) {
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const int k = (t + P::G * q) & (NS - 1);
#pragma unroll
      for (int r = 1; r < R; ++r) w[q + r * NB] = __ldg(&tw[(r - 1) * NS + k]);
    }
  }
}

// One radix pass on the E register-resident points of thread t.
//   v[e] holds x[t + G*e] on entry; on exit v[q + r*NB] holds output r of butterfly j = t + G*q.
template <class P, int LR, int LNS, int DIR>
__device__ __forceinline__ void fft_pass_butterflies(float2 (&v)[P::E], const float2 (&w)[P::E]) {
  constexpr int R = 1 << LR, NB = P::E / R;
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    float2 a[R];
#pragma unroll
    for (int r = 0; r < R; ++r) a[r] = v[q + r * NB];
    if (LNS > 0) {
#pragma unroll
      for (int r = 1; r < R; ++r) a[r] = twmul<DIR>(a[r], w[q + r * NB]);
    }
    Dft<R, DIR>::run(a);
#pragma unroll
    for (int r = 0; r < R; ++r) v[q + r * NB] = a[r];
  }
}

// Twiddles of one butterfly applied from FEW table entries: only w^(2^b k) (b = 0..LR-1) are loaded;
// input r is multiplied by the entry of each set bit of r except that the low three bits use
// products formed once (w^3, w^5, w^6, w^7).  A radix-16 butterfly then issues 4 loads instead of
// 15 (the line kernels are bound by the L1 / shared-memory data pipe, and twiddle loads were a
// quarter of its wavefronts) at the price of 11 more complex multiplies; rounding: at most three
// float32 multiplies on a twiddle path instead of one table entry rounded once.
#ifndef CFD_TW_POW
#define CFD_TW_POW 1
#endif
template <int R, int DIR>
__device__ __forceinline__ void apply_twiddles_pow(float2 (&a)[R], const float2* __restrict__ tw,
                                                   int NS, int k) {
  if constexpr (R == 2) {
    a[1] = twmul<DIR>(a[1], __ldg(&tw[k]));
  } else {
    const float2 w1 = __ldg(&tw[k]), w2 = __ldg(&tw[NS + k]);
    if constexpr (R == 4) {
      a[1] = twmul<DIR>(a[1], w1);
      a[2] = twmul<DIR>(a[2], w2);
      a[3] = twmul<DIR>(a[3], cmul(w1, w2));
    } else {
      const float2 w4 = __ldg(&tw[3 * NS + k]);
      float2 w[8];
      w[1] = w1;
      w[2] = w2;
      w[3] = cmul(w1, w2);
      w[4] = w4;
      w[5] = cmul(w4, w1);
      w[6] = cmul(w4, w2);
      w[7] = cmul(w4, w[3]);
#pragma unroll
      for (int hi = 0; hi < R; hi += 8) {
#pragma unroll
        for (int lo = 1; lo < 8; ++lo) a[hi + lo] = twmul<DIR>(a[hi + lo], w[lo]);
      }
      if constexpr (R >= 16) {
        const float2 w8 = __ldg(&tw[7 * NS + k]);
#pragma unroll
        for (int r = 8; r < R; ++r)
          if (r & 8) a[r] = twmul<DIR>(a[r], w8);
      }
      if constexpr (R >= 32) {
        const float2 w16 = __ldg(&tw[15 * NS + k]);
#pragma unroll
        for (int r = 16; r < R; ++r) a[r] = twmul<DIR>(a[r], w16);
      }
    }
  }
}

template <class P, int LR, int LNS, int DIR>
__device__ __forceinline__ void fft_pass_compute(float2 (&v)[P::E], int t,
                                                 const float2* __restrict__ tw) {
  constexpr int R = 1 << LR, NB = P::E / R, NS = 1 << LNS;
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    float2 a[R];
#pragma unroll
    for (int r = 0; r < R; ++r) a[r] = v[q + r * NB];
    if (LNS > 0) {
      const int k = (t + P::G * q) & (NS - 1);
#if CFD_TW_POW
      apply_twiddles_pow<R, DIR>(a, tw, NS, k);
#else
#pragma unroll
      for (int r = 1; r < R; ++r) a[r] = twmul<DIR>(a[r], __ldg(&tw[(r - 1) * NS + k]));
#endif
    }
    Dft<R, DIR>::run(a);
#pragma unroll
    for (int r = 0; r < R; ++r) v[q + r * NB] = a[r];
  }
}

// Scatter the outputs of a pass to their Stockham positions in the (padded) line buffer.
template <class P, int LR, int LNS>
__device__ __forceinline__ void fft_pass_store(const float2 (&v)[P::E], int t, float2* s) {
  constexpr int R = 1 << LR, NB = P::E / R, NS = 1 << LNS;
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int j = t + P::G * q;
    const int k = j & (NS - 1);
    const int base = ((j - k) << LR) + k;
#pragma unroll
    for (int r = 0; r < R; ++r) s[P::pad(base + r * NS)] = v[q + r * NB];
  }
}

template <class P>
__device__ __forceinline__ void fft_load_regs(float2 (&v)[P::E], int t, const float2* s) {
#pragma unroll
  for (int e = 0; e < P::E; ++e) v[e] = s[P::pad(t + P::G * e)];
}

// Full transforms (all threads of the CTA call these in lock-step, so __syncthreads is valid).
// Data is taken from registers v (x[t + G e]); the result is left in registers with natural index
// t + G*slot: the last pass has Ns = M / R, so output r of butterfly j = t + G*q sits at
// j + r*Ns = t + G*(q + r*NB).  Exchanges go through the padded line buffer `s`.
// Synchronisation of the G threads that own one line.  SyncCta: the whole CTA (always valid when
// every thread runs the same sequence).  SyncLine<G>: only the line's threads -- a warp-level sync
// when a line fits in a warp, otherwise a named barrier (id = 1 + line index, at most 15 lines) --
// so that the lines of a CTA drift apart and overlap each other's memory and compute phases.
struct SyncCta {
  static __device__ __forceinline__ void sync(int) { __syncthreads(); }
};
template <int G>
struct SyncLine {
  static __device__ __forceinline__ void sync(int line) {
    if (G <= 32) {
      __syncwarp();
    } else {
      asm volatile("bar.sync %0, %1;" ::"r"(line + 1), "n"(G) : "memory");
    }
  }
};

// DB = true: two exchange buffers, `s` and `s + alt`, used alternately (exchange number X0 + pass
// picks one).  A pass then writes a buffer nobody can still be reading -- its last readers passed
// the previous pass's barrier -- so the "all reads done" barrier disappears: one barrier per pass,
// and warps that finish a butterfly early store at once instead of idling.
//
// PRETW = true: the twiddles of pass p+1 are loaded before the exchange that follows pass p (E more
// live registers across the exchange -- for kernels that are not capped at 64 registers).
template <class P, int DIR, class SYNC = SyncCta, bool DB = false, int X0 = 0, bool PRETW = false>
struct FftRun {
  template <int PASS>
  static __device__ __forceinline__ void exchange(float2 (&v)[P::E], int t, float2* s, int line,
                                                  int alt) {
    constexpr int LR = P::lr_fwd(PASS);
    constexpr int LNS = P::lns_fwd(PASS);
    float2* buf = s;
    if constexpr (DB) {
      if constexpr ((X0 + PASS) & 1) buf = s + alt;
    } else {
      SYNC::sync(line);  // all reads of the previous layout are done
    }
    fft_pass_store<P, LR, LNS>(v, t, buf);
    SYNC::sync(line);
    fft_load_regs<P>(v, t, buf);
  }
  template <int PASS>
  static __device__ __forceinline__ void passes(float2 (&v)[P::E], int t, float2* s,
                                                const float2* __restrict__ tw, int line, int alt) {
    if constexpr (PASS < P::NP) {
      fft_pass_compute<P, P::lr_fwd(PASS), P::lns_fwd(PASS), DIR>(v, t, tw + P::tw_off_fwd(PASS));
      if constexpr (PASS + 1 < P::NP) {
        exchange<PASS>(v, t, s, line, alt);
        passes<PASS + 1>(v, t, s, tw, line, alt);
      }
    }
  }
  // w holds the twiddles of pass PASS on entry
  template <int PASS>
  static __device__ __forceinline__ void passes_pre(float2 (&v)[P::E], float2 (&w)[P::E], int t,
                                                    float2* s, const float2* __restrict__ tw,
                                                    int line, int alt) {
    if constexpr (PASS < P::NP) {
      fft_pass_butterflies<P, P::lr_fwd(PASS), P::lns_fwd(PASS), DIR>(v, w);
      if constexpr (PASS + 1 < P::NP) {
        fft_pass_twiddles<P, P::lr_fwd(PASS + 1), P::lns_fwd(PASS + 1)>(w, t,
                                                                        tw + P::tw_off_fwd(PASS + 1));
        exchange<PASS>(v, t, s, line, alt);
        passes_pre<PASS + 1>(v, w, t, s, tw, line, alt);
      }
    }
  }
  static __device__ __forceinline__ void run(float2 (&v)[P::E], int t, float2* s,
                                             const float2* __restrict__ tw, int line = 0,
                                             int alt = 0) {
    if constexpr (PRETW) {
      float2 w[P::E];  // pass 0 has Ns = 1: no twiddles
      passes_pre<0>(v, w, t, s, tw, line, alt);
    } else {
      passes<0>(v, t, s, tw, line, alt);
    }
  }
};

}  // namespace cfd
