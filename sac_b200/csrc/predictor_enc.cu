// predictor_enc.cu -- encode-direction predictor as two kernels of specialised warps (sm_100a).
//
// In the encoder the per-sample recurrence of one chain (OLS -> NLMS cascade + RLS -> mix -> bias, the reference's
// Predictor for one coded channel, /root/reference src/libsac/pred.cpp:4-45) falls apart into feed-forward
// recurrences, because every stage is driven by *input* samples (teacher forcing):
//
//   OLS  (ols.cpp)                  p_lpc(t)            depends on inputs only
//   cascade (cascade.h, ls.h, rls)  p_lms(t)            depends on inputs and p_lpc
//   bias (bias.h) + residual        e(t)                depends on inputs and p_lpc + p_lms
//
// ols_kernel      one CTA of 4 warps per chain (light: several CTAs per SM). A block of k samples: regressors (all
//                 lanes) -> one prediction per warp -> IRLS weights (one power per warp) -> covariance, k rank-1
//                 updates per element in one pass -> right-looking LDL^T with the next pivot and its reciprocal
//                 computed ahead of the trailing update -> back substitution. p_lpc(t) goes to HBM (8 B per sample).
// cascade_kernel  one CTA of 8 warps per chain, two warpgroups with their own register budgets (setmaxnreg):
//   warps 0-3   tap warps  128 lane-strided NLMS taps x 4 stages. Weights, power table and mu table of the first
//                          kR0/kR1/kR2/kR3 x 128 taps of each stage live in REGISTERS, the history ring in shared
//                          memory; longer stages continue from shared memory / HBM scratch.
//   warp 4      S  cascade scalars: stage predictions, 2-expert mix, stage targets, NLMS gradients, mix update
//   warp 5      B  bias stage, rounding, residual, cost sums (trails S through a ring in shared memory)
//   warp 6      R  RLS stage (5th cascade stage)
//   taps <-> S and S <-> R meet at named barriers (bar.arrive / bar.sync) twice per sample.
//
// The arithmetic is the canonical order of DESIGN.md section 2 (identical to predictor.cu, which remains the
// decode-direction kernel); the -m gpu tests compare residuals bit for bit with the CPU restatement of that order.
// Reference behaviour restated: src/libsac/pred.cpp:4-45, src/pred/{ols.cpp,ls.h,cascade.h,blend.h,rls.cpp,rls.h,
// bias.h}, src/common/math.h:14-78, src/libsac/libsac.cpp:94-142.
#include "chain.h"
#include "sac_canon_math.h"
#include <cuda_runtime.h>
#include <math_constants.h>
#include <cstddef>
#include <cstdio>

namespace sacb {

using sac_canon::c_exp;
using sac_canon::c_pow;
using sac_canon::c_round;

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kEncThreads = 256;                        // cascade kernel: 2 warpgroups: taps | S, B, R
constexpr int kTeam = 128;
constexpr int kMaxDynSmem = 227 * 1024;                              // OLS kernel: one team of 4 warps
constexpr int kR0 = 12, kR1 = 4, kR2 = 2, kR3 = 1;   // register-resident tap slots per thread and stage
constexpr int kQ = 32;                                // ring depth (samples)
constexpr int kXW = 256;                              // input window (samples, power of two)
constexpr int kXS = 100;                              // row stride of the regressor block
constexpr int kKB = 4;                                // samples per OLS sub-block
enum { kBarB1 = 1, kBarB2 = 2, kBarB3 = 3, kBarB4 = 4, kBarT = 5 };

__device__ __forceinline__ void bar_sync(int id, int cnt) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(cnt) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int cnt) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(cnt) : "memory"); }

__device__ __forceinline__ double shfl_xor(double v, int m) { return __shfl_xor_sync(kFull, v, m); }
__device__ __forceinline__ double shfl_idx(double v, int l) { return __shfl_sync(kFull, v, l); }
__device__ __forceinline__ double butterfly(double v)
{
  v = v + shfl_xor(v, 16);
  v = v + shfl_xor(v, 8);
  v = v + shfl_xor(v, 4);
  v = v + shfl_xor(v, 2);
  v = v + shfl_xor(v, 1);
  return v;
}
__device__ __forceinline__ double sgn(double x) { return (double)((x > 0) - (x < 0)); }
__device__ __forceinline__ double dmax(double a, double b) { return a < b ? b : a; }  // std::max(a,b)
__device__ __forceinline__ double dmin(double a, double b) { return b < a ? b : a; }  // std::min(a,b)

// slmath::dot for n < 8 and the AVX2 layout for 8 <= n <= 11 (math.h:130-161)
__device__ double small_dot(const double *x, const double *y, int n)
{
  double total = 0.0;
  int i = 0;
  if (n >= 8) {
    double s[4];
#pragma unroll
    for (int l = 0; l < 4; l++) s[l] = __fma_rn(x[l], y[l], 0.0) + __fma_rn(x[4 + l], y[4 + l], 0.0);
    total = s[0] + s[1] + s[2] + s[3];
    i = 8;
  }
  double init = 0.0;
  while (n - i >= 4) {
    const double v1 = x[i] * y[i] + x[i + 1] * y[i + 1];
    const double v2 = x[i + 2] * y[i + 2] + x[i + 3] * y[i + 3];
    init = init + (v1 + v2);
    i += 4;
  }
  for (; i < n; i++) init = init + x[i] * y[i];
  return total + init;
}
__device__ __forceinline__ double dot5(const double (&x)[kMixN], const double *y)
{
  const double v1 = x[0] * y[0] + x[1] * y[1];
  const double v2 = x[2] * y[2] + x[3] * y[3];
  double init = 0.0 + (v1 + v2);
  init = init + x[4] * y[4];
  return 0.0 + init;
}

#ifdef SACB_PROFILE_CASC
#define CP(i) do { const long long c__ = clock64(); cp[i] += c__ - clast; clast = c__; } while (0)
#define CP_DECL long long cp[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long clast = clock64()
#else
#define CP(i) do { } while (0)
#define CP_DECL do { } while (0)
#endif

struct EncShared {
  // taps <-> S
  double part[kTapWarps][8];          // [warp][stage] dot, [warp][4+stage] power sum
  double wgrad[kStages];
  double p[kMixN], spow[kStages];
  // S: mix state (cascade.h:11-57, ls.h:214-241, blend.h)
  double v[2][kMixN], eg[2][kMixN], sw[2], rsum[2];
  // S <-> R
  double bp4, p_rls;
  // R: RLS state (rls.cpp)
  double rx[kMaxRls], rw[kMaxRls], rph[kMaxRls], rP[kMaxRls][kMaxRls];
  double S0, S1;
  // B: bias state (bias.h)
  double hist_in[8], hist_d[8], mixw[4][3];
  double cnt[3][64], cval[3][64];
  double bmean, bvar;
  // S -> B ring
  double q2[kQ];
  int q2_pub, q2_con;
  // layout
  double *h[kStages], *ow[kStages], *opw[kStages], *omu[kStages];
  double *fpw[kStages], *fmu[kStages];   // full tables (HBM scratch), used at start-up only
  double sum_pow[kStages];
};

struct OlsShared {
  double X[kKB][kXS];                    // regressor block
  double xo[kXW], xq[kXW];               // input windows as doubles
  double pu[kKB], ff[kKB], wv[kMaxOls];
  double *cov, *W;
  int ld;
  int w_shared;                          // work matrix lies in shared memory (fast LDL path)
};

__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ int ld_vol(const int *p) { return *reinterpret_cast<const volatile int *>(p); }
__device__ __forceinline__ void st_vol(int *p, int v) { *reinterpret_cast<volatile int *>(p) = v; }
__device__ __forceinline__ double ldd_vol(const double *p) { return *reinterpret_cast<const volatile double *>(p); }
__device__ __forceinline__ void spin_until_gt(const int *ctr, int v)
{
  while (ld_vol(ctr) <= v) __nanosleep(40);
  __threadfence_block();                                     // acquire: ring data is read after the counter
}

// ------------------------------------------------------------------------------------------------------------------
// tap warps
template <int R> struct TapRegs { double w[R], pw[R], mu[R]; };

template <int R>
__device__ __forceinline__ void tap_load(TapRegs<R> &r, const double *fpw, const double *fmu, int N, int tl)
{
#pragma unroll
  for (int j = 0; j < R; j++) {
    const int i = tl + kTapThreads * j;
    r.w[j] = 0.0;
    r.pw[j] = i < N ? fpw[i] : 0.0;
    r.mu[j] = i < N ? fmu[i] : 0.0;
  }
}

// The history of a stage is kept twice, back to back (2 x (N+1) entries, h[k] == h[k + N + 1]): the window of the
// last N targets is then always contiguous at h + pos, tap i at h[pos + i], and a thread reaches its taps with
// compile-time offsets from one pointer (the same idea as the reference's RollBuffer2, common/histbuf.h:61-88).
//
// phase A: chain tl accumulates taps tl, tl+128, ... in ascending order (canonical order)
template <int R>
__device__ __forceinline__ void tap_dot(const TapRegs<R> &r, const double *h, const double *ow, const double *opw, int N, int pos,
                                        int tl, double &ad, double &ap)
{
  const double *hp = h + pos + tl;
  ad = 0.0; ap = 0.0;
#pragma unroll
  for (int j = 0; j < R; j++) {
    if (tl + kTapThreads * j < N) {
      const double hv = hp[kTapThreads * j];
      ad = __fma_rn(hv, r.w[j], ad);
      ap = __fma_rn(r.pw[j], hv * hv, ap);
    }
  }
  // taps beyond the register slots (shared memory, or HBM scratch for the largest stages): four per step, all loads
  // issued before the first fma so that their latencies overlap; the fma order per chain stays ascending in i
  int i = tl + kTapThreads * R;
  for (; i + 3 * kTapThreads < N; i += 4 * kTapThreads) {
    const int o = i - kTapThreads * R;
    const double *hq = h + pos + i;
    const double v0 = hq[0], v1 = hq[kTapThreads], v2 = hq[2 * kTapThreads], v3 = hq[3 * kTapThreads];
    const double w0 = ow[o], w1 = ow[o + kTapThreads], w2 = ow[o + 2 * kTapThreads], w3 = ow[o + 3 * kTapThreads];
    const double q0 = opw[o], q1 = opw[o + kTapThreads], q2 = opw[o + 2 * kTapThreads], q3 = opw[o + 3 * kTapThreads];
    ad = __fma_rn(v0, w0, ad); ap = __fma_rn(q0, v0 * v0, ap);
    ad = __fma_rn(v1, w1, ad); ap = __fma_rn(q1, v1 * v1, ap);
    ad = __fma_rn(v2, w2, ad); ap = __fma_rn(q2, v2 * v2, ap);
    ad = __fma_rn(v3, w3, ad); ap = __fma_rn(q3, v3 * v3, ap);
  }
  for (; i < N; i += kTapThreads) {
    const double hv = h[pos + i];
    const int o = i - kTapThreads * R;
    ad = __fma_rn(hv, ow[o], ad);
    ap = __fma_rn(opw[o], hv * hv, ap);
  }
}

// std::min(std::max(w, -10), 10) (ls.h:52) with one comparison: a NaN passes through exactly as in the reference
__device__ __forceinline__ double clamp10(double w)
{
  const double lim = w > 0.0 ? 10.0 : -10.0;
  return fabs(w) > 10.0 ? lim : w;
}

// phase C: w_i = clamp(fma(mutab_i, wgrad*h_i, w_i), +-10) on the pre-push window (ls.h:49-54)
template <int R>
__device__ __forceinline__ void tap_update(TapRegs<R> &r, const double *h, double *ow, const double *omu, int N, int pos, int tl,
                                           double g)
{
  const double *hp = h + pos + tl;
#pragma unroll
  for (int j = 0; j < R; j++) {
    if (tl + kTapThreads * j < N) {
      const double tt = g * hp[kTapThreads * j];
      r.w[j] = clamp10(__fma_rn(r.mu[j], tt, r.w[j]));
    }
  }
  int i = tl + kTapThreads * R;
  for (; i + 3 * kTapThreads < N; i += 4 * kTapThreads) {
    const int o = i - kTapThreads * R;
    const double *hq = h + pos + i;
    const double v0 = hq[0], v1 = hq[kTapThreads], v2 = hq[2 * kTapThreads], v3 = hq[3 * kTapThreads];
    double w0 = ow[o], w1 = ow[o + kTapThreads], w2 = ow[o + 2 * kTapThreads], w3 = ow[o + 3 * kTapThreads];
    const double m0 = omu[o], m1 = omu[o + kTapThreads], m2 = omu[o + 2 * kTapThreads], m3 = omu[o + 3 * kTapThreads];
    w0 = clamp10(__fma_rn(m0, g * v0, w0));
    w1 = clamp10(__fma_rn(m1, g * v1, w1));
    w2 = clamp10(__fma_rn(m2, g * v2, w2));
    w3 = clamp10(__fma_rn(m3, g * v3, w3));
    ow[o] = w0; ow[o + kTapThreads] = w1; ow[o + 2 * kTapThreads] = w2; ow[o + 3 * kTapThreads] = w3;
  }
  for (; i < N; i += kTapThreads) {
    const int o = i - kTapThreads * R;
    const double tt = g * h[pos + i];
    ow[o] = clamp10(__fma_rn(omu[o], tt, ow[o]));
  }
}

__device__ __forceinline__ void tap_warps(EncShared &S, const ChainDesc &d, int tid)
{
  const int tl = tid, lane = tid & 31, tw = tid >> 5;
  const int n = d.n;
  const int N0 = d.vn[0], N1 = d.vn[1], N2 = d.vn[2], N3 = d.vn[3];
  const double *h0 = S.h[0], *h1 = S.h[1], *h2 = S.h[2], *h3 = S.h[3];
  double *ow0 = S.ow[0], *ow1 = S.ow[1], *ow2 = S.ow[2], *ow3 = S.ow[3];
  const double *opw0 = S.opw[0], *opw1 = S.opw[1], *opw2 = S.opw[2], *opw3 = S.opw[3];
  const double *omu0 = S.omu[0], *omu1 = S.omu[1], *omu2 = S.omu[2], *omu3 = S.omu[3];
  TapRegs<kR0> r0; TapRegs<kR1> r1; TapRegs<kR2> r2; TapRegs<kR3> r3;
  tap_load(r0, S.fpw[0], S.fmu[0], N0, tl);
  tap_load(r1, S.fpw[1], S.fmu[1], N1, tl);
  tap_load(r2, S.fpw[2], S.fmu[2], N2, tl);
  tap_load(r3, S.fpw[3], S.fmu[3], N3, tl);
  int p0 = 0, p1 = 0, p2 = 0, p3 = 0;
  CP_DECL;
  for (int t = 0; t < n; t++) {
    CP(7);
    double acc[8];
#ifdef SACB_ABLATE_TAPS
    for (int q = 0; q < 8; q++) acc[q] = 0.0;
    if (false)
#endif
    {
    tap_dot(r0, h0, ow0, opw0, N0, p0, tl, acc[0], acc[4]);
    tap_dot(r1, h1, ow1, opw1, N1, p1, tl, acc[1], acc[5]);
    tap_dot(r2, h2, ow2, opw2, N2, p2, tl, acc[2], acc[6]);
    tap_dot(r3, h3, ow3, opw3, N3, p3, tl, acc[3], acc[7]);
    }
    CP(0);
    // eight butterflies in one: at each of the first three levels a lane keeps half of its values and trades the
    // other half with its partner, so value q ends up summed in lanes 4q..4q+3 -- per value the very same
    // own + partner additions at offsets 16, 8, 4, 2, 1 as butterfly(), with 9 instead of 40 shuffle rounds
    {
      const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
      double a4[4], a2[2], a1;
#pragma unroll
      for (int k = 0; k < 4; k++) a4[k] = (b4 ? acc[4 + k] : acc[k]) + shfl_xor(b4 ? acc[k] : acc[4 + k], 16);
#pragma unroll
      for (int k = 0; k < 2; k++) a2[k] = (b3 ? a4[2 + k] : a4[k]) + shfl_xor(b3 ? a4[k] : a4[2 + k], 8);
      a1 = (b2 ? a2[1] : a2[0]) + shfl_xor(b2 ? a2[0] : a2[1], 4);
      a1 = a1 + shfl_xor(a1, 2);
      a1 = a1 + shfl_xor(a1, 1);
      if ((lane & 3) == 0) S.part[tw][lane >> 2] = a1;
    }
    __threadfence_block();
    bar_arrive(kBarB1, 160);
    CP(1);
    bar_sync(kBarB2, 160);
    CP(2);
    const double g0 = ldd_vol(&S.wgrad[0]), g1 = ldd_vol(&S.wgrad[1]), g2 = ldd_vol(&S.wgrad[2]), g3 = ldd_vol(&S.wgrad[3]);
#ifndef SACB_ABLATE_TAPS
    tap_update(r0, h0, ow0, omu0, N0, p0, tl, g0);
    tap_update(r1, h1, ow1, omu1, N1, p1, tl, g1);
    tap_update(r2, h2, ow2, omu2, N2, p2, tl, g2);
    tap_update(r3, h3, ow3, omu3, N3, p3, tl, g3);
#endif
    p0 = p0 == 0 ? N0 : p0 - 1;                              // S pushed bp[s] at this slot
    p1 = p1 == 0 ? N1 : p1 - 1;
    p2 = p2 == 0 ? N2 : p2 - 1;
    p3 = p3 == 0 ? N3 : p3 - 1;
    CP(3);
  }
#ifdef SACB_PROFILE_CASC
  if (tid == 0 && blockIdx.x < 2)
    printf("chain %d taps: clk/sample dot %.0f reduce+arrive %.0f wait_S %.0f update %.0f\n", blockIdx.x, (double)cp[0] / n, (double)cp[1] / n, (double)cp[2] / n, (double)cp[3] / n);
#endif
}

// ------------------------------------------------------------------------------------------------------------------
// S: cascade scalars (cascade.h:91-118, ls.h:45-56 gradient, ls.h:223-237, blend.h:50-90)
__device__ __forceinline__ void scalar_warp(EncShared &S, const ChainDesc &d, int lane)
{
  const int n = d.n;
  const double alpha = d.proj_alpha;
  const int myN = lane < kStages ? d.vn[lane] : 0;
  double *myh = lane < kStages ? S.h[lane] : nullptr;
  const double my_mu = lane < kStages ? d.vmu[lane] : 0.0;
  const double my_sp = lane < kStages ? S.sum_pow[lane] : 0.0;
  int pos = 0;
  bool bad = false;
  const double *plpc = d.plpc;
  CP_DECL;
  int32_t vcur = lane < n ? __ldg(d.own + lane) : 0;
  int32_t vnext = 32 + lane < n ? __ldg(d.own + 32 + lane) : 0;
  double lcur = lane < n ? __ldg(plpc + lane) : 0.0;
  double lnext = 32 + lane < n ? __ldg(plpc + 32 + lane) : 0.0;
  for (int t = 0; t < n; t++) {
    if ((t & 31) == 0 && t) {
      vcur = vnext; vnext = t + 32 + lane < n ? __ldg(d.own + t + 32 + lane) : 0;
      lcur = lnext; lnext = t + 32 + lane < n ? __ldg(plpc + t + 32 + lane) : 0.0;
    }
    const double val = (double)__shfl_sync(kFull, vcur, t & 31);
    const double p_lpc = shfl_idx(lcur, t & 31);
    // state that does not depend on this sample's predictions
    const double sw0 = S.sw[0], sw1 = S.sw[1];
    double v0[kMixN], v1[kMixN], wi[kMixN];
#pragma unroll
    for (int i = 0; i < kMixN; i++) {
      v0[i] = S.v[0][i]; v1[i] = S.v[1][i];
      wi[i] = dmax(0.0 + ((0.0 + v0[i] * sw0) + v1[i] * sw1), 0.0);
    }
    const double target = val - p_lpc;
    CP(0);
    if (t) bar_sync(kBarB4, 64);                             // p_rls for this sample
    CP(1);
    bar_sync(kBarB1, 160);                                   // tap partial sums
    CP(2);
    if (lane < 8) {
      const double v = (ldd_vol(&S.part[0][lane]) + ldd_vol(&S.part[1][lane])) + (ldd_vol(&S.part[2][lane]) + ldd_vol(&S.part[3][lane]));
      if (lane < 4) S.p[lane] = v; else S.spow[lane - 4] = v;
    }
    if (lane == 8) S.p[4] = t ? ldd_vol(&S.p_rls) : 0.0;
    __syncwarp();
    double p[kMixN];
#pragma unroll
    for (int i = 0; i < kMixN; i++) p[i] = S.p[i];
    // ---- mix predict (cascade.h:36-44, blend.h:25-30) ----
    const double ep0 = dot5(p, v0), ep1 = dot5(p, v1);
    if (!(fabs(ep0) <= 1.7976931348623157e308) || !(fabs(ep1) <= 1.7976931348623157e308)) bad = true;
    const double p_lms = 0.0 + ((0.0 + ep0 * sw0) + ep1 * sw1);
    const double px = p_lpc + p_lms;
    // ---- stage targets (cascade.h:99-113) ----
    double bp[kMixN];
    {
      double prefix = 0.0;
#pragma unroll
      for (int i = 0; i < kMixN; i++) {
        const double pxi = (1.0 - alpha) * prefix + alpha * p_lms;
        bp[i] = target - dmin(dmax(pxi, d.casc_lo), d.casc_hi);
        prefix += wi[i] * p[i];
      }
    }
    // ---- NLMS gradients + history push (ls.h:45-56) ----
    if (lane < kStages) {
      double bpl = bp[0], pl = p[0];
      if (lane == 1) { bpl = bp[1]; pl = p[1]; } else if (lane == 2) { bpl = bp[2]; pl = p[2]; } else if (lane == 3) { bpl = bp[3]; pl = p[3]; }
      S.wgrad[lane] = my_mu * (bpl - pl) * my_sp / (S.spow[lane] + 1.0);
      const int np = pos == 0 ? myN : pos - 1;
      myh[np] = bpl;
      myh[np + myN + 1] = bpl;                               // mirrored copy (see tap_dot)
      pos = np;
    }
    if (lane == 4) S.bp4 = bp[4];
    __threadfence_block();
    bar_arrive(kBarB2, 160);
    bar_arrive(kBarB3, 64);
    CP(3);
    // ---- hand px to the bias warp ----
    if (t >= kQ) spin_until_gt(&S.q2_con, t - kQ);
    CP(5);
    if (lane == 0) { S.q2[t & (kQ - 1)] = px; __threadfence_block(); st_vol(&S.q2_pub, t + 1); }
    // ---- mix update: LS_ADA<L1>, LS_ADA<L2> (ls.h:223-237), BlendExp (blend.h:50-90) ----
#ifndef SACB_ABLATE_POST
    {
      if (lane < 2 * kMixN) {
        const int ex = lane / kMixN, i = lane - ex * kMixN;
        const double er = target - (ex == 0 ? ep0 : ep1);
        const double loss = ex == 0 ? sgn(er) : er;
        double pi_ = p[0];
        if (i == 1) pi_ = p[1]; else if (i == 2) pi_ = p[2]; else if (i == 3) pi_ = p[3]; else if (i == 4) pi_ = p[4];
        const double grad = loss * pi_;
        const double egn = d.mix_beta * S.eg[ex][i] + (1.0 - d.mix_beta) * grad * grad;
        S.eg[ex][i] = egn;
        const double mu_scaled = d.mu_mix / (sqrt(egn) + 1e-5);
        S.v[ex][i] += mu_scaled * grad;
      }
      const double r0 = 0.95 * S.rsum[0] + (1.0 - 0.95) * (-fabs(target - ep0));
      const double r1 = 0.95 * S.rsum[1] + (1.0 - 0.95) * (-fabs(target - ep1));
      const double z0 = 1.0 * r0, z1 = 1.0 * r1;
      const double mz = dmax(dmax(-CUDART_INF, z0), z1);
      const double w0 = c_exp(z0 - mz), w1 = c_exp(z1 - mz);
      double total = 0.0; total += w0; total += w1;
      const double inv_total = 1.0 / total;
      __syncwarp();
      if (lane == 0) { S.rsum[0] = r0; S.rsum[1] = r1; S.sw[0] = w0 * inv_total; S.sw[1] = w1 * inv_total; }
      __syncwarp();
    }
#endif
    CP(4);
  }
#ifdef SACB_PROFILE_CASC
  if (lane == 0 && blockIdx.x < 2)
    printf("chain %d S: clk/sample pre %.0f wait_R %.0f wait_taps %.0f crit %.0f wait_B(ring) %.0f post(mix update) %.0f\n", blockIdx.x, (double)cp[0] / n, (double)cp[1] / n, (double)cp[2] / n, (double)cp[3] / n, (double)cp[5] / n, (double)cp[4] / n);
#endif
  if (lane == 0 && d.flags) *d.flags = bad ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------------------------
// R: RLS stage, UpdateHist(bp[4]) then the prediction for t+1 (rls.cpp:17-65, rls.h:22-36)
__device__ __forceinline__ void rls_warp(EncShared &S, const ChainDesc &d, int lane)
{
  const int n = d.n, lm_n = d.lm_n;
  double p_rls = 0.0;
  CP_DECL;
  for (int t = 0; t < n; t++) {
    CP(0);
    bar_sync(kBarB3, 64);
    CP(1);
#ifdef SACB_ABLATE_R
    if (t + 1 < n) { __threadfence_block(); bar_arrive(kBarB4, 64); }
    continue;
#endif
    const double bp4 = ldd_vol(&S.bp4);
    const double err = bp4 - p_rls;
    if (lane < lm_n) S.rph[lane] = small_dot(S.rP[lane], S.rx, lm_n);
    __syncwarp();
    const double phi = dmax(small_dot(S.rx, S.rph, lm_n), 1e-8);
    const double err2 = err * err;
    const double R = dmax(S.S0 - S.S1, 1e-5);
    const double nis = err2 / (phi + R);
    const double m = c_exp(-d.lm_gamma * nis);
    const double al = 0.99 + (0.999 - 0.99) * m;
    const double denom = 1. / (al + phi);
    const double inv_al = 1.0 / al;
    for (int q = lane; q < lm_n * lm_n; q += 32) {
      const int i = q / lm_n, j = q - i * lm_n;
      if (j <= i) {
        const double mm = S.rph[i] * S.rph[j];
        const double vv = (S.rP[i][j] - denom * mm) * inv_al;
        S.rP[i][j] = vv; S.rP[j][i] = vv;
      }
    }
    double xprev = 0.0;
    if (lane < lm_n) {
      S.rw[lane] += err * (denom * S.rph[lane]);
      xprev = lane > 0 ? S.rx[lane - 1] : bp4;
    }
    __syncwarp();
    if (lane < lm_n) S.rx[lane] = xprev;                     // RollBack (utils.h:330-336)
    if (lane == 0) {
      S.S0 = 0.95 * S.S0 + (1.0 - 0.95) * err2;
      S.S1 = 0.95 * S.S1 + (1.0 - 0.95) * phi;
    }
    __syncwarp();
    p_rls = small_dot(S.rx, S.rw, lm_n);
    if (t + 1 < n) {
      if (lane == 0) S.p_rls = p_rls;
      __threadfence_block();
      bar_arrive(kBarB4, 64);
    }
    CP(2);
  }
#ifdef SACB_PROFILE_CASC
  if (lane == 0 && blockIdx.x < 2)
    printf("chain %d R: clk/sample wait_S %.0f busy %.0f\n", blockIdx.x, (double)cp[1] / n, (double)cp[2] / n);
#endif
}

// ------------------------------------------------------------------------------------------------------------------
// B: bias stage (bias.h:64-162), round / clamp / residual (libsac.cpp:105-108), cost sums
__device__ __forceinline__ void bias_warp(EncShared &S, const ChainDesc &d, int lane)
{
  const int n = d.n;
  long long l1 = 0, sq = 0;
  int32_t vcur = lane < n ? __ldg(d.own + lane) : 0;
  int32_t vnext = 32 + lane < n ? __ldg(d.own + 32 + lane) : 0;
  int32_t ebuf = 0;
  CP_DECL;
  for (int t = 0; t < n; t++) {
    if ((t & 31) == 0 && t) { vcur = vnext; vnext = t + 32 + lane < n ? __ldg(d.own + t + 32 + lane) : 0; }
    const int32_t vali = __shfl_sync(kFull, vcur, t & 31);
    const double val = (double)vali;
    CP(0);
    spin_until_gt(&S.q2_pub, t);
    CP(1);
    const double px = ldd_vol(&S.q2[t & (kQ - 1)]);
    __syncwarp();
    if (lane == 0) st_vol(&S.q2_con, t + 1);
#ifdef SACB_ABLATE_B
    if ((t & 31) == lane) ebuf = (int32_t)px;
    if ((t & 31) == 31 || t == n - 1) { const int base = t & ~31; if (base + lane <= t) d.resid[base + lane] = ebuf; }
    continue;
#endif
    // ---- bias predict (bias.h:64-126) ----
    int ctx0, ctx1, ctx2, mix_ctx;
    {
      const double h0 = S.hist_in[0], h1 = S.hist_in[1], h2 = S.hist_in[2];
      const double d0 = S.hist_d[0], d1 = S.hist_d[1], d2 = S.hist_d[2], d3 = S.hist_d[3], d4 = S.hist_d[4];
      const int b0 = h0 > px ? 0 : 1;
      const int b2 = d0 < 0 ? 0 : 1, b3 = d1 < 0 ? 0 : 1, b4 = d2 < 0 ? 0 : 1;
      const int b5 = d1 < d0 ? 0 : 1, b6 = d2 < d1 ? 0 : 1, b7 = d3 < d2 ? 0 : 1, b8 = d4 < d3 ? 0 : 1;
      const int b9 = fabs(d0) > 32 ? 0 : 1;
      const int b10 = 2 * h0 - h1 > px ? 0 : 1;
      const int b11 = 3 * h0 - 3 * h1 + h2 > px ? 0 : 1;
      double sum = 0;
      sum += fabs(d0); sum += fabs(d1); sum += fabs(d2); sum += fabs(d3); sum += fabs(d4);
      sum /= 5.0;
      mix_ctx = sum > 512 ? 2 : (sum > 32 ? 1 : 0);
      ctx0 = b0 + (b2 << 1) + (b9 << 2) + (b10 << 3) + (b11 << 4);
      ctx1 = b2 + (b3 << 1) + (b4 << 2);
      ctx2 = b5 + (b6 << 1) + (b7 << 2) + (b8 << 3);
    }
    const int myctx = lane == 0 ? ctx0 : (lane == 1 ? ctx1 : ctx2);
    double ptl = 0.0;
    if (lane < 3) ptl = S.cval[lane][myctx] / S.cnt[lane][myctx];
    const double pt0 = shfl_idx(ptl, 0), pt1 = shfl_idx(ptl, 1), pt2 = shfl_idx(ptl, 2);
    const double pbias = 0.0 + (((0.0 + pt0 * S.mixw[mix_ctx][0]) + pt1 * S.mixw[mix_ctx][1]) + pt2 * S.mixw[mix_ctx][2]);
    const double pd = px + pbias;
    int32_t pi;
    {
      const double r = c_round(pd);
      if (!(r >= (double)d.clamp_lo)) pi = d.clamp_lo;
      else if (r > (double)d.clamp_hi) pi = d.clamp_hi;
      else pi = (int32_t)r;
    }
    const int32_t e = vali - pi;
    if ((t & 31) == lane) ebuf = e;
    if ((t & 31) == 31 || t == n - 1) {
      const int base = t & ~31;
      if (base + lane <= t) d.resid[base + lane] = ebuf;
    }
    { const long long ae = e < 0 ? -(long long)e : (long long)e; l1 += ae; sq += (long long)e * (long long)e; }
    // ---- bias update (bias.h:127-162) ----
    {
      const double delta = val - c_round(px);
      const double bv = dmax(0.0, S.bvar), bm = S.bmean;
      const double diff = delta - bm;
      const double z = diff * diff / (bv + 1E-5);
      const double wgt = c_exp(-0.5 * z);
      double hin = 0.0, hdl = 0.0;
      if (lane < 8) { hin = lane > 0 ? S.hist_in[lane - 1] : val; hdl = lane > 0 ? S.hist_d[lane - 1] : delta; }
      __syncwarp();
      if (lane < 8) { S.hist_in[lane] = hin; S.hist_d[lane] = hdl; }
      if (lane < 3) {
        double cv = S.cval[lane][myctx] + wgt * delta;
        double cc = S.cnt[lane][myctx] + wgt;
        if (cc >= (double)d.bias_nscale) { cv *= 0.5; cc *= 0.5; }
        S.cval[lane][myctx] = cv; S.cnt[lane][myctx] = cc;
        const double ptv = lane == 0 ? pt0 : (lane == 1 ? pt1 : pt2);
        S.mixw[mix_ctx][lane] += (d.bias_mu * sgn(delta - pbias)) * sgn(ptv);
      }
      if (lane == 0) {
        const double nm = 0.998 * bm + (1.0 - 0.998) * delta;
        S.bvar = 0.998 * S.bvar + (1.0 - 0.998) * ((delta - bm) * (delta - nm));
        S.bmean = nm;
      }
      __syncwarp();
    }
    CP(2);
  }
#ifdef SACB_PROFILE_CASC
  if (lane == 0 && blockIdx.x < 2)
    printf("chain %d B: clk/sample top %.0f wait_S %.0f busy %.0f\n", blockIdx.x, (double)cp[0] / n, (double)cp[1] / n, (double)cp[2] / n);
#endif
  if (lane == 0) {
    if (d.l1sum) *d.l1sum = l1;
    if (d.sqsum) *d.sqsum = sq;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// O: OLS (ols.cpp:7-57, math.h:14-78) on a team of four warps (128 lanes, own named barrier). Matrices are
// (n+1) x ld row-major squares: the lower triangle holds the covariance (row n = b) resp. the work matrix, the upper
// triangle of the work matrix receives L^T. A block of k samples: regressors (all lanes) -> one prediction per warp
// -> publish -> IRLS weights (one power per warp) -> covariance, k rank-1 updates per element in one pass, 8 x 16
// lane grid -> LDL^T with the next pivot and its reciprocal computed redundantly by every lane ahead of the
// trailing update (one team barrier per column) -> back substitution on warp 0.
__device__ __forceinline__ void team_sync() { bar_sync(kBarT, kTeam); }

__device__ __forceinline__ void ols_team(OlsShared &S, const ChainDesc &d, int tl)
{
  const int lane = tl & 31, tw = tl >> 5;
  const int N = d.n;
  const int n = d.lenA + d.lenB, n1 = n + 1;
  const int ld = S.ld;
  double *cov = S.cov, *W = S.W;
  const double lambda = d.lambda, nu = d.nu, one_m_lambda = 1.0 - d.lambda;
  const int ra = tl >> 4, ca = tl & 15;
  double w0 = 0.0, w1 = 0.0, w2 = 0.0;                        // w[lane], w[lane+32], w[lane+64] (every warp its copy)
  double esum = 0.0;
  int km = 0;
  // input windows as doubles: indices [-64, 128) to start with; outside [0, N) reads as 0 (pred.cpp:17-31)
  for (int q = tl; q < 192; q += kTeam) {
    const int idx = q - 64;
    const bool in = idx >= 0 && idx < N;
    S.xo[idx & (kXW - 1)] = in ? (double)__ldg(d.own + idx) : 0.0;
    S.xq[idx & (kXW - 1)] = in ? (double)__ldg(d.other + idx) : 0.0;
  }
  int fill_end = 128;
  int t = 0;
  while (t < N) {
    if (fill_end < t + 68) {
      if (tl < 64) {
        const int idx = fill_end + tl;
        const bool in = idx < N;
        S.xo[idx & (kXW - 1)] = in ? (double)__ldg(d.own + idx) : 0.0;
        S.xq[idx & (kXW - 1)] = in ? (double)__ldg(d.other + idx) : 0.0;
      }
      fill_end += 64;
    }
    const int kb = min(min(kKB, d.k - km), N - t);
    team_sync();
    // ---- regressors of the sub-block: X[u][0..n), X[u][n] = sample ----
    for (int idx = tl; idx < kb * n1; idx += kTeam) {
      const int u = idx / n1, j = idx - u * n1;
      const int tt = t + u;
      const int sB = max(tt - d.lagB, d.minB) - d.backB;
      double v;
      if (j < d.lenA) v = S.xo[(tt - d.lenA + j) & (kXW - 1)];
      else if (j < n) v = S.xq[(sB + j - d.lenA) & (kXW - 1)];
      else v = S.xo[tt & (kXW - 1)];
      S.X[u][j] = v;
    }
    team_sync();
    // ---- predictions: warp u takes sample u; 32 lane-strided fma chains + butterfly (ols.cpp:22-25) ----
    if (tw < kb) {
      double acc = 0.0;
      if (lane < n) acc = __fma_rn(S.X[tw][lane], w0, acc);
      if (lane + 32 < n) acc = __fma_rn(S.X[tw][lane + 32], w1, acc);
      if (lane + 64 < n) acc = __fma_rn(S.X[tw][lane + 64], w2, acc);
      acc = butterfly(acc);
      if (lane == 0) S.pu[tw] = acc;
    }
    team_sync();
    double pu[kKB];
#pragma unroll
    for (int u = 0; u < kKB; u++) pu[u] = u < kb ? S.pu[u] : 0.0;
    // ---- predictions to HBM ----
    if (tl < kb) d.plpc[t + tl] = S.pu[tl];
    // ---- IRLS weights (ols.cpp:29-36): the running sum is serial and cheap, the powers run one per warp ----
    double es_mine = 1.0;
#pragma unroll
    for (int u = 0; u < kKB; u++) {
      if (u < kb) {
        const double eo = S.X[u][n] - pu[u];
        esum = d.beta_sum * esum + fabs(eo);
        if (tw == u) es_mine = esum;
      }
    }
    if (tw < kb) {
      const double ff_mine = one_m_lambda * c_pow(es_mine + d.beta_add, -d.beta_pow);
      if (lane == 0) S.ff[tw] = ff_mine;
    }
    team_sync();
    double ff[kKB];
#pragma unroll
    for (int u = 0; u < kKB; u++) ff[u] = u < kb ? S.ff[u] : 0.0;
    // ---- covariance: kb rank-1 updates per element in one pass (ols.cpp:38-45), 8 x 16 lane grid ----
    km += kb;
    const bool solve = km >= d.k;
    for (int i = ra; i <= n; i += 8) {
      double xi[kKB];
#pragma unroll
      for (int u = 0; u < kKB; u++) xi[u] = u < kb ? S.X[u][i] : 0.0;
      for (int c = ca; c <= i; c += 32) {
        double v[2], xc[2][kKB];
        bool on[2];
#pragma unroll
        for (int b = 0; b < 2; b++) {
          const int cc = c + 16 * b;
          on[b] = cc <= i && !(i == n && cc == n);
          v[b] = on[b] ? cov[i * ld + cc] : 0.0;
#pragma unroll
          for (int u = 0; u < kKB; u++) xc[b][u] = (on[b] && u < kb) ? S.X[u][cc] : 0.0;
        }
#pragma unroll
        for (int b = 0; b < 2; b++) {
#pragma unroll
          for (int u = 0; u < kKB; u++)
            if (u < kb) v[b] = lambda * v[b] + ff[u] * (xi[u] * xc[b][u]);
        }
#pragma unroll
        for (int b = 0; b < 2; b++) {
          const int cc = c + 16 * b;
          if (on[b]) {
            cov[i * ld + cc] = v[b];
            if (solve) W[i * ld + cc] = (i == cc) ? v[b] + nu : v[b];
          }
        }
      }
    }
    if (solve) {
      km = 0;
      team_sync();
      // ---- right-looking LDL^T of (C + nu I), augmented with b as row n. Every lane carries the pivot chain
      //      (d_j, 1/d_j) redundantly in registers; the owner of element (j+1,j+1) does not store its column-j update
      //      (it is only ever needed as the next pivot), which leaves a single barrier per column ----
      bool ok = true;
      double inv;
      {
        const double d0 = W[0];
        if (d0 < 1e-12) ok = false;
        inv = 1.0 / d0;
      }
      if (S.w_shared) {
        // Fast path, work matrix in shared memory, TWO columns per team barrier. Relative to the pivot (j,j) the element
        // offsets of a lane never change, so a step is a fixed 4 x 4 register tile per lane and 32 x 64 block: all loads
        // are issued first (32-bit shared addresses), then the fmas, then the stores. Per element the arithmetic is the
        // canonical right-looking sequence (column j, then column j+1 with the column-j-updated values):
        //   l_i0 = W_ij * inv_j                 u_i1 = fma(-l_i0, W_{j+1,j}, W_{i,j+1})      l_i1 = u_i1 * inv_{j+1}
        //   W_ic = fma(-l_i1, u_c1, fma(-l_i0, W_cj, W_ic))
        // Every lane carries the pivot chain redundantly: inv_{j+1} at the start of the step, inv_{j+2} (from the
        // rank-2-updated element (j+2,j+2), which its owner does not store) ahead of the tile work.
        const uint32_t ldb = (uint32_t)ld * 8u;
        uint32_t pjj = (uint32_t)__cvta_generic_to_shared(W);
        const uint32_t r0 = (uint32_t)(2 + ra) * ldb, rstep = 8u * ldb;
        const uint32_t q0 = (uint32_t)(2 + ca) * 8u;
        int j = 0;
        while (ok && n - j >= 2) {
          const int m = n - j;                                     // trailing rows rel 1..m (row m = b), columns rel 1..m-1
          const double a10 = lds_f64(pjj + ldb);
          const double l10 = a10 * inv;
          const double d1 = __fma_rn(-l10, a10, lds_f64(pjj + ldb + 8u));
          if (d1 < 1e-12) { ok = false; break; }
          const double inv1 = 1.0 / d1;
          double inv_next = 0.0;
          bool ok_next = true;
          if (m > 2) {                                             // pivot of column j+2 after both updates
            const double a0 = lds_f64(pjj + 2u * ldb), a1 = lds_f64(pjj + 2u * ldb + 8u);
            double e = lds_f64(pjj + 2u * ldb + 16u);
            const double l0 = a0 * inv;
            const double u1 = __fma_rn(-l0, a10, a1);
            const double l1 = u1 * inv1;
            e = __fma_rn(-l0, a0, e);
            e = __fma_rn(-l1, u1, e);
            if (e < 1e-12) ok_next = false;
            inv_next = 1.0 / e;
          }
          for (int rb = 0; rb + 2 <= m; rb += 32) {
            for (int cb = 0; cb + 2 <= m - 1 && cb <= rb + 31; cb += 64) {
              const uint32_t rbase = pjj + (uint32_t)rb * ldb + r0;
              const uint32_t cbase = (uint32_t)cb * 8u + q0;
              const int rr0 = rb + 2 + ra, qq0 = cb + 2 + ca;
              const int delta = qq0 - rr0;                         // element (a,b) is on/below the diagonal iff 8a - 16b >= delta
              bool rv[4], cv[4];
              double l0[4], l1[4], c0[4], c1[4], e[4][4];
#pragma unroll
              for (int a = 0; a < 4; a++) {
                rv[a] = rr0 + 8 * a <= m;
                l0[a] = rv[a] ? lds_f64(rbase + (uint32_t)a * rstep) : 0.0;
                l1[a] = rv[a] ? lds_f64(rbase + (uint32_t)a * rstep + 8u) : 0.0;
              }
#pragma unroll
              for (int b = 0; b < 4; b++) {
                cv[b] = qq0 + 16 * b <= m - 1;
                const uint32_t ca_ = pjj + (uint32_t)(qq0 + 16 * b) * ldb;
                c0[b] = cv[b] ? lds_f64(ca_) : 0.0;
                c1[b] = cv[b] ? lds_f64(ca_ + 8u) : 0.0;
              }
              bool on[4][4];
#pragma unroll
              for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                  on[a][b] = rv[a] && cv[b] && (8 * a - 16 * b >= delta) && !(a == 0 && b == 0 && rr0 == 2 && qq0 == 2);
                  e[a][b] = on[a][b] ? lds_f64(rbase + (uint32_t)a * rstep + cbase + (uint32_t)b * 128u) : 0.0;
                }
#pragma unroll
              for (int a = 0; a < 4; a++) {
                const double la0 = l0[a] * inv;
                l1[a] = __fma_rn(-la0, a10, l1[a]) * inv1;         // l_i1
                l0[a] = la0;
              }
#pragma unroll
              for (int b = 0; b < 4; b++) c1[b] = __fma_rn(-(c0[b] * inv), a10, c1[b]);   // u_c1
#pragma unroll
              for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) e[a][b] = __fma_rn(-l1[a], c1[b], __fma_rn(-l0[a], c0[b], e[a][b]));
#pragma unroll
              for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++)
                  if (on[a][b]) sts_f64(rbase + (uint32_t)a * rstep + cbase + (uint32_t)b * 128u, e[a][b]);
            }
          }
          // L^T into the upper triangle: rows j and j+1 (the lower-triangle columns keep the unscaled values)
          for (int r = 1 + tl; r <= m; r += kTeam) {
            const double a0 = lds_f64(pjj + (uint32_t)r * ldb);
            const double la0 = a0 * inv;
            sts_f64(pjj + (uint32_t)r * 8u, la0);
            if (r >= 2) {
              const double u1 = __fma_rn(-la0, a10, lds_f64(pjj + (uint32_t)r * ldb + 8u));
              sts_f64(pjj + ldb + (uint32_t)r * 8u, u1 * inv1);
            }
          }
          team_sync();
          pjj += 2u * (ldb + 8u);
          j += 2;
          inv = inv_next; ok = ok_next;
        }
        if (ok && n - j == 1) {                                    // last column of an odd order: only the b row is left
          if (tl == 0) sts_f64(pjj + 8u, lds_f64(pjj + ldb) * inv);
          team_sync();
        }
      } else {
        // the host always sizes the launch so that the work matrix fits shared memory (n <= 96: 75 KB)
        __trap();
      }
      if (ok && tw == 0) {
        // back substitution L^T w = z (z = row n of L), columns in descending order, one fma per element
        double y0 = lane < n ? W[lane * ld + n] : 0.0;
        double y1 = lane + 32 < n ? W[(lane + 32) * ld + n] : 0.0;
        double y2 = lane + 64 < n ? W[(lane + 64) * ld + n] : 0.0;
        const int lr = min(lane, n - 1);
        double ln0 = n >= 2 ? W[lr * ld + (n - 1)] : 0.0;     // L[k][lane] for the next step, fetched one step ahead
        for (int k = n - 1; k >= 1; k--) {
          const double lk0 = ln0;
          if (k >= 2) ln0 = W[lr * ld + (k - 1)];
          const int sl = k >> 5;
          double yk = sl == 0 ? y0 : (sl == 1 ? y1 : y2);
          yk = shfl_idx(yk, k & 31);
          if (lane < k) y0 = __fma_rn(-lk0, yk, y0);
          if (k > 32 && lane + 32 < k) y1 = __fma_rn(-W[(lane + 32) * ld + k], yk, y1);
          if (k > 64 && lane + 64 < k) y2 = __fma_rn(-W[(lane + 64) * ld + k], yk, y2);
        }
        if (lane < n) S.wv[lane] = y0;
        if (lane + 32 < n) S.wv[lane + 32] = y1;
        if (lane + 64 < n) S.wv[lane + 64] = y2;
      }
      team_sync();
      if (ok) {
        w0 = lane < n ? S.wv[lane] : 0.0;
        w1 = lane + 32 < n ? S.wv[lane + 32] : 0.0;
        w2 = lane + 64 < n ? S.wv[lane + 64] : 0.0;
      }
    }
    t += kb;
  }
}

// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTeam, 5) ols_kernel(const ChainDesc *__restrict__ descs, const int *__restrict__ idx)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ChainDesc &d = descs[idx ? idx[blockIdx.x] : blockIdx.x];
  OlsShared &S = *reinterpret_cast<OlsShared *>(smem_raw);
  const int tid = threadIdx.x;
  const int n_ols = d.lenA + d.lenB;
  const int ld = (n_ols + 1) | 1;
  if (tid == 0) {
    unsigned int dyn_bytes;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_bytes));
    const size_t head = (sizeof(OlsShared) + 15) & ~size_t(15);
    double *sp = reinterpret_cast<double *>(smem_raw + head);
    long long s_left = ((long long)dyn_bytes - (long long)head) / 8;
    double *gp = d.scratch_ols;
    const long long cnt = (long long)(n_ols + 1) * ld;
    S.ld = ld;
    S.w_shared = cnt <= s_left;
    if (cnt <= s_left) { S.W = sp; sp += cnt; s_left -= cnt; } else { S.W = gp; gp += cnt; }
    if (cnt <= s_left) { S.cov = sp; sp += cnt; s_left -= cnt; } else { S.cov = gp; gp += cnt; }
  }
  __syncthreads();
  for (int i = tid; i < (n_ols + 1) * ld; i += kTeam) { S.cov[i] = 0.0; S.W[i] = 0.0; }
  for (int i = tid; i < kMaxOls; i += kTeam) S.wv[i] = 0.0;
  __syncthreads();
  ols_team(S, d, tid);
}

__global__ void __launch_bounds__(kEncThreads, 2) cascade_kernel(const ChainDesc *__restrict__ descs, const int *__restrict__ idx)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ChainDesc &d = descs[idx ? idx[blockIdx.x] : blockIdx.x];
  EncShared &S = *reinterpret_cast<EncShared *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- carve: shared memory first (histories, overflow taps), the rest in the chain's HBM scratch ----
  if (tid == 0) {
    unsigned int dyn_bytes;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_bytes));
    const size_t head = (sizeof(EncShared) + 15) & ~size_t(15);
    double *sp = reinterpret_cast<double *>(smem_raw + head);
    long long s_left = ((long long)dyn_bytes - (long long)head) / 8;
    double *gp = d.scratch;
    auto take = [&](long long cnt) -> double * {
      double *r;
      if (cnt <= s_left) { r = sp; sp += cnt; s_left -= cnt; }
      else { r = gp; gp += cnt; }
      return r;
    };
    auto take_global = [&](long long cnt) -> double * { double *r = gp; gp += cnt; return r; };
    const int regs[kStages] = {kR0, kR1, kR2, kR3};
    for (int s = 0; s < kStages; s++) { S.fpw[s] = take_global(d.vn[s]); S.fmu[s] = take_global(d.vn[s]); }
    for (int s = kStages - 1; s >= 0; s--) S.h[s] = take(2 * (d.vn[s] + 1));
    for (int s = kStages - 1; s >= 0; s--) {
      const int ov = max(d.vn[s] - kTapThreads * regs[s], 0);
      S.ow[s] = take(ov); S.opw[s] = take(ov); S.omu[s] = take(ov);
    }
  }
  // zero everything up to the layout block
  {
    double *z = reinterpret_cast<double *>(&S);
    const int nz = (int)(offsetof(EncShared, q2_pub) / 8);
    for (int i = tid; i < nz; i += kEncThreads) z[i] = 0.0;
    if (tid == 0) { S.q2_pub = 0; S.q2_con = 0; }
  }
  __syncthreads();
  // ---- tables and initial state ----
  {
    const int regs[kStages] = {kR0, kR1, kR2, kR3};
    for (int s = 0; s < kStages; s++) {
      const int N = d.vn[s];
      const double md = d.vmudecay[s], pd = d.vpowdecay[s];
      double *h = S.h[s], *fpw = S.fpw[s], *fmu = S.fmu[s];
      double *ow = S.ow[s], *opw = S.opw[s], *omu = S.omu[s];
      const int r128 = kTapThreads * regs[s];
      for (int i = tid; i < N; i += kEncThreads) {
        const double pwv = 1.0 / c_pow((double)(1 + i), pd);  // ls.h:39
        const double muv = c_pow(md, (double)i);              // ls.h:41
        fpw[i] = pwv; fmu[i] = muv;
        h[i] = 0.0; h[i + N + 1] = 0.0;
        if (i >= r128) { ow[i - r128] = 0.0; opw[i - r128] = pwv; omu[i - r128] = muv; }
      }
      if (tid == 0) { h[N] = 0.0; h[2 * N + 1] = 0.0; }
    }
  }
  __syncthreads();
  if (tid < kStages) {                                       // sum_powtab accumulates sequentially (ls.h:40)
    const double *fpw = S.fpw[tid];
    const int N = d.vn[tid];
    double sp = 0.0;
    for (int i = 0; i < N; i++) sp += fpw[i];
    S.sum_pow[tid] = sp;
  }
  if (tid < 2 * kMixN) S.v[tid / kMixN][tid % kMixN] = 1.0 / kMixN;        // LSInitType::Uniform (ls.h:199-201)
  if (tid < 2) S.sw[tid] = 1.0 / 2;                                          // blend.h:22
  if (tid < d.lm_n) S.rP[tid][tid] = 1.0 / 1.0;                              // rls.cpp:14-15 (nu=1)
  if (tid < 64) { S.cnt[0][tid] = 4.0; S.cnt[1][tid] = 4.0; S.cnt[2][tid] = 4.0; }   // bias.h:24-28 (freq0=4)
  __syncthreads();

  // register pool of the CTA (256 x 128): taps 176, scalar warps 80
  if (warp < kTapWarps) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
    tap_warps(S, d, tid);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    if (warp == 4) scalar_warp(S, d, lane);
    else if (warp == 5) bias_warp(S, d, lane);
    else if (warp == 6) rls_warp(S, d, lane);
  }
}

} // namespace

size_t predictor_enc_shared_bytes() { return (sizeof(EncShared) + 15) & ~size_t(15); }
size_t predictor_ols_shared_bytes() { return (sizeof(OlsShared) + 15) & ~size_t(15); }

// doubles of HBM scratch a chain may need when nothing but the fixed blocks fit shared memory
long long predictor_enc_scratch_doubles(const int *vn, int n_ols)
{
  long long t = 0;
  for (int s = 0; s < kStages; s++) t += 7LL * vn[s] + 2;
  return t + 16;
}
// doubles of shared memory that keep every array of a chain on chip (histories + taps beyond the register slots)
long long predictor_enc_smem_doubles(const int *vn)
{
  const int regs[kStages] = {kR0, kR1, kR2, kR3};
  long long t = 0;
  for (int s = 0; s < kStages; s++) t += 2 * (vn[s] + 1) + 3LL * (vn[s] > kTapThreads * regs[s] ? vn[s] - kTapThreads * regs[s] : 0);
  return t;
}
long long predictor_ols_scratch_doubles(int n_ols)
{
  const long long ld = (n_ols + 1) | 1;
  return 2LL * (n_ols + 1) * ld + 16;
}

// residuals of every chain: OLS predictions first (ols_kernel -> ChainDesc::plpc), then the cascade
cudaError_t predictor_enc_init_attributes()
{
  cudaError_t e = cudaFuncSetAttribute(cascade_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(ols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
}
cudaError_t launch_predictor_enc(const ChainDesc *d_ols_descs, int nols, const ChainDesc *d_descs, int nchains, int smem_bytes,
                                 int ols_smem_bytes, cudaStream_t stream, cudaEvent_t between)
{
  ols_kernel<<<nols, kTeam, ols_smem_bytes, stream>>>(d_ols_descs, nullptr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (between && (e = cudaEventRecord(between, stream)) != cudaSuccess) return e;
  if (nchains <= 0) return cudaSuccess;                              // OLS stages alone
  cascade_kernel<<<nchains, kEncThreads, smem_bytes, stream>>>(d_descs, nullptr);
  return cudaGetLastError();
}
// the OLS stages / the cascade alone over a subset of the descriptors (chains the search-grade kernels do not take)
cudaError_t launch_ols_canonical(const ChainDesc *d_descs, const int *d_idx, int count, int smem_bytes, cudaStream_t stream)
{
  if (count <= 0) return cudaSuccess;
  ols_kernel<<<count, kTeam, smem_bytes, stream>>>(d_descs, d_idx);
  return cudaGetLastError();
}
cudaError_t launch_cascade_canonical(const ChainDesc *d_descs, const int *d_idx, int count, int smem_bytes, cudaStream_t stream)
{
  if (count <= 0) return cudaSuccess;
  cascade_kernel<<<count, kEncThreads, smem_bytes, stream>>>(d_descs, d_idx);
  return cudaGetLastError();
}

} // namespace sacb
