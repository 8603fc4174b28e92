// predictor_sg.cu -- SEARCH-GRADE encode-direction predictor kernels (sm_100a).
//
// The DDS search only ranks candidates: its <= 1000 objective evaluations per frame need the reference's
// formulas, not the decoder's bit pattern. These kernels evaluate exactly the recurrences of predictor_enc.cu (the reference's
// Predictor, /root/reference src/libsac/pred.cpp:4-45, src/pred/{ols.cpp,ls.h,cascade.h,blend.h,rls.cpp,bias.h}) but are free
// in summation order, fused multiply-adds, reciprocals instead of divisions -- and, above all, in SCHEDULE:
//
//  cascade_sg_kernel  look-ahead NLMS. With h(t) the stage history and w(t) its weights, the prediction of the NEXT sample is
//        p(t+1) = bp(t) w0(t+1) + sum_{j>=1} h_{j-1}(t) w_j(t) + g(t) * sum_{j>=1} mu_j h_{j-1}(t) h_j(t)
//        (w_j(t+1) = w_j(t) + mu_j g(t) h_j(t), ls.h:45-56, as long as no weight hits the +-10 clamp), and the normaliser
//        spow(t+1) = pw_0 bp(t)^2 + sum_{j>=1} pw_j h_{j-1}(t)^2. All three sums depend on history and weights that are known
//        BEFORE the scalar stage of sample t has produced g(t) and bp(t): the tap warps compute them (and apply the weight
//        update of sample t-1) WHILE the scalar warps work on sample t. The serial loop dot -> reduce -> scalars -> update of
//        predictor_enc.cu (8 450 clk/sample, 70 % of it synchronisation) becomes two concurrent streams that meet at ONE
//        CTA barrier per sample. A clamp event (rare; counted) makes the look-ahead inexact: the chain is flagged and the
//        engine re-evaluates it with the canonical kernel.
//        8 tap warps (taps j >= 1 in per-thread contiguous blocks of odd length, weights in registers, tables and history in
//        shared memory) + S (stage predictions, mix, targets, gradients; owns tap 0) + M (mix update) + R (RLS stage, its
//        matrix-vector part computed ahead of the sample) + B (bias stage, residual, trails by one sample).
//  ols_warp_kernel<NP>  orders n <= NP <= 32: ONE WARP per chain, four chains per CTA, lane i owns row i of the covariance and
//        of the work matrix in REGISTERS; right-looking LDL^T exchanges one column per step through a double-buffered
//        shared-memory vector (one __syncwarp per column); L goes to shared memory by rows for the back substitution.
//        Orders above 32 stay on the canonical team kernel (predictor_enc.cu): a 256-thread block-cyclic register kernel and a
//        two-warp rows-in-registers kernel were both measured slower than it on the B200 (profiles/README.md, round 2); the
//        latter is kept behind SACB_OLS_ROWS=1 (ols_rows_kernel) as the record of that experiment.
//
// Results equal the canonical kernels' up to rounding: residuals differ in isolated samples where a prediction lies within
// ~1e-12 of a rounding boundary. tests/test_gpu_grade.py states and checks the tolerance (costs, ranking).
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
constexpr int kSgMaxTapWarps = 8;
enum { kBarPhase = 3, kBarS2MR = 1, kBarMR2S = 2 };                 // barrier 0 stays with __syncthreads()

__device__ __forceinline__ void bar_sync(int id, int cnt) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(cnt) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int cnt) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(cnt) : "memory"); }
__device__ __forceinline__ double shfl_xor(double v, int m) { return __shfl_xor_sync(kFull, v, m); }
__device__ __forceinline__ double shfl_idx(double v, int l) { return __shfl_sync(kFull, v, l); }
__device__ __forceinline__ double sgn(double x) { return (double)((x > 0) - (x < 0)); }
__device__ __forceinline__ double ldd_vol(const double *p) { return *reinterpret_cast<const volatile double *>(p); }

// 1/x for normal x > 0: hardware seed (MUFU.RCP64H, about 8 good bits -- two Newton steps were measured to leave ~1e-9, which
// compounds in the RLS matrix update) + three Newton steps, no special-case handling
__device__ __forceinline__ double rcp_fast(double x)
{
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}

// =====================================================================================================================
// cascade
// =====================================================================================================================
struct SgShared {
  double sums[2][kSgMaxTapWarps][12];    // [t&1][tap warp][3*stage + {B', A', P'}]: look-ahead sums made during sample t
  double g[2][kStages];                  // g_i(t) at [t&1]
  double pxq[2];                         // p_lpc + p_lms of sample t at [t&1] (for B)
  // S -> M, R (same sample)
  double mx_p[kMixN], mx_ep[2], mx_target, bp4;
  // M -> S: expert weights, blend weights, and the combined stage weights max(sum_e s_e v_e[i], 0) (cascade.h:24-34)
  double v[2][kMixN], sw[2], wi[kMixN];
  // R -> S
  double p_rls;
  // B state (bias.h)
  double hist_in[8], hist_d[8], mixw[4][3], cnt[3][64], cval[3][64], bmean, bvar;
  // R: x (history of the stage input) and ph = P x, shared by the lanes of the warp
  double rls_xch[2 * kMaxRls + 4];
  // layout
  double *h[kStages], *mu[kStages], *pw[kStages], *wt[kStages];
  int L[kStages], M[kStages];            // taps per tap thread (odd), ring length
  double sum_pow[kStages];
  int clamped, bad;
};

// TW tap warps + 4 scalar warps; R0..R3 register-resident weights per tap thread and stage
template <int TW, int R0, int R1, int R2, int R3> struct SgCfg {
  static constexpr int tw = TW, tap_threads = 32 * TW, threads = 32 * TW + 128, phase = 32 * TW + 64;
  static constexpr int r0 = R0, r1 = R1, r2 = R2, r3 = R3;
};
using SgSmall = SgCfg<4, 15, 7, 3, 1>;     // 256 threads, 2 CTAs per SM: stages up to 1921 / 897 / 385 / 129 taps (default profile: 1280/256/32/4)
using SgLarge = SgCfg<8, 21, 9, 5, 3>;     // 384 threads, 1 CTA per SM: anything whose tables fit shared memory (longer blocks spill weights to it)

// shared-memory accesses by 32-bit address + immediate offset: one base register per array instead of one 64-bit generic
// address per slot (ptxas otherwise hoists ~100 loop-invariant addresses out of the sample loop: 224 registers)
template <int OFF> __device__ __forceinline__ double lds_o(uint32_t a)
{
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF));
  return v;
}
__device__ __forceinline__ double lds_a(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_a(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// e^x for x <= 0 (all uses: softmax of losses, forgetting factor, bias gate): k = rint(x log2 e), degree-13 Taylor polynomial in
// Estrin form (dependent depth 5 instead of Horner's 13), 2^k by exponent arithmetic; below 2^-1000 the result is 0
__device__ __forceinline__ double exp_neg(double x)
{
  const double kd = rint(x * 0x1.71547652b82fep+0);
  double r = fma(-kd, 0x1.62e42fee00000p-1, x);
  r = fma(-kd, 0x1.a39ef35793c76p-33, r);
  const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
  const double p01 = fma(r, 1.0, 1.0), p23 = fma(r, 0x1.5555555555555p-3, 0.5), p45 = fma(r, 0x1.1111111111111p-7, 0x1.5555555555555p-5);
  const double p67 = fma(r, 0x1.a01a01a01a01ap-13, 0x1.6c16c16c16c17p-10), p89 = fma(r, 0x1.71de3a556c734p-19, 0x1.a01a01a01a01ap-16);
  const double pab = fma(r, 0x1.ae64567f544e4p-26, 0x1.27e4fb7789f5cp-22), pcd = fma(r, 0x1.6124613a86d09p-33, 0x1.1eed8eff8d898p-29);
  const double q0 = fma(r2, p23, p01), q1 = fma(r2, p67, p45), q2 = fma(r2, pab, p89);
  const double o0 = fma(r4, q1, q0), o1 = fma(r4, pcd, q2);
  const double p = fma(r8, o1, o0);
  const int k = (int)kd;
  if (k < -1000) return 0.0;
  return __longlong_as_double(__double_as_longlong(p) + ((long long)k << 52));
}

// one tap: weight update of the previous sample (the clamp is only DETECTED: a chain that meets it is re-evaluated by the
// canonical kernel), then the three look-ahead sums. hp = h_{j+1}(t) = h_j(t-1), hc = h_j(t), hm = h_{j-1}(t).
__device__ __forceinline__ void tap_one(double &w, double hp, double hc, double hm, double m, double pwj, double gprev, double &aB, double &aA,
                                        double &aP, bool &clamped)
{
  const double wn = fma(m * gprev, hp, w);
  clamped |= fabs(wn) > 10.0;
  w = wn;
  aB = fma(hm, wn, aB);
  aA = fma(m * hc, hm, aA);
  aP = fma(pwj, hm * hm, aP);
}

// L taps of one thread, exact length (tables and ring are padded with zeros up to a whole block: no predicates)
template <int R, int L, int Q> struct TapBlock {
  static __device__ __forceinline__ void run(double (&w)[R], uint32_t ah, uint32_t amu, uint32_t apw, double gprev, double hm, double hc, double &aB,
                                             double &aA, double &aP, bool &clamped)
  {
    if constexpr (Q < L && Q < R) {
      const double hp = lds_o<8 * (Q + 1)>(ah);
      tap_one(w[Q], hp, hc, hm, lds_o<8 * Q>(amu), lds_o<8 * Q>(apw), gprev, aB, aA, aP, clamped);
      TapBlock<R, L, Q + 1>::run(w, ah, amu, apw, gprev, hc, hp, aB, aA, aP, clamped);
    }
  }
};
template <int R, int L>
__device__ __forceinline__ void tap_block(double (&w)[R], uint32_t ah, uint32_t amu, uint32_t apw, double gprev, double &aB, double &aA, double &aP,
                                          bool &clamped)
{
  TapBlock<R, L, 0>::run(w, ah, amu, apw, gprev, lds_o<-8>(ah), lds_o<0>(ah), aB, aA, aP, clamped);
}

// one stage of one tap thread for one sample. hw = shared address of h_0(t); tables at amu0 / apw0 / awt0 (index 0); this
// thread's block starts at tap j0 and has L taps (odd, uniform over the CTA), the first R of them in registers.
template <int R>
__device__ __forceinline__ void tap_stage(double (&w)[R], uint32_t hring, int pos, int M, uint32_t amu0, uint32_t apw0, uint32_t awt0, int j0, int L,
                                          double gprev, double &aB, double &aA, double &aP, bool &clamped)
{
  // the ring keeps its first L + 2 slots a second time behind its end: the L + 2 entries h_{j0-1} .. h_{j0+L} of this block are
  // contiguous from slot (pos + j0 - 1) mod M whatever the ring position
  int start = pos + j0 - 1;
  start -= start >= M ? M : 0;
  const uint32_t o0 = 8u * (uint32_t)j0;
  const uint32_t ah = hring + 8u * (uint32_t)(start + 1), amu = amu0 + o0, apw = apw0 + o0;
  switch (L < R ? L : R) {                                           // uniform branch; odd lengths only (sg_block_len)
    case 1: tap_block<R, 1>(w, ah, amu, apw, gprev, aB, aA, aP, clamped); break;
    case 3: if constexpr (R >= 3) tap_block<R, 3>(w, ah, amu, apw, gprev, aB, aA, aP, clamped); break;
    case 5: if constexpr (R >= 5) tap_block<R, 5>(w, ah, amu, apw, gprev, aB, aA, aP, clamped); break;
    case 7: if constexpr (R >= 7) tap_block<R, 7>(w, ah, amu, apw, gprev, aB, aA, aP, clamped); break;
    case 9: if constexpr (R >= 9) tap_block<R, 9>(w, ah, amu, apw, gprev, aB, aA, aP, clamped); break;
    case 11: if constexpr (R >= 11) tap_block<R, 11>(w, ah, amu, apw, gprev, aB, aA, aP, clamped); break;
    case 13: if constexpr (R >= 13) tap_block<R, 13>(w, ah, amu, apw, gprev, aB, aA, aP, clamped); break;
    case 15: if constexpr (R >= 15) tap_block<R, 15>(w, ah, amu, apw, gprev, aB, aA, aP, clamped); break;
    case 17: if constexpr (R >= 17) tap_block<R, 17>(w, ah, amu, apw, gprev, aB, aA, aP, clamped); break;
    case 19: if constexpr (R >= 19) tap_block<R, 19>(w, ah, amu, apw, gprev, aB, aA, aP, clamped); break;
    case 21: if constexpr (R >= 21) tap_block<R, 21>(w, ah, amu, apw, gprev, aB, aA, aP, clamped); break;
    default: break;
  }
  if (L > R) {                                                       // block longer than the register slots: weights in shared memory
    double hm = lds_a(ah + 8u * (uint32_t)(R - 1)), hc = lds_a(ah + 8u * (uint32_t)R);
    for (int q = R; q < L; q++) {
      const uint32_t o = 8u * (uint32_t)q;
      const double hp = lds_a(ah + o + 8u);
      const uint32_t aw = awt0 + o0 + o;
      double wv = lds_a(aw);
      tap_one(wv, hp, hc, hm, lds_a(amu + o), lds_a(apw + o), gprev, aB, aA, aP, clamped);
      sts_a(aw, wv);
      hm = hc; hc = hp;
    }
  }
}

// NV values per lane summed over the warp: at each of the first levels a lane keeps half of its values and trades the other
// half with its partner; value q ends up in the lanes whose upper bits spell q. Writes sums[base + q].
template <int NV>
__device__ __forceinline__ void warp_sums(const double (&acc)[12], int first, int lane, double *out)
{
  static_assert(NV == 8 || NV == 16, "padded value count");
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
  if constexpr (NV == 16) {
    double a8[8], a4[4], a2[2], a1;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const double lo = acc[k], hi = k + 8 < 12 ? acc[k + 8] : 0.0;
      a8[k] = (b4 ? hi : lo) + shfl_xor(b4 ? lo : hi, 16);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) a4[k] = (b3 ? a8[4 + k] : a8[k]) + shfl_xor(b3 ? a8[k] : a8[4 + k], 8);
#pragma unroll
    for (int k = 0; k < 2; k++) a2[k] = (b2 ? a4[2 + k] : a4[k]) + shfl_xor(b2 ? a4[k] : a4[2 + k], 4);
    a1 = (b1 ? a2[1] : a2[0]) + shfl_xor(b1 ? a2[0] : a2[1], 2);
    a1 = a1 + shfl_xor(a1, 1);
    const int slot = (b4 ? 8 : 0) + (b3 ? 4 : 0) + (b2 ? 2 : 0) + (b1 ? 1 : 0);
    if ((lane & 1) == 0 && slot < 12) out[slot] = a1;
  } else {                                                           // the first six values only (stages 0 and 1)
    double a4[4], a2[2], a1;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const double lo = acc[first + k], hi = k + 4 < 6 ? acc[first + k + 4] : 0.0;
      a4[k] = (b4 ? hi : lo) + shfl_xor(b4 ? lo : hi, 16);
    }
#pragma unroll
    for (int k = 0; k < 2; k++) a2[k] = (b3 ? a4[2 + k] : a4[k]) + shfl_xor(b3 ? a4[k] : a4[2 + k], 8);
    a1 = (b2 ? a2[1] : a2[0]) + shfl_xor(b2 ? a2[0] : a2[1], 4);
    a1 = a1 + shfl_xor(a1, 2);
    a1 = a1 + shfl_xor(a1, 1);
    const int slot = (b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0);
    if ((lane & 3) == 0 && slot < 6) out[first + slot] = a1;
  }
}

template <class CFG>
__device__ __forceinline__ void sg_tap_warps(SgShared &S, const ChainDesc &d, int tid)
{
  const int lane = tid & 31, tw = tid >> 5;
  const int n = d.n;
  int M[kStages], j0[kStages], L[kStages], pos[kStages];
  bool mine[kStages];                                                // this thread has a block of the stage
  uint32_t ah[kStages], amu[kStages], apw[kStages], awt[kStages];
#pragma unroll
  for (int s = 0; s < kStages; s++) {
    L[s] = S.L[s]; M[s] = S.M[s];
    j0[s] = 1 + tid * L[s];
    mine[s] = j0[s] < d.vn[s];
    pos[s] = 0;                                                      // ring position of h_0(t)
    ah[s] = smem_u32(S.h[s]); amu[s] = smem_u32(S.mu[s]); apw[s] = smem_u32(S.pw[s]); awt[s] = S.wt[s] ? smem_u32(S.wt[s]) : 0u;
  }
  // stages 2 and 3 are short (default 32 and 4 taps): most warps hold none of their taps and sum six values instead of twelve
  const bool wide = (1 + (tw * 32) * L[2] < d.vn[2]) || (1 + (tw * 32) * L[3] < d.vn[3]);
  double w0[CFG::r0], w1[CFG::r1], w2[CFG::r2], w3[CFG::r3];
#pragma unroll
  for (int q = 0; q < CFG::r0; q++) w0[q] = 0.0;
#pragma unroll
  for (int q = 0; q < CFG::r1; q++) w1[q] = 0.0;
#pragma unroll
  for (int q = 0; q < CFG::r2; q++) w2[q] = 0.0;
#pragma unroll
  for (int q = 0; q < CFG::r3; q++) w3[q] = 0.0;
  bool clamped = false;
  for (int t = 0; t <= n; t++) {
    if (t < n) {
      const double *gp = S.g[(t + 1) & 1];                           // g(t-1)
      double acc[12];
#pragma unroll
      for (int q = 0; q < 12; q++) acc[q] = 0.0;
      if (mine[0]) tap_stage<CFG::r0>(w0, ah[0], pos[0], M[0], amu[0], apw[0], awt[0], j0[0], L[0], ldd_vol(gp), acc[0], acc[1], acc[2], clamped);
      if (mine[1]) tap_stage<CFG::r1>(w1, ah[1], pos[1], M[1], amu[1], apw[1], awt[1], j0[1], L[1], ldd_vol(gp + 1), acc[3], acc[4], acc[5], clamped);
      if (wide) {
        if (mine[2]) tap_stage<CFG::r2>(w2, ah[2], pos[2], M[2], amu[2], apw[2], awt[2], j0[2], L[2], ldd_vol(gp + 2), acc[6], acc[7], acc[8], clamped);
        if (mine[3]) tap_stage<CFG::r3>(w3, ah[3], pos[3], M[3], amu[3], apw[3], awt[3], j0[3], L[3], ldd_vol(gp + 3), acc[9], acc[10], acc[11], clamped);
        warp_sums<16>(acc, 0, lane, &S.sums[t & 1][tw][0]);
      } else {
        warp_sums<8>(acc, 0, lane, &S.sums[t & 1][tw][0]);             // slots 6..11 of this warp stay zero
      }
#pragma unroll
      for (int s = 0; s < kStages; s++) pos[s] = pos[s] == 0 ? M[s] - 1 : pos[s] - 1;   // S pushes bp(t) at this slot during the same phase
    }
    bar_sync(kBarPhase, CFG::phase);
  }
  if (__any_sync(kFull, clamped) && lane == 0) atomicOr(&S.clamped, 1);
}

// S: stage predictions from the look-ahead sums, 2-expert mix, stage targets, gradients, history push (cascade.h:91-118)
template <class CFG>
__device__ __forceinline__ void sg_scalar_warp(SgShared &S, const ChainDesc &d, int lane)
{
  const int n = d.n;
  const double alpha = d.proj_alpha, one_m_alpha = 1.0 - d.proj_alpha;
  const int li = lane & 3;                                           // lanes >= 4 shadow lanes 0..3 (results unused)
  const bool own = lane < kStages;
  const int myM = S.M[li], myE = S.L[li] + 2;                        // ring length, slots repeated behind its end
  double *myh = S.h[li];
  const double my_mu = d.vmu[li], my_sp = S.sum_pow[li];
  const double clo = d.casc_lo, chi = d.casc_hi;
  int pos = 0;
  double w0 = 0.0, bp1 = 0.0, bp2 = 0.0, gprev = 0.0;                // tap 0 of stage li, bp(t-1), bp(t-2), g(t-1)
  bool bad = false;
  const double *plpc = d.plpc;
  int32_t vcur = lane < n ? __ldg(d.own + lane) : 0;
  int32_t vnext = 32 + lane < n ? __ldg(d.own + 32 + lane) : 0;
  double lcur = lane < n ? __ldg(plpc + lane) : 0.0;
  double lnext = 32 + lane < n ? __ldg(plpc + 32 + lane) : 0.0;
  for (int t = 0; t <= n; t++) {
    if (t < n) {
      if ((t & 31) == 0 && t) {
        vcur = vnext; vnext = t + 32 + lane < n ? __ldg(d.own + t + 32 + lane) : 0;
        lcur = lnext; lnext = t + 32 + lane < n ? __ldg(plpc + t + 32 + lane) : 0.0;
      }
      const double val = (double)__shfl_sync(kFull, vcur, t & 31);
      const double p_lpc = shfl_idx(lcur, t & 31);
      const double target = val - p_lpc;
      // ---- part A: needs only the tap warps' sums of the previous phase ----
      double Bs = 0.0, As = 0.0, Ps = 0.0;
      if (t) {
        const double *sp = &S.sums[(t + 1) & 1][0][3 * li];
        double b0 = sp[0], a0 = sp[1], q0 = sp[2];
        double b1 = sp[12], a1 = sp[13], q1 = sp[14];
#pragma unroll
        for (int w = 2; w < CFG::tw; w += 2) {
          b0 += sp[12 * w]; a0 += sp[12 * w + 1]; q0 += sp[12 * w + 2];
          b1 += sp[12 * w + 12]; a1 += sp[12 * w + 13]; q1 += sp[12 * w + 14];
        }
        Bs = b0 + b1; As = a0 + a1; Ps = q0 + q1;
      }
      {                                                              // tap 0: mu_0 = pw_0 = 1 (ls.h:38-42)
        double wn = fma(gprev, bp2, w0);
        wn = fabs(wn) > 10.0 ? (wn > 0.0 ? 10.0 : -10.0) : wn;
        w0 = wn;
      }
      const double p_mine = fma(bp1, w0, fma(gprev, As, Bs));
      const double rs = rcp_fast(fma(bp1, bp1, Ps) + 1.0);           // 1 / (spow + 1)
      double p[kMixN];
      p[0] = shfl_idx(p_mine, 0); p[1] = shfl_idx(p_mine, 1); p[2] = shfl_idx(p_mine, 2); p[3] = shfl_idx(p_mine, 3);
      // ---- part B: needs the mix weights and the RLS prediction made after sample t-1 ----
      if (t) bar_sync(kBarMR2S, 96);
      p[4] = t ? ldd_vol(&S.p_rls) : 0.0;
      const double sw0 = ldd_vol(&S.sw[0]), sw1 = ldd_vol(&S.sw[1]);
      double e0a = 0.0, e0b = 0.0, e1a = 0.0, e1b = 0.0, wi[kMixN];
#pragma unroll
      for (int i = 0; i < kMixN; i++) {
        const double v0 = ldd_vol(&S.v[0][i]), v1 = ldd_vol(&S.v[1][i]);
        if (i & 1) { e0b = fma(p[i], v0, e0b); e1b = fma(p[i], v1, e1b); } else { e0a = fma(p[i], v0, e0a); e1a = fma(p[i], v1, e1a); }
        wi[i] = ldd_vol(&S.wi[i]);
      }
      const double ep0 = e0a + e0b, ep1 = e1a + e1b;
      if (!(fabs(ep0) <= 1.7976931348623157e308) || !(fabs(ep1) <= 1.7976931348623157e308)) bad = true;
      const double p_lms = fma(ep1, sw1, ep0 * sw0);
      double bp[kMixN];
      {
        const double apl = alpha * p_lms;
        double prefix = 0.0;
#pragma unroll
        for (int i = 0; i < kMixN; i++) {
          bp[i] = target - fmin(fmax(fma(one_m_alpha, prefix, apl), clo), chi);
          prefix = fma(wi[i], p[i], prefix);
        }
      }
      // what M and R wait for goes out first
      if (lane == 4) S.bp4 = bp[4];
      if (lane >= 8 && lane < 8 + kMixN) { const int k = lane - 8; S.mx_p[k] = k == 0 ? p[0] : (k == 1 ? p[1] : (k == 2 ? p[2] : (k == 3 ? p[3] : p[4]))); }
      if (lane == 16) { S.mx_ep[0] = ep0; S.mx_ep[1] = ep1; S.mx_target = target; }
      __threadfence_block();
      bar_arrive(kBarS2MR, 96);
      // ... then what the tap warps and B need at the next phase
      const double bpl = li == 0 ? bp[0] : (li == 1 ? bp[1] : (li == 2 ? bp[2] : bp[3]));
      const double g = my_mu * (bpl - p_mine) * my_sp * rs;
      if (own) {
        S.g[t & 1][lane] = g;
        const int np = pos == 0 ? myM - 1 : pos - 1;
        myh[np] = bpl;
        if (np < myE) myh[np + myM] = bpl;                           // the first L + 2 slots are kept twice (see tap_stage)
        pos = np;
      }
      if (lane == 16) S.pxq[t & 1] = p_lpc + p_lms;
      bp2 = bp1; bp1 = bpl; gprev = g;
    }
    bar_sync(kBarPhase, CFG::phase);
  }
  if (bad && lane == 0) atomicOr(&S.bad, 1);
}

// M: LS_ADA<L1>, LS_ADA<L2> (ls.h:223-237) on lanes 0..9, BlendExp (blend.h:50-90) on all lanes
__device__ __forceinline__ void sg_mix_warp(SgShared &S, const ChainDesc &d, int lane)
{
  const int n = d.n;
  const int ex = lane < kMixN ? 0 : 1, i = min(lane - ex * kMixN, kMixN - 1);
  const bool act = lane < 2 * kMixN;
  double v = 1.0 / kMixN, eg = 0.0;                                  // own expert weight and AdaGrad state (LSInitType::Uniform)
  double r0 = 0.0, r1 = 0.0;
  const double beta = d.mix_beta, beta1 = 1.0 - d.mix_beta, mu = d.mu_mix;
  for (int t = 0; t < n; t++) {
    bar_sync(kBarS2MR, 96);
    const double target = S.mx_target, ep0 = S.mx_ep[0], ep1 = S.mx_ep[1];
    const double pi = S.mx_p[i];
    {
      const double er = target - (ex == 0 ? ep0 : ep1);
      const double loss = ex == 0 ? sgn(er) : er;
      const double grad = loss * pi;
      eg = fma(beta, eg, beta1 * grad * grad);
      v = fma(mu * rcp_fast(sqrt(eg) + 1e-5), grad, v);
    }
    r0 = fma(0.95, r0, (1.0 - 0.95) * (-fabs(target - ep0)));
    r1 = fma(0.95, r1, (1.0 - 0.95) * (-fabs(target - ep1)));
    const double mz = fmax(r0, r1);
    const double e0 = exp_neg(r0 - mz), e1 = exp_neg(r1 - mz);
    const double inv_total = rcp_fast(e0 + e1);
    const double s0 = e0 * inv_total, s1 = e1 * inv_total;
    // combined stage weight of input i: both experts' new weights meet in lane i
    const double vo = __shfl_sync(kFull, v, (lane + kMixN) & 31);    // lane i < 5 receives expert 1's weight of input i
    if (act) S.v[ex][i] = v;
    if (lane < kMixN) S.wi[lane] = fmax(fma(vo, s1, v * s0), 0.0);
    if (lane == 16) { S.sw[0] = s0; S.sw[1] = s1; }
    __threadfence_block();
    if (t + 1 < n) bar_arrive(kBarMR2S, 96);
  }
}

// R: RLS stage with adaptive forgetting (rls.cpp:17-65, rls.h:22-36). Lane j < m owns row j of P and w_j. Everything that
// does not depend on the sample's stage input bp4 is computed ahead of it: ph = P x, phi = x^T ph, and the two dot products
// of the NEXT prediction  p(t+1) = x_new . (w + c ph) = bp4 (w_0 + c ph_0) + Q1 + c Q2,  x_new = [bp4, x_0 .. x_{m-2}].
__device__ __forceinline__ void sg_rls_warp(SgShared &S, const ChainDesc &d, int lane, double *xch)
{
  const int n = d.n, m = d.lm_n;
  const int row = min(lane, m - 1);
  double *rx = xch, *rph = xch + kMaxRls + 2;                        // x (history of the stage input), ph = P x: shared by the lanes
  double P[kMaxRls];
#pragma unroll
  for (int k = 0; k < kMaxRls; k++) P[k] = (k == row) ? 1.0 : 0.0;
  if (lane < kMaxRls + 2) { rx[lane] = 0.0; rph[lane] = 0.0; }
  __syncwarp();
  double w = 0.0, S0 = 0.0, S1 = 0.0, p_rls = 0.0;
  const double gamma = d.lm_gamma;
  for (int t = 0; t < n; t++) {
    double phr = 0.0;
#pragma unroll
    for (int k = 0; k < kMaxRls; k++) if (k < m) phr = fma(P[k], rx[k], phr);
    if (lane < m) rph[lane] = phr;
    __syncwarp();
    double phi = 0.0;
#pragma unroll
    for (int k = 0; k < kMaxRls; k++) if (k < m) phi = fma(rx[k], rph[k], phi);
    phi = fmax(phi, 1e-8);
    const double rnis = rcp_fast(phi + fmax(S0 - S1, 1e-5));
    const double xprev = (row > 0 && lane < m) ? rx[row - 1] : 0.0;  // x_new[row] for rows 1..m-1
    double q1 = xprev * w, q2 = xprev * phr;                         // lanes 1..m-1 contribute; lane 0 and lanes >= m carry 0
    q1 += shfl_xor(q1, 8); q2 += shfl_xor(q2, 8);
    q1 += shfl_xor(q1, 4); q2 += shfl_xor(q2, 4);
    q1 += shfl_xor(q1, 2); q2 += shfl_xor(q2, 2);
    q1 += shfl_xor(q1, 1); q2 += shfl_xor(q2, 1);
    const double Q1 = shfl_idx(q1, 0), Q2 = shfl_idx(q2, 0);         // sums over lanes 0..15 (m <= 10)
    const double w_0 = shfl_idx(w, 0), ph_0 = rph[0];
    bar_sync(kBarS2MR, 96);
    const double bp4 = ldd_vol(&S.bp4);
    const double err = bp4 - p_rls;
    const double err2 = err * err;
    const double mm = exp_neg(-gamma * (err2 * rnis));
    const double al = fma(0.999 - 0.99, mm, 0.99);
    const double denom = rcp_fast(al + phi), inv_al = rcp_fast(al);
    const double c = err * denom;
    p_rls = fma(bp4, fma(c, ph_0, w_0), fma(c, Q2, Q1));
    if (lane == 0) S.p_rls = p_rls;
    __threadfence_block();
    if (t + 1 < n) bar_arrive(kBarMR2S, 96);
    // ---- off the critical path: w, P, x, S0, S1 ----
    w = fma(c, phr, w);                                              // w_row += err * denom * ph_row
    // symmetric by construction, as the reference's one-sided update mirrored (rls.cpp:44-52): ph_i * ph_j commutes
#pragma unroll
    for (int k = 0; k < kMaxRls; k++) if (k < m) P[k] = fma(-denom, phr * rph[k], P[k]) * inv_al;
    __syncwarp();                                                    // every lane has read rx / rph
    if (lane < m) rx[lane] = row > 0 ? xprev : bp4;
    __syncwarp();
    S0 = fma(0.95, S0, (1.0 - 0.95) * err2);
    S1 = fma(0.95, S1, (1.0 - 0.95) * phi);
  }
}

// B: bias stage (bias.h:64-162), round / clamp / residual (libsac.cpp:105-108), cost sums; handles sample t-1 during phase t
template <class CFG>
__device__ __forceinline__ void sg_bias_warp(SgShared &S, const ChainDesc &d, int lane)
{
  const int n = d.n;
  long long l1 = 0, sq = 0;
  int32_t vcur = lane < n ? __ldg(d.own + lane) : 0;
  int32_t vnext = 32 + lane < n ? __ldg(d.own + 32 + lane) : 0;
  int32_t ebuf = 0;
  const double nscale = (double)d.bias_nscale, bmu = d.bias_mu;
  for (int ph = 0; ph <= n; ph++) {
    if (ph > 0) {
      const int t = ph - 1;
      if ((t & 31) == 0 && t) { vcur = vnext; vnext = t + 32 + lane < n ? __ldg(d.own + t + 32 + lane) : 0; }
      const int32_t vali = __shfl_sync(kFull, vcur, t & 31);
      const double val = (double)vali;
      const double px = S.pxq[t & 1];
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
        // bias.h:103-108 compares sum/5 with 32 and 512; the deltas are integers, so sum/5 > 32 <=> sum > 160 exactly
        // (a multiplication by 0.2 is NOT equivalent: 160 * 0.2 > 32 in binary floating point)
        const double sum = fabs(d0) + fabs(d1) + fabs(d2) + fabs(d3) + fabs(d4);
        mix_ctx = sum > 2560 ? 2 : (sum > 160 ? 1 : 0);
        ctx0 = b0 + (b2 << 1) + (b9 << 2) + (b10 << 3) + (b11 << 4);
        ctx1 = b2 + (b3 << 1) + (b4 << 2);
        ctx2 = b5 + (b6 << 1) + (b7 << 2) + (b8 << 3);
      }
      const int tb = min(lane, 2);
      const int myctx = tb == 0 ? ctx0 : (tb == 1 ? ctx1 : ctx2);
      const double cv = S.cval[tb][myctx], cc = S.cnt[tb][myctx];
      const double ptl = cv * rcp_fast(cc);
      const double pt0 = shfl_idx(ptl, 0), pt1 = shfl_idx(ptl, 1), pt2 = shfl_idx(ptl, 2);
      const double pbias = fma(pt2, S.mixw[mix_ctx][2], fma(pt1, S.mixw[mix_ctx][1], pt0 * S.mixw[mix_ctx][0]));
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
      {
        const double delta = val - c_round(px);
        const double bv = fmax(0.0, S.bvar), bm = S.bmean;
        const double diff = delta - bm;
        const double z = diff * diff * rcp_fast(bv + 1E-5);
        const double wgt = exp_neg(-0.5 * z);
        double hin = 0.0, hdl = 0.0;
        if (lane < 8) { hin = lane > 0 ? S.hist_in[lane - 1] : val; hdl = lane > 0 ? S.hist_d[lane - 1] : delta; }
        __syncwarp();
        if (lane < 8) { S.hist_in[lane] = hin; S.hist_d[lane] = hdl; }
        if (lane < 3) {
          double ncv = fma(wgt, delta, cv);
          double ncc = cc + wgt;
          if (ncc >= nscale) { ncv *= 0.5; ncc *= 0.5; }
          S.cval[lane][myctx] = ncv; S.cnt[lane][myctx] = ncc;
          const double ptv = lane == 0 ? pt0 : (lane == 1 ? pt1 : pt2);
          S.mixw[mix_ctx][lane] += (bmu * sgn(delta - pbias)) * sgn(ptv);
        }
        if (lane == 0) {
          const double nm = fma(0.998, bm, (1.0 - 0.998) * delta);
          S.bvar = fma(0.998, S.bvar, (1.0 - 0.998) * ((delta - bm) * (delta - nm)));
          S.bmean = nm;
        }
        __syncwarp();
      }
    }
    bar_sync(kBarPhase, CFG::phase);
  }
  if (lane == 0) {
    if (d.l1sum) *d.l1sum = l1;
    if (d.sqsum) *d.sqsum = sq;
  }
}

__host__ __device__ inline int sg_block_len(int N, int tap_threads)  // taps j = 1..N-1 in blocks of odd length over the tap threads
{
  const int taps = N - 1;
  int L = (taps + tap_threads - 1) / tap_threads;
  if (L < 1) L = 1;
  return L | 1;
}
__host__ __device__ inline int sg_padded_taps(int N, int L)          // taps 1..np: whole blocks of L covering 1..N-1
{
  const int taps = N - 1 > 0 ? N - 1 : 0;
  return ((taps + L - 1) / L) * L;
}

template <class CFG>
__global__ void __launch_bounds__(CFG::threads, (CFG::tw == 4 ? 2 : 1)) cascade_sg_kernel(const ChainDesc *__restrict__ descs, const int *__restrict__ idx)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ChainDesc &d = descs[idx ? idx[blockIdx.x] : blockIdx.x];
  SgShared &S = *reinterpret_cast<SgShared *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    const size_t head = (sizeof(SgShared) + 15) & ~size_t(15);
    double *sp = reinterpret_cast<double *>(smem_raw + head);
    const int regs[kStages] = {CFG::r0, CFG::r1, CFG::r2, CFG::r3};
    for (int s = 0; s < kStages; s++) {
      const int N = d.vn[s];
      const int L = sg_block_len(N, CFG::tap_threads), np = sg_padded_taps(N, L);
      S.L[s] = L; S.M[s] = np + 3;                                   // ring: h_0 .. h_{np+1} are read, one more slot takes the push
      S.h[s] = sp; sp += (np + 3) + (L + 2);
      S.mu[s] = sp; sp += np + 1;
      S.pw[s] = sp; sp += np + 1;
      if (L > regs[s]) { S.wt[s] = sp; sp += np + 1; } else S.wt[s] = nullptr;
    }
    S.clamped = 0; S.bad = 0;
  }
  {
    double *z = reinterpret_cast<double *>(&S);
    const int nz = (int)(offsetof(SgShared, h) / 8);
    for (int i = tid; i < nz; i += CFG::threads) z[i] = 0.0;
  }
  __syncthreads();
  for (int s = 0; s < kStages; s++) {
    const int N = d.vn[s], np = S.M[s] - 3;
    const double md = d.vmudecay[s], pd = d.vpowdecay[s];
    double *h = S.h[s], *mu = S.mu[s], *pw = S.pw[s], *wt = S.wt[s];
    for (int i = tid; i <= np; i += CFG::threads) {                  // entries beyond the stage's taps are zero: padded taps do nothing
      pw[i] = i < N ? 1.0 / c_pow((double)(1 + i), pd) : 0.0;        // ls.h:39 (the canonical table values)
      mu[i] = i < N ? c_pow(md, (double)i) : 0.0;                    // ls.h:41
      if (wt) wt[i] = 0.0;
    }
    for (int i = tid; i < (np + 3) + (S.L[s] + 2); i += CFG::threads) h[i] = 0.0;
  }
  __syncthreads();
  if (tid < kStages) {                                               // sum_powtab accumulates sequentially (ls.h:40)
    const double *pw = S.pw[tid];
    const int N = d.vn[tid];
    double sp = 0.0;
    for (int i = 0; i < N; i++) sp += pw[i];
    S.sum_pow[tid] = sp;
  }
  if (tid < 2 * kMixN) S.v[tid / kMixN][tid % kMixN] = 1.0 / kMixN;
  if (tid < kMixN) S.wi[tid] = 1.0 / kMixN;                          // max(0.5 * 0.2 + 0.5 * 0.2, 0)
  if (tid < 2) S.sw[tid] = 0.5;
  if (tid < 64) { S.cnt[0][tid] = 4.0; S.cnt[1][tid] = 4.0; S.cnt[2][tid] = 4.0; }
  __syncthreads();
  if (warp < CFG::tw) sg_tap_warps<CFG>(S, d, tid);
  else if (warp == CFG::tw) sg_scalar_warp<CFG>(S, d, lane);
  else if (warp == CFG::tw + 1) sg_mix_warp(S, d, lane);
  else if (warp == CFG::tw + 2) sg_rls_warp(S, d, lane, S.rls_xch);
  else sg_bias_warp<CFG>(S, d, lane);
  __syncthreads();
  if (tid == 0 && d.flags) *d.flags = (S.bad ? 1 : 0) | (S.clamped ? 2 : 0);
}

// =====================================================================================================================
// OLS
// =====================================================================================================================
constexpr int kOXW = 256;                                            // input window (samples, power of two)
constexpr int kOKB = 4;                                              // samples per block

// ---------------------------------------------------------------------------------------------------------------------
// ols_warp_kernel<NP>: orders n <= NP <= 32, ONE WARP per chain, four independent chains per CTA, no CTA barrier in the loop.
// Lane i owns row i of the covariance and of the work matrix in REGISTERS (rows n..NP-1 are padding: zero regressors, pivot
// nu); the right-hand side is a column (one register per lane), so the forward substitution rides along with the
// elimination. Per column step a lane publishes its entry of column j through a double-buffered shared-memory vector, one
// __syncwarp, every lane reads the pivot and the column (128-bit broadcast loads) and applies the rank-1 update to its row:
// ~1 300 instructions per 32 x 32 solve for the whole chain (the block-cyclic kernel above spends ~30 000 over 8 warps).
// L goes to shared memory by rows for the back substitution.
struct OlsWarpShared {
  double xo[kOXW], xq[kOXW];
  double Xs[kOKB][32];
  double col[2][36];                                                 // column j at [0, 32), right-hand side of the pivot row at [32]
  double Ls[32][33];
};

template <int NP>
__global__ void __launch_bounds__(128, (NP <= 16 ? 4 : (NP <= 24 ? 3 : 2))) ols_warp_kernel(const ChainDesc *__restrict__ descs, const int *__restrict__ idx, int count)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int ci = blockIdx.x * 4 + wib;
  if (ci >= count) return;
  const ChainDesc &d = descs[idx ? idx[ci] : ci];
  OlsWarpShared &S = reinterpret_cast<OlsWarpShared *>(smem_raw)[wib];
  const int N = d.n;
  const int n = d.lenA + d.lenB;
  const double lambda = d.lambda, nu = d.nu, one_m_lambda = 1.0 - d.lambda;
  const int lenA = d.lenA, lagB = d.lagB, minB = d.minB, backB = d.backB;
  double cv[NP], W[NP];
#pragma unroll
  for (int c = 0; c < NP; c++) { cv[c] = 0.0; W[c] = 0.0; }
  double bcv = 0.0, wgt = 0.0, esum = 0.0;
  int km = 0;
  for (int q = lane; q < 192; q += 32) {
    const int ix = q - 64;
    const bool in = ix >= 0 && ix < N;
    S.xo[ix & (kOXW - 1)] = in ? (double)__ldg(d.own + ix) : 0.0;
    S.xq[ix & (kOXW - 1)] = in ? (double)__ldg(d.other + ix) : 0.0;
  }
  __syncwarp();
  int fill_end = 128;
  int t = 0;
  while (t < N) {
    if (fill_end < t + 68) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int ix = fill_end + lane + 32 * h;
        const bool in = ix < N;
        S.xo[ix & (kOXW - 1)] = in ? (double)__ldg(d.own + ix) : 0.0;
        S.xq[ix & (kOXW - 1)] = in ? (double)__ldg(d.other + ix) : 0.0;
      }
      fill_end += 64;
      __syncwarp();
    }
    const int kb = min(min(kOKB, d.k - km), N - t);
    // ---- this lane's regressor element of the block's samples (pred.cpp:17-31), the samples themselves ----
    double xi[kOKB], val[kOKB];
#pragma unroll
    for (int u = 0; u < kOKB; u++) {
      const int tt = t + u;
      const int sB = max(tt - lagB, minB) - backB;
      double v = 0.0;
      if (u < kb) {
        if (lane < lenA) v = S.xo[(tt - lenA + lane) & (kOXW - 1)];
        else if (lane < n) v = S.xq[(sB + lane - lenA) & (kOXW - 1)];
      }
      xi[u] = v;
      val[u] = u < kb ? S.xo[tt & (kOXW - 1)] : 0.0;
      S.Xs[u][lane] = v;
    }
    // ---- predictions (ols.cpp:22-25) ----
    double pu[kOKB];
#pragma unroll
    for (int u = 0; u < kOKB; u++) pu[u] = xi[u] * wgt;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
#pragma unroll
      for (int u = 0; u < kOKB; u++) pu[u] += shfl_xor(pu[u], o);
    }
    if (lane < kb) d.plpc[t + lane] = lane == 0 ? pu[0] : (lane == 1 ? pu[1] : (lane == 2 ? pu[2] : pu[3]));
    // ---- IRLS weights (ols.cpp:29-36): lane u computes the power of sample u ----
    double es_mine = 1.0;
#pragma unroll
    for (int u = 0; u < kOKB; u++)
      if (u < kb) {
        esum = fma(d.beta_sum, esum, fabs(val[u] - pu[u]));
        if (lane == u) es_mine = esum;
      }
    const double ffl = one_m_lambda * c_pow(es_mine + d.beta_add, -d.beta_pow);
    // ---- covariance: one rank-kb update of the lane's row and of its right-hand side (ols.cpp:38-45) ----
    km += kb;
    const bool solve = km >= d.k;
    {
      double fx[kOKB], lamk = 1.0;
#pragma unroll
      for (int u = kOKB - 1; u >= 0; u--) {
        const double f = u < kb ? shfl_idx(ffl, u) * lamk : 0.0;
        fx[u] = f * xi[u];
        if (u < kb) lamk *= lambda;
      }
      __syncwarp();                                                  // Xs complete
#pragma unroll
      for (int c = 0; c < NP; c += 2) {
        double2 x[kOKB];
#pragma unroll
        for (int u = 0; u < kOKB; u++) x[u] = *reinterpret_cast<const double2 *>(&S.Xs[u][c]);
        double a = cv[c] * lamk, b = cv[c + 1] * lamk;
#pragma unroll
        for (int u = 0; u < kOKB; u++) { a = fma(fx[u], x[u].x, a); b = fma(fx[u], x[u].y, b); }
        cv[c] = a; cv[c + 1] = b;
      }
      double bb = bcv * lamk;
#pragma unroll
      for (int u = 0; u < kOKB; u++) bb = fma(fx[u], val[u], bb);
      bcv = bb;
      __syncwarp();                                                  // Xs may be rewritten by the next block
    }
    if (solve) {
      km = 0;
#pragma unroll
      for (int c = 0; c < NP; c++) W[c] = cv[c] + (c == lane ? nu : 0.0);
      double y = bcv, z = 0.0;
      bool ok = true;
#pragma unroll
      for (int j = 0; j < NP; j++) {
        double *cb = S.col[j & 1];
        cb[lane] = W[j];
        if (lane == j) cb[32] = y;
        __syncwarp();
        const double dj = cb[j], yj = cb[32];
        if (dj < 1e-12) ok = false;
        const double inv = rcp_fast(dj);
        const double lij = W[j] * inv;
        if (lane == j) z = y * inv;                                  // (D^-1 L^-1 b)_j
        if (lane > j) { y = fma(-lij, yj, y); S.Ls[lane][j] = lij; }
        // trailing update of this lane's row: W[i][c] -= l_ij * W[c][j] (entries above the diagonal are never used)
        int c = j + 1;
        if (c & 1) { if (c < NP) W[c] = fma(-lij, cb[c], W[c]); c++; }
#pragma unroll
        for (; c + 1 < NP; c += 2) {
          const double2 u = *reinterpret_cast<const double2 *>(&cb[c]);
          W[c] = fma(-lij, u.x, W[c]);
          W[c + 1] = fma(-lij, u.y, W[c + 1]);
        }
        if (c < NP) W[c] = fma(-lij, cb[c], W[c]);
      }
      __syncwarp();
      // ---- back substitution L^T w = z: columns in descending order, row k of L is contiguous ----
      double lk = lane < NP - 1 ? S.Ls[NP - 1][lane] : 0.0;
#pragma unroll
      for (int k = NP - 1; k >= 1; k--) {
        const double cur = lk;
        if (k >= 2) lk = lane < k - 1 ? S.Ls[k - 1][lane] : 0.0;
        const double zk = shfl_idx(z, k);
        if (lane < k) z = fma(-cur, zk, z);
      }
      if (ok) wgt = z;
      __syncwarp();
    }
    t += kb;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// ols_rows_kernel<NW>: orders 33 .. 32 NW (NW = 2, 3), NW warps per chain, thread i owns ROW i of the work matrix in registers
// (rows n .. 32 NW - 1 are padding), the covariance lives in shared memory (lower triangle, packed by rows) and is updated
// once per block of k samples. Elimination as in ols_warp_kernel -- a column per step through a double-buffered shared
// vector -- with a named barrier over the NW warps per step; finished 8-column groups are skipped with one uniform branch
// each. L goes to shared memory (packed lower triangle) for the back substitution on warp 0.
template <int NW> struct OlsRowsShared {
  double xo[kOXW], xq[kOXW];
  double Xs[kOKB][32 * NW];
  double col[2][32 * NW + 4];
  double part[NW][kOKB], pu[kOKB], ff[kOKB];
  double z[32 * NW], wv[32 * NW];
};
__host__ __device__ inline size_t tri_index(int i) { return (size_t)i * (size_t)(i + 1) / 2; }   // row i of a packed lower triangle (diagonal included)

template <int NW>
__global__ void __launch_bounds__(32 * NW, (NW == 2 ? 4 : 2)) ols_rows_kernel(const ChainDesc *__restrict__ descs, const int *__restrict__ idx)
{
  constexpr int NR = 32 * NW;                                        // rows = columns held
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ChainDesc &d = descs[idx ? idx[blockIdx.x] : blockIdx.x];
  OlsRowsShared<NW> &S = *reinterpret_cast<OlsRowsShared<NW> *>(smem_raw);
  double *cov = reinterpret_cast<double *>(smem_raw + ((sizeof(OlsRowsShared<NW>) + 15) & ~size_t(15)));   // packed lower triangle, NR rows
  double *Ls = cov + tri_index(NR);                                  // packed strictly-lower triangle: row i at i(i-1)/2
  const int tid = threadIdx.x, lane = tid & 31, tw = tid >> 5;
  const int N = d.n;
  const int n = d.lenA + d.lenB;
  const double lambda = d.lambda, nu = d.nu, one_m_lambda = 1.0 - d.lambda;
  const int lenA = d.lenA, lagB = d.lagB, minB = d.minB, backB = d.backB;
  double *myrow = cov + tri_index(tid);                              // cov[tid][0 .. tid]
  for (int c = 0; c <= tid; c++) myrow[c] = 0.0;
  double W[NR];
#pragma unroll
  for (int c = 0; c < NR; c++) W[c] = 0.0;
  double bcv = 0.0, wgt = 0.0, esum = 0.0;
  int km = 0;
  for (int q = tid; q < 192; q += NR) {
    const int ix = q - 64;
    const bool in = ix >= 0 && ix < N;
    S.xo[ix & (kOXW - 1)] = in ? (double)__ldg(d.own + ix) : 0.0;
    S.xq[ix & (kOXW - 1)] = in ? (double)__ldg(d.other + ix) : 0.0;
  }
  int fill_end = 128;
  int t = 0;
  auto team = [&]() { bar_sync(1, NR); };
  team();
  while (t < N) {
    if (fill_end < t + 68) {
      if (tid < 64) {
        const int ix = fill_end + tid;
        const bool in = ix < N;
        S.xo[ix & (kOXW - 1)] = in ? (double)__ldg(d.own + ix) : 0.0;
        S.xq[ix & (kOXW - 1)] = in ? (double)__ldg(d.other + ix) : 0.0;
      }
      fill_end += 64;
      team();
    }
    const int kb = min(min(kOKB, d.k - km), N - t);
    double xi[kOKB], val[kOKB];
#pragma unroll
    for (int u = 0; u < kOKB; u++) {
      const int tt = t + u;
      const int sB = max(tt - lagB, minB) - backB;
      double v = 0.0;
      if (u < kb) {
        if (tid < lenA) v = S.xo[(tt - lenA + tid) & (kOXW - 1)];
        else if (tid < n) v = S.xq[(sB + tid - lenA) & (kOXW - 1)];
      }
      xi[u] = v;
      val[u] = u < kb ? S.xo[tt & (kOXW - 1)] : 0.0;
      S.Xs[u][tid] = v;
    }
    // ---- predictions: per-warp butterflies, then the NW partial sums ----
    {
      double pp[kOKB];
#pragma unroll
      for (int u = 0; u < kOKB; u++) pp[u] = xi[u] * wgt;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
#pragma unroll
        for (int u = 0; u < kOKB; u++) pp[u] += shfl_xor(pp[u], o);
      }
      if (lane < kOKB) S.part[tw][lane] = lane == 0 ? pp[0] : (lane == 1 ? pp[1] : (lane == 2 ? pp[2] : pp[3]));
    }
    team();                                                          // Xs and the partial sums are complete
    double pu[kOKB];
#pragma unroll
    for (int u = 0; u < kOKB; u++) {
      double s = S.part[0][u];
#pragma unroll
      for (int w = 1; w < NW; w++) s += S.part[w][u];
      pu[u] = s;
    }
    if (tid < kb) d.plpc[t + tid] = tid == 0 ? pu[0] : (tid == 1 ? pu[1] : (tid == 2 ? pu[2] : pu[3]));
    // ---- IRLS weights (ols.cpp:29-36): thread u computes the power of sample u ----
    double es_mine = 1.0;
#pragma unroll
    for (int u = 0; u < kOKB; u++)
      if (u < kb) {
        esum = fma(d.beta_sum, esum, fabs(val[u] - pu[u]));
        if (tid == u) es_mine = esum;
      }
    if (tid < kOKB) S.ff[tid] = one_m_lambda * c_pow(es_mine + d.beta_add, -d.beta_pow);
    team();
    // ---- covariance (ols.cpp:38-45): one rank-kb update of this thread's row (columns 0 .. tid) and right-hand side ----
    km += kb;
    const bool solve = km >= d.k;
    double lamk = 1.0;
    {
      double fx[kOKB];
#pragma unroll
      for (int u = kOKB - 1; u >= 0; u--) {
        const double f = u < kb ? S.ff[u] * lamk : 0.0;
        fx[u] = f * xi[u];
        if (u < kb) lamk *= lambda;
      }
      const int cend = min(tid, n - 1);                              // rows beyond n (and the part right of the diagonal) stay zero
      for (int c = 0; c <= cend; c++) {
        double a = myrow[c] * lamk;
#pragma unroll
        for (int u = 0; u < kOKB; u++) a = fma(fx[u], S.Xs[u][c], a);
        myrow[c] = a;
      }
      double bb = bcv * lamk;
#pragma unroll
      for (int u = 0; u < kOKB; u++) bb = fma(fx[u], val[u], bb);
      bcv = bb;
    }
    if (solve) {
      km = 0;
      team();                                                        // every row of the covariance is up to date
      // this thread's row of C + nu I, FULL row (the part right of the diagonal comes from the other rows' columns)
#pragma unroll
      for (int c = 0; c < NR; c++) {
        double v = 0.0;
        if (c <= tid) v = myrow[c];
        else if (c < n && tid < n) v = cov[tri_index(c) + tid];
        W[c] = v + (c == tid ? nu : 0.0);
      }
      double y = bcv, z = 0.0;
      bool ok = true;
      double wj = W[0];                                              // this row's entry of the pivot column, picked up during the previous step
      for (int j = 0; j < n; j++) {
        double *cb = S.col[j & 1];
        cb[tid] = wj;
        if (tid == j) cb[NR] = y;
        team();
        const double dj = cb[j], yj = cb[NR];
        if (dj < 1e-12) ok = false;
        const double inv = rcp_fast(dj);
        const double lij = wj * inv;
        if (tid == j) z = y * inv;
        if (tid > j) { y = fma(-lij, yj, y); Ls[tri_index(tid - 1) + j] = lij; }
        const int jn = j + 1, gn = jn >> 3;
        double wnext = 0.0;
#pragma unroll
        for (int g = 0; g < NR / 8; g++)
          if (8 * g + 7 > j) {                                       // uniform: groups left of the pivot are finished
#pragma unroll
            for (int k = 0; k < 8; k += 2) {
              const double2 u = *reinterpret_cast<const double2 *>(&cb[8 * g + k]);
              if (8 * g + k > j) W[8 * g + k] = fma(-lij, u.x, W[8 * g + k]);
              if (8 * g + k + 1 > j) W[8 * g + k + 1] = fma(-lij, u.y, W[8 * g + k + 1]);
            }
            if (g == gn) {                                           // uniform: the group of the next pivot column
              // W[jn] with a runtime jn, as predicated moves (a C++ select chain makes the compiler move W to local memory)
#pragma unroll
              for (int k = 0; k < 8; k++)
                asm("{ .reg .pred p; setp.eq.s32 p, %2, %3; selp.f64 %0, %1, %0, p; }" : "+d"(wnext) : "d"(W[8 * g + k]), "r"(jn), "r"(8 * g + k));
            }
          }
        wj = wnext;
      }
      S.z[tid] = z;
      team();
      if (tw == 0) {
        // back substitution L^T w = z on warp 0: slots lane, lane + 32, ...; row k of L is contiguous
        double zz[NW];
#pragma unroll
        for (int s = 0; s < NW; s++) zz[s] = lane + 32 * s < n ? S.z[lane + 32 * s] : 0.0;
        for (int k = n - 1; k >= 1; k--) {
          const double *row = Ls + tri_index(k - 1);
          const int sl = k >> 5;
          double zk = zz[0];
#pragma unroll
          for (int s = 1; s < NW; s++) if (sl == s) zk = zz[s];
          zk = shfl_idx(zk, k & 31);
#pragma unroll
          for (int s = 0; s < NW; s++) if (lane + 32 * s < k) zz[s] = fma(-row[lane + 32 * s], zk, zz[s]);
        }
#pragma unroll
        for (int s = 0; s < NW; s++) S.wv[lane + 32 * s] = zz[s];
      }
      team();
      if (ok) wgt = tid < n ? S.wv[tid] : 0.0;
    }
    team();                                                          // Xs, part, ff may be rewritten
    t += kb;
  }
}

} // namespace

// ---- host side -------------------------------------------------------------------------------------------------------
// shared memory a chain needs in cascade_sg_kernel<SMALL/LARGE>; 0 = not eligible for that variant
size_t cascade_sg_smem_bytes(const int *vn, int large)
{
  const int regs_s[kStages] = {SgSmall::r0, SgSmall::r1, SgSmall::r2, SgSmall::r3};
  const int regs_l[kStages] = {SgLarge::r0, SgLarge::r1, SgLarge::r2, SgLarge::r3};
  size_t doubles = 0;
  for (int s = 0; s < kStages; s++) {
    const int N = vn[s], L = sg_block_len(N, large ? SgLarge::tap_threads : SgSmall::tap_threads), np = sg_padded_taps(N, L);
    doubles += (size_t)(np + 3) + (size_t)(L + 2) + 2 * (size_t)(np + 1);
    if (L > (large ? regs_l[s] : regs_s[s])) doubles += (size_t)(np + 1);   // blocks longer than the register slots keep the rest of their weights here
  }
  return ((sizeof(SgShared) + 15) & ~size_t(15)) + doubles * 8 + 64;
}
// OLS kernel classes: orders up to 32 run one warp per chain (ols_warp_kernel<16 / 24 / 32>, classes 16 / 24 / 32); 64 = the
// two-warp experiment (ols_rows_kernel<2>); 0 = no search-grade kernel for this order (canonical team kernel)
int ols_sg_class(int n_ols)
{
  if (n_ols <= 16) return 16;
  if (n_ols <= 24) return 24;
  if (n_ols <= 32) return 32;
  if (n_ols <= 64) return 64;
  return 0;
}
size_t ols_sg_smem_bytes(int n_ols)
{
  if (n_ols <= 32) return 4 * sizeof(OlsWarpShared) + 64;
  if (n_ols <= 64) return ((sizeof(OlsRowsShared<2>) + 15) & ~size_t(15)) + (tri_index(64) + tri_index(63)) * 8 + 64;
  return 0;
}

cudaError_t predictor_sg_init_attributes()
{
  const int big = 227 * 1024;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(cascade_sg_kernel<SgSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(cascade_sg_kernel<SgLarge>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(ols_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(ols_warp_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(ols_warp_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(ols_warp_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
}

cudaError_t launch_ols_sg(const ChainDesc *d_descs, const int *d_idx, int count, int nb_class, int smem_bytes, cudaStream_t stream)
{
  if (count <= 0) return cudaSuccess;
  if (nb_class == 64) { ols_rows_kernel<2><<<count, 64, smem_bytes, stream>>>(d_descs, d_idx); return cudaGetLastError(); }
  const int grid = (count + 3) / 4;
  if (nb_class == 16) ols_warp_kernel<16><<<grid, 128, smem_bytes, stream>>>(d_descs, d_idx, count);
  else if (nb_class == 24) ols_warp_kernel<24><<<grid, 128, smem_bytes, stream>>>(d_descs, d_idx, count);
  else if (nb_class == 32) ols_warp_kernel<32><<<grid, 128, smem_bytes, stream>>>(d_descs, d_idx, count);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}
cudaError_t launch_cascade_sg(const ChainDesc *d_descs, const int *d_idx, int count, int large, int smem_bytes, cudaStream_t stream)
{
  if (count <= 0) return cudaSuccess;
  if (large) cascade_sg_kernel<SgLarge><<<count, SgLarge::threads, smem_bytes, stream>>>(d_descs, d_idx);
  else cascade_sg_kernel<SgSmall><<<count, SgSmall::threads, smem_bytes, stream>>>(d_descs, d_idx);
  return cudaGetLastError();
}

} // namespace sacb
