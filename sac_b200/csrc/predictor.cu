// predictor.cu -- DECODE direction of the per-sample fp64 recurrence (OLS -> 4-stage NLMS cascade + RLS -> 2-expert
// mix -> bias): in the decoder every stage needs the sample just reconstructed, so the recurrence cannot be split
// the way predictor_enc.cu splits the encoder's. One CTA per chain (coded channel), 160 threads:
//
//   warp 0      "scalar warp": OLS (regressor, dot, covariance, LDL^T), stage targets, mix/blend, RLS, bias,
//               round/clamp/residual -- everything that is O(n_ols^2) or O(1) per sample
//   warps 1..4  "tap warps": 128 lane-strided NLMS taps x 4 stages: predict dot + power sum (phase A) and the
//               weight update (phase C)
//
// Per sample the two groups meet at two named barriers (bar.arrive / bar.sync, no __syncthreads):
//   tap: A(t) -> arrive B1 .......... sync B2 -> C(t) -> A(t+1) ...
//   w0 : OLS dot(t) -> sync B1 -> predict/residual/targets/push -> arrive B2 -> updates(t), OLS solve, x(t+1) ...
// so the scalar warp's updates overlap the tap warps' C(t)+A(t+1).
//
// Arithmetic is the canonical order of DESIGN.md (fma placement and reduction trees spelled out; compiled with
// --fmad=false): the -m gpu tests compare residuals bit-for-bit with the CPU restatement kept under oracle/.
// Reference behaviour restated here: /root/reference src/libsac/pred.cpp:4-45, src/pred/{ols.cpp,ls.h,cascade.h,
// blend.h,rls.cpp,rls.h,bias.h}, src/common/math.h:14-78, src/libsac/libsac.cpp:94-199.
#include "chain.h"
#include "sac_canon_math.h"
#include <cuda_runtime.h>
#include <math_constants.h>
#include <cstddef>

namespace sacb {

using sac_canon::c_exp;
using sac_canon::c_pow;
using sac_canon::c_round;

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kBarB1 = 1, kBarB2 = 2;

__device__ __forceinline__ void bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"((int)kBlockThreads) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"((int)kBlockThreads) : "memory"); }

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

// slmath::dot for n < 8 and the AVX2 layout for 8 <= n <= 11 (math.h:130-161), scalar, for the tiny vectors of the
// mix / RLS / bias stages. x and y may live in shared memory.
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

// slmath::dot, n = 5, first operand in registers
__device__ __forceinline__ double dot5(const double (&x)[kMixN], const double *y)
{
  const double v1 = x[0] * y[0] + x[1] * y[1];
  const double v2 = x[2] * y[2] + x[3] * y[3];
  double init = 0.0 + (v1 + v2);
  init = init + x[4] * y[4];
  return 0.0 + init;
}

struct Stage {
  double *h, *w, *mu, *pw;   // history (capacity n+1, circular), weights, mu-decay table, power table
  int n, pos;                // pos: index of the newest element
  double mu_s, sum_pow;
};

// everything the scalar warp owns; lives in shared memory
struct Scalar {
  double part[kTapWarps][8];          // tap-warp partial sums: [warp][stage] dot, [warp][4+stage] s2pow
  double wgrad[kStages];
  double p[kMixN], spow[kStages], bp[kMixN];
  // mix (cascade.h:11-57, ls.h:214-241, blend.h)
  double v[2][kMixN], eg[2][kMixN], ep[2], sw[2], rsum[2];
  // RLS (rls.cpp)
  double rx[kMaxRls], rw[kMaxRls], rph[kMaxRls], rP[kMaxRls][kMaxRls];
  double S0, S1;
  // bias (bias.h)
  double hist_in[8], hist_d[8], pt[3], mixw[4][3];
  double cnt[3][64], cval[3][64];
  double bmean, bvar;
  // OLS
  double x[kMaxOls + 1];              // regressor, x[n] = current sample (augmented row)
  double ow[kMaxOls], od[kMaxOls + 1], lcol[kMaxOls + 1];
  double esum;
  Stage st[kStages];
};

__device__ __forceinline__ int wrap(int idx, int cap) { return idx >= cap ? idx - cap : idx; }

template <bool DECODE> __device__ __forceinline__ int32_t load_sample(const int32_t *p)
{
  if (DECODE) return *reinterpret_cast<const volatile int32_t *>(p);   // written during this launch (other CTA / own lane 0)
  return __ldg(p);
}

// packed lower-triangular index, row-major, rows 0..n (row n holds b)
__device__ __forceinline__ int tri(int i, int c) { return ((i * (i + 1)) >> 1) + c; }

// Remap::Unmap (map.cpp:188-202) in O(1): the |m|-th used sample value above (m > 0) or below (m < 0) the re-based
// prediction, from the cumulative counts C(v) = #{used u <= v} and the ascending list U of the used values (sparse.cu)
__device__ __forceinline__ int32_t unmap_residual(const ChainDesc &d, int32_t pi, int32_t m)
{
  if (m == 0) return 0;
  const int pred = pi + d.unmap_mean;
  const int nused = __ldg(d.unmap_cum + 65536);
  auto C = [&](int v) { return v < -32768 ? 0 : __ldg(d.unmap_cum + min(v, 32768) + 32768); };
  int idx = m > 0 ? C(pred) + m - 1 : C(pred - 1) + m;
  idx = min(max(idx, 0), nused - 1);          // a damaged stream can ask for a rank that does not exist (the reference would not return)
  return __ldg(d.unmap_list + idx) - pred;
}

// ------------------------------------------------------------------------------------------------------------------
template <bool DECODE>
__global__ void __launch_bounds__(kBlockThreads, 3) predictor_kernel(const ChainDesc *__restrict__ descs)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ChainDesc &d = descs[blockIdx.x];
  Scalar &S = *reinterpret_cast<Scalar *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_ols = d.lenA + d.lenB;
  const int ntri = ((n_ols + 1) * (n_ols + 2)) >> 1;   // augmented (n+1 rows)

  // ---- carve arrays: shared memory first (most re-used first), remainder in the chain's HBM scratch ----
  __shared__ double *s_cov, *s_wrk;
  if (tid == 0) {
    unsigned int dyn_bytes;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_bytes));
    double *sp = reinterpret_cast<double *>(smem_raw + ((sizeof(Scalar) + 15) & ~size_t(15)));
    long long s_left = ((long long)dyn_bytes - (long long)((sizeof(Scalar) + 15) & ~size_t(15))) / 8;
    double *gp = d.scratch;
    auto take = [&](long long cnt) -> double * {
      double *r;
      if (cnt <= s_left) { r = sp; sp += cnt; s_left -= cnt; }
      else { r = gp; gp += cnt; }
      return r;
    };
    for (int s = kStages - 1; s >= 1; s--) {
      Stage &st = S.st[s];
      st.n = d.vn[s];
      st.h = take(st.n + 1); st.w = take(st.n); st.pw = take(st.n); st.mu = take(st.n);
    }
    Stage &s0 = S.st[0];
    s0.n = d.vn[0];
    s0.w = take(s0.n); s0.h = take(s0.n + 1);
    s_cov = take(ntri); s_wrk = take(ntri);
    s0.pw = take(s0.n); s0.mu = take(s0.n);
  }
  __syncthreads();
  double *cov = s_cov, *wrk = s_wrk;

  // ---- initial state ----
  for (int s = 0; s < kStages; s++) {
    const Stage st = S.st[s];
    const double md = d.vmudecay[s], pd = d.vpowdecay[s];
    for (int i = tid; i < st.n; i += kBlockThreads) {
      st.h[i] = 0.0; st.w[i] = 0.0;
      st.pw[i] = 1.0 / c_pow((double)(1 + i), pd);        // ls.h:39
      st.mu[i] = c_pow(md, (double)i);                     // ls.h:41
    }
    if (tid == 0) st.h[st.n] = 0.0;
  }
  for (int i = tid; i < ntri; i += kBlockThreads) cov[i] = 0.0;
  {
    double *z = reinterpret_cast<double *>(&S);
    const int nz = (int)(offsetof(Scalar, st) / 8);
    for (int i = tid; i < nz; i += kBlockThreads) z[i] = 0.0;
  }
  __syncthreads();
  if (tid < kStages) {                                     // sum_powtab accumulates sequentially (ls.h:40)
    Stage &st = S.st[tid];
    double sp = 0.0;
    for (int i = 0; i < st.n; i++) sp += st.pw[i];
    st.sum_pow = sp; st.mu_s = d.vmu[tid]; st.pos = 0;
  }
  if (tid < 2 * kMixN) { S.v[tid / kMixN][tid % kMixN] = 1.0 / kMixN; }   // LSInitType::Uniform (ls.h:199-201)
  if (tid < 2) S.sw[tid] = 1.0 / 2;                                        // blend.h:22
  if (tid < d.lm_n) S.rP[tid][tid] = 1.0 / 1.0;                            // rls.cpp:14-15 (nu=1)
  if (tid < 64) { S.cnt[0][tid] = 4.0; S.cnt[1][tid] = 4.0; S.cnt[2][tid] = 4.0; }   // bias.h:24-28 (freq0=4)
  __syncthreads();

  const int n = d.n;

  if (warp != 0) {
    // =============================== tap warps =====================================================================
    const int tl = tid - 32;          // 0..127
    const int tw = warp - 1;
    int pos[kStages];
#pragma unroll
    for (int s = 0; s < kStages; s++) pos[s] = 0;
    for (int t = 0; t < n; t++) {
      // ---- phase A: lane-strided dot and power sum per stage ----
      double acc[8];
#pragma unroll
      for (int s = 0; s < kStages; s++) {
        const Stage &st = S.st[s];
        const int N = st.n, cap = N + 1;
        const double *h = st.h, *w = st.w, *pw = st.pw;
        double ad = 0.0, ap = 0.0;
        int hi = wrap(pos[s] + tl, cap);                    // only threads with tl < N iterate: pos+tl < 2*cap
        for (int i = tl; i < N; i += kTapThreads) {
          const double hv = h[hi];
          ad = __fma_rn(hv, w[i], ad);
          ap = __fma_rn(pw[i], hv * hv, ap);
          hi = wrap(hi + kTapThreads, cap);                 // a second iteration implies cap > 128
        }
        acc[s] = ad; acc[4 + s] = ap;
      }
#pragma unroll
      for (int q = 0; q < 8; q++) acc[q] = butterfly(acc[q]);
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 8; q++) S.part[tw][q] = acc[q];
      }
      __threadfence_block();
      bar_arrive(kBarB1);
      bar_sync(kBarB2);
      // ---- phase C: w_i = clamp(fma(mutab_i, wgrad*h_i, w_i), +-10) on the pre-push window (ls.h:49-54) ----
#pragma unroll
      for (int s = 0; s < kStages; s++) {
        const Stage &st = S.st[s];
        const int N = st.n, cap = N + 1;
        const double *h = st.h, *mu = st.mu;
        double *w = st.w;
        const double g = S.wgrad[s];
        int hi = wrap(pos[s] + tl, cap);
        for (int i = tl; i < N; i += kTapThreads) {
          const double tt = g * h[hi];
          double wn = __fma_rn(mu[i], tt, w[i]);
          wn = dmin(dmax(wn, -10.0), 10.0);
          w[i] = wn;
          hi = wrap(hi + kTapThreads, cap);
        }
        pos[s] = pos[s] == 0 ? N : pos[s] - 1;              // the scalar warp pushed bp[s] at this slot
      }
    }
    return;
  }

  // ================================= scalar warp =====================================================================
  const int32_t *own = d.own, *other = d.other;
  const double lambda = d.lambda, nu = d.nu;
  const double one_m_lambda = 1.0 - lambda;
  const int lm_n = d.lm_n;
  const double alpha = d.proj_alpha;
  int km = 0;
  long long l1 = 0, sq = 0;
  bool bad = false;
  double p_rls = 0.0;

  // regressor for sample t into S.x[0..n_ols)  (pred.cpp:17-31)
  auto build_x = [&](int t) {
    const int sB = max(t - d.lagB, d.minB) - d.backB;
    for (int j = lane; j < n_ols; j += 32) {
      int idx; const int32_t *src;
      if (j < d.lenA) { idx = t - d.lenA + j; src = own; } else { idx = sB + (j - d.lenA); src = other; }
      S.x[j] = (idx >= 0 && idx < n) ? (double)load_sample<DECODE>(src + idx) : 0.0;
    }
  };
  auto wait_other = [&](int t) {
    if (DECODE && d.wait_ctr) {
      const int need = min(n, max(t + d.wait_add, 0));
      if (lane == 0) { while (*reinterpret_cast<const volatile int *>(d.wait_ctr) < need) __nanosleep(64); }
      __syncwarp();
      __threadfence();
    }
  };

  wait_other(0);
  build_x(0);
  __syncwarp();

  for (int t = 0; t < n; t++) {
    // ---- OLS predict: 32 lane-strided fma chains + butterfly (ols.cpp:22-25) ----
    double p_lpc = 0.0;
    for (int j = lane; j < n_ols; j += 32) p_lpc = __fma_rn(S.x[j], S.ow[j], p_lpc);
    p_lpc = butterfly(p_lpc);

    bar_sync(kBarB1);
    // ---- stage predictions from the tap warps' partial sums ----
    if (lane < 8) {
      const double v = (S.part[0][lane] + S.part[1][lane]) + (S.part[2][lane] + S.part[3][lane]);
      if (lane < 4) S.p[lane] = v; else S.spow[lane - 4] = v;
    }
    if (lane == 8) S.p[4] = p_rls;
    __syncwarp();
    double p[kMixN];
#pragma unroll
    for (int i = 0; i < kMixN; i++) p[i] = S.p[i];
    // ---- mix predict (cascade.h:36-44, blend.h:25-30) ----
    const double ep0 = dot5(p, S.v[0]), ep1 = dot5(p, S.v[1]);
    if (!(fabs(ep0) <= 1.7976931348623157e308) || !(fabs(ep1) <= 1.7976931348623157e308)) bad = true;
    const double sw0 = S.sw[0], sw1 = S.sw[1];
    const double p_lms = 0.0 + ((0.0 + ep0 * sw0) + ep1 * sw1);
    const double px = p_lpc + p_lms;
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
    // ---- round, clamp, residual (libsac.cpp:105-108; NaN / overflow -> lo) ----
    int32_t pi;
    {
      const double r = c_round(pd);
      if (!(r >= (double)d.clamp_lo)) pi = d.clamp_lo;
      else if (r > (double)d.clamp_hi) pi = d.clamp_hi;
      else pi = (int32_t)r;
    }
    int32_t vali, e;
    if (DECODE) {
      e = d.err_in[t];
      if (d.unmap_cum) e = unmap_residual(d, pi, e);
      vali = pi + e;
      if (lane == 0) d.own_out[t] = vali;
    } else {
      vali = __ldg(own + t);
      e = vali - pi;
      if (lane == 0) d.resid[t] = e;
    }
    const double val = (double)vali;
    { const long long ae = e < 0 ? -(long long)e : (long long)e; l1 += ae; sq += (long long)e * (long long)e; }
    // ---- stage targets (cascade.h:99-113) ----
    const double target = val - p_lpc;
    double bp[kMixN];
    {
      double prefix = 0.0;
#pragma unroll
      for (int i = 0; i < kMixN; i++) {
        const double wi = dmax(0.0 + ((0.0 + S.v[0][i] * sw0) + S.v[1][i] * sw1), 0.0);
        const double pxi = (1.0 - alpha) * prefix + alpha * p_lms;
        bp[i] = target - dmin(dmax(pxi, d.casc_lo), d.casc_hi);
        prefix += wi * p[i];
      }
    }
    // ---- NLMS gradients + history push (ls.h:45-56) ----
    if (lane < kStages) {
      Stage &st = S.st[lane];
      double bpl = bp[0], pl = p[0];
      if (lane == 1) { bpl = bp[1]; pl = p[1]; } else if (lane == 2) { bpl = bp[2]; pl = p[2]; } else if (lane == 3) { bpl = bp[3]; pl = p[3]; }
      S.wgrad[lane] = st.mu_s * (bpl - pl) * st.sum_pow / (S.spow[lane] + 1.0);
      const int np = st.pos == 0 ? st.n : st.pos - 1;
      st.h[np] = bpl;
      st.pos = np;
    }
    __threadfence_block();
    bar_arrive(kBarB2);

    // ================= updates for sample t (overlap the tap warps' C(t) + A(t+1)) =================================
    // ---- RLS stage: UpdateHist(bp[4]) (rls.cpp:29-65, rls.h:22-36) ----
    {
      const double err = bp[4] - p_rls;
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
        xprev = lane > 0 ? S.rx[lane - 1] : bp[4];
      }
      __syncwarp();
      if (lane < lm_n) S.rx[lane] = xprev;                 // RollBack (utils.h:330-336)
      if (lane == 0) {
        S.S0 = 0.95 * S.S0 + (1.0 - 0.95) * err2;
        S.S1 = 0.95 * S.S1 + (1.0 - 0.95) * phi;
      }
      __syncwarp();
      p_rls = small_dot(S.rx, S.rw, lm_n);                 // prediction for t+1 (rls.cpp:17-21)
    }
    // ---- mix update: LS_ADA<L1>, LS_ADA<L2> (ls.h:223-237), BlendExp (blend.h:50-90) ----
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
    }
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
    }
    // ---- OLS update (ols.cpp:27-57): augmented covariance rows 0..n (row n = b) ----
    {
      if (lane == 0) S.x[n_ols] = val;
      const double eo = val - p_lpc;
      const double es = d.beta_sum * S.esum + fabs(eo);
      const double c = c_pow(es + d.beta_add, -d.beta_pow);
      const double ff = one_m_lambda * c;
      __syncwarp();
      if (lane == 0) S.esum = es;
      {
        int i = 0, c2 = lane;                               // walk packed (row i, col c2) in steps of 32
        while (c2 > i) { c2 -= (i + 1); i++; }
        for (int q = lane; q < ntri; q += 32) {
          if (!(i == n_ols && c2 == n_ols))
            cov[q] = lambda * cov[q] + ff * (S.x[i] * S.x[c2]);
          c2 += 32;
          while (c2 > i) { c2 -= (i + 1); i++; }
        }
      }
      km++;
      if (km >= d.k) {
        km = 0;
        __syncwarp();
        // right-looking LDL^T of (C + nu I) on the augmented matrix; row n carries b -> z = D^-1 L^-1 b
        {
          int i = 0, c2 = lane;
          while (c2 > i) { c2 -= (i + 1); i++; }
          for (int q = lane; q < ntri; q += 32) {
            double vq = cov[q];
            if (i == c2 && i < n_ols) vq = vq + nu;
            wrk[q] = vq;
            c2 += 32;
            while (c2 > i) { c2 -= (i + 1); i++; }
          }
        }
        __syncwarp();
        bool ok = true;
        for (int j = 0; j < n_ols; j++) {
          const double dj = wrk[tri(j, j)];
          if (dj < 1e-12) { ok = false; break; }
          const double inv = 1.0 / dj;
          for (int i = j + 1 + lane; i <= n_ols; i += 32) S.lcol[i] = wrk[tri(i, j)] * inv;
          __syncwarp();
          {
            const int m = n_ols - j;                        // trailing rows j+1..n_ols, cols j+1..row
            const int cnt = (m * (m + 1)) >> 1;
            int i = 0, c2 = lane;
            while (c2 > i) { c2 -= (i + 1); i++; }
            for (int q = lane; q < cnt; q += 32) {
              const int ri = j + 1 + i, cc = j + 1 + c2;
              if (!(ri == n_ols && cc == n_ols)) {
                const int o = tri(ri, cc);
                wrk[o] = __fma_rn(-S.lcol[ri], wrk[tri(cc, j)], wrk[o]);
              }
              c2 += 32;
              while (c2 > i) { c2 -= (i + 1); i++; }
            }
          }
          __syncwarp();
          for (int i = j + 1 + lane; i <= n_ols; i += 32) wrk[tri(i, j)] = S.lcol[i];
          __syncwarp();
        }
        if (ok) {
          // back substitution L^T w = z, columns applied in descending order, one fma per element
          for (int i = lane; i < n_ols; i += 32) S.od[i] = wrk[tri(n_ols, i)];
          __syncwarp();
          for (int kk = n_ols - 1; kk >= 0; kk--) {
            const double wk = S.od[kk];
            for (int i = lane; i < kk; i += 32) S.od[i] = __fma_rn(-wrk[tri(kk, i)], wk, S.od[i]);
            __syncwarp();
          }
          for (int i = lane; i < n_ols; i += 32) S.ow[i] = S.od[i];
        }
      }
    }
    __syncwarp();
    if (DECODE) {
      __threadfence();
      if (lane == 0) *reinterpret_cast<volatile int *>(d.progress) = t + 1;
    }
    if (t + 1 < n) { wait_other(t + 1); build_x(t + 1); }
    __syncwarp();
  }
  if (lane == 0) {
    if (d.l1sum) *d.l1sum = l1;
    if (d.sqsum) *d.sqsum = sq;
    if (d.flags) *d.flags = bad ? 1 : 0;
  }
}

} // namespace

size_t predictor_scalar_bytes() { return (sizeof(Scalar) + 15) & ~size_t(15); }

cudaError_t predictor_init_attributes()
{
  return cudaFuncSetAttribute(predictor_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
}
// decode direction only: the encode direction is predictor_enc.cu (ols_kernel + cascade_kernel)
cudaError_t launch_predictor_decode(const ChainDesc *d_descs, int nchains, int smem_bytes, cudaStream_t stream)
{
  predictor_kernel<true><<<nchains, kBlockThreads, smem_bytes, stream>>>(d_descs);
  return cudaGetLastError();
}

} // namespace sacb
