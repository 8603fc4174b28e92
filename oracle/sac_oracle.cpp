// TEST INFRASTRUCTURE ONLY -- never linked into, loaded by or executed from the product path. See sac_oracle.h.
//
// Plain scalar restatement of the reference algorithm. "ref <file>:<lines>" comments give the reference code each
// function follows (paths relative to /root/reference/src). Build: oracle/Makefile (-ffp-contract=off, so that
// every fused multiply-add below is an explicit fma() and nothing else is fused).
#include "sac_oracle.h"
#include "sac_canon_math.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <random>
#include <vector>

namespace {

int g_order = SACO_ORDER_REF;
int g_math = SACO_MATH_LIBM;

inline bool b200() { return g_order == SACO_ORDER_B200; }
inline double m_exp(double x) { return g_math == SACO_MATH_CANON ? sac_canon::c_exp(x) : std::exp(x); }
inline double m_pow(double x, double y) { return g_math == SACO_MATH_CANON ? sac_canon::c_pow(x, y) : std::pow(x, y); }
inline double m_round(double x) { return g_math == SACO_MATH_CANON ? sac_canon::c_round(x) : std::round(x); }
inline double sgn(double x) { return (double)((x > 0) - (x < 0)); } // ref common/utils.h:156-163

typedef std::vector<double> vec;

// ---------------------------------------------------------------------------------------------------------------
// inner products
// ---------------------------------------------------------------------------------------------------------------

// ref common/math.h:130-161 (slmath::dot): AVX2 path = two 4-lane fma accumulators over blocks of 8, lane-wise sum,
// left-to-right horizontal sum; tail by std::transform_reduce (libstdc++ numeric:378-391: blocks of 4 as
// (a0+a1)+(a2+a3) added to init, then one by one), result added to the AVX total.
double tail_reduce(const double *x, const double *y, size_t i, size_t n, bool sq)
{
  double init = 0.0;
  auto term = [&](size_t k) { return sq ? y[k] * (x[k] * x[k]) : x[k] * y[k]; };
  while (n - i >= 4) {
    double v1 = term(i) + term(i + 1);
    double v2 = term(i + 2) + term(i + 3);
    init = init + (v1 + v2);
    i += 4;
  }
  for (; i < n; i++) init = init + term(i);
  return init;
}
double dot_ref(const double *x, const double *y, size_t n)
{
  double total = 0.0;
  size_t i = 0;
  if (n >= 8) {
    double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    for (; i + 8 <= n; i += 8)
      for (int l = 0; l < 4; l++) {
        s1[l] = std::fma(x[i + l], y[i + l], s1[l]);
        s2[l] = std::fma(x[i + 4 + l], y[i + 4 + l], s2[l]);
      }
    for (int l = 0; l < 4; l++) s1[l] = s1[l] + s2[l];
    total = s1[0] + s1[1] + s1[2] + s1[3];
  }
  total += tail_reduce(x, y, i, n, false);
  return total;
}
// ref common/math.h:164-191 (slmath::calc_s2pow): one 4-lane accumulator, fma(pow, x*x, sum)
double s2pow_ref(const double *x, const double *powtab, size_t n)
{
  double spow = 0.0;
  size_t i = 0;
  if (n >= 8) {
    double s[4] = {0, 0, 0, 0};
    for (; i + 4 <= n; i += 4)
      for (int l = 0; l < 4; l++) s[l] = std::fma(powtab[i + l], x[i + l] * x[i + l], s[l]);
    spow = s[0] + s[1] + s[2] + s[3];
  }
  spow += tail_reduce(x, powtab, i, n, true);
  return spow;
}

// B200 order (DESIGN.md "canonical arithmetic"): V lane-strided fma chains (element i feeds chain i mod V, in
// ascending i), then per group of 32 chains a butterfly (offsets 16,8,4,2,1), then the group sums pairwise.
double tree_b200(double *a, int V)
{
  for (int g = 0; g < V; g += 32)
    for (int off = 16; off >= 1; off >>= 1)
      for (int l = 0; l < off; l++) a[g + l] = a[g + l] + a[g + l + off];
  int ng = V / 32;
  double s[8];
  for (int g = 0; g < ng; g++) s[g] = a[g * 32];
  for (; ng > 1; ng >>= 1)
    for (int g = 0; g < ng / 2; g++) s[g] = s[2 * g] + s[2 * g + 1];
  return s[0];
}
double dot_b200(const double *x, const double *y, size_t n, int V)
{
  double a[128];
  for (int l = 0; l < V; l++) a[l] = 0.0;
  for (size_t i = 0; i < n; i++) a[i % V] = std::fma(x[i], y[i], a[i % V]);
  return tree_b200(a, V);
}
double s2pow_b200(const double *x, const double *powtab, size_t n, int V)
{
  double a[128];
  for (int l = 0; l < V; l++) a[l] = 0.0;
  for (size_t i = 0; i < n; i++) a[i % V] = std::fma(powtab[i], x[i] * x[i], a[i % V]);
  return tree_b200(a, V);
}

enum { V_NLMS = 128, V_OLS = 32 };

// ---------------------------------------------------------------------------------------------------------------
// OLS  (ref pred/ols.cpp:7-57, common/math.h:14-78 LDLT, common/utils.h:39-73 RunSumGEO)
// ---------------------------------------------------------------------------------------------------------------
struct Ols {
  int n, kmax, km = 0;
  double lambda, nu, beta_sum, beta_pow, beta_add;
  double esum = 0.0, pred = 0.0;
  vec x, w, b, D, invD, y, z;
  std::vector<vec> mcov, L;
  Ols(int n_, int kmax_, double lambda_, double nu_, double bsum, double bpow, double badd)
      : n(n_), kmax(kmax_), lambda(lambda_), nu((1.0 - lambda_) * nu_), beta_sum(bsum), beta_pow(bpow), beta_add(badd),
        x(n_), w(n_), b(n_), D(n_), invD(n_), y(n_), z(n_), mcov(n_, vec(n_)), L(n_, vec(n_)) {}

  double Predict()
  {
    pred = b200() ? dot_b200(x.data(), w.data(), n, V_OLS) : dot_ref(x.data(), w.data(), n);
    return pred;
  }
  // ref math.h:21-51: left-looking LDL^T of (A + nu I); pivot < 1e-12 -> fail
  bool FactorRef()
  {
    for (int i = 0; i < n; i++) { std::fill_n(L[i].begin(), i, 0.0); D[i] = 0.0; }
    for (int j = 0; j < n; j++) {
      double dj = mcov[j][j] + nu;
      for (int k = 0; k < j; k++) dj -= L[j][k] * L[j][k] * D[k];
      if (dj < 1e-12) return false;
      const double inv = 1.0 / dj;
      D[j] = dj; invD[j] = inv;
      for (int i = j + 1; i < n; i++) {
        double lij = mcov[i][j];
        for (int k = 0; k < j; k++) lij -= L[i][k] * L[j][k] * D[k];
        L[i][j] = lij * inv;
      }
    }
    return true;
  }
  // ref math.h:53-73
  void SolveRef()
  {
    for (int i = 0; i < n; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[i][k] * y[k]; y[i] = s; }
    for (int i = 0; i < n; i++) z[i] = y[i] * invD[i];
    for (int i = n - 1; i >= 0; i--) { double s = z[i]; for (int k = i + 1; k < n; k++) s -= L[k][i] * w[k]; w[i] = s; }
  }
  // B200 order: right-looking LDL^T on the matrix augmented with b as row n (so the forward substitution and the
  // D^-1 scaling fall out of the factorisation: row n of L is z = D^-1 L^-1 b). Same pivots and column scaling as
  // the reference; the rank-1 trailing update uses the unscaled column W[c][j] (= L[c][j]*D[j] up to rounding) and
  // one fma per element, columns applied in ascending j. Back substitution column-oriented, k descending.
  std::vector<vec> W;
  vec lcol;
  bool FactorSolveB200()
  {
    if (W.empty()) { W.assign(n + 1, vec(n + 1)); lcol.assign(n + 1, 0.0); }
    for (int i = 0; i < n; i++) { for (int c = 0; c <= i; c++) W[i][c] = mcov[i][c]; W[i][i] = W[i][i] + nu; }
    for (int c = 0; c < n; c++) W[n][c] = b[c];
    for (int j = 0; j < n; j++) {
      const double dj = W[j][j];
      if (dj < 1e-12) return false;
      const double inv = 1.0 / dj;
      for (int i = j + 1; i <= n; i++) lcol[i] = W[i][j] * inv;
      for (int i = j + 1; i <= n; i++)
        for (int c = j + 1; c <= i && c < n; c++) W[i][c] = std::fma(-lcol[i], W[c][j], W[i][c]);
      for (int i = j + 1; i <= n; i++) W[i][j] = lcol[i];
    }
    for (int i = 0; i < n; i++) y[i] = W[n][i];
    for (int k = n - 1; k >= 0; k--)
      for (int i = 0; i < k; i++) y[i] = std::fma(-W[k][i], y[k], y[i]);
    for (int i = 0; i < n; i++) w[i] = y[i];
    return true;
  }
  // ref ols.cpp:27-57
  void Update(double val)
  {
    const double e = val - pred;
    esum = beta_sum * esum + std::fabs(e);
    const double c = m_pow(esum + beta_add, -beta_pow);
    const double ff = (1.0 - lambda) * c;
    for (int j = 0; j < n; j++) {
      const double xj = x[j];
      for (int i = 0; i <= j; i++) mcov[j][i] = lambda * mcov[j][i] + ff * (xj * x[i]);
      b[j] = lambda * b[j] + ff * (xj * val);
    }
    km++;
    if (km >= kmax) {
      if (b200()) FactorSolveB200();
      else { if (FactorRef()) SolveRef(); }
      km = 0;
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// NLMS stage (ref pred/ls.h:10-62, common/histbuf.h:61-88)
// ---------------------------------------------------------------------------------------------------------------
struct Nlms {
  int n;
  double mu, sum_powtab = 0.0, pred = 0.0;
  vec h, w, mutab, powtab; // h[0] newest
  Nlms(int n_, double mu_, double mu_decay, double pow_decay) : n(n_), mu(mu_), h(n_), w(n_), mutab(n_), powtab(n_)
  {
    for (int i = 0; i < n; i++) {                       // ref ls.h:37-42
      powtab[i] = 1.0 / m_pow((double)(1 + i), pow_decay);
      sum_powtab += powtab[i];
      mutab[i] = m_pow(mu_decay, (double)i);
    }
  }
  double Predict()
  {
    pred = b200() ? dot_b200(h.data(), w.data(), n, V_NLMS) : dot_ref(h.data(), w.data(), n);
    return pred;
  }
  void Update(double val)                               // ref ls.h:45-56
  {
    const double spow = b200() ? s2pow_b200(h.data(), powtab.data(), n, V_NLMS) : s2pow_ref(h.data(), powtab.data(), n);
    const double wgrad = mu * (val - pred) * sum_powtab / (spow + 1.0);
    for (int i = 0; i < n; i++) {
      const double t = wgrad * h[i];
      double wn = b200() ? std::fma(mutab[i], t, w[i]) : w[i] + mutab[i] * t;
      w[i] = std::min(std::max(wn, -10.0), 10.0);
    }
    std::memmove(&h[1], &h[0], (n - 1) * sizeof(double)); // push: newest at index 0
    h[0] = val;
  }
};

// ---------------------------------------------------------------------------------------------------------------
// RLS 5th stage with adaptive forgetting (ref pred/rls.cpp:6-65, pred/rls.h:13-40)
// ---------------------------------------------------------------------------------------------------------------
struct Rls {
  int n;
  double gamma, beta, S0 = 0.0, S1 = 0.0, px = 0.0;
  vec x, w, ph;
  std::vector<vec> P;
  Rls(int n_, double gamma_, double beta_) : n(n_), gamma(gamma_), beta(beta_), x(n_), w(n_), ph(n_), P(n_, vec(n_))
  {
    for (int i = 0; i < n; i++) P[i][i] = 1.0 / 1.0;
  }
  double Predict() { px = dot_ref(x.data(), w.data(), n); return px; }
  void UpdateHist(double val)
  {
    const double err = val - px;
    for (int i = 0; i < n; i++) ph[i] = dot_ref(P[i].data(), x.data(), n);
    const double phi = std::max(dot_ref(x.data(), ph.data(), n), 1e-8);
    const double err2 = err * err;
    const double R = std::max(S0 - S1, 1e-5);                       // rls.h:22-30
    const double nis = err2 / (phi + R);
    const double m = m_exp(-gamma * nis);
    const double alpha = 0.99 + (0.999 - 0.99) * m;
    const double denom = 1. / (alpha + phi);
    const double inv_alpha = 1.0 / alpha;
    for (int i = 0; i < n; i++)
      for (int j = 0; j <= i; j++) {
        const double mm = ph[i] * ph[j];
        const double v = (P[i][j] - denom * mm) * inv_alpha;
        P[i][j] = P[j][i] = v;
      }
    for (int i = 0; i < n; i++) w[i] += err * (denom * ph[i]);
    S0 = beta * S0 + (1.0 - beta) * err2;                           // rls.h:31-36
    S1 = beta * S1 + (1.0 - beta) * phi;
    if (n) { std::memmove(&x[1], &x[0], (n - 1) * sizeof(double)); x[0] = val; } // utils.h:330-336
  }
};

// ---------------------------------------------------------------------------------------------------------------
// 2-expert mix: LS_ADA<L1>, LS_ADA<L2> blended by softmax of EMA loss
// (ref pred/cascade.h:11-69, pred/ls.h:171-241, pred/blend.h:11-96)
// ---------------------------------------------------------------------------------------------------------------
struct Mix {
  int n;
  double mu, beta, beta1;
  vec wv[2], eg[2];
  double ep[2] = {0, 0}, sw[2] = {0.5, 0.5}, rsum[2] = {0, 0};
  bool bad = false;
  Mix(int n_, double mu_, double beta_) : n(n_), mu(mu_), beta(beta_), beta1(1.0 - beta_)
  {
    for (int e = 0; e < 2; e++) { wv[e].assign(n, 1.0 / n); eg[e].assign(n, 0.0); }
  }
  double GetWeight(int idx) const                                   // cascade.h:24-34
  {
    double ew[2] = {wv[0][idx], wv[1][idx]};
    return std::max(dot_ref(ew, sw, 2), 0.0);
  }
  double Predict(const double *p)                                   // cascade.h:36-44, blend.h:25-30
  {
    for (int e = 0; e < 2; e++) {
      ep[e] = dot_ref(p, wv[e].data(), n);
      if (!std::isfinite(ep[e])) bad = true;
    }
    return dot_ref(ep, sw, 2);
  }
  void Update(const double *p, double target)                       // cascade.h:45-49
  {
    for (int e = 0; e < 2; e++) {                                   // ls.h:223-237
      const double err = target - ep[e];
      const double loss = e == 0 ? sgn(err) : err;
      for (int i = 0; i < n; i++) {
        const double grad = loss * p[i];
        eg[e][i] = beta * eg[e][i] + beta1 * grad * grad;
        const double mu_scaled = mu / (std::sqrt(eg[e][i]) + 1e-5);
        wv[e][i] += mu_scaled * grad;
      }
    }
    for (int e = 0; e < 2; e++) {                                   // blend.h:50-61, utils.h:46-52 (alpha=.95)
      const double loss = std::fabs(target - ep[e]);
      rsum[e] = 0.95 * rsum[e] + (1.0 - 0.95) * (-loss);
    }
    double zm[2], max_z = -std::numeric_limits<double>::infinity(); // blend.h:72-90 (beta=1.0)
    for (int e = 0; e < 2; e++) { zm[e] = 1.0 * rsum[e]; max_z = std::max(max_z, zm[e]); }
    double total = 0.0;
    for (int e = 0; e < 2; e++) { sw[e] = m_exp(zm[e] - max_z); total += sw[e]; }
    const double inv_total = 1.0 / total;
    for (int e = 0; e < 2; e++) sw[e] *= inv_total;
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Cascade (ref pred/cascade.h:75-131)
// ---------------------------------------------------------------------------------------------------------------
struct Cascade {
  double lo, hi, p_alpha, pred = 0.0;
  std::vector<Nlms> st;
  Rls lm;
  Mix mix;
  double p[5], bp[5];
  Cascade(int lo_, int hi_, const int *vn, const double *vmu, const double *vmudecay, const double *vpowdecay,
          double mu_mix, double mu_mix_beta, int lm_n, double lm_alpha, double proj_alpha)
      : lo(lo_), hi(hi_), p_alpha(proj_alpha), lm(lm_n, lm_alpha, 0.95), mix(5, mu_mix, mu_mix_beta)
  {
    for (int i = 0; i < 4; i++) st.emplace_back(vn[i], vmu[i], vmudecay[i], vpowdecay[i]);
  }
  double Predict()
  {
    for (int i = 0; i < 4; i++) p[i] = st[i].Predict();
    p[4] = lm.Predict();
    return (pred = mix.Predict(p));
  }
  void Update(double target)
  {
    double p_prefix = 0.0;
    for (int i = 0; i <= 4; i++) {
      const double w = mix.GetWeight(i);
      const double px = (1.0 - p_alpha) * p_prefix + p_alpha * pred;
      bp[i] = target - std::min(std::max(px, lo), hi);
      p_prefix += w * p[i];
    }
    for (int i = 0; i < 4; i++) st[i].Update(bp[i]);
    lm.UpdateHist(bp[4]);
    mix.Update(p, target);
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Bias estimator (ref pred/bias.h:16-175, common/utils.h:75-110 RunMeanVar, pred/ls.h:279-292 SSLMS)
// ---------------------------------------------------------------------------------------------------------------
struct Bias {
  struct Cnt { double cnt = 4.0, val = 0.0; };
  double mu;
  int nscale;
  double hist_input[8] = {0}, hist_delta[8] = {0}, pt[3] = {0, 0, 0};
  double mixw[4][3] = {{0}};
  Cnt cnt0[64], cnt1[64], cnt2[64];
  int ctx0 = 0, ctx1 = 0, ctx2 = 0, mix_ctx = 0;
  double px = 0.0, pbias = 0.0, mean = 0.0, var = 0.0;
  Bias(double mu_, int nb_scale) : mu(mu_), nscale(1 << nb_scale) {}
  void CalcContext(double p)                                        // bias.h:64-113
  {
    const double *hi = hist_input, *hd = hist_delta;
    int b0 = hi[0] > p ? 0 : 1;
    int b2 = hd[0] < 0 ? 0 : 1, b3 = hd[1] < 0 ? 0 : 1, b4 = hd[2] < 0 ? 0 : 1;
    int b5 = hd[1] < hd[0] ? 0 : 1, b6 = hd[2] < hd[1] ? 0 : 1, b7 = hd[3] < hd[2] ? 0 : 1, b8 = hd[4] < hd[3] ? 0 : 1;
    int b9 = (std::fabs(hd[0])) > 32 ? 0 : 1;
    int b10 = 2 * hi[0] - hi[1] > p ? 0 : 1;
    int b11 = 3 * hi[0] - 3 * hi[1] + hi[2] > p ? 0 : 1;
    double sum = 0;
    for (int i = 0; i < 5; i++) sum += std::fabs(hd[i]);
    sum /= 5.0;
    int t = 0;
    if (sum > 512) t = 2; else if (sum > 32) t = 1;
    ctx0 = b0 + (b2 << 1) + (b9 << 2) + (b10 << 3) + (b11 << 4);
    ctx1 = b2 + (b3 << 1) + (b4 << 2);
    ctx2 = b5 + (b6 << 1) + (b7 << 2) + (b8 << 3);
    mix_ctx = t;
  }
  double Predict(double pred)                                       // bias.h:114-126
  {
    px = pred;
    CalcContext(pred);
    pt[0] = cnt0[ctx0].val / cnt0[ctx0].cnt;
    pt[1] = cnt1[ctx1].val / cnt1[ctx1].cnt;
    pt[2] = cnt2[ctx2].val / cnt2[ctx2].cnt;
    pbias = dot_ref(pt, mixw[mix_ctx], 3);
    return px + pbias;
  }
  void upd(Cnt &c, double delta, double w)                          // bias.h:33-40
  {
    c.val += w * delta;
    c.cnt += w;
    if (c.cnt >= nscale) { c.val *= 0.5; c.cnt *= 0.5; }
  }
  void Update(double val)                                           // bias.h:127-162
  {
    const double delta = val - m_round(px);
    std::memmove(&hist_input[1], &hist_input[0], 7 * sizeof(double)); hist_input[0] = val;
    std::memmove(&hist_delta[1], &hist_delta[0], 7 * sizeof(double)); hist_delta[0] = delta;
    const double v = std::max(0.0, var);
    const double diff = delta - mean;
    const double z = diff * diff / (v + 1E-5);
    const double w = m_exp(-0.5 * z);
    upd(cnt0[ctx0], delta, w); upd(cnt1[ctx1], delta, w); upd(cnt2[ctx2], delta, w);
    const double old_mean = mean;                                   // utils.h:92-96 (alpha=.998)
    mean = 0.998 * mean + (1.0 - 0.998) * delta;
    var = 0.998 * var + (1.0 - 0.998) * ((delta - old_mean) * (delta - mean));
    const double wf = mu * sgn(delta - pbias);                      // ls.h:285-291
    for (int i = 0; i < 3; i++) mixw[mix_ctx][i] += wf * sgn(pt[i]);
  }
};

// ---------------------------------------------------------------------------------------------------------------
// parameters: profile vector -> Predictor::tparam (ref libsac/libsac.cpp:37-92, libsac/profile.cpp:3-89)
// ---------------------------------------------------------------------------------------------------------------
struct Coef { float vmin, vmax, vdef; };
const Coef kBase[58] = {
    {0.99f, 0.9999f, 0.998f}, {1.0f, 100.0f, 25.0f}, {0.001f, 1.0f, 0.1f}, {0.001f, 1.0f, 0.12f}, {0.001f, 1.0f, 0.06f},
    {0.001f, 1.0f, 0.04f}, {0.98f, 1.0f, 1.0f}, {0.0f, 1.0f, 0.8f}, {0.0f, 1.0f, 0.8f}, {0.0f, 32.0f, 0.0f},
    {0.0005f, 0.05f, 0.005f}, {0.8f, 0.9999f, 0.95f}, {0.99f, 0.9999f, 0.998f}, {1.0f, 100.0f, 25.0f}, {0.001f, 1.0f, 0.1f},
    {0.001f, 1.0f, 0.12f}, {0.001f, 1.0f, 0.06f}, {0.001f, 1.0f, 0.04f}, {0.98f, 1.0f, 1.0f}, {0.0f, 1.0f, 0.8f},
    {0.0f, 1.0f, 0.8f}, {0.0f, 1.0f, 0.8f}, {0.0005f, 0.05f, 0.005f}, {0.8f, 0.9999f, 0.95f}, {4.0f, 32.0f, 16.0f},
    {4.0f, 32.0f, 16.0f}, {0.0f, 32.0f, 8.0f}, {-32.0f, 32.0f, 8.0f}, {256.0f, 8192.0f, 1280.0f}, {32.0f, 4096.0f, 256.0f},
    {4.0f, 2048.0f, 32.0f}, {256.0f, 8192.0f, 1280.0f}, {32.0f, 4096.0f, 256.0f}, {4.0f, 2048.0f, 32.0f}, {0.0f, 1.0f, 0.5f},
    {0.1f, 2.0f, 0.8f}, {0.1f, 10.0f, 2.0f}, {2.0f, 1024.0f, 4.0f}, {2.0f, 1024.0f, 4.0f}, {0.98f, 1.0f, 1.0f},
    {0.98f, 1.0f, 1.0f}, {1.0f, 10.0f, 4.0f}, {0.1f, 10.0f, 5.0f}, {0.001f, 0.005f, 0.0015f}, {0.001f, 0.005f, 0.0015f},
    {4.0f, 10.0f, 5.0f}, {0.98f, 1.0f, 1.0f}, {0.98f, 1.0f, 1.0f}, {0.98f, 1.0f, 1.0f}, {0.98f, 1.0f, 1.0f},
    {0.0f, 1.0f, 0.8f}, {0.0f, 1.0f, 0.8f}, {0.0f, 1.0f, 0.8f}, {0.0f, 1.0f, 0.5f}, {0.1f, 2.0f, 0.8f},
    {0.1f, 10.0f, 2.0f}, {0.0f, 0.5f, 0.1f}, {0.0f, 0.5f, 0.1f}};

struct Param {
  int nA, nB, nM0, nS0, nS1, k, ch_ref, lm_n, bias_scale;
  int vn[2][4];
  double vmu[2][4], vmudecay[2][4], vpowdecay[2][4];
  double lambda[2], ols_nu[2], mu_mix[2], mu_mix_beta[2], beta_sum[2], beta_pow[2], beta_add[2];
  double bias_mu[2], lm_alpha, proj_alpha[2];
};
Param set_param(const float *g, int k)
{
  Param p;
  p.k = k;
  auto G = [&](int i) { return (double)g[i]; };
  auto R = [&](int i) { return (int)std::round(G(i)); };
  const int in0[4] = {28, 29, 30, 37}, in1[4] = {31, 32, 33, 38};
  const int imu0[4] = {2, 3, 4, 5}, imu1[4] = {14, 15, 16, 17};
  const int imd0[4] = {6, 39, 46, 47}, imd1[4] = {18, 40, 48, 49};
  const int ipd0[4] = {7, 8, 50, 51}, ipd1[4] = {19, 20, 21, 52};
  for (int i = 0; i < 4; i++) {
    p.vn[0][i] = R(in0[i]); p.vn[1][i] = R(in1[i]);
    p.vmu[0][i] = G(imu0[i]) / double(p.vn[0][i]); p.vmu[1][i] = G(imu1[i]) / double(p.vn[1][i]);
    p.vmudecay[0][i] = G(imd0[i]); p.vmudecay[1][i] = G(imd1[i]);
    p.vpowdecay[0][i] = G(ipd0[i]); p.vpowdecay[1][i] = G(ipd1[i]);
  }
  p.lambda[0] = G(0); p.ols_nu[0] = G(1); p.mu_mix[0] = G(10); p.mu_mix_beta[0] = G(11);
  p.lambda[1] = G(12); p.ols_nu[1] = G(13); p.mu_mix[1] = G(22); p.mu_mix_beta[1] = G(23);
  p.nA = R(24); p.nB = R(25); p.nS0 = R(26); p.nS1 = R(27); p.nM0 = R(9);
  p.beta_sum[0] = G(34); p.beta_pow[0] = G(35); p.beta_add[0] = G(36);
  p.beta_sum[1] = G(53); p.beta_pow[1] = G(54); p.beta_add[1] = G(55);
  p.proj_alpha[0] = G(56); p.proj_alpha[1] = G(57);
  p.lm_n = R(41); p.lm_alpha = G(42);
  p.bias_mu[0] = G(43); p.bias_mu[1] = G(44);
  p.bias_scale = R(45);
  p.ch_ref = 0;
  if (p.nS1 < 0) { p.nS1 = -p.nS1; p.ch_ref = 1; }
  return p;
}

// ---------------------------------------------------------------------------------------------------------------
// Predictor (ref libsac/pred.cpp:4-45) and the frame loops (ref libsac/libsac.cpp:94-199)
// ---------------------------------------------------------------------------------------------------------------
struct Predictor {
  Param p;
  Ols ols[2];
  Cascade lms[2];
  Bias be[2];
  double p_lpc[2] = {0, 0}, p_lms[2] = {0, 0};
  Predictor(const Param &q, const int32_t *mm)
      : p(q),
        ols{Ols(q.nA + q.nM0, q.k, q.lambda[0], q.ols_nu[0], q.beta_sum[0], q.beta_pow[0], q.beta_add[0]),
            Ols(q.nB + q.nS0 + q.nS1, q.k, q.lambda[1], q.ols_nu[1], q.beta_sum[1], q.beta_pow[1], q.beta_add[1])},
        lms{Cascade(mm[0], mm[1], q.vn[0], q.vmu[0], q.vmudecay[0], q.vpowdecay[0], q.mu_mix[0], q.mu_mix_beta[0], q.lm_n, q.lm_alpha, q.proj_alpha[0]),
            Cascade(mm[2], mm[3], q.vn[1], q.vmu[1], q.vmudecay[1], q.vpowdecay[1], q.mu_mix[1], q.mu_mix_beta[1], q.lm_n, q.lm_alpha, q.proj_alpha[1])},
        be{Bias(q.bias_mu[0], q.bias_scale), Bias(q.bias_mu[1], q.bias_scale)} {}
  void fill0(const int32_t *s0, int idx0, const int32_t *s1, int idx1)          // pred.cpp:17-23
  {
    double *buf = ols[0].x.data(); int bp = 0;
    for (int i = idx0 - p.nA; i < idx0; i++) buf[bp++] = (i >= 0) ? s0[i] : 0.0;
    for (int i = idx1 - p.nM0; i < idx1; i++) buf[bp++] = (i >= 0) ? s1[i] : 0.0;
  }
  void fill1(const int32_t *s0, const int32_t *s1, int idx1, int n)             // pred.cpp:25-31
  {
    double *buf = ols[1].x.data(); int bp = 0;
    for (int i = idx1 - p.nB; i < idx1; i++) buf[bp++] = (i >= 0) ? s1[i] : 0.0;
    for (int i = idx1 - p.nS0; i < idx1 + p.nS1; i++) buf[bp++] = (i >= 0 && i < n) ? s0[i] : 0.0;
  }
  double predict(int ch)
  {
    p_lpc[ch] = ols[ch].Predict();
    p_lms[ch] = lms[ch].Predict();
    return be[ch].Predict(p_lpc[ch] + p_lms[ch]);
  }
  void update(int ch, double val)
  {
    ols[ch].Update(val);
    lms[ch].Update(val - p_lpc[ch]);
    be[ch].Update(val);
  }
  bool bad() const { return lms[0].mix.bad || lms[1].mix.bad; }
};

inline int32_t round_clamp(double pd, int32_t lo, int32_t hi)      // libsac.cpp:106 (NaN / overflow -> lo, as cvttsd2si)
{
  const double r = m_round(pd);
  if (!(r >= (double)lo)) return lo;
  if (r > (double)hi) return hi;
  return (int32_t)r;
}

int predict_frame(int nch, const int32_t *const *planes, int from, int n, const float *prof, int k, const int32_t *mm,
                  int32_t *const *err)
{
  Param param = set_param(prof, k);
  int32_t mm4[4] = {mm[0], mm[1], nch == 2 ? mm[2] : mm[0], nch == 2 ? mm[3] : mm[1]};
  Predictor pr(param, mm4);
  if (nch == 1) {
    const int32_t *src = planes[0] + from;
    for (int idx = 0; idx < n; idx++) {
      pr.fill0(src, idx, src, idx);
      const double pd = pr.predict(0);
      const int32_t pi = round_clamp(pd, mm[0], mm[1]);
      err[0][idx] = src[idx] - pi;
      pr.update(0, src[idx]);
    }
  } else {
    const int ch0 = param.ch_ref, ch1 = 1 - ch0;
    // note (libsac.cpp:98-100,106): Range r0/r1 follow channel index 0/1 while clamps follow the coded channel
    const int32_t *src0 = planes[ch0] + from, *src1 = planes[ch1] + from;
    int idx0 = 0, idx1 = 0;
    while (idx0 < n || idx1 < n) {
      if (idx0 < n) {
        pr.fill0(src0, idx0, src1, idx1);
        const double pd = pr.predict(0);
        const int32_t pi = round_clamp(pd, mm[2 * ch0], mm[2 * ch0 + 1]);
        err[ch0][idx0] = src0[idx0] - pi;
        pr.update(0, src0[idx0]);
        idx0++;
      }
      if (idx0 >= param.nS1) {
        pr.fill1(src0, src1, idx1, n);
        const double pd = pr.predict(1);
        const int32_t pi = round_clamp(pd, mm[2 * ch1], mm[2 * ch1 + 1]);
        err[ch1][idx1] = src1[idx1] - pi;
        pr.update(1, src1[idx1]);
        idx1++;
      }
    }
  }
  return pr.bad() ? 1 : 0;
}

struct Remap;
int32_t remap_unmap(const Remap *m, int32_t pred, int32_t merr);
// maps[ch] != nullptr: the channel's residuals are rank-mapped (libsac.cpp:157-160), mean[ch] re-bases the prediction
int unpredict_frame(int nch, int n, const float *prof, const int32_t *mm, const int32_t *const *err, int32_t *const *dst,
                    const Remap *const *maps = nullptr, const int32_t *mean = nullptr)
{
  auto residual = [&](int ch, int32_t pi, int32_t e) { return (maps && maps[ch]) ? remap_unmap(maps[ch], pi + mean[ch], e) : e; };
  Param param = set_param(prof, 1);
  int32_t mm4[4] = {mm[0], mm[1], nch == 2 ? mm[2] : mm[0], nch == 2 ? mm[3] : mm[1]};
  Predictor pr(param, mm4);
  if (nch == 1) {
    int32_t *d = dst[0];
    for (int idx = 0; idx < n; idx++) {
      pr.fill0(d, idx, d, idx);
      const int32_t pi = round_clamp(pr.predict(0), mm[0], mm[1]);
      d[idx] = pi + residual(0, pi, err[0][idx]);
      pr.update(0, d[idx]);
    }
  } else {
    const int ch0 = param.ch_ref, ch1 = 1 - ch0;
    int32_t *d0 = dst[ch0], *d1 = dst[ch1];
    int idx0 = 0, idx1 = 0;
    while (idx0 < n || idx1 < n) {
      if (idx0 < n) {
        pr.fill0(d0, idx0, d1, idx1);
        const int32_t pi = round_clamp(pr.predict(0), mm[2 * ch0], mm[2 * ch0 + 1]);
        d0[idx0] = pi + residual(ch0, pi, err[ch0][idx0]);
        pr.update(0, d0[idx0]);
        idx0++;
      }
      if (idx0 >= param.nS1) {
        pr.fill1(d0, d1, idx1, n);
        const int32_t pi = round_clamp(pr.predict(1), mm[2 * ch1], mm[2 * ch1 + 1]);
        d1[idx1] = pi + residual(ch1, pi, err[ch1][idx1]);
        pr.update(1, d1[idx1]);
        idx1++;
      }
    }
  }
  return pr.bad() ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------------
// integer helpers (ref common/utils.h:248-273)
// ---------------------------------------------------------------------------------------------------------------
inline int32_t S2U(int32_t v) { if (v < 0) return 2 * (-v); if (v > 0) return 2 * v - 1; return v; }
inline int32_t U2S(int32_t v) { return (v & 1) ? ((v + 1) >> 1) : -(v >> 1); }
inline int iLog2(int v) { int nb = 0; while (v >>= 1) nb++; return nb; }

// ---------------------------------------------------------------------------------------------------------------
// probability model pieces (ref model/model.h, counter.h:17-69, domain.h:7-61, mixer.h:59-101, sse.h:84-125)
// ---------------------------------------------------------------------------------------------------------------
enum { PBITS = 15, PSCALE = 1 << 15, PSCALEh = PSCALE >> 1, PSCALEm = PSCALE - 1 };

struct Tables {
  int fwd[PSCALE];
  int inv[4095];
  int divt[PSCALE];
  int tmin, tmax;
  Tables()
  {
    for (int i = 0; i < PSCALE; i++) {                              // domain.h:17-22
      double f = std::max(i, 1) / (double)PSCALE;
      double q = std::log(f / (1.0 - f)) * 256;
      fwd[i] = (int)std::round(q);
    }
    tmin = fwd[0]; tmax = fwd[PSCALE - 1];
    for (int i = -2047; i <= 2047; i++) {                           // domain.h:27-31
      double q = PSCALE / (1.0 + std::exp(-double(i) / 256.0));
      inv[i + 2047] = (int)std::round(q);
    }
    for (int i = 0; i < PSCALE; i++) divt[i] = PSCALE / (i + 3);    // counter.h:40-51
  }
  int Inv(int x) const { return x < -2047 ? 1 : (x > 2047 ? PSCALEm : inv[x + 2047]); }
};
const Tables &T() { static Tables t; return t; }

inline int idiv_signed(int val, int s) { return val < 0 ? -(((-val) + (1 << (s - 1))) >> s) : (val + (1 << (s - 1))) >> s; }
inline int idiv_signed64(int64_t val, int s) { return (int)(val < 0 ? -(((-val) + (1 << (s - 1))) >> s) : (val + (1 << (s - 1))) >> s); }

struct CounterLimit {                                               // counter.h:53-69
  uint16_t p1 = PSCALEh, counter = 0;
  void update(int bit, int limit)
  {
    if (counter < limit) counter++;
    int dp = bit ? ((PSCALE - p1) * T().divt[counter]) >> PBITS : -((p1 * T().divt[counter]) >> PBITS);
    p1 = (uint16_t)std::clamp(p1 + dp, 1, (int)PSCALEm);
  }
};
struct Counter16 {                                                  // counter.h:17-38, update(bit,L)
  uint16_t p1 = PSCALEh;
  void update(int bit, int L)
  {
    int err = (bit << PBITS) - p1;
    int px = int(p1) + idiv_signed(L * err, PBITS);
    p1 = (uint16_t)std::clamp(px, 1, (int)PSCALEm);
  }
};
struct MixLogistic {                                                // mixer.h:59-101
  int n = 0, w[5] = {0, 0, 0, 0, 0}, x[5] = {0, 0, 0, 0, 0}, pd = 0;
  int Predict(const int *p)
  {
    int64_t sum = 0;
    for (int i = 0; i < n; i++) { x[i] = (int16_t)T().fwd[p[i]]; sum += int64_t(w[i] * x[i]); }
    int s = idiv_signed64(sum, 16);
    pd = (int16_t)std::clamp(T().Inv(s), 1, (int)PSCALEm);
    return pd;
  }
  void Update(int bit, int rate)
  {
    int err = (bit << PBITS) - pd;
    for (int i = 0; i < n; i++) {
      int de = idiv_signed(x[i] * err, 12);
      w[i] = std::clamp(w[i] + idiv_signed(de * rate, 12), -(1 << 19), (1 << 19) - 1);
    }
  }
};
template <int N> struct SsenlT {                                    // sse.h:84-125 (SSENL<N>)
  int tscale, xscale, lb = 0, p_quant = 0;
  Counter16 Map[2][N + 1];
  SsenlT()
  {
    tscale = T().tmax; xscale = (2 * tscale) / (N - 1);
    if (xscale == 0) xscale = 1;
    for (int i = 0; i <= N; i++) { int x = T().Inv(i * xscale - tscale); Map[0][i].p1 = x; Map[1][i].p1 = x; }
  }
  int Predict(int p1)
  {
    int pq = std::min(2 * tscale, std::max(0, T().fwd[p1] + tscale));
    p_quant = pq / xscale;
    int p_mod = pq - p_quant * xscale;
    int pl = Map[lb][p_quant].p1, ph = Map[lb][p_quant + 1].p1;
    int px = (pl * (xscale - p_mod) + ph * p_mod) / xscale;
    return std::clamp(px, 1, (int)PSCALEm);
  }
  void Update(int bit, int rate)
  {
    Map[lb][p_quant].update(bit, rate);
    Map[lb][p_quant + 1].update(bit, rate);
    lb = bit;
  }
};
typedef SsenlT<15> Ssenl;

// ---------------------------------------------------------------------------------------------------------------
// range coder (ref model/range.cpp:54-92)
// ---------------------------------------------------------------------------------------------------------------
struct RangeEnc {
  std::vector<uint8_t> out;
  uint32_t range = 0xFFFFFFFFu, FFNum = 0, Cache = 0;
  uint64_t lowc = 0;
  void ShiftLow()
  {
    uint32_t Carry = uint32_t(lowc >> 32), low = uint32_t(lowc);
    if (low < 0xFF000000U || Carry) {
      out.push_back((uint8_t)(Cache + Carry));
      for (; FFNum != 0; FFNum--) out.push_back((uint8_t)(Carry - 1));
      Cache = low >> 24;
    } else FFNum++;
    lowc = (uint64_t)(uint32_t)(low << 8);
  }
  void Encode(uint32_t p1, int bit)
  {
    const uint32_t rnew = (uint32_t)((uint64_t(range) * ((PSCALE - p1) << (32 - PBITS))) >> 32);
    if (bit) { range -= rnew; lowc += rnew; } else range = rnew;
    while (range < 0x01000000U) { range <<= 8; ShiftLow(); }
  }
  void Stop() { for (int i = 0; i < 5; i++) ShiftLow(); }
};
struct RangeDec {
  const uint8_t *in; int len, pos = 0;
  uint32_t range = 0xFFFFFFFFu, code = 0;
  int get() { return pos < len ? in[pos++] : -1; }                  // bufio.h:18-21
  RangeDec(const uint8_t *p, int n) : in(p), len(n) { for (int i = 0; i < 5; i++) code = (code << 8) + get(); }
  int Decode(uint32_t p1)
  {
    const uint32_t rnew = (uint32_t)((uint64_t(range) * ((PSCALE - p1) << (32 - PBITS))) >> 32);
    int bit = (code >= rnew);
    if (bit) { range -= rnew; code -= rnew; } else range = rnew;
    while (range < 0x01000000U) { range <<= 8; code = (code << 8) + get(); }
    return bit;
  }
};

// ---------------------------------------------------------------------------------------------------------------
// sparse PCM: which of the 2 x 32768 magnitudes occur (Remap, ref libsac/map.h:36-49, map.cpp:104-202) and their
// context-coded transmission (MapEncoder, map.h:10-34, map.cpp:3-102)
// ---------------------------------------------------------------------------------------------------------------
struct Remap {
  static const int scale = 1 << 15;
  int vmin = 0, vmax = 0;
  std::vector<uint8_t> usedl, usedh;
  Remap() : usedl(scale + 1, 0), usedh(scale + 1, 0) {}
  void Analyse(const int32_t *src, int n)                              // map.cpp:126-157 (raw samples, before mean removal)
  {
    for (int i = 0; i < n; i++) {
      int val = src[i];
      if (val > 0) { if (val <= scale) { if (val > vmax) vmax = val; usedh[val] = 1; } }
      else if (val < 0) { val = -val; if (val <= scale) { if (val > vmin) vmin = val; usedl[val] = 1; } }
    }
  }
  bool isUsed(int val) const                                           // map.cpp:159-166
  {
    if (val > scale) return false;
    if (val < -scale) return false;
    if (val > 0) return usedh[val];
    if (val < 0) return usedl[-val];
    return true;
  }
  int32_t Map(int32_t pred, int32_t err) const                         // map.cpp:175-186
  {
    int sg = 1;
    if (err == 0) return 0;
    if (err < 0) { err = -err; sg = -1; }
    int merr = 0;
    for (int i = 1; i <= err; i++) if (isUsed(pred + (sg * i))) merr++;
    return sg * merr;
  }
  int32_t Unmap(int32_t pred, int32_t merr) const                      // map.cpp:188-202
  {
    int sg = 1;
    if (merr == 0) return 0;
    if (merr < 0) { merr = -merr; sg = -1; }
    int err = 1, terr = 0;
    while (1) {
      if (isUsed(pred + (sg * err))) terr++;
      if (terr == merr) break;
      err++;
    }
    return sg * err;
  }
};

int32_t remap_unmap(const Remap *m, int32_t pred, int32_t merr) { return m->Unmap(pred, merr); }

struct MapCoder {
  Counter16 cnt[24], cctx[256];
  Counter16 *pc1 = nullptr, *pc2 = nullptr, *pc3 = nullptr, *pc4 = nullptr, *px = nullptr;
  MixLogistic mixl[4], mixh[4], finalmix, *mix = nullptr;
  SsenlT<32> sse0;                                                      // only sse[0] is ever addressed (map.cpp:83-99)
  std::vector<uint8_t> &ul, &uh;
  MapCoder(std::vector<uint8_t> &usedl, std::vector<uint8_t> &usedh) : ul(usedl), uh(usedh)
  {
    for (auto &m : mixl) m.n = 5;
    for (auto &m : mixh) m.n = 5;
    finalmix.n = 2;
  }
  int PredictLow(int i)                                                // map.cpp:9-29
  {
    const int ctx1 = ul[i - 1], ctx2 = uh[i - 1], ctx3 = i > 1 ? ul[i - 2] : 0;
    pc1 = &cnt[ctx1]; pc2 = &cnt[2 + ctx2]; pc3 = &cnt[4 + (ctx1 << 1) + ctx3]; pc4 = &cnt[8 + (ctx1 << 1) + ctx2];
    int sctx = ul[i - 1];
    if (i > 1) sctx += (ul[i - 2] << 1);
    if (i > 2) sctx += (ul[i - 3] << 2);
    if (i > 3) sctx += (ul[i - 4] << 3);
    px = &cctx[sctx];
    mix = &mixl[ctx1 + (ctx3 << 1)];
    const int p[5] = {pc1->p1, pc2->p1, pc3->p1, pc4->p1, px->p1};
    return mix->Predict(p);
  }
  int PredictHigh(int i)                                               // map.cpp:31-52
  {
    const int ctx1 = uh[i - 1], ctx2 = ul[i], ctx3 = i > 1 ? uh[i - 2] : 0;
    pc1 = &cnt[12 + ctx1]; pc2 = &cnt[12 + 2 + ctx2]; pc3 = &cnt[12 + 4 + (ctx1 << 1) + ctx3]; pc4 = &cnt[12 + 8 + (ctx1 << 1) + ctx2];
    int sctx = uh[i - 1];
    if (i > 1) sctx += (uh[i - 2] << 1);
    if (i > 2) sctx += (uh[i - 3] << 2);
    if (i > 3) sctx += (uh[i - 4] << 3);
    px = &cctx[32 + sctx];
    mix = &mixh[ctx1 + (ctx3 << 1)];
    const int p[5] = {pc1->p1, pc2->p1, pc3->p1, pc4->p1, px->p1};
    return mix->Predict(p);
  }
  void Update(int bit)                                                 // map.cpp:54-62
  {
    pc1->update(bit, 500); pc2->update(bit, 500); pc3->update(bit, 500); pc4->update(bit, 500); px->update(bit, 500);
    mix->Update(bit, 1000);
  }
  int PredictSSE(int p1) { const int vp[2] = {sse0.Predict(p1), p1}; return finalmix.Predict(vp); }
  void UpdateSSE(int bit) { sse0.Update(bit, 300); finalmix.Update(bit, 500); }
  void Encode(RangeEnc &rc)                                            // map.cpp:76-89
  {
    for (int i = 1; i <= 1 << 15; i++) {
      int bit = ul[i];
      rc.Encode(PredictSSE(PredictLow(i)), bit); Update(bit); UpdateSSE(bit);
      bit = uh[i];
      rc.Encode(PredictSSE(PredictHigh(i)), bit); Update(bit); UpdateSSE(bit);
    }
  }
  void Decode(RangeDec &rc)                                            // map.cpp:91-102
  {
    for (int i = 1; i <= 1 << 15; i++) {
      int bit = rc.Decode(PredictSSE(PredictLow(i)));
      Update(bit); ul[i] = bit; UpdateSSE(bit);
      bit = rc.Decode(PredictSSE(PredictHigh(i)));
      Update(bit); uh[i] = bit; UpdateSSE(bit);
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// bitplane coder (ref libsac/vle.cpp:3-261, libsac/vle.h:42-85)
// ---------------------------------------------------------------------------------------------------------------
int predict_laplace(uint32_t avg_sum, int bpn)                      // vle.cpp:70-79
{
  double p_l = 0.0;
  if (avg_sum > 0) {
    double theta = m_exp(-1.0 / avg_sum);
    p_l = 1.0 - 1.0 / (1 + m_pow(theta, (double)(1 << bpn)));
  }
  return std::min(std::max((int)m_round(p_l * PSCALE), 1), (int)PSCALEm);
}

struct Bitplane {
  std::vector<CounterLimit> csig0, csig1, cref0, cref1, cref2, cref3, p_laplace;
  std::vector<MixLogistic> lmixref, lmixsig;
  MixLogistic ssemix;
  std::vector<Ssenl> sse;
  std::vector<int> msb;
  int *pabuf = nullptr;
  int maxbpn, numsamples, bpn = 0, sample = 0, pestimate = 0, sigst[17];
  uint32_t state = 0;
  Ssenl *psse1 = nullptr, *psse2 = nullptr;
  CounterLimit *pl = nullptr, *pc1 = nullptr, *pc2 = nullptr, *pc3 = nullptr, *pc4 = nullptr;
  MixLogistic *plmix = nullptr;

  Bitplane(int maxbpn_, int n) : csig0(1 << 16), csig1(128), cref0(64), cref1(256), cref2(64), cref3(512), p_laplace(32),
                                 lmixref(256), lmixsig(256), sse(160), msb(n), maxbpn(maxbpn_), numsamples(n)
  {
    for (auto &m : lmixref) m.n = 5;
    for (auto &m : lmixsig) m.n = 3;
    ssemix.n = 2;
    for (int i = 0; i < 32; i++) {                                  // vle.cpp:16-22 (theta=.99; 1<<i as int, i=31 wraps)
      int sh = (int)(1u << i);
      int p = std::min(std::max((int)m_round((1.0 - 1.0 / (1 + m_pow(0.99, (double)sh))) * PSCALE), 1), (int)PSCALEm);
      p_laplace[i].p1 = (uint16_t)p;
    }
  }
  static uint32_t bmask(int i) { return ~((1u << i) - 1); }
  uint32_t GetAvgSum(int n)                                         // vle.cpp:54-68
  {
    uint64_t nsum = 0; int nidx = 0;
    for (int k = sample - n; k <= sample + n; k++)
      if (k >= 0 && k < numsamples) {
        int val = pabuf[k];
        val &= k < sample ? bmask(bpn) : bmask(bpn + 1);
        nsum += val; nidx++;
      }
    return nidx > 0 ? (uint32_t)((nsum + (nidx - 1)) / nidx) : 0;
  }
  void GetSigState(int i)                                           // vle.cpp:33-52
  {
    sigst[0] = msb[i];
    for (int d = 1; d <= 8; d++) {
      sigst[2 * d - 1] = i > d - 1 ? msb[i - d] : 0;
      sigst[2 * d] = i < numsamples - d ? msb[i + d] : 0;
    }
  }
  int PredictRef()                                                  // vle.cpp:81-129
  {
    int val = pabuf[sample];
    int lval = sample > 0 ? pabuf[sample - 1] : 0, lval2 = sample > 1 ? pabuf[sample - 2] : 0;
    int nval = sample < (numsamples - 1) ? pabuf[sample + 1] : 0, nval2 = sample < (numsamples - 2) ? pabuf[sample + 2] : 0;
    int b0 = val >> (bpn + 1), b1 = lval >> bpn, b2 = nval >> (bpn + 1), b3 = lval2 >> bpn, b4 = nval2 >> (bpn + 1);
    int c0 = (b0 << 1) < b1, c1 = b0 < b2, c2 = (b0 << 1) < b3, c3 = b0 < b4;
    int x0 = (val >> (bpn + 1)) << 1, x1 = lval >> bpn, x2 = (nval >> (bpn + 1)) << 1, x3 = lval2 >> bpn, x4 = (nval2 >> (bpn + 1)) << 1;
    int xm = (x0 + x1 + x2 + x3 + x4) / 5;
    int d0 = x0 > xm, d1 = x1 > xm;
    int ctx1 = (b0 & 15) + ((b1 & 15) << 4) + ((b2 & 15) << 8);
    int ctx2 = (c0 + (c1 << 1) + (c2 << 2) + (c3 << 3)) + (d0 << 4) + (d1 << 5);
    int ctx3 = sigst[1] + sigst[2] + sigst[3] + sigst[4] + sigst[5] + sigst[6] + sigst[7] + sigst[8];
    pl = &p_laplace[bpn]; pc1 = &cref0[msb[sample]]; pc2 = &cref1[ctx1 & 255]; pc3 = &cref2[ctx2]; pc4 = &cref3[ctx3];
    int pctx = ((((pestimate >> 12) << 1) + d0) << 1) + (b0 & 1);
    plmix = &lmixref[pctx];
    int in[5] = {pestimate, pl->p1, pc1->p1, pc2->p1, pc3->p1};
    return plmix->Predict(in);
  }
  void UpdateRef(int bit)                                           // vle.cpp:131-140
  {
    pl->update(bit, 150); pc1->update(bit, 150); pc2->update(bit, 150); pc3->update(bit, 150); pc4->update(bit, 150);
    plmix->Update(bit, 800);
    state = (state << 1) + 0;
  }
  int PredictSig()                                                  // vle.cpp:143-175
  {
    int ctx1 = 0;
    for (int i = 0; i < 16; i++) if (sigst[i + 1]) ctx1 += 1 << i;
    int n1 = 0, n2 = 0;
    for (int i = 1; i <= 32; i++) {
      if (sample - i >= 0) { if (msb[sample - i]) n1++; if (msb[sample - i] > bpn) n2++; }
      if (sample + i < numsamples - 1) { if (msb[sample + i]) n1++; if (msb[sample + i] > bpn) n2++; }
    }
    pl = &p_laplace[bpn]; pc1 = &csig0[ctx1]; pc2 = &csig1[n2];
    int mixctx = ((state & 15) << 3) + ((n1 >= 3 ? 3 : n1) << 1) + (n2 > 0 ? 1 : 0);
    plmix = &lmixsig[mixctx];
    int in[3] = {pl->p1, pc1->p1, pc2->p1};
    return plmix->Predict(in);
  }
  void UpdateSig(int bit)                                           // vle.cpp:177-186
  {
    pl->update(bit, 150); pc1->update(bit, 300); pc2->update(bit, 300);
    plmix->Update(bit, 700);
    state = (state << 1) + 1;
  }
  int PredictSSE(int p1)                                            // vle.cpp:188-197
  {
    int ctx1 = ((pestimate >> 11) << 1) + (sigst[0] ? 1 : 0);
    int ctx2 = 32 + (sigst[0] ? 1 : 0) + ((sigst[1] ? 1 : 0) << 1) + ((sigst[2] ? 1 : 0) << 2) + ((sigst[3] ? 1 : 0) << 3) +
               ((sigst[4] ? 1 : 0) << 4) + ((sigst[5] ? 1 : 0) << 5) + ((sigst[6] ? 1 : 0) << 6);
    psse1 = &sse[ctx1]; psse2 = &sse[ctx2];
    int pr1 = psse1->Predict(p1);
    int pr2 = psse2->Predict(pr1);
    int in[2] = {(pr1 + pr2 + 1) >> 1, p1};
    return ssemix.Predict(in);
  }
  void UpdateSSE(int bit) { psse1->Update(bit, 250); psse2->Update(bit, 250); ssemix.Update(bit, 250); } // vle.cpp:199-204

  template <class Coder> void Run(Coder &&code, int32_t *abuf, bool decode)   // vle.cpp:206-261
  {
    pabuf = abuf;
    if (decode) for (int i = 0; i < numsamples; i++) abuf[i] = 0;
    for (bpn = maxbpn; bpn >= 0; bpn--) {
      state = 0;
      for (sample = 0; sample < numsamples; sample++) {
        pestimate = predict_laplace(GetAvgSum(32), bpn);
        GetSigState(sample);
        int bit = decode ? 0 : (pabuf[sample] >> bpn) & 1;
        if (sigst[0]) {
          int p = PredictSSE(PredictRef());
          bit = code(p, bit);
          UpdateRef(bit); UpdateSSE(bit);
          if (decode && bit) abuf[sample] += (1 << bpn);
        } else {
          int p = PredictSSE(PredictSig());
          bit = code(p, bit);
          UpdateSig(bit); UpdateSSE(bit);
          if (bit) { if (decode) abuf[sample] += (1 << bpn); msb[sample] = bpn; }
        }
      }
    }
    if (decode) for (int i = 0; i < numsamples; i++) abuf[i] = U2S(abuf[i]);
  }
};

void bitplane_encode_into(RangeEnc &rc, const int32_t *ubuf, int n, int maxbpn)
{
  std::vector<int32_t> tmp(ubuf, ubuf + n);
  Bitplane bc(maxbpn, n);
  bc.Run([&](int p, int bit) { rc.Encode(p, bit); return bit; }, tmp.data(), false);
}
std::vector<uint8_t> bitplane_encode(const int32_t *ubuf, int n, int maxbpn)
{
  RangeEnc rc;
  bitplane_encode_into(rc, ubuf, n, maxbpn);
  rc.Stop();
  return rc.out;
}
// EncodeMonoFrame_Mapped (libsac.cpp:214-228): the used-value map, then the rank-mapped residuals, ONE range coder
std::vector<uint8_t> mapped_encode(Remap &map, const int32_t *ubuf_map, int n, int maxbpn_map)
{
  RangeEnc rc;
  MapCoder me(map.usedl, map.usedh);
  me.Encode(rc);
  bitplane_encode_into(rc, ubuf_map, n, maxbpn_map);
  rc.Stop();
  return rc.out;
}

// ---------------------------------------------------------------------------------------------------------------
// cost functions (ref libsac/cost.h:15-176)
// ---------------------------------------------------------------------------------------------------------------
double cost_calc(int kind, const int32_t *buf, int n)
{
  if (n == 0) return 0.0;
  switch (kind) {
  case SACO_COST_L1: { int64_t s = 0; for (int i = 0; i < n; i++) s += (int64_t)std::fabs((double)buf[i]); return s / (double)n; }
  case SACO_COST_RMS: { int64_t s = 0; for (int i = 0; i < n; i++) s += buf[i] * buf[i]; return std::sqrt(s / (double)n); }
  case SACO_COST_GOLOMB: {                                          // cost.h:43-66 (alpha=.97, RunWeight)
    double rm = 0.0; int64_t nbits = 0;
    for (int i = 0; i < n; i++) {
      const int32_t m = std::max((int32_t)rm, 1);
      const int32_t u = S2U(buf[i]);
      nbits += u / m + 1;
      if (m > 1) nbits += 32 - __builtin_clz((unsigned)m);
      rm = 0.97 * rm + u;
    }
    return nbits / 8.;
  }
  case SACO_COST_ENTROPY: {                                         // cost.h:70-116
    int32_t mn = std::numeric_limits<int32_t>::max(), mx = std::numeric_limits<int32_t>::min();
    for (int i = 0; i < n; i++) { mx = std::max(mx, buf[i]); mn = std::min(mn, buf[i]); }
    std::vector<int> counts((size_t)(mx - mn) + 1);
    for (int i = 0; i < n; i++) ++counts[buf[i] - mn];
    const double invs = 1.0 / (double)n;
    double ent = 0.0;
    if (counts.size() < (size_t)n) {
      for (int c : counts) { if (c == 0) continue; ent += c * std::log2(c * invs); }
    } else {
      for (int i = 0; i < n; i++) ent += std::log2(counts[buf[i] - mn] * invs);
    }
    return -ent / 8.0;
  }
  case SACO_COST_BITPLANE: {                                        // cost.h:144-175
    std::vector<int32_t> u(n); int vmax = 0;
    for (int i = 0; i < n; i++) { u[i] = S2U(buf[i]); vmax = std::max(vmax, u[i]); }
    return (double)bitplane_encode(u.data(), n, iLog2(vmax)).size();
  }
  }
  return -1.0;
}

// ---------------------------------------------------------------------------------------------------------------
// DDS (ref opt/dds.cpp:12-119, opt/opt.cpp:5,111-116,156-166, opt/ssc.h:6-60, common/rand.h)
// ---------------------------------------------------------------------------------------------------------------
struct Dds {
  std::mt19937 eng{0};
  int ndim, nfunc_max;
  const double *xmin, *xmax;
  double r01() { return std::uniform_real_distribution<double>{0, 1}(eng); }
  double rnorm() { return std::normal_distribution<double>{0.0, 1.0}(eng); }
  double reflect(double xnew, double lo, double hi)
  {
    if (xnew < lo) { xnew = lo + (lo - xnew); if (xnew > hi) xnew = lo; }
    if (xnew > hi) { xnew = hi - (xnew - hi); if (xnew < lo) xnew = hi; }
    return xnew;
  }
  vec candidate(const vec &x, int nfunc, double sigma)
  {
    std::vector<int> J;
    double p = 1.0 - std::log(nfunc) / std::log(nfunc_max);
    for (int i = 0; i < ndim; i++) if (r01() < p) J.push_back(i);
    if (!J.size()) J.push_back((int)std::uniform_int_distribution<uint32_t>{0u, (uint32_t)(ndim - 1)}(eng));
    vec xt = x;
    for (int k : J) {
      double s = sigma * (xmax[k] - xmin[k]);
      xt[k] = reflect(x[k] + s * rnorm(), xmin[k], xmax[k]);
    }
    return xt;
  }
};

double dds_run(int ndim, const double *xmin, const double *xmax, const double *xstart, int nfunc_max, int num_threads,
               double sigma_init, saco_cost_cb cb, void *user, double *xbest)
{
  Dds d; d.ndim = ndim; d.nfunc_max = nfunc_max; d.xmin = xmin; d.xmax = xmax;
  vec xb(xstart, xstart + ndim);
  double fb = cb(xb.data(), ndim, user);
  double sigma = sigma_init;
  int nfunc = 1;
  if (num_threads <= 0) {                                           // run_single, SSC0(3,50)
    int nsucc = 0, nfail = 0;
    while (nfunc < nfunc_max) {
      vec xg = d.candidate(xb, nfunc, sigma);
      double fg = cb(xg.data(), ndim, user);
      nfunc++;
      bool ok = fg < fb;
      if (ok) { xb = xg; fb = fg; nsucc++; nfail = 0; } else { nsucc = 0; nfail++; }
      if (nsucc >= 3) { sigma *= 2.0; nsucc = 0; } else if (nfail >= 50) { sigma /= 2.0; nfail = 0; }
      sigma = std::clamp(sigma, 0.05, 0.5);
    }
  } else {                                                          // run_mt, SSC1(0.05,0.10,0.05)
    double p_succ = 0.05;
    while (nfunc < nfunc_max) {
      const int nt = std::min(nfunc_max - nfunc, num_threads);
      std::vector<vec> xg(nt); vec fg(nt);
      for (int i = 0; i < nt; i++) { xg[i] = d.candidate(xb, nfunc, sigma); nfunc++; }
      for (int i = 0; i < nt; i++) fg[i] = cb(xg[i].data(), ndim, user);
      const double fb_old = fb; int nsucc = 0;
      for (int i = 0; i < nt; i++) if (fg[i] < fb_old) { nsucc++; if (fg[i] < fb) { fb = fg[i]; xb = xg[i]; } }
      const double lambda = nsucc / (double)nt;
      p_succ = (1.0 - 0.10) * p_succ + 0.10 * lambda;
      sigma = sigma * std::exp(0.05 * (p_succ - 0.05) / (1.0 - 0.05));
      sigma = std::clamp(sigma, 0.05, 0.25);
    }
  }
  std::copy(xb.begin(), xb.end(), xbest);
  return fb;
}

inline void put32(uint8_t *b, uint32_t v) { b[0] = v & 0xff; b[1] = (v >> 8) & 0xff; b[2] = (v >> 16) & 0xff; b[3] = (v >> 24) & 0xff; }
inline uint32_t get32(const uint8_t *b) { return b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24); }

} // namespace

// =================================================================================================================
extern "C" {

void saco_set_modes(int order, int math) { g_order = order; g_math = math; }

int saco_base_profile(float *vmin, float *vmax, float *vdef)
{
  for (int i = 0; i < 58; i++) { vmin[i] = kBase[i].vmin; vmax[i] = kBase[i].vmax; vdef[i] = kBase[i].vdef; }
  return 58;
}

int saco_predict_frame(int nch, const int32_t *s0, const int32_t *s1, int from, int n, const float *profile58, int k,
                       const int32_t *minmax, int32_t *e0, int32_t *e1)
{
  const int32_t *planes[2] = {s0, s1};
  int32_t *err[2] = {e0, e1};
  return predict_frame(nch, planes, from, n, profile58, k, minmax, err);
}
int saco_unpredict_frame(int nch, int n, const float *profile58, const int32_t *minmax, const int32_t *e0,
                         const int32_t *e1, int32_t *s0, int32_t *s1)
{
  const int32_t *err[2] = {e0, e1};
  int32_t *dst[2] = {s0, s1};
  return unpredict_frame(nch, n, profile58, minmax, err, dst);
}

double saco_cost(int kind, const int32_t *buf, int n) { return cost_calc(kind, buf, n); }

int saco_bitplane_encode(const int32_t *ubuf, int n, int maxbpn, uint8_t *out, int cap)
{
  std::vector<uint8_t> b = bitplane_encode(ubuf, n, maxbpn);
  if (out && (int)b.size() <= cap) std::copy(b.begin(), b.end(), out);
  return (int)b.size();
}
void saco_bitplane_decode(const uint8_t *in, int nbytes, int n, int maxbpn, int32_t *out)
{
  RangeDec rd(in, nbytes);
  Bitplane bc(maxbpn, n);
  bc.Run([&](int p, int) { return rd.Decode(p); }, out, true);
}
int saco_predict_laplace(uint32_t avg_sum, int bpn) { return predict_laplace(avg_sum, bpn); }
void saco_logdomain_tables(int *fwd, int *inv)
{
  std::copy_n(T().fwd, PSCALE, fwd);
  std::copy_n(T().inv, 4095, inv);
}
int saco_range_encode(const uint16_t *p1s, const uint8_t *bits, int n, uint8_t *out, int cap)
{
  RangeEnc rc;
  for (int i = 0; i < n; i++) rc.Encode(p1s[i], bits[i]);
  rc.Stop();
  if (out && (int)rc.out.size() <= cap) std::copy(rc.out.begin(), rc.out.end(), out);
  return (int)rc.out.size();
}

double saco_dds_run(int ndim, const double *xmin, const double *xmax, const double *xstart, int nfunc_max,
                    int num_threads, double sigma_init, saco_cost_cb cb, void *user, double *xbest)
{
  return dds_run(ndim, xmin, xmax, xstart, nfunc_max, num_threads, sigma_init, cb, user, xbest);
}

// ref libsac/libsac.cpp:443-479 (Predict), :365-427 (Optimize), :429-441 (S2U), :201-278 (Encode incl. the sparse-pcm
// choice: rank-mapped residuals when the L1 ratio exceeds 1.05 AND the mapped record is smaller), :507-578 (WriteEncoded)
// mapped_out[ch] (optional) receives 1 where the channel was coded rank-mapped.
int saco_encode_frame2(int nch, int n, const int32_t *s0_in, const int32_t *s1_in, float *profile_io, const int *cfg,
                       int sparse_pcm, uint8_t *out, int cap, int *mapped_out)
{
  const int optimize = cfg[0]; const double fraction = cfg[1] / 1e6; const int maxnfunc = cfg[2], nthreads = cfg[3];
  const double sigma = cfg[4] / 1e6; const int optk = cfg[5], cost_kind = cfg[6], framesize = cfg[7];
  std::vector<int32_t> s[2];
  s[0].assign(s0_in, s0_in + n);
  if (nch == 2) s[1].assign(s1_in, s1_in + n);
  int32_t mean[2] = {0, 0}, mm[4] = {0, 0, 0, 0};
  Remap maps[2];
  for (int ch = 0; ch < nch; ch++) {                                // libsac.cpp:626-651, 448-458
    if (sparse_pcm) maps[ch].Analyse(s[ch].data(), n);              // on the raw samples, before the mean is removed
    int64_t sum = 0;
    int32_t mn = std::numeric_limits<int32_t>::max(), mx = std::numeric_limits<int32_t>::min();
    for (int i = 0; i < n; i++) { sum += s[ch][i]; mx = std::max(mx, s[ch][i]); mn = std::min(mn, s[ch][i]); }
    mean[ch] = n ? (int)std::floor(sum / (double)n) : 0;
    if (mean[ch] != 0) { for (int i = 0; i < n; i++) s[ch][i] -= mean[ch]; mn -= mean[ch]; mx -= mean[ch]; }
    mm[2 * ch] = mn; mm[2 * ch + 1] = mx;
  }
  const int32_t *planes[2] = {s[0].data(), nch == 2 ? s[1].data() : nullptr};
  if (optimize) {
    const int nopt = std::min(n, (int)std::ceil(framesize * fraction));
    const int start = (n - nopt) / 2;
    std::vector<int> idx;
    for (int i = 0; i < 58; i++) if (i != 56 && i != 57) idx.push_back(i);
    const int nd = (int)idx.size();
    vec xmin(nd), xmax(nd), xs(nd), xb(nd);
    for (int i = 0; i < nd; i++) { xmin[i] = kBase[idx[i]].vmin; xmax[i] = kBase[idx[i]].vmax; xs[i] = profile_io[idx[i]]; }
    struct Ctx { int nch, start, nopt, optk, cost_kind; const int32_t *const *planes; const int32_t *mm; const float *base; const std::vector<int> *idx; } ctx
        {nch, start, nopt, optk, cost_kind, planes, mm, profile_io, &idx};
    auto cb = [](const double *x, int nd_, void *u) -> double {
      Ctx &c = *static_cast<Ctx *>(u);
      float prof[58];
      std::copy_n(c.base, 58, prof);
      for (int i = 0; i < nd_; i++) prof[(*c.idx)[i]] = (float)x[i];   // vdef is a float (profile.h:70-72)
      std::vector<int32_t> e0(c.nopt), e1(c.nopt);
      int32_t *err[2] = {e0.data(), e1.data()};
      predict_frame(c.nch, c.planes, c.start, c.nopt, prof, c.optk, c.mm, err);
      double cost = 0.0;
      for (int ch = 0; ch < c.nch; ch++) cost += cost_calc(c.cost_kind, err[ch], c.nopt);
      return cost;
    };
    dds_run(nd, xmin.data(), xmax.data(), xs.data(), maxnfunc, nthreads, sigma, cb, &ctx, xb.data());
    for (int i = 0; i < nd; i++) profile_io[idx[i]] = (float)xb[i];
  }
  std::vector<int32_t> e[2] = {std::vector<int32_t>(n), std::vector<int32_t>(nch == 2 ? n : 0)};
  int32_t *err[2] = {e[0].data(), e[1].data()};
  predict_frame(nch, planes, 0, n, profile_io, 1, mm, err);
  std::vector<uint8_t> rec(4 + 58 * 4);
  put32(rec.data(), (uint32_t)n);
  for (int i = 0; i < 58; i++) { uint32_t ix; std::memcpy(&ix, &profile_io[i], 4); put32(&rec[4 + 4 * i], ix); }
  for (int ch = 0; ch < nch; ch++) {
    std::vector<int32_t> u(n); int32_t emax = 0;
    for (int i = 0; i < n; i++) { u[i] = S2U(e[ch][i]); emax = std::max(emax, u[i]); }
    const int maxbpn = iLog2(emax);
    std::vector<uint8_t> payload = bitplane_encode(u.data(), n, maxbpn);
    int flag = maxbpn;
    if (mapped_out) mapped_out[ch] = 0;
    if (sparse_pcm) {                                               // CalcRemapError, libsac.cpp:230-251, 259-276
      std::vector<int32_t> em(n), um(n); int32_t emax_map = 0;
      for (int i = 0; i < n; i++) {
        const int32_t pred = (s[ch][i] - e[ch][i]) + mean[ch];      // pred[ch][i] = pi + mean (libsac.cpp:107)
        em[i] = maps[ch].Map(pred, e[ch][i]); um[i] = S2U(em[i]); emax_map = std::max(emax_map, um[i]);
      }
      const int maxbpn_map = iLog2(emax_map);
      const double ent1 = cost_calc(SACO_COST_L1, e[ch].data(), n), ent2 = cost_calc(SACO_COST_L1, em.data(), n);
      const double r = ent2 != 0.0 ? ent1 / ent2 : 1.0;
      if (r > 1.05) {
        std::vector<uint8_t> pm = mapped_encode(maps[ch], um.data(), n, maxbpn_map);
        if (pm.size() < payload.size()) { payload.swap(pm); flag = (1 << 9) | maxbpn_map; if (mapped_out) mapped_out[ch] = 1; }
      }
    }
    uint8_t hdr[18];
    put32(hdr, (uint32_t)payload.size()); put32(hdr + 4, (uint32_t)mean[ch]); put32(hdr + 8, (uint32_t)mm[2 * ch]);
    put32(hdr + 12, (uint32_t)mm[2 * ch + 1]); hdr[16] = flag & 0xff; hdr[17] = (flag >> 8) & 0xff;
    rec.insert(rec.end(), hdr, hdr + 18);
    rec.insert(rec.end(), payload.begin(), payload.end());
  }
  if (out && (int)rec.size() <= cap) std::copy(rec.begin(), rec.end(), out);
  return (int)rec.size();
}

int saco_encode_frame(int nch, int n, const int32_t *s0_in, const int32_t *s1_in, float *profile_io, const int *cfg,
                      uint8_t *out, int cap)
{
  return saco_encode_frame2(nch, n, s0_in, s1_in, profile_io, cfg, 1 /* FrameCoder::tsac_cfg default, libsac.h:34 */, out, cap, nullptr);
}

// used-value map alone: Remap::Analyse + MapEncoder::Encode + Stop (map.cpp:76-89, 126-157) -> bytes; and back
int saco_map_encode(const int32_t *raw, int n, uint8_t *out, int cap)
{
  Remap m; m.Analyse(raw, n);
  RangeEnc rc; MapCoder me(m.usedl, m.usedh); me.Encode(rc); rc.Stop();
  if (out && (int)rc.out.size() <= cap) std::copy(rc.out.begin(), rc.out.end(), out);
  return (int)rc.out.size();
}
void saco_map_decode(const uint8_t *in, int nbytes, uint8_t *usedl /*[32769]*/, uint8_t *usedh /*[32769]*/)
{
  Remap m; RangeDec rd(in, nbytes); MapCoder me(m.usedl, m.usedh); me.Decode(rd);
  std::copy(m.usedl.begin(), m.usedl.end(), usedl); std::copy(m.usedh.begin(), m.usedh.end(), usedh);
}
// Remap::Map / Unmap element-wise against the map of `raw` (map.cpp:175-202)
void saco_remap(const int32_t *raw, int nraw, const int32_t *pred, const int32_t *err, int n, int unmap, int32_t *out)
{
  Remap m; m.Analyse(raw, nraw);
  for (int i = 0; i < n; i++) out[i] = unmap ? m.Unmap(pred[i], err[i]) : m.Map(pred[i], err[i]);
}

// ref libsac/libsac.cpp:580-593 (ReadEncoded), :280-298 (DecodeMonoFrame), :144-199 (UnpredictFrame)
int saco_decode_frame(int nch, const uint8_t *in, int len, int32_t *s0, int32_t *s1, int *n_out)
{
  (void)len;
  int pos = 0;
  const int n = (int)get32(in); pos += 4;
  float prof[58];
  for (int i = 0; i < 58; i++) { uint32_t ix = get32(in + pos); std::memcpy(&prof[i], &ix, 4); pos += 4; }
  std::vector<int32_t> e[2];
  int32_t mean[2] = {0, 0}, mm[4] = {0, 0, 0, 0};
  Remap maps[2]; bool mapped[2] = {false, false};
  for (int ch = 0; ch < nch; ch++) {
    const int blocksize = (int)get32(in + pos);
    mean[ch] = (int32_t)get32(in + pos + 4); mm[2 * ch] = (int32_t)get32(in + pos + 8); mm[2 * ch + 1] = (int32_t)get32(in + pos + 12);
    const int flag = in[pos + 16] | (in[pos + 17] << 8);             // ReadBlockHeader, libsac.cpp:551-564
    const int maxbpn = flag & 0xff;
    mapped[ch] = (flag >> 9) != 0;
    pos += 18;
    e[ch].resize(n);
    RangeDec rd(in + pos, blocksize);
    if (mapped[ch]) { MapCoder me(maps[ch].usedl, maps[ch].usedh); me.Decode(rd); }   // DecodeMonoFrame, libsac.cpp:280-298
    Bitplane bc(maxbpn, n);
    bc.Run([&](int p, int) { return rd.Decode(p); }, e[ch].data(), true);
    pos += blocksize;
  }
  const int32_t *err[2] = {e[0].data(), e[1].data()};
  int32_t *dst[2] = {s0, s1};
  const Remap *mp[2] = {mapped[0] ? &maps[0] : nullptr, mapped[1] ? &maps[1] : nullptr};
  unpredict_frame(nch, n, prof, mm, err, dst, mp, mean);
  for (int ch = 0; ch < nch; ch++)
    if (mean[ch] != 0) for (int i = 0; i < n; i++) dst[ch][i] += mean[ch];
  *n_out = n;
  return pos;
}

} // extern "C"
