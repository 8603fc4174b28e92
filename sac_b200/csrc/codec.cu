// codec.cu -- frame and file level of the codec on top of the engine:
//   FrameCoder::Predict / Optimize / Encode / WriteEncoded   (/root/reference src/libsac/libsac.cpp:365-578)
//   FrameCoder::ReadEncoded / Decode / Unpredict             (src/libsac/libsac.cpp:144-199,280-298,580-593)
//   Codec::EncodeFile / DecodeFile                            (src/libsac/libsac.cpp:782-883)
//   OptDDS, SSC0/SSC1                                         (src/opt/dds.cpp, src/opt/ssc.h, src/opt/opt.cpp:111-166)
//   Wav / Chunks / Sac containers, MD5                        (src/file/wav.cpp, src/file/sac.cpp, src/common/md5.cpp)
// The search loop and the containers are host code; every sample-touching step runs in the kernels.
#include "codec.h"
#include "engine.h"
#include <algorithm>
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <array>
#include <atomic>
#include <memory>
#include <mutex>
#include <cstdlib>
#include <string>
#include <thread>

namespace sacb {

// =====================================================================================================================
// DDS
// =====================================================================================================================
DdsSearch::DdsSearch(int D, const double *xmin, const double *xmax, const double *xstart, int nfunc_max, int num_threads,
                     double sigma_init, int spec)
    : D_(D), nfunc_max_(nfunc_max), num_threads_(num_threads), xmin_(xmin, xmin + D), xmax_(xmax, xmax + D), xb_(xstart, xstart + D),
      sigma_(sigma_init), spec_(num_threads <= 0 ? std::max(spec, 1) : 1)
{
}
double DdsSearch::reflect(double xnew, double lo, double hi) const      // opt.cpp:156-166
{
  if (xnew < lo) { xnew = lo + (lo - xnew); if (xnew > hi) xnew = lo; }
  if (xnew > hi) { xnew = hi - (xnew - hi); if (xnew < lo) xnew = hi; }
  return xnew;
}
std::vector<double> DdsSearch::candidate(int nfunc)                     // dds.cpp:12-31
{
  std::vector<int> J;
  const double p = 1.0 - std::log(nfunc) / std::log(nfunc_max_);
  for (int i = 0; i < D_; i++)
    if (std::uniform_real_distribution<double>{0, 1}(eng_) < p) J.push_back(i);
  if (J.empty()) J.push_back((int)std::uniform_int_distribution<uint32_t>{0u, (uint32_t)(D_ - 1)}(eng_));
  std::vector<double> xt = xb_;
  for (int k : J) {
    const double s = sigma_ * (xmax_[k] - xmin_[k]);                    // opt.cpp:111-116
    xt[k] = reflect(xb_[k] + s * std::normal_distribution<double>{0.0, 1.0}(eng_), xmin_[k], xmax_[k]);
  }
  return xt;
}
void DdsSearch::ssc0_step(bool ok)                                      // SSC0(3,50)::update (ssc.h:6-37)
{
  if (ok) { nsucc_ += 1; nfail_ = 0; } else { nsucc_ = 0; nfail_ += 1; }
  if (nsucc_ >= 3) { sigma_ = sigma_ * 2.0; nsucc_ = 0; } else if (nfail_ >= 50) { sigma_ = sigma_ / 2.0; nfail_ = 0; }
  sigma_ = std::clamp(sigma_, 0.05, 0.5);
}
// Sequential search (run_single, dds.cpp:33-60) in speculative batches. A step of run_single draws its candidate from
// (xb, sigma, nfunc, generator) and only then looks at the cost; most steps fail (the incumbent stays). propose() therefore
// draws the next `width` candidates as run_single would if every one of them failed -- same generator stream, same SSC0
// bookkeeping -- and consume() walks the costs in order: up to the first success they ARE run_single's steps; the
// success is applied, the generator is rewound to where it stood after that candidate was drawn, and the rest of the batch
// (drawn around the old incumbent) is dropped. The accepted sequence, every sigma and the result are those of run_single
// evaluated one candidate at a time; only the number of evaluations spent is larger.
void DdsSearch::propose(std::vector<std::vector<double>> &cands)
{
  cands.clear();
  pending_.clear();
  spec_pts_.clear();
  if (num_threads_ <= 0) {
    if (!started_) cands.push_back(xb_);
    int width = 1;
    if (spec_ > 1) {
      // expected useful steps of a batch of w: (1 - (1-q)^w) / q; a width of about 1/q keeps two thirds of the work useful
      width = std::clamp((int)std::lround(1.0 / std::max(q_est_, 1e-3)), 2, spec_);
      if (!started_) width = std::max(width - 1, 1);
    } else if (!started_) width = 0;
    const double sigma0 = sigma_; const int ns0 = nsucc_, nf0 = nfail_, nfunc0 = nfunc_;
    const std::mt19937 rng0 = eng_;
    int nf = started_ ? nfunc_ : 1;
    for (int i = 0; i < width && nf < nfunc_max_; i++) {
      SpecPoint sp;
      sp.sigma = sigma_; sp.nsucc = nsucc_; sp.nfail = nfail_; sp.nfunc = nf;
      pending_.push_back(candidate(nf));
      sp.rng_after = eng_;
      spec_pts_.push_back(sp);
      ssc0_step(false);                                                 // assume it fails
      nf++;
    }
    // the speculated bookkeeping is re-applied by consume(); restore what propose() may not change
    sigma_ = sigma0; nsucc_ = ns0; nfail_ = nf0; nfunc_ = nfunc0;
    if (pending_.empty()) eng_ = rng0;
    cands.insert(cands.end(), pending_.begin(), pending_.end());
    evaluated_ += (long long)cands.size();
    return;
  }
  if (!started_) {
    cands.push_back(xb_);
    // run_mt draws its first generation from the start vector, sigma_init and the evaluation counter alone
    // (dds.cpp:66-83): none of it depends on f(xstart), which the selection only needs afterwards (dds.cpp:88-98).
    // The start vector and the first generation therefore go out as ONE batch -- same candidates, same random draws,
    // same selection, one launch set less per frame.
    nfunc_ = 1;
    const int nt = std::min(nfunc_max_ - nfunc_, num_threads_);
    for (int i = 0; i < nt; i++) { pending_.push_back(candidate(nfunc_)); nfunc_++; }
    cands.insert(cands.end(), pending_.begin(), pending_.end());
    evaluated_ += (long long)cands.size();
    return;
  }
  const int nt = std::min(nfunc_max_ - nfunc_, num_threads_);
  for (int i = 0; i < nt; i++) { cands.push_back(candidate(nfunc_)); nfunc_++; }
  pending_ = cands;
  evaluated_ += (long long)cands.size();
}
void DdsSearch::consume(const double *costs)
{
  if (num_threads_ <= 0) {                                              // run_single + SSC0
    if (!started_) { fb_ = costs[0]; started_ = true; nfunc_ = 1; costs += 1; }
    for (size_t i = 0; i < pending_.size(); i++) {
      const SpecPoint &sp = spec_pts_[i];
      sigma_ = sp.sigma; nsucc_ = sp.nsucc; nfail_ = sp.nfail;        // equal to the running state: kept for clarity
      const bool ok = costs[i] < fb_;
      if (trace_on_) trace_.emplace_back(costs[i], pending_[i]);
      if (ok) { xb_ = pending_[i]; fb_ = costs[i]; }
      ssc0_step(ok);
      nfunc_ = sp.nfunc + 1;
      q_est_ = std::clamp(0.97 * q_est_ + 0.03 * (ok ? 1.0 : 0.0), 0.03, 0.5);
      if (ok) { eng_ = sp.rng_after; break; }                           // later candidates were drawn around the old incumbent
    }
    return;
  }
  if (!started_) {
    fb_ = costs[0]; started_ = true;
    costs += 1;                                                        // the first generation travelled with the start vector
    if (pending_.empty()) return;
  }
  const int nt = (int)pending_.size();
  {                                                                     // run_mt + SSC1 (dds.cpp:88-98, ssc.h:40-60)
    const double fb_old = fb_;
    int nsucc = 0;
    for (int i = 0; i < nt; i++) {
      if (trace_on_) trace_.emplace_back(costs[i], pending_[i]);
      if (costs[i] < fb_old) { nsucc++; if (costs[i] < fb_) { fb_ = costs[i]; xb_ = pending_[i]; } }
    }
    const double lambda = nsucc / static_cast<double>(nt);
    p_succ_ = (1.0 - 0.10) * p_succ_ + 0.10 * lambda;
    sigma_ = sigma_ * std::exp(0.05 * (p_succ_ - 0.05) / (1.0 - 0.05));
    sigma_ = std::clamp(sigma_, 0.05, 0.25);
  }
}

// =====================================================================================================================
// DE (src/opt/de.cpp). Random draws are made in the reference's order with freshly constructed libstdc++
// distributions (common/rand.h), so that the trial vectors equal the reference's for the same costs.
// =====================================================================================================================
DeSearch::DeSearch(int D, const double *xmin, const double *xmax, const double *xstart, int nfunc_max, double sigma_init)
    : D_(D), nfunc_max_(nfunc_max), xmin_(xmin, xmin + D), xmax_(xmax, xmax + D), sigma_init_(sigma_init)
{
  xb_.first = 0.0;
  xb_.second.assign(xstart, xstart + D);
}
double DeSearch::reflect(double xnew, double lo, double hi) const      // opt.cpp:156-166
{
  if (xnew < lo) { xnew = lo + (lo - xnew); if (xnew > hi) xnew = lo; }
  if (xnew > hi) { xnew = hi - (xnew - hi); if (xnew < lo) xnew = hi; }
  return xnew;
}
std::vector<int> DeSearch::select_k_unique_except(int n, int ie, int k)    // de.cpp:13-31
{
  std::vector<int> r;
  if (k >= n - 1) return r;
  std::vector<int> e(n);
  for (int i = 0; i < n; i++) e[i] = i;
  e.erase(std::remove(e.begin(), e.end(), ie), e.end());
  for (int i = 0; i < k; i++) {
    const int idx = (int)std::uniform_int_distribution<uint32_t>{0u, (uint32_t)(e.size() - 1)}(eng_);
    const int val = e[idx];
    r.push_back(val);
    e.erase(std::remove(e.begin(), e.end(), val), e.end());
  }
  return r;
}
void DeSearch::propose(std::vector<std::vector<double>> &cands)
{
  cands.clear();
  if (phase_ == 0) { cands.push_back(xb_.second); return; }
  if (phase_ == 1) {                                                    // de.cpp:91-104: NP-1 normal samples around the start
    pop_.assign(kNP, Point());
    pop_[0] = xb_;
    for (int a = 1; a < kNP; a++) {
      std::vector<double> xt(D_);
      for (int i = 0; i < D_; i++) {                                    // Opt::gen_norm (opt.cpp:111-116)
        const double s = sigma_init_ * (xmax_[i] - xmin_[i]);
        const double xnew = xb_.second[i] + s * std::normal_distribution<double>{0.0, 1.0}(eng_);
        xt[i] = reflect(xnew, xmin_[i], xmax_[i]);
      }
      pop_[a].second = xt;
      cands.push_back(xt);
    }
    return;
  }
  // ---- one generation (de.cpp:122-143) ----
  std::sort(pop_.begin(), pop_.end(), [](const Point &a, const Point &b) { return a.first < b.first; });   // CURPBEST
  const int num_agents = std::min(nfunc_max_ - nfunc_, (int)pop_.size());
  gen_.assign(num_agents, Point());
  gen_mut_.assign(num_agents, std::pair<double, double>());
  for (int ia = 0; ia < num_agents; ia++) {                             // generate_candidate (de.cpp:33-71)
    const double tCR = std::clamp(std::normal_distribution<double>{mCR_, 0.1}(eng_), 0.01, 1.0);
    const double tF = std::clamp(std::cauchy_distribution<double>{mF_, 0.1}(eng_), 0.01, 1.0);
    const int R = (int)std::uniform_int_distribution<uint32_t>{0u, (uint32_t)(D_ - 1)}(eng_);
    const std::vector<int> v = select_k_unique_except((int)pop_.size(), ia, 2);
    const int np = std::min(kNpBest, (int)pop_.size() - 1);
    const int xp = np > 0 ? (int)std::uniform_int_distribution<uint32_t>{0u, (uint32_t)np}(eng_) : 0;
    const std::vector<double> &xbest = pop_[xp].second, &xcur = pop_[ia].second, &x1 = pop_[v[0]].second, &x2 = pop_[v[1]].second;
    std::vector<double> xm(D_), xtrial(D_);
    for (int i = 0; i < D_; i++) {                                      // mut_curbest (de.cpp:175-183)
      const double y = xcur[i] + tF * (xbest[i] - xcur[i]) + tF * (x1[i] - x2[i]);
      xm[i] = reflect(y, xmin_[i], xmax_[i]);
    }
    for (int i = 0; i < D_; i++) {                                      // binomial cross-over: the event is drawn for every i
      const bool ev = std::uniform_real_distribution<double>{0, 1}(eng_) < tCR;
      xtrial[i] = (ev || i == R) ? xm[i] : xcur[i];
    }
    gen_mut_[ia] = {tCR, tF};
    gen_[ia].second = xtrial;
    cands.push_back(xtrial);
  }
}
void DeSearch::consume(const double *costs)
{
  if (phase_ == 0) { xb_.first = costs[0]; nfunc_ = 1; phase_ = 1; return; }
  if (phase_ == 1) {                                                    // de.cpp:106-111
    for (int a = 1; a < kNP; a++) pop_[a].first = costs[a - 1];
    nfunc_ += kNP - 1;
    for (int a = 1; a < kNP; a++) if (pop_[a].first < xb_.first) xb_ = pop_[a];
    phase_ = 2;
    return;
  }
  const int num_agents = (int)gen_.size();
  for (int ia = 0; ia < num_agents; ia++) gen_[ia].first = costs[ia];
  nfunc_ += num_agents;
  std::vector<double> CR_succ, F_succ;                                  // greedy selection (de.cpp:146-158)
  for (int ia = 0; ia < num_agents; ia++)
    if (gen_[ia].first < pop_[ia].first) {
      pop_[ia] = gen_[ia];
      CR_succ.push_back(gen_mut_[ia].first);
      F_succ.push_back(gen_mut_[ia].second);
      if (pop_[ia].first < xb_.first) xb_ = pop_[ia];
    }
  if (nfunc_ >= nfunc_max_) return;
  double mean = 0.0;                                                    // MathUtils::mean / meanL (utils.h:283-304)
  if (!CR_succ.empty()) { double s = 0.0; for (double x : CR_succ) s += x; mean = s / (double)CR_succ.size(); }
  double meanl = 0.0;
  if (!F_succ.empty()) { double s0 = 0.0, s1 = 0.0; for (double x : F_succ) { s0 += (x * x); s1 += x; } meanl = s1 > 0.0 ? s0 / s1 : 0.0; }
  mCR_ = (1.0 - 0.1) * mCR_ + 0.1 * mean;
  mF_ = (1.0 - 0.1) * mF_ + 0.1 * meanl;
}

// =====================================================================================================================
// CMA (src/opt/cma.cpp): (1+1)-ES with covariance adaptation; same draws and operation order as the reference
// =====================================================================================================================
namespace {
// slmath::dot (common/math.h:130-161): two 4-lane fma accumulators over blocks of 8, lane sums, then libstdc++'s
// transform_reduce shape for the tail (blocks of 4 as (a0+a1)+(a2+a3), then one by one)
double slmath_dot(const double *x, const double *y, size_t n)
{
  double total = 0.0;
  size_t i = 0;
  if (n >= 8) {
    double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    for (; i + 8 <= n; i += 8)
      for (int l = 0; l < 4; l++) { s1[l] = std::fma(x[i + l], y[i + l], s1[l]); s2[l] = std::fma(x[i + 4 + l], y[i + 4 + l], s2[l]); }
    for (int l = 0; l < 4; l++) s1[l] = s1[l] + s2[l];
    total = s1[0] + s1[1] + s1[2] + s1[3];
  }
  double init = 0.0;
  while (n - i >= 4) {
    const double v1 = x[i] * y[i] + x[i + 1] * y[i + 1];
    const double v2 = x[i + 2] * y[i + 2] + x[i + 3] * y[i + 3];
    init = init + (v1 + v2);
    i += 4;
  }
  for (; i < n; i++) init = init + x[i] * y[i];
  total += init;
  return total;
}
} // namespace

CmaSearch::CmaSearch(int D, const double *xmin, const double *xmax, const double *xstart, int nfunc_max, double sigma_init)
    : D_(D), nfunc_max_(nfunc_max), xmin_(xmin, xmin + D), xmax_(xmax, xmax + D), xb_(xstart, xstart + D), pc_(D, 0.0), az_(D, 0.0),
      mcov_(D, std::vector<double>(D, 0.0)), G_(D, std::vector<double>(D, 0.0)), sigma_(sigma_init)
{
  const double n = (double)D;
  dinv_ = 1.0 / (1.0 + n / 2.0);                                        // SSC1 p_d = 1/d
  p_target_ = 2.0 / 11.0;
  cp_ = 1.0 / 12.0;
  cc_ = 2.0 / (n + 2.0);
  ccov_ = 2.0 / (n * n + 6.0);
  p_succ_ = p_target_;
  for (int i = 0; i < D; i++) mcov_[i][i] = 1.0;
}
double CmaSearch::reflect(double xnew, double lo, double hi) const
{
  if (xnew < lo) { xnew = lo + (lo - xnew); if (xnew > hi) xnew = lo; }
  if (xnew > hi) { xnew = hi - (xnew - hi); if (xnew < lo) xnew = hi; }
  return xnew;
}
void CmaSearch::propose(std::vector<std::vector<double>> &cands)
{
  cands.clear();
  if (!started_) { cands.push_back(xb_); return; }
  // chol.Factor(mcov, 0.1) (math.h:89-111); a failed factorisation leaves G partially updated, as in the reference
  {
    const double ftol = 1E-8;
    for (int i = 0; i < D_; i++) std::copy_n(mcov_[i].begin(), i + 1, G_[i].begin());
    for (int i = 0; i < D_; i++) {
      for (int j = 0; j < i; j++) {
        double sum = G_[i][j];
        for (int k = 0; k < j; k++) sum -= (G_[i][k] * G_[j][k]);
        G_[i][j] = sum / G_[j][j];
      }
      double sum = G_[i][i] + 0.1;
      for (int k = 0; k < i; k++) sum -= (G_[i][k] * G_[i][k]);
      if (sum > ftol) G_[i][i] = std::sqrt(sum); else break;
    }
  }
  std::vector<double> z(D_);
  for (auto &r : z) r = std::normal_distribution<double>{0.0, 1.0}(eng_);
  for (int i = 0; i < D_; i++) az_[i] = slmath_dot(G_[i].data(), z.data(), (size_t)D_);   // slmath::mul(chol.G, z): full rows of G
  xgen_.assign(D_, 0.0);
  for (int i = 0; i < D_; i++) {
    const double scale = (xmax_[i] - xmin_[i]) * sigma_;
    const double xnew = xb_[i] + scale * az_[i];
    xgen_[i] = reflect(xnew, xmin_[i], xmax_[i]);
  }
  cands.push_back(xgen_);
}
void CmaSearch::consume(const double *costs)
{
  if (!started_) { fb_ = costs[0]; nfunc_ = 1; started_ = true; return; }
  const double fn = costs[0];
  const double lambda = (fn < fb_) ? 1.0 : 0.0;
  p_succ_ = (1.0 - cp_) * p_succ_ + cp_ * lambda;                       // SSC1::update (ssc.h:49-55)
  sigma_ = sigma_ * std::exp(dinv_ * (p_succ_ - p_target_) / (1.0 - p_target_));
  sigma_ = std::clamp(sigma_, 0.05, 0.25);
  if (fn < fb_) {
    fb_ = fn;
    xb_ = xgen_;
    const double a = 1.0 - cc_, b = std::sqrt(cc_ * (2.0 - cc_));       // update_cov (cma.cpp:49-53)
    for (int i = 0; i < D_; i++) pc_[i] = a * pc_[i] + b * az_[i];
    for (int j = 0; j < D_; j++)
      for (int i = 0; i < D_; i++) mcov_[j][i] = (1.0 - ccov_) * mcov_[j][i] + ccov_ * (pc_[j] * pc_[i]);
  }
  nfunc_++;
}

// =====================================================================================================================
// MD5
// =====================================================================================================================
namespace {
const uint32_t kMd5K[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1,
    0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453,
    0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942,
    0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05,
    0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d,
    0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
const int kMd5S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20,
                       4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
inline uint32_t rol(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }
void md5_block(Md5 &m, const uint8_t *p)
{
  uint32_t w[16];
  for (int i = 0; i < 16; i++) w[i] = p[4 * i] | (p[4 * i + 1] << 8) | (p[4 * i + 2] << 16) | ((uint32_t)p[4 * i + 3] << 24);
  uint32_t a = m.a, b = m.b, c = m.c, d = m.d;
  for (int i = 0; i < 64; i++) {
    uint32_t f; int g;
    if (i < 16) { f = (b & c) | (~b & d); g = i; }
    else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
    else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
    else { f = c ^ (b | ~d); g = (7 * i) & 15; }
    const uint32_t t = d; d = c; c = b;
    b = b + rol(a + f + kMd5K[i] + w[g], kMd5S[i]);
    a = t;
  }
  m.a += a; m.b += b; m.c += c; m.d += d;
}
} // namespace
Md5::Md5() : a(0x67452301), b(0xefcdab89), c(0x98badcfe), d(0x10325476), len(0), fill(0) {}
void Md5::update(const uint8_t *p, size_t n)
{
  len += n;
  while (n) {
    const size_t take = std::min(n, (size_t)(64 - fill));
    std::memcpy(buf + fill, p, take);
    fill += (int)take; p += take; n -= take;
    if (fill == 64) { md5_block(*this, buf); fill = 0; }
  }
}
void Md5::final(uint8_t out[16])
{
  const uint64_t bits = len * 8;
  uint8_t pad[72] = {0x80};
  const size_t padlen = (fill < 56) ? (56 - fill) : (120 - fill);
  update(pad, padlen);
  uint8_t lb[8];
  for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (8 * i));
  update(lb, 8);
  const uint32_t v[4] = {a, b, c, d};
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out[4 * i + j] = (uint8_t)(v[i] >> (8 * j));
}

// =====================================================================================================================
// WAV / metadata
// =====================================================================================================================
namespace {
inline uint32_t get32(const uint8_t *b) { return b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24); }
inline uint16_t get16(const uint8_t *b) { return (uint16_t)(b[0] | (b[1] << 8)); }
inline void put32(uint8_t *b, uint32_t v) { b[0] = v & 0xff; b[1] = (v >> 8) & 0xff; b[2] = (v >> 16) & 0xff; b[3] = (v >> 24) & 0xff; }
inline void put16(uint8_t *b, uint16_t v) { b[0] = v & 0xff; b[1] = (v >> 8) & 0xff; }
inline void push32(std::vector<uint8_t> &o, uint32_t v) { uint8_t b[4]; put32(b, v); o.insert(o.end(), b, b + 4); }
inline void push16(std::vector<uint8_t> &o, uint16_t v) { uint8_t b[2]; put16(b, v); o.insert(o.end(), b, b + 2); }
inline uint32_t word_align(uint32_t n) { return (n & 1) ? n + 1 : n; }
constexpr uint32_t kRIFF = 0x46464952, kWAVE = 0x45564157, kFMT = 0x20746d66, kDATA = 0x61746164;
} // namespace

uint32_t WavInfo::metadatasize() const
{
  uint32_t t = 0;
  for (auto &c : chunks) t += 8 + (uint32_t)c.data.size();
  return t;
}

// Wav::ReadHeader (wav.cpp:167-263)
int wav_parse(const uint8_t *f, size_t len, WavInfo &wi)
{
  if (len < 12 || get32(f) != kRIFF || get32(f + 8) != kWAVE) { set_error("input is not a valid .wav file"); return SAC_E_FORMAT; }
  wi.chunks.clear();
  wi.chunks.push_back({kRIFF, get32(f + 4), std::vector<uint8_t>(f + 8, f + 12)});
  size_t pos = 12;
  bool have_fmt = false, have_data = false;
  while (true) {
    if (pos + 8 > len) { set_error("could not read wav chunk"); return SAC_E_FORMAT; }   // also after 'data': stray bytes are an error (wav.cpp:182-183)
    const uint32_t id = get32(f + pos), cs = get32(f + pos + 4);
    pos += 8;
    if (id == kFMT) {
      if ((cs != 16 && cs != 18 && cs != 40) || pos + cs > len) { set_error("invalid fmt-chunk size"); return SAC_E_FORMAT; }
      const uint8_t *b = f + pos;
      int audioformat = get16(b);
      wi.nch = get16(b + 2); wi.samplerate = (int)get32(b + 4); wi.blockalign = get16(b + 12); wi.bits = get16(b + 14);
      if (cs >= 18) {
        const int cb = get16(b + 16);
        if (cb >= 22 && cs >= 40) { wi.bits = get16(b + 18); audioformat = get16(b + 24); }
      }
      wi.chunks.push_back({id, cs, std::vector<uint8_t>(b, b + cs)});
      pos += cs;
      if (audioformat != 1) { set_error("only PCM format supported"); return SAC_E_UNSUPPORTED; }
      have_fmt = true;
    } else if (id == kDATA) {
      if (!have_fmt || wi.blockalign <= 0) { set_error("data chunk before fmt chunk"); return SAC_E_FORMAT; }
      wi.chunks.push_back({id, cs, {}});
      wi.data_pos = pos;
      wi.numsamples = cs / wi.blockalign;
      const size_t endofdata = pos + word_align(cs);
      have_data = true;
      if (endofdata >= len) {
        if (endofdata > len) wi.numsamples = (uint32_t)((len - pos) / wi.blockalign);   // truncated data chunk (wav.cpp:238-243)
        break;
      }
      // Deliberate deviation: the reference seeks by chunksize (wav.cpp:246-247), which lands on the pad byte of an
      // odd-sized data chunk and misparses every chunk after it; its own decoder re-inserts that pad byte
      // (libsac.cpp:880-881), so the aligned skip is what round-trips.
      pos += word_align(cs);
    } else {
      const uint32_t rs = word_align(cs);
      if (pos + rs > len) { set_error("truncated wav chunk"); return SAC_E_FORMAT; }
      wi.chunks.push_back({id, cs, std::vector<uint8_t>(f + pos, f + pos + rs)});
      pos += rs;
    }
    if (pos == len) break;
  }
  if (!have_data || !have_fmt) { set_error("wav file without fmt/data chunk"); return SAC_E_FORMAT; }
  return SAC_OK;
}

// Wav::ReadSamples (wav.cpp:77-124)
void wav_unpack(const WavInfo &wi, const uint8_t *pcm, int first, int count, std::vector<std::vector<int32_t>> &planes)
{
  const int cs = wi.blockalign / wi.nch;
  const uint8_t *p = pcm + (size_t)first * wi.blockalign;
  for (int i = 0; i < count; i++)
    for (int k = 0; k < wi.nch; k++) {
      int32_t v = 0;
      if (cs == 1) { v = (int32_t)p[0] - 128; }
      else if (cs == 2) { v = (int16_t)((p[1] << 8) | p[0]); }
      else if (cs == 3) { int32_t s = ((int32_t)p[2] << 24) | ((int32_t)p[1] << 16) | ((int32_t)p[0] << 8); v = s >> 8; }
      planes[k][i] = v;
      p += cs;
    }
}
// Wav::WriteSamples (wav.cpp:126-164)
void wav_pack(const WavInfo &wi, const std::vector<std::vector<int32_t>> &planes, int count, std::vector<uint8_t> &out)
{
  const int cs = wi.blockalign / wi.nch;
  const size_t o0 = out.size();
  out.resize(o0 + (size_t)count * wi.blockalign);
  uint8_t *p = out.data() + o0;
  for (int i = 0; i < count; i++)
    for (int k = 0; k < wi.nch; k++) {
      const int32_t s = planes[k][i];
      if (cs == 1) p[0] = (uint8_t)((s + 128) & 0xff);
      else if (cs == 2) { p[0] = s & 0xff; p[1] = (s >> 8) & 0xff; }
      else if (cs == 3) { p[0] = s & 0xff; p[1] = (s >> 8) & 0xff; p[2] = (s >> 16) & 0xff; }
      p += cs;
    }
}
// Chunks::PackMetaData / UnpackMetaData (wav.cpp:24-53)
void pack_metadata(const WavInfo &wi, std::vector<uint8_t> &out)
{
  for (auto &c : wi.chunks) { push32(out, c.id); push32(out, c.csize); out.insert(out.end(), c.data.begin(), c.data.end()); }
}
int unpack_metadata(const uint8_t *p, size_t n, WavInfo &wi)
{
  size_t ofs = 0;
  wi.chunks.clear();
  while (ofs + 8 <= n) {
    const uint32_t id = get32(p + ofs), cs = get32(p + ofs + 4);
    ofs += 8;
    if (id == kRIFF) { if (ofs + 4 > n) return SAC_E_FORMAT; wi.chunks.push_back({id, cs, std::vector<uint8_t>(p + ofs, p + ofs + 4)}); ofs += 4; }
    else if (id == kDATA) wi.chunks.push_back({id, cs, {}});
    else {
      const uint32_t ws = word_align(cs);
      if (ofs + ws > n) return SAC_E_FORMAT;
      wi.chunks.push_back({id, cs, std::vector<uint8_t>(p + ofs, p + ofs + ws)});
      ofs += ws;
    }
  }
  return ofs == n ? SAC_OK : SAC_E_FORMAT;
}

// =====================================================================================================================
// frames
// =====================================================================================================================
namespace {

struct FrameWork {
  int n = 0;
  int32_t mean[2] = {0, 0}, mm[4] = {0, 0, 0, 0};
  Window *win = nullptr;
  float profile[kProfileSize];
};

// AnalyseMonoChannel + zero-mean (libsac.cpp:626-651, 452-458)
void analyse_channel(std::vector<int32_t> &s, int zero_mean, int32_t &mean, int32_t &mn, int32_t &mx)
{
  int64_t sum = 0;
  mn = std::numeric_limits<int32_t>::max(); mx = std::numeric_limits<int32_t>::min();
  for (int32_t v : s) { sum += v; mx = std::max(mx, v); mn = std::min(mn, v); }
  mean = s.empty() ? 0 : (int)std::floor(sum / (double)s.size());
  if (!zero_mean) mean = 0;
  else if (mean != 0) { for (auto &v : s) v -= mean; mn -= mean; mx -= mean; }
}

} // namespace

static int frames_encode_seq(Engine *e, const sac_cfg &cfg, int nch, int max_framesize, int nframes, const int32_t *const *planes,
                             const int *numsamples, float *profile_io, std::vector<uint8_t> &out,
                             const sac_window *const *resident = nullptr, const int32_t *resident_means = nullptr);

// frame_parallel == 2: every frame of the call is encoded by its own host thread on its own helper engine (own
// stream and pools), all concurrently on this GPU. A single frame's kernels are bound by their slowest chain and
// leave most of the machine idle; several frames in flight fill it. Frames start from the base profile (the
// reference's --opt-reset semantics, cmdline.cpp:193).
static int frames_encode_streams(Engine *e, const sac_cfg &cfg, int nch, int max_framesize, int nframes, const int32_t *const *planes,
                                 const int *numsamples, float *profile_io, std::vector<uint8_t> &out,
                                 const sac_window *const *resident, const int32_t *resident_means)
{
  const int par = std::min(std::clamp(cfg.inflight, 1, 64), nframes);   // frames in flight: one worker (host thread + helper engine) each
  const int kBase = 8;                                               // helper slots 0..3 serve final passes of the main engine
  std::vector<std::vector<uint8_t>> outs(nframes);
  std::vector<int> rcs(nframes, SAC_OK);
  std::vector<std::string> errs(nframes);
  std::vector<std::array<float, kProfileSize>> profs(nframes);
  sac_cfg one = cfg;
  one.frame_parallel = 0; one.reset = 1;
  for (int i = 0; i < par; i++) if (!e->helper(kBase + i)) return SAC_E_CUDA;
  // workers pull the next frame as soon as they are done with one: the GPU never waits for the slowest frame of a group
  std::atomic<int> next{0};
  std::vector<std::thread> th;
  for (int w = 0; w < par; w++) {
    th.emplace_back([&, w]() {
      Engine *h = e->helpers[kBase + w];
      cudaSetDevice(h->device);
      for (;;) {
        const int f = next.fetch_add(1);
        if (f >= nframes) break;
        for (int i = 0; i < kProfileSize; i++) profs[f][i] = kBaseProfile[i][2];
        rcs[f] = frames_encode_seq(h, one, nch, max_framesize, 1, planes ? planes + (size_t)f * nch : nullptr, numsamples ? numsamples + f : nullptr,
                                   profs[f].data(), outs[f], resident ? resident + f : nullptr, resident_means ? resident_means + (size_t)f * nch : nullptr);
        if (rcs[f]) errs[f] = sac_last_error();
      }
    });
  }
  for (auto &t : th) t.join();
  for (int f = 0; f < nframes; f++) {
    if (rcs[f]) { set_error(errs[f]); return rcs[f]; }
    out.insert(out.end(), outs[f].begin(), outs[f].end());
  }
  std::memcpy(profile_io, profs[nframes - 1].data(), sizeof(float) * kProfileSize);
  return SAC_OK;
}

static int frames_encode(Engine *e, const sac_cfg &cfg, int nch, int max_framesize, int nframes, const int32_t *const *planes,
                         const int *numsamples, float *profile_io, std::vector<uint8_t> &out,
                         const sac_window *const *resident = nullptr, const int32_t *resident_means = nullptr)
{
  if (cfg.search != SAC_SEARCH_DDS && cfg.search != SAC_SEARCH_DE && cfg.search != SAC_SEARCH_CMA) {
    set_error("unknown search method");
    return SAC_E_ARG;
  }
  if (cfg.frame_parallel == 2 && nframes > 1)
    return frames_encode_streams(e, cfg, nch, max_framesize, nframes, planes, numsamples, profile_io, out, resident, resident_means);
  return frames_encode_seq(e, cfg, nch, max_framesize, nframes, planes, numsamples, profile_io, out, resident, resident_means);
}

static int frames_encode_seq(Engine *e, const sac_cfg &cfg, int nch, int max_framesize, int nframes, const int32_t *const *planes,
                             const int *numsamples, float *profile_io, std::vector<uint8_t> &out,
                             const sac_window *const *resident, const int32_t *resident_means)
{
  // arithmetic of the search evaluations for the duration of this call (final passes are canonical); the engine's own setting
  // (sac_engine_set_grade, used by sac_predict / sac_eval_*) is restored on every way out
  struct GradeGuard { Engine *e; int prev; ~GradeGuard() { e->grade = prev; } } grade_guard{e, e->grade};
  e->grade = cfg.grade ? 1 : 0;
  std::vector<FrameWork> fw(nframes);
  struct Cleanup { std::vector<FrameWork> &f; bool own; ~Cleanup() { if (own) for (auto &x : f) if (x.win) sac_window_destroy(reinterpret_cast<sac_window *>(x.win)); } } cleanup{fw, resident == nullptr};
  // ---- analysis + upload (or windows already resident in HBM) ----
  for (int f = 0; f < nframes; f++) {
    FrameWork &w = fw[f];
    if (resident) {
      w.win = const_cast<Window *>(reinterpret_cast<const Window *>(resident[f]));
      if (!w.win || w.win->nch != nch) { set_error("resident window mismatch"); return SAC_E_ARG; }
      w.n = w.win->numsamples;
      for (int ch = 0; ch < nch; ch++) { w.mean[ch] = resident_means ? resident_means[f * nch + ch] : 0; w.mm[2 * ch] = w.win->minmax[2 * ch]; w.mm[2 * ch + 1] = w.win->minmax[2 * ch + 1]; }
      if (w.n <= 0 || w.n > max_framesize) { set_error("frame length out of range"); return SAC_E_ARG; }
      continue;
    }
    w.n = numsamples[f];
    if (w.n <= 0 || w.n > max_framesize) { set_error("frame length out of range"); return SAC_E_ARG; }
    std::vector<std::vector<int32_t>> s(nch);
    const int32_t *pp[2] = {nullptr, nullptr};
    for (int ch = 0; ch < nch; ch++) {
      s[ch].assign(planes[f * nch + ch], planes[f * nch + ch] + w.n);
      analyse_channel(s[ch], cfg.zero_mean, w.mean[ch], w.mm[2 * ch], w.mm[2 * ch + 1]);
      pp[ch] = s[ch].data();
    }
    w.win = reinterpret_cast<Window *>(sac_window_create(reinterpret_cast<sac_engine *>(e), nch, pp, w.n, w.mm));
    if (!w.win) return SAC_E_CUDA;
  }
  // ---- final pass (k=1, whole frame) + payload emission of one frame on a helper engine: it is launched as soon as
  //      the frame's profile is known and runs on its own high-priority stream while the main stream searches the
  //      next frame; results are collected at the end (WriteEncoded order) ----
  struct Final {
    Engine *eng = nullptr; bool launched = false; int nchains = 0; std::vector<int> cj, cc; std::vector<size_t> boff;
    std::vector<int> sp_of; int nsp = 0; size_t sp_res_off = 0; std::vector<size_t> sp_bytes_off;   // sparse-PCM side (sparse.h)
  };
  std::vector<Final> fin(nframes);
  const int kHelpers = 4;
  auto collect_final = [&](int f) -> int {
    Final &F = fin[f];
    if (!F.launched || !F.eng) return SAC_OK;
    Engine *h = F.eng;
    SACB_CUDA(h->wait());
    for (int c = 0; c < F.nchains; c++)
      if (h->h_flags.p[c]) { set_error("final pass: predictor state became non-finite"); return SAC_E_UNSUPPORTED; }
    // serialise (WriteEncoded / WriteBlockHeader / EncodeProfile, libsac.cpp:507-578)
    std::vector<uint8_t> payload;
    push32(out, (uint32_t)fw[f].n);
    for (int i = 0; i < kProfileSize; i++) { uint32_t ix; std::memcpy(&ix, &fw[f].profile[i], 4); push32(out, ix); }
    for (int ch = 0; ch < nch; ch++) {
      int c = -1;
      for (int q = 0; q < F.nchains; q++) if (F.cc[q] == ch) c = q;
      long long nb = h->h_sums.p[2 * F.nchains + c];
      const int maxbpn = h->h_flags.p[F.nchains + c];
      const uint8_t *src = h->d_bytes.p + F.boff[c];
      uint16_t flag = (uint16_t)(maxbpn & 0xff);
      const int k = F.sp_of[c];
      if (k >= 0 && h->h_sparse.p[k].go && h->h_sparse.p[k].nbytes < nb) {     // size_mapped < size_normal (libsac.cpp:267-275)
        if (cfg.verbose > 0) std::fprintf(stderr, "  sparse frame %lld -> %lld (%lld)\n", nb, h->h_sparse.p[k].nbytes, h->h_sparse.p[k].nbytes - nb);
        nb = h->h_sparse.p[k].nbytes;
        src = h->d_sparse.p + F.sp_bytes_off[k];
        flag = (uint16_t)((1 << 9) | (h->h_sparse.p[k].maxbpn & 0xff));        // WriteBlockHeader, libsac.cpp:536-541
      }
      payload.resize((size_t)nb);
      SACB_CUDA(cudaMemcpyAsync(payload.data(), src, (size_t)nb, cudaMemcpyDeviceToHost, h->stream));
      SACB_CUDA(cudaStreamSynchronize(h->stream));
      push32(out, (uint32_t)nb); push32(out, (uint32_t)fw[f].mean[ch]); push32(out, (uint32_t)fw[f].mm[2 * ch]); push32(out, (uint32_t)fw[f].mm[2 * ch + 1]);
      push16(out, flag);
      out.insert(out.end(), payload.begin(), payload.end());
    }
    F.eng = nullptr;
    return SAC_OK;
  };
  auto launch_final = [&](int f) -> int {
    Final &F = fin[f];
    // a helper is reused only after the frame that used it has been collected; frames are collected in order
    if (f >= kHelpers) { for (int g = 0; g <= f - kHelpers; g++) { int rc = collect_final(g); if (rc) return rc; } }
    Engine *h = e->helper(f % kHelpers);
    if (!h) return SAC_E_CUDA;
    F.eng = h;
    std::vector<Job> jobs(1);
    jobs[0].win = fw[f].win; jobs[0].from = 0; jobs[0].n = fw[f].n; jobs[0].k = 1;
    std::memcpy(jobs[0].profile, fw[f].profile, sizeof(float) * kProfileSize);
    h->begin_call();
    size_t stride;
    int rc = h->run_predict(jobs, F.cj, F.cc, stride);
    if (rc) return rc;
    const int nchains = F.nchains = (int)F.cj.size();
    F.sp_of.assign(nchains, -1); F.nsp = 0;
    if (cfg.sparse_pcm)
      for (int c = 0; c < nchains; c++) {
        const int ch = F.cc[c];
        const long long lo = (long long)fw[f].mm[2 * ch] + fw[f].mean[ch], hi = (long long)fw[f].mm[2 * ch + 1] + fw[f].mean[ch];
        if (lo >= -kRemapScale && hi <= kRemapScale) F.sp_of[c] = F.nsp++;
      }
    SACB_CUDA(h->h_bpjobs.reserve(nchains + F.nsp));
    SACB_CUDA(h->d_bpjobs.reserve(nchains + F.nsp));
    SACB_CUDA(h->d_csig0.reserve((size_t)65536 * (nchains + F.nsp)));
    F.boff.assign(nchains + 1, 0);
    for (int c = 0; c < nchains; c++) F.boff[c + 1] = F.boff[c] + (((size_t)fw[f].n * 4 + 1024 + 15) & ~size_t(15));
    SACB_CUDA(h->d_bytes.reserve(F.boff[nchains]));
    for (int c = 0; c < nchains; c++) {
      BpJob &b = h->h_bpjobs.p[c];
      std::memset(&b, 0, sizeof(b));
      b.buf = h->d_resid.p + (size_t)c * stride; b.n = fw[f].n; b.signed_input = 1; b.maxbpn = -1;
      b.csig0 = h->d_csig0.p + (size_t)65536 * c; b.out = h->d_bytes.p + F.boff[c];
      b.nbytes = h->d_sums.p + 2 * (size_t)nchains + c; b.maxbpn_out = h->d_flags.p + nchains + c;
    }
    // ---- sparse-PCM (FrameCoder::EncodeMonoFrame, libsac.cpp:253-278): the channel is ALSO coded as rank distances
    //      among the sample values that occur, behind the coded map of those values, when the L1 ratio calls for it;
    //      the shorter record is kept at collection. The reference's map covers |value| <= 32768 only and its mapped
    //      records of wider samples do not decode (map.cpp:126-166): such channels stay unmapped here. ----
    int njobs_bp = nchains;
    if (F.nsp > 0) {
      SparseJobs SJ;
      std::memset(&SJ, 0, sizeof(SJ));
      SJ.n = F.nsp;
      const size_t nsp = (size_t)F.nsp;
      const size_t o_rc = 0, o_res = o_rc + sizeof(RcInit) * nsp, o_used = (o_res + sizeof(SparseOut) * nsp + 127) & ~size_t(127);
      const size_t o_cum = o_used + kRemapUsedBytes * nsp, cumb = ((size_t)kRemapDom * 4 + 127) & ~size_t(127);
      const size_t o_list = o_cum + cumb * nsp, o_em = o_list + cumb * nsp, emb = stride * 4;
      const size_t o_bytes = o_em + emb * nsp, bytesb = ((size_t)fw[f].n * 4 + 1024 + kMapBytesMax + 127) & ~size_t(127);
      SACB_CUDA(h->d_sparse.reserve(o_bytes + bytesb * nsp));
      SACB_CUDA(cudaMemsetAsync(h->d_sparse.p, 0, o_cum, h->stream));
      F.sp_res_off = o_res; F.sp_bytes_off.assign(nsp, 0);
      for (int c = 0; c < nchains; c++) {
        const int k = F.sp_of[c];
        if (k < 0) continue;
        const int ch = F.cc[c];
        SparseJob &j = SJ.j[k];
        j.s = fw[f].win->d_planes[ch]; j.e = h->d_resid.p + (size_t)c * stride; j.n = fw[f].n; j.mean = fw[f].mean[ch];
        j.em = reinterpret_cast<int32_t *>(h->d_sparse.p + o_em + emb * k);
        j.used = h->d_sparse.p + o_used + kRemapUsedBytes * k;
        j.cum = reinterpret_cast<int32_t *>(h->d_sparse.p + o_cum + cumb * k);
        j.ulist = reinterpret_cast<int32_t *>(h->d_sparse.p + o_list + cumb * k);
        j.bytes = h->d_sparse.p + o_bytes + bytesb * k; F.sp_bytes_off[k] = o_bytes + bytesb * k;
        j.rc = reinterpret_cast<RcInit *>(h->d_sparse.p + o_rc) + k;
        j.res = reinterpret_cast<SparseOut *>(h->d_sparse.p + o_res) + k;
        BpJob &b = h->h_bpjobs.p[nchains + k];
        std::memset(&b, 0, sizeof(b));
        b.buf = j.em; b.n = fw[f].n; b.signed_input = 1; b.maxbpn = -1;
        b.csig0 = h->d_csig0.p + (size_t)65536 * (nchains + k); b.out = j.bytes;
        b.nbytes = &j.res->nbytes; b.maxbpn_out = &j.res->maxbpn; b.rc_init = j.rc;
      }
      SACB_CUDA(launch_sparse_encode(h->bt, SJ, h->stream));
      h->launches += 4; h->last_launches[2] += 4;
      njobs_bp = nchains + F.nsp;
    }
    SACB_CUDA(cudaMemcpyAsync(h->d_bpjobs.p, h->h_bpjobs.p, sizeof(BpJob) * njobs_bp, cudaMemcpyHostToDevice, h->stream));
    SACB_CUDA(launch_bitplane(h->bt, h->d_bpjobs.p, njobs_bp, 1, h->stream));
    h->launches++; h->last_launches[1]++;
    if (F.nsp > 0) {
      SACB_CUDA(h->h_sparse.reserve((size_t)F.nsp));
      SACB_CUDA(cudaMemcpyAsync(h->h_sparse.p, h->d_sparse.p + F.sp_res_off, sizeof(SparseOut) * F.nsp, cudaMemcpyDeviceToHost, h->stream));
    }
    SACB_CUDA(h->h_sums.reserve((size_t)3 * nchains));
    SACB_CUDA(h->h_flags.reserve((size_t)2 * nchains));
    SACB_CUDA(cudaMemcpyAsync(h->h_sums.p, h->d_sums.p, sizeof(long long) * 3 * nchains, cudaMemcpyDeviceToHost, h->stream));
    SACB_CUDA(cudaMemcpyAsync(h->h_flags.p, h->d_flags.p, sizeof(int) * 2 * nchains, cudaMemcpyDeviceToHost, h->stream));
    F.launched = true;
    return SAC_OK;
  };

  // ---- search (FrameCoder::Optimize, libsac.cpp:365-427; Predict :461-476) ----
  std::vector<int> dims;
  for (int i = 0; i < kProfileSize; i++) if (i != 56 && i != 57) dims.push_back(i);
  const int D = (int)dims.size();
  std::vector<double> xmin(D), xmax(D);
  for (int i = 0; i < D; i++) { xmin[i] = kBaseProfile[dims[i]][0]; xmax[i] = kBaseProfile[dims[i]][1]; }
  float cur[kProfileSize];
  std::memcpy(cur, profile_io, sizeof(cur));
  auto base_reset = [&](float *p) { for (int i = 0; i < kProfileSize; i++) p[i] = kBaseProfile[i][2]; };

  if (cfg.optimize && cfg.maxnfunc > 0 && cfg.fraction > 0.0) {
    const int groups = cfg.frame_parallel == 1 ? 1 : nframes;     // frames searched together per group
    for (int g = 0; g < groups; g++) {
      const int f0 = cfg.frame_parallel == 1 ? 0 : g, f1 = cfg.frame_parallel == 1 ? nframes : g + 1;
      std::vector<std::unique_ptr<Search>> ss;
      std::vector<int> wfrom, wn;
      for (int f = f0; f < f1; f++) {
        if (cfg.reset) base_reset(fw[f].profile); else std::memcpy(fw[f].profile, cur, sizeof(cur));
        std::vector<double> xs(D);
        for (int i = 0; i < D; i++) xs[i] = fw[f].profile[dims[i]];
        if (cfg.search == SAC_SEARCH_DE) ss.emplace_back(new DeSearch(D, xmin.data(), xmax.data(), xs.data(), cfg.maxnfunc, cfg.sigma));
        else if (cfg.search == SAC_SEARCH_CMA) ss.emplace_back(new CmaSearch(D, xmin.data(), xmax.data(), xs.data(), cfg.maxnfunc, 0.0));   // cma_cfg.sigma_init stays 0 (cmdline.cpp:232-235)
        else {
          DdsSearch *ds = new DdsSearch(D, xmin.data(), xmax.data(), xs.data(), cfg.maxnfunc, cfg.num_threads, cfg.sigma, cfg.spec);
          if (std::getenv("SACB_TRACE_DDS")) ds->trace(true);
          ss.emplace_back(ds);
        }
        const int nopt = std::min(fw[f].n, (int)std::ceil(max_framesize * cfg.fraction));   // libsac.cpp:367-368
        wn.push_back(nopt); wfrom.push_back((fw[f].n - nopt) / 2);
      }
      const auto t_search0 = std::chrono::steady_clock::now();
      while (true) {
        std::vector<const sac_window *> wins;
        std::vector<int> from, nn, owner;
        std::vector<float> bases;
        std::vector<double> X;
        std::vector<std::vector<std::vector<double>>> cands(ss.size());
        for (size_t si = 0; si < ss.size(); si++) {
          if (ss[si]->done()) continue;
          ss[si]->propose(cands[si]);
          for (auto &c : cands[si]) {
            wins.push_back(reinterpret_cast<const sac_window *>(fw[f0 + si].win));
            from.push_back(wfrom[si]); nn.push_back(wn[si]); owner.push_back((int)si);
            bases.insert(bases.end(), fw[f0 + si].profile, fw[f0 + si].profile + kProfileSize);
            X.insert(X.end(), c.begin(), c.end());
          }
        }
        if (wins.empty()) break;
        std::vector<double> cost(wins.size()), thr(wins.size());
        for (size_t j = 0; j < wins.size(); j++) thr[j] = ss[owner[j]]->accept_below();
        e->redo_below = thr.data();
        int rc = sac_eval_jobs(reinterpret_cast<sac_engine *>(e), (int)wins.size(), wins.data(), from.data(), nn.data(), bases.data(),
                               dims.data(), D, X.data(), cfg.cost_kind, cfg.optk, cost.data());
        e->redo_below = nullptr;
        if (rc) return rc;
        size_t off = 0;
        for (size_t si = 0; si < ss.size(); si++) {
          if (cands[si].empty()) continue;
          ss[si]->consume(&cost[off]);
          off += cands[si].size();
          if (cfg.verbose > 1)
            std::fprintf(stderr, "  frame %d DDS %5d: %0.4f s=%0.3f  batch %d  t=%.2fs\n", f0 + (int)si, ss[si]->nfunc(), ss[si]->best_cost(), ss[si]->sigma(),
                         (int)cands[si].size(), std::chrono::duration<double>(std::chrono::steady_clock::now() - t_search0).count());
        }
      }
      if (const char *tp = std::getenv("SACB_TRACE_DDS")) {           // probe: the steps of the sequential search, one line each
        static std::mutex trace_mu;
        std::lock_guard<std::mutex> lk(trace_mu);
        if (FILE *tf = std::fopen(tp, "a")) {
          for (int f = f0; f < f1; f++)
            if (auto *ds = dynamic_cast<DdsSearch *>(ss[f - f0].get())) {
              int step = 1;
              for (const auto &pr : ds->traced()) {
                std::fprintf(tf, "{\"frame_n\": %d, \"step\": %d, \"cost\": %.17g, \"x\": [", fw[f].n, step++, pr.first);
                for (size_t i = 0; i < pr.second.size(); i++) std::fprintf(tf, "%s%.9g", i ? ", " : "", pr.second[i]);
                std::fprintf(tf, "]}\n");
              }
            }
          std::fclose(tf);
        }
      }
      for (int f = f0; f < f1; f++) {
        const auto &xb = ss[f - f0]->best_x();
        for (int i = 0; i < D; i++) fw[f].profile[dims[i]] = (float)xb[i];                // libsac.cpp:418-420
        std::memcpy(cur, fw[f].profile, sizeof(cur));
        int rc = launch_final(f);                                    // overlaps the next group's search
        if (rc) return rc;
      }
    }
  } else {
    for (int f = 0; f < nframes; f++) std::memcpy(fw[f].profile, cur, sizeof(cur));
  }
  std::memcpy(profile_io, cur, sizeof(cur));

  // ---- whatever final passes are still outstanding, then collect and serialise in frame order ----
  for (int f = 0; f < nframes; f++)
    if (!fin[f].launched) { int rc = launch_final(f); if (rc) return rc; }
  for (int f = 0; f < nframes; f++) { int rc = collect_final(f); if (rc) return rc; }
  return SAC_OK;
}

// one frame record -> planes (ReadEncoded, Decode, Unpredict)
static long long frame_decode(Engine *e, int nch, const uint8_t *in, long long len, std::vector<std::vector<int32_t>> &planes, int cap,
                              int *n_out)
{
  const long long hdr = 4 + 4LL * kProfileSize;
  if (len < hdr) { set_error("truncated frame record"); return SAC_E_FORMAT; }
  const int n = (int)get32(in);
  if (n <= 0 || n > cap) { set_error("frame record: bad sample count"); return SAC_E_FORMAT; }
  float prof[kProfileSize];
  for (int i = 0; i < kProfileSize; i++) { const uint32_t ix = get32(in + 4 + 4 * i); std::memcpy(&prof[i], &ix, 4); }
  long long pos = hdr;
  int32_t mean[2] = {0, 0}, mm[4] = {0, 0, 0, 0};
  int maxbpn[2] = {0, 0};
  bool mapped[2] = {false, false};
  long long poff[2] = {0, 0}, plen[2] = {0, 0};
  for (int ch = 0; ch < nch; ch++) {
    if (pos + 18 > len) { set_error("truncated block header"); return SAC_E_FORMAT; }
    plen[ch] = get32(in + pos);
    mean[ch] = (int32_t)get32(in + pos + 4); mm[2 * ch] = (int32_t)get32(in + pos + 8); mm[2 * ch + 1] = (int32_t)get32(in + pos + 12);
    const uint16_t flag = get16(in + pos + 16);
    mapped[ch] = (flag >> 9) != 0;                                  // ReadBlockHeader, libsac.cpp:551-564
    maxbpn[ch] = flag & 0xff;
    if (maxbpn[ch] > 30) { set_error("block header: bit plane count out of range"); return SAC_E_FORMAT; }   // the model has 32 planes (vle.h), as sac_bitplane_decode checks
    pos += 18;
    poff[ch] = pos;
    if (pos + plen[ch] > len) { set_error("truncated payload"); return SAC_E_FORMAT; }
    pos += plen[ch];
  }
  if (nch == 1) { mm[2] = mm[0]; mm[3] = mm[1]; }
  SACB_CUDA(cudaSetDevice(e->device));
  e->begin_call();
  const size_t stride = ((size_t)n + 31) & ~size_t(31);
  // device buffers: residuals (d_resid), decoded planes (d_scratch tail is not used here: separate allocation)
  SACB_CUDA(e->d_resid.reserve(stride * 2 * nch));                 // [0,nch): residuals, [nch,2nch): decoded samples
  SACB_CUDA(e->d_csig0.reserve((size_t)65536 * nch));
  SACB_CUDA(e->h_bpjobs.reserve(nch));
  SACB_CUDA(e->d_bpjobs.reserve(nch));
  size_t bytes_need = 0;
  for (int ch = 0; ch < nch; ch++) bytes_need += (((size_t)plen[ch] + 15) & ~size_t(15)) + stride;
  SACB_CUDA(e->d_bytes.reserve(bytes_need + 64));
  size_t bo = 0;
  for (int ch = 0; ch < nch; ch++) {
    BpJob &b = e->h_bpjobs.p[ch];
    std::memset(&b, 0, sizeof(b));
    b.buf = e->d_resid.p + (size_t)ch * stride; b.n = n; b.maxbpn = maxbpn[ch];
    b.csig0 = e->d_csig0.p + (size_t)65536 * ch;
    b.in = e->d_bytes.p + bo; b.in_len = plen[ch];
    SACB_CUDA(cudaMemcpyAsync(e->d_bytes.p + bo, in + poff[ch], (size_t)plen[ch], cudaMemcpyHostToDevice, e->stream));
    bo += ((size_t)plen[ch] + 15) & ~size_t(15);
    b.msb = e->d_bytes.p + bo;
    bo += stride;
  }
  // ---- sparse-PCM channels: the payload opens with the coded map of the used sample values (DecodeMonoFrame,
  //      libsac.cpp:280-298); the bitplane decoder continues the same range decoder ----
  SparseJobs SJ;
  std::memset(&SJ, 0, sizeof(SJ));
  int sp_of[2] = {-1, -1};
  for (int ch = 0; ch < nch; ch++) if (mapped[ch]) sp_of[ch] = SJ.n++;
  if (SJ.n > 0) {
    const size_t nsp = (size_t)SJ.n, o_used = (sizeof(RcInit) * nsp + 127) & ~size_t(127), o_cum = o_used + kRemapUsedBytes * nsp;
    const size_t cumb = ((size_t)kRemapDom * 4 + 127) & ~size_t(127), o_list = o_cum + cumb * nsp;
    SACB_CUDA(e->d_sparse.reserve(o_list + cumb * nsp));
    SACB_CUDA(cudaMemsetAsync(e->d_sparse.p, 0, o_cum, e->stream));
    for (int ch = 0; ch < nch; ch++) {
      const int k = sp_of[ch];
      if (k < 0) continue;
      SparseJob &j = SJ.j[k];
      j.n = n; j.mean = mean[ch];
      j.used = e->d_sparse.p + o_used + kRemapUsedBytes * k;
      j.cum = reinterpret_cast<int32_t *>(e->d_sparse.p + o_cum + cumb * k);
      j.ulist = reinterpret_cast<int32_t *>(e->d_sparse.p + o_list + cumb * k);
      j.bytes = const_cast<uint8_t *>(e->h_bpjobs.p[ch].in); j.in_len = plen[ch];
      j.rc = reinterpret_cast<RcInit *>(e->d_sparse.p) + k;
      e->h_bpjobs.p[ch].rc_init = j.rc;
    }
    SACB_CUDA(launch_sparse_decode(e->bt, SJ, e->stream));
    e->launches += 2; e->last_launches[2] += 2;
  }
  SACB_CUDA(cudaMemcpyAsync(e->d_bpjobs.p, e->h_bpjobs.p, sizeof(BpJob) * nch, cudaMemcpyHostToDevice, e->stream));
  SACB_CUDA(cudaEventRecord(e->ev[2], e->stream));
  SACB_CUDA(launch_bitplane(e->bt, e->d_bpjobs.p, nch, 2, e->stream));
  SACB_CUDA(cudaEventRecord(e->ev[3], e->stream));
  e->launches++; e->last_launches[1]++;
  // ---- predictor in decode direction: the two channels are coupled through progress counters ----
  const HostParam hp = map_profile(prof);
  SACB_CUDA(e->h_descs.reserve(nch));
  SACB_CUDA(e->d_descs.reserve(nch));
  SACB_CUDA(e->d_flags.reserve(8));
  SACB_CUDA(e->d_sums.reserve(8));
  SACB_CUDA(cudaMemsetAsync(e->d_flags.p, 0, sizeof(int) * 8, e->stream));
  long long soff[3] = {0, 0, 0};
  for (int cc = 0; cc < nch; cc++) soff[cc + 1] = soff[cc] + ((chain_scratch_doubles(hp, cc, nch) + 1) & ~1LL);
  SACB_CUDA(e->d_scratch.reserve((size_t)soff[nch]));
  int32_t *dplanes[2] = {e->d_resid.p + (size_t)nch * stride, e->d_resid.p + (size_t)(nch + 1) * stride};
  for (int cc = 0; cc < nch; cc++) {
    ChainDesc &d = e->h_descs.p[cc];
    const int32_t *cpl[2] = {dplanes[0], dplanes[1]};
    const int actual = fill_chain(d, hp, nch, cc, 1, cpl, 0, n, mm);
    d.own_out = dplanes[actual];
    d.err_in = e->d_resid.p + (size_t)actual * stride;
    if (sp_of[actual] >= 0) {                                      // UnpredictFrame, libsac.cpp:157-160
      d.unmap_cum = SJ.j[sp_of[actual]].cum; d.unmap_list = SJ.j[sp_of[actual]].ulist; d.unmap_mean = mean[actual];
    }
    d.resid = nullptr;
    d.scratch = e->d_scratch.p + soff[cc]; d.scratch_doubles = soff[cc + 1] - soff[cc];
    d.l1sum = nullptr; d.sqsum = nullptr; d.flags = e->d_flags.p + cc;
    d.progress = e->d_flags.p + 4 + cc;
    d.wait_ctr = nullptr; d.wait_add = 0;
    if (nch == 2 && d.lenB > 0) {
      d.wait_ctr = e->d_flags.p + 4 + (1 - cc);
      d.wait_add = cc == 0 ? -(std::max(hp.nS1, 1) - 1) : hp.nS1;
    }
  }
  SACB_CUDA(cudaMemcpyAsync(e->d_descs.p, e->h_descs.p, sizeof(ChainDesc) * nch, cudaMemcpyHostToDevice, e->stream));
  SACB_CUDA(cudaEventRecord(e->ev[0], e->stream));
  SACB_CUDA(launch_predictor_decode(e->d_descs.p, nch, e->smem_bytes, e->stream));
  SACB_CUDA(cudaEventRecord(e->ev[1], e->stream));
  e->launches++; e->last_launches[0]++;
  for (int ch = 0; ch < nch; ch++) {
    planes[ch].resize(n);
    SACB_CUDA(cudaMemcpyAsync(planes[ch].data(), dplanes[ch], sizeof(int32_t) * n, cudaMemcpyDeviceToHost, e->stream));
  }
  SACB_CUDA(cudaStreamSynchronize(e->stream));
  { float ms = 0; cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]); e->last_ms[0] += ms; cudaEventElapsedTime(&ms, e->ev[2], e->ev[3]); e->last_ms[1] += ms; }
  for (int ch = 0; ch < nch; ch++)
    if (mean[ch] != 0) for (auto &v : planes[ch]) v += mean[ch];     // libsac.cpp:194-198
  *n_out = n;
  return pos;
}

// =====================================================================================================================
// adaptive frame split (Codec::Analyse / PushState, libsac.cpp:695-780; SparsePCM::Analyse, sparse.h:25-70)
// =====================================================================================================================
namespace {
struct SubFrame { int state = -1, start = 0, length = 0; };        // Codec::tsub_frame (libsac.h:87-91)

// SparsePCM::Analyse: ratio of the L1 norm of the values to the L1 norm of their rank distances among the used values
double sparse_cost(const int32_t *buf, int n)
{
  int32_t minval = std::numeric_limits<int32_t>::max(), maxval = std::numeric_limits<int32_t>::min();
  for (int i = 0; i < n; i++) { if (buf[i] > maxval) maxval = buf[i]; if (buf[i] < minval) minval = buf[i]; }
  const int N = (int)((long long)maxval - (long long)minval + 1);
  std::vector<int> used(N, 0), prefix(N + 1, 0);
  for (int i = 0; i < n; i++) used[buf[i] - minval] = 1;
  for (int i = 0; i < N; i++) prefix[i + 1] = prefix[i] + used[i];
  double sum0 = 0, sum1 = 0;
  for (int i = 0; i < n; i++) {
    const int32_t val = buf[i];
    int rank = 0;                                                     // val2rank_fast(val, p = 0)
    if (val != 0) {
      const int tidx = val - minval, pidx = 0 - minval;
      if (val > 0) rank = prefix[tidx + 1] - prefix[std::clamp(pidx + 1, 0, N)];
      else rank = prefix[tidx] - prefix[std::clamp(pidx, 0, N)];
    }
    sum0 += std::fabs((double)val);
    sum1 += std::fabs((double)rank);
  }
  return sum1 > 0 ? sum0 / sum1 : 0;
}

// Adaptive split in three passes over plain data (the reference interleaves them in Codec::Analyse / PushState,
// libsac.cpp:704-780; frame boundaries must come out identical, tests/golden/golden_split.json):
//   1. class of every block of `blocksamples`: sparse (rank-mapped cost ratio above 1.35) or not;
//   2. run-length encoding of the classes;
//   3. runs -> sub-frames. A run becomes a sub-frame when the next run of the other class starts, unless it is shorter than
//      min_len and a sub-frame already exists: then it is added to that sub-frame instead.
// One behaviour of the reference is reproduced on purpose because it decides frame boundaries: when a short run is absorbed
// like that, the reference keeps it as its open run and does not account the blocks of the run that follows -- each of them
// adds the short run's length to the previous sub-frame once more -- and a later run of the short run's own class continues it.
// With the codec's own parameters (blocks and minimum length both 3 s) only the ragged last block of a read can be short, and it
// has no successor, so sub-frame lengths always add up to the samples read there.
struct ClassRun { int state, length, nblocks; };

std::vector<ClassRun> class_runs(const std::vector<std::vector<int32_t>> &samples, int blocksamples, int samples_read)
{
  std::vector<ClassRun> runs;
  for (int done = 0; done < samples_read; done += blocksamples) {
    const int len = std::min(blocksamples, samples_read - done);
    double cost = 0;
    for (const auto &plane : samples) cost += sparse_cost(plane.data() + done, len);
    const int cls = (cost / (double)samples.size()) > 1.35;
    if (!runs.empty() && runs.back().state == cls) { runs.back().length += len; runs.back().nblocks++; }
    else runs.push_back({cls, len, 1});
  }
  return runs;
}

std::vector<SubFrame> analyse_subframes(const std::vector<std::vector<int32_t>> &samples, int blocksamples, int min_len, int samples_read)
{
  std::vector<SubFrame> sub;
  const std::vector<ClassRun> runs = class_runs(samples, blocksamples, samples_read);
  if (runs.empty()) return sub;
  SubFrame open;
  open.state = runs[0].state; open.start = 0; open.length = runs[0].length;
  auto absorbed = [&](const SubFrame &r) { return r.length < min_len && !sub.empty(); };
  for (size_t k = 1; k < runs.size(); k++) {
    const ClassRun &r = runs[k];
    if (r.state == open.state) open.length += r.length;                       // continues a short run that was absorbed before
    else if (absorbed(open)) sub.back().length += open.length * r.nblocks;    // see above: once per block of the following run
    else {
      sub.push_back(open);
      open.start += open.length; open.state = r.state; open.length = r.length;
    }
  }
  if (open.length) { if (absorbed(open)) sub.back().length += open.length; else sub.push_back(open); }
  return sub;
}
} // namespace

// =====================================================================================================================
// files
// =====================================================================================================================
// Everything of Codec::EncodeFile that happens on the host before the first frame is coded (libsac.cpp:782-835): WAV
// header parse, .sac header + metadata + MD5 of the PCM bytes, reads of max_framesize samples split into sub-frames.
struct ContainerPlan {
  WavInfo wi;
  size_t md5pos = 0;
  uint8_t md5[16];
  int max_framesize = 0;
  std::vector<std::vector<int32_t>> store;                           // frame f, channel ch at store[f * nch + ch]
  std::vector<int> ns;                                               // samples per frame
};
static int container_plan(const sac_cfg &cfg, const uint8_t *wav, size_t wav_len, std::vector<uint8_t> &out, ContainerPlan &cp, bool keep_samples)
{
  WavInfo &wi = cp.wi;
  int rc = wav_parse(wav, wav_len, wi);
  if (rc) return rc;
  if (!(wi.bits <= 24 && (wi.nch == 1 || wi.nch == 2)) || wi.blockalign <= 0 || wi.blockalign % wi.nch) {
    set_error("unsupported input format: must be 1-16 bit, mono/stereo, pcm");          // cmdline.cpp:253-262
    return SAC_E_UNSUPPORTED;
  }
  {
    // the sample container must be the 1..3 bytes wav_unpack / wav_pack handle AND the width the decoder rebuilds from the
    // bit depth alone (blockalign = nch * ceil(bits / 8), wav.cpp:126-164): 24 valid bits in a 32-bit container or a depth of
    // 0 bits would otherwise be coded as silence under the MD5 of the real PCM. (The reference prints "error: unknown csize"
    // for such input and writes a file that cannot be restored; here the input is refused.)
    const int cs = wi.blockalign / wi.nch;
    if (wi.bits < 1 || cs < 1 || cs > 3 || cs != (wi.bits + 7) / 8) {
      set_error("unsupported input format: sample container does not match the bit depth (must be 1-16 bit, mono/stereo, pcm)");
      return SAC_E_UNSUPPORTED;
    }
  }
  const int max_framesize = cp.max_framesize = cfg.max_framelen * wi.samplerate;
  if (max_framesize <= 0) { set_error("bad frame length"); return SAC_E_ARG; }
  // ---- .sac header (sac.cpp:15-38) + MD5 ----
  out.clear();
  out.push_back('S'); out.push_back('A'); out.push_back('C'); out.push_back('2');
  push16(out, (uint16_t)wi.nch); push32(out, (uint32_t)wi.samplerate); push16(out, (uint16_t)wi.bits); push32(out, wi.numsamples);
  out.push_back((uint8_t)cfg.max_framelen); out.push_back((uint8_t)SAC_ARITH_CANONICAL);   // byte 17: arithmetic variant (sac_b200.h)
  push32(out, wi.metadatasize());
  pack_metadata(wi, out);
  cp.md5pos = out.size();
  const uint8_t *pcm = wav + wi.data_pos;
  const size_t pcm_bytes = (size_t)wi.numsamples * wi.blockalign;
  Md5 md5;
  md5.update(pcm, pcm_bytes);
  md5.final(cp.md5);
  out.insert(out.end(), cp.md5, cp.md5 + 16);
  // ---- reads of max_framesize samples, each split into sub-frames (Codec::EncodeFile, libsac.cpp:805-835) ----
  cp.store.clear(); cp.ns.clear();
  for (uint32_t first = 0; first < wi.numsamples; first += (uint32_t)max_framesize) {
    const int nread = (int)std::min<uint32_t>((uint32_t)max_framesize, wi.numsamples - first);
    std::vector<std::vector<int32_t>> pl(wi.nch, std::vector<int32_t>(nread));
    wav_unpack(wi, pcm, (int)first, nread, pl);
    std::vector<SubFrame> sub;
    if (cfg.adapt_block) sub = analyse_subframes(pl, wi.samplerate * 3, wi.samplerate * 3, nread);
    else { SubFrame s; s.state = 0; s.start = 0; s.length = nread; sub.push_back(s); }
    for (const SubFrame &s : sub) {
      if (s.length <= 0) continue;
      if (keep_samples)
        for (int ch = 0; ch < wi.nch; ch++) cp.store.emplace_back(pl[ch].begin() + s.start, pl[ch].begin() + s.start + s.length);
      cp.ns.push_back(s.length);
    }
  }
  return SAC_OK;
}

static int encode_image(Engine *e, const sac_cfg &cfg, const uint8_t *wav, size_t wav_len, std::vector<uint8_t> &out, sac_file_stats *st)
{
  const auto t0 = std::chrono::steady_clock::now();
  ContainerPlan cp;
  int rc = container_plan(cfg, wav, wav_len, out, cp, true);
  if (rc) return rc;
  const WavInfo &wi = cp.wi;
  const int nframes = (int)cp.ns.size();
  std::vector<const int32_t *> ptrs((size_t)nframes * wi.nch);
  for (size_t i = 0; i < ptrs.size(); i++) ptrs[i] = cp.store[i].data();
  float prof[kProfileSize];
  for (int i = 0; i < kProfileSize; i++) prof[i] = kBaseProfile[i][2];
  if (nframes > 0) {
    rc = frames_encode(e, cfg, wi.nch, cp.max_framesize, nframes, ptrs.data(), cp.ns.data(), prof, out);
    if (rc) return rc;
  }
  if (st) {
    std::memset(st, 0, sizeof(*st));
    st->in_bytes = (long long)wav_len; st->out_bytes = (long long)out.size(); st->numsamples = (int)wi.numsamples; st->nch = wi.nch;
    st->samplerate = wi.samplerate; st->bits = wi.bits; st->nframes = nframes; std::memcpy(st->md5, cp.md5, 16); st->md5_ok = 1;
    st->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  return SAC_OK;
}

// One file on SEVERAL GPUs of the box, from the C++ host (no collective: frames are independent searches under --opt-reset
// semantics, which this path implies, and records are concatenated in frame order, libsac.cpp:565-578). Frames are dealt
// to the engines round-robin; every engine encodes its share with its frames in flight (frame_parallel = 2) on its own
// host thread. With one engine this is encode_image with reset = 1.
static int encode_image_multi(Engine *const *engs, int neng, const sac_cfg &cfg_in, const uint8_t *wav, size_t wav_len, std::vector<uint8_t> &out,
                              sac_file_stats *st)
{
  const auto t0 = std::chrono::steady_clock::now();
  sac_cfg cfg = cfg_in;
  cfg.reset = 1; cfg.frame_parallel = 2;
  ContainerPlan cp;
  int rc = container_plan(cfg, wav, wav_len, out, cp, true);
  if (rc) return rc;
  const WavInfo &wi = cp.wi;
  const int nframes = (int)cp.ns.size();
  const int used = std::max(1, std::min(neng, nframes));
  std::vector<std::vector<int>> share(used);
  for (int f = 0; f < nframes; f++) share[f % used].push_back(f);
  std::vector<std::vector<uint8_t>> recs(used);
  std::vector<int> rcs(used, SAC_OK);
  std::vector<std::string> errs(used);
  std::vector<std::thread> th;
  for (int g = 0; g < used; g++)
    th.emplace_back([&, g]() {
      Engine *e = engs[g];
      cudaSetDevice(e->device);
      const std::vector<int> &fs = share[g];
      if (fs.empty()) return;
      std::vector<const int32_t *> ptrs;
      std::vector<int> ns;
      for (int f : fs) {
        for (int ch = 0; ch < wi.nch; ch++) ptrs.push_back(cp.store[(size_t)f * wi.nch + ch].data());
        ns.push_back(cp.ns[f]);
      }
      float prof[kProfileSize];
      for (int i = 0; i < kProfileSize; i++) prof[i] = kBaseProfile[i][2];
      rcs[g] = frames_encode(e, cfg, wi.nch, cp.max_framesize, (int)fs.size(), ptrs.data(), ns.data(), prof, recs[g]);
      if (rcs[g]) errs[g] = sac_last_error();
    });
  for (auto &t : th) t.join();
  for (int g = 0; g < used; g++) if (rcs[g]) { set_error(errs[g]); return rcs[g]; }
  // frame f is record number f / used of engine f % used: walk the records (u32 numsamples, 58 f32, per channel 18 B header + payload)
  std::vector<size_t> pos(used, 0);
  for (int f = 0; f < nframes; f++) {
    const std::vector<uint8_t> &r = recs[f % used];
    size_t p = pos[f % used], q = p + 4 + 4 * (size_t)kProfileSize;
    for (int ch = 0; ch < wi.nch; ch++) { if (q + 18 > r.size()) { set_error("internal: short frame record"); return SAC_E_FORMAT; } q += 18 + (size_t)get32(r.data() + q); }
    if (q > r.size()) { set_error("internal: short frame record"); return SAC_E_FORMAT; }
    out.insert(out.end(), r.begin() + (long)p, r.begin() + (long)q);
    pos[f % used] = q;
  }
  if (st) {
    std::memset(st, 0, sizeof(*st));
    st->in_bytes = (long long)wav_len; st->out_bytes = (long long)out.size(); st->numsamples = (int)wi.numsamples; st->nch = wi.nch;
    st->samplerate = wi.samplerate; st->bits = wi.bits; st->nframes = nframes; std::memcpy(st->md5, cp.md5, 16); st->md5_ok = 1;
    st->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  return SAC_OK;
}

static int decode_image(Engine *e, const uint8_t *sac, size_t len, std::vector<uint8_t> &out, sac_file_stats *st)
{
  const auto t0 = std::chrono::steady_clock::now();
  if (len < 22 + 16 || !(sac[0] == 'S' && sac[1] == 'A' && sac[2] == 'C' && sac[3] == '2')) { set_error("input is not a valid .sac file"); return SAC_E_FORMAT; }
  WavInfo wi;
  wi.nch = get16(sac + 4); wi.samplerate = (int)get32(sac + 6); wi.bits = get16(sac + 10); wi.numsamples = get32(sac + 12);
  const int max_framelen = sac[16];
  const uint32_t mdsize = get32(sac + 18);
  if (22 + (size_t)mdsize + 16 > len || wi.nch < 1 || wi.nch > 2 || wi.bits < 1 || wi.bits > 24) { set_error("corrupt .sac header"); return SAC_E_FORMAT; }
  const int arith = sac[17];
  if (arith != SAC_ARITH_REFERENCE && arith != SAC_ARITH_CANONICAL) { set_error("unknown arithmetic variant in the .sac header (byte 17)"); return SAC_E_UNSUPPORTED; }
  {
    const long long mfs = (long long)max_framelen * (long long)wi.samplerate;   // crafted headers must not overflow the frame capacity
    if (max_framelen <= 0 || wi.samplerate <= 0 || mfs > (1LL << 28)) { set_error("corrupt .sac header: frame length out of range"); return SAC_E_FORMAT; }
  }
  if (unpack_metadata(sac + 22, mdsize, wi)) { set_error("unpackmetadata mismatch"); return SAC_E_FORMAT; }
  size_t pos = 22 + mdsize;
  uint8_t want[16];
  std::memcpy(want, sac + pos, 16);
  pos += 16;
  const int csize = (wi.bits + 7) / 8;                                // Wav(AudioFile&) ctor (wav.cpp:61-70)
  wi.blockalign = wi.nch * csize;
  const int max_framesize = max_framelen * wi.samplerate;
  // ---- WAV header chunks up to and including 'data' (Wav::WriteHeader, wav.cpp:265-281) ----
  out.clear();
  size_t chunkpos = 0;
  auto write_header = [&]() {
    while (chunkpos < wi.chunks.size()) {
      const WavChunk &c = wi.chunks[chunkpos++];
      push32(out, c.id); push32(out, c.csize);
      if (c.id == kDATA) break;
      out.insert(out.end(), c.data.begin(), c.data.end());
    }
  };
  write_header();
  Md5 md5;
  long long todo = wi.numsamples, data_bytes = 0;
  int nframes = 0;
  std::vector<std::vector<int32_t>> planes(wi.nch);
  while (todo > 0) {
    int n = 0;
    const long long used = frame_decode(e, wi.nch, sac + pos, (long long)(len - pos), planes, max_framesize, &n);
    if (used < 0) return (int)used;
    pos += (size_t)used;
    const size_t o0 = out.size();
    wav_pack(wi, planes, n, out);
    md5.update(out.data() + o0, out.size() - o0);
    data_bytes += (long long)(out.size() - o0);
    todo -= n;
    nframes++;
  }
  if (data_bytes & 1) out.push_back(0);                               // libsac.cpp:880-881
  write_header();
  uint8_t dig[16];
  md5.final(dig);
  if (st) {
    std::memset(st, 0, sizeof(*st));
    st->in_bytes = (long long)len; st->out_bytes = (long long)out.size(); st->numsamples = (int)wi.numsamples; st->nch = wi.nch;
    st->samplerate = wi.samplerate; st->bits = wi.bits; st->nframes = nframes; std::memcpy(st->md5, dig, 16);
    st->md5_ok = std::memcmp(dig, want, 16) == 0;
    st->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  if (std::memcmp(dig, want, 16) != 0) {
    // never hand out audio that is not the encoder's input (a lossless archiver must not corrupt silently)
    set_error(arith == SAC_ARITH_REFERENCE
                  ? "audio MD5 mismatch: the file was written by a build of the reference (arithmetic variant 0) whose fp64 predictor "
                    "arithmetic differs from this decoder's canonical arithmetic; decode it with the build that wrote it"
                  : "audio MD5 mismatch: the stream is damaged");
    return SAC_E_MD5;
  }
  return SAC_OK;
}

static int read_file(const char *path, std::vector<uint8_t> &buf)
{
  std::ifstream f(path, std::ios::binary);
  if (!f) { set_error(std::string("could not open ") + path); return SAC_E_IO; }
  f.seekg(0, std::ios::end);
  const std::streamoff sz = f.tellg();
  f.seekg(0);
  buf.resize((size_t)sz);
  if (sz) f.read(reinterpret_cast<char *>(buf.data()), sz);
  return f ? SAC_OK : SAC_E_IO;
}
static int write_file(const char *path, const std::vector<uint8_t> &buf)
{
  std::ofstream f(path, std::ios::binary);
  if (!f) { set_error(std::string("could not create ") + path); return SAC_E_IO; }
  f.write(reinterpret_cast<const char *>(buf.data()), (std::streamsize)buf.size());
  return f ? SAC_OK : SAC_E_IO;
}

} // namespace sacb

using namespace sacb;

extern "C" {

double sac_dds_run(int D, const double *xmin, const double *xmax, const double *xstart, int nfunc_max, int num_threads,
                   double sigma_init, sac_eval_fn eval, void *user, double *xbest)
{
  if (D <= 0 || !xmin || !xmax || !xstart || !eval || nfunc_max < 1) { set_error("sac_dds_run: bad argument"); return std::numeric_limits<double>::quiet_NaN(); }
  DdsSearch s(D, xmin, xmax, xstart, nfunc_max, num_threads, sigma_init);
  std::vector<std::vector<double>> cands;
  std::vector<double> X, cost;
  do {
    s.propose(cands);
    X.clear();
    for (auto &c : cands) X.insert(X.end(), c.begin(), c.end());
    cost.assign(cands.size(), 0.0);
    if (eval(X.data(), (int)cands.size(), D, cost.data(), user)) { set_error("sac_dds_run: evaluator aborted"); break; }
    s.consume(cost.data());
  } while (!s.done());
  if (xbest) std::copy(s.best_x().begin(), s.best_x().end(), xbest);
  return s.best_cost();
}

double sac_dds_run_spec(int D, const double *xmin, const double *xmax, const double *xstart, int nfunc_max, double sigma_init, int spec,
                        sac_eval_fn eval, void *user, double *xbest, long long *evaluated)
{
  if (D <= 0 || !xmin || !xmax || !xstart || !eval || nfunc_max < 1) { set_error("sac_dds_run_spec: bad argument"); return std::numeric_limits<double>::quiet_NaN(); }
  DdsSearch s(D, xmin, xmax, xstart, nfunc_max, 0, sigma_init, spec);
  std::vector<std::vector<double>> cands;
  std::vector<double> X, cost;
  do {
    s.propose(cands);
    if (cands.empty()) break;
    X.clear();
    for (auto &c : cands) X.insert(X.end(), c.begin(), c.end());
    cost.assign(cands.size(), 0.0);
    if (eval(X.data(), (int)cands.size(), D, cost.data(), user)) { set_error("sac_dds_run_spec: evaluator aborted"); break; }
    s.consume(cost.data());
  } while (!s.done());
  if (xbest) std::copy(s.best_x().begin(), s.best_x().end(), xbest);
  if (evaluated) *evaluated = s.evaluated();
  return s.best_cost();
}

double sac_de_run(int D, const double *xmin, const double *xmax, const double *xstart, int nfunc_max, double sigma_init, sac_eval_fn eval,
                  void *user, double *xbest)
{
  if (D <= 0 || !xmin || !xmax || !xstart || !eval || nfunc_max < 1) { set_error("sac_de_run: bad argument"); return std::numeric_limits<double>::quiet_NaN(); }
  DeSearch s(D, xmin, xmax, xstart, nfunc_max, sigma_init);
  std::vector<std::vector<double>> cands;
  std::vector<double> X, cost;
  do {
    s.propose(cands);
    if (cands.empty()) break;
    X.clear();
    for (auto &c : cands) X.insert(X.end(), c.begin(), c.end());
    cost.assign(cands.size(), 0.0);
    if (eval(X.data(), (int)cands.size(), D, cost.data(), user)) { set_error("sac_de_run: evaluator aborted"); break; }
    s.consume(cost.data());
  } while (!s.done());
  if (xbest) std::copy(s.best_x().begin(), s.best_x().end(), xbest);
  return s.best_cost();
}

double sac_cma_run(int D, const double *xmin, const double *xmax, const double *xstart, int nfunc_max, double sigma_init, sac_eval_fn eval,
                   void *user, double *xbest)
{
  if (D <= 0 || !xmin || !xmax || !xstart || !eval || nfunc_max < 1) { set_error("sac_cma_run: bad argument"); return std::numeric_limits<double>::quiet_NaN(); }
  CmaSearch s(D, xmin, xmax, xstart, nfunc_max, sigma_init);
  std::vector<std::vector<double>> cands;
  std::vector<double> cost(1);
  do {
    s.propose(cands);
    if (eval(cands[0].data(), 1, D, cost.data(), user)) { set_error("sac_cma_run: evaluator aborted"); break; }
    s.consume(cost.data());
  } while (!s.done());
  if (xbest) std::copy(s.best_x().begin(), s.best_x().end(), xbest);
  return s.best_cost();
}

int sac_analyse_subframes(int nch, const int32_t *const *planes, int numsamples, int samplerate, int *start, int *length, int *state, int cap)
{
  if (nch < 1 || nch > 2 || !planes || numsamples <= 0 || samplerate <= 0) { set_error("sac_analyse_subframes: bad argument"); return SAC_E_ARG; }
  std::vector<std::vector<int32_t>> pl(nch);
  for (int ch = 0; ch < nch; ch++) pl[ch].assign(planes[ch], planes[ch] + numsamples);
  const std::vector<SubFrame> sub = analyse_subframes(pl, samplerate * 3, samplerate * 3, numsamples);
  for (int i = 0; i < (int)sub.size() && i < cap; i++) {
    if (start) start[i] = sub[i].start;
    if (length) length[i] = sub[i].length;
    if (state) state[i] = sub[i].state;
  }
  return (int)sub.size();
}

int sac_bitplane_encode(sac_engine *h, const int32_t *resid, int n, int *maxbpn_io, uint8_t *out, long long cap, long long *out_len)
{
  Engine *e = reinterpret_cast<Engine *>(h);
  if (!e || !resid || n <= 0 || !out_len) { set_error("sac_bitplane_encode: bad argument"); return SAC_E_ARG; }
  SACB_CUDA(cudaSetDevice(e->device));
  e->begin_call();
  const size_t stride = ((size_t)n + 31) & ~size_t(31), bcap = (size_t)n * 4 + 1024;
  SACB_CUDA(e->d_resid.reserve(stride));
  SACB_CUDA(e->d_csig0.reserve(65536));
  SACB_CUDA(e->d_bytes.reserve(bcap));
  SACB_CUDA(e->d_sums.reserve(4));
  SACB_CUDA(e->d_flags.reserve(4));
  SACB_CUDA(e->h_bpjobs.reserve(1));
  SACB_CUDA(e->d_bpjobs.reserve(1));
  SACB_CUDA(cudaMemcpyAsync(e->d_resid.p, resid, sizeof(int32_t) * n, cudaMemcpyHostToDevice, e->stream));
  BpJob &b = e->h_bpjobs.p[0];
  std::memset(&b, 0, sizeof(b));
  b.buf = e->d_resid.p; b.n = n; b.signed_input = 1; b.maxbpn = maxbpn_io ? *maxbpn_io : -1;
  b.csig0 = e->d_csig0.p; b.out = e->d_bytes.p; b.nbytes = e->d_sums.p; b.maxbpn_out = e->d_flags.p;
  SACB_CUDA(cudaMemcpyAsync(e->d_bpjobs.p, e->h_bpjobs.p, sizeof(BpJob), cudaMemcpyHostToDevice, e->stream));
  SACB_CUDA(cudaEventRecord(e->ev[2], e->stream));
  SACB_CUDA(launch_bitplane(e->bt, e->d_bpjobs.p, 1, 1, e->stream));
  SACB_CUDA(cudaEventRecord(e->ev[3], e->stream));
  e->launches++; e->last_launches[1]++;
  long long nb = 0; int mb = 0;
  SACB_CUDA(cudaMemcpyAsync(&nb, e->d_sums.p, sizeof(nb), cudaMemcpyDeviceToHost, e->stream));
  SACB_CUDA(cudaMemcpyAsync(&mb, e->d_flags.p, sizeof(mb), cudaMemcpyDeviceToHost, e->stream));
  SACB_CUDA(cudaStreamSynchronize(e->stream));
  { float ms = 0; cudaEventElapsedTime(&ms, e->ev[2], e->ev[3]); e->last_ms[1] += ms; }
  *out_len = nb;
  if (maxbpn_io) *maxbpn_io = mb;
  if (!out || nb > cap) { set_error("output buffer too small"); return SAC_E_ARG; }
  SACB_CUDA(cudaMemcpy(out, e->d_bytes.p, (size_t)nb, cudaMemcpyDeviceToHost));
  return SAC_OK;
}

int sac_bitplane_decode(sac_engine *h, const uint8_t *payload, long long len, int n, int maxbpn, int32_t *resid_out)
{
  Engine *e = reinterpret_cast<Engine *>(h);
  if (!e || !payload || len < 0 || n <= 0 || maxbpn < 0 || maxbpn > 30 || !resid_out) { set_error("sac_bitplane_decode: bad argument"); return SAC_E_ARG; }
  SACB_CUDA(cudaSetDevice(e->device));
  e->begin_call();
  const size_t stride = ((size_t)n + 31) & ~size_t(31), pl = ((size_t)len + 15) & ~size_t(15);
  SACB_CUDA(e->d_resid.reserve(stride));
  SACB_CUDA(e->d_csig0.reserve(65536));
  SACB_CUDA(e->d_bytes.reserve(pl + stride + 16));
  SACB_CUDA(e->h_bpjobs.reserve(1));
  SACB_CUDA(e->d_bpjobs.reserve(1));
  if (len) SACB_CUDA(cudaMemcpyAsync(e->d_bytes.p, payload, (size_t)len, cudaMemcpyHostToDevice, e->stream));
  BpJob &b = e->h_bpjobs.p[0];
  std::memset(&b, 0, sizeof(b));
  b.buf = e->d_resid.p; b.n = n; b.maxbpn = maxbpn; b.csig0 = e->d_csig0.p; b.in = e->d_bytes.p; b.in_len = len; b.msb = e->d_bytes.p + pl;
  SACB_CUDA(cudaMemcpyAsync(e->d_bpjobs.p, e->h_bpjobs.p, sizeof(BpJob), cudaMemcpyHostToDevice, e->stream));
  SACB_CUDA(cudaEventRecord(e->ev[2], e->stream));
  SACB_CUDA(launch_bitplane(e->bt, e->d_bpjobs.p, 1, 2, e->stream));
  SACB_CUDA(cudaEventRecord(e->ev[3], e->stream));
  e->launches++; e->last_launches[1]++;
  SACB_CUDA(cudaMemcpyAsync(resid_out, e->d_resid.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, e->stream));
  SACB_CUDA(cudaStreamSynchronize(e->stream));
  { float ms = 0; cudaEventElapsedTime(&ms, e->ev[2], e->ev[3]); e->last_ms[1] += ms; }
  return SAC_OK;
}

void sac_cfg_default(sac_cfg *c)
{
  std::memset(c, 0, sizeof(*c));
  c->optimize = 0; c->fraction = 0; c->maxnfunc = 0; c->num_threads = 0; c->sigma = 0.2; c->optk = 4;
  c->cost_kind = SAC_COST_ENTROPY; c->reset = 0; c->zero_mean = 1; c->sparse_pcm = 1; c->max_framelen = 20; c->adapt_block = 1;
  c->frame_parallel = 0; c->verbose = 0; c->search = SAC_SEARCH_DDS;
  c->spec = 16; c->inflight = 4; c->grade = 0;
}
int sac_cfg_preset(sac_cfg *c, const char *name)                        // cmdline.cpp:127-156
{
  std::string s(name ? name : "");
  for (auto &ch : s) ch = (char)std::tolower(ch);
  if (s == "normal") { c->optimize = 0; }
  else if (s == "high") { c->optimize = 1; c->fraction = 0.1; c->maxnfunc = 100; c->sigma = 0.20; }
  else if (s == "veryhigh") { c->optimize = 1; c->fraction = 0.2; c->maxnfunc = 300; c->sigma = 0.25; }
  else if (s == "extrahigh") { c->optimize = 1; c->fraction = 0.2; c->maxnfunc = 600; c->sigma = 0.25; }
  else if (s == "best") { c->optimize = 1; c->fraction = 0.50; c->maxnfunc = 1000; c->sigma = 0.25; c->cost_kind = SAC_COST_BITPLANE; }
  else if (s == "insane") { c->optimize = 1; c->fraction = 0.50; c->maxnfunc = 1500; c->sigma = 0.25; c->cost_kind = SAC_COST_BITPLANE; }
  else { set_error("unknown preset"); return SAC_E_ARG; }
  return SAC_OK;
}

int sac_frames_encode(sac_engine *h, const sac_cfg *cfg, int nch, int max_framesize, int nframes, const int32_t *const *planes,
                      const int *numsamples, float *profile_io, uint8_t *out, long long cap, long long *out_len)
{
  Engine *e = reinterpret_cast<Engine *>(h);
  if (!e || !cfg || nch < 1 || nch > 2 || nframes <= 0 || !planes || !numsamples || !profile_io || !out_len) { set_error("sac_frames_encode: bad argument"); return SAC_E_ARG; }
  std::vector<uint8_t> buf;
  int rc = frames_encode(e, *cfg, nch, max_framesize, nframes, planes, numsamples, profile_io, buf);
  if (rc) return rc;
  *out_len = (long long)buf.size();
  if ((long long)buf.size() > cap || !out) { set_error("output buffer too small"); return SAC_E_ARG; }
  std::memcpy(out, buf.data(), buf.size());
  return SAC_OK;
}

int sac_frames_encode_resident(sac_engine *h, const sac_cfg *cfg, int nch, int max_framesize, int nframes, const sac_window *const *wins,
                               const int32_t *means, float *profile_io, uint8_t *out, long long cap, long long *out_len)
{
  Engine *e = reinterpret_cast<Engine *>(h);
  if (!e || !cfg || nch < 1 || nch > 2 || nframes <= 0 || !wins || !profile_io || !out_len) { set_error("sac_frames_encode_resident: bad argument"); return SAC_E_ARG; }
  std::vector<uint8_t> buf;
  int rc = frames_encode(e, *cfg, nch, max_framesize, nframes, nullptr, nullptr, profile_io, buf, wins, means);
  if (rc) return rc;
  *out_len = (long long)buf.size();
  if ((long long)buf.size() > cap || !out) { set_error("output buffer too small"); return SAC_E_ARG; }
  std::memcpy(out, buf.data(), buf.size());
  return SAC_OK;
}

long long sac_frame_decode(sac_engine *h, int nch, const uint8_t *in, long long len, int32_t *const *planes_out, int cap_samples, int *numsamples)
{
  Engine *e = reinterpret_cast<Engine *>(h);
  if (!e || nch < 1 || nch > 2 || !in || !planes_out || !numsamples) { set_error("sac_frame_decode: bad argument"); return SAC_E_ARG; }
  std::vector<std::vector<int32_t>> planes(nch);
  int n = 0;
  const long long used = frame_decode(e, nch, in, len, planes, cap_samples, &n);
  if (used < 0) return used;
  for (int ch = 0; ch < nch; ch++) std::memcpy(planes_out[ch], planes[ch].data(), sizeof(int32_t) * (size_t)n);
  *numsamples = n;
  return used;
}

int sac_frame_stats(const int32_t *samples, int n, int zero_mean, int32_t *out3)
{
  if (!samples || n <= 0 || !out3) { set_error("sac_frame_stats: bad argument"); return SAC_E_ARG; }
  std::vector<int32_t> s(samples, samples + n);
  analyse_channel(s, zero_mean, out3[0], out3[1], out3[2]);
  return SAC_OK;
}

int sac_container_plan(const sac_cfg *cfg, const uint8_t *wav, long long wav_len, uint8_t *out, long long cap, long long *out_len,
                       int *frame_lengths, int cap_frames, sac_file_stats *st)
{
  if (!cfg || !wav || wav_len <= 0) { set_error("null argument"); return SAC_E_ARG; }
  std::vector<uint8_t> o;
  ContainerPlan cp;
  int rc = container_plan(*cfg, wav, (size_t)wav_len, o, cp, false);
  if (rc) return rc;
  if (out_len) *out_len = (long long)o.size();
  if (out) { if ((long long)o.size() > cap) { set_error("output buffer too small"); return SAC_E_ARG; } std::memcpy(out, o.data(), o.size()); }
  const int nframes = (int)cp.ns.size();
  if (frame_lengths) { if (nframes > cap_frames) { set_error("frame_lengths too small"); return SAC_E_ARG; } std::copy(cp.ns.begin(), cp.ns.end(), frame_lengths); }
  if (st) {
    std::memset(st, 0, sizeof(*st));
    st->in_bytes = wav_len; st->out_bytes = (long long)o.size(); st->numsamples = (int)cp.wi.numsamples; st->nch = cp.wi.nch;
    st->samplerate = cp.wi.samplerate; st->bits = cp.wi.bits; st->nframes = nframes; std::memcpy(st->md5, cp.md5, 16); st->md5_ok = 1;
  }
  return SAC_OK;
}

int sac_encode_memory(sac_engine *h, const sac_cfg *cfg, const uint8_t *wav, long long wav_len, uint8_t *out, long long cap,
                      long long *out_len, sac_file_stats *st)
{
  Engine *e = reinterpret_cast<Engine *>(h);
  if (!e || !cfg || !wav || wav_len <= 0 || !out_len) { set_error("sac_encode_memory: bad argument"); return SAC_E_ARG; }
  std::vector<uint8_t> buf;
  int rc = encode_image(e, *cfg, wav, (size_t)wav_len, buf, st);
  if (rc) return rc;
  *out_len = (long long)buf.size();
  if (!out || (long long)buf.size() > cap) { set_error("output buffer too small"); return SAC_E_ARG; }
  std::memcpy(out, buf.data(), buf.size());
  return SAC_OK;
}
int sac_decode_memory(sac_engine *h, const uint8_t *sac, long long sac_len, uint8_t *out, long long cap, long long *out_len, sac_file_stats *st)
{
  Engine *e = reinterpret_cast<Engine *>(h);
  if (!e || !sac || sac_len <= 0 || !out_len) { set_error("sac_decode_memory: bad argument"); return SAC_E_ARG; }
  std::vector<uint8_t> buf;
  int rc = decode_image(e, sac, (size_t)sac_len, buf, st);
  if (rc) { *out_len = 0; return rc; }
  *out_len = (long long)buf.size();
  if (!out || (long long)buf.size() > cap) { set_error("output buffer too small"); return SAC_E_ARG; }
  std::memcpy(out, buf.data(), buf.size());
  return SAC_OK;
}
int sac_encode_file(sac_engine *h, const sac_cfg *cfg, const char *wav_path, const char *sac_path, sac_file_stats *st)
{
  Engine *e = reinterpret_cast<Engine *>(h);
  if (!e || !cfg || !wav_path || !sac_path) { set_error("sac_encode_file: bad argument"); return SAC_E_ARG; }
  std::vector<uint8_t> in, out;
  int rc = read_file(wav_path, in);
  if (rc) return rc;
  rc = encode_image(e, *cfg, in.data(), in.size(), out, st);
  if (rc) return rc;
  return write_file(sac_path, out);
}
int sac_encode_file_multi(sac_engine *const *engines, int nengines, const sac_cfg *cfg, const char *wav_path, const char *sac_path, sac_file_stats *st)
{
  if (!engines || nengines < 1 || !cfg || !wav_path || !sac_path) { set_error("sac_encode_file_multi: bad argument"); return SAC_E_ARG; }
  for (int i = 0; i < nengines; i++) if (!engines[i]) { set_error("sac_encode_file_multi: null engine"); return SAC_E_ARG; }
  std::vector<uint8_t> in, out;
  int rc = read_file(wav_path, in);
  if (rc) return rc;
  rc = encode_image_multi(reinterpret_cast<Engine *const *>(engines), nengines, *cfg, in.data(), in.size(), out, st);
  if (rc) return rc;
  return write_file(sac_path, out);
}

// Codec::EncodeFile over a batch (config C4): files are independent, so every engine (GPU) takes the next file off a queue
// ordered longest first and encodes it with its frames in flight; nothing is exchanged between the GPUs.
int sac_encode_files(sac_engine *const *engines, int nengines, const sac_cfg *cfg, int nfiles, const char *const *wav_paths,
                     const char *const *sac_paths, sac_file_stats *stats, int *status)
{
  if (!engines || nengines < 1 || !cfg || nfiles < 0 || (nfiles && (!wav_paths || !sac_paths))) { set_error("sac_encode_files: bad argument"); return SAC_E_ARG; }
  for (int i = 0; i < nengines; i++) if (!engines[i]) { set_error("sac_encode_files: null engine"); return SAC_E_ARG; }
  for (int i = 0; i < nfiles; i++) if (!wav_paths[i] || !sac_paths[i]) { set_error("sac_encode_files: null path"); return SAC_E_ARG; }
  for (int i = 0; i < nfiles; i++)
    for (int j = 0; j < i; j++)
      if (std::string(sac_paths[i]) == sac_paths[j]) { set_error(std::string("sac_encode_files: two inputs map to the same output ") + sac_paths[i]); return SAC_E_ARG; }
  std::vector<std::pair<long long, int>> order;                      // (size, index), longest first
  for (int i = 0; i < nfiles; i++) {
    std::ifstream f(wav_paths[i], std::ios::binary | std::ios::ate);
    order.push_back({f ? (long long)f.tellg() : -1, i});
  }
  std::sort(order.begin(), order.end(), [](const std::pair<long long, int> &a, const std::pair<long long, int> &b) { return a.first != b.first ? a.first > b.first : a.second < b.second; });
  std::vector<int> rcs(nfiles, SAC_OK);
  std::vector<std::string> errs(nfiles);
  std::atomic<int> next{0};
  std::vector<std::thread> th;
  sac_cfg c = *cfg;
  if (c.frame_parallel == 0) { c.frame_parallel = 2; c.reset = 1; }   // frames of a file in flight together
  for (int g = 0; g < std::min(nengines, std::max(nfiles, 1)); g++)
    th.emplace_back([&, g]() {
      Engine *e = reinterpret_cast<Engine *>(engines[g]);
      cudaSetDevice(e->device);
      for (;;) {
        const int k = next.fetch_add(1);
        if (k >= nfiles) break;
        const int i = order[k].second;
        rcs[i] = sac_encode_file(engines[g], &c, wav_paths[i], sac_paths[i], stats ? stats + i : nullptr);
        if (rcs[i]) errs[i] = sac_last_error();
      }
    });
  for (auto &t : th) t.join();
  int first_bad = SAC_OK;
  for (int i = 0; i < nfiles; i++) {
    if (status) status[i] = rcs[i];
    if (rcs[i] && first_bad == SAC_OK) { first_bad = rcs[i]; set_error(std::string(wav_paths[i]) + ": " + errs[i]); }
  }
  return first_bad;
}

int sac_decode_file(sac_engine *h, const char *sac_path, const char *wav_path, sac_file_stats *st)
{
  Engine *e = reinterpret_cast<Engine *>(h);
  if (!e || !sac_path || !wav_path) { set_error("sac_decode_file: bad argument"); return SAC_E_ARG; }
  std::vector<uint8_t> in, out;
  int rc = read_file(sac_path, in);
  if (rc) return rc;
  rc = decode_image(e, in.data(), in.size(), out, st);
  if (rc) return rc;
  return write_file(wav_path, out);
}

} // extern "C"
