// TEST INFRASTRUCTURE ONLY -- never linked into, loaded by or executed from the product path.
//
// Function-level oracle: a thin extern "C" shim over the reference's own classes, compiled together with the
// reference sources where they lie under /root/reference (see oracle/Makefile; output oracle/_ref/libsacref*.so).
// No reference code is copied here: the shim only constructs the reference's objects and calls their methods.
// FrameCoder's hot-path members (PredictFrame, Optimize, SetParam -- src/libsac/libsac.h:58-77) are private, so
// the reference headers are included with `private` opened up; std headers are pulled in first so that only the
// reference's own declarations are affected.
#include <algorithm>
#include <array>
#include <atomic>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <format>
#include <fstream>
#include <functional>
#include <future>
#include <iomanip>
#include <iostream>
#include <memory>
#include <numeric>
#include <random>
#include <span>
#include <sstream>
#include <string>
#include <thread>
#include <variant>
#include <vector>
#include <immintrin.h>

#define private public
#define protected public
#include "libsac/libsac.h"
#include "libsac/pred.h"
#include "libsac/cost.h"
#include "libsac/vle.h"
#include "opt/dds.h"
#include "opt/de.h"
#include "opt/cma.h"
#include "opt/ssc.h"
#undef private
#undef protected

extern "C" {

// ---- profile (src/libsac/profile.cpp:3-89) -------------------------------------------------------------------
int ref_base_profile(float *vmin, float *vmax, float *vdef)
{
  SacProfile p;
  int n = p.LoadBaseProfile();
  for (int i = 0; i < n; i++) { vmin[i] = p.coefs[i].vmin; vmax[i] = p.coefs[i].vmax; vdef[i] = p.coefs[i].vdef; }
  return n;
}

// ---- FrameCoder (src/libsac/libsac.cpp) -----------------------------------------------------------------------
struct RefFrame {
  FrameCoder::tsac_cfg cfg;
  std::unique_ptr<FrameCoder> fc;
  int nch, framesize;
};

// cost_kind: FrameCoder::SearchCost order L1,RMS,Entropy,Golomb,Bitplane (libsac.h:14)
void *ref_frame_new(int nch, int framesize, int optimize, double fraction, int maxnfunc, int num_threads,
                    double sigma, int optk, int cost_kind, int reset, int sparse_pcm, int zero_mean)
{
  auto *h = new RefFrame();
  h->nch = nch; h->framesize = framesize;
  auto &c = h->cfg;
  c.optimize = optimize; c.sparse_pcm = sparse_pcm; c.zero_mean = zero_mean; c.mt_mode = 0;
  c.ocfg.fraction = fraction; c.ocfg.maxnfunc = maxnfunc; c.ocfg.num_threads = num_threads;
  c.ocfg.sigma = sigma; c.ocfg.optk = optk; c.ocfg.reset = reset;
  c.ocfg.optimize_cost = static_cast<FrameCoder::SearchCost>(cost_kind);
  c.ocfg.dds_cfg.nfunc_max = maxnfunc;      // as CmdLine::Parse does (src/cmdline.cpp:222-226)
  c.ocfg.dds_cfg.num_threads = num_threads;
  c.ocfg.dds_cfg.sigma_init = sigma;
  h->fc = std::make_unique<FrameCoder>(nch, framesize, c);
  return h;
}
void ref_frame_free(void *hp) { delete static_cast<RefFrame *>(hp); }

void ref_frame_set_samples(void *hp, int ch, const int32_t *src, int n)
{
  auto *h = static_cast<RefFrame *>(hp);
  std::copy_n(src, n, h->fc->samples[ch].begin());
  h->fc->SetNumSamples(n);
}
void ref_frame_set_mt(void *hp, int mt_mode) { static_cast<RefFrame *>(hp)->fc->cfg.mt_mode = mt_mode; }

// mean/min/max + zero-mean, exactly the prologue of FrameCoder::Predict (libsac.cpp:445-459) without the search
void ref_frame_analyse(void *hp)
{
  auto *h = static_cast<RefFrame *>(hp);
  FrameCoder &f = *h->fc;
  for (int ch = 0; ch < h->nch; ch++) {
    f.AnalyseMonoChannel(ch, f.numsamples_);
    if (f.cfg.zero_mean == 0) f.framestats[ch].mean = 0;
    else if (f.framestats[ch].mean != 0) {
      for (int i = 0; i < f.numsamples_; i++) f.samples[ch][i] -= f.framestats[ch].mean;
      f.framestats[ch].minval -= f.framestats[ch].mean;
      f.framestats[ch].maxval -= f.framestats[ch].mean;
    }
  }
}
void ref_frame_get_stats(void *hp, int ch, int32_t *out /*mean,min,max,maxbpn,blocksize,enc_mapped*/)
{
  auto &s = static_cast<RefFrame *>(hp)->fc->framestats[ch];
  out[0] = s.mean; out[1] = s.minval; out[2] = s.maxval; out[3] = s.maxbpn; out[4] = s.blocksize; out[5] = s.enc_mapped;
}
void ref_frame_set_stats(void *hp, int ch, int32_t mean, int32_t minval, int32_t maxval)
{
  auto &s = static_cast<RefFrame *>(hp)->fc->framestats[ch];
  s.mean = mean; s.minval = minval; s.maxval = maxval;
}
void ref_frame_set_profile(void *hp, const float *vdef)
{
  auto &p = static_cast<RefFrame *>(hp)->fc->base_profile;
  for (size_t i = 0; i < p.coefs.size(); i++) p.coefs[i].vdef = vdef[i];
}
void ref_frame_get_profile(void *hp, float *vdef)
{
  auto &p = static_cast<RefFrame *>(hp)->fc->base_profile;
  for (size_t i = 0; i < p.coefs.size(); i++) vdef[i] = p.coefs[i].vdef;
}

// FrameCoder::PredictFrame (libsac.cpp:94-142) on window [from, from+n) with the frame's current base profile
// overridden by vdef (58 floats). optimize!=0 -> k=optk and no pred[] write. Residuals to e[ch][n].
void ref_frame_predict_window(void *hp, const float *vdef, int from, int n, int optimize, int32_t *e0, int32_t *e1)
{
  auto *h = static_cast<RefFrame *>(hp);
  FrameCoder &f = *h->fc;
  SacProfile prof = f.base_profile;
  for (size_t i = 0; i < prof.coefs.size(); i++) prof.coefs[i].vdef = vdef[i];
  FrameCoder::tch_samples err(h->nch, std::vector<int32_t>(n));
  f.PredictFrame(prof, err, from, n, optimize != 0);
  std::copy_n(err[0].begin(), n, e0);
  if (h->nch > 1) std::copy_n(err[1].begin(), n, e1);
}

// full FrameCoder::Predict() (analysis + DDS search + final pass + S2U) and Encode()
void ref_frame_predict(void *hp) { static_cast<RefFrame *>(hp)->fc->Predict(); }
void ref_frame_encode(void *hp)
{
  auto *h = static_cast<RefFrame *>(hp);
  h->fc->Encode();
  for (int ch = 0; ch < h->nch; ch++) h->fc->framestats[ch].blocksize = h->fc->encoded[ch].GetBufPos();
}
int ref_frame_get_error(void *hp, int ch, int32_t *dst)
{
  auto *h = static_cast<RefFrame *>(hp);
  int n = h->fc->numsamples_;
  std::copy_n(h->fc->error[ch].begin(), n, dst);
  return n;
}
int ref_frame_get_encoded(void *hp, int ch, uint8_t *dst, int cap)
{
  auto *h = static_cast<RefFrame *>(hp);
  int n = h->fc->encoded[ch].GetBufPos();
  if (dst && n <= cap) std::copy_n(h->fc->encoded[ch].GetBuf().begin(), n, dst);
  return n;
}

// sparse-PCM side of a frame (libsac.cpp:214-298, map.cpp): maxbpn of the rank-mapped residuals; loading a coded block
// back and running the reference's Decode() + Unpredict() on it
int ref_frame_get_maxbpn_map(void *hp, int ch) { return static_cast<RefFrame *>(hp)->fc->framestats[ch].maxbpn_map; }
void ref_frame_load_block(void *hp, int ch, const uint8_t *payload, int nbytes, int32_t mean, int32_t minval,
                          int32_t maxval, int maxbpn, int enc_mapped)
{
  auto *h = static_cast<RefFrame *>(hp);
  auto &st = h->fc->framestats[ch];
  st.mean = mean; st.minval = minval; st.maxval = maxval; st.maxbpn = maxbpn; st.enc_mapped = enc_mapped != 0; st.blocksize = nbytes;
  auto &b = h->fc->encoded[ch].GetBuf();
  if ((int)b.size() < nbytes + 8) b.resize(nbytes + 8);
  std::copy_n(payload, nbytes, b.begin());
}
void ref_frame_decode(void *hp, int n)
{
  auto *h = static_cast<RefFrame *>(hp);
  h->fc->SetNumSamples(n);
  h->fc->Decode();
  h->fc->Unpredict();
}
int ref_frame_get_samples(void *hp, int ch, int32_t *dst)
{
  auto *h = static_cast<RefFrame *>(hp);
  int n = h->fc->numsamples_;
  std::copy_n(h->fc->samples[ch].begin(), n, dst);
  return n;
}
// MapEncoder alone on the used-value map of `raw` (Remap::Analyse, map.cpp:126-157; MapEncoder::Encode, map.cpp:76-89)
int ref_map_encode(const int32_t *raw, int n, uint8_t *out, int cap)
{
  Remap m; m.Reset();
  std::vector<int32_t> tmp(raw, raw + n);
  m.Analyse(tmp.data(), n);
  BufIO buf;
  RangeCoderSH rc(buf);
  rc.Init();
  MapEncoder me(rc, m.usedl, m.usedh);
  me.Encode();
  rc.Stop();
  int nb = (int)buf.GetBufPos();
  if (out && nb <= cap) std::copy_n(buf.GetBuf().begin(), nb, out);
  return nb;
}
void ref_remap(const int32_t *raw, int nraw, const int32_t *pred, const int32_t *err, int n, int unmap, int32_t *out)
{
  Remap m; m.Reset();
  std::vector<int32_t> tmp(raw, raw + nraw);
  m.Analyse(tmp.data(), nraw);
  for (int i = 0; i < n; i++) out[i] = unmap ? m.Unmap(pred[i], err[i]) : m.Map(pred[i], err[i]);
}

// FrameCoder::SetParam (libsac.cpp:37-92) flattened in the order of sac_profile_params (include/sac_b200.h)
int ref_set_param(void *hp, const float *vdef, double *out)
{
  auto *h = static_cast<RefFrame *>(hp);
  SacProfile prof = h->fc->base_profile;
  for (size_t i = 0; i < prof.coefs.size(); i++) prof.coefs[i].vdef = vdef[i];
  Predictor::tparam p;
  h->fc->SetParam(p, prof, false);
  int o = 0;
  const int ints[8] = {p.nA, p.nB, p.nM0, p.nS0, p.nS1, p.ch_ref, p.lm_n, p.bias_scale0};
  for (int v : ints) out[o++] = v;
  for (int i = 0; i < 4; i++) out[o++] = p.vn0[i];
  for (int i = 0; i < 4; i++) out[o++] = p.vn1[i];
  for (int i = 0; i < 4; i++) out[o++] = p.vmu0[i];
  for (int i = 0; i < 4; i++) out[o++] = p.vmu1[i];
  for (int i = 0; i < 4; i++) out[o++] = p.vmudecay0[i];
  for (int i = 0; i < 4; i++) out[o++] = p.vmudecay1[i];
  for (int i = 0; i < 4; i++) out[o++] = p.vpowdecay0[i];
  for (int i = 0; i < 4; i++) out[o++] = p.vpowdecay1[i];
  out[o++] = p.lambda0; out[o++] = p.lambda1; out[o++] = p.ols_nu0; out[o++] = p.ols_nu1;
  out[o++] = p.mu_mix0; out[o++] = p.mu_mix1; out[o++] = p.mu_mix_beta0; out[o++] = p.mu_mix_beta1;
  out[o++] = p.beta_sum0; out[o++] = p.beta_pow0; out[o++] = p.beta_add0;
  out[o++] = p.beta_sum1; out[o++] = p.beta_pow1; out[o++] = p.beta_add1;
  out[o++] = p.bias_mu0; out[o++] = p.bias_mu1; out[o++] = p.lm_alpha; out[o++] = p.proj_alpha0; out[o++] = p.proj_alpha1;
  return o;
}

// ---- cost functions (src/libsac/cost.h) ----------------------------------------------------------------------
double ref_cost(int kind, const int32_t *buf, int n)
{
  std::unique_ptr<CostFunction> c;
  switch (kind) {
    case 0: c = std::make_unique<CostL1>(); break;
    case 1: c = std::make_unique<CostRMS>(); break;
    case 2: c = std::make_unique<CostEntropy>(); break;
    case 3: c = std::make_unique<CostGolomb>(); break;
    case 4: c = std::make_unique<CostBitplane>(); break;
    default: return -1.0;
  }
  return c->Calc(std::span<const int32_t>(buf, static_cast<size_t>(n)));
}

// ---- bitplane coder + range coder (src/libsac/vle.cpp, src/model/range.cpp) -----------------------------------
int ref_bitplane_encode(const int32_t *ubuf, int n, int maxbpn, uint8_t *out, int cap)
{
  std::vector<int32_t> tmp(ubuf, ubuf + n);
  BufIO io;
  RangeCoderSH rc(io);
  rc.Init();
  BitplaneCoder bc(maxbpn, n);
  bc.Encode(rc.encode_p1, tmp.data());
  rc.Stop();
  int nb = io.GetBufPos();
  if (out && nb <= cap) std::copy_n(io.GetBuf().begin(), nb, out);
  return nb;
}
// residuals come back signed (BitplaneCoder::Decode applies U2S, vle.cpp:260)
void ref_bitplane_decode(const uint8_t *in, int nbytes, int n, int maxbpn, int32_t *out)
{
  BufIO io(nbytes + 16);
  std::copy_n(in, nbytes, io.GetBuf().begin());
  io.Reset();
  RangeCoderSH rc(io, 1);
  rc.Init();
  BitplaneCoder bc(maxbpn, n);
  bc.Decode(rc.decode_p1, out);
}
// the per-decision probability trace (p1, bit) of an encode -- for model parity without the coder
int ref_bitplane_trace(const int32_t *ubuf, int n, int maxbpn, uint16_t *p1s, uint8_t *bits, long cap)
{
  std::vector<int32_t> tmp(ubuf, ubuf + n);
  long cnt = 0;
  BitplaneCoder bc(maxbpn, n);
  bc.Encode([&](uint32_t p1, int bit) { if (cnt < cap) { p1s[cnt] = p1; bits[cnt] = bit; } cnt++; }, tmp.data());
  return cnt <= cap ? (int)cnt : -1;
}
// BitplaneCoder::PredictLaplace (vle.cpp:70-79) as a pure function of (avg_sum, bpn)
int ref_predict_laplace(uint32_t avg_sum, int bpn)
{
  BitplaneCoder bc(1, 1);
  bc.bpn = bpn;
  return bc.PredictLaplace(avg_sum);
}
int ref_range_encode(const uint16_t *p1s, const uint8_t *bits, int n, uint8_t *out, int cap)
{
  BufIO io;
  RangeCoderSH rc(io);
  rc.Init();
  for (int i = 0; i < n; i++) rc.EncodeBitOne(p1s[i], bits[i]);
  rc.Stop();
  int nb = io.GetBufPos();
  if (out && nb <= cap) std::copy_n(io.GetBuf().begin(), nb, out);
  return nb;
}
void ref_logdomain_tables(int *fwd /*[32768]*/, int *inv /*[4095]*/)
{
  for (int i = 0; i < PSCALE; i++) fwd[i] = myDomain.Fwd(i);
  for (int i = LogDomain::dmin; i <= LogDomain::dmax; i++) inv[i - LogDomain::dmin] = myDomain.Inv(i);
}

// ---- DDS (src/opt/dds.cpp) -----------------------------------------------------------------------------------
typedef double (*ref_cost_cb)(const double *x, int n, void *user);
// runs OptDDS::run; trace (if non-null) receives every evaluated vector in call order [nfunc_max][ndim]
double ref_dds_run(int ndim, const double *xmin, const double *xmax, const double *xstart, int nfunc_max,
                   int num_threads, double sigma_init, ref_cost_cb cb, void *user, double *xbest)
{
  Opt::box_const pb(ndim);
  for (int i = 0; i < ndim; i++) { pb[i].xmin = xmin[i]; pb[i].xmax = xmax[i]; }
  OptDDS::DDSCfg cfg;
  cfg.nfunc_max = nfunc_max; cfg.num_threads = num_threads; cfg.sigma_init = sigma_init;
  OptDDS dds(cfg, pb, false);
  vec1D xs(xstart, xstart + ndim);
  Opt::ppoint r = dds.run([&](const vec1D &x) { return cb(x.data(), ndim, user); }, xs);
  std::copy_n(r.second.begin(), ndim, xbest);
  return r.first;
}
// the candidate generator alone: n candidates around x for nfunc = nfunc0.., fresh OptDDS (seed 0)
void ref_dds_candidates(int ndim, const double *xmin, const double *xmax, const double *x, int nfunc_max, int nfunc0,
                        int count, double sigma, double *out)
{
  Opt::box_const pb(ndim);
  for (int i = 0; i < ndim; i++) { pb[i].xmin = xmin[i]; pb[i].xmax = xmax[i]; }
  OptDDS::DDSCfg cfg;
  cfg.nfunc_max = nfunc_max;
  OptDDS dds(cfg, pb, false);
  vec1D xv(x, x + ndim);
  for (int c = 0; c < count; c++) {
    vec1D cand = dds.generate_candidate(xv, nfunc0 + c, sigma);
    std::copy_n(cand.begin(), ndim, out + (size_t)c * ndim);
  }
}
// ---- DE (src/opt/de.cpp), one evaluation thread so that the trace is in generation order --------------------------
double ref_de_run(int ndim, const double *xmin, const double *xmax, const double *xstart, int nfunc_max, double sigma_init,
                  ref_cost_cb cb, void *user, double *xbest)
{
  Opt::box_const pb(ndim);
  for (int i = 0; i < ndim; i++) { pb[i].xmin = xmin[i]; pb[i].xmax = xmax[i]; }
  OptDE::DECfg cfg;                              // NP 30, CR .5, F .5, c .1, CURPBEST, INIT_NORM, pbest .1 (de.h:18-31)
  cfg.nfunc_max = nfunc_max; cfg.num_threads = 1; cfg.sigma_init = sigma_init;
  OptDE de(cfg, pb, false);
  vec1D xs(xstart, xstart + ndim);
  Opt::ppoint r = de.run([&](const vec1D &x) { return cb(x.data(), ndim, user); }, xs);
  std::copy_n(r.second.begin(), ndim, xbest);
  return r.first;
}
// ---- CMA (src/opt/cma.cpp) -----------------------------------------------------------------------------------------
double ref_cma_run(int ndim, const double *xmin, const double *xmax, const double *xstart, int nfunc_max, double sigma_init,
                   ref_cost_cb cb, void *user, double *xbest)
{
  Opt::box_const pb(ndim);
  for (int i = 0; i < ndim; i++) { pb[i].xmin = xmin[i]; pb[i].xmax = xmax[i]; }
  OptCMA::CMACfg cfg;
  cfg.nfunc_max = nfunc_max; cfg.num_threads = 1; cfg.sigma_init = sigma_init;
  OptCMA cma(cfg, pb, false);
  vec1D xs(xstart, xstart + ndim);
  Opt::ppoint r = cma.run([&](const vec1D &x) { return cb(x.data(), ndim, user); }, xs);
  std::copy_n(r.second.begin(), ndim, xbest);
  return r.first;
}
double ref_ssc0_trace(const int *succ, int n, double sigma, double *out)
{
  SSC0 s(3, 50);
  for (int i = 0; i < n; i++) { sigma = s.update(sigma, succ[i] ? 1.0 : 0.0); out[i] = sigma; }
  return sigma;
}
double ref_ssc1_trace(const double *lambda, int n, double sigma, double *out)
{
  SSC1 s(0.05, 0.10, 0.05);
  for (int i = 0; i < n; i++) { sigma = s.update(sigma, lambda[i]); out[i] = sigma; }
  return sigma;
}

} // extern "C"
