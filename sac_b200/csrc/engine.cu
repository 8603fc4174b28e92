// engine.cu -- per-GPU engine: pools in HBM, descriptor building (FrameCoder::SetParam), batched kernel launches,
// and the compute half of the C ABI (include/sac_b200.h).
#include "engine.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>

namespace sacb {

static thread_local std::string g_err;
static std::atomic<long long> g_dedup_totals[3];                    // logical chains, evaluated chains, evaluated OLS stages
void set_error(const std::string &msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char *what)
{
  set_error(std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what);
  return SAC_E_CUDA;
}

// SacProfile::LoadBaseProfile (profile.cpp:3-89): {vmin, vmax, vdef} per coefficient index
const float kBaseProfile[kProfileSize][3] = {
    {0.99f, 0.9999f, 0.998f}, {1.0f, 100.0f, 25.0f},    {0.001f, 1.0f, 0.1f},      {0.001f, 1.0f, 0.12f},    {0.001f, 1.0f, 0.06f},
    {0.001f, 1.0f, 0.04f},    {0.98f, 1.0f, 1.0f},      {0.0f, 1.0f, 0.8f},        {0.0f, 1.0f, 0.8f},       {0.0f, 32.0f, 0.0f},
    {0.0005f, 0.05f, 0.005f}, {0.8f, 0.9999f, 0.95f},   {0.99f, 0.9999f, 0.998f},  {1.0f, 100.0f, 25.0f},    {0.001f, 1.0f, 0.1f},
    {0.001f, 1.0f, 0.12f},    {0.001f, 1.0f, 0.06f},    {0.001f, 1.0f, 0.04f},     {0.98f, 1.0f, 1.0f},      {0.0f, 1.0f, 0.8f},
    {0.0f, 1.0f, 0.8f},       {0.0f, 1.0f, 0.8f},       {0.0005f, 0.05f, 0.005f},  {0.8f, 0.9999f, 0.95f},   {4.0f, 32.0f, 16.0f},
    {4.0f, 32.0f, 16.0f},     {0.0f, 32.0f, 8.0f},      {-32.0f, 32.0f, 8.0f},     {256.0f, 8192.0f, 1280.0f}, {32.0f, 4096.0f, 256.0f},
    {4.0f, 2048.0f, 32.0f},   {256.0f, 8192.0f, 1280.0f}, {32.0f, 4096.0f, 256.0f}, {4.0f, 2048.0f, 32.0f},  {0.0f, 1.0f, 0.5f},
    {0.1f, 2.0f, 0.8f},       {0.1f, 10.0f, 2.0f},      {2.0f, 1024.0f, 4.0f},     {2.0f, 1024.0f, 4.0f},    {0.98f, 1.0f, 1.0f},
    {0.98f, 1.0f, 1.0f},      {1.0f, 10.0f, 4.0f},      {0.1f, 10.0f, 5.0f},       {0.001f, 0.005f, 0.0015f}, {0.001f, 0.005f, 0.0015f},
    {4.0f, 10.0f, 5.0f},      {0.98f, 1.0f, 1.0f},      {0.98f, 1.0f, 1.0f},       {0.98f, 1.0f, 1.0f},      {0.98f, 1.0f, 1.0f},
    {0.0f, 1.0f, 0.8f},       {0.0f, 1.0f, 0.8f},       {0.0f, 1.0f, 0.8f},        {0.0f, 1.0f, 0.5f},       {0.1f, 2.0f, 0.8f},
    {0.1f, 10.0f, 2.0f},      {0.0f, 0.5f, 0.1f},       {0.0f, 0.5f, 0.1f}};

// FrameCoder::SetParam (libsac.cpp:37-92). Values are clamped to the structural limits of the kernels (the DDS box
// keeps them inside anyway: profile.cpp:47-53,66-67).
HostParam map_profile(const float *g)
{
  HostParam p;
  auto G = [&](int i) { return (double)g[i]; };
  auto R = [&](int i) { return (int)std::round(G(i)); };
  const int in0[4] = {28, 29, 30, 37}, in1[4] = {31, 32, 33, 38};
  const int imu0[4] = {2, 3, 4, 5}, imu1[4] = {14, 15, 16, 17};
  const int imd0[4] = {6, 39, 46, 47}, imd1[4] = {18, 40, 48, 49};
  const int ipd0[4] = {7, 8, 50, 51}, ipd1[4] = {19, 20, 21, 52};
  for (int i = 0; i < 4; i++) {
    p.vn[0][i] = std::clamp(R(in0[i]), 1, 8192); p.vn[1][i] = std::clamp(R(in1[i]), 1, 8192);
    p.vmu[0][i] = G(imu0[i]) / double(p.vn[0][i]); p.vmu[1][i] = G(imu1[i]) / double(p.vn[1][i]);
    p.vmudecay[0][i] = G(imd0[i]); p.vmudecay[1][i] = G(imd1[i]);
    p.vpowdecay[0][i] = G(ipd0[i]); p.vpowdecay[1][i] = G(ipd1[i]);
  }
  p.lambda[0] = G(0); p.ols_nu[0] = G(1); p.mu_mix[0] = G(10); p.mu_mix_beta[0] = G(11);
  p.lambda[1] = G(12); p.ols_nu[1] = G(13); p.mu_mix[1] = G(22); p.mu_mix_beta[1] = G(23);
  p.nA = std::clamp(R(24), 1, 32); p.nB = std::clamp(R(25), 1, 32); p.nS0 = std::clamp(R(26), 0, 32);
  p.nS1 = std::clamp(R(27), -32, 32); p.nM0 = std::clamp(R(9), 0, 32);
  p.beta_sum[0] = G(34); p.beta_pow[0] = G(35); p.beta_add[0] = G(36);
  p.beta_sum[1] = G(53); p.beta_pow[1] = G(54); p.beta_add[1] = G(55);
  p.proj_alpha[0] = G(56); p.proj_alpha[1] = G(57);
  p.lm_n = std::clamp(R(41), 1, (int)kMaxRls); p.lm_alpha = G(42);
  p.bias_mu[0] = G(43); p.bias_mu[1] = G(44);
  p.bias_scale = std::clamp(R(45), 0, 30);
  p.ch_ref = 0;
  if (p.nS1 < 0) { p.nS1 = -p.nS1; p.ch_ref = 1; }
  return p;
}

static int ols_order(const HostParam &hp, int cc) { return cc == 0 ? hp.nA + hp.nM0 : hp.nB + hp.nS0 + hp.nS1; }

long long chain_scratch_doubles(const HostParam &hp, int cc, int)
{
  long long t = 0;
  for (int s = 0; s < 4; s++) t += 4LL * hp.vn[cc][s] + 1;
  const long long n = ols_order(hp, cc);
  t += (n + 1) * (n + 2);
  // the encode-direction kernel has its own layout; reserve for whichever is larger
  return std::max(t + 8, predictor_enc_scratch_doubles(hp.vn[cc], (int)n));
}

int fill_chain(ChainDesc &d, const HostParam &hp, int nch, int cc, int k, const int32_t *const *planes, int from, int n,
               const int32_t *mm)
{
  std::memset(&d, 0, sizeof(d));
  int actual;
  if (nch == 1) {
    actual = 0;
    d.own = planes[0] + from; d.other = planes[0] + from;
    d.lenA = hp.nA; d.lenB = hp.nM0; d.lagB = 0; d.minB = 0; d.backB = hp.nM0;
  } else if (cc == 0) {
    actual = hp.ch_ref;
    d.own = planes[actual] + from; d.other = planes[1 - actual] + from;
    d.lenA = hp.nA; d.lenB = hp.nM0; d.lagB = std::max(hp.nS1, 1) - 1; d.minB = 0; d.backB = hp.nM0;
  } else {
    actual = 1 - hp.ch_ref;
    d.own = planes[actual] + from; d.other = planes[1 - actual] + from;
    d.lenA = hp.nB; d.lenB = hp.nS0 + hp.nS1; d.lagB = 0; d.minB = -(1 << 30); d.backB = hp.nS0;
  }
  d.n = n;
  d.k = k;
  d.lambda = hp.lambda[cc];
  d.nu = (1.0 - hp.lambda[cc]) * hp.ols_nu[cc];                      // ols.cpp:11
  d.beta_sum = hp.beta_sum[cc]; d.beta_pow = hp.beta_pow[cc]; d.beta_add = hp.beta_add[cc];
  for (int s = 0; s < 4; s++) {
    d.vn[s] = hp.vn[cc][s]; d.vmu[s] = hp.vmu[cc][s]; d.vmudecay[s] = hp.vmudecay[cc][s]; d.vpowdecay[s] = hp.vpowdecay[cc][s];
  }
  d.mu_mix = hp.mu_mix[cc]; d.mix_beta = hp.mu_mix_beta[cc];
  d.lm_n = hp.lm_n; d.lm_gamma = hp.lm_alpha; d.proj_alpha = hp.proj_alpha[cc];
  // Cascade ranges follow the coded-channel index, residual clamps the actual channel (libsac.cpp:98-102,106)
  const int ri = (nch == 2) ? cc : 0;
  d.casc_lo = (double)mm[2 * ri]; d.casc_hi = (double)mm[2 * ri + 1];
  d.clamp_lo = mm[2 * actual]; d.clamp_hi = mm[2 * actual + 1];
  d.bias_mu = hp.bias_mu[cc]; d.bias_nscale = 1 << hp.bias_scale;
  return actual;
}

// ---- Engine ----------------------------------------------------------------------------------------------------------
Engine *Engine::helper(int i)
{
  while ((int)helpers.size() <= i) {
    Engine *h = new Engine();
    if (h->init(device, this) != SAC_OK) { delete h; return nullptr; }
    helpers.push_back(h);
  }
  return helpers[i];
}
long long Engine::total_launches() const
{
  long long t = launches;
  for (const Engine *h : helpers) t += h->total_launches();          // helpers of helpers run the final passes of frames in flight
  return t;
}

int Engine::init(int dev, const Engine *parent)
{
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
    set_error("no CUDA device visible: libsac_b200 has no CPU fallback");
    return SAC_E_NODEVICE;
  }
  if (dev < 0 || dev >= cnt) { set_error("device index out of range"); return SAC_E_ARG; }
  device = dev;
  SACB_CUDA(cudaSetDevice(dev));
  {
    // dynamic shared-memory limits of every kernel, once per process (launches come from several host threads)
    static std::once_flag once_dev[64];
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once_dev[dev & 63], [] {
      attr_err = predictor_enc_init_attributes();
      if (attr_err == cudaSuccess) attr_err = predictor_sg_init_attributes();
      if (attr_err == cudaSuccess) attr_err = predictor_init_attributes();
      if (attr_err == cudaSuccess) attr_err = bitplane_init_attributes();
    });
    SACB_CUDA(attr_err);
  }
  if (parent) {
    int lo = 0, hi = 0;
    SACB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    SACB_CUDA(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, hi));
    is_helper = true;
  } else {
    SACB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  }
  for (auto &e : ev) SACB_CUDA(cudaEventCreate(&e));
  SACB_CUDA(cudaEventCreateWithFlags(&ev_wait, cudaEventBlockingSync | cudaEventDisableTiming));
  {
    // pools allocate stream-ordered on this engine's stream; freed blocks stay in the device's pool (no trim to the OS)
    static std::once_flag once_pool[64];
    std::call_once(once_pool[dev & 63], [dev] {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ULL;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
      }
    });
    d_descs.s = d_scratch.s = d_scratch_ols.s = d_plpc.s = d_resid.s = d_sums.s = d_flags.s = d_bpjobs.s = d_csig0.s = d_hist.s = d_cost.s = d_bytes.s =
        d_sparse.s = d_idx.s = stream;
  }
  for (auto &s2 : side) SACB_CUDA(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
  SACB_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
  for (auto &e : ev_join) SACB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  if (const char *s = std::getenv("SAC_B200_SMEM_KB")) smem_bytes = std::clamp(std::atoi(s), 24, 226) * 1024;
  if (const char *s = std::getenv("SAC_B200_ENC_SMEM_KB")) enc_smem_bytes = std::clamp(std::atoi(s), 24, 226) * 1024;
  if (const char *s = std::getenv("SAC_B200_OLS_SMEM_KB")) ols_smem_bytes = std::clamp(std::atoi(s), 12, 226) * 1024;
  if (const char *s = std::getenv("SAC_B200_DEDUP")) dedup = std::atoi(s) != 0;
  if (parent) {
    dedup = parent->dedup; grade = parent->grade;
    enc_smem_bytes = parent->enc_smem_bytes; ols_smem_bytes = parent->ols_smem_bytes; ols_smem_cap_bytes = parent->ols_smem_cap_bytes; smem_bytes = parent->smem_bytes;
    bt = parent->bt;                                                 // shared device tables (owned by the parent)
    return SAC_OK;
  }
  SACB_CUDA(bt.init(stream));
  return SAC_OK;
}
void Engine::destroy()
{
  cudaSetDevice(device);
  for (Engine *h : helpers) { h->destroy(); delete h; }
  helpers.clear();
  if (stream) cudaStreamSynchronize(stream);
  d_descs.release(); h_descs.release(); d_scratch.release(); d_scratch_ols.release(); d_plpc.release(); d_resid.release(); d_sums.release(); d_flags.release();
  d_idx.release(); h_idx.release();
  h_sums.release(); h_flags.release(); d_bpjobs.release(); h_bpjobs.release(); d_csig0.release(); d_hist.release();
  d_cost.release(); h_cost.release(); d_bytes.release(); h_stage.release(); d_sparse.release(); h_sparse.release();
  if (!is_helper) bt.destroy();
  for (auto &e : ev) if (e) cudaEventDestroy(e);
  if (ev_wait) cudaEventDestroy(ev_wait);
  for (auto &s2 : side) if (s2) { cudaStreamSynchronize(s2); cudaStreamDestroy(s2); s2 = nullptr; }
  if (ev_fork) cudaEventDestroy(ev_fork);
  for (auto &e : ev_join) if (e) cudaEventDestroy(e);
  if (stream) cudaStreamDestroy(stream);
  stream = nullptr;
}
cudaError_t Engine::wait()
{
  cudaError_t e = cudaEventRecord(ev_wait, stream);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(ev_wait);
}
void Engine::begin_call()
{
  for (int i = 0; i < 4; i++) { last_ms[i] = 0; last_launches[i] = 0; }
}

// Chains of one call that are bit-identical computations are evaluated once. Late in a DDS search a candidate differs
// from the incumbent in a handful of the 56 dimensions, so most candidates of a generation leave one channel's
// parameters -- or at least its 8 OLS parameters -- untouched: those chains (resp. their OLS stage) coincide.
//   slot_of[c]  unique chain ("slot") of logical chain c = (job, coded channel); residuals, sums, flags are per slot
//   OLS stage   one ols_kernel CTA per distinct (planes, range, k, OLS parameters); slots share its p_lpc plane
int Engine::run_predict(const std::vector<Job> &jobs, std::vector<int> &chain_job, std::vector<int> &chain_ch, size_t &stride, int grade_req)
{
  SACB_CUDA(cudaSetDevice(device));
  chain_job.clear(); chain_ch.clear();
  slot_of.clear(); slot_rep.clear();
  size_t maxn = 0;
  int nlog = 0;
  for (auto &j : jobs) { nlog += j.win->nch; maxn = std::max(maxn, (size_t)j.n); }
  stride = (maxn + 31) & ~size_t(31);
  // ---- logical chains -> slots ----
  std::vector<HostParam> hps(jobs.size());
  std::vector<ChainDesc> descs;                                      // one per slot, inputs only
  std::vector<int> slot_job, slot_cc;
  {
    std::unordered_map<std::string, int> seen;
    for (size_t ji = 0; ji < jobs.size(); ji++) {
      const Job &j = jobs[ji];
      if (j.from < 0 || j.n <= 0 || j.from + j.n > j.win->numsamples) { set_error("window range outside the frame"); return SAC_E_ARG; }
      hps[ji] = map_profile(j.profile);
      for (int cc = 0; cc < j.win->nch; cc++) {
        ChainDesc d;
        const int actual = fill_chain(d, hps[ji], j.win->nch, cc, j.k, j.win->d_planes, j.from, j.n, j.win->minmax);
        const int c = (int)chain_job.size();
        chain_job.push_back((int)ji); chain_ch.push_back(actual);
        int u = -1;
        if (dedup) {
          const std::string key(reinterpret_cast<const char *>(&d), sizeof(d));   // inputs only: outputs are still zero
          auto it = seen.find(key);
          if (it != seen.end()) u = it->second; else seen.emplace(key, (int)descs.size());
        }
        if (u < 0) { u = (int)descs.size(); descs.push_back(d); slot_rep.push_back(c); slot_job.push_back((int)ji); slot_cc.push_back(cc); }
        slot_of.push_back(u);
      }
    }
  }
  const int nu = nslots = (int)descs.size();
  // ---- distinct OLS stages among the slots ----
  struct OlsKey { const int32_t *own, *other; int n, lenA, lenB, lagB, minB, backB, k, pad; double lambda, nu, beta_sum, beta_pow, beta_add; };
  std::vector<int> ols_of(nu), ols_rep;
  {
    std::unordered_map<std::string, int> seen;
    for (int u = 0; u < nu; u++) {
      const ChainDesc &d = descs[u];
      OlsKey k;
      std::memset(&k, 0, sizeof(k));
      k.own = d.own; k.other = d.other; k.n = d.n; k.lenA = d.lenA; k.lenB = d.lenB; k.lagB = d.lagB; k.minB = d.minB; k.backB = d.backB; k.k = d.k;
      k.lambda = d.lambda; k.nu = d.nu; k.beta_sum = d.beta_sum; k.beta_pow = d.beta_pow; k.beta_add = d.beta_add;
      int v = -1;
      if (dedup) {
        const std::string key(reinterpret_cast<const char *>(&k), sizeof(k));
        auto it = seen.find(key);
        if (it != seen.end()) v = it->second; else seen.emplace(key, (int)ols_rep.size());
      }
      if (v < 0) { v = (int)ols_rep.size(); ols_rep.push_back(u); }
      ols_of[u] = v;
    }
  }
  const int nv = (int)ols_rep.size();
  last_unique[0] = nlog; last_unique[1] = nu; last_unique[2] = nv;
  g_dedup_totals[0] += nlog; g_dedup_totals[1] += nu; g_dedup_totals[2] += nv;
  SACB_CUDA(h_descs.reserve(nu + nv));
  SACB_CUDA(d_descs.reserve(nu + nv));
  SACB_CUDA(d_resid.reserve(stride * nu));
  SACB_CUDA(d_sums.reserve((size_t)3 * nu));
  SACB_CUDA(d_flags.reserve((size_t)4 * nu));
  SACB_CUDA(d_plpc.reserve(stride * nv));
  // scratch offsets, shared-memory requests
  std::vector<long long> soff(nu + 1, 0), ooff(nv + 1, 0);
  size_t need_w = 0, need_both = 0, casc_need = 0;
  for (int u = 0; u < nu; u++) {
    const HostParam &hp = hps[slot_job[u]];
    soff[u + 1] = soff[u] + ((chain_scratch_doubles(hp, slot_cc[u], jobs[slot_job[u]].win->nch) + 1) & ~1LL);
    casc_need = std::max(casc_need, (size_t)predictor_enc_smem_doubles(hp.vn[slot_cc[u]]) * 8);
  }
  for (int v = 0; v < nv; v++) {
    const int u = ols_rep[v];
    const size_t n = (size_t)ols_order(hps[slot_job[u]], slot_cc[u]), ld = (n + 1) | 1, mat = (n + 1) * ld * 8;
    ooff[v + 1] = ooff[v] + ((predictor_ols_scratch_doubles((int)n) + 1) & ~1LL);
    need_w = std::max(need_w, mat); need_both = std::max(need_both, 2 * mat);
  }
  SACB_CUDA(d_scratch.reserve((size_t)soff[nu]));
  SACB_CUDA(d_scratch_ols.reserve((size_t)ooff[nv]));
  // cascade kernel: histories and overflow taps of the largest chain, capped so that two CTAs still fit an SM.
  // OLS kernel: work matrix always (fast LDL path), covariance too while both stay within the cap
  const int casc_smem = (int)std::min<size_t>(predictor_enc_shared_bytes() + casc_need + 64, (size_t)enc_smem_bytes);
  const size_t ols_head = predictor_ols_shared_bytes();
  const size_t ols_cap = (size_t)ols_smem_cap_bytes;
  int ols_smem = (int)std::max<size_t>(std::min(ols_head + need_both, ols_cap), std::min<size_t>(ols_head + need_w, 200 * 1024));
  ols_smem = std::max(ols_smem, ols_smem_bytes);
  for (int u = 0; u < nu; u++) {
    ChainDesc &d = h_descs.p[u];
    d = descs[u];
    d.resid = d_resid.p + (size_t)u * stride;
    d.scratch = d_scratch.p + soff[u];
    d.scratch_doubles = soff[u + 1] - soff[u];
    d.plpc = d_plpc.p + (size_t)ols_of[u] * stride;
    d.l1sum = d_sums.p + u; d.sqsum = d_sums.p + nu + u; d.flags = d_flags.p + u;
  }
  for (int v = 0; v < nv; v++) {                                    // OLS descriptors follow the slot descriptors
    ChainDesc &d = h_descs.p[nu + v];
    d = descs[ols_rep[v]];
    d.scratch_ols = d_scratch_ols.p + ooff[v];
    d.plpc = d_plpc.p + (size_t)v * stride;
  }
  SACB_CUDA(cudaMemcpyAsync(d_descs.p, h_descs.p, sizeof(ChainDesc) * (nu + nv), cudaMemcpyHostToDevice, stream));
  SACB_CUDA(cudaEventRecord(ev[0], stream));
  ev4_recorded = true;
  if (grade_req == 0) {
    SACB_CUDA(launch_predictor_enc(d_descs.p + nu, nv, d_descs.p, nu, casc_smem, ols_smem, stream, ev[4]));
    SACB_CUDA(cudaEventRecord(ev[1], stream));
    launches += 2; last_launches[0] += 2;                           // ols_kernel + cascade_kernel
    return SAC_OK;
  }
  // ---- search-grade kernels: OLS stages by matrix size class, cascades by what fits registers / shared memory ----
  // (probe switch SACB_SG_PARTS: bit 0 = search-grade OLS, bit 1 = search-grade cascade; default both)
  static const int sg_parts = [] { const char *v = std::getenv("SACB_SG_PARTS"); return v ? std::atoi(v) : 3; }();
  {
    // OLS: orders up to 32 on the one-warp register kernels (classes 16 / 24 / 32), larger ones on the canonical team kernel
    // (measured: the 256-thread block-cyclic search-grade kernel loses to it, profiles/README.md). Cascade: small (two CTAs per SM),
    // large (one), canonical fallback for chains whose tables exceed shared memory.
    const int kOlsClasses = 5;
    // ols_sg_class values: one warp per chain for orders <= 32; 0 = canonical team kernel for larger orders. (A two-warp
    // rows-in-registers kernel for orders 33..64 exists, ols_rows_kernel<2>, class 64, SACB_OLS_ROWS=1: measured SLOWER than the
    // canonical kernel on the B200 -- 257 ms against 230 ms for a whole heterogeneous generation, profiles/README.md -- so it is off.)
    static const bool use_rows = [] { const char *v = std::getenv("SACB_OLS_ROWS"); return v && v[0] == '1'; }();
    const int cls_id[kOlsClasses] = {16, 24, 32, 64, 0};
    std::vector<int> ols_cls[kOlsClasses], casc[3];
    size_t ols_sm[kOlsClasses] = {0, 0, 0, 0, 0}, casc_sm[3] = {0, 0, 0};
    size_t need_w2 = 0, need_both2 = 0;
    for (int v = 0; v < nv; v++) {
      const int u = ols_rep[v];
      const int n = ols_order(hps[slot_job[u]], slot_cc[u]);
      const int c = n <= 16 ? 0 : (n <= 24 ? 1 : (n <= 32 ? 2 : ((n <= 64 && use_rows) ? 3 : 4)));
      ols_cls[c].push_back(nu + v);
      if (c < 4) ols_sm[c] = std::max(ols_sm[c], ols_sg_smem_bytes(n));
      else { const size_t ld = ((size_t)n + 1) | 1, mat = ((size_t)n + 1) * ld * 8; need_w2 = std::max(need_w2, mat); need_both2 = std::max(need_both2, 2 * mat); }
    }
    if (!ols_cls[4].empty())
      ols_sm[4] = std::max<size_t>(std::max<size_t>(std::min(ols_head + need_both2, ols_cap), std::min<size_t>(ols_head + need_w2, 200 * 1024)), (size_t)ols_smem_bytes);
    const size_t kSmallCap = 112 * 1024, kLargeCap = 226 * 1024;     // two CTAs per SM / one
    for (int u = 0; u < nu; u++) {
      const int *vn = hps[slot_job[u]].vn[slot_cc[u]];
      const size_t ss = cascade_sg_smem_bytes(vn, 0), sl = cascade_sg_smem_bytes(vn, 1);
      int c = 2;
      if (ss && ss <= kSmallCap) c = 0; else if (sl && sl <= kLargeCap) c = 1;
      if (!(sg_parts & 2)) c = 2;
      casc[c].push_back(u);
      casc_sm[c] = std::max(casc_sm[c], c == 0 ? ss : (c == 1 ? sl : predictor_enc_shared_bytes() + (size_t)predictor_enc_smem_doubles(vn) * 8 + 64));
    }
    casc_sm[2] = std::min(casc_sm[2], (size_t)enc_smem_bytes);
    sg_stats[0] += (long long)casc[0].size(); sg_stats[1] += (long long)casc[1].size(); sg_stats[2] += (long long)casc[2].size();
    SACB_CUDA(h_idx.reserve((size_t)nu + nv));
    SACB_CUDA(d_idx.reserve((size_t)nu + nv));
    size_t off = 0, ooff[kOlsClasses], coff[3];
    for (int c = 0; c < kOlsClasses; c++) { ooff[c] = off; for (int x : ols_cls[c]) h_idx.p[off++] = x; }
    for (int c = 0; c < 3; c++) { coff[c] = off; for (int x : casc[c]) h_idx.p[off++] = x; }
    SACB_CUDA(cudaMemcpyAsync(d_idx.p, h_idx.p, sizeof(int) * off, cudaMemcpyHostToDevice, stream));
    // fork: the classes run side by side (a launch ends with its slowest chain; serialised they would add up)
    int used = 0;
    auto fork_stream = [&](int k) -> cudaStream_t { return k == 0 ? stream : side[(k - 1) % kSide]; };
    auto join_all = [&]() -> int {
      for (int k = 1; k < used && k <= kSide; k++) {
        SACB_CUDA(cudaEventRecord(ev_join[k - 1], side[k - 1]));
        SACB_CUDA(cudaStreamWaitEvent(stream, ev_join[k - 1], 0));
      }
      used = 0;
      return SAC_OK;
    };
    auto next_stream = [&](cudaStream_t &out) -> int {
      const int k = used++;
      out = fork_stream(k);
      if (k >= 1 && k <= kSide) SACB_CUDA(cudaStreamWaitEvent(out, ev_fork, 0));
      return SAC_OK;
    };
    SACB_CUDA(cudaEventRecord(ev_fork, stream));
    if (!(sg_parts & 1)) {                                          // probe: canonical OLS kernel over all stages
      cudaStream_t s2; int rc2 = next_stream(s2); if (rc2) return rc2;
      SACB_CUDA(launch_ols_canonical(d_descs.p + nu, nullptr, nv, ols_smem, s2));
      launches++; last_launches[0]++;
    }
    for (int c = kOlsClasses - 1; c >= 0 && (sg_parts & 1); c--) {  // the longest-running class first, on the main stream
      if (ols_cls[c].empty()) continue;
      cudaStream_t s2; int rc2 = next_stream(s2); if (rc2) return rc2;
      if (cls_id[c] == 0) SACB_CUDA(launch_ols_canonical(d_descs.p, d_idx.p + ooff[c], (int)ols_cls[c].size(), (int)ols_sm[c], s2));
      else SACB_CUDA(launch_ols_sg(d_descs.p, d_idx.p + ooff[c], (int)ols_cls[c].size(), cls_id[c], (int)ols_sm[c], s2));
      launches++; last_launches[0]++;
    }
    { int rc2 = join_all(); if (rc2) return rc2; }
    SACB_CUDA(cudaEventRecord(ev[4], stream));
    SACB_CUDA(cudaEventRecord(ev_fork, stream));
    for (int c = 2; c >= 0; c--) {
      if (casc[c].empty()) continue;
      cudaStream_t s2; int rc2 = next_stream(s2); if (rc2) return rc2;
      if (c == 2) SACB_CUDA(launch_cascade_canonical(d_descs.p, d_idx.p + coff[2], (int)casc[2].size(), (int)casc_sm[2], s2));
      else SACB_CUDA(launch_cascade_sg(d_descs.p, d_idx.p + coff[c], (int)casc[c].size(), c, (int)casc_sm[c], s2));
      launches++; last_launches[0]++;
    }
    { int rc2 = join_all(); if (rc2) return rc2; }
    SACB_CUDA(cudaEventRecord(ev[1], stream));
  }
  return SAC_OK;
}

int Engine::run_cost(int cost_kind, const std::vector<Job> &jobs, const std::vector<int> &chain_job, const std::vector<int> &chain_ch,
                     size_t stride, double *cost)
{
  const int nlog = (int)chain_job.size();
  const int nchains = nslots;                                       // everything below is per slot (unique chain)
  auto rep_job = [&](int u) -> const Job & { return jobs[chain_job[slot_rep[u]]]; };
  SACB_CUDA(h_sums.reserve((size_t)3 * nchains));
  SACB_CUDA(h_flags.reserve((size_t)2 * nchains));
  SACB_CUDA(h_cost.reserve(nchains));
  SACB_CUDA(d_cost.reserve(nchains));
  std::vector<double> ccost(nchains, 0.0);
  SACB_CUDA(cudaEventRecord(ev[2], stream));
  if (cost_kind == SAC_COST_ENTROPY || cost_kind == SAC_COST_GOLOMB) {
    // ns / ranges ride in the flags pool's second half
    std::vector<int> meta(2 * nchains);
    size_t maxbins = 0;
    for (int c = 0; c < nchains; c++) {
      const Job &j = rep_job(c);
      meta[c] = j.n;
      const int ch = chain_ch[slot_rep[c]];
      const long long R = (long long)j.win->minmax[2 * ch + 1] - (long long)j.win->minmax[2 * ch];
      if (cost_kind == SAC_COST_ENTROPY && R > (1 << 20)) { set_error("entropy cost: residual range above 2^20 not supported"); return SAC_E_UNSUPPORTED; }
      meta[nchains + c] = (int)std::max<long long>(R, 1);
      maxbins = std::max(maxbins, (size_t)(2 * meta[nchains + c] + 1));
    }
    DevBuf<int> &dm = d_flags;                                      // [0,n) flags, [2n,3n) ns, [3n,4n) ranges (reserved in run_predict)
    SACB_CUDA(cudaMemcpyAsync(dm.p + 2 * nchains, meta.data(), sizeof(int) * 2 * nchains, cudaMemcpyHostToDevice, stream));
    if (cost_kind == SAC_COST_ENTROPY) {
      const size_t hs = (maxbins + 31) & ~size_t(31);
      SACB_CUDA(d_hist.reserve(hs * nchains));
      SACB_CUDA(launch_entropy(d_resid.p, stride, dm.p + 2 * nchains, dm.p + 3 * nchains, nchains, d_hist.p, hs, d_cost.p, stream));
    } else {
      SACB_CUDA(launch_golomb(d_resid.p, stride, dm.p + 2 * nchains, nchains, d_cost.p, stream));
    }
    launches++; last_launches[2]++;
    SACB_CUDA(cudaMemcpyAsync(h_cost.p, d_cost.p, sizeof(double) * nchains, cudaMemcpyDeviceToHost, stream));
    SACB_CUDA(cudaMemcpyAsync(h_flags.p, d_flags.p, sizeof(int) * nchains, cudaMemcpyDeviceToHost, stream));
    SACB_CUDA(cudaEventRecord(ev[3], stream));
    SACB_CUDA(wait());
    for (int c = 0; c < nchains; c++) ccost[c] = (h_flags.p[c] & 1) ? std::numeric_limits<double>::infinity() : h_cost.p[c];
    float ms = 0; cudaEventElapsedTime(&ms, ev[2], ev[3]); last_ms[2] += ms;
  } else if (cost_kind == SAC_COST_BITPLANE) {
    SACB_CUDA(h_bpjobs.reserve(nchains));
    SACB_CUDA(d_bpjobs.reserve(nchains));
    SACB_CUDA(d_csig0.reserve((size_t)65536 * nchains));
    for (int c = 0; c < nchains; c++) {
      BpJob &b = h_bpjobs.p[c];
      std::memset(&b, 0, sizeof(b));
      b.buf = d_resid.p + (size_t)c * stride; b.n = rep_job(c).n; b.signed_input = 1; b.maxbpn = -1;
      b.csig0 = d_csig0.p + (size_t)65536 * c; b.nbytes = d_sums.p + 2 * (size_t)nchains + c; b.maxbpn_out = nullptr;
    }
    SACB_CUDA(cudaMemcpyAsync(d_bpjobs.p, h_bpjobs.p, sizeof(BpJob) * nchains, cudaMemcpyHostToDevice, stream));
    SACB_CUDA(launch_bitplane(bt, d_bpjobs.p, nchains, 0, stream));
    launches++; last_launches[1]++;
    SACB_CUDA(cudaEventRecord(ev[3], stream));
    SACB_CUDA(cudaMemcpyAsync(h_sums.p, d_sums.p, sizeof(long long) * 3 * nchains, cudaMemcpyDeviceToHost, stream));
    SACB_CUDA(cudaMemcpyAsync(h_flags.p, d_flags.p, sizeof(int) * nchains, cudaMemcpyDeviceToHost, stream));
    SACB_CUDA(wait());
    for (int c = 0; c < nchains; c++)
      ccost[c] = (h_flags.p[c] & 1) ? std::numeric_limits<double>::infinity() : (double)h_sums.p[2 * nchains + c];
    float ms = 0; cudaEventElapsedTime(&ms, ev[2], ev[3]); last_ms[1] += ms;
  } else if (cost_kind == SAC_COST_L1 || cost_kind == SAC_COST_RMS) {
    SACB_CUDA(cudaMemcpyAsync(h_sums.p, d_sums.p, sizeof(long long) * 3 * nchains, cudaMemcpyDeviceToHost, stream));
    SACB_CUDA(cudaMemcpyAsync(h_flags.p, d_flags.p, sizeof(int) * nchains, cudaMemcpyDeviceToHost, stream));
    SACB_CUDA(wait());
    for (int c = 0; c < nchains; c++) {
      const double n = (double)rep_job(c).n;
      double v = cost_kind == SAC_COST_L1 ? h_sums.p[c] / n : std::sqrt(h_sums.p[nchains + c] / n);   // cost.h:15-41
      ccost[c] = (h_flags.p[c] & 1) ? std::numeric_limits<double>::infinity() : v;
    }
  } else { set_error("unknown cost kind"); return SAC_E_ARG; }
  {
    float ms = 0, ms_ols = 0;
    cudaEventElapsedTime(&ms, ev[0], ev[1]); last_ms[0] += ms;
    if (ev4_recorded) { cudaEventElapsedTime(&ms_ols, ev[0], ev[4]); last_ms[3] += ms_ols; last_launches[3]++; }
    ev4_recorded = false;
    (void)cudaGetLastError();                                       // an optional timing query must not leave an error pending
    total_ms[0] += ms_ols; total_ms[1] += ms - ms_ols; total_calls++;
    if (cost_kind == SAC_COST_BITPLANE) { float b = 0; cudaEventElapsedTime(&b, ev[2], ev[3]); total_ms[2] += b; }
    else if (cost_kind == SAC_COST_ENTROPY || cost_kind == SAC_COST_GOLOMB) { float b = 0; cudaEventElapsedTime(&b, ev[2], ev[3]); total_ms[3] += b; }
  }
  for (size_t j = 0; j < jobs.size(); j++) cost[j] = 0.0;
  job_inexact.assign(jobs.size(), 0);
  // sum over channels (FrameCoder::GetCost, libsac.cpp:358-361; a two-term sum is order-independent)
  for (int c = 0; c < nlog; c++) {
    cost[chain_job[c]] += ccost[slot_of[c]];
    if (h_flags.p[slot_of[c]] & 2) job_inexact[chain_job[c]] = 1;   // search-grade cascade: look-ahead met the weight clamp
  }
  return SAC_OK;
}

} // namespace sacb

// =====================================================================================================================
using namespace sacb;

extern "C" {

const char *sac_version(void) { return "sac_b200 0.1 (bitstream: Sac v0.7.25 container, canonical-arithmetic model)"; }
const char *sac_last_error(void) { return g_err.c_str(); }
int sac_device_count(void)
{
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) return 0;
  return c;
}

sac_engine *sac_engine_create(int device)
{
  Engine *e = new Engine();
  if (e->init(device) != SAC_OK) { delete e; return nullptr; }
  return reinterpret_cast<sac_engine *>(e);
}
void sac_engine_destroy(sac_engine *h)
{
  if (!h) return;
  Engine *e = reinterpret_cast<Engine *>(h);
  e->destroy();
  delete e;
}
long long sac_engine_launches(const sac_engine *h) { return reinterpret_cast<const Engine *>(h)->total_launches(); }
void sac_engine_last_timing(const sac_engine *h, double *out_ms, long long *out_launches)
{
  const Engine *e = reinterpret_cast<const Engine *>(h);
  for (int i = 0; i < 4; i++) { if (out_ms) out_ms[i] = e->last_ms[i]; if (out_launches) out_launches[i] = e->last_launches[i]; }
}

void sac_engine_total_timing(const sac_engine *h, double *out_ms4, long long *out_calls)
{
  const Engine *e = reinterpret_cast<const Engine *>(h);
  for (int i = 0; i < 4; i++) out_ms4[i] = e->total_ms[i];
  long long calls = e->total_calls;
  for (const Engine *x : e->helpers) {
    double t[4]; long long c = 0;
    sac_engine_total_timing(reinterpret_cast<const sac_engine *>(x), t, &c);
    for (int i = 0; i < 4; i++) out_ms4[i] += t[i];
    calls += c;
  }
  if (out_calls) *out_calls = calls;
}

int sac_engine_set_grade(sac_engine *h, int grade)
{
  Engine *e = reinterpret_cast<Engine *>(h);
  if (!e) return -1;
  const int prev = e->grade;
  e->grade = grade ? 1 : 0;
  for (Engine *x : e->helpers) sac_engine_set_grade(reinterpret_cast<sac_engine *>(x), grade);
  return prev;
}
void sac_engine_grade_stats(const sac_engine *h, long long *out4)
{
  const Engine *e = reinterpret_cast<const Engine *>(h);
  for (int i = 0; i < 4; i++) out4[i] = e->sg_stats[i];
  for (const Engine *x : e->helpers) { long long t[4]; sac_engine_grade_stats(reinterpret_cast<const sac_engine *>(x), t); for (int i = 0; i < 4; i++) out4[i] += t[i]; }
}

int sac_engine_set_dedup(sac_engine *h, int on)
{
  Engine *e = reinterpret_cast<Engine *>(h);
  if (!e) return -1;
  const int prev = e->dedup ? 1 : 0;
  e->dedup = on != 0;
  for (Engine *x : e->helpers) sac_engine_set_dedup(reinterpret_cast<sac_engine *>(x), on);
  return prev;
}
void sac_dedup_totals(long long *out3)
{
  for (int i = 0; i < 3; i++) out3[i] = g_dedup_totals[i].load();
}

double sac_fp64_peak_gflops(sac_engine *h)
{
  Engine *e = reinterpret_cast<Engine *>(h);
  if (!e) return -1.0;
  cudaSetDevice(e->device);
  const int blocks = 148 * 8, iters = 1 << 15;
  if (e->d_cost.reserve((size_t)blocks * 256) != cudaSuccess) return -1.0;
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e->ev[2], e->stream);
    if (launch_dfma_peak(e->d_cost.p, blocks, iters, e->stream) != cudaSuccess) return -1.0;
    cudaEventRecord(e->ev[3], e->stream);
    cudaStreamSynchronize(e->stream);
    e->launches++;
    float ms = 0;
    cudaEventElapsedTime(&ms, e->ev[2], e->ev[3]);
    const double gf = 2.0 * 8.0 * iters * (double)blocks * 256.0 / (ms * 1e-3) / 1e9;
    if (rep > 0 && gf > best) best = gf;
  }
  return best;
}

void sac_model_tables(int16_t *stretch, int16_t *squash) { compute_logdomain_tables(stretch, squash); }

int sac_base_profile(float *vmin, float *vmax, float *vdef)
{
  for (int i = 0; i < kProfileSize; i++) {
    if (vmin) vmin[i] = kBaseProfile[i][0];
    if (vmax) vmax[i] = kBaseProfile[i][1];
    if (vdef) vdef[i] = kBaseProfile[i][2];
  }
  return kProfileSize;
}

sac_window *sac_window_create(sac_engine *h, int nch, const int32_t *const *planes, int numsamples, const int32_t *minmax)
{
  Engine *e = reinterpret_cast<Engine *>(h);
  if (!e || nch < 1 || nch > 2 || numsamples <= 0 || !planes || !minmax) { set_error("sac_window_create: bad argument"); return nullptr; }
  cudaSetDevice(e->device);
  Window *w = new Window();
  w->eng = e; w->device = e->device; w->stream = e->stream; w->nch = nch; w->numsamples = numsamples;
  w->d_planes[0] = w->d_planes[1] = nullptr;
  for (int i = 0; i < 2 * nch; i++) w->minmax[i] = minmax[i];
  if (nch == 1) { w->minmax[2] = minmax[0]; w->minmax[3] = minmax[1]; }
  if (e->h_stage.reserve(numsamples) != cudaSuccess) { delete w; set_error("pinned staging allocation failed"); return nullptr; }
  for (int ch = 0; ch < nch; ch++) {
    if (cudaMallocAsync(reinterpret_cast<void **>(&w->d_planes[ch]), sizeof(int32_t) * (size_t)numsamples, e->stream) != cudaSuccess) { set_error("HBM allocation failed"); sac_window_destroy(reinterpret_cast<sac_window *>(w)); return nullptr; }
    std::memcpy(e->h_stage.p, planes[ch], sizeof(int32_t) * (size_t)numsamples);
    cudaMemcpyAsync(w->d_planes[ch], e->h_stage.p, sizeof(int32_t) * (size_t)numsamples, cudaMemcpyHostToDevice, e->stream);
    cudaStreamSynchronize(e->stream);
  }
  return reinterpret_cast<sac_window *>(w);
}
void sac_window_destroy(sac_window *h)
{
  Window *w = reinterpret_cast<Window *>(h);
  if (!w) return;
  cudaSetDevice(w->device);
  // stream-ordered on the per-thread default stream: every user of the planes has been waited for by the time a window is
  // destroyed, and -- unlike cudaFree -- this does not synchronise the device under the other frames in flight
  for (int ch = 0; ch < 2; ch++) if (w->d_planes[ch]) cudaFreeAsync(w->d_planes[ch], cudaStreamPerThread);
  delete w;
}

int sac_predict(sac_engine *h, const sac_window *wh, const float *profiles, int P, int from, int n, int k, int32_t *resid, int *flags)
{
  Engine *e = reinterpret_cast<Engine *>(h);
  const Window *w = reinterpret_cast<const Window *>(wh);
  if (!e || !w || !profiles || P <= 0 || !resid || k < 1) { set_error("sac_predict: bad argument"); return SAC_E_ARG; }
  e->begin_call();
  std::vector<Job> jobs(P);
  for (int p = 0; p < P; p++) {
    jobs[p].win = w; jobs[p].from = from; jobs[p].n = n; jobs[p].k = k;
    std::memcpy(jobs[p].profile, profiles + (size_t)p * kProfileSize, sizeof(float) * kProfileSize);
  }
  std::vector<int> cj, cc;
  size_t stride;
  int rc = e->run_predict(jobs, cj, cc, stride, e->grade);
  if (rc) return rc;
  const int nchains = (int)cj.size();
  SACB_CUDA(e->h_flags.reserve((size_t)2 * nchains));
  SACB_CUDA(cudaMemcpyAsync(e->h_flags.p, e->d_flags.p, sizeof(int) * e->nslots, cudaMemcpyDeviceToHost, e->stream));
  for (int c = 0; c < nchains; c++) {
    int32_t *dst = resid + ((size_t)cj[c] * w->nch + cc[c]) * n;
    SACB_CUDA(cudaMemcpyAsync(dst, e->d_resid.p + (size_t)e->slot_of[c] * stride, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, e->stream));
  }
  SACB_CUDA(cudaStreamSynchronize(e->stream));
  { float ms = 0; cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]); e->last_ms[0] += ms; }
  if (flags) {
    for (int p = 0; p < P; p++) flags[p] = 0;
    for (int c = 0; c < nchains; c++) flags[cj[c]] |= e->h_flags.p[e->slot_of[c]];
  }
  return SAC_OK;
}

int sac_eval_jobs(sac_engine *h, int njobs, const sac_window *const *wins, const int *from, const int *n, const float *bases,
                  const int *dims, int D, const double *X, int cost_kind, int optk, double *cost)
{
  Engine *e = reinterpret_cast<Engine *>(h);
  if (!e || njobs <= 0 || !wins || !from || !n || !bases || !X || !cost || D < 0 || D > kProfileSize || optk < 1) {
    set_error("sac_eval_jobs: bad argument");
    return SAC_E_ARG;
  }
  e->begin_call();
  std::vector<Job> jobs(njobs);
  for (int j = 0; j < njobs; j++) {
    jobs[j].win = reinterpret_cast<const Window *>(wins[j]);
    if (!jobs[j].win) { set_error("sac_eval_jobs: null window"); return SAC_E_ARG; }
    jobs[j].from = from[j]; jobs[j].n = n[j]; jobs[j].k = optk;
    std::memcpy(jobs[j].profile, bases + (size_t)j * kProfileSize, sizeof(float) * kProfileSize);
    for (int i = 0; i < D; i++) {
      const int idx = dims ? dims[i] : (i < 56 ? i : -1);
      if (idx < 0 || idx >= kProfileSize) { set_error("sac_eval_jobs: bad dimension index"); return SAC_E_ARG; }
      jobs[j].profile[idx] = (float)X[(size_t)j * D + i];           // vdef is a float (profile.h:70-72, libsac.cpp:394)
    }
  }
  // Experiment switch (off by default, SACB_LPT=1): evaluate the jobs longest-expected-first. CTAs are dispatched in slot
  // order = order of first appearance, so this lets the chains with thousands of taps / high OLS orders start first when
  // the launch has more chains than resident CTA slots. Costs are scattered back to the caller's order; results are
  // unaffected (every job is an independent computation).
  static const bool lpt = [] { const char *v = std::getenv("SACB_LPT"); return v && v[0] == '1'; }();
  std::vector<int> order;
  std::vector<double> cost_sorted;
  if (lpt && njobs > 1) {
    std::vector<double> est(njobs);
    for (int j = 0; j < njobs; j++) {
      const HostParam hp = map_profile(jobs[j].profile);
      double w = 0.0;
      for (int cc = 0; cc < jobs[j].win->nch; cc++) {
        const double no = (double)ols_order(hp, cc);
        double taps = 0.0;
        for (int s2 = 0; s2 < 4; s2++) taps += hp.vn[cc][s2];
        w = std::max(w, (double)jobs[j].n * (taps / 128.0 + no * no * no / (8.0 * optk) + 40.0));   // the job ends with its slower channel
      }
      est[j] = w;
    }
    order.resize(njobs);
    for (int j = 0; j < njobs; j++) order[j] = j;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return est[a] > est[b]; });
    std::vector<Job> sorted(njobs);
    for (int i = 0; i < njobs; i++) sorted[i] = jobs[order[i]];
    jobs.swap(sorted);
    cost_sorted.resize(njobs);
  }
  std::vector<int> cj, cc;
  size_t stride;
  int rc = e->run_predict(jobs, cj, cc, stride, e->grade);
  if (rc) return rc;
  double *cdst = order.empty() ? cost : cost_sorted.data();
  rc = e->run_cost(cost_kind, jobs, cj, cc, stride, cdst);
  if (rc) return rc;
  if (e->grade) {
    // chains whose look-ahead met the +-10 weight clamp are not search-grade exact: those jobs go through the canonical kernels
    // ... those that could be accepted, that is: a candidate whose weights run into the clamp is almost always a poor one, and a
    // cost at or above the search's incumbent is a failure whatever its exact value
    std::vector<int> redo;
    for (int j = 0; j < njobs; j++) {
      if (!e->job_inexact[j]) continue;
      const int orig = order.empty() ? j : order[j];
      if (e->redo_below && cdst[j] >= e->redo_below[orig]) continue;
      redo.push_back(j);
    }
    if (!redo.empty()) {
      e->sg_stats[3] += (long long)redo.size();
      std::vector<Job> rj;
      for (int j : redo) rj.push_back(jobs[j]);
      std::vector<double> rcost(redo.size());
      rc = e->run_predict(rj, cj, cc, stride, 0);
      if (rc) return rc;
      rc = e->run_cost(cost_kind, rj, cj, cc, stride, rcost.data());
      if (rc) return rc;
      for (size_t i = 0; i < redo.size(); i++) cdst[redo[i]] = rcost[i];
    }
  }
  if (!order.empty()) for (int i = 0; i < njobs; i++) cost[order[i]] = cost_sorted[i];
  return SAC_OK;
}

int sac_profile_params(const float *profile58, double *out, int cap)
{
  if (!profile58 || !out || cap < 59) { set_error("sac_profile_params: bad argument"); return SAC_E_ARG; }
  const HostParam p = map_profile(profile58);
  int o = 0;
  const int ints[8] = {p.nA, p.nB, p.nM0, p.nS0, p.nS1, p.ch_ref, p.lm_n, p.bias_scale};
  for (int v : ints) out[o++] = v;
  for (int c = 0; c < 2; c++) for (int i = 0; i < 4; i++) out[o++] = p.vn[c][i];
  for (int c = 0; c < 2; c++) for (int i = 0; i < 4; i++) out[o++] = p.vmu[c][i];
  for (int c = 0; c < 2; c++) for (int i = 0; i < 4; i++) out[o++] = p.vmudecay[c][i];
  for (int c = 0; c < 2; c++) for (int i = 0; i < 4; i++) out[o++] = p.vpowdecay[c][i];
  for (int c = 0; c < 2; c++) out[o++] = p.lambda[c];
  for (int c = 0; c < 2; c++) out[o++] = p.ols_nu[c];
  for (int c = 0; c < 2; c++) out[o++] = p.mu_mix[c];
  for (int c = 0; c < 2; c++) out[o++] = p.mu_mix_beta[c];
  for (int c = 0; c < 2; c++) { out[o++] = p.beta_sum[c]; out[o++] = p.beta_pow[c]; out[o++] = p.beta_add[c]; }
  for (int c = 0; c < 2; c++) out[o++] = p.bias_mu[c];
  out[o++] = p.lm_alpha;
  for (int c = 0; c < 2; c++) out[o++] = p.proj_alpha[c];
  return o;                                                          // 59
}

int sac_eval_population(sac_engine *h, const sac_window *w, int from, int n, const float *base_profile, const int *dims, int D,
                        const double *X, int P, int cost_kind, int optk, double *cost)
{
  if (P <= 0 || !base_profile) { set_error("sac_eval_population: bad argument"); return SAC_E_ARG; }
  std::vector<const sac_window *> wins(P, w);
  std::vector<int> fr(P, from), nn(P, n);
  std::vector<float> bases((size_t)P * kProfileSize);
  for (int p = 0; p < P; p++) std::memcpy(&bases[(size_t)p * kProfileSize], base_profile, sizeof(float) * kProfileSize);
  return sac_eval_jobs(h, P, wins.data(), fr.data(), nn.data(), bases.data(), dims, D, X, cost_kind, optk, cost);
}

int sac_cost(sac_engine *h, int cost_kind, const int32_t *bufs, int count, int n, double *cost)
{
  Engine *e = reinterpret_cast<Engine *>(h);
  if (!e || !bufs || count <= 0 || n <= 0 || !cost) { set_error("sac_cost: bad argument"); return SAC_E_ARG; }
  e->begin_call();
  SACB_CUDA(cudaSetDevice(e->device));
  // one mono pseudo-window so that run_cost's bookkeeping applies; residuals are uploaded in place of predictor output
  Window w; w.eng = e; w.device = e->device; w.stream = e->stream; w.nch = 1; w.numsamples = n; w.d_planes[0] = w.d_planes[1] = nullptr;
  int32_t mn = 0, mx = 0;
  for (size_t i = 0; i < (size_t)count * n; i++) { mn = std::min(mn, bufs[i]); mx = std::max(mx, bufs[i]); }
  const int32_t R = std::max(-(long long)mn, (long long)mx) > 0 ? (int32_t)std::max(-(long long)mn, (long long)mx) : 1;
  w.minmax[0] = 0; w.minmax[1] = R; w.minmax[2] = 0; w.minmax[3] = R;
  std::vector<Job> jobs(count);
  std::vector<int> cj(count), cc(count, 0);
  for (int c = 0; c < count; c++) { jobs[c].win = &w; jobs[c].from = 0; jobs[c].n = n; jobs[c].k = 1; cj[c] = c; }
  const size_t stride = ((size_t)n + 31) & ~size_t(31);
  SACB_CUDA(e->d_resid.reserve(stride * count));
  SACB_CUDA(e->d_sums.reserve((size_t)3 * count));
  SACB_CUDA(e->d_flags.reserve((size_t)4 * count));
  SACB_CUDA(cudaMemsetAsync(e->d_flags.p, 0, sizeof(int) * 4 * count, e->stream));
  std::vector<long long> sums(3 * (size_t)count, 0);
  for (int c = 0; c < count; c++) {
    const int32_t *b = bufs + (size_t)c * n;
    long long l1 = 0, sq = 0;
    for (int i = 0; i < n; i++) { l1 += std::llabs((long long)b[i]); sq += (long long)b[i] * b[i]; }
    sums[c] = l1; sums[count + c] = sq;
    SACB_CUDA(cudaMemcpyAsync(e->d_resid.p + (size_t)c * stride, b, sizeof(int32_t) * n, cudaMemcpyHostToDevice, e->stream));
  }
  SACB_CUDA(cudaMemcpyAsync(e->d_sums.p, sums.data(), sizeof(long long) * 3 * count, cudaMemcpyHostToDevice, e->stream));
  SACB_CUDA(cudaEventRecord(e->ev[0], e->stream));
  SACB_CUDA(cudaEventRecord(e->ev[1], e->stream));
  SACB_CUDA(cudaStreamSynchronize(e->stream));
  e->slot_of.resize(count); e->slot_rep.resize(count); e->nslots = count;
  for (int c = 0; c < count; c++) { e->slot_of[c] = c; e->slot_rep[c] = c; }
  return e->run_cost(cost_kind, jobs, cj, cc, stride, cost);
}

} // extern "C"
