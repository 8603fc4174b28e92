// engine.h -- the per-GPU engine behind the C ABI: HBM pools, model tables, batched launches.
#ifndef SAC_B200_ENGINE_H
#define SAC_B200_ENGINE_H
#include "bitplane.h"
#include "chain.h"
#include "sparse.h"
#include "sac_b200.h"
#include <cuda_runtime.h>
#include <algorithm>
#include <string>
#include <vector>

namespace sacb {

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);
#define SACB_CUDA(call)                                              \
  do {                                                               \
    cudaError_t e__ = (call);                                        \
    if (e__ != cudaSuccess) return ::sacb::cuda_fail(e__, #call);    \
  } while (0)

// Grow-only device buffer, STREAM-ORDERED (cudaMallocAsync / cudaFreeAsync on the owning engine's stream): cudaFree is a
// device-wide synchronisation, and with ten frames in flight (a host thread, an engine and pools each) every pool that grew
// and every frame that finished stalled all the others behind the longest kernel on the device (seconds). All users of a
// buffer are ordered after the engine's stream (side streams fork from it by events), so stream order is sufficient.
template <class T> struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;
  cudaStream_t s = nullptr;
  cudaError_t reserve(size_t n)
  {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeAsync(p, s);
    p = nullptr; cap = 0;
    size_t want = n + n / 4;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&p), want * sizeof(T), s);
    if (e != cudaSuccess) { (void)cudaGetLastError(); want = n; e = cudaMallocAsync(reinterpret_cast<void **>(&p), want * sizeof(T), s); }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeAsync(p, s); p = nullptr; cap = 0; }
};
template <class T> struct PinBuf {
  T *p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n)
  {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    const size_t want = std::max<size_t>(2 * n, 4096 / sizeof(T));   // pinned allocations synchronise too: grow rarely
    cudaError_t e = cudaMallocHost(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Window {           // sac_window
  struct Engine *eng;
  int device;             // kept here: the window may outlive its engine
  cudaStream_t stream;    // the planes are allocated stream-ordered on the creating engine's stream
  int nch, numsamples;
  int32_t minmax[4];
  int32_t *d_planes[2];   // HBM, mean-free
};

// one evaluation job: a profile on a window range
struct Job {
  const Window *win;
  int from, n, k;
  float profile[kProfileSize];
};

struct Engine {           // sac_engine
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // ev[4]: between ols_kernel and cascade_kernel
  BitplaneTables bt;
  int smem_bytes = 72 * 1024;        // decode-direction kernel (predictor.cu)
  int enc_smem_bytes = 100 * 1024;   // cascade kernel (predictor_enc.cu): cap of the per-launch request, 2 CTAs per SM
  int ols_smem_bytes = 28 * 1024;    // OLS kernel, minimum request (both matrices up to n = 32)
  int ols_smem_cap_bytes = 100 * 1024; // OLS kernel: above this the covariance goes to HBM scratch, the work matrix stays
  long long launches = 0;
  double last_ms[4] = {0, 0, 0, 0};          // predictor (ols + cascade), bitplane, other cost kernels, ols alone
  double total_ms[4] = {0, 0, 0, 0};         // since creation, CUDA events on this engine's stream: ols, cascade, bitplane, other cost kernels
  long long total_calls = 0;                 // population evaluations timed into total_ms
  bool ev4_recorded = false;                 // ev[4] (between OLS and cascade) belongs to the call being timed
  // side streams: the kernel classes of one evaluation (OLS by order, cascade by size) run side by side, fork / join by events
  enum { kSide = 4 };
  cudaStream_t side[kSide] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[kSide] = {nullptr, nullptr, nullptr, nullptr};
  long long last_launches[4] = {0, 0, 0, 0};

  DevBuf<ChainDesc> d_descs;
  PinBuf<ChainDesc> h_descs;
  DevBuf<double> d_scratch;
  DevBuf<double> d_scratch_ols;
  DevBuf<double> d_plpc;          // OLS predictions per chain (encode direction)
  DevBuf<int32_t> d_resid;
  DevBuf<long long> d_sums;       // per chain: l1, sq, nbytes
  DevBuf<int> d_flags;            // per chain: flags, maxbpn
  PinBuf<long long> h_sums;
  PinBuf<int> h_flags;
  DevBuf<BpJob> d_bpjobs;
  PinBuf<BpJob> h_bpjobs;
  DevBuf<uint32_t> d_csig0;
  DevBuf<unsigned int> d_hist;
  DevBuf<double> d_cost;
  PinBuf<double> h_cost;
  DevBuf<uint8_t> d_bytes;        // payload / msb scratch
  PinBuf<int32_t> h_stage;        // pinned staging for sample uploads
  DevBuf<uint8_t> d_sparse;       // sparse-PCM side of a final pass / a decode: coder hand-over, used map, counts, mapped residuals, payload
  PinBuf<SparseOut> h_sparse;

  // helper engines: own stream (high priority) and pools on the same device, used for the final pass of a frame
  // while the main stream already searches the next frame; they share this engine's model tables
  std::vector<Engine *> helpers;
  bool is_helper = false;
  Engine *helper(int i);
  long long total_launches() const;

  // waits for the stream without spinning (a blocking-sync event): the waits of this path last seconds, and several
  // host threads (frames in flight, one rank per GPU) share the host cores
  cudaEvent_t ev_wait = nullptr;
  cudaError_t wait();

  // arithmetic of population evaluations (sac_cfg::grade): 1 = search-grade kernels (predictor_sg.cu); final passes and the
  // decoder always use the canonical kernels
  int grade = 0;
  DevBuf<int> d_idx;              // descriptor indices per kernel class (search-grade launches)
  PinBuf<int> h_idx;
  const double *redo_below = nullptr;   // set by the search driver around sac_eval_jobs: per job, the cost below which an inexact (clamp-flagged)
                                        // search-grade result matters -- only those jobs are re-evaluated canonically; null: all of them
  std::vector<char> job_inexact;  // per job of the last run_cost: a chain hit the weight clamp under look-ahead (re-evaluate canonically)
  long long sg_stats[4] = {0, 0, 0, 0};   // chains through cascade_sg small / large / canonical fallback, jobs re-evaluated after a clamp

  // chain de-duplication of the last run_predict (see engine.cu): logical chain -> slot, slot -> a representative
  bool dedup = true;
  std::vector<int> slot_of, slot_rep;
  int nslots = 0;
  long long last_unique[3] = {0, 0, 0};   // logical chains, unique chains, unique OLS stages of the last launch

  int init(int dev, const Engine *parent = nullptr);
  void destroy();
  // residuals of every chain of `jobs` into d_resid (layout: chain c at c*stride); returns 0
  int run_predict(const std::vector<Job> &jobs, std::vector<int> &chain_job, std::vector<int> &chain_ch, size_t &stride, int grade = 0);
  // per-job cost (sum over channels) for residuals left by run_predict
  int run_cost(int cost_kind, const std::vector<Job> &jobs, const std::vector<int> &chain_job, const std::vector<int> &chain_ch,
               size_t stride, double *cost);
  void begin_call();
};

// kernels (predictor.cu, cost.cu)
cudaError_t launch_predictor_decode(const ChainDesc *d_descs, int nchains, int smem_bytes, cudaStream_t stream);
cudaError_t launch_predictor_enc(const ChainDesc *d_ols_descs, int nols, const ChainDesc *d_descs, int nchains, int smem_bytes,
                                 int ols_smem_bytes, cudaStream_t stream, cudaEvent_t between = nullptr);
long long predictor_enc_scratch_doubles(const int *vn, int n_ols);
long long predictor_ols_scratch_doubles(int n_ols);
size_t predictor_ols_shared_bytes();
long long predictor_enc_smem_doubles(const int *vn);
cudaError_t predictor_enc_init_attributes();
cudaError_t launch_cascade_canonical(const ChainDesc *d_descs, const int *d_idx, int count, int smem_bytes, cudaStream_t stream);
cudaError_t launch_ols_canonical(const ChainDesc *d_descs, const int *d_idx, int count, int smem_bytes, cudaStream_t stream);
// search-grade kernels (predictor_sg.cu)
cudaError_t predictor_sg_init_attributes();
size_t cascade_sg_smem_bytes(const int *vn, int large);          // 0: the chain does not fit that variant
int ols_sg_class(int n_ols);                                     // 16 / 24 / 32 (one warp per chain), 64 (experiment), 0 = canonical kernel
size_t ols_sg_smem_bytes(int n_ols);
cudaError_t launch_ols_sg(const ChainDesc *d_descs, const int *d_idx, int count, int nb_class, int smem_bytes, cudaStream_t stream);
cudaError_t launch_cascade_sg(const ChainDesc *d_descs, const int *d_idx, int count, int large, int smem_bytes, cudaStream_t stream);
cudaError_t predictor_init_attributes();
size_t predictor_enc_shared_bytes();
cudaError_t launch_entropy(const int32_t *resid, size_t stride, const int *ns, const int *ranges, int nchains, unsigned int *hist,
                           size_t hist_stride, double *out, cudaStream_t stream);
cudaError_t launch_dfma_peak(double *out, int blocks, int iters, cudaStream_t stream);
cudaError_t launch_golomb(const int32_t *resid, size_t stride, const int *ns, int nchains, double *out, cudaStream_t stream);

// host-side parameter mapping (FrameCoder::SetParam, libsac.cpp:37-92)
struct HostParam {
  int nA, nB, nM0, nS0, nS1, ch_ref, lm_n, bias_scale;
  int vn[2][4];
  double vmu[2][4], vmudecay[2][4], vpowdecay[2][4];
  double lambda[2], ols_nu[2], mu_mix[2], mu_mix_beta[2], beta_sum[2], beta_pow[2], beta_add[2];
  double bias_mu[2], lm_alpha, proj_alpha[2];
};
HostParam map_profile(const float *profile);
long long chain_scratch_doubles(const HostParam &hp, int coded_ch, int nch);
// fills the encoder descriptor of coded channel `cc` (0/1); returns the actual channel index it codes
int fill_chain(ChainDesc &d, const HostParam &hp, int nch, int cc, int k, const int32_t *const *planes, int from, int n,
               const int32_t *minmax);

extern const float kBaseProfile[kProfileSize][3];

} // namespace sacb
#endif
