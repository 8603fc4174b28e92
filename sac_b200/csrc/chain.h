// chain.h -- descriptors shared by host glue and kernels.
//
// A "chain" is one serial fp64 recurrence: one coded channel of one candidate parameter vector over one frame
// window (the reference's Predictor for one channel, /root/reference src/libsac/pred.cpp:4-45). In the encoder the
// two channels of a stereo frame are independent chains because both regressors read input samples only
// (SURVEY.md fact 3); in the decoder they are coupled through progress counters.
#ifndef SAC_B200_CHAIN_H
#define SAC_B200_CHAIN_H
#include <stdint.h>

namespace sacb {

enum { kStages = 4, kMixN = 5, kMaxOls = 96, kMaxRls = 10, kProfileSize = 58, kSearchDims = 56 };
enum { kTapThreads = 128, kTapWarps = 4, kBlockThreads = 160 };
enum CostKind { kCostL1 = 0, kCostRMS = 1, kCostEntropy = 2, kCostGolomb = 3, kCostBitplane = 4 };

struct ChainDesc {
  // --- sample planes (device pointers to the start of the window, i.e. plane + from) ---
  const int32_t *own;       // this chain's channel
  const int32_t *other;     // the other channel (== own for mono)
  int32_t *own_out;         // decoder: decoded samples are written here (== own); encoder: null
  const int32_t *err_in;    // decoder: residuals to add; encoder: null
  int32_t *resid;           // encoder: residual output [n]; decoder: null
  int n;                    // window length
  // --- OLS regressor x = [own[t-lenA..t-1] | other[sB..sB+lenB-1]], sB = max(t-lagB, minB)-backB (pred.cpp:17-31)
  int lenA, lenB, lagB, minB, backB;
  int k;                    // solve interval (libsac.cpp:39-40)
  double lambda, nu, beta_sum, beta_pow, beta_add;   // nu already scaled by (1-lambda) (ols.cpp:11)
  // --- NLMS cascade (libsac.cpp:44-61) ---
  int vn[kStages];
  double vmu[kStages], vmudecay[kStages], vpowdecay[kStages];
  double mu_mix, mix_beta;
  int lm_n;
  double lm_gamma, proj_alpha;
  double casc_lo, casc_hi;  // Cascade Range (follows the coded-channel index, libsac.cpp:98-102)
  // --- bias stage ---
  double bias_mu;
  int bias_nscale;
  int clamp_lo, clamp_hi;   // residual clamp (follows the actual channel, libsac.cpp:106)
  // --- spill space in HBM for arrays that do not fit the CTA's shared memory ---
  double *scratch;
  long long scratch_doubles;
  double *scratch_ols;      // encode direction: OLS matrices that do not fit the OLS kernel's shared memory
  double *plpc;             // encode direction: OLS predictions p_lpc[n], written by ols_kernel, read by cascade_kernel
  // --- outputs ---
  long long *l1sum;         // sum |e|
  long long *sqsum;         // sum e*e
  int *flags;               // bit0: a prediction became non-finite (cascade.h:40)
  // --- decoder coupling: wait until *wait_ctr >= need(t) before sample t; publish own progress ---
  int *progress;            // number of decoded samples of this channel
  const int *wait_ctr;      // progress of the other channel (null: no dependency)
  int wait_add;             // need(t) = min(n, max(t + wait_add, 0))  (ch0: -lag; ch1: +nS1)
  // --- decoder, sparse-PCM channel: err_in holds rank distances among the used sample values (libsac.cpp:157-160,
  //     map.cpp:188-202); cumulative counts and the sorted used values as built by sparse.cu. null: plain residuals ---
  const int32_t *unmap_cum;
  const int32_t *unmap_list;
  int unmap_mean;           // the prediction is re-based to raw sample values: pi + mean
};

} // namespace sacb
#endif
