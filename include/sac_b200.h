/* sac_b200.h -- C ABI of libsac_b200.so, the B200-native encode path of the Sac lossless audio codec.
 *
 * The reference (slmdev/sac v0.7.25, one CPU binary) has no plugin interface; this ABI sits behind the two seams of
 * its frame coder and is what a cgo/ctypes/C++ host binds:
 *
 *   population seam   Opt::eval_points_mt(func, points)          /root/reference src/opt/opt.h:25, src/opt/opt.cpp:11-43
 *                     with func = FrameCoder::Optimize's cost lambda  src/libsac/libsac.cpp:389-397
 *                     -> sac_eval_population / sac_eval_jobs
 *   search            Opt::run(func, xstart) for OptDDS          src/opt/opt.h:20, src/opt/dds.cpp:33-119
 *                     -> sac_dds_run (host; any evaluator through a callback)
 *   frame seam        FrameCoder::{SetNumSamples,Predict,Encode,WriteEncoded}  src/libsac/libsac.h:45-56
 *                     as driven by Codec::EncodeFile             src/libsac/libsac.cpp:822-829
 *                     -> sac_frames_encode ; inverse (ReadEncoded,Decode,Unpredict) -> sac_frame_decode
 *   file              Codec::EncodeFile / DecodeFile             src/libsac/libsac.cpp:782-883
 *                     -> sac_encode_file / sac_decode_file
 *
 * Conventions: plain pointers and sizes, caller owns every buffer, integer status (0 = ok, <0 = error, see
 * sac_last_error()), no exceptions cross the boundary. All sample planes are planar int32 HOST memory unless a
 * function says otherwise. Every compute entry point needs a CUDA device (sm_100a); without one sac_engine_create
 * fails with SAC_E_NODEVICE -- there is no CPU fallback.
 */
#ifndef SAC_B200_H
#define SAC_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SAC_PROFILE_SIZE 58   /* src/libsac/profile.cpp:10 */
#define SAC_SEARCH_DIMS 56    /* all but coefficients 56,57 (src/libsac/libsac.cpp:469-475) */

enum { SAC_OK = 0, SAC_E_NODEVICE = -1, SAC_E_CUDA = -2, SAC_E_ARG = -3, SAC_E_IO = -4, SAC_E_FORMAT = -5, SAC_E_UNSUPPORTED = -6,
       SAC_E_MD5 = -7 /* decoded audio does not hash to the MD5 in the header: no output is delivered */ };
/* Byte 17 of the .sac header (reserved, 0 in files of the reference, src/file/sac.cpp:15-38) names the ARITHMETIC the frame
 * payloads were predicted in: the decoder re-runs the predictor and must repeat the encoder's fp64 bit pattern (SURVEY.md
 * fact 1). 0 = some build of the reference (libm, compiler-dependent contraction), 1 = this library's canonical arithmetic
 * (include/sac_canon_math.h). Files written here carry 1; the reference ignores the byte. Decoding a variant-0 file is
 * attempted (short material often agrees) and, like every decode, succeeds only if the audio MD5 matches; other values are refused. */
#define SAC_ARITH_REFERENCE 0
#define SAC_ARITH_CANONICAL 1
/* FrameCoder::SearchCost (src/libsac/libsac.h:14) */
enum { SAC_COST_L1 = 0, SAC_COST_RMS = 1, SAC_COST_ENTROPY = 2, SAC_COST_GOLOMB = 3, SAC_COST_BITPLANE = 4 };

typedef struct sac_engine sac_engine; /* one per GPU: stream, HBM pools, model tables */
typedef struct sac_window sac_window; /* samples of one (sub)frame resident in HBM + its stats */

const char *sac_version(void);
const char *sac_last_error(void);
int sac_device_count(void);

/* ---- engine ---------------------------------------------------------------------------------------------------- */
sac_engine *sac_engine_create(int device);
void sac_engine_destroy(sac_engine *);
/* number of kernels this engine has launched since creation (bench.py's gpu_launches) */
long long sac_engine_launches(const sac_engine *);
/* device time (ms, CUDA events on the engine's stream) and launches of the last call, by kernel class:
 * [0] predictor (ols_kernel + cascade_kernel) [1] bitplane [2] entropy/other [3] ols_kernel alone; out_ms[4], out_launches[4] */
void sac_engine_last_timing(const sac_engine *, double *out_ms, long long *out_launches);

/* device time (ms) of every population evaluation since the engine was created, by kernel class, summed over the engine and its
 * helper engines (frames in flight): [0] OLS kernels [1] cascade kernels [2] bitplane cost kernel [3] other cost kernels. CUDA
 * events on each engine's own stream; with several streams in flight the spans overlap, so the SHARES are meaningful, the sum
 * exceeds the wall time. *out_calls (may be null): evaluations timed. */
void sac_engine_total_timing(const sac_engine *, double *out_ms4, long long *out_calls);

/* Exact de-duplication (on by default): chains of one call whose inputs are identical -- same planes, range, k and
 * channel parameters -- are evaluated once, and OLS stages with identical OLS parameters are computed once and shared
 * (late in a DDS search most candidates leave one channel untouched). Results are bit-identical with and without.
 * Returns the previous setting. sac_dedup_totals: {chains requested, chains evaluated, OLS stages evaluated} since
 * process start. */
int sac_engine_set_dedup(sac_engine *, int on);
void sac_dedup_totals(long long *out3);

/* Arithmetic of sac_predict / sac_eval_population / sac_eval_jobs on this engine: 0 = canonical (default; the decoder's bits),
 * 1 = search-grade kernels (same formulas; summation order, fused operations and schedule are free: costs equal the canonical
 * ones except where a residual flips at a rounding boundary). sac_frames_encode sets it from sac_cfg::grade for
 * the search and always runs the final pass canonically. Returns the previous value. flags of sac_predict: bit 0 non-finite
 * prediction, bit 1 (search grade only) a weight met the +-10 clamp under look-ahead (sac_eval_* re-evaluate such jobs canonically).
 * sac_engine_grade_stats: chains through the small / large search-grade cascade, the canonical fallback, jobs re-evaluated. */
int sac_engine_set_grade(sac_engine *, int grade);
void sac_engine_grade_stats(const sac_engine *, long long *out4);

/* measured DFMA throughput of the device (GFLOP/s, 2 flop per fma; CUDA events): the fp64 roofline denominator */
double sac_fp64_peak_gflops(sac_engine *);

/* ---- profile (SacProfile::LoadBaseProfile, src/libsac/profile.cpp:3-89) ------------------------------------------ */
int sac_base_profile(float *vmin, float *vmax, float *vdef); /* 58 each; returns 58 */

/* LogDomain stretch[32768] / squash[4095] tables of the bitplane model (src/model/domain.h:7-61), canonical math, host only */
void sac_model_tables(int16_t *stretch, int16_t *squash);

/* ---- windows ---------------------------------------------------------------------------------------------------- */
/* planes[ch][numsamples] must already be mean-free (FrameCoder::Predict, libsac.cpp:452-458);
 * minmax = {min0,max0[,min1,max1]} after mean removal. Copies to HBM. */
sac_window *sac_window_create(sac_engine *, int nch, const int32_t *const *planes, int numsamples, const int32_t *minmax);
void sac_window_destroy(sac_window *);

/* ---- FrameCoder::PredictFrame (libsac.cpp:94-142) for P profiles at once ------------------------------------------
 * profiles[P][58]; window [from, from+n); k = OLS solve interval; resid[P][nch][n] (host) receives the residuals
 * in channel-index order. flags[P] (may be null): 1 if a prediction became non-finite. */
int sac_predict(sac_engine *, const sac_window *, const float *profiles, int P, int from, int n, int k, int32_t *resid,
                int *flags);

/* ---- CostFunction::Calc (src/libsac/cost.h) on host residual arrays: bufs[count][n] -> cost[count] ---------------- */
int sac_cost(sac_engine *, int cost_kind, const int32_t *bufs, int count, int n, double *cost);

/* ---- BitplaneCoder::Encode / Decode over RangeCoderSH (src/libsac/vle.h:51-53, src/model/range.h:45-60) -----------
 * resid[n]: signed residuals (S2U-mapped inside, as FrameCoder::CnvError_S2U). maxbpn_io: in <0 = derive from the
 * data (iLog2 of the largest mapped value), out = the value used. Payload to out (cap bytes), length to *out_len. */
int sac_bitplane_encode(sac_engine *, const int32_t *resid, int n, int *maxbpn_io, uint8_t *out, long long cap, long long *out_len);
int sac_bitplane_decode(sac_engine *, const uint8_t *payload, long long len, int n, int maxbpn, int32_t *resid_out);

/* FrameCoder::SetParam (src/libsac/libsac.cpp:37-92) alone, host only: the 58 profile floats -> the predictor's parameters
 * as this library hands them to its kernels. out[59] (doubles): nA,nB,nM0,nS0,nS1,ch_ref,lm_n,bias_scale | vn ch0[4],ch1[4] |
 * vmu[2][4] | vmudecay[2][4] | vpowdecay[2][4] | lambda[2] | ols_nu[2] | mu_mix[2] | mu_mix_beta[2] |
 * beta_sum,beta_pow,beta_add per channel | bias_mu[2] | lm_alpha | proj_alpha[2]. Returns 59. */
int sac_profile_params(const float *profile58, double *out, int cap);

/* ---- population evaluation -------------------------------------------------------------------------------------
 * cost[p] = sum over channels of Cost(PredictFrame(profile with dims[i] <- (float)X[p][i], window, k=optk)).
 * Non-finite predictor state gives cost = +INF. One call in flight per engine. */
int sac_eval_population(sac_engine *, const sac_window *, int from, int n, const float *base_profile, const int *dims,
                        int D, const double *X, int P, int cost_kind, int optk, double *cost);
/* the same over several windows in one batch: job j evaluates X[j] on wins[j] with bases[j] (58 floats each) */
int sac_eval_jobs(sac_engine *, int njobs, const sac_window *const *wins, const int *from, const int *n,
                  const float *bases, const int *dims, int D, const double *X, int cost_kind, int optk, double *cost);

/* ---- DDS (OptDDS, src/opt/dds.cpp) ------------------------------------------------------------------------------
 * eval(X[P][D], P, cost[P], user) fills costs; returns non-zero to abort. num_threads<=0: run_single (SSC0);
 * >0: run_mt with generations of num_threads (SSC1). mt19937 seed 0 as src/opt/opt.cpp:5. Returns best cost. */
typedef int (*sac_eval_fn)(const double *X, int P, int D, double *cost, void *user);
double sac_dds_run(int D, const double *xmin, const double *xmax, const double *xstart, int nfunc_max, int num_threads,
                   double sigma_init, sac_eval_fn eval, void *user, double *xbest);
/* run_single (num_threads == 0, the reference's default search) in speculative batches: up to `spec` candidates per call of
 * eval are drawn as run_single would draw them if each of them failed; costs are walked in order, the first success is
 * applied, the generator rewound to that point and the rest of the batch dropped. Accepted sequence, step sizes and
 * result equal sac_dds_run(..., num_threads = 0, ...); *evaluated (may be null) receives the candidates spent. */
double sac_dds_run_spec(int D, const double *xmin, const double *xmax, const double *xstart, int nfunc_max, double sigma_init,
                        int spec, sac_eval_fn eval, void *user, double *xbest, long long *evaluated);
/* OptDE::run (src/opt/de.cpp:80-172; NP = 30, current-to-pbest/1/bin, adaptive CR and F) over the same evaluator seam:
 * the start vector, then the 29 initial samples, then generations of up to 30 trial vectors */
double sac_de_run(int D, const double *xmin, const double *xmax, const double *xstart, int nfunc_max, double sigma_init,
                  sac_eval_fn eval, void *user, double *xbest);
/* OptCMA::run (src/opt/cma.cpp:55-92): (1+1)-CMA-ES, one evaluation per step (P is always 1) */
double sac_cma_run(int D, const double *xmin, const double *xmax, const double *xstart, int nfunc_max, double sigma_init,
                   sac_eval_fn eval, void *user, double *xbest);

/* ---- frame coding ------------------------------------------------------------------------------------------------ */
enum { SAC_SEARCH_DDS = 0, SAC_SEARCH_DE = 1, SAC_SEARCH_CMA = 2 };   /* FrameCoder::SearchMethod (src/libsac/libsac.h) */
/* Codec::Analyse for one read of `numsamples` raw samples (host only, no GPU): 3-s blocks are classed by the sparse-PCM
 * cost ratio (> 1.35), runs of equal class become sub-frames (start/length/state, at most `cap` written). Returns
 * the number of sub-frames. */
int sac_analyse_subframes(int nch, const int32_t *const *planes, int numsamples, int samplerate, int *start, int *length,
                          int *state, int cap);

typedef struct sac_cfg {        /* FrameCoder::tsac_cfg / toptim_cfg (src/libsac/libsac.h:19-44) */
  int optimize;                 /* 0 = --normal */
  double fraction;              /* window fraction of max_framesize */
  int maxnfunc;                 /* DDS evaluations per frame */
  int num_threads;              /* DDS generation size (0 = sequential run_single) */
  double sigma;                 /* initial DDS radius */
  int optk;                     /* OLS solve interval during search (4) */
  int cost_kind;                /* SAC_COST_* */
  int reset;                    /* --opt-reset: every frame starts from the base profile */
  int zero_mean;                /* 1 */
  int sparse_pcm;               /* 1 (default): a channel whose sample values are sparse is also coded as rank distances among the used values behind the coded value map, and the shorter record is kept (libsac.cpp:253-278, map.cpp) */
  int max_framelen;             /* seconds (20) */
  int adapt_block;              /* adaptive sub-frame split of every max_framelen read (Codec::Analyse, libsac.cpp:726-780) */
  int frame_parallel;           /* B200 extension (implies --opt-reset semantics): 1 = all frames of a call share each generation's launches,
                                   2 = every frame runs on its own stream / host thread, all concurrently on the GPU */
  int verbose;
  int search;                   /* SAC_SEARCH_*: --opt-cfg=dds|de|cma (src/cmdline.cpp:195-206) */
  int spec;                     /* B200 extension, sequential DDS only (num_threads == 0, the reference's default): candidates per
                                   speculative batch (sac_dds_run_spec). Results do not depend on it; 1 = one candidate per launch */
  int inflight;                 /* frame_parallel == 2: frames in flight on the GPU (1..64) */
  int grade;                    /* arithmetic of the SEARCH evaluations: 0 = canonical (bit-exact with the decoder's),
                                   1 = search-grade kernels (same formulas, free summation order / fused operations; the final
                                   pass and the bitstream are always canonical) */
} sac_cfg;
void sac_cfg_default(sac_cfg *);
/* presets of src/cmdline.cpp:127-156: "normal","high","veryhigh","extrahigh","best","insane" */
int sac_cfg_preset(sac_cfg *, const char *name);

/* Predict()+Encode()+WriteEncoded() for nframes frames of one stream. planes[f*nch+ch] -> raw samples of frame f
 * (numsamples[f] each). profile_io[58]: in = start profile of the first frame, out = profile after the last.
 * Frame records are appended to out (cap bytes); *out_len receives the total. */
int sac_frames_encode(sac_engine *, const sac_cfg *, int nch, int max_framesize, int nframes,
                      const int32_t *const *planes, const int *numsamples, float *profile_io, uint8_t *out, long long cap,
                      long long *out_len);
/* the same with the frames already resident in HBM (sac_window_create on mean-free planes; means[f*nch+ch] go into
 * the block headers): the timed region of bench.py's device-resident throughput */
int sac_frames_encode_resident(sac_engine *, const sac_cfg *, int nch, int max_framesize, int nframes,
                               const sac_window *const *wins, const int32_t *means, float *profile_io, uint8_t *out,
                               long long cap, long long *out_len);
/* one frame record -> samples (mean restored). Returns bytes consumed (>0) or <0. planes_out[ch][cap_samples]. */
long long sac_frame_decode(sac_engine *, int nch, const uint8_t *in, long long len, int32_t *const *planes_out,
                           int cap_samples, int *numsamples);

/* ---- files (Codec::EncodeFile / DecodeFile, WAV and .sac containers of src/file/) ------------------------------- */
typedef struct sac_file_stats { long long in_bytes, out_bytes; int numsamples, nch, samplerate, bits, nframes; double seconds; uint8_t md5[16]; int md5_ok; } sac_file_stats;
int sac_encode_file(sac_engine *, const sac_cfg *, const char *wav_path, const char *sac_path, sac_file_stats *);
int sac_decode_file(sac_engine *, const char *sac_path, const char *wav_path, sac_file_stats *);
/* Several GPUs of one box from the C++ host, one engine per device (sac_engine_create(d)), no collective:
 *   sac_encode_file_multi  ONE file; its frames are dealt round-robin to the engines, each engine's share in flight on its GPU,
 *                          records concatenated in frame order (libsac.cpp:565-578). Implies --opt-reset semantics (frames are
 *                          independent searches, cmdline.cpp:193); the bytes equal sac_encode_file's with reset = 1 on one GPU.
 *   sac_encode_files       a BATCH of files (Codec::EncodeFile per file, libsac.cpp:782-855): every engine takes the next file of a
 *                          longest-first queue. status[i] (may be null) receives each file's code; returns the first failure.
 *                          Two inputs that map to the same output path are refused. */
int sac_encode_file_multi(sac_engine *const *engines, int nengines, const sac_cfg *, const char *wav_path, const char *sac_path,
                          sac_file_stats *);
int sac_encode_files(sac_engine *const *engines, int nengines, const sac_cfg *, int nfiles, const char *const *wav_paths,
                     const char *const *sac_paths, sac_file_stats *stats, int *status);
/* Frame prologue of one channel (FrameCoder::AnalyseMonoChannel + zero-mean, libsac.cpp:626-651, 452-458), host only:
 * out3 = {mean (0 when !zero_mean), min, max} with min/max taken after the mean is removed -- the block header's fields. */
int sac_frame_stats(const int32_t *samples, int n, int zero_mean, int32_t *out3);
/* Host part of Codec::EncodeFile alone (libsac.cpp:782-835; wav.cpp:167-263, sac.cpp:15-38; NO GPU needed): parses the
 * WAV image and writes what precedes the first frame record of the .sac file -- header, metadata (all WAV chunks), MD5 of
 * the PCM bytes -- to out[0,*out_len); frame_lengths[0,st->nframes) receives the samples per frame record (reads of
 * max_framelen seconds, split by Codec::Analyse when cfg->adapt_block). sac_encode_memory = this + sac_frames_encode. */
int sac_container_plan(const sac_cfg *, const uint8_t *wav, long long wav_len, uint8_t *out, long long cap, long long *out_len,
                       int *frame_lengths, int cap_frames, sac_file_stats *);
/* in-memory variants (host buffers) used by bench.py's e2e leg */
int sac_encode_memory(sac_engine *, const sac_cfg *, const uint8_t *wav, long long wav_len, uint8_t *out, long long cap,
                      long long *out_len, sac_file_stats *);
int sac_decode_memory(sac_engine *, const uint8_t *sac, long long sac_len, uint8_t *out, long long cap, long long *out_len,
                      sac_file_stats *);

#ifdef __cplusplus
}
#endif
#endif
