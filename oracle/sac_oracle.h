/* TEST INFRASTRUCTURE ONLY -- never linked into, loaded by or executed from the product path.
 *
 * sac_oracle: CPU restatement of the reference's (slmdev/sac v0.7.25, /root/reference) per-frame encode path --
 * DDS search, OLS + NLMS cascade + bias predictor, cost functions, bitplane model, range coder, frame
 * serialisation -- in plain scalar C++. Every function cites the reference file:line it follows.
 *
 * Parity pin: with order=SACO_ORDER_REF and math=SACO_MATH_LIBM the restatement reproduces, bit for bit, the
 * reference's own classes compiled here from /root/reference with -ffp-contract=off (oracle/_ref/libsacref_nc.so,
 * built by oracle/Makefile) and the goldens of BASELINE.md section 2 (tests/test_oracle_pin.py,
 * tests/golden/). With order=SACO_ORDER_B200 and math=SACO_MATH_CANON it evaluates the same algorithm in the
 * summation order and elementary functions the CUDA kernels use (DESIGN.md "canonical arithmetic"), which is what
 * the -m gpu parity tests compare against bit-exactly.
 */
#ifndef SAC_ORACLE_H
#define SAC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { SACO_ORDER_REF = 0, SACO_ORDER_B200 = 1 };
enum { SACO_MATH_LIBM = 0, SACO_MATH_CANON = 1 };
enum { SACO_COST_L1 = 0, SACO_COST_RMS = 1, SACO_COST_ENTROPY = 2, SACO_COST_GOLOMB = 3, SACO_COST_BITPLANE = 4 };

void saco_set_modes(int order, int math);

/* profile (src/libsac/profile.cpp:3-89): fills 58 entries, returns 58 */
int saco_base_profile(float *vmin, float *vmax, float *vdef);

/* FrameCoder::PredictFrame (src/libsac/libsac.cpp:94-142), encoder direction.
 * planes: s0[, s1] already mean-free; window [from, from+n); minmax = {min0,max0,min1,max1}; k = OLS solve interval.
 * Residuals to e0[, e1] (n each). Returns 0, or 1 if an expert prediction became non-finite (cascade.h:40). */
int saco_predict_frame(int nch, const int32_t *s0, const int32_t *s1, int from, int n, const float *profile58,
                       int k, const int32_t *minmax, int32_t *e0, int32_t *e1);
/* FrameCoder::UnpredictFrame (libsac.cpp:144-199) without the sparse-map branch: residuals -> samples (mean-free) */
int saco_unpredict_frame(int nch, int n, const float *profile58, const int32_t *minmax, const int32_t *e0,
                         const int32_t *e1, int32_t *s0, int32_t *s1);

/* cost functions (src/libsac/cost.h) */
double saco_cost(int kind, const int32_t *buf, int n);

/* BitplaneCoder + RangeCoderSH (src/libsac/vle.cpp, src/model/range.cpp). ubuf = S2U-mapped residuals. */
int saco_bitplane_encode(const int32_t *ubuf, int n, int maxbpn, uint8_t *out, int cap);
void saco_bitplane_decode(const uint8_t *in, int nbytes, int n, int maxbpn, int32_t *out_signed);
int saco_predict_laplace(uint32_t avg_sum, int bpn);
void saco_logdomain_tables(int *fwd /*[32768]*/, int *inv /*[4095]*/);
int saco_range_encode(const uint16_t *p1s, const uint8_t *bits, int n, uint8_t *out, int cap);

/* OptDDS (src/opt/dds.cpp) with std::mt19937(0) as the reference (src/opt/opt.cpp:5) */
typedef double (*saco_cost_cb)(const double *x, int n, void *user);
double saco_dds_run(int ndim, const double *xmin, const double *xmax, const double *xstart, int nfunc_max,
                    int num_threads, double sigma_init, saco_cost_cb cb, void *user, double *xbest);

/* FrameCoder::Predict + Encode + WriteEncoded for one frame without search (optimize=0) or with a sequential /
 * population DDS search; writes the frame record (Appendix A of SURVEY.md) to out, returns its length.
 * samples are the raw planes (mean is removed inside, libsac.cpp:443-459). profile_io: in = start profile,
 * out = profile used. cfg: {optimize, fraction*1e6, maxnfunc, num_threads, sigma*1e6, optk, cost_kind, max_framesize} */
int saco_encode_frame(int nch, int n, const int32_t *s0, const int32_t *s1, float *profile_io, const int *cfg,
                      uint8_t *out, int cap);
/* same with the sparse-PCM choice explicit (tsac_cfg::sparse_pcm, libsac.h:34; saco_encode_frame uses the default 1):
 * a channel is coded rank-mapped when sum|e| / sum|Map(e)| > 1.05 and the mapped record is shorter (libsac.cpp:253-278);
 * mapped_out[ch] (optional) tells which. */
int saco_encode_frame2(int nch, int n, const int32_t *s0, const int32_t *s1, float *profile_io, const int *cfg,
                       int sparse_pcm, uint8_t *out, int cap, int *mapped_out);
/* sparse-PCM pieces (libsac/map.cpp): the coded used-value map of raw samples (Remap::Analyse + MapEncoder::Encode),
 * its decoder (flags by magnitude, [32769] each), and Remap::Map / Unmap element-wise against the map of `raw` */
int saco_map_encode(const int32_t *raw, int n, uint8_t *out, int cap);
void saco_map_decode(const uint8_t *in, int nbytes, uint8_t *usedl, uint8_t *usedh);
void saco_remap(const int32_t *raw, int nraw, const int32_t *pred, const int32_t *err, int n, int unmap, int32_t *out);
/* inverse: parses one frame record (mapped blocks included), returns bytes consumed; samples (mean restored) to s0[, s1] */
int saco_decode_frame(int nch, const uint8_t *in, int len, int32_t *s0, int32_t *s1, int *n_out);

#ifdef __cplusplus
}
#endif
#endif
