// codec.h -- host side of the frame / file codec: DDS search driver, frame records, WAV and .sac containers, MD5.
#ifndef SAC_B200_CODEC_H
#define SAC_B200_CODEC_H
#include <cstdint>
#include <limits>
#include <random>
#include <string>
#include <utility>
#include <vector>

namespace sacb {

// ---- searches as resumable state machines so that several of them (frames) can share one GPU batch per generation:
// propose() hands out the candidates of the next generation (first call: the start vector alone), consume() takes
// their costs in the same order.
class Search {
public:
  virtual ~Search() {}
  virtual bool done() const = 0;
  virtual void propose(std::vector<std::vector<double>> &cands) = 0;
  virtual void consume(const double *costs) = 0;
  virtual const std::vector<double> &best_x() const = 0;
  virtual double best_cost() const = 0;
  virtual double sigma() const = 0;
  virtual int nfunc() const = 0;
  // a candidate can only change the search if its cost is below this (+inf while no cost is known)
  virtual double accept_below() const { return std::numeric_limits<double>::infinity(); }
};

// DDS (OptDDS::run_single / run_mt, /root/reference src/opt/dds.cpp:33-106)
class DdsSearch : public Search {
public:
  // spec > 1 (sequential search only, num_threads <= 0): speculative batches, see propose()
  DdsSearch(int D, const double *xmin, const double *xmax, const double *xstart, int nfunc_max, int num_threads, double sigma_init,
            int spec = 1);
  bool done() const override { return started_ && nfunc_ >= nfunc_max_; }
  double accept_below() const override { return started_ ? fb_ : std::numeric_limits<double>::infinity(); }
  long long evaluated() const { return evaluated_; }   // candidates handed out so far (>= nfunc() when speculating)
  // every step of the sequential search as (candidate, cost), in order -- recorded when trace(true) was called (probes)
  void trace(bool on) { trace_on_ = on; }
  const std::vector<std::pair<double, std::vector<double>>> &traced() const { return trace_; }
  void propose(std::vector<std::vector<double>> &cands) override;
  void consume(const double *costs) override;
  const std::vector<double> &best_x() const override { return xb_; }
  double best_cost() const override { return fb_; }
  double sigma() const override { return sigma_; }
  int nfunc() const override { return nfunc_; }

private:
  std::vector<double> candidate(int nfunc);
  double reflect(double xnew, double lo, double hi) const;
  std::mt19937 eng_{0};                       // Opt::rand, seed 0 (opt.cpp:5)
  int D_, nfunc_max_, num_threads_;
  std::vector<double> xmin_, xmax_, xb_;
  double fb_ = 0.0, sigma_;
  int nfunc_ = 0;
  bool started_ = false;
  int nsucc_ = 0, nfail_ = 0;                 // SSC0(3,50)
  double p_succ_ = 0.05;                      // SSC1(0.05,0.10,0.05)
  std::vector<std::vector<double>> pending_;
  // speculative sequential search: state of run_single as it would be BEFORE the outcome of pending_[i] is known, under
  // the assumption that every earlier candidate of the batch failed, and the generator state after drawing pending_[i]
  struct SpecPoint { double sigma; int nsucc, nfail, nfunc; std::mt19937 rng_after; };
  std::vector<SpecPoint> spec_pts_;
  int spec_ = 1;
  double q_est_ = 0.125;                      // running estimate of the per-step success rate (sizes the batches); typical: 0.07 .. 0.15
  long long evaluated_ = 0;
  bool trace_on_ = false;
  std::vector<std::pair<double, std::vector<double>>> trace_;
  void ssc0_step(bool ok);
};

// Differential evolution, the reference's --opt-cfg=de (OptDE, src/opt/de.cpp, de.h: JADE-style current-to-pbest/1/bin,
// NP = 30, adaptive CR / F). A generation is NP trial vectors: it maps onto one population evaluation as is.
class DeSearch : public Search {
public:
  DeSearch(int D, const double *xmin, const double *xmax, const double *xstart, int nfunc_max, double sigma_init);
  bool done() const override { return phase_ == 2 && nfunc_ >= nfunc_max_; }
  void propose(std::vector<std::vector<double>> &cands) override;
  void consume(const double *costs) override;
  const std::vector<double> &best_x() const override { return xb_.second; }
  double best_cost() const override { return xb_.first; }
  double sigma() const override { return mF_; }
  int nfunc() const override { return nfunc_; }

private:
  typedef std::pair<double, std::vector<double>> Point;    // Opt::ppoint (opt.h:14)
  double reflect(double xnew, double lo, double hi) const;
  std::vector<int> select_k_unique_except(int n, int ie, int k);
  std::mt19937 eng_{0};                       // Opt::rand, seed 0 (opt.cpp:5)
  int D_, nfunc_max_;
  std::vector<double> xmin_, xmax_;
  double sigma_init_;
  // DECfg defaults (de.h:18-31)
  static constexpr int kNP = 30;
  static constexpr int kNpBest = 2;           // clamp(round(0.1 * 30) - 1, 0, 29)
  double mCR_ = 0.5, mF_ = 0.5;
  std::vector<Point> pop_, gen_;
  std::vector<std::pair<double, double>> gen_mut_;
  Point xb_;
  int nfunc_ = 0;
  int phase_ = 0;                             // 0: start vector, 1: initial population, 2: generations
};

// (1+1)-CMA-ES, the reference's --opt-cfg=cma (OptCMA, src/opt/cma.cpp, cma.h; Cholesky of the covariance with 0.1
// regularisation per step, SSC1 step-size control). Strictly one evaluation per step.
class CmaSearch : public Search {
public:
  CmaSearch(int D, const double *xmin, const double *xmax, const double *xstart, int nfunc_max, double sigma_init);
  bool done() const override { return started_ && nfunc_ >= nfunc_max_; }
  void propose(std::vector<std::vector<double>> &cands) override;
  void consume(const double *costs) override;
  const std::vector<double> &best_x() const override { return xb_; }
  double best_cost() const override { return fb_; }
  double sigma() const override { return sigma_; }
  int nfunc() const override { return nfunc_; }

private:
  double reflect(double xnew, double lo, double hi) const;
  std::mt19937 eng_{0};
  int D_, nfunc_max_;
  std::vector<double> xmin_, xmax_, xb_, pc_, az_, xgen_;
  std::vector<std::vector<double>> mcov_, G_;
  double fb_ = 0.0, sigma_;
  double cc_, ccov_, p_target_, cp_, dinv_, p_succ_;     // CMAParams (cma.h:21-37), SSC1 state
  int nfunc_ = 0;
  bool started_ = false;
};

// ---- MD5 (RFC 1321) over the raw PCM bytes, as the reference stores in the .sac header ----
struct Md5 {
  uint32_t a, b, c, d;
  uint64_t len;
  uint8_t buf[64];
  int fill;
  Md5();
  void update(const uint8_t *p, size_t n);
  void final(uint8_t out[16]);
};

// ---- WAV (src/file/wav.cpp) ----
struct WavChunk { uint32_t id, csize; std::vector<uint8_t> data; };
struct WavInfo {
  int nch = 0, samplerate = 0, bits = 0, blockalign = 0;
  uint32_t numsamples = 0;
  std::vector<WavChunk> chunks;               // every chunk in file order ('data' without payload)
  size_t data_pos = 0;
  uint32_t metadatasize() const;
};
// parses the RIFF structure of a complete file image; returns 0 or a negative SAC_E_* code
int wav_parse(const uint8_t *file, size_t len, WavInfo &wi);
void wav_unpack(const WavInfo &wi, const uint8_t *pcm, int first, int count, std::vector<std::vector<int32_t>> &planes);
void wav_pack(const WavInfo &wi, const std::vector<std::vector<int32_t>> &planes, int count, std::vector<uint8_t> &out);
void pack_metadata(const WavInfo &wi, std::vector<uint8_t> &out);
int unpack_metadata(const uint8_t *p, size_t n, WavInfo &wi);

} // namespace sacb
#endif
