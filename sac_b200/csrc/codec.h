// codec.h -- host side of the frame / file codec: DDS search driver, frame records, WAV and .sac containers, MD5.
#ifndef SAC_B200_CODEC_H
#define SAC_B200_CODEC_H
#include <cstdint>
#include <random>
#include <string>
#include <vector>

namespace sacb {

// ---- DDS as a resumable state machine (OptDDS::run_single / run_mt, /root/reference src/opt/dds.cpp:33-106) so that
// several searches (frames) can share one GPU batch per generation.
class DdsSearch {
public:
  DdsSearch(int D, const double *xmin, const double *xmax, const double *xstart, int nfunc_max, int num_threads, double sigma_init);
  bool done() const { return started_ && nfunc_ >= nfunc_max_; }
  // candidates of the next generation (first call: the start vector alone)
  void propose(std::vector<std::vector<double>> &cands);
  void consume(const double *costs);
  const std::vector<double> &best_x() const { return xb_; }
  double best_cost() const { return fb_; }
  double sigma() const { return sigma_; }
  int nfunc() const { return nfunc_; }

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
