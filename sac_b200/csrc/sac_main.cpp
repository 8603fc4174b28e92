// sac_main.cpp -- the `sac` command line of the reference (/root/reference src/main.cpp, src/cmdline.cpp), same flags,
// driving libsac_b200 through its C ABI. Options the B200 path adds: --gpu=N, --frame-parallel.
#include "sac_b200.h"
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

static const char kHelp[] =
    "usage: sac [--options] input output\n\n"
    "  --encode            encode input.wav to output.sac (def)\n"
    "    --normal|high|veryhigh|extrahigh compression (def=normal)\n"
    "    --best            you asked for it\n\n"
    "  --decode            decode input.sac to output.wav\n"
    "  --list              list info about input.sac\n"
    "  --listfull          verbose info about input\n"
    "  --verbose           verbose output\n\n"
    "  supported types: 1-16 bit, mono/stereo pcm\n"
    "  advanced options    (automatically set)\n"
    "   --optimize=#       frame-based optimization\n"
    "     no|s,n,c,k       s=[0,1.0],n=[0,10000]\n"
    "                      c=[l1,rms,glb,ent,bpn] k=[1,32]\n"
    "   --opt-cfg=#        configure optimization method\n"
    "     dds|de|cma,nt,s  nt=generation size (GPU batch; dds default: an eighth of the evaluation\n"
    "                      budget, at most 128; 0 = the reference's sequential search), s=search radius (def=0.2)\n"
    "   --opt-reset        reset opt params at frame boundaries\n"
    "   --mt-mode=n        accepted, ignored (parallelism is the GPU's)\n"
    "   --zero-mean        zero-mean input\n"
    "   --adapt-block      adaptive frame splitting (default), --adapt-block=no disables it\n"
    "   --framelen=n       def=20 seconds\n"
    "   --sparse-pcm       enable pcm modelling (default), --sparse-pcm=no disables it\n"
    "  B200 options\n"
    "   --gpu=n            CUDA device (def=0)\n"
    "   --frame-parallel[=1|2]  search all frames of the file concurrently (implies --opt-reset semantics):\n"
    "                      1 = shared launches per generation, 2 = one stream and host thread per frame (default 2)\n";

static std::string upper(std::string s) { for (auto &c : s) c = (char)std::toupper((unsigned char)c); return s; }
static std::vector<std::string> split(const std::string &s, char d)
{
  std::vector<std::string> out; std::string cur;
  for (char c : s) { if (c == d) { if (!cur.empty()) out.push_back(cur); cur.clear(); } else cur += c; }
  if (!cur.empty()) out.push_back(cur);
  return out;
}
static uint32_t rd32(const uint8_t *b) { return b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24); }
static uint16_t rd16(const uint8_t *b) { return (uint16_t)(b[0] | (b[1] << 8)); }
static std::string time_str(long long numsamples, long long sr)
{
  int h = 0, m = 0, s = 0, ms = 0;
  if (numsamples > 0 && sr > 0) {
    while (numsamples >= 3600 * sr) { ++h; numsamples -= 3600 * sr; }
    while (numsamples >= 60 * sr) { ++m; numsamples -= 60 * sr; }
    while (numsamples >= sr) { ++s; numsamples -= sr; }
    ms = (int)std::round((numsamples * 1000.) / sr);
  }
  char b[64]; std::snprintf(b, sizeof b, "%02d:%02d:%02d.%d", h, m, s, ms);
  return b;
}
static void print_md5(const uint8_t d[16]) { for (int i = 0; i < 16; i++) std::printf("%x", (int)d[i]); }   // cmdline.cpp:313 (no zero padding)

static int list_file(const std::string &path, bool full, int mt_mode)
{
  std::ifstream f(path, std::ios::binary);
  if (!f) { std::cout << "could not open\n"; return 1; }
  std::vector<uint8_t> b((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  std::cout << "ok (" << b.size() << " Bytes)\n";
  if (b.size() < 38 || std::memcmp(b.data(), "SAC2", 4)) { std::cout << "warning: input is not a valid .sac file\n"; return 1; }
  const int nch = rd16(&b[4]), sr = (int)rd32(&b[6]), bits = rd16(&b[10]);
  const uint32_t ns = rd32(&b[12]);
  const int fl = b[16];
  const uint32_t md = rd32(&b[18]);
  const double bps = (b.size() * 8.0) / ((double)ns * nch);
  std::printf("  WAVE  Codec: PCM (%d kbps)\n  %dHz %d Bit  %s\n  %u Samples [%s]\n", (int)std::round((sr * nch * bps) / 1000), sr, bits,
              nch == 1 ? "Mono" : "Stereo", ns, time_str(ns, sr).c_str());
  std::printf("  Profile: mt%d %ds\n  Ratio:   %.3f bps\n\n  Audio MD5: ", mt_mode, fl, bps);   // cmdline.cpp:307-311
  size_t pos = 22 + (size_t)md;
  if (pos + 16 > b.size()) { std::printf("\nwarning: input is not a valid .sac file\n"); return 1; }   // metadata size beyond the file
  print_md5(&b[pos]);
  std::printf("\n");
  pos += 16;
  if (full) {                                                         // Codec::ScanFrames (libsac.cpp:659-693)
    int frame = 1, coef = 0, blk = 0;
    while (pos + 4 <= b.size()) {
      std::printf("Frame %d: %u samples \n", frame, rd32(&b[pos]));
      pos += 4 + 58 * 4; coef += 58 * 4;
      for (int ch = 0; ch < nch && pos + 18 <= b.size(); ch++) {
        const uint32_t bs = rd32(&b[pos]);
        const uint16_t flag = rd16(&b[pos + 16]);
        std::printf("  Channel %d: %u bytes\n    Bpn: %d, sparse_pcm: %d\n    mean: %d, min: %d, max: %d\n", ch, bs, flag & 0xff, flag >> 9,
                    (int32_t)rd32(&b[pos + 4]), (int32_t)rd32(&b[pos + 8]), (int32_t)rd32(&b[pos + 12]));
        pos += 18 + bs; blk += 18;
      }
      frame++;
    }
    std::printf("Frames   %d\nHdr_size %d (coefs %d,block %d)\n", frame - 1, coef + blk, coef, blk);
  }
  return 0;
}

int main(int argc, const char *argv[])
{
  std::cout << "Sac-B200 v0.1 - Lossless Audio Coder, B200-native encode path (model of Sac v0.7.25 by Sebastian Lehmann)\n\n";
  if (argc < 2) { std::cout << kHelp; return 1; }
  enum { ENCODE, DECODE, LIST, LISTFULL } mode = ENCODE;
  sac_cfg cfg;
  sac_cfg_default(&cfg);
  std::string in, out;
  bool first = true;
  int gpu = 0, gpus = 1;
  int mt_mode = 2;                                                    // tsac_cfg default (libsac.h:19-44); listings echo it
  bool gen_given = false;                                             // --opt-cfg=...,N seen
  for (int k = 1; k < argc; k++) {
    const std::string param = argv[k];
    const std::string up = upper(param);
    std::string key = up, val;
    const size_t eq = up.find('=');
    if (eq != std::string::npos) { key = up.substr(0, eq); val = up.substr(eq + 1); }
    if (param.size() > 1 && param[0] == '-' && param[1] == '-') {
      if (key == "--ENCODE") mode = ENCODE;
      else if (key == "--DECODE") mode = DECODE;
      else if (key == "--LIST") mode = LIST;
      else if (key == "--LISTFULL") mode = LISTFULL;
      else if (key == "--VERBOSE") cfg.verbose = val.size() ? std::max(0, std::atoi(val.c_str())) : 1;
      else if (key == "--NORMAL" || key == "--HIGH" || key == "--VERYHIGH" || key == "--EXTRAHIGH" || key == "--BEST" || key == "--INSANE")
        sac_cfg_preset(&cfg, key.c_str() + 2);
      else if (key == "--OPTIMIZE") {
        if (val == "NO" || val == "0") cfg.optimize = 0;
        else {
          auto vs = split(val, ',');
          if (vs.size() >= 2) {
            cfg.fraction = std::clamp(std::atof(vs[0].c_str()), 0., 1.);
            cfg.maxnfunc = std::clamp(std::atoi(vs[1].c_str()), 0, 50000);
            if (vs.size() >= 3) {
              if (vs[2] == "L1") cfg.cost_kind = SAC_COST_L1;
              else if (vs[2] == "RMS") cfg.cost_kind = SAC_COST_RMS;
              else if (vs[2] == "GLB") cfg.cost_kind = SAC_COST_GOLOMB;
              else if (vs[2] == "ENT") cfg.cost_kind = SAC_COST_ENTROPY;
              else if (vs[2] == "BPN") cfg.cost_kind = SAC_COST_BITPLANE;
              else std::cerr << "warning: unknown cost function '" << vs[2] << "'\n";
            }
            if (vs.size() >= 4) cfg.optk = std::clamp(std::atoi(vs[3].c_str()), 1, 32);
            cfg.optimize = (cfg.fraction > 0. && cfg.maxnfunc > 0) ? 1 : 0;
          } else std::cerr << "unknown option: " << val << '\n';
        }
      } else if (key == "--FRAMELEN") { if (val.size()) cfg.max_framelen = std::max(0, std::atoi(val.c_str())); }
      else if (key == "--MT-MODE") { if (val.size()) mt_mode = std::max(0, std::min(2, std::atoi(val.c_str()))); }   // cmdline.cpp:185-187 (only echoed)
      else if (key == "--SPARSE-PCM") cfg.sparse_pcm = !(val == "NO" || val == "0");
      else if (key == "--STEREO-MS") {}
      else if (key == "--OPT-RESET") cfg.reset = 1;
      else if (key == "--OPT-CFG") {
        auto vs = split(val, ',');
        if (vs.size() >= 1) {                                            // cmdline.cpp:198-204
          if (vs[0] == "DDS") cfg.search = SAC_SEARCH_DDS;
          else if (vs[0] == "DE") cfg.search = SAC_SEARCH_DE;
          else if (vs[0] == "CMA") cfg.search = SAC_SEARCH_CMA;
          else std::cerr << "  warning: invalid opt='" << vs[0] << "'\n";
        }
        if (vs.size() >= 2) { cfg.num_threads = std::clamp(std::atoi(vs[1].c_str()), 0, 4096); gen_given = true; }
        if (vs.size() >= 3) cfg.sigma = std::clamp(std::atof(vs[2].c_str()), 0., 1.);
      } else if (key == "--ADAPT-BLOCK") cfg.adapt_block = !(val == "NO" || val == "0");
      else if (key == "--ZERO-MEAN") cfg.zero_mean = !(val == "NO" || val == "0");
      else if (key == "--GPU") gpu = std::atoi(val.c_str());
      else if (key == "--GPUS") gpus = (val == "ALL" || val.empty()) ? sac_device_count() : std::clamp(std::atoi(val.c_str()), 1, 64);   // frames of the file across the box's GPUs (implies --opt-reset)
      else if (key == "--FRAME-PARALLEL") cfg.frame_parallel = val.empty() ? 2 : std::max(0, std::min(2, std::atoi(val.c_str())));
      else if (key == "--SPEC") cfg.spec = std::clamp(std::atoi(val.c_str()), 1, 256);          // candidates per speculative batch
      else if (key == "--INFLIGHT") cfg.inflight = std::clamp(std::atoi(val.c_str()), 1, 64);   // frames in flight (--frame-parallel=2)
      else if (key == "--GRADE") cfg.grade = std::clamp(std::atoi(val.c_str()), 0, 1);          // 1: search-grade kernels for the search
      else std::cerr << "warning: unknown option '" << param << "'\n";
    } else {
      if (first) { in = param; first = false; } else out = param;
    }
  }
  if (mode == LIST || mode == LISTFULL) {
    std::cout << "Open: '" << in << "': ";
    const int rc = list_file(in, mode == LISTFULL, mt_mode);
    std::printf("\n  Time:    [00:00:00]\n");                          // cmdline.cpp:355-356
    return rc;
  }
  // The reference's default search is sequential (toptim_cfg::num_threads = 0, libsac.h:26): one candidate per step. On a GPU that
  // is a chain of ~100 dependent launches per frame even in speculative batches (sac_cfg::spec), each as long as its slowest
  // candidate: minutes for seconds of audio. Unless the user says otherwise the CLI therefore runs the reference's own population
  // variant (OptDDS::run_mt, dds.cpp:63-106) in about 8 generations (an eighth of the evaluation budget per generation, at most
  // 128: --best -> 125, --high -> 13). --opt-cfg=dds,0 selects the reference's sequential search, executed in speculative batches
  // (same accepted sequence and result as one candidate at a time).
  if (mode == ENCODE && cfg.optimize && cfg.search == SAC_SEARCH_DDS && !gen_given) cfg.num_threads = std::clamp((cfg.maxnfunc + 7) / 8, 1, 128);
  // console output mirrors CmdLine::Process (cmdline.cpp:245-358): Open / PrintWav / Create / PrintMode / MD5 / ratio line
  const auto t_all = std::chrono::steady_clock::now();
  std::vector<uint8_t> img;
  std::cout << "Open: '" << in << "': ";
  {
    std::ifstream f(in, std::ios::binary);
    if (!f) { std::cout << "could not open\n"; return 1; }
    img.assign((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  }
  std::cout << "ok (" << img.size() << " Bytes)\n";
  sac_file_stats st;
  std::memset(&st, 0, sizeof(st));
  int rc;
  sac_engine *eng = nullptr;
  if (mode == ENCODE) {
    // host half first (WAV parse, header, MD5, frame plan: no GPU involved), so that a bad input is reported as the reference does
    if (sac_container_plan(&cfg, img.data(), (long long)img.size(), nullptr, 0, nullptr, nullptr, 0, &st)) {
      const std::string err = sac_last_error();
      if (err.find("unsupported input format") != std::string::npos) {
        std::printf("  WAVE  Codec: PCM\n");
        std::cerr << "unsupported input format\nmust be 1-16 bit, mono/stereo, pcm\n";                 // cmdline.cpp:253-262
      } else std::cout << "warning: input is not a valid .wav file\n";
      return 1;
    }
    std::printf("  WAVE  Codec: PCM (%d kbps)\n  %dHz %d Bit  %s\n  %d Samples [%s]\n", (st.samplerate * st.nch * st.bits) / 1000, st.samplerate,
                st.bits, st.nch == 1 ? "Mono" : "Stereo", st.numsamples, time_str(st.numsamples, st.samplerate).c_str());
    eng = sac_engine_create(gpu);
    if (!eng) { std::cerr << "error: " << sac_last_error() << "\n"; return 1; }
    std::cout << "Create: '" << out << "': ";
    { std::ofstream probe(out, std::ios::binary); if (!probe) { std::cout << "could not create\n"; sac_engine_destroy(eng); return 1; } }
    std::cout << "ok\n";
    std::printf("  Profile: mt%d %ds%s%s%s\n", mt_mode, cfg.max_framelen, cfg.adapt_block ? " ab" : "", cfg.zero_mean ? " zero-mean" : "",
                cfg.sparse_pcm ? " sparse-pcm" : "");
    if (cfg.optimize) {
      const char *cs[] = {"L1", "rms", "ent", "glb", "bpn"};
      std::printf("  Optimize: %s %.1f%%,n=%d,%s,k=%d\n", cfg.search == SAC_SEARCH_DE ? "DE" : (cfg.search == SAC_SEARCH_CMA ? "" : "DDS"),
                  cfg.fraction * 100.0, cfg.maxnfunc, cs[cfg.cost_kind], cfg.optk);
    }
    if (cfg.optimize)                                                                                   // not in the reference: how the GPU is fed
      std::printf("  B200: %s %d, frame-parallel %d (in flight %d), search grade %d, gpu %d\n", cfg.num_threads > 0 || cfg.search != SAC_SEARCH_DDS ? "generation" : "speculative batch",
                  cfg.search == SAC_SEARCH_DE ? 30 : (cfg.search == SAC_SEARCH_CMA ? 1 : (cfg.num_threads > 0 ? cfg.num_threads : cfg.spec)), cfg.frame_parallel, cfg.inflight, cfg.grade, gpu);
    std::printf("\n");
    if (gpus > 1) {
      std::vector<sac_engine *> engs{eng};
      const int ndev = sac_device_count();
      for (int d = 0; d < ndev && (int)engs.size() < gpus; d++) {
        if (d == gpu) continue;
        sac_engine *x = sac_engine_create(d);
        if (!x) { std::cerr << "error: " << sac_last_error() << "\n"; return 1; }
        engs.push_back(x);
      }
      std::printf("  B200: %d GPUs, frames dealt round-robin (--opt-reset semantics)\n\n", (int)engs.size());
      rc = sac_encode_file_multi(engs.data(), (int)engs.size(), &cfg, in.c_str(), out.c_str(), &st);
      for (size_t i = 1; i < engs.size(); i++) sac_engine_destroy(engs[i]);
    } else rc = sac_encode_file(eng, &cfg, in.c_str(), out.c_str(), &st);
    if (rc) { std::cerr << "error: " << sac_last_error() << "\n"; sac_engine_destroy(eng); return 1; }
    std::printf("  %d/%d: 100.0%%\n", st.numsamples, st.numsamples);
    std::printf("  MD5:     "); print_md5(st.md5); std::printf("\n");
    const double r = st.out_bytes * 100.0 / st.in_bytes, bps = (st.out_bytes * 8.) / ((double)st.numsamples * st.nch);
    const double xr = st.seconds > 0 ? (st.numsamples / (double)st.samplerate) / st.seconds : 0.0;
    std::printf("\n  %lld->%lld=%.1f%% (%.3f bps)  %.3fx\n", st.in_bytes, st.out_bytes, r, bps, xr);
  } else {
    if (img.size() < 38 || std::memcmp(img.data(), "SAC2", 4)) { std::cout << "warning: input is not a valid .sac file\n"; return 1; }
    {
      const int nch = rd16(&img[4]), sr = (int)rd32(&img[6]), bits = rd16(&img[10]), fl = img[16];
      const uint32_t ns = rd32(&img[12]), md = rd32(&img[18]);
      const double bps = (img.size() * 8.0) / ((double)ns * nch);
      std::printf("  WAVE  Codec: PCM (%d kbps)\n  %dHz %d Bit  %s\n  %u Samples [%s]\n", (int)std::round((sr * nch * bps) / 1000), sr, bits,
                  nch == 1 ? "Mono" : "Stereo", ns, time_str(ns, sr).c_str());
      std::printf("  Profile: mt%d %ds\n  Ratio:   %.3f bps\n\n  Audio MD5: ", mt_mode, fl, bps);            // cmdline.cpp:307-311
      if (22 + (size_t)md + 16 <= img.size()) print_md5(&img[22 + md]);
      std::printf("\n");
    }
    eng = sac_engine_create(gpu);
    if (!eng) { std::cerr << "error: " << sac_last_error() << "\n"; return 1; }
    std::cout << "Create: '" << out << "': ";
    rc = sac_decode_file(eng, in.c_str(), out.c_str(), &st);
    if (rc == SAC_E_MD5) {                                                                             // nothing was written: the audio is not the encoder's input
      std::cout << "ok\n";
      std::printf("\n  Audio MD5: Error ("); print_md5(st.md5); std::printf(")\n");
      std::cerr << "error: " << sac_last_error() << "\n";
      std::remove(out.c_str());
      sac_engine_destroy(eng);
      return 1;
    }
    if (rc) { std::cout << "could not create\n"; std::cerr << "error: " << sac_last_error() << "\n"; sac_engine_destroy(eng); return 1; }
    std::cout << "ok\n";
    std::printf("  %d/%d: 100.0%%\n", st.numsamples, st.numsamples);
    const double xr = st.seconds > 0 ? (st.numsamples / (double)st.samplerate) / st.seconds : 0.0;
    std::printf("\n  Speed %.3fx\n  Audio MD5: ", xr);
    if (st.md5_ok) std::printf("ok\n"); else { std::printf("Error ("); print_md5(st.md5); std::printf(")\n"); }
  }
  st.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_all).count();
  { const long long t = (long long)std::llround(st.seconds); std::printf("\n  Time:    [%02d:%02d:%02d]\n", (int)(t / 3600), (int)(t / 60) % 60, (int)(t % 60)); }
  sac_engine_destroy(eng);
  return rc ? 1 : 0;
}
