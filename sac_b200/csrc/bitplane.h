// bitplane.h -- job descriptors and launchers of the bitplane coder kernels (bitplane.cu)
#ifndef SAC_B200_BITPLANE_H
#define SAC_B200_BITPLANE_H
#include <stdint.h>
#include <vector>
#include <cuda_runtime.h>

namespace sacb {

// Range-coder state handed from the sparse-PCM map coder (sparse.cu) to the bitplane coder that continues the same
// payload (EncodeMonoFrame_Mapped / DecodeMonoFrame, libsac.cpp:214-228, 280-298).
struct RcInit {
  unsigned long long lowc;   // encoder: low incl. carry
  long long nbytes;          // encoder: bytes already written to `out`; decoder: bytes already consumed from `in`
  uint32_t range, ffnum, cache, code;
  int go;                    // encoder: 0 = the cost ratio did not call for a mapped record; the bitplane job returns at once
  int pad;
};

struct BpJob {
  int32_t *buf;          // encode/cost: residuals (signed if signed_input, mapped in place) ; decode: output (signed)
  int n;
  int signed_input;      // 1: apply S2U first (MathUtils::S2U)
  int maxbpn;            // encode: -1 = derive from max ; decode: from the block header
  uint32_t *csig0;       // 65536 counters of the significance-pattern context (HBM scratch, per job)
  uint8_t *out;          // encode mode 1: payload bytes
  long long *nbytes;     // payload length (cost: the byte count)
  int *maxbpn_out;
  const uint8_t *in;     // decode: payload
  long long in_len;
  uint8_t *msb;          // decode: n bytes scratch
  const RcInit *rc_init; // null: a fresh coder; else continue from this state (written by an earlier kernel of the stream)
};

struct BitplaneTables {
  int16_t *d_stretch = nullptr, *d_squash = nullptr;
  uint16_t *d_laplace = nullptr;
  int lap_bits = 0;
  std::vector<int16_t> h_stretch, h_squash;
  cudaError_t init(cudaStream_t stream);
  void destroy();
};

// mode 0: byte count only (CostBitplane), 1: emit payload, 2: decode
cudaError_t launch_bitplane(const BitplaneTables &bt, const BpJob *d_jobs, int njobs, int mode, cudaStream_t stream);
size_t bitplane_state_bytes();
cudaError_t bitplane_init_attributes();
void compute_logdomain_tables(int16_t *stretch /*[32768]*/, int16_t *squash /*[4095]*/);

} // namespace sacb
#endif
