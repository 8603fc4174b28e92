// sparse.h -- sparse-PCM side of a frame: which sample values occur (Remap), residuals as rank distances among the
// used values (Map / Unmap), and the context-mixing coder that transmits the used-value map (MapEncoder).
// Ref: src/libsac/map.h:10-49, map.cpp:3-202, libsac.cpp:157-160 (Unmap on decode), 214-278 (mapped record, choice).
#ifndef SAC_B200_SPARSE_H
#define SAC_B200_SPARSE_H
#include "bitplane.h"
#include <cstdint>
#include <cuda_runtime.h>

namespace sacb {

constexpr int kRemapScale = 1 << 15;                 // Remap::scale (map.cpp:104)
constexpr int kRemapDom = 2 * kRemapScale + 1;       // values -32768..32768 live at index v + 32768
constexpr size_t kRemapUsedBytes = 65664;            // kRemapDom rounded up to 128
constexpr size_t kMapBytesMax = 20480;               // 131072 binary decisions cannot code to more than 16.4 KB + flush

// what the encode side leaves for the host, per channel
struct SparseOut {
  long long nbytes;      // mapped payload length (written by the bitplane job that continues the map coder)
  long long l1[2];       // sum |e|, sum |Map(e)|  (CostL1 numerators, cost.h:15-27)
  int maxbpn;            // of the rank-mapped residuals
  int go;                // 1: ratio > 1.05, a mapped payload was produced
};

struct SparseJob {
  const int32_t *s;      // mean-free samples of the frame (window plane)
  const int32_t *e;      // encode: residuals of the final pass
  int32_t *em;           // encode: rank-mapped residuals (signed) out
  int n;
  int mean;              // raw value = s + mean (the map is taken on raw samples, libsac.cpp:448-451)
  uint8_t *used;         // [kRemapUsedBytes] used flags by value index; index 32768 (value 0) stays 0 for the map coder
  int32_t *cum;          // [kRemapDom] number of used values <= v (value 0 counts as used, map.cpp:159-166)
  int32_t *ulist;        // [kRemapDom] the used values in ascending order
  uint8_t *bytes;        // encode: payload buffer (map bytes first); decode: the channel's payload
  long long in_len;      // decode: payload length
  RcInit *rc;            // coder state handed to the bitplane job
  SparseOut *res;        // encode
};
struct SparseJobs { SparseJob j[2]; int n; };

// encode: used map of the raw samples, cumulative counts, rank-mapped residuals + L1 sums, then -- only where
// sum|e| / sum|Map(e)| > 1.05 -- the coded map into bytes[] and the coder state into *rc (rc->go tells the bitplane job).
// The caller zeroes used[], *rc and *res beforehand (same stream).
cudaError_t launch_sparse_encode(const BitplaneTables &bt, const SparseJobs &jobs, cudaStream_t stream);
// decode: the map from the head of the payload (MapEncoder::Decode), cumulative counts and sorted list for Unmap,
// coder state into *rc. The caller zeroes used[] beforehand.
cudaError_t launch_sparse_decode(const BitplaneTables &bt, const SparseJobs &jobs, cudaStream_t stream);

} // namespace sacb
#endif
