// model_dev.cuh -- device-side pieces of the reference's probability model that more than one coder uses (the
// bitplane coder in bitplane.cu, the sparse-PCM map coder in sparse.cu): integer helpers, LogDomain table access,
// LinearCounter16, the binary range coder. Ref: src/model/counter.h:31-37, domain.h:7-61, range.cpp:54-92.
#ifndef SAC_B200_MODEL_DEV_CUH
#define SAC_B200_MODEL_DEV_CUH
#include <cstdint>
#include <cuda_runtime.h>

namespace sacb {
namespace {

constexpr int PBITS = 15, PSCALE = 1 << 15, PSCALEm = PSCALE - 1;

__device__ __forceinline__ int idiv_s(int val, int s) { return val < 0 ? -(((-val) + (1 << (s - 1))) >> s) : (val + (1 << (s - 1))) >> s; }
__device__ __forceinline__ int idiv_s64(long long val, int s)
{
  return (int)(val < 0 ? -(((-val) + (1LL << (s - 1))) >> s) : (val + (1LL << (s - 1))) >> s);
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

struct Tables { const int16_t *stretch; const int16_t *squash; const uint16_t *laplace; int lap_bits; };

// stretch / squash point into the CTA's shared-memory copies (stage_tables): they sit on the serial chain of every decision
__device__ __forceinline__ int stretch(const Tables &T, int p) { return T.stretch[p]; }
__device__ __forceinline__ int squash(const Tables &T, int x)
{
  if (x < -2047) return 1;
  if (x > 2047) return PSCALEm;
  return T.squash[x + 2047];
}

// LinearCounter16::update(bit, L) (counter.h:31-37)
__device__ __forceinline__ int counter16_upd(int p1, int bit, int L)
{
  const int err = (bit << PBITS) - p1;
  return clampi(p1 + idiv_s(L * err, PBITS), 1, PSCALEm);
}

// the adaptive chain for one binary decision; all lanes execute it with identical operands
struct Coder {
  uint32_t range, ffnum, cache;
  unsigned long long lowc;
  long long nbytes;
  uint8_t *out;
};

template <int MODE /*0 cost, 1 encode*/>
__device__ __forceinline__ void shift_low(Coder &rc, int lane)
{
  if (MODE == 1) {
    const uint32_t carry = (uint32_t)(rc.lowc >> 32), low = (uint32_t)rc.lowc;
    if (low < 0xFF000000u || carry) {
      if (lane == 0) rc.out[rc.nbytes] = (uint8_t)(rc.cache + carry);
      rc.nbytes++;
      for (; rc.ffnum != 0; rc.ffnum--) { if (lane == 0) rc.out[rc.nbytes] = (uint8_t)(carry - 1); rc.nbytes++; }
      rc.cache = low >> 24;
    } else rc.ffnum++;
    rc.lowc = (unsigned long long)(uint32_t)(low << 8);
  } else {
    rc.nbytes++;                                                   // every ShiftLow eventually emits exactly one byte
  }
}
template <int MODE>
__device__ __forceinline__ void rc_encode(Coder &rc, int p1, int bit, int lane)
{
  const uint32_t rnew = __umulhi(rc.range, (uint32_t)(PSCALE - p1) << (32 - PBITS));   // range.h:23
  if (bit) { rc.range -= rnew; if (MODE == 1) rc.lowc += rnew; } else rc.range = rnew;
  while (rc.range < 0x01000000u) { rc.range <<= 8; shift_low<MODE>(rc, lane); }
}

} // namespace
} // namespace sacb
#endif
