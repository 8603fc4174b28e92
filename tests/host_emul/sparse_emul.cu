// TEST INFRASTRUCTURE. Compiles the DEVICE code of sac_b200/csrc/sparse.cu (the map coder's model step, its contexts, the
// range coder of model_dev.cuh, the O(1) rank-mapping formulas) for the HOST by re-declaring __device__ as
// __host__ __device__, runs it on the CPU and compares with the CPU restatement (oracle/liboracle.so, itself pinned to
// the reference): map payload bytes, Map() and Unmap() values. Lets the no-GPU suite check the arithmetic of the CUDA
// sources themselves. Built and run by tests/test_host_emulation.py; prints one line per trial and "ALL EQUAL".
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#undef __device__
#define __device__ __location__(host) __location__(device)
static inline __location__(host) __location__(device) unsigned emul_umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
#define __umulhi emul_umulhi
#include "sparse.cu"
#undef __umulhi
#undef __device__
#define __device__ __location__(device)
using namespace sacb;

extern "C" int saco_map_encode(const int32_t *raw, int n, uint8_t *out, int cap);
extern "C" void saco_remap(const int32_t *raw, int nraw, const int32_t *pred, const int32_t *err, int n, int unmap, int32_t *out);
extern "C" void saco_logdomain_tables(int *fwd, int *inv);
extern "C" void saco_set_modes(int order, int math);

int main()
{
  saco_set_modes(1, 1);
  std::vector<int> fwd(32768), inv(4095);
  saco_logdomain_tables(fwd.data(), inv.data());
  std::vector<int16_t> st(32768), sq(4095);
  for (int i = 0; i < 32768; i++) st[i] = (int16_t)fwd[i];
  for (int i = 0; i < 4095; i++) sq[i] = (int16_t)inv[i];
  const Tables T{st.data(), sq.data(), nullptr, 0};
  srand(5);
  int failures = 0;
  for (int trial = 0; trial < 5; trial++) {
    const int n = trial == 4 ? 300 : 20000;
    std::vector<int32_t> raw(n);
    const int step = trial == 0 ? 16 : trial == 1 ? 3 : trial == 3 ? 7 : 1;
    for (int i = 0; i < n; i++) raw[i] = ((rand() % 60000) - 30000) / step * step + (trial == 3 ? 5 : 0);
    if (trial == 2) for (int i = 0; i < n; i++) raw[i] = (rand() % 4000) - 2000;          // dense around zero
    if (trial == 4) { raw[0] = 32768; raw[1] = -32768; raw[2] = 1; raw[3] = -1; }         // the table's edges
    std::vector<uint8_t> used(kRemapUsedBytes, 0);
    for (int i = 0; i < n; i++) { const int v = raw[i]; if (v != 0 && v >= -kScale && v <= kScale) used[v + kScale] = 1; }   // remap_mark_kernel
    // ---- body of map_encode_kernel ----
    static MapState S;
    map_state_init(S, T);
    std::vector<uint8_t> out(65536);
    Coder rc;
    rc.range = 0xFFFFFFFFu; rc.ffnum = 0; rc.cache = 0; rc.lowc = 0; rc.nbytes = 0; rc.out = out.data();
    MapStep m;
    for (int i = 1; i <= kScale; i++) {
      int bit = used[kScale - i];
      map_ctx_low(S, m, used.data(), i);
      rc_encode<1>(rc, m.predict(S, T), bit, 0);
      m.update(S, bit);
      bit = used[kScale + i];
      map_ctx_high(S, m, used.data(), i);
      rc_encode<1>(rc, m.predict(S, T), bit, 0);
      m.update(S, bit);
    }
    for (int i = 0; i < 5; i++) shift_low<1>(rc, 0);
    std::vector<uint8_t> ob(65536);
    const int nb = saco_map_encode(raw.data(), n, ob.data(), 65536);
    const bool eq = rc.nbytes == nb && !memcmp(out.data(), ob.data(), nb);
    // ---- remap_scan_kernel's tables (serial here), cum_at() of remap_map_kernel, unmap formula of predictor.cu ----
    std::vector<int32_t> cum(kDom), ul(kDom);
    int run = 0;
    for (int i = 0; i < kDom; i++) { const int u = (i == kScale) ? 1 : used[i]; if (u) ul[run] = i - kScale; run += u; cum[i] = run; }
    const int N = 5000;
    std::vector<int32_t> pred(N), err(N), mo(N), uo(N);
    for (int i = 0; i < N; i++) { pred[i] = (rand() % 66000) - 33000; err[i] = (rand() % 801) - 400; if (i % 50 == 0) err[i] *= 100; }
    saco_remap(raw.data(), n, pred.data(), err.data(), N, 0, mo.data());
    int bad = 0, bad2 = 0;
    for (int i = 0; i < N; i++) {
      const int e = err[i], p = pred[i];
      int mm = 0;
      if (e > 0) mm = cum_at(cum.data(), p + e) - cum_at(cum.data(), p);
      else if (e < 0) mm = -(cum_at(cum.data(), p - 1) - cum_at(cum.data(), p + e - 1));
      bad += mm != mo[i];
    }
    // Unmap is only defined for ranks that exist: feed Map's own output back (the reference does not return otherwise)
    std::vector<int32_t> mok(N);
    for (int i = 0; i < N; i++) mok[i] = mo[i];
    saco_remap(raw.data(), n, pred.data(), mok.data(), N, 1, uo.data());
    for (int i = 0; i < N; i++) {
      const int mm = mok[i], p = pred[i];
      int e = 0;
      if (mm != 0) {
        const int nused = cum[65536];
        auto C = [&](int v) { return v < -32768 ? 0 : cum[std::min(v, 32768) + 32768]; };
        int idx = mm > 0 ? C(p) + mm - 1 : C(p - 1) + mm;
        idx = std::min(std::max(idx, 0), nused - 1);
        e = ul[idx] - p;
      }
      bad2 += e != uo[i];
    }
    printf("trial %d: map bytes device-code %lld oracle %d %s; Map mismatches %d, Unmap mismatches %d of %d\n", trial, rc.nbytes, nb,
           eq ? "EQUAL" : "DIFF", bad, bad2, N);
    failures += !eq + bad + bad2;
  }
  if (!failures) printf("ALL EQUAL\n");
  return failures ? 1 : 0;
}
