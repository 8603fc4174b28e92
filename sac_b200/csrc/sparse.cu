// sparse.cu -- sparse-PCM kernels (see sparse.h). Integer arithmetic throughout; results are the reference's bit for bit.
//
//   remap_mark_kernel   Remap::Analyse (map.cpp:126-157): one flag per occurring raw value, |v| <= 32768
//   remap_scan_kernel   cumulative counts + sorted list of the used values. Map/Unmap walk the value axis one step at a
//                       time in the reference (map.cpp:175-202); with the cumulative counts both are O(1):
//                         Map(pred, e>0)   = C(pred+e) - C(pred)          Map(pred, e<0)   = -(C(pred-1) - C(pred+e-1))
//                         Unmap(pred, m>0) = U[C(pred)+m-1] - pred        Unmap(pred, m<0) = U[C(pred-1)+m] - pred
//                       where C(v) = #{used u <= v} (0 counts as used, nothing beyond +-32768 does), U = used values ascending
//   remap_map_kernel    CalcRemapError (libsac.cpp:230-251): rank-mapped residuals and the two L1 sums
//   map_encode_kernel   MapEncoder::Encode (map.cpp:9-89): 2 x 32768 binary decisions through 5 counters -> mixer ->
//                       SSE -> final mixer -> range coder. One serial adaptive chain of 131 072 decisions (about 1 % of a
//                       frame's decisions): one thread, state in shared memory.
//   map_decode_kernel   MapEncoder::Decode (map.cpp:91-102)
#include "sparse.h"
#include "model_dev.cuh"

namespace sacb {

namespace {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kScale = kRemapScale, kDom = kRemapDom;
constexpr int kMapTScale = 2662, kMapXScale = (2 * kMapTScale) / 31;   // SSENL<32>: tscale = myDomain.max, xscale = 2*tscale/(N-1)
constexpr int kRateCnt = 500, kRateSse = 300, kRateMix = 1000, kRateFinal = 500;   // map.h:11-14

__global__ void remap_mark_kernel(SparseJobs J)
{
  const SparseJob &j = J.j[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < j.n; i += gridDim.x * blockDim.x) {
    const int v = j.s[i] + j.mean;
    if (v != 0 && v >= -kScale && v <= kScale) j.used[v + kScale] = 1;
  }
}

__global__ void __launch_bounds__(1024) remap_scan_kernel(SparseJobs J)
{
  const SparseJob &j = J.j[blockIdx.x];
  __shared__ int wsum[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kPer = (kDom + 1023) / 1024;
  const int i0 = min(tid * kPer, kDom), i1 = min(i0 + kPer, kDom);
  int cnt = 0;
  for (int i = i0; i < i1; i++) cnt += (i == kScale) ? 1 : (int)j.used[i];
  int x = cnt;
  for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(kFullMask, x, o); if (lane >= o) x += y; }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane];
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(kFullMask, w, o); if (lane >= o) w += y; }
    wsum[lane] = w;
  }
  __syncthreads();
  int run = x - cnt + (warp ? wsum[warp - 1] : 0);
  for (int i = i0; i < i1; i++) {
    const int u = (i == kScale) ? 1 : (int)j.used[i];
    if (u) j.ulist[run] = i - kScale;
    run += u;
    j.cum[i] = run;
  }
}

__device__ __forceinline__ int cum_at(const int32_t *cum, int v) { return v < -kScale ? 0 : cum[min(v, kScale) + kScale]; }

__global__ void __launch_bounds__(256) remap_map_kernel(SparseJobs J)
{
  const SparseJob &j = J.j[blockIdx.y];
  long long s1 = 0, s2 = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < j.n; i += gridDim.x * blockDim.x) {
    const int e = j.e[i];
    const int pred = (j.s[i] - e) + j.mean;                        // pred[ch][i] = pi + mean (libsac.cpp:107)
    int m = 0;
    if (e > 0) m = cum_at(j.cum, pred + e) - cum_at(j.cum, pred);
    else if (e < 0) m = -(cum_at(j.cum, pred - 1) - cum_at(j.cum, pred + e - 1));
    j.em[i] = m;
    s1 += e < 0 ? -(long long)e : (long long)e;
    s2 += m < 0 ? -(long long)m : (long long)m;
  }
  for (int o = 16; o >= 1; o >>= 1) { s1 += __shfl_xor_sync(kFullMask, s1, o); s2 += __shfl_xor_sync(kFullMask, s2, o); }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(reinterpret_cast<unsigned long long *>(&j.res->l1[0]), (unsigned long long)s1);
    atomicAdd(reinterpret_cast<unsigned long long *>(&j.res->l1[1]), (unsigned long long)s2);
  }
}

// ---- MapEncoder ----------------------------------------------------------------------------------------------
struct MapState {
  int mixl[4][5], mixh[4][5];                                      // NMixLogistic(5) x 4 per half (map.cpp:4)
  int fm[2];                                                       // finalmix(2)
  uint16_t cnt[24], cctx[64];                                      // LinearCounter16, p1 = PSCALE/2; cctx: low [0,16), high [32,48)
  uint16_t sse[2][34];                                             // SSENL<32> sse[0] (the only one addressed, map.cpp:80-99)
  int lb;
};

__device__ void map_state_init(MapState &S, const Tables &T)
{
  for (int i = 0; i < 4; i++) for (int k = 0; k < 5; k++) { S.mixl[i][k] = 0; S.mixh[i][k] = 0; }
  S.fm[0] = S.fm[1] = 0;
  for (int i = 0; i < 24; i++) S.cnt[i] = PSCALE >> 1;
  for (int i = 0; i < 64; i++) S.cctx[i] = PSCALE >> 1;
  for (int i = 0; i <= 32; i++) { const int x = squash(T, i * kMapXScale - kMapTScale); S.sse[0][i] = (uint16_t)x; S.sse[1][i] = (uint16_t)x; }
  S.lb = 0;
}

// one decision: forward path (PredictLow/High + PredictSSE) then, once the bit is known, Update + UpdateSSE
struct MapStep {
  uint16_t *c[5];
  int *w;
  int x[5], pd, q, xs0, xs1, pf;
  __device__ __forceinline__ int predict(MapState &S, const Tables &T)
  {
    long long sum = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) { x[k] = stretch(T, *c[k]); sum += (long long)(w[k] * x[k]); }        // mixer.h:78-88
    pd = clampi(squash(T, idiv_s64(sum, 16)), 1, PSCALEm);
    const int pq = min(2 * kMapTScale, max(0, stretch(T, pd) + kMapTScale));                          // sse.h:102-113
    q = pq / kMapXScale;
    const int pmod = pq - q * kMapXScale;
    const int pl = S.sse[S.lb][q], ph = S.sse[S.lb][q + 1];
    const int px = clampi((pl * (kMapXScale - pmod) + ph * pmod) / kMapXScale, 1, PSCALEm);
    xs0 = stretch(T, px); xs1 = x_of_pd(T);
    pf = clampi(squash(T, idiv_s64((long long)(S.fm[0] * xs0) + (long long)(S.fm[1] * xs1), 16)), 1, PSCALEm);
    return pf;
  }
  __device__ __forceinline__ int x_of_pd(const Tables &T) const { return stretch(T, pd); }
  __device__ __forceinline__ void update(MapState &S, int bit)
  {
#pragma unroll
    for (int k = 0; k < 5; k++) *c[k] = (uint16_t)counter16_upd(*c[k], bit, kRateCnt);                // map.cpp:54-62
    const int err = (bit << PBITS) - pd;
#pragma unroll
    for (int k = 0; k < 5; k++) w[k] = clampi(w[k] + idiv_s(idiv_s(x[k] * err, 12) * kRateMix, 12), -(1 << 19), (1 << 19) - 1);
    S.sse[S.lb][q] = (uint16_t)counter16_upd(S.sse[S.lb][q], bit, kRateSse);                          // sse.h:114-119
    S.sse[S.lb][q + 1] = (uint16_t)counter16_upd(S.sse[S.lb][q + 1], bit, kRateSse);
    S.lb = bit;
    const int ef = (bit << PBITS) - pf;
    S.fm[0] = clampi(S.fm[0] + idiv_s(idiv_s(xs0 * ef, 12) * kRateFinal, 12), -(1 << 19), (1 << 19) - 1);
    S.fm[1] = clampi(S.fm[1] + idiv_s(idiv_s(xs1 * ef, 12) * kRateFinal, 12), -(1 << 19), (1 << 19) - 1);
  }
};

// contexts (map.cpp:9-52). ul(i) = used[32768 - i], uh(i) = used[32768 + i]; index 0 of either is never set.
__device__ __forceinline__ void map_ctx_low(MapState &S, MapStep &m, const uint8_t *used, int i)
{
  const uint8_t *ul = used + kScale, *uh = used + kScale;
  const int c1 = ul[-(i - 1)], c2 = uh[i - 1], c3 = i > 1 ? ul[-(i - 2)] : 0;
  m.c[0] = &S.cnt[c1]; m.c[1] = &S.cnt[2 + c2]; m.c[2] = &S.cnt[4 + (c1 << 1) + c3]; m.c[3] = &S.cnt[8 + (c1 << 1) + c2];
  int sctx = c1;
  if (i > 1) sctx += ul[-(i - 2)] << 1;
  if (i > 2) sctx += ul[-(i - 3)] << 2;
  if (i > 3) sctx += ul[-(i - 4)] << 3;
  m.c[4] = &S.cctx[sctx];
  m.w = S.mixl[c1 + (c3 << 1)];
}
__device__ __forceinline__ void map_ctx_high(MapState &S, MapStep &m, const uint8_t *used, int i)
{
  const uint8_t *ul = used + kScale, *uh = used + kScale;
  const int c1 = uh[i - 1], c2 = ul[-i], c3 = i > 1 ? uh[i - 2] : 0;
  m.c[0] = &S.cnt[12 + c1]; m.c[1] = &S.cnt[12 + 2 + c2]; m.c[2] = &S.cnt[12 + 4 + (c1 << 1) + c3]; m.c[3] = &S.cnt[12 + 8 + (c1 << 1) + c2];
  int sctx = c1;
  if (i > 1) sctx += uh[i - 2] << 1;
  if (i > 2) sctx += uh[i - 3] << 2;
  if (i > 3) sctx += uh[i - 4] << 3;
  m.c[4] = &S.cctx[32 + sctx];
  m.w = S.mixh[c1 + (c3 << 1)];
}

__global__ void __launch_bounds__(32) map_encode_kernel(SparseJobs J, Tables T)
{
  __shared__ MapState S;
  const SparseJob &j = J.j[blockIdx.x];
  if (threadIdx.x != 0) return;
  // CalcRemapError's ratio and the r > 1.05 gate (libsac.cpp:243-250, 265): CostL1 = sum / n in double
  const double ent1 = (double)j.res->l1[0] / (double)j.n, ent2 = (double)j.res->l1[1] / (double)j.n;
  const double r = ent2 != 0.0 ? ent1 / ent2 : 1.0;
  if (!(r > 1.05)) { j.res->go = 0; j.rc->go = 0; return; }
  map_state_init(S, T);
  Coder rc;
  rc.range = 0xFFFFFFFFu; rc.ffnum = 0; rc.cache = 0; rc.lowc = 0; rc.nbytes = 0; rc.out = j.bytes;
  MapStep m;
  const uint8_t *used = j.used;
  for (int i = 1; i <= kScale; i++) {
    int bit = used[kScale - i];
    map_ctx_low(S, m, used, i);
    rc_encode<1>(rc, m.predict(S, T), bit, 0);
    m.update(S, bit);
    bit = used[kScale + i];
    map_ctx_high(S, m, used, i);
    rc_encode<1>(rc, m.predict(S, T), bit, 0);
    m.update(S, bit);
  }
  RcInit I;
  I.lowc = rc.lowc; I.nbytes = rc.nbytes; I.range = rc.range; I.ffnum = rc.ffnum; I.cache = rc.cache; I.code = 0; I.go = 1; I.pad = 0;
  *j.rc = I;
  j.res->go = 1;
}

__global__ void __launch_bounds__(32) map_decode_kernel(SparseJobs J, Tables T)
{
  __shared__ MapState S;
  const SparseJob &j = J.j[blockIdx.x];
  if (threadIdx.x != 0) return;
  map_state_init(S, T);
  const uint8_t *in = j.bytes;
  const long long in_len = j.in_len;
  long long ipos = 0;
  uint32_t range = 0xFFFFFFFFu, code = 0;
  auto getb = [&]() -> uint32_t { const uint32_t b = ipos < in_len ? in[ipos] : 0xffffffffu; ipos++; return b; };   // bufio.h:18-21
  for (int i = 0; i < 5; i++) code = (code << 8) + getb();
  auto decode = [&](int p1) -> int {                                                                  // range.cpp:73-80
    const uint32_t rnew = __umulhi(range, (uint32_t)(PSCALE - p1) << (32 - PBITS));
    const int b = code >= rnew;
    if (b) { range -= rnew; code -= rnew; } else range = rnew;
    while (range < 0x01000000u) { range <<= 8; code = (code << 8) + getb(); }
    return b;
  };
  MapStep m;
  uint8_t *used = j.used;
  for (int i = 1; i <= kScale; i++) {
    map_ctx_low(S, m, used, i);
    int bit = decode(m.predict(S, T));
    m.update(S, bit);
    used[kScale - i] = (uint8_t)bit;
    map_ctx_high(S, m, used, i);
    bit = decode(m.predict(S, T));
    m.update(S, bit);
    used[kScale + i] = (uint8_t)bit;
  }
  RcInit I;
  I.lowc = 0; I.nbytes = ipos; I.range = range; I.ffnum = 0; I.cache = 0; I.code = code; I.go = 1; I.pad = 0;
  *j.rc = I;
}

} // namespace

cudaError_t launch_sparse_encode(const BitplaneTables &bt, const SparseJobs &jobs, cudaStream_t stream)
{
  if (jobs.n <= 0) return cudaSuccess;
  Tables T{bt.d_stretch, bt.d_squash, bt.d_laplace, bt.lap_bits};
  int nmax = 0;
  for (int i = 0; i < jobs.n; i++) nmax = jobs.j[i].n > nmax ? jobs.j[i].n : nmax;
  const int blocks = (nmax + 255) / 256 < 592 ? (nmax + 255) / 256 : 592;
  remap_mark_kernel<<<dim3(blocks, jobs.n), 256, 0, stream>>>(jobs);
  remap_scan_kernel<<<jobs.n, 1024, 0, stream>>>(jobs);
  remap_map_kernel<<<dim3(blocks, jobs.n), 256, 0, stream>>>(jobs);
  map_encode_kernel<<<jobs.n, 32, 0, stream>>>(jobs, T);
  return cudaGetLastError();
}

cudaError_t launch_sparse_decode(const BitplaneTables &bt, const SparseJobs &jobs, cudaStream_t stream)
{
  if (jobs.n <= 0) return cudaSuccess;
  Tables T{bt.d_stretch, bt.d_squash, bt.d_laplace, bt.lap_bits};
  map_decode_kernel<<<jobs.n, 32, 0, stream>>>(jobs, T);
  remap_scan_kernel<<<jobs.n, 1024, 0, stream>>>(jobs);
  return cudaGetLastError();
}

} // namespace sacb
