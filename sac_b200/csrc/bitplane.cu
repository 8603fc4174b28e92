// bitplane.cu -- the residual entropy coder (bitplanes, most significant first; adaptive counters -> logistic
// mixers -> two SSE stages -> final mixer -> binary range coder) as sm_100a kernels, one warp per channel stream.
//
// Restates /root/reference src/libsac/vle.cpp:3-261 (BitplaneCoder), src/model/counter.h:17-69, domain.h:7-61,
// mixer.h:59-101, sse.h:84-125, range.cpp:54-92, src/libsac/cost.h:144-175 (CostBitplane). Integer arithmetic is
// reproduced exactly; tests compare payload bytes with the reference's own BitplaneCoder.
//
// Encoder structure ("contexts are a pure function of the residual array"): per plane the
// stream is walked in chunks of 32 samples. All 32 lanes of the first pipeline warp compute, in parallel, everything
// about "their" sample that does not depend on adaptive state -- windowed magnitude mean (prefix sums), Laplace
// estimate (table), significance patterns and counts (ballots), refinement contexts -- and prefetch the one
// large-table counter (csig0, 64K entries, HBM/L2). The adaptive chain of the 32 decisions then runs as a pipeline
// of four warps (counters -> mixer -> SSE -> final mixer + coder), see bitplane_pipe_kernel.
// The decoder cannot look ahead (contexts depend on bits decoded in the current plane) and evaluates each
// decision cooperatively across one warp instead.
#include "bitplane.h"
#include "model_dev.cuh"
#include "sac_canon_math.h"
#include <cuda_runtime.h>
#include <cstdlib>

namespace sacb {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kTScale = 2662, kXScale = 380;                       // SSENL<15>: tscale=myDomain.max, xscale=2*tscale/14
constexpr int kLimP = 150, kLimSig = 300, kLimRef = 150;           // vle.h:43-45
constexpr int kRateRef = 800, kRateSig = 700, kRateSse = 250, kRateSseMix = 250;   // vle.h:46-49

__constant__ uint16_t c_div[304];                                  // PSCALE/(i+3)   (counter.h:40-51)
__constant__ uint16_t c_plap_init[32];                             // vle.cpp:16-22
__constant__ uint16_t c_sse_init[16];                              // sse.h:92-97

// per-stream adaptive state in shared memory
struct BpState {
  uint32_t plap[32];                                               // counters: p1 | count<<16
  uint32_t csig1[128], cref0[32], cref1[256], cref2[64], cref3[256];
  int mixref[32][5];
  int mixsig[128][3];
  int ssemix[2];
  uint16_t sse[160][2][16];
  uint32_t sse_lb[160];
};


// LinearCounterLimit::update (counter.h:58-68)
__device__ __forceinline__ uint32_t counter_upd(uint32_t c, int bit, int limit)
{
  int p1 = (int)(c & 0xffffu), cnt = (int)(c >> 16);
  if (cnt < limit) cnt++;
  const int dv = c_div[cnt];
  const int dp = bit ? ((PSCALE - p1) * dv) >> PBITS : -((p1 * dv) >> PBITS);
  p1 = clampi(p1 + dp, 1, PSCALEm);
  return (uint32_t)p1 | ((uint32_t)cnt << 16);
}

// BitplaneCoder::PredictLaplace (vle.cpp:70-79) in canonical math
__device__ __forceinline__ int laplace_calc(uint32_t avg_sum, int bpn)
{
  double p_l = 0.0;
  if (avg_sum > 0) {
    const double theta = sac_canon::c_exp(-1.0 / (double)avg_sum);
    p_l = 1.0 - 1.0 / (1 + sac_canon::c_pow(theta, (double)(1 << bpn)));
  }
  return min(max((int)sac_canon::c_round(p_l * PSCALE), 1), PSCALEm);
}
__device__ __forceinline__ int laplace_p(const Tables &T, uint32_t avg_sum, int bpn)
{
  if (bpn <= T.lap_bits - 1 && avg_sum < (1u << T.lap_bits)) return __ldg(T.laplace + ((size_t)bpn << T.lap_bits) + avg_sum);
  return laplace_calc(avg_sum, bpn);
}

struct Decision {                                                  // unpacked per-sample record
  int is_ref, bit, pe, sse2;
  int ctx1, n2, mixctx;                                            // sig branch
  int tb, c1, c2, c3, pctx;                                        // ref branch
};

// mixer + SSE + final mix for one decision. cs = current csig0 counter value (sig branch). Returns final p1 and
// performs every model update; new csig0 value in cs.
template <class CodeFn>
__device__ __forceinline__ int model_step(BpState &S, const Tables &T, const Decision &d, int bpn, uint32_t &cs, CodeFn &&code)
{
  int x[5], nin, p_mix;
  int *w;
  uint32_t *pc1 = nullptr, *pc2 = nullptr, *pc3 = nullptr, *pc4 = nullptr;
  const uint32_t plv = S.plap[bpn];
  if (d.is_ref) {                                                  // vle.cpp:81-129
    pc1 = &S.cref0[d.tb]; pc2 = &S.cref1[d.c1]; pc3 = &S.cref2[d.c2]; pc4 = &S.cref3[d.c3];
    w = S.mixref[d.pctx];
    nin = 5;
    x[0] = stretch(T, d.pe); x[1] = stretch(T, plv & 0xffff); x[2] = stretch(T, *pc1 & 0xffff);
    x[3] = stretch(T, *pc2 & 0xffff); x[4] = stretch(T, *pc3 & 0xffff);
  } else {                                                         // vle.cpp:159-175
    pc2 = &S.csig1[d.n2];
    w = S.mixsig[d.mixctx];
    nin = 3;
    x[0] = stretch(T, plv & 0xffff); x[1] = stretch(T, cs & 0xffff); x[2] = stretch(T, *pc2 & 0xffff);
    x[3] = 0; x[4] = 0;
  }
  {
    long long sum = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) if (i < nin) sum += (long long)(w[i] * x[i]);
    p_mix = clampi(squash(T, idiv_s64(sum, 16)), 1, PSCALEm);      // mixer.h:74-85
  }
  // ---- SSE (vle.cpp:188-197, sse.h:99-112) ----
  const int sse1 = ((d.pe >> 11) << 1) + d.is_ref;
  const int sse2 = d.sse2;
  const int lb1 = S.sse_lb[sse1], lb2 = S.sse_lb[sse2];
  int q1, q2, pr1, pr2;
  {
    const int pq = min(2 * kTScale, max(0, stretch(T, p_mix) + kTScale));
    q1 = pq / kXScale;
    const int pm = pq - q1 * kXScale;
    const int pl = S.sse[sse1][lb1][q1], ph = S.sse[sse1][lb1][q1 + 1];
    pr1 = clampi((pl * (kXScale - pm) + ph * pm) / kXScale, 1, PSCALEm);
  }
  {
    const int pq = min(2 * kTScale, max(0, stretch(T, pr1) + kTScale));
    q2 = pq / kXScale;
    const int pm = pq - q2 * kXScale;
    const int pl = S.sse[sse2][lb2][q2], ph = S.sse[sse2][lb2][q2 + 1];
    pr2 = clampi((pl * (kXScale - pm) + ph * pm) / kXScale, 1, PSCALEm);
  }
  const int xs0 = stretch(T, (pr1 + pr2 + 1) >> 1), xs1 = stretch(T, p_mix);
  const int pfin = clampi(squash(T, idiv_s64((long long)(S.ssemix[0] * xs0) + (long long)(S.ssemix[1] * xs1), 16)), 1, PSCALEm);

  const int bit = code(pfin);

  // ---- updates (vle.cpp:131-140,177-186,199-204) ----
  S.plap[bpn] = counter_upd(plv, bit, kLimP);
  {
    const int err = (bit << PBITS) - p_mix;
    const int rate = d.is_ref ? kRateRef : kRateSig;
#pragma unroll
    for (int i = 0; i < 5; i++)
      if (i < nin) {
        const int de = idiv_s(x[i] * err, 12);
        w[i] = clampi(w[i] + idiv_s(de * rate, 12), -(1 << 19), (1 << 19) - 1);
      }
  }
  if (d.is_ref) {
    *pc1 = counter_upd(*pc1, bit, kLimRef);
    *pc2 = counter_upd(*pc2, bit, kLimRef);
    *pc3 = counter_upd(*pc3, bit, kLimRef);
    *pc4 = counter_upd(*pc4, bit, kLimRef);
  } else {
    cs = counter_upd(cs, bit, kLimSig);
    *pc2 = counter_upd(*pc2, bit, kLimSig);
  }
  S.sse[sse1][lb1][q1] = (uint16_t)counter16_upd(S.sse[sse1][lb1][q1], bit, kRateSse);
  S.sse[sse1][lb1][q1 + 1] = (uint16_t)counter16_upd(S.sse[sse1][lb1][q1 + 1], bit, kRateSse);
  S.sse_lb[sse1] = bit;
  S.sse[sse2][lb2][q2] = (uint16_t)counter16_upd(S.sse[sse2][lb2][q2], bit, kRateSse);
  S.sse[sse2][lb2][q2 + 1] = (uint16_t)counter16_upd(S.sse[sse2][lb2][q2 + 1], bit, kRateSse);
  S.sse_lb[sse2] = bit;
  {
    const int err = (bit << PBITS) - pfin;
    S.ssemix[0] = clampi(S.ssemix[0] + idiv_s(idiv_s(xs0 * err, 12) * kRateSseMix, 12), -(1 << 19), (1 << 19) - 1);
    S.ssemix[1] = clampi(S.ssemix[1] + idiv_s(idiv_s(xs1 * err, 12) * kRateSseMix, 12), -(1 << 19), (1 << 19) - 1);
  }
  return bit;
}

__device__ void init_state(BpState &S, uint32_t *csig0, int lane)
{
  const uint32_t c0 = (uint32_t)(PSCALE >> 1);
  for (int i = lane; i < 32; i += 32) { S.plap[i] = c_plap_init[i]; S.cref0[i] = c0; }
  for (int i = lane; i < 128; i += 32) S.csig1[i] = c0;
  for (int i = lane; i < 256; i += 32) { S.cref1[i] = c0; S.cref3[i] = c0; }
  for (int i = lane; i < 64; i += 32) S.cref2[i] = c0;
  for (int i = lane; i < 32 * 5; i += 32) (&S.mixref[0][0])[i] = 0;
  for (int i = lane; i < 128 * 3; i += 32) (&S.mixsig[0][0])[i] = 0;
  if (lane < 2) S.ssemix[lane] = 0;
  for (int i = lane; i < 160 * 32; i += 32) (&S.sse[0][0][0])[i] = c_sse_init[i & 15];
  for (int i = lane; i < 160; i += 32) S.sse_lb[i] = 0;
  uint4 *c4 = reinterpret_cast<uint4 *>(csig0);
  const uint4 v4 = make_uint4(c0, c0, c0, c0);
  for (int i = lane; i < 65536 / 4; i += 32) c4[i] = v4;
  __syncwarp();
}

__device__ __forceinline__ uint32_t scan_incl(uint32_t v, int lane)
{
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(kFull, v, o); if (lane >= o) v += t; }
  return v;
}
__device__ __forceinline__ int topbit(uint32_t x) { return 31 - __clz(x); }   // -1 for 0

constexpr int kWarpsPerCta = 2;
constexpr int kStretchBytes = PSCALE * 2, kSquashBytes = 4096 * 2;   // shared-memory copies of the LogDomain tables

// copies the LogDomain tables to the head of the CTA's dynamic shared memory; returns the table view to use
__device__ __forceinline__ Tables stage_tables(const Tables &G, unsigned char *smem)
{
  const uint4 *gs = reinterpret_cast<const uint4 *>(G.stretch);
  uint4 *ss = reinterpret_cast<uint4 *>(smem);
  for (int i = threadIdx.x; i < kStretchBytes / 16; i += blockDim.x) ss[i] = __ldg(gs + i);
  int16_t *sq = reinterpret_cast<int16_t *>(smem + kStretchBytes);
  for (int i = threadIdx.x; i < 4095; i += blockDim.x) sq[i] = __ldg(G.squash + i);
  __syncthreads();
  Tables T = G;
  T.stretch = reinterpret_cast<const int16_t *>(smem);
  T.squash = sq;
  return T;
}

// ------------------------------------------------------------------------------------------------------------------
// Pipelined encoder / byte counter: one CTA (4 warps) per stream. The adaptive chain of a decision splits into four
// recurrences that only feed forward, so each gets its own warp and they run one chunk (32 decisions) apart:
//
//   W1  contexts of the chunk (all lanes, as above) + the bit-driven counters and their stretch() inputs
//   W2  logistic mixer (vle.cpp:120-129,166-175, mixer.h:59-101)
//   W3  the two SSE stages (vle.cpp:188-197, sse.h:84-125)
//   W4  final mixer + range coder (vle.cpp:197-204, range.cpp:66-92)
//
// Records travel through double-buffered rings in shared memory (one slot per decision, published per chunk).
// Inside a warp the forward path is evaluated uniformly; the state updates of one decision (up to 5 counters,
// 5 mixer weights, 4 SSE knots) are spread over lanes, one update each. Arithmetic is identical to model_step.
struct PipeShared {
  BpState st;
  uint32_t mixpad[4];                                              // a sig-branch mixer reads weights [3],[4] (times 0)
  uint16_t divtab[304];
  uint32_t dummy;
  uint4 r12[64];
  uint2 r23[64];
  uint2 r34[64];
  int pub12, con12, pub23, con23, pub34, con34;
  uint32_t red[4];
};

__device__ __forceinline__ int pld(const int *p) { return *reinterpret_cast<const volatile int *>(p); }
__device__ __forceinline__ void pst(int *p, int v) { *reinterpret_cast<volatile int *>(p) = v; }
__device__ __forceinline__ void pipe_wait_gt(const int *ctr, int v)
{
  while (pld(ctr) <= v) __nanosleep(40);
  __threadfence_block();
}
__device__ __forceinline__ void pipe_publish(int *ctr, int v, int lane)
{
  __syncwarp();
  __threadfence_block();
  if (lane == 0) pst(ctr, v);
}
__device__ __forceinline__ int sx16(uint32_t v) { return (int)(int16_t)(v & 0xffffu); }

// LinearCounterLimit::update with the divisor table in shared memory (lanes update different counters)
__device__ __forceinline__ uint32_t counter_upd_s(const uint16_t *divtab, uint32_t c, int bit, int limit)
{
  int p1 = (int)(c & 0xffffu), cnt = (int)(c >> 16);
  if (cnt < limit) cnt++;
  const int dv = divtab[cnt];
  const int dp = bit ? ((PSCALE - p1) * dv) >> PBITS : -((p1 * dv) >> PBITS);
  p1 = clampi(p1 + dp, 1, PSCALEm);
  return (uint32_t)p1 | ((uint32_t)cnt << 16);
}

constexpr int kPipeThreads = 128;                                  // threads per stream (4 pipeline warps)
#ifndef SACB_PIPE_STREAMS
#define SACB_PIPE_STREAMS 2
#endif
constexpr int kPipeStreams = SACB_PIPE_STREAMS;                    // streams per CTA: they share the 72 KB of LogDomain tables (make EXTRA=-DSACB_PIPE_STREAMS=n to experiment)
static_assert(kPipeStreams >= 1 && kPipeStreams <= 7, "named barriers 1..7 and 227 KB of shared memory bound the streams per CTA");

template <int MODE>
__global__ void __launch_bounds__(kPipeThreads * kPipeStreams) bitplane_pipe_kernel(const BpJob *__restrict__ jobs, int njobs, Tables TG)
{
  extern __shared__ __align__(16) unsigned char bp_smem[];
  const Tables T = stage_tables(TG, bp_smem);
  const int sidx = threadIdx.x / kPipeThreads;                     // stream of this CTA
  const int job = blockIdx.x * kPipeStreams + sidx;
  PipeShared &P = reinterpret_cast<PipeShared *>(bp_smem + kStretchBytes + kSquashBytes)[sidx];
  BpState &S = P.st;
  const int tid = threadIdx.x - sidx * kPipeThreads, lane = tid & 31, warp = tid >> 5;
  if (job >= njobs) return;
  const BpJob J = jobs[job];
  if (J.rc_init && !J.rc_init->go) return;                         // mapped record not wanted (whole stream leaves, before any stream barrier)
  const int sbar = 1 + sidx;                                       // named barrier of this stream
  const int n = J.n;
  int32_t *u = J.buf;

  // ---- S2U map (utils.h:248-253) and maxbpn (libsac.cpp:429-441, cost.h:150-158), all threads ----
  uint32_t vmax = 0;
  if (J.signed_input) {
    for (int i = tid; i < n; i += kPipeThreads) {
      const int32_t v = u[i];
      const int32_t m = v < 0 ? 2 * (-v) : (v > 0 ? 2 * v - 1 : 0);
      u[i] = m;
      vmax = max(vmax, (uint32_t)m);
    }
  } else {
    for (int i = tid; i < n; i += kPipeThreads) vmax = max(vmax, (uint32_t)u[i]);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) vmax = max(vmax, __shfl_xor_sync(kFull, vmax, o));
  if (lane == 0) P.red[warp] = vmax;
  // ---- state ----
  {
    const uint32_t c0 = (uint32_t)(PSCALE >> 1);
    for (int i = tid; i < 32; i += kPipeThreads) { S.plap[i] = c_plap_init[i]; S.cref0[i] = c0; }
    for (int i = tid; i < 128; i += kPipeThreads) S.csig1[i] = c0;
    for (int i = tid; i < 256; i += kPipeThreads) { S.cref1[i] = c0; S.cref3[i] = c0; }
    for (int i = tid; i < 64; i += kPipeThreads) S.cref2[i] = c0;
    for (int i = tid; i < 32 * 5; i += kPipeThreads) (&S.mixref[0][0])[i] = 0;
    for (int i = tid; i < 128 * 3; i += kPipeThreads) (&S.mixsig[0][0])[i] = 0;
    for (int i = tid; i < 160 * 32; i += kPipeThreads) (&S.sse[0][0][0])[i] = c_sse_init[i & 15];
    for (int i = tid; i < 160; i += kPipeThreads) S.sse_lb[i] = 0;
    for (int i = tid; i < 304; i += kPipeThreads) P.divtab[i] = c_div[i];
    if (tid < 4) P.mixpad[tid] = 0;
    if (tid == 0) { S.ssemix[0] = 0; S.ssemix[1] = 0; P.dummy = 0; P.pub12 = 0; P.con12 = 0; P.pub23 = 0; P.con23 = 0; P.pub34 = 0; P.con34 = 0; }
    uint4 *c4 = reinterpret_cast<uint4 *>(J.csig0);
    const uint4 v4 = make_uint4(c0, c0, c0, c0);
    for (int i = tid; i < 65536 / 4; i += kPipeThreads) c4[i] = v4;
  }
  asm volatile("bar.sync %0, %1;" ::"r"(sbar), "r"(kPipeThreads) : "memory");
  vmax = max(max(P.red[0], P.red[1]), max(P.red[2], P.red[3]));
  const int maxbpn = J.maxbpn >= 0 ? J.maxbpn : max(topbit(vmax), 0);
  const int nchunks = (n + 31) >> 5;
  // mixer weights as one int array: mixref rows (5 ints) then mixsig rows (3 ints); BpState keeps them adjacent
  int *mixw = &S.mixref[0][0];
  constexpr int kSigBase = 32 * 5;

  if (warp == 0) {
    // ================================ W1: contexts + counters ======================================================
    uint32_t *sw = &S.plap[0];                                     // plap | csig1 | cref0 | cref1 | cref2 | cref3 (adjacent)
    constexpr int oCsig1 = 32, oCref0 = 160, oCref1 = 192, oCref2 = 448, oCref3 = 512;
    const int oDummy = (int)(&P.dummy - sw);
    int q = 0;
    for (int bpn = maxbpn; bpn >= 0; bpn--) {
      const int shA = bpn + 1, shB = max(bpn, 1);
      const uint32_t himask = ~((2u << bpn) - 1u);
      uint32_t u_prev = 0, u_cur = 0, u_next = lane < n ? (uint32_t)u[lane] : 0u;
      uint32_t pre_cur = 0, tot_cur = 0, hi_cur = 0, suf_prev = 0;
      uint32_t pre_next, tot_next;
      {
        const uint32_t h = u_next & himask;
        pre_next = scan_incl(h, lane);
        tot_next = __shfl_sync(kFull, pre_next, 31);
      }
      uint32_t A_prev = 0, A_cur = 0, A_next = __ballot_sync(kFull, (u_next >> shA) != 0);
      uint32_t B_prev = 0, B_cur = 0;
      uint32_t K_prev = 0, K_cur = 0;
      for (int c = 0; c < nchunks; c++, q++) {
        const int base = c << 5, s = base + lane;
        suf_prev = tot_cur - pre_cur + hi_cur;
        u_prev = u_cur; u_cur = u_next;
        pre_cur = pre_next; tot_cur = tot_next; hi_cur = u_cur & himask;
        A_prev = A_cur; A_cur = A_next; B_prev = B_cur; K_prev = K_cur;
        {
          const int sn = s + 32;
          u_next = sn < n ? (uint32_t)u[sn] : 0u;
          const uint32_t h = u_next & himask;
          pre_next = scan_incl(h, lane);
          tot_next = __shfl_sync(kFull, pre_next, 31);
          A_next = __ballot_sync(kFull, (u_next >> shA) != 0);
        }
        B_cur = __ballot_sync(kFull, (u_cur >> shB) != 0);
        K_cur = __ballot_sync(kFull, ((u_cur >> bpn) & 1u) != 0);
        const unsigned long long AL = ((unsigned long long)A_cur << 32) | A_prev;
        const unsigned long long BL = ((unsigned long long)B_cur << 32) | B_prev;
        const unsigned long long KL = ((unsigned long long)K_cur << 32) | K_prev;
        const unsigned long long AR = ((unsigned long long)A_next << 32) | A_cur;
        const uint32_t a_left = (uint32_t)(AL >> lane), b_left = (uint32_t)(BL >> lane), k_left = (uint32_t)(KL >> lane);
        const uint32_t a_right = (uint32_t)(AR >> (lane + 1));
        uint32_t a_right_cnt = a_right;
        { const int dl = (n - 1) - (s + 1); if (dl >= 0 && dl < 32) a_right_cnt &= ~(1u << dl); }
        auto leftbit = [&](uint32_t wbits, int dd) { return (int)((wbits >> (32 - dd)) & 1u); };
        auto rightbit = [&](uint32_t wbits, int dd) { return (int)((wbits >> (dd - 1)) & 1u); };
        const int selfA = (int)((A_cur >> lane) & 1u);
        const int bit_l = (int)((u_cur >> bpn) & 1u);
        const uint32_t nsum = suf_prev + tot_cur + pre_next + ((uint32_t)__popc(k_left) << bpn);
        const int nidx = min(s + 32, n - 1) - max(s - 32, 0) + 1;
        const uint32_t avg = s < n ? (nsum + (uint32_t)(nidx - 1)) / (uint32_t)nidx : 0u;
        const int pe_l = laplace_p(T, avg, bpn);
        auto nb = [&](int dd) -> uint32_t {
          const int sl = lane + dd;
          const uint32_t vc = __shfl_sync(kFull, u_cur, sl & 31);
          const uint32_t vp = __shfl_sync(kFull, u_prev, sl & 31);
          const uint32_t vn = __shfl_sync(kFull, u_next, sl & 31);
          return sl < 0 ? vp : (sl > 31 ? vn : vc);
        };
        const uint32_t l1 = nb(-1), l2 = nb(-2), l3 = nb(-3), l4 = nb(-4);
        const uint32_t r1 = nb(1), r2 = nb(2), r3 = nb(3), r4 = nb(4);
        uint32_t rec0, rec1, cs_val = 0;
        const int sse2_l = 32 + selfA + (leftbit(b_left, 1) << 1) + (rightbit(a_right, 1) << 2) + (leftbit(b_left, 2) << 3) +
                           (rightbit(a_right, 2) << 4) + (leftbit(b_left, 3) << 5) + (rightbit(a_right, 3) << 6);
        rec0 = (uint32_t)pe_l | ((uint32_t)bit_l << 15) | ((uint32_t)selfA << 16) | ((uint32_t)sse2_l << 17);
        if (selfA) {
          const int b0 = (int)(u_cur >> (bpn + 1)), b1 = (int)(l1 >> bpn), b2 = (int)(r1 >> (bpn + 1));
          const int b3 = (int)(l2 >> bpn), b4 = (int)(r2 >> (bpn + 1));
          const int c0 = (b0 << 1) < b1, c1 = b0 < b2, c2 = (b0 << 1) < b3, c3 = b0 < b4;
          const int x0 = b0 << 1, x1 = b1, x2 = b2 << 1, x3 = b3, x4 = b4 << 1;
          const int xm = (x0 + x1 + x2 + x3 + x4) / 5;
          const int d0 = x0 > xm, d1 = x1 > xm;
          const int cc1 = (b0 & 15) + ((b1 & 15) << 4);
          const int cc2 = (c0 + (c1 << 1) + (c2 << 2) + (c3 << 3)) + (d0 << 4) + (d1 << 5);
          auto msbL = [&](uint32_t v) { return (v >> shB) != 0 ? topbit(v) : 0; };
          auto msbR = [&](uint32_t v) { return (v >> shA) != 0 ? topbit(v) : 0; };
          const int cc3 = msbL(l1) + msbR(r1) + msbL(l2) + msbR(r2) + msbL(l3) + msbR(r3) + msbL(l4) + msbR(r4);
          const int pctx = ((((pe_l >> 12) << 1) + d0) << 1) + (b0 & 1);
          rec1 = (uint32_t)topbit(u_cur) | ((uint32_t)cc1 << 5) | ((uint32_t)cc2 << 13) | ((uint32_t)cc3 << 19) | ((uint32_t)pctx << 27);
        } else {
          int ctx1 = 0;
#pragma unroll
          for (int dd = 1; dd <= 8; dd++) ctx1 |= (leftbit(b_left, dd) << (2 * dd - 2)) | (rightbit(a_right, dd) << (2 * dd - 1));
          const int n1 = __popc(b_left) + __popc(a_right_cnt);
          const int n2 = __popc(a_left) + __popc(a_right_cnt);
          int st4 = 0;
#pragma unroll
          for (int dd = 1; dd <= 4; dd++) st4 |= ((s - dd >= 0 && !leftbit(a_left, dd)) ? 1 : 0) << (dd - 1);
          const int mixctx = (st4 << 3) + (min(n1, 3) << 1) + (n2 > 0 ? 1 : 0);
          rec1 = (uint32_t)ctx1 | ((uint32_t)n2 << 16) | ((uint32_t)mixctx << 23);
          if (s < n) cs_val = J.csig0[ctx1];
        }
        // ---- ring slot free? then the serial part: counters of each decision, inputs to the mixer ----
        if (q >= 2) pipe_wait_gt(&P.con12, q - 2);
        uint4 *ring = P.r12 + ((q & 1) << 5);
        const int cnt = min(32, n - base);
        for (int j = 0; j < cnt; j++) {
          const uint32_t w0 = __shfl_sync(kFull, rec0, j), w1 = __shfl_sync(kFull, rec1, j);
          const uint32_t cs = __shfl_sync(kFull, cs_val, j);
          const int pe = (int)(w0 & 0x7fff), bit = (int)((w0 >> 15) & 1), is_ref = (int)((w0 >> 16) & 1), sse2 = (int)((w0 >> 17) & 0xff);
          const int sse1 = ((pe >> 11) << 1) + is_ref;
          // lane roles. lane 0: p_laplace[bpn]; ref: lanes 1-4 cref0..3, lane 6 carries pestimate; sig: lane 1 csig1, lanes >= 5 csig0
          int off = oDummy, limit = kLimSig, pos = -1, woff;
          if (is_ref) {
            const int tb = (int)(w1 & 31), c1 = (int)((w1 >> 5) & 0xff), c2 = (int)((w1 >> 13) & 0x3f), c3 = (int)((w1 >> 19) & 0xff);
            woff = (int)(w1 >> 27) * 5;
            limit = kLimRef;
            if (lane == 1) { off = oCref0 + tb; pos = 2; }
            else if (lane == 2) { off = oCref1 + c1; pos = 3; }
            else if (lane == 3) { off = oCref2 + c2; pos = 4; }
            else if (lane == 4) off = oCref3 + c3;
            else if (lane == 6) pos = 0;
            if (lane == 0) pos = 1;
          } else {
            const int n2 = (int)((w1 >> 16) & 0x7f);
            woff = kSigBase + (int)((w1 >> 23) & 0x7f) * 3;
            if (lane == 1) { off = oCsig1 + n2; pos = 2; }
            else if (lane == 2) pos = 3;
            else if (lane == 3) pos = 4;
            else if (lane == 5) pos = 1;
            if (lane == 0) pos = 0;
          }
          if (lane == 0) { off = bpn; limit = kLimP; }
          uint32_t v = sw[off];
          if (lane >= 5) v = (is_ref && lane == 6) ? (uint32_t)pe : cs;
          int sx = stretch(T, (int)(v & 0xffffu));
          if (!is_ref && (lane == 2 || lane == 3)) sx = 0;
          uint16_t *slot16 = reinterpret_cast<uint16_t *>(ring + j);
          if (pos >= 0) slot16[pos] = (uint16_t)sx;
          if (lane == 7) {
            uint32_t *slot32 = reinterpret_cast<uint32_t *>(ring + j);
            slot32[3] = (uint32_t)woff | ((uint32_t)is_ref << 10) | ((uint32_t)bit << 11) | ((uint32_t)sse1 << 12) | ((uint32_t)sse2 << 20);
          }
          const uint32_t nv = counter_upd_s(P.divtab, v, bit, limit);
          if (lane < 5 && off != oDummy) sw[off] = nv;
          if (!is_ref) {
            const uint32_t csn = __shfl_sync(kFull, nv, 31);
            const int ctx1 = (int)(w1 & 0xffff);
            if (lane == 5) J.csig0[ctx1] = csn;
            if (!selfA && lane > j && (int)(rec1 & 0xffff) == ctx1) cs_val = csn;   // keep prefetched copies coherent
          }
          __syncwarp();
        }
        pipe_publish(&P.pub12, q + 1, lane);
      }
    }
  } else if (warp == 1) {
    // ================================ W2: mixer =====================================================================
    int q = 0;
    for (int bpn = maxbpn; bpn >= 0; bpn--) {
      for (int c = 0; c < nchunks; c++, q++) {
        const int cnt = min(32, n - (c << 5));
        pipe_wait_gt(&P.pub12, q);
        if (q >= 2) pipe_wait_gt(&P.con23, q - 2);
        const uint4 *rin = P.r12 + ((q & 1) << 5);
        uint2 *rout = P.r23 + ((q & 1) << 5);
        for (int j = 0; j < cnt; j++) {
          const uint4 r = rin[j];
          const int x0 = sx16(r.x), x1 = sx16(r.x >> 16), x2 = sx16(r.y), x3 = sx16(r.y >> 16), x4 = sx16(r.z);
          const int woff = (int)(r.w & 0x3ff), is_ref = (int)((r.w >> 10) & 1), bit = (int)((r.w >> 11) & 1);
          int *w = mixw + woff;
          long long sum = (long long)(w[0] * x0) + (long long)(w[1] * x1) + (long long)(w[2] * x2);
          if (is_ref) sum += (long long)(w[3] * x3) + (long long)(w[4] * x4);
          const int p_mix = clampi(squash(T, idiv_s64(sum, 16)), 1, PSCALEm);
          const int stp = stretch(T, p_mix);
          if (lane == 0) rout[j] = make_uint2((uint32_t)p_mix | ((uint32_t)(stp & 0xffff) << 16), r.w >> 11);
          const int err = (bit << PBITS) - p_mix;
          const int rate = is_ref ? kRateRef : kRateSig;
          const int nin = is_ref ? 5 : 3;
          __syncwarp();                                            // every lane has read the weights
          if (lane < nin) {
            const int xi = lane == 0 ? x0 : (lane == 1 ? x1 : (lane == 2 ? x2 : (lane == 3 ? x3 : x4)));
            const int de = idiv_s(xi * err, 12);
            w[lane] = clampi(w[lane] + idiv_s(de * rate, 12), -(1 << 19), (1 << 19) - 1);
          }
          __syncwarp();
        }
        __syncwarp();
        __threadfence_block();
        if (lane == 0) { pst(&P.con12, q + 1); pst(&P.pub23, q + 1); }
      }
    }
  } else if (warp == 2) {
    // ================================ W3: SSE =======================================================================
    int q = 0;
    for (int bpn = maxbpn; bpn >= 0; bpn--) {
      for (int c = 0; c < nchunks; c++, q++) {
        const int cnt = min(32, n - (c << 5));
        pipe_wait_gt(&P.pub23, q);
        if (q >= 2) pipe_wait_gt(&P.con34, q - 2);
        const uint2 *rin = P.r23 + ((q & 1) << 5);
        uint2 *rout = P.r34 + ((q & 1) << 5);
        for (int j = 0; j < cnt; j++) {
          const uint2 r = rin[j];
          const int stp = sx16(r.x >> 16);
          const int bit = (int)(r.y & 1), sse1 = (int)((r.y >> 1) & 0xff), sse2 = (int)((r.y >> 9) & 0xff);
          const int lb1 = S.sse_lb[sse1], lb2 = S.sse_lb[sse2];
          uint16_t *k1 = &S.sse[sse1][lb1][0], *k2 = &S.sse[sse2][lb2][0];
          int q1, q2, pr1, pr2;
          {
            const int pq = min(2 * kTScale, max(0, stp + kTScale));
            q1 = pq / kXScale;
            const int pm = pq - q1 * kXScale;
            const int pl = k1[q1], ph = k1[q1 + 1];
            pr1 = clampi((pl * (kXScale - pm) + ph * pm) / kXScale, 1, PSCALEm);
          }
          {
            const int pq = min(2 * kTScale, max(0, stretch(T, pr1) + kTScale));
            q2 = pq / kXScale;
            const int pm = pq - q2 * kXScale;
            const int pl = k2[q2], ph = k2[q2 + 1];
            pr2 = clampi((pl * (kXScale - pm) + ph * pm) / kXScale, 1, PSCALEm);
          }
          const int xs0 = stretch(T, (pr1 + pr2 + 1) >> 1);
          if (lane == 0) rout[j] = make_uint2((uint32_t)(xs0 & 0xffff) | ((uint32_t)(stp & 0xffff) << 16), (uint32_t)bit);
          __syncwarp();                                            // every lane has read the knots
          if (lane < 4) {
            uint16_t *kp = (lane < 2 ? k1 + q1 : k2 + q2) + (lane & 1);
            *kp = (uint16_t)counter16_upd(*kp, bit, kRateSse);
          } else if (lane == 4) S.sse_lb[sse1] = bit;
          else if (lane == 5) S.sse_lb[sse2] = bit;
          __syncwarp();
        }
        __syncwarp();
        __threadfence_block();
        if (lane == 0) { pst(&P.con23, q + 1); pst(&P.pub34, q + 1); }
      }
    }
  } else {
    // ================================ W4: final mixer + coder =======================================================
    Coder rc;
    rc.range = 0xFFFFFFFFu; rc.ffnum = 0; rc.cache = 0; rc.lowc = 0; rc.nbytes = 0; rc.out = J.out;
    if (J.rc_init) {                                               // the map coder's bytes are already in J.out[0, nbytes)
      const RcInit I = *J.rc_init;
      rc.range = I.range; rc.ffnum = I.ffnum; rc.cache = I.cache; rc.lowc = I.lowc; rc.nbytes = I.nbytes;
    }
    int m0 = 0, m1 = 0;
    int q = 0;
    for (int bpn = maxbpn; bpn >= 0; bpn--) {
      for (int c = 0; c < nchunks; c++, q++) {
        const int cnt = min(32, n - (c << 5));
        pipe_wait_gt(&P.pub34, q);
        const uint2 *rin = P.r34 + ((q & 1) << 5);
        const uint2 mine = lane < cnt ? rin[lane] : make_uint2(0u, 0u);
        __syncwarp();
        __threadfence_block();
        if (lane == 0) pst(&P.con34, q + 1);                       // records are in registers now
        for (int j = 0; j < cnt; j++) {
          const uint32_t rx = __shfl_sync(kFull, mine.x, j), ry = __shfl_sync(kFull, mine.y, j);
          const int xs0 = sx16(rx), xs1 = sx16(rx >> 16), bit = (int)(ry & 1);
          const int pfin = clampi(squash(T, idiv_s64((long long)(m0 * xs0) + (long long)(m1 * xs1), 16)), 1, PSCALEm);
          rc_encode<MODE>(rc, pfin, bit, lane);
          const int err = (bit << PBITS) - pfin;
          m0 = clampi(m0 + idiv_s(idiv_s(xs0 * err, 12) * kRateSseMix, 12), -(1 << 19), (1 << 19) - 1);
          m1 = clampi(m1 + idiv_s(idiv_s(xs1 * err, 12) * kRateSseMix, 12), -(1 << 19), (1 << 19) - 1);
        }
      }
    }
    for (int i = 0; i < 5; i++) shift_low<MODE>(rc, lane);         // RangeCoderSH::Stop (range.cpp:61-64)
    if (lane == 0) {
      if (J.nbytes) *J.nbytes = rc.nbytes;
      if (J.maxbpn_out) *J.maxbpn_out = maxbpn;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// decoder: one warp per stream, each decision evaluated cooperatively (vle.cpp:233-261)
__global__ void __launch_bounds__(kWarpsPerCta * 32) bitplane_decode_kernel(const BpJob *__restrict__ jobs, int njobs, Tables TG)
{
  extern __shared__ __align__(16) unsigned char bp_smem[];
  const Tables T = stage_tables(TG, bp_smem);
  BpState *states = reinterpret_cast<BpState *>(bp_smem + kStretchBytes + kSquashBytes);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int job = blockIdx.x * kWarpsPerCta + wib;
  if (job >= njobs) return;
  const BpJob J = jobs[job];
  BpState &S = states[wib];
  const int n = J.n, maxbpn = J.maxbpn;
  volatile int32_t *buf = J.buf;                                   // magnitudes accumulate here
  volatile uint8_t *msb = J.msb;                                   // plane at which a sample became significant
  for (int i = lane; i < n; i += 32) { buf[i] = 0; msb[i] = 0; }
  init_state(S, J.csig0, lane);

  const uint8_t *in = J.in;
  const long long in_len = J.in_len;
  long long ipos = 0;
  uint32_t range = 0xFFFFFFFFu, code = 0;
  auto getb = [&]() -> uint32_t { const uint32_t b = ipos < in_len ? in[ipos] : 0xffffffffu; ipos++; return b; };
  if (J.rc_init) { const RcInit I = *J.rc_init; range = I.range; code = I.code; ipos = I.nbytes; }   // after MapEncoder::Decode
  else for (int i = 0; i < 5; i++) code = (code << 8) + getb();

  for (int bpn = maxbpn; bpn >= 0; bpn--) {
    uint32_t state = 0;
    const uint32_t mL = ~((1u << bpn) - 1u), mR = ~((2u << bpn) - 1u);
    for (int s = 0; s < n; s++) {
      // lanes cover offsets -32..-1 (k = s-32+lane) and +1..+32 (k = s+1+lane); lane 0 also covers k = s
      const int kl = s - 32 + lane, kr = s + 1 + lane;
      const uint32_t vl = kl >= 0 ? (uint32_t)buf[kl] : 0u;
      const uint32_t vr = kr < n ? (uint32_t)buf[kr] : 0u;
      const int ml = kl >= 0 ? msb[kl] : 0;
      const int mr = kr < n ? msb[kr] : 0;
      const uint32_t vs = (uint32_t)buf[s];
      const int ms = msb[s];
      uint32_t part = (vl & mL) + (vr & mR) + (lane == 0 ? (vs & mR) : 0u);
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(kFull, part, o);
      const int nidx = min(s + 32, n - 1) - max(s - 32, 0) + 1;
      const uint32_t avg = (part + (uint32_t)(nidx - 1)) / (uint32_t)nidx;
      const int pe = laplace_p(T, avg, bpn);
      const uint32_t sigL = __ballot_sync(kFull, ml != 0);         // bit i <-> k = s-32+i
      const uint32_t sigR = __ballot_sync(kFull, mr != 0);         // bit i <-> k = s+1+i
      const uint32_t bigL = __ballot_sync(kFull, ml > bpn);
      const uint32_t bigR = __ballot_sync(kFull, mr > bpn);
      uint32_t sigRc = sigR, bigRc = bigR;
      { const int dl = (n - 1) - (s + 1); if (dl >= 0 && dl < 32) { sigRc &= ~(1u << dl); bigRc &= ~(1u << dl); } }
      auto Lv = [&](int dd) { return __shfl_sync(kFull, vl, 32 - dd); };   // value at s-dd
      auto Rv = [&](int dd) { return __shfl_sync(kFull, vr, dd - 1); };
      auto Lm = [&](int dd) { return __shfl_sync(kFull, ml, 32 - dd); };
      auto Rm = [&](int dd) { return __shfl_sync(kFull, mr, dd - 1); };
      Decision d;
      d.pe = pe; d.bit = 0; d.is_ref = ms != 0;
      d.sse2 = 32 + (ms ? 1 : 0) + (int)((sigL >> 31) & 1) * 2 + (int)(sigR & 1) * 4 + (int)((sigL >> 30) & 1) * 8 + (int)((sigR >> 1) & 1) * 16 +
               (int)((sigL >> 29) & 1) * 32 + (int)((sigR >> 2) & 1) * 64;
      d.ctx1 = d.n2 = d.mixctx = d.tb = d.c1 = d.c2 = d.c3 = d.pctx = 0;
      uint32_t cs = 0;
      if (d.is_ref) {
        const uint32_t l1 = Lv(1), l2 = Lv(2), r1 = Rv(1), r2 = Rv(2);
        const int b0 = (int)(vs >> (bpn + 1)), b1 = (int)(l1 >> bpn), b2 = (int)(r1 >> (bpn + 1)), b3 = (int)(l2 >> bpn), b4 = (int)(r2 >> (bpn + 1));
        const int c0 = (b0 << 1) < b1, c1 = b0 < b2, c2 = (b0 << 1) < b3, c3 = b0 < b4;
        const int x0 = b0 << 1, x1 = b1, x2 = b2 << 1, x3 = b3, x4 = b4 << 1;
        const int xm = (x0 + x1 + x2 + x3 + x4) / 5;
        const int d0 = x0 > xm, d1 = x1 > xm;
        d.tb = ms; d.c1 = (b0 & 15) + ((b1 & 15) << 4);
        d.c2 = (c0 + (c1 << 1) + (c2 << 2) + (c3 << 3)) + (d0 << 4) + (d1 << 5);
        d.c3 = Lm(1) + Rm(1) + Lm(2) + Rm(2) + Lm(3) + Rm(3) + Lm(4) + Rm(4);
        d.pctx = ((((pe >> 12) << 1) + d0) << 1) + (b0 & 1);
      } else {
        int ctx1 = 0;
#pragma unroll
        for (int dd = 1; dd <= 8; dd++) ctx1 |= (int)((sigL >> (32 - dd)) & 1) << (2 * dd - 2) | (int)((sigR >> (dd - 1)) & 1) << (2 * dd - 1);
        const int n1 = __popc(sigL) + __popc(sigRc), n2 = __popc(bigL) + __popc(bigRc);
        d.ctx1 = ctx1; d.n2 = n2;
        d.mixctx = (int)((state & 15) << 3) + (min(n1, 3) << 1) + (n2 > 0 ? 1 : 0);
        cs = J.csig0[ctx1];
      }
      const int bit = model_step(S, T, d, bpn, cs, [&](int p) {
        const uint32_t rnew = __umulhi(range, (uint32_t)(PSCALE - p) << (32 - PBITS));   // range.cpp:73-80
        const int b = code >= rnew;
        if (b) { range -= rnew; code -= rnew; } else range = rnew;
        while (range < 0x01000000u) { range <<= 8; code = (code << 8) + getb(); }
        return b;
      });
      if (d.is_ref) state = (state << 1);
      else { state = (state << 1) + 1; if (lane == 0) J.csig0[d.ctx1] = cs; }
      if (bit && lane == 0) {
        buf[s] = (int32_t)(vs + (1u << bpn));
        if (!d.is_ref) msb[s] = (uint8_t)bpn;
      }
      __syncwarp();
    }
  }
  for (int i = lane; i < n; i += 32) {                             // U2S (utils.h:254-259)
    const int32_t v = buf[i];
    buf[i] = (v & 1) ? ((v + 1) >> 1) : -(v >> 1);
  }
}

// fills the Laplace table rows in canonical math
__global__ void laplace_table_kernel(uint16_t *tab, int lap_bits)
{
  const size_t total = (size_t)lap_bits << lap_bits;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int bpn = (int)(i >> lap_bits);
    const uint32_t avg = (uint32_t)(i & ((1u << lap_bits) - 1));
    tab[i] = (uint16_t)laplace_calc(avg, bpn);
  }
}

} // namespace

size_t bitplane_state_bytes() { return sizeof(BpState); }

// ---- host side ---------------------------------------------------------------------------------------------------
// LogDomain (domain.h:17-31) in canonical math; equality with the reference's libm tables is asserted in tests
void compute_logdomain_tables(int16_t *st, int16_t *sq)
{
  using namespace sac_canon;
  for (int i = 0; i < PSCALE; i++) {
    const double f = (i > 1 ? i : 1) / (double)PSCALE;
    st[i] = (int16_t)c_round(c_log(f / (1.0 - f)) * 256);
  }
  for (int i = -2047; i <= 2047; i++) sq[i + 2047] = (int16_t)c_round(PSCALE / (1.0 + c_exp(-double(i) / 256.0)));
}

cudaError_t BitplaneTables::init(cudaStream_t stream)
{
  using namespace sac_canon;
  std::vector<int16_t> st(PSCALE), sq(4095);
  compute_logdomain_tables(st.data(), sq.data());
  uint16_t dv[304], pl[32], si[16];
  for (int i = 0; i < 304; i++) dv[i] = (uint16_t)(PSCALE / (i + 3));
  for (int i = 0; i < 32; i++) {
    const int sh = (int)(1u << i);
    int p = (int)c_round((1.0 - 1.0 / (1 + c_pow(0.99, (double)sh))) * PSCALE);
    pl[i] = (uint16_t)(p < 1 ? 1 : (p > PSCALEm ? PSCALEm : p));
  }
  for (int i = 0; i < 16; i++) {
    const int x = i * kXScale - kTScale;
    si[i] = (uint16_t)(x < -2047 ? 1 : (x > 2047 ? PSCALEm : sq[x + 2047]));
  }
  h_stretch = st; h_squash = sq;
  cudaError_t e;
  if ((e = cudaMemcpyToSymbolAsync(c_div, dv, sizeof(dv), 0, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbolAsync(c_plap_init, pl, sizeof(pl), 0, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbolAsync(c_sse_init, si, sizeof(si), 0, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&d_stretch, PSCALE * 2)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&d_squash, 4095 * 2)) != cudaSuccess) return e;
  lap_bits = 18;
  if ((e = cudaMalloc(&d_laplace, ((size_t)lap_bits << lap_bits) * 2)) != cudaSuccess) return e;
  if ((e = cudaMemcpyAsync(d_stretch, st.data(), PSCALE * 2, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
  if ((e = cudaMemcpyAsync(d_squash, sq.data(), 4095 * 2, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
  laplace_table_kernel<<<592, 256, 0, stream>>>(d_laplace, lap_bits);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  return cudaStreamSynchronize(stream);
}
void BitplaneTables::destroy()
{
  cudaFree(d_stretch); cudaFree(d_squash); cudaFree(d_laplace);
  d_stretch = nullptr; d_squash = nullptr; d_laplace = nullptr;
}

cudaError_t bitplane_init_attributes()
{
  const int smem = (int)(sizeof(BpState) * kWarpsPerCta) + kStretchBytes + kSquashBytes;
  const int psmem = (int)sizeof(PipeShared) * kPipeStreams + kStretchBytes + kSquashBytes;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(bitplane_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(bitplane_pipe_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, psmem)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(bitplane_pipe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, psmem);
}

cudaError_t launch_bitplane(const BitplaneTables &bt, const BpJob *d_jobs, int njobs, int mode, cudaStream_t stream)
{
  Tables T{bt.d_stretch, bt.d_squash, bt.d_laplace, bt.lap_bits};
  const int grid = (njobs + kWarpsPerCta - 1) / kWarpsPerCta;
  const int smem = (int)(sizeof(BpState) * kWarpsPerCta) + kStretchBytes + kSquashBytes;
  if (mode != 2) {
    const int psmem = (int)sizeof(PipeShared) * kPipeStreams + kStretchBytes + kSquashBytes;
    const int pgrid = (njobs + kPipeStreams - 1) / kPipeStreams;
    if (mode == 0) bitplane_pipe_kernel<0><<<pgrid, kPipeThreads * kPipeStreams, psmem, stream>>>(d_jobs, njobs, T);
    else bitplane_pipe_kernel<1><<<pgrid, kPipeThreads * kPipeStreams, psmem, stream>>>(d_jobs, njobs, T);
    return cudaGetLastError();
  }
  bitplane_decode_kernel<<<grid, kWarpsPerCta * 32, smem, stream>>>(d_jobs, njobs, T);
  return cudaGetLastError();
}

} // namespace sacb
