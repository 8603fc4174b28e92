// cost.cu -- the cheap DDS objectives on residual arrays resident in HBM:
//   CostEntropy (/root/reference src/libsac/cost.h:70-116): order-0 empirical entropy in bytes
//   CostGolomb  (src/libsac/cost.h:43-66): adaptive Golomb bit count with a running mean (alpha = .97)
// CostL1 / CostRMS come from the sums the predictor kernel already accumulates; CostBitplane is bitplane.cu.
#include "engine.h"

namespace sacb {
namespace {

// one CTA per chain: histogram in HBM (zeroed here), then sum c*log2(c/n) over the alphabet
__global__ void __launch_bounds__(256) entropy_kernel(const int32_t *__restrict__ resid, size_t stride, const int *__restrict__ ns,
                                                     const int *__restrict__ ranges, unsigned int *hist, size_t hist_stride,
                                                     double *__restrict__ out)
{
  const int c = blockIdx.x;
  const int n = ns[c], R = ranges[c];           // residuals lie in [-R, R]
  const int32_t *e = resid + (size_t)c * stride;
  unsigned int *h = hist + (size_t)c * hist_stride;
  const int bins = 2 * R + 1;
  for (int i = threadIdx.x; i < bins; i += blockDim.x) h[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int v = e[i] + R;
    v = min(max(v, 0), bins - 1);
    atomicAdd(&h[v], 1u);
  }
  __syncthreads();
  const double invs = 1.0 / (double)n;
  double acc = 0.0;
  for (int i = threadIdx.x; i < bins; i += blockDim.x) {
    const unsigned int cnt = h[i];
    if (cnt) acc += (double)cnt * log2((double)cnt * invs);
  }
  __shared__ double red[8];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; w++) s += red[w];
    out[c] = n ? -s / 8.0 : 0.0;
  }
}

// one thread per chain (a strictly serial scalar recurrence)
__global__ void golomb_kernel(const int32_t *__restrict__ resid, size_t stride, const int *__restrict__ ns, int nchains,
                              double *__restrict__ out)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchains) return;
  const int n = ns[c];
  const int32_t *e = resid + (size_t)c * stride;
  double rm = 0.0;
  long long nbits = 0;
  for (int i = 0; i < n; i++) {
    const int m = max((int)rm, 1);
    const int v = e[i];
    const int u = v < 0 ? 2 * (-v) : (v > 0 ? 2 * v - 1 : 0);
    nbits += u / m + 1;
    if (m > 1) nbits += 32 - __clz((unsigned)m);
    rm = 0.97 * rm + (double)u;
  }
  out[c] = n ? (double)nbits / 8. : 0.0;
}

// 8 independent DFMA chains per thread: the fp64 pipe's issue-bound peak
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double seed)
{
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
    a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

} // namespace

cudaError_t launch_dfma_peak(double *out, int blocks, int iters, cudaStream_t stream)
{
  dfma_peak_kernel<<<blocks, 256, 0, stream>>>(out, iters, 1.0);
  return cudaGetLastError();
}

cudaError_t launch_entropy(const int32_t *resid, size_t stride, const int *ns, const int *ranges, int nchains, unsigned int *hist,
                           size_t hist_stride, double *out, cudaStream_t stream)
{
  entropy_kernel<<<nchains, 256, 0, stream>>>(resid, stride, ns, ranges, hist, hist_stride, out);
  return cudaGetLastError();
}
cudaError_t launch_golomb(const int32_t *resid, size_t stride, const int *ns, int nchains, double *out, cudaStream_t stream)
{
  golomb_kernel<<<(nchains + 63) / 64, 64, 0, stream>>>(resid, stride, ns, nchains, out);
  return cudaGetLastError();
}

} // namespace sacb
