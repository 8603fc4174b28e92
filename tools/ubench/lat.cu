// latency micro-benchmarks for the serial chains of the predictor (sm_100a): clocks per dependent operation
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }
__global__ void k(double *out, long long *t, int iters, double seed, int mode)
{
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = seed + i;
  __syncthreads();
  double a = seed + threadIdx.x * 1e-9, b = 1.0000001, c = 1e-9;
  int idx = threadIdx.x & 31;
  long long t0 = clk();
  if (mode == 0) for (int i = 0; i < iters; i++) a = __fma_rn(a, b, c);                       // dependent DFMA
  else if (mode == 1) for (int i = 0; i < iters; i++) a = 1.0 / a + 1.5;                        // dependent DDIV + DADD
  else if (mode == 2) for (int i = 0; i < iters; i++) { a = __shfl_xor_sync(0xffffffffu, a, 1) + c; }   // shfl(double)+DADD
  else if (mode == 3) for (int i = 0; i < iters; i++) { asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x)); }     // named barrier, 4 warps
  else if (mode == 4) for (int i = 0; i < iters; i++) { idx = (int)sm[idx & 1023] & 1023; }     // dependent LDS.64 + cvt
  else if (mode == 5) for (int i = 0; i < iters; i++) { sm[threadIdx.x] = a; __syncwarp(); a = sm[threadIdx.x ^ 1] + c; __syncwarp(); }  // st->ld
  else if (mode == 6) for (int i = 0; i < iters; i++) a = a * b;                                // dependent DMUL
  else if (mode == 7) for (int i = 0; i < iters; i++) { idx = idx * 3 + 1; idx &= 0xffff; }     // dependent IMAD+LOP
  else if (mode == 8) for (int i = 0; i < iters; i++) a = sqrt(a) + 1.5;
  else if (mode == 9) for (int i = 0; i < iters; i++) { __threadfence_block(); a = a + c; }
  long long t1 = clk();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + idx;
  if (threadIdx.x == 0) t[mode] = t1 - t0;
}
int main()
{
  double *o; long long *t, h[16];
  cudaMalloc(&o, 8 * 1024); cudaMalloc(&t, 8 * 16);
  const char *names[] = {"dep DFMA", "dep DDIV+DADD", "shfl.f64+DADD", "bar.sync (all warps)", "dep LDS.64+cvt", "STS->syncwarp->LDS+DADD->syncwarp", "dep DMUL", "IMAD+LOP", "DSQRT+DADD", "membar.cta+DADD"};
  const int iters = 4096;
  for (int threads : {32, 128}) {
    for (int m = 0; m < 10; m++) { k<<<1, threads>>>(o, t, iters, 1.25, m); }
    cudaDeviceSynchronize();
    for (int m = 0; m < 10; m++) { k<<<1, threads>>>(o, t, iters, 1.25, m); }
    cudaDeviceSynchronize();
    cudaMemcpy(h, t, sizeof(h), cudaMemcpyDeviceToHost);
    for (int m = 0; m < 10; m++) printf("threads=%d  %-36s %.1f clk/iter\n", threads, names[m], (double)h[m] / iters);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
