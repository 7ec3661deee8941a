// Micro-benchmark: sustained DFMA / DADD issue rate per SM (independent chains, no memory traffic).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, int iters, double x, double y) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = fma(a[i], x, y);       // DFMA, 3 operands
      else if (MODE == 1) a[i] = a[i] + x;         // DADD
      else a[i] = fma(x, y, a[i]);                 // DFMA accumulate form (v * sel + acc)
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 12345.678) out[0] = s;
}
template <int MODE>
void run(const char* name, int sms, double mhz) {
  double* d; cudaMalloc(&d, 8);
  const int iters = 4096, blocks = sms * 8, threads = 256;
  k<MODE><<<blocks, threads>>>(d, 16, 1.0000001, 1e-9);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)blocks * threads * iters * 8;
  printf("%s: %.3f ms, %.2f lane-ops/clk/SM (at %.0f MHz), %.2f Tops/s\n", name, ms, ops / (ms * 1e-3) / sms / (mhz * 1e6), mhz,
         ops / (ms * 1e-3) / 1e12);
  cudaFree(d);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1e3;
  printf("%s, %d SMs, %.0f MHz\n", p.name, p.multiProcessorCount, mhz);
  run<0>("DFMA a*x+y", p.multiProcessorCount, mhz);
  run<1>("DADD a+x", p.multiProcessorCount, mhz);
  run<2>("DFMA x*y+a", p.multiProcessorCount, mhz);
  return 0;
}
