// Microbenchmark: vector FP64 (DFMA), F2F.F64.F32 and FFMA issue rates per SM on this GPU (lanes per clock per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu && ./fp64_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  float f0 = (float)a0, f1 = (float)a1, f2 = (float)a2, f3 = (float)a3, f4 = (float)a4, f5 = (float)a5, f6 = (float)a6, f7 = (float)a7;
  const double m = 1.0000001, c = 1e-9;
  const float mf = 1.0000001f, cf = 1e-9f;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    } else if (MODE == 1) {
      f0 = fmaf(f0, mf, cf); f1 = fmaf(f1, mf, cf); f2 = fmaf(f2, mf, cf); f3 = fmaf(f3, mf, cf);
      f4 = fmaf(f4, mf, cf); f5 = fmaf(f5, mf, cf); f6 = fmaf(f6, mf, cf); f7 = fmaf(f7, mf, cf);
    } else {
      // 8 conversions f32 -> f64 (+ 8 FADD to keep them alive and dependent)
      a0 = (double)f0; a1 = (double)f1; a2 = (double)f2; a3 = (double)f3; a4 = (double)f4; a5 = (double)f5; a6 = (double)f6; a7 = (double)f7;
      f0 += (float)__double2hiint(a0); f1 += (float)__double2hiint(a1); f2 += (float)__double2hiint(a2); f3 += (float)__double2hiint(a3);
      f4 += (float)__double2hiint(a4); f5 += (float)__double2hiint(a5); f6 += (float)__double2hiint(a6); f7 += (float)__double2hiint(a7);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7;
}
template <int MODE>
void run(const char* name, int sms, double ghz) {
  double* out;
  const int blocks = sms * 2, threads = 512, iters = 4096;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, iters, 1.0);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, iters, 1.0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)blocks * threads * iters * 8;
  printf("%s: %.3f ms, %.1f lane-ops per clock per SM (at %.2f GHz)\n", name, ms, ops / (ms * 1e-3) / sms / (ghz * 1e9), ghz);
  cudaFree(out);
}
int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  printf("%s, %d SMs, %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
  run<0>("DFMA", p.multiProcessorCount, ghz);
  run<1>("FFMA", p.multiProcessorCount, ghz);
  run<2>("F2F.F64.F32 (+I2F+FADD)", p.multiProcessorCount, ghz);
  return 0;
}
