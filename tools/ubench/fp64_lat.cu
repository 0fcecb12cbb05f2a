// Micro-benchmark (profiling aid, not part of the library): latency of dependent f64 / conversion chains and per-SM
// throughput of DADD / F2F on the device at hand. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, double a, double b, float f) {
  double x = a;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1024; i++) { x = x + b; x = x + b; x = x + b; x = x + b; }
  long long t1 = clock64();
  double y = a;
#pragma unroll 1
  for (int i = 0; i < 1024; i++) { y = y * b; y = y * b; y = y * b; y = y * b; }
  long long t2 = clock64();
  double z = a;
#pragma unroll 1
  for (int i = 0; i < 1024; i++) { z = (double)(float)z; z = z + b; z = (double)(float)z; z = z + b; }
  long long t3 = clock64();
  float g = f;
#pragma unroll 1
  for (int i = 0; i < 1024; i++) { g = g + f; g = g + f; g = g + f; g = g + f; }
  long long t4 = clock64();
  out[threadIdx.x] = x + y + z + g;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
}
// throughput: many independent chains per thread, full SM occupancy
template <int MODE>
__global__ void thr(double* out, double a, double b) {
  double x[8];
  for (int k = 0; k < 8; k++) x[k] = a + k + threadIdx.x;
#pragma unroll 1
  for (int i = 0; i < 512; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (MODE == 0) x[k] = x[k] + b;
      else if (MODE == 1) x[k] = (double)(float)x[k];
      else x[k] = x[k] * b + a;
    }
  }
  double s = 0;
  for (int k = 0; k < 8; k++) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1 << 24); cudaMallocManaged(&cyc, 64);
  lat<<<1, 32>>>(out, cyc, 1.0, 1.0000001, 1.5f);
  cudaDeviceSynchronize();
  printf("latency per op (cycles): DADD %.1f  DMUL %.1f  [F2F.F32.F64 + F2F.F64.F32 + DADD] %.1f  FADD %.1f\n", cyc[0] / 4096.0, cyc[1] / 4096.0, cyc[2] / 2048.0, cyc[3] / 4096.0);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  for (int mode = 0; mode < 3; mode++) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      if (mode == 0) thr<0><<<sms * 4, 512>>>(out, 1.0, 1.0000001);
      else if (mode == 1) thr<1><<<sms * 4, 512>>>(out, 1.0, 1.0000001);
      else thr<2><<<sms * 4, 512>>>(out, 1.0, 1.0000001);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)sms * 4 * 512 * 512 * 8 * (mode == 1 ? 2 : (mode == 2 ? 2 : 1));
    double clk = p.clockRate * 1e3;
    printf("mode %d (%s): %.3f ms, %.1f thread-ops/clk/SM (clock %.0f MHz)\n", mode, mode == 0 ? "DADD" : mode == 1 ? "F2F pair" : "DMUL+DADD", ms,
           ops / (ms * 1e-3) / clk / sms, clk / 1e6);
  }
  return 0;
}
