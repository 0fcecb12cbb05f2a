// Micro-benchmark + exhaustive check (profiling aid): throughput of F2F conversions vs integer-pipe widening / rounding,
// and bit-equality of the integer forms with the hardware conversions.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o cvt_rate cvt_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// (double)f for every f32 that is zero or normal; `bad` is set for denormals / inf / nan (caller takes the F2F path)
__device__ __forceinline__ double widen_int(float f, bool& bad) {
  const unsigned b = __float_as_uint(f);
  const unsigned ab = b & 0x7fffffffu;
  const unsigned e = ab >> 23;
  bad = (e == 0u && ab != 0u) || e == 255u;
  const unsigned hi = (b & 0x80000000u) | ((ab >> 3) + (ab ? 0x38000000u : 0u));
  return __hiloint2double((int)hi, (int)(b << 29));
}
// (double)(float)x for x whose f32 image is zero or normal; bad otherwise
__device__ __forceinline__ double round_int(double x, bool& bad) {
  unsigned lo = (unsigned)__double2loint(x), hi = (unsigned)__double2hiint(x);
  const unsigned u = hi << 1;  // exponent + top mantissa bits, sign dropped
  const bool inrange = (u - (897u << 21)) <= ((1149u - 897u) << 21) + 0x1fffffu;
  const bool zero = (u | lo) == 0u;
  bad = !(inrange || zero);
  const unsigned t = (lo >> 29) & 1u;
  const unsigned long long v = ((unsigned long long)hi << 32 | lo) + 0x0fffffffull + t;
  return __longlong_as_double((long long)(v & ~0x1fffffffull));
}

template <int MODE>
__global__ void thr(double* out, float fa, double da) {
  double x[8]; float f[8];
  for (int k = 0; k < 8; k++) { x[k] = da + k + threadIdx.x; f[k] = fa + k + threadIdx.x; }
#pragma unroll 1
  for (int i = 0; i < 512; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (MODE == 0) { x[k] = (double)f[k] + x[k]; f[k] = __uint_as_float(__float_as_uint(f[k]) ^ (unsigned)i); }        // F2F.F64.F32 + DADD
      else if (MODE == 1) { f[k] = (float)x[k]; x[k] = x[k] + da; asm volatile("" ::"f"(f[k])); }                        // F2F.F32.F64 + DADD
      else if (MODE == 2) { bool b; x[k] = widen_int(f[k], b) + x[k]; f[k] = __uint_as_float(__float_as_uint(f[k]) ^ (unsigned)i); }
      else if (MODE == 3) { bool b; x[k] = round_int(x[k], b) + da; }
      else { x[k] = x[k] + da; f[k] = __uint_as_float(__float_as_uint(f[k]) ^ (unsigned)i); }                            // DADD + LOP baseline
    }
  }
  double s = 0;
  for (int k = 0; k < 8; k++) s += x[k] + f[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void check_widen(unsigned long long* nbad, unsigned long long* nslow) {
  const unsigned long long n = 1ull << 32;
  unsigned long long bad = 0, slow = 0;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    const float f = __uint_as_float((unsigned)i);
    bool b;
    const double w = widen_int(f, b);
    if (b) { slow++; continue; }
    if (__double_as_longlong(w) != __double_as_longlong((double)f)) bad++;
  }
  atomicAdd(nbad, bad); atomicAdd(nslow, slow);
}
__device__ unsigned long long mix(unsigned long long z) { z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
__global__ void check_round(unsigned long long* nbad, unsigned long long* nslow, unsigned long long per) {
  unsigned long long bad = 0, slow = 0;
  const unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  for (unsigned long long i = 0; i < per; i++) {
    unsigned long long r = mix(tid * per + i + 0x9E3779B97F4A7C15ull);
    // half of the samples: arbitrary bit patterns; the rest: exponents near the f32 range edges and halfway cases
    unsigned long long bits = r;
    const unsigned sel = (unsigned)(mix(r) & 7);
    if (sel >= 4) {
      const unsigned long long e = sel == 4 ? 897 + (r % 3) - 1 : sel == 5 ? 1149 + (r % 4) - 1 : sel == 6 ? 1023 + (long long)(r % 60) - 30 : (r % 2047);
      bits = (r & 0x800fffffffffffffull) | (e << 52);
      if (mix(r + 1) & 1) bits = (bits & ~0x1fffffffull) | ((mix(r + 2) & 1) ? 0x10000000ull : 0x0fffffffull + (mix(r + 3) & 3));  // ties and neighbours
    }
    const double x = __longlong_as_double((long long)bits);
    bool b;
    const double w = round_int(x, b);
    if (b) { slow++; continue; }
    if (__double_as_longlong(w) != __double_as_longlong((double)(float)x)) bad++;
  }
  atomicAdd(nbad, bad); atomicAdd(nslow, slow);
}

int main() {
  double* out; cudaMalloc(&out, 1 << 24);
  unsigned long long* c; cudaMallocManaged(&c, 64);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  const char* names[] = {"F2F.F64.F32 + DADD", "F2F.F32.F64 + DADD", "int widen + DADD", "int round + DADD", "DADD + LOP (baseline)"};
  for (int mode = 0; mode < 5; mode++) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      switch (mode) {
        case 0: thr<0><<<sms * 4, 512>>>(out, 1.5f, 1.0000001); break;
        case 1: thr<1><<<sms * 4, 512>>>(out, 1.5f, 1.0000001); break;
        case 2: thr<2><<<sms * 4, 512>>>(out, 1.5f, 1.0000001); break;
        case 3: thr<3><<<sms * 4, 512>>>(out, 1.5f, 1.0000001); break;
        default: thr<4><<<sms * 4, 512>>>(out, 1.5f, 1.0000001); break;
      }
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
    }
    const double iters = (double)sms * 4 * 512 * 512 * 8;
    printf("%-26s %.3f ms  -> %.2f clk per warp-iteration per SMSP\n", names[mode], ms, ms * 1e-3 * p.clockRate * 1e3 / (iters / 32 / (sms * 4)));
  }
  c[0] = c[1] = 0;
  check_widen<<<sms * 8, 256>>>(c, c + 1);
  cudaDeviceSynchronize();
  printf("widen_int vs (double)f over all 2^32 floats: mismatches %llu, sent to the slow path %llu\n", c[0], c[1]);
  c[0] = c[1] = 0;
  check_round<<<sms * 8, 256>>>(c, c + 1, 1 << 14);
  cudaDeviceSynchronize();
  printf("round_int vs (double)(float)x over %.2e doubles: mismatches %llu, sent to the slow path %llu\n", (double)sms * 8 * 256 * (1 << 14), c[0], c[1]);
  return 0;
}
