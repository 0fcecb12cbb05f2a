// Micro-benchmark (profiling aid, not part of the library): cycles per projected Gauss-Seidel row update in the arithmetic
// of k_gs_exact (f64 on f32-stored operands, no FMA, f32 rounding of the body deltas), rows read from shared memory, for
// 1 / 4 / 16 warps per SM. MODE 0: one lane per unit (27 widenings + 12 roundings per row); MODE 1: two lanes per unit
// (lane A: body i, lane B: body j; one f64 shuffle pair per row).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o row_step row_step.cu
#include <cstdio>
#include <cuda_runtime.h>
struct f3 { float x, y, z; };
__device__ __forceinline__ double vdot(const f3& a, const f3& b) { return ((double)a.x * (double)b.x + (double)a.y * (double)b.y) + (double)a.z * (double)b.z; }
__device__ __forceinline__ f3 vaddscaled(const f3& a, double s, const f3& b) {
  f3 r; r.x = (float)((double)a.x + s * (double)b.x); r.y = (float)((double)a.y + s * (double)b.y); r.z = (float)((double)a.z + s * (double)b.z); return r;
}
__device__ __forceinline__ double shfl_xor_f64(double v, int m) {
  return __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(v), m), __shfl_xor_sync(0xffffffffu, __double2loint(v), m));
}
// (double)(float)x without the conversion pipe: x + M - M with M = 1.5 * 2^(e + 29), e = max(exponent of x, -126), rounds x at
// the ulp of its f32 image with the adder's round-to-nearest-even; the sign is copied back for results that round to zero.
// `bad` collects the exponent field so the caller can reject |x| >= 2^128 (f32 overflow), inf and nan.
__device__ __forceinline__ double round32(double x, unsigned& bad) {
  const unsigned hi = (unsigned)__double2hiint(x);
  const unsigned eb = hi & 0x7ff00000u;
  bad = max(bad, eb);
  const unsigned mh = max(eb, (1023u - 126u) << 20) + ((29u << 20) | 0x00080000u);
  const double M = __hiloint2double((int)mh, 0);
  const double t = (x + M) - M;
  return __hiloint2double((int)(((unsigned)__double2hiint(t) & 0x7fffffffu) | (hi & 0x80000000u)), __double2loint(t));
}
__global__ void check_round(unsigned long long seed, unsigned long long* nbad, unsigned long long* ntested) {
  unsigned long long bad = 0, n = 0;
  unsigned long long z = seed + (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull;
  for (int i = 0; i < 4096; i++) {
    z += 0x9E3779B97F4A7C15ull;
    unsigned long long r = z; r = (r ^ (r >> 30)) * 0xBF58476D1CE4E5B9ull; r = (r ^ (r >> 27)) * 0x94D049BB133111EBull; r ^= r >> 31;
    // exponent in the f32 range incl. subnormals and a margin below; mantissa: random, or a tie / near-tie pattern
    const unsigned long long sign = r >> 63;
    const unsigned long long e = 1023 - 160 + (r >> 40) % 290;  // 2^-160 .. 2^129
    unsigned long long m = r & 0xfffffffffffffull;
    const int kind = (r >> 52) & 7;
    if (kind == 0) m = (m & ~0x1fffffffull) | 0x10000000ull;        // exact tie
    else if (kind == 1) m = (m & ~0x1fffffffull) | 0x10000001ull;   // just above
    else if (kind == 2) m = (m & ~0x1fffffffull) | 0x0fffffffull;   // just below
    else if (kind == 3) m = (m | 0xfffffe0000000ull);               // carries into the next binade
    double x = __longlong_as_double((long long)((sign << 63) | (e << 52) | m));
    if (i == 0) x = 0.0; if (i == 1) x = -0.0;
    unsigned flag = 0;
    const double y = round32(x, flag);
    const double ref = (double)(float)x;
    if (flag >= 0x47f00000u) continue;  // caller's overflow guard
    n++;
    if (__double_as_longlong(y) != __double_as_longlong(ref)) bad++;
  }
  atomicAdd(nbad, bad); atomicAdd(ntested, n);
}
struct Row { float4 c0, c1, c2, c3; double Bv, invC, eps, bound; };
template <int MODE>
__global__ void k(const Row* rows, int nRows, int reps, float* out, long long* cyc, int activeLanes) {
  extern __shared__ Row s_rows[];
  for (int i = threadIdx.x; i < nRows; i += blockDim.x) s_rows[i] = rows[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  if (lane >= activeLanes) return;  // partial warps: does the conversion pipe charge for inactive lanes?
  f3 vA = {0.01f * lane, 0.02f, 0.03f}, wA = {0.1f, 0.2f, 0.3f}, vB = {0.f, 0.f, 0.f}, wB = {0.05f, 0.01f, 0.02f};
  double lam = 0.0, acc = 0.0;
  double dstate[12] = {0.01 * lane, 0.02, 0.03, 0.1, 0.2, 0.3, 0, 0, 0, 0.05, 0.01, 0.02};
  unsigned ovf = 0;
  const double imA = 1.0, imB = 0.5;
  const long long t0 = clock64();
  for (int rep = 0; rep < reps; rep++) {
#pragma unroll 1
    for (int r = 0; r < nRows; r++) {
      const Row& q = s_rows[r];
      const float4 q0 = q.c0, q1 = q.c1, q2 = q.c2, q3 = q.c3;
      const double Bv = q.Bv, invC = q.invC, eps = q.eps, bound = q.bound;
      f3 n = {q0.x, q0.y, q0.z}, rA = {q1.x, q1.y, q1.z}, rB = {q2.x, q2.y, q2.z}, iB = {q3.x, q3.y, q3.z}, iA = {q1.w, q2.w, q3.w};
      f3 sA = {-n.x, -n.y, -n.z};
      if (MODE == 0) {
        const double gwl = (vdot(vA, sA) + vdot(wA, rA)) + (vdot(vB, n) + vdot(wB, rB));
        double dl = invC * (Bv - gwl - eps * lam);
        if (lam + dl < 0.0) dl = -lam; else if (lam + dl > bound) dl = bound - lam;
        lam += dl;
        vA = vaddscaled(vA, imA * dl, sA); wA = vaddscaled(wA, dl, iA);
        vB = vaddscaled(vB, imB * dl, n); wB = vaddscaled(wB, dl, iB);
        acc += dl > 0.0 ? dl : -dl;
      } else if (MODE == 2) {
        // f64 state that always holds f32-representable values, rounded with round32 (no conversions on the chain)
        double* st = dstate;
        const double nx = q0.x, ny = q0.y, nz = q0.z, rAx = q1.x, rAy = q1.y, rAz = q1.z, rBx = q2.x, rBy = q2.y, rBz = q2.z;
        const double iAx = q1.w, iAy = q2.w, iAz = q3.w, iBx = q3.x, iBy = q3.y, iBz = q3.z;
        const double gwl = (((st[0] * -nx + st[1] * -ny) + st[2] * -nz) + ((st[3] * rAx + st[4] * rAy) + st[5] * rAz)) +
                           (((st[6] * nx + st[7] * ny) + st[8] * nz) + ((st[9] * rBx + st[10] * rBy) + st[11] * rBz));
        double dl = invC * (Bv - gwl - eps * lam);
        if (lam + dl < 0.0) dl = -lam; else if (lam + dl > bound) dl = bound - lam;
        lam += dl;
        const double dA = imA * dl, dB = imB * dl;
        st[0] = round32(st[0] + dA * -nx, ovf); st[1] = round32(st[1] + dA * -ny, ovf); st[2] = round32(st[2] + dA * -nz, ovf);
        st[3] = round32(st[3] + dl * iAx, ovf); st[4] = round32(st[4] + dl * iAy, ovf); st[5] = round32(st[5] + dl * iAz, ovf);
        st[6] = round32(st[6] + dB * nx, ovf); st[7] = round32(st[7] + dB * ny, ovf); st[8] = round32(st[8] + dB * nz, ovf);
        st[9] = round32(st[9] + dl * iBx, ovf); st[10] = round32(st[10] + dl * iBy, ovf); st[11] = round32(st[11] + dl * iBz, ovf);
        acc += dl > 0.0 ? dl : -dl;
      } else {
        const bool a = (lane & 1) == 0;
        const f3 jl = a ? sA : n, jr = a ? rA : rB, ju = a ? iA : iB;
        double p = vdot(vA, jl) + vdot(wA, jr);
        const double gwl = p + shfl_xor_f64(p, 1);
        double dl = invC * (Bv - gwl - eps * lam);
        if (lam + dl < 0.0) dl = -lam; else if (lam + dl > bound) dl = bound - lam;
        lam += dl;
        vA = vaddscaled(vA, (a ? imA : imB) * dl, jl); wA = vaddscaled(wA, dl, ju);
        acc += dl > 0.0 ? dl : -dl;
      }
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = vA.x + wA.y + vB.z + wB.x + (float)acc + (float)lam + (float)(dstate[0] + dstate[4] + dstate[8] + dstate[11]) + ovf;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  const int nRows = 30, reps = 200;
  Row h[nRows];
  for (int i = 0; i < nRows; i++) {
    h[i].c0 = make_float4(0.3f, 0.5f + 0.01f * i, -0.2f, 0.f); h[i].c1 = make_float4(0.1f, -0.2f, 0.05f * i, 0.3f);
    h[i].c2 = make_float4(-0.1f, 0.2f, 0.15f, 0.1f); h[i].c3 = make_float4(0.2f, 0.1f, -0.3f, 0.2f);
    h[i].Bv = 0.1 * (i % 3 - 1); h[i].invC = 0.7; h[i].eps = 1e-3; h[i].bound = 1e6;
  }
  Row* d; float* out; long long* cyc;
  cudaMalloc(&d, sizeof(h)); cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int mode = 0; mode < 3; mode++)
    for (int threads : {32, 128, 256, 512}) {
      long long c = 0;
      for (int it = 0; it < 2; it++) {
        if (mode == 0) k<0><<<148, threads, sizeof(h)>>>(d, nRows, reps, out, cyc, 32); else if (mode == 1) k<1><<<148, threads, sizeof(h)>>>(d, nRows, reps, out, cyc, 32); else k<2><<<148, threads, sizeof(h)>>>(d, nRows, reps, out, cyc, 32);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("mode %d (%s), %2d warps/SM: %.1f cycles per row step (%s)\n", mode, mode == 0 ? "1 lane/unit" : mode == 1 ? "2 lanes/unit" : "1 lane/unit, f64 state + add-magic rounding", threads / 32,
             (double)c / (nRows * reps), cudaGetErrorString(cudaGetLastError()));
    }
  {
    unsigned long long *cnt, hc[2] = {0, 0};
    cudaMalloc(&cnt, 16); cudaMemset(cnt, 0, 16);
    for (int it = 0; it < 64; it++) check_round<<<148 * 8, 256>>>(0x1234567ull + it * 0x51ull, cnt, cnt + 1);
    cudaDeviceSynchronize();
    cudaMemcpy(hc, cnt, 16, cudaMemcpyDeviceToHost);
    printf("round32 vs (double)(float)x over %llu doubles (ties, near-ties, carries, subnormal images, +-0): mismatches %llu (%s)\n", hc[1], hc[0], cudaGetErrorString(cudaGetLastError()));
  }
  for (int threads : {32})
    for (int act : {32, 16, 8, 4, 1}) {
      long long c = 0;
      for (int it = 0; it < 2; it++) { k<0><<<148, threads, sizeof(h)>>>(d, nRows, reps, out, cyc, act); cudaDeviceSynchronize(); }
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("mode 0, %2d warps/SM, %2d active lanes per warp: %.1f cycles per row step\n", threads / 32, act, (double)c / (nRows * reps));
    }
  return 0;
}
