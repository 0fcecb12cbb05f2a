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
struct Row { float4 c0, c1, c2, c3; double Bv, invC, eps, bound; };
template <int MODE>
__global__ void k(const Row* rows, int nRows, int reps, float* out, long long* cyc) {
  extern __shared__ Row s_rows[];
  for (int i = threadIdx.x; i < nRows; i += blockDim.x) s_rows[i] = rows[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  f3 vA = {0.01f * lane, 0.02f, 0.03f}, wA = {0.1f, 0.2f, 0.3f}, vB = {0.f, 0.f, 0.f}, wB = {0.05f, 0.01f, 0.02f};
  double lam = 0.0, acc = 0.0;
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
  out[blockIdx.x * blockDim.x + threadIdx.x] = vA.x + wA.y + vB.z + wB.x + (float)acc + (float)lam;
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
  for (int mode = 0; mode < 2; mode++)
    for (int threads : {32, 128, 256, 512}) {
      long long c = 0;
      for (int it = 0; it < 2; it++) {
        if (mode == 0) k<0><<<148, threads, sizeof(h)>>>(d, nRows, reps, out, cyc); else k<1><<<148, threads, sizeof(h)>>>(d, nRows, reps, out, cyc);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("mode %d (%s), %2d warps/SM: %.1f cycles per row step (%s)\n", mode, mode ? "2 lanes/unit" : "1 lane/unit", threads / 32,
             (double)c / (nRows * reps), cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
