// k_sph.cuh — SPHSystem.update (lib/objects/sph_system.dart:62-163), a World.subsystem that runs after gravity and
// before the broadphase (world_class.dart:472-475). SURVEY.md 8f rank 4.
//
// The reference is two sequential loops over the particles with an O(N) neighbour search each; every particle only
// writes its own density / pressure (first loop) and its own force (second loop), so one thread per particle reproduces
// it exactly: the neighbour list is never materialised, both kernels re-walk the particle list in its order (neighbours
// in list order, the particle itself last). The second loop reads pressures[j] / densities[j] with j = the POSITION in
// the neighbour list (sph_system.dart:131-133,142), not the neighbour's own index - reproduced as written.
// math.pow(x, 2 | 3) is evaluated correctly rounded (q*q; pow3_cr), pow(h, 9) comes from the host's libm (SphDev.h9).
#pragma once
#include "world.cuh"

struct SphDev {
  const int* particles;  // body indices, SPHSystem.particles order
  double* densities;
  double* pressures;
  int n;
  double density, h, h9, cs, viscosity, eps;
};

__device__ __forceinline__ double pow3_cr(double x) {  // x^3 with one rounding: exact square (hi + lo), double-double times x
  const double hi = x * x, lo = __fma_rn(x, x, -hi);
  const double p = hi * x, e = __fma_rn(hi, x, -p);
  return p + (e + lo * x);
}

__global__ void __launch_bounds__(128) k_sph_density(BodyArrays B, SphDev S) {
  const double r2 = S.h * S.h;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S.n; i += gridDim.x * blockDim.x) {
    const int bi = S.particles[i];
    const f3 pp = ld3(B.pos[bi]);
    double sum = 0.0;
    for (int k = 0; k <= S.n; k++) {  // k == n: the particle itself, appended after its neighbours (:75)
      const int bk = k < S.n ? S.particles[k] : bi;
      if (k < S.n) {
        if (bk == bi) continue;
        const f3 dist = vsub(ld3(B.pos[bk]), pp);  // getNeighbors :55-58
        if (!(vlen2(dist) < r2)) continue;
      }
      const f3 dist = vsub(pp, ld3(B.pos[bk]));
      const double len = vlen(dist);
      const double weight = (315.0 / (64.0 * 3.141592653589793 * S.h9)) * pow3_cr(S.h * S.h - len * len);  // w(), :166-170
      sum += B.mass[bk] * weight;
    }
    S.densities[i] = sum;
    S.pressures[i] = S.cs * S.cs * (sum - S.density);
  }
}

__global__ void __launch_bounds__(128) k_sph_forces(BodyArrays B, SphDev S) {
  const double r2 = S.h * S.h, h = S.h;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S.n; i += gridDim.x * blockDim.x) {
    const int bi = S.particles[i];
    const f3 pp = ld3(B.pos[bi]), pv = ld3(B.vel[bi]);
    const double di = S.densities[i], pi = S.pressures[i];
    f3 aPressure = mk3(0.0, 0.0, 0.0), aVisc = aPressure;
    int j = 0;  // position in the neighbour list
    for (int k = 0; k <= S.n; k++) {
      const int bk = k < S.n ? S.particles[k] : bi;
      if (k < S.n) {
        if (bk == bi) continue;
        const f3 dist = vsub(ld3(B.pos[bk]), pp);
        if (!(vlen2(dist) < r2)) continue;
      }
      const f3 rVec = vsub(pp, ld3(B.pos[bk]));
      const double r = vlen(rVec);
      const double mk = B.mass[bk];
      const double dj = S.densities[j], pj = S.pressures[j];  // as written: indexed by list position
      const double pij = -mk * (pi / (di * di + S.eps) + pj / (dj * dj + S.eps));
      const double q = h * h - r * r;
      f3 gradW = vscale(945.0 / (32.0 * 3.141592653589793 * S.h9) * (q * q), rVec);  // gradw(), :173-177
      gradW = vscale(pij, gradW);
      aPressure = vadd(aPressure, gradW);
      f3 u = vsub(ld3(B.vel[bk]), pv);
      u = vscale((1.0 / (0.0001 + di * dj)) * S.viscosity * mk, u);
      const double nabla = (945.0 / (32.0 * 3.141592653589793 * S.h9)) * (h * h - r * r) * (7 * r * r - 3 * h * h);  // nablaw(), :180-184
      u = vscale(nabla, u);
      aVisc = vadd(aVisc, u);
      j++;
    }
    const double m = B.mass[bi];
    aVisc = vscale(m, aVisc);
    aPressure = vscale(m, aPressure);
    f3 f = ld3(B.force[bi]);
    f = vadd(f, aVisc);
    f = vadd(f, aPressure);
    B.force[bi] = st3(f);
  }
}
