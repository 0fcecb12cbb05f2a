// k_gs_world_exact.cuh — COLORED (exact) solve of a batch of small independent worlds: one warp per world, the world's
// rows resident in shared memory for the whole solve.
//
// GSSolver's arithmetic (gs_solver.dart:76-108: f64 on f32-stored operands, no FMA, every Vector3 store rounds to float)
// over the colour order of include/cannon_cuda.h, one independent solve per world with its own tolerance exit
// (gs_solver.dart:105) - what separate World objects would do (lib/world/world_class.dart:392-431).
//   * The execution order is grouped by (world, colour) (k_world_count / k_world_fill), so a world's units - and, rows being
//     built in execution order, its rows - are contiguous colour by colour.
//   * A block is one warp and owns one world. At the start it packs the world's rows into shared memory once (96-byte GxRow
//     records + the multipliers, the unit records and the bodies' delta vectors): ten iterations then run without touching
//     global memory. 57 KB per block lets four worlds share an SM (592 worlds in flight on a B200: one shard of config 4 at
//     8 GPUs). Rows beyond the shared capacity (a world's last, sparsest colours) and oversize unit / body tables stay in
//     global memory - slower, same result.
//   * Lanes take the units of a colour (independent by construction); one __syncwarp() separates colours: no grid barrier,
//     no atomics, no dependence on the other worlds of the batch.
// Measured on config 4 (4096 worlds x 64 bodies, 1.93e6 rows, 13 colours; profiles/README.md): the grid-wide level sweep
// k_gs needs 3.36 ms per solve (130 grid barriers) and 1.39 ms for a 512-world shard; eight lanes per world with the rows
// read from global memory every iteration 3.57 / 1.50 ms (a lone warp per SM waiting for L2); this kernel 2.73 / 0.46 ms.
// A world's solve is one warp's dependent chain (~0.4 ms for 10 iterations of 13 colours, conversion-pipe latency bound);
// the full batch is 4096 / 592 = 7 rounds of that - shared memory per world, not arithmetic, caps the concurrency.
// Tried and rejected: two lanes per unit (one per body, an f64 shuffle pair per row; tools/ubench/row_step.cu promises 348
// instead of 545 clk per row for a lone warp): bit-exact, but 0.50 / 2.91 ms here - the shuffles and the second lane's
// shared-memory reads cost more than the halved conversions save.
#pragma once
#include "k_gs_exact.cuh"

#define GWX_SMEM_BYTES 57088            // dynamic shared memory per block: four blocks per SM
#define GWX_MAXB 80                     // bodies whose (vlambda, wlambda) pairs live in shared memory
#define GWX_MAXU 112                    // unit records in shared memory
#define GWX_ROW_BYTES 104               // GxRow + multiplier
#define GWX_MAXR (((GWX_SMEM_BYTES - GWX_MAXB * 32 - GWX_MAXU * 40) / GWX_ROW_BYTES) & ~1)  // even: the tables behind the rows stay 16-byte aligned

struct __align__(8) GwxUnit { int bi, bj, fl, r0, r1, pad; double imA, imB; };

// a row of the unpacked arrays as the packed record of k_gs_exact (same bound codes as finish_row)
__device__ __forceinline__ void gwx_pack(const RowArrays& R, int r, GxRow& q) {
  const float4 n = R.n[r], rA = R.rA[r], rB = R.rB[r], iA = R.iA[r], iB = R.iB[r];
  q.nx = n.x; q.ny = n.y; q.nz = n.z;
  q.rAx = rA.x; q.rAy = rA.y; q.rAz = rA.z; q.rBx = rB.x; q.rBy = rB.y; q.rBz = rB.z;
  q.iAx = iA.x; q.iAy = iA.y; q.iAz = iA.z; q.iBx = iB.x; q.iBy = iB.y; q.iBz = iB.z;
  q.B = R.B[r]; q.invC = R.invC[r]; q.eps = R.eps[r];
  const double minF = R.minF[r], maxF = R.maxF[r];
  int bc = GXB_GENERAL;
  q.bound = maxF;
  if (__double_as_longlong(minF) == 0LL) bc = GXB_POS;
  else if (__double_as_longlong(minF) == __double_as_longlong(-maxF)) bc = GXB_SYM;
  else if (__double_as_longlong(maxF) == 0LL) { bc = GXB_NEG; q.bound = -minF; }
  const int kind = R.kind[r];
  q.code = ((kind == ROW_ROT || kind == ROW_MOTOR) ? 1 : 0) | (bc << 2);
}
__device__ __forceinline__ void gwx_widen(const GxRow& q, double lam, GxJ& j) {
  j.Bv = q.B; j.invC = q.invC; j.eps = q.eps; j.lam = lam;
  j.nx = (double)q.nx; j.ny = (double)q.ny; j.nz = (double)q.nz;
  const bool rot = q.code & 1;
  j.sAx = rot ? 0.0 : -j.nx; j.sAy = rot ? 0.0 : -j.ny; j.sAz = rot ? 0.0 : -j.nz;
  j.rAx = (double)q.rAx; j.rAy = (double)q.rAy; j.rAz = (double)q.rAz;
  j.rBx = (double)q.rBx; j.rBy = (double)q.rBy; j.rBz = (double)q.rBz;
  j.iAx = (double)q.iAx; j.iAy = (double)q.iAy; j.iAz = (double)q.iAz;
  j.iBx = (double)q.iBx; j.iBy = (double)q.iBy; j.iBz = (double)q.iBz;
  const int bc = q.code >> 2;
  j.general = bc == GXB_GENERAL;
  j.mn = bc == GXB_POS ? 0.0 : -q.bound;
  j.mx = bc == GXB_NEG ? 0.0 : q.bound;
}

__global__ void __launch_bounds__(32) k_gs_world_exact(RowArrays R, BodyArrays B, UnitArrays U, SolveParams P, GsStats G,
                                                       const int* __restrict__ binStart, const int* __restrict__ worldBody,
                                                       const int* __restrict__ nLevelsPtr) {
  extern __shared__ __align__(128) unsigned char s_dyn[];
  GxRow* const s_rows = (GxRow*)s_dyn;                                      // GWX_MAXR records
  double* const s_lam = (double*)(s_dyn + (size_t)GWX_MAXR * sizeof(GxRow));  // their multipliers
  float4* const s_vw = (float4*)(s_dyn + (size_t)GWX_MAXR * GWX_ROW_BYTES);
  GwxUnit* const s_units = (GwxUnit*)(s_dyn + (size_t)GWX_MAXR * GWX_ROW_BYTES + GWX_MAXB * 32);
  __shared__ int s_bs[GR_LV + 1];  // first unit of every colour of this world: read once, not once per colour and iteration
  const int lane = threadIdx.x, wd = blockIdx.x;
  const int nLevels = *nLevelsPtr;
  for (int i = lane; i <= GR_LV; i += 32) s_bs[i] = binStart[(size_t)wd * GR_LV + i];
  __syncwarp();
  const int* const bs = s_bs;
  const int a0 = bs[0], nU = bs[GR_LV] - a0;
  if (nU <= 0) { if (lane == 0) G.worldIters[wd] = 0; return; }
  const int b0 = worldBody[wd], nB = worldBody[wd + 1] - b0;
  const int r0w = U.eRowBase[a0], nR = U.eRowBase[a0 + nU] - r0w;
  const int nRs = min(nR, GWX_MAXR);     // rows [r0w, r0w + nRs) are staged, the rest is read from the global arrays
  const bool unitsShared = nU <= GWX_MAXU, bodiesShared = nB <= GWX_MAXB;
  float4* const vw = bodiesShared ? s_vw : (B.vlam + 2 * (size_t)b0);  // (vlambda, wlambda) pairs of the world's bodies
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = lane; i < 2 * nB; i += 32) vw[i] = z4;
  for (int i = lane; i < nRs; i += 32) { GxRow q; gwx_pack(R, r0w + i, q); s_rows[i] = q; s_lam[i] = 0.0; }  // k_rows_build left lambda at 0
  if (unitsShared)
    for (int i = lane; i < nU; i += 32) {
      GwxUnit m;
      const int a = a0 + i;
      m.bi = U.eBi[a]; m.bj = U.eBj[a]; m.fl = U.eFlags[a]; m.r0 = U.eRowBase[a]; m.r1 = U.eRowBase[a + 1]; m.pad = 0;
      m.imA = U.eImA[a]; m.imB = U.eImB[a];
      s_units[i] = m;
    }
  __syncwarp();
  int iter = 0;
  for (; iter != P.maxIter; iter++) {
    double local = 0.0;
    const int lvEnd = max(nLevels, 1);
    for (int lvl = 0; lvl < lvEnd; lvl++) {
      const int bin = min(lvl, GR_LV - 1);
      const int u0 = bs[bin], u1 = bs[bin + 1];
      for (int a = u0 + lane; a < u1; a += 32) {
        // colours beyond GR_LV share the last bin: this level's units only
        if (nLevels > GR_LV && bin == GR_LV - 1 && U.eLevel[a] != lvl) continue;
        GwxUnit m;
        if (unitsShared) m = s_units[a - a0];
        else { m.bi = U.eBi[a]; m.bj = U.eBj[a]; m.fl = U.eFlags[a]; m.r0 = U.eRowBase[a]; m.r1 = U.eRowBase[a + 1]; m.imA = U.eImA[a]; m.imB = U.eImB[a]; }
        if (m.r0 == m.r1) continue;
        const int ia = 2 * (m.bi - b0), ib = 2 * (m.bj - b0);
        // a body that is not movable keeps vlambda = wlambda = 0: never read or written (it may belong to nobody's table)
        f3 vA = ld3((m.fl & 1) ? vw[ia] : z4), wA = ld3((m.fl & 1) ? vw[ia + 1] : z4);
        f3 vB = ld3((m.fl & 2) ? vw[ib] : z4), wB = ld3((m.fl & 2) ? vw[ib + 1] : z4);
        double acc = 0.0;
        if (m.r1 - r0w <= nRs) {
          // every row of the unit is staged (the common case): software-pipelined like k_gs_exact - the next row's record is
          // read and widened in the shadow of this row's dependent chain; one basic block, no movable-body branches (an
          // immovable body has invMassSolve = 0 and I^-1 r = 0: its never-stored deltas stay +0 by arithmetic)
          const int k0 = m.r0 - r0w, nr = m.r1 - m.r0;
          GxJ cur, nxt;
          gwx_widen(s_rows[k0], s_lam[k0], cur);
          for (int rr = 0; rr < nr; rr++) {
            const int kn = k0 + min(rr + 1, nr - 1);
            gwx_widen(s_rows[kn], s_lam[kn], nxt);
            const double gwl = (gx_dot(vA, cur.sAx, cur.sAy, cur.sAz) + gx_dot(wA, cur.rAx, cur.rAy, cur.rAz)) +
                               (gx_dot(vB, cur.nx, cur.ny, cur.nz) + gx_dot(wB, cur.rBx, cur.rBy, cur.rBz));
            double dl = cur.invC * (cur.Bv - gwl - cur.eps * cur.lam);
            double mn = cur.mn, mx = cur.mx;
            if (cur.general) { mn = R.minF[m.r0 + rr]; mx = R.maxF[m.r0 + rr]; }
            if (cur.lam + dl < mn) dl = mn - cur.lam;
            else if (cur.lam + dl > mx) dl = mx - cur.lam;
            s_lam[k0 + rr] = cur.lam + dl;
            vA = gx_axpy(vA, m.imA * dl, cur.sAx, cur.sAy, cur.sAz); wA = gx_axpy(wA, dl, cur.iAx, cur.iAy, cur.iAz);
            vB = gx_axpy(vB, m.imB * dl, cur.nx, cur.ny, cur.nz); wB = gx_axpy(wB, dl, cur.iBx, cur.iBy, cur.iBz);
            acc += dl > 0.0 ? dl : -dl;
            cur = nxt;
          }
        } else
        for (int r = m.r0; r < m.r1; r++) {
          const int k = r - r0w;
          GxJ j;
          if (k < nRs) gwx_widen(s_rows[k], s_lam[k], j);
          else { GxRow q; gwx_pack(R, r, q); gwx_widen(q, R.lambda[r], j); }
          // one projected Gauss-Seidel row update (gs_solver.dart:88-102, equation_class.dart:95-105,151-169)
          const double gwl = (gx_dot(vA, j.sAx, j.sAy, j.sAz) + gx_dot(wA, j.rAx, j.rAy, j.rAz)) +
                             (gx_dot(vB, j.nx, j.ny, j.nz) + gx_dot(wB, j.rBx, j.rBy, j.rBz));
          double dl = j.invC * (j.Bv - gwl - j.eps * j.lam);
          double mn = j.mn, mx = j.mx;
          if (j.general) { mn = R.minF[r]; mx = R.maxF[r]; }
          if (j.lam + dl < mn) dl = mn - j.lam;
          else if (j.lam + dl > mx) dl = mx - j.lam;
          if (k < nRs) s_lam[k] = j.lam + dl; else R.lambda[r] = j.lam + dl;
          if (m.fl & 1) { vA = gx_axpy(vA, m.imA * dl, j.sAx, j.sAy, j.sAz); wA = gx_axpy(wA, dl, j.iAx, j.iAy, j.iAz); }
          if (m.fl & 2) { vB = gx_axpy(vB, m.imB * dl, j.nx, j.ny, j.nz); wB = gx_axpy(wB, dl, j.iBx, j.iBy, j.iBz); }
          acc += dl > 0.0 ? dl : -dl;
        }
        if (m.fl & 1) { vw[ia] = st3(vA); vw[ia + 1] = st3(wA); }
        if (m.fl & 2) { vw[ib] = st3(vB); vw[ib + 1] = st3(wB); }
        local += acc;
      }
      if (!bodiesShared) __threadfence_block();
      __syncwarp();
    }
    // tolerance test of this world (gs_solver.dart:99-107): the sum is order-insensitive for the comparison against tol^2
    double tot = local;
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (tot * tot < P.tol2) break;
  }
  if (bodiesShared) for (int i = lane; i < 2 * nB; i += 32) B.vlam[2 * (size_t)b0 + i] = vw[i];
  for (int i = lane; i < nRs; i += 32) R.lambda[r0w + i] = s_lam[i];
  if (lane == 0) { G.worldIters[wd] = iter; atomicMax(G.itersDone, iter); }
}
