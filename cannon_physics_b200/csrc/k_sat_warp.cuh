// k_sat_warp.cuh — warp-per-task separating-axis test + clipping for hull/hull and hull/heightfield-pillar tasks.
//
// The reference's convexConvex (lib/world/narrow_phase.dart:1981-2043) is a long sequential loop: up to
// |facesA| + |facesB| + |edgesA|*|edgesB| axis tests (48 for box-box, 85 for box-pillar, 120 for two 8-segment
// cylinders; lib/rigid_body_shapes/convex_polyhedron.dart:232-356), each projecting both hulls. Here one warp
// owns one task:
//   1. lanes rotate the face normals / unique edges of both hulls into shared memory once,
//   2. the axis list is dealt round-robin to the lanes; every lane keeps its own (depth, axis index) minimum,
//   3. a warp vote detects a separating axis, a lexicographic (depth, index) shuffle reduction reproduces the
//      sequential "first strictly smaller depth wins" rule bit for bit,
//   4. lane 0 clips the incident face (Sutherland-Hodgman, polygons in shared memory), lanes emit the contacts.
// Every arithmetic expression is the one the sequential path (k_narrowphase.cuh) evaluates, so contact counts
// and geometry stay bit-identical to the oracle; only the scheduling changed.
#pragma once
#include <cooperative_groups.h>

#include "k_narrowphase.cuh"
namespace cg = cooperative_groups;

#define SAT_MAXF 32
#define SAT_MAXE 32
#define SAT_GROUP 8  // lanes per task (a tile of the warp); 4 tasks share a warp
#define SAT_TILES 8            // tiles per CTA

struct SatScratch {
  f3 nA[SAT_MAXF], nB[SAT_MAXF];  // world face normals
  f3 eA[SAT_MAXE], eB[SAT_MAXE];  // world unique edges
  f3 pa[NP_MAXPOLY], pb[NP_MAXPOLY];
  double depth[NP_MAXPOLY];
  f3 cand[18];                    // pillar edge candidates
  int candMask[18];               // earlier candidates each one is almostEquals to
  PillarStore pil;
  int kept, overflow, closestA;
  f3 nrm;
};

// project with a precomputed local origin (ConvexPolyhedron.project, convex_polyhedron.dart:843-883)
__device__ __forceinline__ void hull_project_o(const HullView& H, const f3& axis, const q4& quat, const f3& localOrigin, double& mx, double& mn) {
  const f3 localAxis = qrot(qnegw(quat), axis);
  const double add = vdot(localOrigin, localAxis);
  mn = mx = vdot(ld3(H.v[0]), localAxis);
  for (int i = 1; i < H.nV; i++) {
    const double val = vdot(ld3(H.v[i]), localAxis);
    if (val > mx) mx = val;
    if (val < mn) mn = val;
  }
  mn -= add;
  mx -= add;
  if (mn > mx) { const double t = mn; mn = mx; mx = t; }
}

// clipFaceAgainstHull with the world normals taken from shared memory (same values as qrot(quat, n))
__device__ inline int clip_hulls_s(SatScratch& S, const HullView& HA, const f3& posA, const HullView& HB, const f3& posB, const q4& quatB,
                                   const f3& sep, bool& overflow) {
  int closestB = -1;
  double dmax = -INFINITY;
  for (int f = 0; f < HB.nF; f++) {
    const double d = vdot(S.nB[f], sep);
    if (d > dmax) { dmax = d; closestB = f; }
  }
  if (closestB < 0) return 0;
  f3* pa = S.pa;
  f3* pb = S.pb;
  int nIn = 0;
  {
    const int o = HB.fvOff[closestB], L = HB.fvOff[closestB + 1] - o;
    for (int i = 0; i < L && nIn < NP_MAXPOLY; i++) pa[nIn++] = vadd(posB, qrot(quatB, ld3(HB.v[HB.fvIdx[o + i]])));
    if (L > NP_MAXPOLY) overflow = true;
  }
  int closestA = -1;
  double dmin = INFINITY;
  for (int f = 0; f < HA.nF; f++) {
    const double d = vdot(S.nA[f], sep);
    if (d < dmin) { dmin = d; closestA = f; }
  }
  if (closestA < 0) return 0;
  const int numVerticesA = HA.fvOff[closestA + 1] - HA.fvOff[closestA];
  const int co = HA.fcOff[closestA], nConn = HA.fcOff[closestA + 1] - co;
  f3* in = pa;
  f3* out = pb;
  for (int i = 0; i < numVerticesA; i++) {
    const int otherFace = (nConn > i) ? HA.fcIdx[co + i] : 0;
    const f3 pn = S.nA[otherFace];
    const double pc = HA.pc[otherFace] - vdot(pn, posA);
    int nOut = 0;
    if (nIn >= 2) {
      f3 firstVertex = in[nIn - 1];
      double nDotFirst = vdot(pn, firstVertex) + pc;
      for (int vi = 0; vi < nIn; vi++) {
        const f3 lastVertex = in[vi];
        const double nDotLast = vdot(pn, lastVertex) + pc;
        if (nDotFirst < 0) {
          if (nOut < NP_MAXPOLY) out[nOut++] = (nDotLast < 0) ? lastVertex : vlerp(firstVertex, lastVertex, nDotFirst / (nDotFirst - nDotLast));
          else overflow = true;
        } else if (nDotLast < 0) {
          if (nOut + 1 < NP_MAXPOLY) {
            out[nOut++] = vlerp(firstVertex, lastVertex, nDotFirst / (nDotFirst - nDotLast));
            out[nOut++] = lastVertex;
          } else overflow = true;
        }
        firstVertex = lastVertex;
        nDotFirst = nDotLast;
      }
    }
    f3* t = in; in = out; out = t;
    nIn = nOut;
  }
  const f3 nrm = S.nA[closestA];
  S.nrm = nrm;
  const double planeEq = HA.pc[closestA] - vdot(nrm, posA);
  int kept = 0;
  for (int i = 0; i < nIn; i++) {
    double depth = vdot(nrm, in[i]) + planeEq;
    if (depth <= -100.0) depth = -100.0;
    if (depth <= 100.0 && depth <= 1e-6) {
      const f3 p = in[i];
      out[kept] = p;
      S.depth[kept] = depth;
      kept++;
    }
  }
  if (out != pa)
    for (int i = 0; i < kept; i++) pa[i] = out[i];
  return kept;
}

template <bool PILLAR>
__global__ void __launch_bounds__(SAT_TILES * SAT_GROUP, 8) k_np_hull_warp(BodyArrays B, ShapeTables T, NpArrays A, int* clipOverflow) {
  __shared__ SatScratch s_scr[SAT_TILES];
  cg::thread_block_tile<SAT_GROUP> tile = cg::tiled_partition<SAT_GROUP>(cg::this_thread_block());
  const int lane = tile.thread_rank(), tib = threadIdx.x / SAT_GROUP;
  SatScratch& S = s_scr[tib];
  const int TYPE = PILLAR ? NP_HPIL : NP_HH;
  const int nb = (*A.nTasks <= A.taskCap) ? A.bucketCount[TYPE] : 0;
  const int tilesPerGrid = gridDim.x * SAT_TILES;
  for (int u = blockIdx.x * SAT_TILES + tib; u < nb; u += tilesPerGrid) {
    tile.sync();
    TaskCtx c;
    load_task(B, T, A, A.bucket[A.bucketStart[TYPE] + u], c);
    RawOut o; o.A = A; o.task = c.task;
    const HullView HA = hull_view(T, c.si.hull);
    HullView HB;
    f3 xB;
    bool upper = false;
    if (PILLAR) {
      const HfDev hf = T.hfs[c.sj.hf];
      const int2 cell = A.taskCell[c.task];
      upper = (c.info >> 4) & 1;
      // Heightfield.getConvexTrianglePillar (heightfield.dart:330-487) spread over the tile: vertices by lane 0,
      // one face normal per lane, one edge candidate per lane, duplicate removal in candidate order
      // (computeNormals / computeEdges, convex_polyhedron.dart:110-185)
      if (lane == 0) {
        f3 off, pv[6];
        double bsr;
        pillar_bounds(T, hf, cell.x, cell.y, upper, off, bsr, pv);
        for (int i = 0; i < 6; i++) S.pil.v[i] = st3(pv[i]);
        S.pil.bsr = bsr;
        S.nrm = off;
      }
      tile.sync();
      const f3 off = S.nrm;
      const int* fv = upper ? c_pillarUpper : c_pillarLower;
      for (int f = lane; f < 5; f += SAT_GROUP) {
        const int o0 = c_pillarFvOff[f];
        const f3 va = ld3(S.pil.v[fv[o0]]), vb = ld3(S.pil.v[fv[o0 + 1]]), vc = ld3(S.pil.v[fv[o0 + 2]]);
        f3 nn = vcross(vsub(vc, vb), vsub(vb, va));
        if (!(nn.x == 0.f && nn.y == 0.f && nn.z == 0.f)) vnormalize(nn);
        nn = vneg(nn);
        S.pil.n[f] = st3(nn);
        S.pil.pc[f] = -vdot(nn, va);
      }
      for (int e = lane; e < 18; e += SAT_GROUP) {
        int f = 0;
        while (e >= c_pillarFvOff[f + 1]) f++;
        const int o0 = c_pillarFvOff[f], L = c_pillarFvOff[f + 1] - o0, j = e - o0;
        f3 ev = vsub(ld3(S.pil.v[fv[o0 + j]]), ld3(S.pil.v[fv[o0 + (j + 1) % L]]));
        vnormalize(ev);
        S.cand[e] = ev;
      }
      tile.sync();
      for (int e = lane; e < 18; e += SAT_GROUP) {
        int m = 0;
        const f3 ev = S.cand[e];
        for (int p = 0; p < e; p++)
          if (valmost_eq(S.cand[p], ev)) m |= 1 << p;
        S.candMask[e] = m;
      }
      tile.sync();
      if (lane == 0) {
        int nE = 0, keep = 0;
        for (int k = 0; k < 18; k++)
          if ((S.candMask[k] & keep) == 0) { keep |= 1 << k; S.pil.e[nE++] = st3(S.cand[k]); }
        S.pil.nE = nE;
      }
      tile.sync();
      xB = to_world_point(c.xj, c.qj, off);
      HB = pillar_view(S.pil, upper);
    } else {
      HB = hull_view(T, c.sj.hull);
      xB = c.xj;
    }
    int kept = 0;
    f3 sep; sep.x = sep.y = sep.z = 0.f;
    bool candidate = PILLAR ? (vdist(c.xi, xB) < HB.bsr + HA.bsr) : true;
    if (candidate && (vdist(c.xi, xB) > HA.bsr + HB.bsr)) candidate = false;  // convexConvex's own bounding test (:1999)
    if (candidate && (HA.nF > SAT_MAXF || HB.nF > SAT_MAXF || HA.nE > SAT_MAXE || HB.nE > SAT_MAXE)) {
      continue;  // oversized hulls are left to the sequential kernels (k_np_hull_hull / k_np_hull_pillar, oversizeOnly)
    }
    if (candidate) {
      for (int i = lane; i < HA.nF; i += SAT_GROUP) S.nA[i] = qrot(c.qi, ld3(HA.n[i]));
      for (int i = lane; i < HB.nF; i += SAT_GROUP) S.nB[i] = qrot(c.qj, ld3(HB.n[i]));
      for (int i = lane; i < HA.nE; i += SAT_GROUP) S.eA[i] = qrot(c.qi, ld3(HA.e[i]));
      for (int i = lane; i < HB.nE; i += SAT_GROUP) S.eB[i] = qrot(c.qj, ld3(HB.e[i]));
      tile.sync();
      f3 zero; zero.x = zero.y = zero.z = 0.f;
      const f3 oA = to_local_point(c.xi, c.qi, zero), oB = to_local_point(xB, c.qj, zero);
      const int nfa = HA.hasAxes ? (PILLAR ? 1 : HA.nF) : 0;  // heightfieldConvex passes faceListA = [0] (:2062)
      const int nfb = HB.hasAxes ? HB.nF : 0;
      const int nAxes = nfa + nfb + HA.nE * HB.nE;
      double best = INFINITY;
      int bestIdx = 0x7fffffff;
      f3 bestAxis = zero;
      bool separated = false;
      // rounds of SAT_GROUP axes with a vote after each round: a separating axis ends the task early (same result
      // as the sequential `return false`, convex_polyhedron.dart:264-267)
      for (int base = 0; base < nAxes && !separated; base += SAT_GROUP) {
        const int t = base + lane;
        bool sepHere = false;
        if (t < nAxes) {
          f3 axis;
          bool valid = true;
          if (t < nfa) axis = S.nA[t];
          else if (t < nfa + nfb) axis = S.nB[t - nfa];
          else {
            const int e = t - nfa - nfb;
            axis = vcross(S.eA[e / HB.nE], S.eB[e % HB.nE]);
            if (valmost_zero(axis)) valid = false;
            else vnormalize(axis);
          }
          if (valid) {
            double maxA, minA, maxB, minB;
            hull_project_o(HA, axis, c.qi, oA, maxA, minA);
            hull_project_o(HB, axis, c.qj, oB, maxB, minB);
            if (maxA < minB || maxB < minA) sepHere = true;
            else {
              const double d0 = maxA - minB, d1 = maxB - minA;
              const double d = d0 < d1 ? d0 : d1;
              if (d < best) { best = d; bestIdx = t; bestAxis = axis; }
            }
          }
        }
        separated = tile.any(sepHere);
      }
      if (!separated) {
        // lexicographic (depth, index) minimum == the sequential loop's "first strictly smaller depth"
        double rb = best;
        int ri = bestIdx;
        for (int off = SAT_GROUP / 2; off > 0; off >>= 1) {
          const double ob = tile.shfl_xor(rb, off);
          const int oi = tile.shfl_xor(ri, off);
          if (ob < rb || (ob == rb && oi < ri)) { rb = ob; ri = oi; }
        }
        const unsigned winMask = tile.ballot(bestIdx == ri && ri != 0x7fffffff);
        if (winMask) {
          const int src = __ffs(winMask) - 1;
          sep.x = tile.shfl(bestAxis.x, src);
          sep.y = tile.shfl(bestAxis.y, src);
          sep.z = tile.shfl(bestAxis.z, src);
        }
        const f3 deltaC = vsub(xB, c.xi);
        if (vdot(deltaC, sep) > 0.0) sep = vneg(sep);
        if (lane == 0) {
          bool ovf = false;
          S.kept = clip_hulls_s(S, HA, c.xi, HB, xB, c.qj, sep, ovf);
          if (ovf) atomicExch(clipOverflow, 1);
        }
        tile.sync();
        kept = S.kept;
      }
    }
    // emission: lane 0 reserves the block in the raw pool, lanes write one contact each
    tile.sync();
    if (lane == 0) {
      raw_alloc(o, kept);
      S.kept = o.A.taskCnt[o.task];  // 0 if the pool overflowed
      S.closestA = o.start;
    }
    tile.sync();
    kept = S.kept;
    const int start = S.closestA;
    const f3 ni = vneg(sep);
    const f3 nrm = S.nrm;
    for (int j = lane; j < kept; j += SAT_GROUP) {
      const f3 q = vscale(S.depth[j], vneg(nrm));
      f3 ri = vadd(S.pa[j], q);
      f3 rj = S.pa[j];
      ri = vsub(ri, c.xi);
      rj = vsub(rj, xB);
      ri = vsub(vadd(ri, c.xi), c.xi);
      rj = vsub(vadd(rj, xB), c.xj);
      A.rawRi[start + j] = st3(ri);
      A.rawRj[start + j] = st3(rj);
      A.rawNi[start + j] = st3(ni);
    }
  }
}
