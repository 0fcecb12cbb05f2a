// k_sat_warp.cuh — tile-per-task separating-axis test + clipping for hull/hull and hull/heightfield-pillar tasks.
//
// The reference's convexConvex (lib/world/narrow_phase.dart:1981-2043) is a long sequential loop: up to
// |facesA| + |facesB| + |edgesA|*|edgesB| axis tests (48 for box-box, 85 for box-pillar, 141 for cylinder-pillar;
// lib/rigid_body_shapes/convex_polyhedron.dart:232-356), each projecting both hulls. Here an 8-lane tile owns one
// task (four tasks per warp) and the work is two launches per task type (k_np_hull_warp<PILLAR, PHASE>):
//   PHASE 0  lanes rotate the face normals / pruned unique edges of both hulls into shared memory and widen the
//            vertices to f64 once; the pruned axis list (+- copies removed, see below) is taken in rounds of 8 with a
//            vote after each round (a separating axis closes the task); a lexicographic (depth, index) shuffle
//            reduction reproduces the sequential "first strictly smaller depth wins" rule bit for bit; surviving
//            tasks are queued with their axis,
//   PHASE 1  the queued tasks are clipped by the whole tile (Sutherland-Hodgman with a prefix sum per pass, polygons
//            in shared memory) and the lanes emit the contacts.
// Every arithmetic expression is the one the sequential path (k_narrowphase.cuh) evaluates, so contact counts
// and geometry stay bit-identical to the oracle; only the scheduling changed.
#pragma once
#include <cooperative_groups.h>

#include "k_narrowphase.cuh"
namespace cg = cooperative_groups;

#define SAT_GROUP 8  // lanes per task (a tile of the warp); 4 tasks share a warp
#define SAT_TILES 8            // tiles per CTA
#ifndef SAT_PILLAR_CTAS
#define SAT_PILLAR_CTAS 8      // resident CTAs per SM the hull / pillar kernels are compiled for. Measured on the settled 100k
                               // pile: 8 (96 registers) narrowphase 1.288 ms, 10 -> 1.317 ms, 12 (80 registers) -> 1.330 ms: more resident
                               // warps only queue on the FP64 pipe, the extra register pressure costs more than the latency hiding gains
#endif

// The B side of a hull / pillar task is a heightfield pillar (5 faces, <= 7 pruned edges, 6 vertices): its tables are
// sized for that, which takes a CTA's scratch from 23 KB to 18 KB (shared memory no longer bounds the occupancy; see
// SAT_PILLAR_CTAS for why it is not raised).
template <int NFB, int NEB, int NVB>
struct SatScratchT {
  f3 nA[SAT_MAXF], nB[NFB];  // world face normals
  f3 eA[SAT_MAXE], eB[NEB];  // world unique edges
  union {                          // three phases of a task that never overlap
    struct { double vA[SAT_MAXV][3], vB[NVB][3]; } v;                                   // axis loop: local vertices, widened once
    struct { f3 pa[NP_MAXPOLY], pb[NP_MAXPOLY]; double depth[NP_MAXPOLY]; } c;          // clipping + emission
  } u;
  int kept, overflow, closestA;
  f3 nrm;
};

// Axis pruning that cannot change the result. If a world axis is numerically +-equal to an earlier axis of the same
// loop, every expression downstream (Quaternion.vmult, cross2, normalize, project, testSepAxis; quaternion.dart:21-44,
// convex_polyhedron.dart:360-384,843-883) is odd in the axis, IEEE rounding is symmetric under negation, and testSepAxis
// is symmetric in (max, -min): the later axis is separated iff the earlier one is and has exactly the same depth, so
// with findSeparatingAxis' strict `d < dmin` (convex_polyhedron.dart:266,302,328) it can never be selected. The same
// holds for the edge loop when one of the two edges is a +-copy. A box therefore contributes 3 face axes and 3 edges
// instead of 6 and 6, a heightfield pillar 7 edges instead of 14: box-box 48 -> 15 axes, box-pillar 90 -> 27.
// The pruned lists are built once: per hull on the host (cannon_world_set_shapes: +-equal in the local frame implies
// +-equal after Quaternion.vmult), per heightfield pillar by k_pillars_build.

// ConvexPolyhedron.project (convex_polyhedron.dart:843-883) with a precomputed local origin and the hull's local
// vertices already widened to double in shared memory. float -> double is exact and max / min do not depend on the
// order of the candidates, so two vertices are evaluated per trip for instruction-level parallelism.
__device__ __forceinline__ void hull_project_w(const double (*v)[3], int nV, const f3& axis, const q4& quat, const f3& localOrigin, double& mx,
                                               double& mn) {
  const f3 localAxis = qrot(qnegw(quat), axis);
  const double ax = W(localAxis.x), ay = W(localAxis.y), az = W(localAxis.z);
  const double add = vdot(localOrigin, localAxis);
  {
    double s = v[0][0] * ax;
    s += v[0][1] * ay;
    s += v[0][2] * az;
    mn = mx = s;
  }
  int i = 1;
  for (; i + 1 < nV; i += 2) {
    double s0 = v[i][0] * ax, s1 = v[i + 1][0] * ax;
    s0 += v[i][1] * ay; s1 += v[i + 1][1] * ay;
    s0 += v[i][2] * az; s1 += v[i + 1][2] * az;
    mx = fmax(mx, fmax(s0, s1));
    mn = fmin(mn, fmin(s0, s1));
  }
  if (i < nV) {
    double s = v[i][0] * ax;
    s += v[i][1] * ay;
    s += v[i][2] * az;
    mx = fmax(mx, s);
    mn = fmin(mn, s);
  }
  mn -= add;
  mx -= add;
  if (mn > mx) { const double t = mn; mn = mx; mx = t; }
}

// clipAgainstHull / clipFaceAgainstHull / clipFaceAgainstPlane (convex_polyhedron.dart:189-227,417-587) by the whole
// tile: the face searches are lexicographic reductions (first extremum wins, like the sequential scans), each
// Sutherland-Hodgman pass gives one polygon edge to a lane and places its 0 / 1 / 2 output vertices with a prefix sum,
// so the output polygon has the sequential order and every vertex is produced by the sequential expression.
template <class Tile, class SatScratch>
__device__ inline int clip_hulls_tile(const Tile& tile, int lane, SatScratch& S, const HullView& HA, const f3& posA, const HullView& HB,
                                      const f3& posB, const q4& quatB, const f3& sep, bool& overflow) {
  f3* pa = S.u.c.pa;
  f3* pb = S.u.c.pb;
  double* nd = S.u.c.depth;
  double bd = -INFINITY;
  int closestB = 0x7fffffff;
  for (int f = lane; f < HB.nF; f += SAT_GROUP) {
    const double d = vdot(S.nB[f], sep);
    if (d > bd) { bd = d; closestB = f; }
  }
  double ad = INFINITY;
  int closestA = 0x7fffffff;
  for (int f = lane; f < HA.nF; f += SAT_GROUP) {
    const double d = vdot(S.nA[f], sep);
    if (d < ad) { ad = d; closestA = f; }
  }
  for (int off = SAT_GROUP / 2; off > 0; off >>= 1) {
    const double ob = tile.shfl_xor(bd, off), oa = tile.shfl_xor(ad, off);
    const int ib = tile.shfl_xor(closestB, off), ia = tile.shfl_xor(closestA, off);
    if (ib != 0x7fffffff && (closestB == 0x7fffffff || ob > bd || (ob == bd && ib < closestB))) { bd = ob; closestB = ib; }
    if (ia != 0x7fffffff && (closestA == 0x7fffffff || oa < ad || (oa == ad && ia < closestA))) { ad = oa; closestA = ia; }
  }
  if (closestB == 0x7fffffff || closestA == 0x7fffffff) return 0;
  int nIn;
  {
    const int o = HB.fvOff[closestB], L = HB.fvOff[closestB + 1] - o;
    nIn = min(L, NP_MAXPOLY);
    for (int i = lane; i < nIn; i += SAT_GROUP) pa[i] = vadd(posB, qrot(quatB, ld3(HB.v[HB.fvIdx[o + i]])));
    if (L > NP_MAXPOLY) overflow = true;
  }
  const int numVerticesA = HA.fvOff[closestA + 1] - HA.fvOff[closestA];
  const int co = HA.fcOff[closestA], nConn = HA.fcOff[closestA + 1] - co;
  f3* in = pa;
  f3* out = pb;
  tile.sync();
  for (int i = 0; i < numVerticesA; i++) {
    const int otherFace = (nConn > i) ? HA.fcIdx[co + i] : 0;
    const f3 pn = S.nA[otherFace];
    const double pc = HA.pc[otherFace] - vdot(pn, posA);
    int nOut = 0;
    if (nIn >= 2) {
      for (int v = lane; v < nIn; v += SAT_GROUP) nd[v] = vdot(pn, in[v]) + pc;
      tile.sync();
      for (int base = 0; base < nIn; base += SAT_GROUP) {
        const int vi = base + lane;
        int cnt = 0;
        double nDotFirst = 0.0, nDotLast = 0.0;
        if (vi < nIn) {
          nDotFirst = nd[vi == 0 ? nIn - 1 : vi - 1];
          nDotLast = nd[vi];
          if (nDotFirst < 0) cnt = 1;
          else if (nDotLast < 0) cnt = 2;
        }
        int incl = cnt;
        for (int o = 1; o < SAT_GROUP; o <<= 1) {
          const int t = tile.shfl_up(incl, o);
          if (lane >= o) incl += t;
        }
        const int pos = nOut + incl - cnt;
        if (cnt == 1) {
          if (pos < NP_MAXPOLY) {
            const f3 lastVertex = in[vi];
            out[pos] = (nDotLast < 0) ? lastVertex : vlerp(in[vi == 0 ? nIn - 1 : vi - 1], lastVertex, nDotFirst / (nDotFirst - nDotLast));
          } else overflow = true;
        } else if (cnt == 2) {
          if (pos + 1 < NP_MAXPOLY) {
            const f3 lastVertex = in[vi];
            out[pos] = vlerp(in[vi == 0 ? nIn - 1 : vi - 1], lastVertex, nDotFirst / (nDotFirst - nDotLast));
            out[pos + 1] = lastVertex;
          } else overflow = true;
        }
        nOut += tile.shfl(incl, SAT_GROUP - 1);
      }
      nOut = min(nOut, NP_MAXPOLY);
    }
    tile.sync();
    f3* t = in; in = out; out = t;
    nIn = nOut;
  }
  const f3 nrm = S.nA[closestA];
  if (lane == 0) S.nrm = nrm;
  const double planeEq = HA.pc[closestA] - vdot(nrm, posA);
  // keep the points at or below the reference face (depth <= 1e-6, clamped at -100), in polygon order
  int kept = 0;
  for (int base = 0; base < nIn; base += SAT_GROUP) {
    const int i = base + lane;
    bool keep = false;
    double depth = 0.0;
    f3 p; p.x = p.y = p.z = 0.f;
    if (i < nIn) {
      p = in[i];
      depth = vdot(nrm, p) + planeEq;
      if (depth <= -100.0) depth = -100.0;
      keep = depth <= 100.0 && depth <= 1e-6;
    }
    const unsigned m = tile.ballot(keep);
    tile.sync();  // every lane has read in[] of this chunk before out[] (which may alias earlier chunks only) is written
    if (keep) {
      const int pos = kept + __popc(m & ((1u << lane) - 1u));
      out[pos] = p;
      nd[pos] = depth;
    }
    kept += __popc(m);
  }
  tile.sync();
  if (out != pa) {
    for (int i = lane; i < kept; i += SAT_GROUP) pa[i] = out[i];
    tile.sync();
  }
  return kept;
}

// Two launches per task type. PHASE 0 runs the separating-axis test of every task and either closes the task (no contact)
// or queues it with its separating axis; PHASE 1 clips and emits the queued tasks. One fused kernel was ~53 KB of code
// walked by four divergent tiles per warp - far beyond the 32 KB L1.5 instruction cache (`no_instruction` was the second
// largest stall); each phase fits, and phase 1 only sees tasks that all take the same path.
template <bool PILLAR, int PHASE>
__global__ void __launch_bounds__(SAT_TILES * SAT_GROUP, PILLAR ? SAT_PILLAR_CTAS : 8) k_np_hull_warp(BodyArrays B, ShapeTables T, NpArrays A, int* clipOverflow) {
  typedef SatScratchT<PILLAR ? 8 : SAT_MAXF, PILLAR ? 8 : SAT_MAXE, PILLAR ? 8 : SAT_MAXV> SatScratch;
  __shared__ SatScratch s_scr[SAT_TILES];
  cg::thread_block_tile<SAT_GROUP> tile = cg::tiled_partition<SAT_GROUP>(cg::this_thread_block());
  const int lane = tile.thread_rank(), tib = threadIdx.x / SAT_GROUP;
  SatScratch& S = s_scr[tib];
  const int TYPE = PILLAR ? NP_HPIL : NP_HH;
  const int* const clipList = A.clipList + (size_t)(PILLAR ? 1 : 0) * A.taskCap;
  int* const nClip = A.nClip + (PILLAR ? 1 : 0);
  const int nb = (*A.nTasks <= A.taskCap) ? (PHASE == 0 ? A.bucketCount[TYPE] : min(*nClip, A.taskCap)) : 0;
  const int tilesPerGrid = gridDim.x * SAT_TILES;
  for (int u = blockIdx.x * SAT_TILES + tib; u < nb; u += tilesPerGrid) {
    tile.sync();
    TaskCtx c;
    load_task(B, T, A, PHASE == 0 ? A.bucket[A.bucketStart[TYPE] + u] : clipList[u], c);
    RawOut o; o.A = A; o.task = c.task;
    const HullView HA = hull_view(T, c.si.hull);
    HullView HB;
    f3 xB;
    bool upper = false;
    if (PILLAR) {
      // the pillar (vertices, face normals, plane constants, pruned unique edges) was built when the shapes were set
      const HfDev hf = T.hfs[c.sj.hf];
      const int2 cell = A.taskCell[c.task];
      upper = (c.info >> 4) & 1;
      const PillarRec* R = pillar_rec(T, hf, cell.x, cell.y, upper);
      xB = to_world_point(c.xj, c.qj, ld3(R->off));
      HB = pillar_view_rec(R, upper);
    } else {
      HB = hull_view(T, c.sj.hull);
      xB = c.xj;
    }
    int kept = 0;
    f3 sep; sep.x = sep.y = sep.z = 0.f;
    if (PHASE == 1) {
      // a queued task: its separating axis is known, clipping needs the world face normals of both hulls
      sep = ld3(A.taskSep[c.task]);
      for (int i = lane; i < HA.nF; i += SAT_GROUP) S.nA[i] = qrot(c.qi, ld3(HA.n[i]));
      for (int i = lane; i < HB.nF; i += SAT_GROUP) S.nB[i] = qrot(c.qj, ld3(HB.n[i]));
      tile.sync();
      if (A.debug != 1) {
        bool ovf = false;
        kept = clip_hulls_tile(tile, lane, S, HA, c.xi, HB, xB, c.qj, sep, ovf);
        if (ovf) atomicExch(clipOverflow, 1);
      }
    } else {
    bool needClip = false;
    bool candidate = PILLAR ? (vdist(c.xi, xB) < HB.bsr + HA.bsr) : true;
    if (candidate && (vdist(c.xi, xB) > HA.bsr + HB.bsr)) candidate = false;  // convexConvex's own bounding test (:1999)
    if (candidate && sat_oversize(HA, HB)) {
      continue;  // oversized hulls are left to the sequential kernels (k_np_hull_hull / k_np_hull_pillar, oversizeOnly)
    }
    if (A.debug == 3) candidate = false;
    if (candidate) {
      for (int i = lane; i < HA.nF; i += SAT_GROUP) S.nA[i] = qrot(c.qi, ld3(HA.n[i]));
      for (int i = lane; i < HB.nF; i += SAT_GROUP) S.nB[i] = qrot(c.qj, ld3(HB.n[i]));
      for (int i = lane; i < HA.nEk; i += SAT_GROUP) S.eA[i] = qrot(c.qi, ld3(HA.ek[i]));
      for (int i = lane; i < HB.nEk; i += SAT_GROUP) S.eB[i] = qrot(c.qj, ld3(HB.ek[i]));
      for (int i = lane; i < HA.nV; i += SAT_GROUP) { const float4 v = HA.v[i]; S.u.v.vA[i][0] = v.x; S.u.v.vA[i][1] = v.y; S.u.v.vA[i][2] = v.z; }
      for (int i = lane; i < HB.nV; i += SAT_GROUP) { const float4 v = HB.v[i]; S.u.v.vB[i][0] = v.x; S.u.v.vB[i][1] = v.y; S.u.v.vB[i][2] = v.z; }
      tile.sync();
      const int nEA = HA.nEk, nEB = HB.nEk;
      const int nFA = PILLAR ? 1 : HA.nFk, nFB = HB.nFk;
      f3 zero; zero.x = zero.y = zero.z = 0.f;
      const f3 oA = to_local_point(c.xi, c.qi, zero), oB = to_local_point(xB, c.qj, zero);
      const int nfa = HA.hasAxes ? nFA : 0;  // heightfieldConvex passes faceListA = [0] (:2062)
      const int nfb = HB.hasAxes ? nFB : 0;
      const int nAxes = nfa + nfb + nEA * nEB;
      double best = INFINITY;
      int bestIdx = 0x7fffffff;
      f3 bestAxis = zero;
      bool separated = A.debug == 2;
      // rounds of SAT_GROUP axes with a vote after each round: a separating axis ends the task early (same result
      // as the sequential `return false`, convex_polyhedron.dart:264-267)
      for (int base = 0; base < nAxes && !separated; base += SAT_GROUP) {
        const int t = base + lane;
        bool sepHere = false;
        if (t < nAxes) {
          f3 axis;
          bool valid = true;
          if (t < nfa) axis = S.nA[PILLAR ? 0 : HA.fk[t]];
          else if (t < nfa + nfb) axis = S.nB[HB.fk[t - nfa]];
          else {
            const int e = t - nfa - nfb;
            axis = vcross(S.eA[e / nEB], S.eB[e % nEB]);
            if (valmost_zero(axis)) valid = false;
            else vnormalize(axis);
          }
          if (valid) {
            double maxA, minA, maxB, minB;
            hull_project_w(S.u.v.vA, HA.nV, axis, c.qi, oA, maxA, minA);
            hull_project_w(S.u.v.vB, HB.nV, axis, c.qj, oB, maxB, minB);
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
        needClip = true;
      }
    }
    if (lane == 0) {
      if (needClip) {  // queue for the clipping launch
        const int q = atomicAdd(nClip, 1);
        if (q < A.taskCap) A.clipList[(size_t)(PILLAR ? 1 : 0) * A.taskCap + q] = c.task;
        A.taskSep[c.task] = st3(sep);
      } else {
        raw_alloc(o, 0);  // separated (or outside the bounding test): the task is closed without contacts
      }
    }
    continue;
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
      const f3 q = vscale(S.u.c.depth[j], vneg(nrm));
      f3 ri = vadd(S.u.c.pa[j], q);
      f3 rj = S.u.c.pa[j];
      ri = vsub(ri, c.xi);
      rj = vsub(rj, xB);
      ri = vsub(vadd(ri, c.xi), c.bxi);
      rj = vsub(vadd(rj, xB), c.bxj);
      A.rawRi[start + j] = st3(ri);
      A.rawRj[start + j] = st3(rj);
      A.rawNi[start + j] = st3(ni);
    }
  }
}
