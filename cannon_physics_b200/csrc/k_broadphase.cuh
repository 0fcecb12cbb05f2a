// k_broadphase.cuh — pair finding (K2).
//
// The reference's three broadphases define the pair SET and the pair ORDER:
//   Naive  naive_broadphase.dart:21-32   all i>j pairs passing needBroadphaseCollision + the bounding test,
//                                        emitted i-major, j ascending
//   SAP    sap_broadphase.dart:138-189   stable sort of the persistent axis list by aabb.lowerBound[axis],
//                                        forward sweep that `continue`s on filtered pairs and `break`s on
//                                        pos[axis] +- boundingRadius
//   Grid   grid_broadphase.dart:59-239   (intended semantics) Naive set restricted to bodies sharing a bin
// The device path reproduces set and order exactly but never does O(N^2) work for Naive/Grid: bodies with a
// finite bounding radius are hashed into a uniform grid (cell >= 2*r_max), the rest ("big": planes,
// heightfields) are tested against everything. Each body counts its partners j<i, an exclusive scan assigns
// output ranges, and a second pass writes them sorted by j — a canonical order independent of scheduling.
#pragma once
#include "world.cuh"

#define BP_MAXNB 160  // per-thread partner buffer; longer lists fall back to repeated selection

struct BpParams {
  int n;                // bodies
  int kind;             // CANNON_BP_*
  int useBoxes;
  int nWorlds;
  int hashMask;         // H-1
  double cell;          // uniform-grid cell edge
  int nBig;
  // GridBroadphase parameters
  int gnx, gny, gnz;
  double gxmin, gymin, gzmin, gxmult, gymult, gzmult, gbx, gby, gbz, gBinRadius;
  int sapAxis;
};

struct BpArrays {
  int* nbCache;         // [n * BP_CACHE] partners found by the counting pass of k_bp_small, indexed by bucket position
  // per body
  int4* cellc;          // (cx,cy,cz,world) of small bodies
  int* binLo;           // packed 10:10:10 GridBroadphase bin range
  int* binHi;
  uint32_t* skey;       // sort keys / values (hash bucket -> body)
  uint32_t* sval;
  int* cellStart;       // H+1 entries
  int* cellEnd;
  // data gathered into bucket order
  float4* spos;
  double* srad;
  int4* smeta;          // group, mask, (staticOrSleeping | world<<1), body index
  int4* scell;
  const int* bigList;   // indices of big bodies, ascending (batch: grouped by world)
  const int* bigWorldStart;  // nWorlds+1 (batch) or {0,nBig}
  const int* worldStart;     // nWorlds+1: first body index of each world (bodies of a world are contiguous)
  int* counts;          // partners per body (by body index)
  int* offs;            // exclusive scan of counts
  // SAP
  uint32_t* sapKey;
  uint32_t* sapList;    // persistent axisList (body indices in sorted order)
};

__device__ __forceinline__ uint32_t cell_hash(int cx, int cy, int cz, int w) {
  return ((uint32_t)cx * 73856093u) ^ ((uint32_t)cy * 19349663u) ^ ((uint32_t)cz * 83492791u) ^ ((uint32_t)w * 2654435761u);
}

__device__ __forceinline__ int grid_bin_lo(double v, double mn, double mult, int n) {
  double t = (v - mn) * mult;
  long long q;
  if (!(t > -1e15)) q = -1000000000000000LL;
  else if (!(t < 1e15)) q = 1000000000000000LL;
  else q = (long long)floor(t);
  return q < 0 ? 0 : (q >= n ? n - 1 : (int)q);
}
__device__ __forceinline__ int grid_bin_hi(double v, double mn, double mult, int n) {
  double t = (v - mn) * mult;
  long long q;
  if (!(t > -1e15)) q = -1000000000000000LL;
  else if (!(t < 1e15)) q = 1000000000000000LL;
  else q = (long long)ceil(t);
  return q < 0 ? 0 : (q >= n ? n - 1 : (int)q);
}

// per body: uniform-grid cell + hash key, GridBroadphase bin range
__global__ void __launch_bounds__(256) k_bp_cells(BodyArrays B, ShapeTables T, BpParams P, BpArrays A) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += gridDim.x * blockDim.x) {
    const float4 p = B.pos[i];
    const int w = P.nWorlds > 1 ? B.world[i] : 0;
    uint32_t key;
    if (B.flags[i] & BF_BIG) {
      key = (uint32_t)P.hashMask + 1u;  // sorts behind every bucket
      A.cellc[i] = make_int4(0, 0, 0, -1);
    } else {
      const double inv = 1.0 / P.cell;
      double fx = floor(W(p.x) * inv), fy = floor(W(p.y) * inv), fz = floor(W(p.z) * inv);
      fx = fmin(fmax(fx, -1.0e9), 1.0e9); fy = fmin(fmax(fy, -1.0e9), 1.0e9); fz = fmin(fmax(fz, -1.0e9), 1.0e9);
      const int cx = (int)fx, cy = (int)fy, cz = (int)fz;
      A.cellc[i] = make_int4(cx, cy, cz, w);
      key = cell_hash(cx, cy, cz, w) & (uint32_t)P.hashMask;
    }
    A.skey[i] = key;
    A.sval[i] = (uint32_t)i;
    if (P.kind == CANNON_BP_GRID) {
      // addBoxToBins, grid_broadphase.dart:100-148: spheres use pos +- radius, everything else the body AABB
      const int sh = B.shape[i];
      double x0, y0, z0, x1, y1, z1;
      const int st = sh >= 0 ? T.shapes[sh].type : -1;
      if (st == CANNON_SHAPE_SPHERE) {
        const double r = T.shapes[sh].radius;
        x0 = W(p.x) - r; y0 = W(p.y) - r; z0 = W(p.z) - r;
        x1 = W(p.x) + r; y1 = W(p.y) + r; z1 = W(p.z) + r;
      } else {
        const float4 lo = B.aabbLo[i], hi = B.aabbHi[i];
        x0 = W(lo.x); y0 = W(lo.y); z0 = W(lo.z);
        x1 = W(hi.x); y1 = W(hi.y); z1 = W(hi.z);
      }
      const int lx = grid_bin_lo(x0, P.gxmin, P.gxmult, P.gnx), ly = grid_bin_lo(y0, P.gymin, P.gymult, P.gny),
                lz = grid_bin_lo(z0, P.gzmin, P.gzmult, P.gnz);
      const int hx = grid_bin_hi(x1, P.gxmin, P.gxmult, P.gnx), hy = grid_bin_hi(y1, P.gymin, P.gymult, P.gny),
                hz = grid_bin_hi(z1, P.gzmin, P.gzmult, P.gnz);
      A.binLo[i] = lx | (ly << 10) | (lz << 20);
      A.binHi[i] = hx | (hy << 10) | (hz << 20);
      if (st == CANNON_SHAPE_PLANE) A.binLo[i] |= (1 << 30);  // plane: membership decided per bin (grid_broadphase.dart:165-195)
    }
  }
}

__global__ void __launch_bounds__(256) k_bp_ranges(BpParams P, BpArrays A) {
  const int nSmall = P.n - P.nBig;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nSmall; k += gridDim.x * blockDim.x) {
    const uint32_t key = A.skey[k];
    if (k == 0 || A.skey[k - 1] != key) A.cellStart[key] = k;
    if (k == nSmall - 1 || A.skey[k + 1] != key) A.cellEnd[key] = k + 1;
  }
}

__global__ void __launch_bounds__(256) k_bp_reorder(BodyArrays B, BpParams P, BpArrays A) {
  const int nSmall = P.n - P.nBig;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nSmall; k += gridDim.x * blockDim.x) {
    const int i = (int)A.sval[k];
    A.spos[k] = B.pos[i];
    A.srad[k] = B.brad[i];
    const int sos = (B.type[i] == CANNON_BODY_STATIC || B.sleep[i] == CANNON_SLEEPING) ? 1 : 0;
    A.smeta[k] = make_int4(B.group[i], B.mask[i], sos, i);
    A.scell[k] = A.cellc[i];
  }
}

struct BpSelf {
  float4 pos;
  double rad;
  int group, mask, sos, idx, world;
  float4 lo, hi;
  int binLo, binHi;
};

// GridBroadphase plane membership for the bin range of the other body (SURVEY.md §5.9-4 intended semantics):
// the plane is in bin (xi,yi,zi) iff d.dot(n) < binRadius with d accumulated in float like the reference's
// Vector3 `d` (grid_broadphase.dart:176-195). d.dot(n) is monotone in each bin index, so the minimum over
// the range sits at the corner picked by the signs of n.
__device__ inline bool grid_plane_shares(const BodyArrays& B, const BpParams& P, int plane, int lo, int hi) {
  f3 z; z.x = 0.f; z.y = 0.f; z.z = 1.f;
  const f3 n = qrot(ldq(B.quat[plane]), z);
  const float4 pp = B.pos[plane];
  const int xi = W(n.x) > 0 ? (lo & 1023) : (hi & 1023);
  const int yi = W(n.y) > 0 ? ((lo >> 10) & 1023) : ((hi >> 10) & 1023);
  const int zi = W(n.z) > 0 ? ((lo >> 20) & 1023) : ((hi >> 20) & 1023);
  float dx = (float)(P.gxmin + P.gbx * 0.5 - W(pp.x));
  float dy = (float)(P.gymin + P.gby * 0.5 - W(pp.y));
  float dz = (float)(P.gzmin + P.gbz * 0.5 - W(pp.z));
  for (int k = 0; k < xi; k++) dx = (float)(W(dx) + P.gbx);
  for (int k = 0; k < yi; k++) dy = (float)(W(dy) + P.gby);
  for (int k = 0; k < zi; k++) dz = (float)(W(dz) + P.gbz);
  f3 d; d.x = dx; d.y = dy; d.z = dz;
  return vdot(d, n) < P.gBinRadius;
}

// needBroadphaseCollision (broadphase.dart:44-63) + intersectionTest (:67-102) [+ Grid bin sharing]
__device__ __forceinline__ bool bp_test(const BodyArrays& B, const BpParams& P, const BpArrays& A, const BpSelf& s, const float4& opos,
                                        double orad, int ogroup, int omask, int osos, int oidx) {
  if ((s.group & omask) == 0 || (ogroup & s.mask) == 0) return false;
  if (s.sos && osos) return false;
  if (P.useBoxes) {  // aabb.dart:131-147
    const float4 l2 = B.aabbLo[oidx], u2 = B.aabbHi[oidx];
    const float4 l1 = s.lo, u1 = s.hi;
    const bool ox = (l2.x <= u1.x && u1.x <= u2.x) || (l1.x <= u2.x && u2.x <= u1.x);
    const bool oy = (l2.y <= u1.y && u1.y <= u2.y) || (l1.y <= u2.y && u2.y <= u1.y);
    const bool oz = (l2.z <= u1.z && u1.z <= u2.z) || (l1.z <= u2.z && u2.z <= u1.z);
    if (!(ox && oy && oz)) return false;
  } else {
    f3 r = vsub(ld3(opos), ld3(s.pos));
    const double sum = s.rad + orad;
    if (!(vlen2(r) < sum * sum)) return false;
  }
  if (P.kind == CANNON_BP_GRID) {
    const int lo2 = A.binLo[oidx], hi2 = A.binHi[oidx];
    const bool plane1 = (s.binLo >> 30) & 1, plane2 = (lo2 >> 30) & 1;
    if (plane1 && plane2) return true;
    if (plane1) return grid_plane_shares(B, P, s.idx, lo2, hi2);
    if (plane2) return grid_plane_shares(B, P, oidx, s.binLo, s.binHi);
    const bool sx = (s.binLo & 1023) <= (hi2 & 1023) && (lo2 & 1023) <= (s.binHi & 1023);
    const bool sy = ((s.binLo >> 10) & 1023) <= ((hi2 >> 10) & 1023) && ((lo2 >> 10) & 1023) <= ((s.binHi >> 10) & 1023);
    const bool sz = ((s.binLo >> 20) & 1023) <= ((hi2 >> 20) & 1023) && ((lo2 >> 20) & 1023) <= ((s.binHi >> 20) & 1023);
    if (!(sx && sy && sz)) return false;
  }
  return true;
}

#define BP_CACHE 16  // partners remembered by the counting pass (per body, global scratch): the emit pass of a body with
                     // at most this many partners sorts and writes them without walking its 27 cells again
struct NbState {
  int mode;   // 0 count, 1 collect into buf, 2 select the smallest j > last, 3 count + remember the first BP_CACHE
  int count;
  int last, best;
  int* buf;
};
__device__ __forceinline__ void nb_accept(NbState& S, int j) {
  if (S.mode == 0) S.count++;
  else if (S.mode == 3) { if (S.count < BP_CACHE) S.buf[S.count] = j; S.count++; }
  else if (S.mode == 1) { if (S.count < BP_MAXNB) S.buf[S.count] = j; S.count++; }
  else if (j > S.last && j < S.best) S.best = j;
}

__device__ inline void bp_load_self(const BodyArrays& B, const BpParams& P, const BpArrays& A, int i, BpSelf& s) {
  s.pos = B.pos[i];
  s.rad = B.brad[i];
  s.group = B.group[i];
  s.mask = B.mask[i];
  s.sos = (B.type[i] == CANNON_BODY_STATIC || B.sleep[i] == CANNON_SLEEPING) ? 1 : 0;
  s.idx = i;
  s.world = P.nWorlds > 1 ? B.world[i] : 0;
  if (P.useBoxes) { s.lo = B.aabbLo[i]; s.hi = B.aabbHi[i]; }
  if (P.kind == CANNON_BP_GRID) { s.binLo = A.binLo[i]; s.binHi = A.binHi[i]; }
}

// partners j < i of a small body i: 27 grid cells + big bodies
__device__ inline void bp_enum_small(const BodyArrays& B, const BpParams& P, const BpArrays& A, const BpSelf& s, const int4 c, NbState& S) {
  for (int dz = -1; dz <= 1; dz++)
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++) {
        const int cx = c.x + dx, cy = c.y + dy, cz = c.z + dz;
        const uint32_t h = cell_hash(cx, cy, cz, c.w) & (uint32_t)P.hashMask;
        const int e = A.cellEnd[h];
        for (int k = A.cellStart[h]; k < e; k++) {
          const int4 m = A.smeta[k];
          if (m.w >= s.idx) continue;
          const int4 oc = A.scell[k];
          if (oc.x != cx || oc.y != cy || oc.z != cz || oc.w != c.w) continue;  // other cell hashed into this bucket
          if (bp_test(B, P, A, s, A.spos[k], A.srad[k], m.x, m.y, m.z, m.w)) nb_accept(S, m.w);
        }
      }
  const int b0 = A.bigWorldStart[s.world], b1 = A.bigWorldStart[s.world + 1];
  for (int t = b0; t < b1; t++) {
    const int j = A.bigList[t];
    if (j >= s.idx) break;  // ascending
    const int osos = (B.type[j] == CANNON_BODY_STATIC || B.sleep[j] == CANNON_SLEEPING) ? 1 : 0;
    if (bp_test(B, P, A, s, B.pos[j], B.brad[j], B.group[j], B.mask[j], osos, j)) nb_accept(S, j);
  }
}

// pass 0: counts; pass 1: write pairs (i, j ascending) at offs[i]
__global__ void __launch_bounds__(128) k_bp_small(BodyArrays B, BpParams P, BpArrays A, int pass, int* __restrict__ p1, int* __restrict__ p2,
                                                  int cap, int* __restrict__ overflow) {
  const int nSmall = P.n - P.nBig;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nSmall; k += gridDim.x * blockDim.x) {
    const int i = (int)A.sval[k];
    BpSelf s;
    bp_load_self(B, P, A, i, s);
    const int4 c = A.scell[k];
    NbState S;
    if (pass == 0) {
      S.mode = A.nbCache ? 3 : 0; S.count = 0; S.buf = A.nbCache ? A.nbCache + (size_t)k * BP_CACHE : nullptr;
      bp_enum_small(B, P, A, s, c, S);
      A.counts[i] = S.count;
      continue;
    }
    const int cnt = A.counts[i];
    if (cnt == 0) continue;
    const int off = A.offs[i];
    if (off + cnt > cap) { atomicMax(overflow, off + cnt); continue; }
    if (cnt <= BP_CACHE && A.nbCache) {
      int buf[BP_CACHE];
      const int4* src = (const int4*)(A.nbCache + (size_t)k * BP_CACHE);
#pragma unroll
      for (int q = 0; q < BP_CACHE / 4; q++) {
        const int4 v = src[q];
        buf[4 * q] = v.x; buf[4 * q + 1] = v.y; buf[4 * q + 2] = v.z; buf[4 * q + 3] = v.w;
      }
      // rank sort by j (partners are distinct): no data-dependent indexing, the buffer stays in registers
#pragma unroll
      for (int a = 0; a < BP_CACHE; a++) {
        if (a < cnt) {
          int rank = 0;
#pragma unroll
          for (int b = 0; b < BP_CACHE; b++) rank += (b < cnt && buf[b] < buf[a]) ? 1 : 0;
          p1[off + rank] = i; p2[off + rank] = buf[a];
        }
      }
    } else if (cnt <= BP_MAXNB) {
      int buf[BP_MAXNB];
      S.mode = 1; S.count = 0; S.buf = buf;
      bp_enum_small(B, P, A, s, c, S);
      for (int a = 1; a < cnt; a++) {  // insertion sort by j
        const int v = buf[a];
        int b = a - 1;
        while (b >= 0 && buf[b] > v) { buf[b + 1] = buf[b]; b--; }
        buf[b + 1] = v;
      }
      for (int a = 0; a < cnt; a++) { p1[off + a] = i; p2[off + a] = buf[a]; }
    } else {
      int last = -1;
      for (int a = 0; a < cnt; a++) {
        S.mode = 2; S.last = last; S.best = 0x7fffffff; S.buf = nullptr;
        bp_enum_small(B, P, A, s, c, S);
        last = S.best;
        p1[off + a] = i; p2[off + a] = last;
      }
    }
  }
}

// one block per big body b: partners are all j < b of its world; ordered block compaction
// NaiveBroadphase of a batch of small worlds (naive_broadphase.dart:14-33 inside every world): a warp owns body i and tests
// the bodies j < i of ITS world, 32 at a time; ballots place the accepted pairs in j order, so the output is the reference's
// i-major / j-ascending list without the hash grid, its radix sort and the big-body pass (config 4: 64-body worlds - the
// grid machinery was ~20 launches and 0.2 ms for 2016 candidate pairs per world). pass 0 counts, pass 1 (after the scan) emits.
__global__ void __launch_bounds__(128) k_bp_world_all(BodyArrays B, BpParams P, BpArrays A, int n, int pass, int* __restrict__ p1, int* __restrict__ p2,
                                                      int cap, int* __restrict__ overflow) {
  const int lane = threadIdx.x & 31;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += (gridDim.x * blockDim.x) >> 5) {
    BpSelf s;
    bp_load_self(B, P, A, i, s);
    const int j0 = A.worldStart[s.world];
    const int off = pass ? A.offs[i] : 0;
    const int cnt = pass ? A.counts[i] : 0;
    const bool writable = pass && (off + cnt <= cap);
    if (pass && !writable && lane == 0 && cnt > 0) atomicMax(overflow, off + cnt);
    int run = 0;
    for (int base = j0; base < i; base += 32) {
      const int j = base + lane;
      bool ok = false;
      if (j < i) {
        const int osos = (B.type[j] == CANNON_BODY_STATIC || B.sleep[j] == CANNON_SLEEPING) ? 1 : 0;
        ok = bp_test(B, P, A, s, B.pos[j], B.brad[j], B.group[j], B.mask[j], osos, j);
      }
      const unsigned bits = __ballot_sync(0xffffffffu, ok);
      if (ok && writable) {
        const int k = off + run + __popc(bits & ((1u << lane) - 1u));
        p1[k] = i; p2[k] = j;
      }
      run += __popc(bits);
    }
    if (pass == 0 && lane == 0) A.counts[i] = run;
  }
}

__global__ void __launch_bounds__(256) k_bp_big(BodyArrays B, BpParams P, BpArrays A, int pass, int* __restrict__ p1, int* __restrict__ p2,
                                                int cap, int* __restrict__ overflow) {
  __shared__ int s_run;
  for (int t = blockIdx.x; t < P.nBig; t += gridDim.x) {
    const int i = A.bigList[t];
    BpSelf s;
    bp_load_self(B, P, A, i, s);
    const int j0 = A.worldStart[s.world];
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    const int off = pass ? A.offs[i] : 0;
    const int cnt = pass ? A.counts[i] : 0;
    const bool writable = pass && (off + cnt <= cap);
    if (pass && !writable && threadIdx.x == 0 && cnt > 0) atomicMax(overflow, off + cnt);
    for (int base = j0; base < i; base += blockDim.x) {
      const int j = base + threadIdx.x;
      int ok = 0;
      if (j < i) {
        const int osos = (B.type[j] == CANNON_BODY_STATIC || B.sleep[j] == CANNON_SLEEPING) ? 1 : 0;
        ok = bp_test(B, P, A, s, B.pos[j], B.brad[j], B.group[j], B.mask[j], osos, j) ? 1 : 0;
      }
      int tot;
      const int ex = block_excl_scan(ok, &tot);
      const int run = s_run;
      if (ok && writable) { p1[off + run + ex] = i; p2[off + run + ex] = j; }
      __syncthreads();
      if (threadIdx.x == 0) s_run = run + tot;
      __syncthreads();
    }
    if (pass == 0 && threadIdx.x == 0) A.counts[i] = s_run;
    __syncthreads();
  }
}

// ---- SAP -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sap_keys(BodyArrays B, BpParams P, BpArrays A) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < P.n; k += gridDim.x * blockDim.x) {
    const int b = (int)A.sapList[k];
    const float4 lo = B.aabbLo[b];
    float v = P.sapAxis == 0 ? lo.x : (P.sapAxis == 1 ? lo.y : lo.z);
    v = v + 0.0f;  // -0 -> +0: the reference's `<=` compares them equal
    A.sapKey[k] = float_to_ordered(v);
  }
}

// one warp per sorted position i: forward sweep, lanes test 32 consecutive j; `continue` on filtered pairs,
// `break` at the first unfiltered j with pos_j - r_j >= pos_i + r_i (sap_broadphase.dart:149-165)
__global__ void __launch_bounds__(256) k_sap_sweep(BodyArrays B, BpParams P, BpArrays A, int pass, int* __restrict__ p1, int* __restrict__ p2,
                                                   int cap, int* __restrict__ overflow) {
  const int lane = threadIdx.x & 31;
  const int warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
  for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < P.n; k += warpsPerGrid) {
    const int i = (int)A.sapList[k];
    BpSelf s;
    bp_load_self(B, P, A, i, s);
    const double pi = P.sapAxis == 0 ? W(s.pos.x) : (P.sapAxis == 1 ? W(s.pos.y) : W(s.pos.z));
    const double boundA2 = pi + s.rad;
    int run = 0;
    const int off = pass ? A.offs[k] : 0;
    const int cnt = pass ? A.counts[k] : 0;
    const bool writable = pass && (off + cnt <= cap);
    if (pass && !writable && lane == 0 && cnt > 0) atomicMax(overflow, off + cnt);
    for (int base = k + 1; base < P.n; base += 32) {
      const int kk = base + lane;
      bool brk = false, ok = false;
      int j = -1;
      if (kk < P.n) {
        j = (int)A.sapList[kk];
        const int ogroup = B.group[j], omask = B.mask[j];
        const int osos = (B.type[j] == CANNON_BODY_STATIC || B.sleep[j] == CANNON_SLEEPING) ? 1 : 0;
        const bool need = !((s.group & omask) == 0 || (ogroup & s.mask) == 0) && !(s.sos && osos);
        if (need) {
          const float4 op = B.pos[j];
          const double orad = B.brad[j];
          const double pj = P.sapAxis == 0 ? W(op.x) : (P.sapAxis == 1 ? W(op.y) : W(op.z));
          const double boundB1 = pj - orad;
          if (!(boundB1 < boundA2)) brk = true;
          else ok = bp_test(B, P, A, s, op, orad, ogroup, omask, osos, j);
        }
      }
      const unsigned brkMask = __ballot_sync(0xffffffffu, brk);
      const unsigned before = brkMask ? ((1u << (__ffs(brkMask) - 1)) - 1u) : 0xffffffffu;
      const unsigned okMask = __ballot_sync(0xffffffffu, ok) & before;
      if (ok && ((before >> lane) & 1u) && writable) {
        const int r = run + __popc(okMask & ((1u << lane) - 1u));
        p1[off + r] = i; p2[off + r] = j;
      }
      run += __popc(okMask);
      if (brkMask) break;
    }
    if (pass == 0 && lane == 0) A.counts[k] = run;
  }
}
