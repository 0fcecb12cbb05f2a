// k_narrowphase.cuh — contact generation (K3), lib/world/narrow_phase.dart:634-2178 for the in-scope shapes.
//
// Pipeline (all counts stay on the device):
//   k_np_tasks (pass 0/1)  per pair: common prologue of getContacts (:670-716) -> number of resolver tasks
//                          (1 for ordinary pairs, 2 per heightfield cell of the index window) -> scan -> task list
//                          in the reference's loop order, bucketed by resolver type
//   k_np_<resolver>        one kernel per resolver type over its bucket; contacts go to a raw pool
//   scan(task counts)      canonical contact offsets (pair order, then the resolver's emission order)
//   k_np_finalize          raw pool -> final contact SoA + material-derived parameters (:492-586)
// so the final contact order equals the reference's regardless of how threads were scheduled.
#pragma once
#include "world.cuh"

enum { NP_SS = 0, NP_SP = 1, NP_SB = 2, NP_SH = 3, NP_PH = 4, NP_HH = 5, NP_SPIL = 6, NP_HPIL = 7,
       NP_SPT = 8, NP_PPT = 9, NP_HPT = 10, NP_PILPT = 11,  // sphere / plane / hull / heightfield pillar against a Particle
       NP_STM = 12, NP_PTM = 13,  // sphere / plane against a Trimesh
       NP_NTYPES = 14 };
__device__ __forceinline__ bool np_pillar_code(int c) { return c == NP_SPIL || c == NP_HPIL || c == NP_PILPT; }
#define NP_MAXPOLY 40

struct NpArrays {
  const int* p1;
  const int* p2;
  const int* nPairs;       // device count
  int* pairTasks;          // tasks per pair, then exclusive scan (in place into pairTaskOff)
  unsigned long long* pairMask;  // heightfield pairs: surviving pillars of the index window (pass 0 -> pass 1)
  int* pairTaskOff;
  int* nTasks;
  int* taskPair;
  int* taskInfo;           // type | upper<<4
  int2* taskCell;
  int* bucket;             // [taskCap], one segment per resolver type
  int* bucketCount;        // [NP_NTYPES] tasks per type (pass 0)
  int* bucketStart;        // [NP_NTYPES] exclusive scan of bucketCount
  int* bucketCursor;       // [NP_NTYPES] fill cursors (pass 1)
  int* taskCnt;            // contacts per task
  int* taskRaw;            // start in the raw pool
  int* taskOff;            // exclusive scan of taskCnt
  int* rawCount;
  float4 *rawRi, *rawRj, *rawNi;
  int taskCap, contactCap;
  int debug;               // profiling aid (CANNON_NP_DEBUG): 1 skip clipping, 2 skip the axis loop, 3 skip after pillar build
  int* overflowTasks;
  int* overflowContacts;
  int* taskHit;            // contact events on: per task, "a justTest resolver returned true" (nullptr otherwise: justTest pairs make no tasks)
  int* unsupported;        // set when a pair only the reference's unfinished trimesh resolvers would handle passes the prologue
  // tile SAT kernel (k_sat_warp.cuh): tasks that passed the separating-axis test, queued for the clipping launch
  int* clipList;           // [2][taskCap]: hull/hull, hull/pillar
  int* nClip;              // [2]
  float4* taskSep;         // separating axis of a queued task
};

struct ContactArrays {
  int* nContacts;
  int *bi, *bj;
  float4 *ri, *rj, *ni;
  double *rest, *mu, *slip;            // restitution, friction coefficient, slip force mu*|g|*m_red
  double *ca, *cb, *ceps;              // contact SPOOK
  double *fb, *feps;                   // friction SPOOK (a is unused: g == 0)
  int* enabled;
  int* row;                            // solver row of the contact equation, -1 if filtered
  int* task;                           // resolver task (manifold) the contact came from
};

struct HullView {
  const float4* v; int nV;
  const float4* n; const double* pc; int nF;
  const int* fvOff; const int* fvIdx;   // fvOff[f]..fvOff[f+1]
  const int* fcOff; const int* fcIdx;
  const float4* e; int nE;
  const float4* ek; int nEk;   // edges / faces left after dropping +-copies (tile SAT kernel only)
  const int* fk; int nFk;
  int hasAxes;
  double bsr;
};

__device__ __forceinline__ HullView hull_view(const ShapeTables& T, int hull) {
  const HullDev h = T.hulls[hull];
  HullView H;
  H.v = T.verts + h.vOff; H.nV = h.nV;
  H.n = T.fnormals + h.fOff; H.pc = T.fplanec + h.fOff; H.nF = h.nF;
  H.fvOff = T.fvOff + h.fOff + hull; H.fvIdx = T.fvIdx;
  H.fcOff = T.fcOff + h.fOff + hull; H.fcIdx = T.fcIdx;
  H.e = T.edges + h.eOff; H.nE = h.nE;
  H.ek = T.edgesK + h.ekOff; H.nEk = h.nEk;
  H.fk = T.facesK + h.fkOff; H.nFk = h.nFk;
  H.hasAxes = h.hasAxes;
  H.bsr = h.bsr;
  return H;
}

__device__ __forceinline__ bool is_hull_type(int t) { return t >= CANNON_SHAPE_BOX && t <= CANNON_SHAPE_SIZED_PLANE; }  // box + the ConvexPolyhedron subclasses

// heightfield index window: sphereHeightfield :1318-1362 / heightfieldConvex :2070-2113 (+ getRectMinMax, heightfield.dart:146-162)
__device__ inline bool hf_window(const ShapeTables& T, const HfDev& hf, const f3& local, double radius, int& iMinX, int& iMaxX, int& iMinY,
                                 int& iMaxY) {
  const double wd = (double)hf.esize;
  iMinX = (int)floor((W(local.x) - radius) / wd) - 1;
  iMaxX = (int)ceil((W(local.x) + radius) / wd) + 1;
  iMinY = (int)floor((W(local.y) - radius) / wd) - 1;
  iMaxY = (int)ceil((W(local.y) + radius) / wd) + 1;
  if (iMaxX < 0 || iMaxY < 0 || iMinX > hf.nx || iMinY > hf.ny) return false;
  if (iMinX < 0) iMinX = 0;
  if (iMaxX < 0) iMaxX = 0;
  if (iMinY < 0) iMinY = 0;
  if (iMaxY < 0) iMaxY = 0;
  if (iMinX >= hf.nx) iMinX = hf.nx - 1;
  if (iMaxX >= hf.nx) iMaxX = hf.nx - 1;
  if (iMaxY >= hf.ny) iMaxY = hf.ny - 1;
  if (iMinY >= hf.ny) iMinY = hf.ny - 1;
  // the window maximum below never exceeds the heightfield's maximum: a shape above that (most of a pile) is rejected
  // by the test after the loop whatever the window holds, so the loop (up to 36 scattered loads per pair) is skipped
  if (W(local.z) - radius > hf.maxV) return false;
  double mx = hf.minV;
  for (int i = iMinX; i <= iMaxX; i++)
    for (int j = iMinY; j <= iMaxY; j++) {
      const double h = T.hfdata[hf.dataOff + (size_t)i * hf.ny + j];
      if (h > mx) mx = h;
    }
  if (W(local.z) - radius > mx || W(local.z) + radius < hf.minV) return false;
  return true;
}

// Offset and bounding-sphere radius of one triangle pillar (same arithmetic as build_pillar below): enough for the
// `distanceTo(worldPillarOffset) < pillar.boundingSphereRadius + r` gate of narrow_phase.dart:1374-1377,2121-2124,
// so rejected pillars never become tasks.
__device__ inline void pillar_bounds(const ShapeTables& T, const HfDev& hf, int xi, int yi, bool upper, f3& offset, double& bsr, f3* v) {
  const double es = (double)hf.esize;
  const double* d = T.hfdata + hf.dataOff;
  const double h00 = d[(size_t)xi * hf.ny + yi], h10 = d[(size_t)(xi + 1) * hf.ny + yi], h01 = d[(size_t)xi * hf.ny + yi + 1],
               h11 = d[(size_t)(xi + 1) * hf.ny + yi + 1];
  const double h = (fmin(fmin(h00, h10), fmin(h01, h11)) - hf.minV) / 2 + hf.minV;
  if (!upper) {
    offset = mk3((xi + 0.25) * es, (yi + 0.25) * es, h);
    v[0] = mk3(-0.25 * es, -0.25 * es, h00 - h);
    v[1] = mk3(0.75 * es, -0.25 * es, h10 - h);
    v[2] = mk3(-0.25 * es, 0.75 * es, h01 - h);
    v[3] = mk3(-0.25 * es, -0.25 * es, -h - 1);
    v[4] = mk3(0.75 * es, -0.25 * es, -h - 1);
    v[5] = mk3(-0.25 * es, 0.75 * es, -h - 1);
  } else {
    offset = mk3((xi + 0.75) * es, (yi + 0.75) * es, h);
    v[0] = mk3(0.25 * es, 0.25 * es, h11 - h);
    v[1] = mk3(-0.75 * es, 0.25 * es, h01 - h);
    v[2] = mk3(0.25 * es, -0.75 * es, h10 - h);
    v[3] = mk3(0.25 * es, 0.25 * es, -h - 1);
    v[4] = mk3(-0.75 * es, 0.25 * es, -h - 1);
    v[5] = mk3(0.25 * es, -0.75 * es, -h - 1);
  }
  double max2 = 0;
  for (int i = 0; i < 6; i++) {
    const double n2 = vlen2(v[i]);
    if (n2 > max2) max2 = n2;
  }
  bsr = sqrt(max2);
}

// min/max of `nV` local vertices along a world axis, ConvexPolyhedron.project (convex_polyhedron.dart:843-883) with the
// axis-independent local origin hoisted out
__device__ __forceinline__ void project_verts(const float4* __restrict__ gv, const f3* __restrict__ lv, int nV, const f3& axis, const q4& quat,
                                              const f3& localOrigin, double& mx, double& mn) {
  const f3 localAxis = qrot(qnegw(quat), axis);
  const double add = vdot(localOrigin, localAxis);
  mn = mx = vdot(gv ? ld3(gv[0]) : lv[0], localAxis);
  for (int i = 1; i < nV; i++) {
    const double val = vdot(gv ? ld3(gv[i]) : lv[i], localAxis);
    if (val > mx) mx = val;
    if (val < mn) mn = val;
  }
  mn -= add;
  mx -= add;
  if (mn > mx) { const double t = mn; mn = mx; mx = t; }
}

// Exact early rejection of a hull/pillar task: convexConvex reports no contact as soon as ANY axis of its test
// set separates the hulls (convex_polyhedron.dart:264-267,338-341). The set always contains face 0 of the hull
// (when it has `uniqueAxes`) and cross(hull edge, pillar vertical edge (0,0,1)); testing just those with the
// very same projection arithmetic kills most window pillars without changing a single result.
// The axes of that test depend on the hull and on the heightfield's orientation only, not on the pillar: the warp
// evaluates each axis, its hull projection and its image in the heightfield frame ONCE per pair (one axis per lane,
// QsAxis in shared memory) and every lane then projects only the six vertices of its own pillar.
struct QsAxis {
  double lx, ly, lz; // qrot(qnegw(qP), axis): the axis in the heightfield's local frame, widened once (float -> double is exact):
                     // every pillar of the pair multiplies it with its six vertices, and the conversions are the scarce pipe
  double maxA, minA; // hull projection (ConvexPolyhedron.project of the hull on the axis)
  int valid;
};
#define QS_MAX_AXES 33

// what k_np_tasks needs about one heightfield pair, fetched by the pair's own lane and read by the whole warp
struct HfPairCtx {
  f3 xf, xs, oA;
  q4 qs, qf;
  double rFirst;
  unsigned long long mask;  // survivor mask cached by pass 0 (windows of <= 64 pillars)
  HfDev hf;
  HullDev hd;
  int hullTask;
  f3 lc;          // sphere tasks: the sphere centre in the heightfield frame
  double margin;  // and the slack of sphere_pillar_far
};

// Conservative rejection of a sphere / pillar task. sphereConvex (narrow_phase.dart:1039-1257) reports a contact only when
// a vertex, a face point or an edge point of the pillar lies within the sphere radius of the centre, i.e. when the
// centre is closer than R to the pillar; the pillar lies inside the box [x0,x1] x [y0,y1] x (-inf, top] of its triangle
// in the heightfield frame. A centre farther than R + margin from that box cannot produce a contact; the margin
// (1 % of the cell plus 1e-4 of the coordinate magnitudes) is three orders of magnitude above the f32 rounding of the
// reference's world-frame arithmetic, so no reachable contact is ever dropped and the surviving tasks run the
// unchanged resolver. Rejected pillars never become tasks (nine tenths of a resting sphere's window).
__device__ __forceinline__ bool sphere_pillar_far(const f3& lc, double R, double margin, const f3& po, const f3* pv) {
  const double ox = W(po.x), oy = W(po.y), oz = W(po.z);
  const double x0 = fmin(fmin(W(pv[0].x), W(pv[1].x)), W(pv[2].x)) + ox, x1 = fmax(fmax(W(pv[0].x), W(pv[1].x)), W(pv[2].x)) + ox;
  const double y0 = fmin(fmin(W(pv[0].y), W(pv[1].y)), W(pv[2].y)) + oy, y1 = fmax(fmax(W(pv[0].y), W(pv[1].y)), W(pv[2].y)) + oy;
  const double zt = fmax(fmax(W(pv[0].z), W(pv[1].z)), W(pv[2].z)) + oz;
  const double cx = W(lc.x), cy = W(lc.y), cz = W(lc.z);
  const double dx = fmax(0.0, fmax(x0 - cx, cx - x1)), dy = fmax(0.0, fmax(y0 - cy, cy - y1)), dz = fmax(0.0, cz - zt);
  const double r = R + margin;
  return dx * dx + dy * dy + dz * dz > r * r;
}

__device__ __forceinline__ bool pillar_quick_separated_pre(const QsAxis* __restrict__ ax, int nAx, const f3* pv, const f3& xP, const q4& qP) {
  f3 zero; zero.x = zero.y = zero.z = 0.f;
  const f3 oP = to_local_point(xP, qP, zero);
  const double ox = W(oP.x), oy = W(oP.y), oz = W(oP.z);
  // vdot(v, localAxis) = (vx * lx + vy * ly) + vz * lz on the widened components, exactly as dmath.cuh evaluates it
  double px[6], py[6], pz[6];  // the pillar's vertices widened once for all axes
#pragma unroll
  for (int i = 0; i < 6; i++) { px[i] = W(pv[i].x); py[i] = W(pv[i].y); pz[i] = W(pv[i].z); }
  for (int t = 0; t < nAx; t++) {
    if (!ax[t].valid) continue;
    const double lx = ax[t].lx, ly = ax[t].ly, lz = ax[t].lz;
    const double add = ox * lx + oy * ly + oz * lz;
    double mn, mx;
    mn = mx = px[0] * lx + py[0] * ly + pz[0] * lz;
#pragma unroll
    for (int i = 1; i < 6; i++) {
      const double val = px[i] * lx + py[i] * ly + pz[i] * lz;
      if (val > mx) mx = val;
      if (val < mn) mn = val;
    }
    mn -= add;
    mx -= add;
    if (mn > mx) { const double tt = mn; mn = mx; mx = tt; }
    if (ax[t].maxA < mn || mx < ax[t].minA) return true;
  }
  return false;
}

// one quick-separation axis of a hull / heightfield pair: face 0 of the hull (t == 0 when it has uniqueAxes) or
// cross(hull edge, heightfield z), its hull projection and its image in the heightfield frame
__device__ __forceinline__ QsAxis qs_axis(const ShapeTables& T, const HullDev& hd, const q4& qf, const f3& oA, const q4& qs, int t) {
  f3 up; up.x = 0.f; up.y = 0.f; up.z = 1.f;
  const int e = t - (hd.hasAxes ? 1 : 0);
  f3 axis;
  bool valid = true;
  if (e < 0) axis = qrot(qf, ld3(T.fnormals[hd.fOff]));
  else {
    axis = vcross(qrot(qf, ld3(T.edgesK[hd.ekOff + e])), qrot(qs, up));
    if (valmost_zero(axis)) valid = false;
    else vnormalize(axis);
  }
  QsAxis q;
  q.valid = valid ? 1 : 0;
  q.maxA = q.minA = 0.0;
  q.lx = W(axis.x); q.ly = W(axis.y); q.lz = W(axis.z);
  if (valid) {
    project_verts(T.verts + hd.vOff, nullptr, hd.nV, axis, qf, oA, q.maxA, q.minA);
    const f3 lax = qrot(qnegw(qs), axis);
    q.lx = W(lax.x); q.ly = W(lax.y); q.lz = W(lax.z);
  }
  return q;
}
#define QS_HOIST_AXES 6  // pairs with at most this many quick axes (box 4, 8-segment cylinder 6) get them computed up front

// which body plays "i" for the resolver: lower ShapeType index first, equal types swapped (narrow_phase.dart:706-710)
__device__ __forceinline__ void np_order(int a, int b, int ta, int tb, int& first, int& second) {
  if (ta < tb) { first = a; second = b; } else { first = b; second = a; }
}

// Pass 0 counts the tasks of every pair, pass 1 (after the scan) writes them. Ordinary pairs are handled by their
// own thread. Heightfield pairs are expanded *warp-cooperatively*: the warp takes its heightfield pairs one at a time
// and every lane tests one pillar of the index window (bounding gate + exact quick separation), a ballot gives the
// survivor mask in the reference's loop order (i, j, lower/upper), so counts and emission order need no atomics.
__global__ void __launch_bounds__(128) k_np_tasks(BodyArrays B, ShapeTables T, NpArrays A, int pass, int hoistBytes) {
  __shared__ QsAxis s_qs[4][QS_MAX_AXES];
  __shared__ HfPairCtx s_ctx[4][32];
  __shared__ int s_hoistOff[4 * 33];
  extern __shared__ __align__(16) unsigned char s_dynTasks[];  // pass 0: QsAxis[4][32][QS_HOIST_AXES]
  QsAxis* const s_hoist = hoistBytes ? (QsAxis*)s_dynTasks : nullptr;
  QsAxis* const qs_ax = s_qs[threadIdx.x >> 5];
  const int np = *A.nPairs;
  const int lane = threadIdx.x & 31;
  const int warpStart = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 5;
  const int stride = gridDim.x * blockDim.x;
  for (int kb = warpStart; kb < np; kb += stride) {
    const int k = kb + lane;
    int nt = 0, code = -1, jt = 0;
    int first = 0, second = 0;
    int iMinX = 0, iMaxX = 0, iMinY = 0, iMaxY = 0;
    bool hfPair = false;
    if (k < np) {
      const int a = A.p1[k], b = A.p2[k];
      const int sa = B.shape[a], sb = B.shape[b];
      if (sa >= 0 && sb >= 0) {
        const ShapeDev si = T.shapes[sa], sj = T.shapes[sb];
        const int tya = B.type[a], tyb = B.type[b];
        const bool justTest = (tya == CANNON_BODY_KINEMATIC && tyb == CANNON_BODY_STATIC) || (tya == CANNON_BODY_STATIC && tyb == CANNON_BODY_KINEMATIC) ||
                              (tya == CANNON_BODY_KINEMATIC && tyb == CANNON_BODY_KINEMATIC);
        const bool maskOk = (si.mask & sj.group) != 0 && (sj.mask & si.group) != 0;
        const f3 xi = ld3(B.pos[a]), xj = ld3(B.pos[b]);
        // kinematic / static pairs run their resolver in justTest mode (narrow_phase.dart:706-716): no equations, only the
        // overlap keepers behind the contact events - so they become tasks only when events are on (bit 5 of taskInfo)
        jt = justTest ? 1 : 0;
        if (maskOk && (!justTest || A.taskHit != nullptr) && !(vdist(xi, xj) > si.bsr + sj.bsr)) {
          int lo = si.type, hi = sj.type;
          first = a; second = b;
          if (!(lo < hi)) { int t = lo; lo = hi; hi = t; first = b; second = a; }
          if (hi == CANNON_SHAPE_TRIMESH) {  // narrow_phase.dart:212-238 (heightfield-trimesh has no key)
            if (lo == CANNON_SHAPE_SPHERE) code = NP_STM;
            else if (lo == CANNON_SHAPE_PLANE) code = NP_PTM;
            else if (lo != CANNON_SHAPE_HEIGHTFIELD) atomicExch(A.unsupported, 1);  // boxTrimesh / trimeshConvex / particleTrimesh / trimeshTrimesh
          } else if (hi == CANNON_SHAPE_PARTICLE) {  // narrow_phase.dart:196-211 (particle-particle has no key)
            if (lo == CANNON_SHAPE_SPHERE) code = NP_SPT;
            else if (lo == CANNON_SHAPE_PLANE) code = NP_PPT;
            else if (is_hull_type(lo)) code = NP_HPT;
            else if (lo == CANNON_SHAPE_HEIGHTFIELD) { code = NP_PILPT; const int t = first; first = second; second = t; }  // expansion: `first` is the small shape
          } else if (lo == CANNON_SHAPE_SPHERE) {
            if (hi == CANNON_SHAPE_SPHERE) code = NP_SS;
            else if (hi == CANNON_SHAPE_PLANE) code = NP_SP;
            else if (hi == CANNON_SHAPE_BOX) code = NP_SB;
            else if (is_hull_type(hi)) code = NP_SH;
            else if (hi == CANNON_SHAPE_HEIGHTFIELD) code = NP_SPIL;
          } else if (lo == CANNON_SHAPE_PLANE) {
            if (is_hull_type(hi)) code = NP_PH;
          } else if (is_hull_type(lo)) {
            // `convexSizedPlane` is the one key of narrow_phase.dart:336-473 that can never match its lower-cased name
            if (is_hull_type(hi)) code = (lo == CANNON_SHAPE_CONVEX && hi == CANNON_SHAPE_SIZED_PLANE) ? -1 : NP_HH;
            else if (hi == CANNON_SHAPE_HEIGHTFIELD) code = NP_HPIL;
          }
          if (code >= 0) nt = 1;
          if (np_pillar_code(code)) {
            const ShapeDev s1 = T.shapes[B.shape[first]], s2 = T.shapes[B.shape[second]];
            const HfDev hf = T.hfs[s2.hf];
            const f3 local = to_local_point(ld3(B.pos[second]), ldq(B.quat[second]), ld3(B.pos[first]));
            const double radius = code == NP_SPIL ? s1.radius : (code == NP_PILPT ? s1.bsr : T.hulls[s1.hull].bsr);
            nt = 0;
            hfPair = hf_window(T, hf, local, radius, iMinX, iMaxX, iMinY, iMaxY) && iMaxX > iMinX && iMaxY > iMinY;
          }
        }
      }
    }
    // ---- warp-cooperative expansion of the heightfield pairs of this warp ----
    // every heightfield pair's own lane fetches what the warp will need (all lanes in parallel: one round of memory
    // latency per batch instead of one per pair); pass 1 only needs the survivor mask cached by pass 0
    HfPairCtx* const ctxs = s_ctx[threadIdx.x >> 5];
    __syncwarp();  // the previous batch is done with the scratch
    if (hfPair) {
      HfPairCtx& X = ctxs[lane];
      const bool cachedOwn = pass && (iMaxX - iMinX) * (iMaxY - iMinY) * 2 <= 64;
      X.mask = cachedOwn ? A.pairMask[k] : 0ull;
      if (!cachedOwn) {
        const ShapeDev s1 = T.shapes[B.shape[first]], s2 = T.shapes[B.shape[second]];
        X.hf = T.hfs[s2.hf];
        X.xf = ld3(B.pos[first]); X.xs = ld3(B.pos[second]);
        X.qs = ldq(B.quat[second]);
        X.hullTask = code == NP_HPIL;
        if (X.hullTask) {
          X.hd = T.hulls[s1.hull];
          X.qf = ldq(B.quat[first]);
          f3 zero; zero.x = zero.y = zero.z = 0.f;
          X.oA = to_local_point(X.xf, X.qf, zero);
          X.rFirst = X.hd.bsr;
        } else {
          X.rFirst = s1.bsr;
          X.lc = to_local_point(X.xs, X.qs, X.xf);
          if (code == NP_PILPT) X.margin = INFINITY;  // heightfieldParticle (:2421-2440) has the bounding gate only
          else X.margin = fmax(0.0, s1.radius - s1.bsr) + 0.01 * (double)X.hf.esize + 1e-4 * (fabs(W(X.xf.x)) + fabs(W(X.xf.y)) + fabs(W(X.xf.z)) + fabs(W(X.xs.x)) + fabs(W(X.xs.y)) +
                                                         fabs(W(X.xs.z)) + fabs(W(X.lc.x)) + fabs(W(X.lc.y)) + fabs(W(X.lc.z)));
        }
      }
    }
    __syncwarp();
    // quick-separation axes of ALL hull / heightfield pairs of the batch, one (pair, axis) item per lane: the serial
    // loop below then starts with its axes ready instead of computing 4-6 of them with 4-6 lanes per pair
    const bool hoisted = pass == 0 && s_hoist != nullptr;
    if (hoisted) {
      int myAx = 0;
      if (hfPair && code == NP_HPIL) {
        const HullDev& h = ctxs[lane].hd;
        if (h.nE <= 32 && h.nF <= 32) { myAx = (h.hasAxes ? 1 : 0) + h.nEk; if (myAx > QS_HOIST_AXES) myAx = 0; }
      }
      int incl = myAx;
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      int* const ex = s_hoistOff + (threadIdx.x >> 5) * 33;
      ex[lane] = incl - myAx;
      if (lane == 31) ex[32] = total;
      __syncwarp();
      for (int item = lane; item < total; item += 32) {
        int lo = 0, hi = 32;  // last pair slot whose first item is <= item (slots without axes have empty ranges)
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (ex[mid] <= item) lo = mid; else hi = mid; }
        while (ex[lo + 1] <= item) lo++;  // skip empty slots that share the offset
        const HfPairCtx& X = ctxs[lo];
        s_hoist[((size_t)(threadIdx.x >> 5) * 32 + lo) * QS_HOIST_AXES + (item - ex[lo])] = qs_axis(T, X.hd, X.qf, X.oA, X.qs, item - ex[lo]);
      }
      __syncwarp();
    }
    unsigned todo = __ballot_sync(0xffffffffu, hfPair);
    const int myOff = (pass && k < np) ? A.pairTaskOff[k] : 0;
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int pCode = __shfl_sync(0xffffffffu, code, src) | (__shfl_sync(0xffffffffu, jt, src) << 5), pk = kb + src;
      const int x0 = __shfl_sync(0xffffffffu, iMinX, src), x1 = __shfl_sync(0xffffffffu, iMaxX, src);
      const int y0 = __shfl_sync(0xffffffffu, iMinY, src), y1 = __shfl_sync(0xffffffffu, iMaxY, src);
      const int pOff = __shfl_sync(0xffffffffu, myOff, src);
      const int wy = y1 - y0, nP = (x1 - x0) * wy * 2;
      const HfPairCtx& X = ctxs[src];
      const HfDev hf = X.hf;
      const f3 xf = X.xf, xs = X.xs;
      const q4 qs = X.qs;
      const bool hullTask = (pCode & 15) == NP_HPIL;
      const double rFirst = X.rFirst;
      HullDev hd;
      q4 qf;
      f3 oA;
      if (hullTask) { hd = X.hd; qf = X.qf; oA = X.oA; }
      const bool cached = pass && nP <= 64;
      const unsigned long long cachedMask = X.mask;
      // quick-separation axes of this pair (face 0 of the hull when it has uniqueAxes, then hull edge x heightfield z)
      const bool quick = hullTask && !cached && hd.nE <= 32 && hd.nF <= 32;
      int nAx = 0;
      const QsAxis* qs_use = qs_ax;
      if (quick) {
        nAx = (hd.hasAxes ? 1 : 0) + hd.nEk;  // +-copies of an edge separate exactly when the first copy does
        if (hoisted && nAx <= QS_HOIST_AXES) {
          qs_use = s_hoist + ((size_t)(threadIdx.x >> 5) * 32 + src) * QS_HOIST_AXES;  // computed before the loop
        } else {
          __syncwarp();  // the previous pair's lanes are done with the scratch
          for (int t = lane; t < nAx; t += 32) qs_ax[t] = qs_axis(T, hd, qf, oA, qs, t);
          __syncwarp();
        }
      }
      unsigned long long mask = 0ull;
      int count = 0;
      for (int base = 0; base < nP; base += 32) {
        const int pidx = base + lane;
        bool alive = false;
        int ci = 0, cj = 0, up = 0;
        if (pidx < nP) {
          const int cell = pidx >> 1;
          up = pidx & 1;
          ci = x0 + cell / wy;
          cj = y0 + cell % wy;
          if (cached) alive = (cachedMask >> pidx) & 1ull;
          else {
            // offset, bounding radius and vertices of the pillar from the table built by k_pillars_build
            const PillarRec* R = T.pillars + hf.pilOff + (((long long)ci * (hf.ny - 1) + cj) * 2 + up);
            f3 po = ld3(R->off), pv[6];
            const double pr = R->bsr;
#pragma unroll
            for (int i = 0; i < 6; i++) pv[i] = ld3(R->v[i]);
            const f3 wpo = to_world_point(xs, qs, po);
            alive = vdist(xf, wpo) < pr + rFirst;
            if (alive && !hullTask) alive = !sphere_pillar_far(X.lc, rFirst, X.margin, po, pv);
            if (alive && quick) alive = !pillar_quick_separated_pre(qs_use, nAx, pv, wpo, qs);
          }
        }
        const unsigned bits = __ballot_sync(0xffffffffu, alive);
        if (base < 64) mask |= (unsigned long long)bits << base;
        if (pass && alive) {
          const int t = pOff + count + __popc(bits & ((1u << lane) - 1u));
          if (t < A.taskCap) {
            A.taskPair[t] = pk;
            A.taskInfo[t] = pCode | (up << 4);
            A.taskCell[t] = make_int2(ci, cj);
          }
        }
        count += __popc(bits);
      }
      if (lane == src) { nt = count; if (!pass) A.pairMask[pk] = mask; }
    }
    // bucket bookkeeping: one atomic per (warp, resolver type) instead of one per pair
    const bool valid = k < np;
    const bool fits = valid && nt > 0 && (pass == 0 || myOff + nt <= A.taskCap);
    if (pass && valid && nt > 0 && !fits) atomicMax(A.overflowTasks, myOff + nt);
    const int gcode = fits ? code : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, gcode);
    const int leader = __ffs(peers) - 1;
    const unsigned lt = (1u << lane) - 1u;
    if (pass == 0) {
      if (valid) A.pairTasks[k] = nt;
      if (gcode >= 0) {
        const int total = __reduce_add_sync(peers, nt);
        if (lane == leader) atomicAdd(&A.bucketCount[gcode], total);
      }
      continue;
    }
    // exclusive prefix of nt inside the lane's type group (heightfield pairs carry several tasks each)
    int before = 0;
    if (np_pillar_code(gcode)) {
      for (unsigned m = peers; m; m &= m - 1) {  // the same trips for every lane of the group
        const int src = __ffs(m) - 1;
        const int v = __shfl_sync(peers, nt, src);
        if (src < lane) before += v;
      }
    } else {
      before = __popc(peers & lt);
    }
    if (gcode < 0) continue;
    int slot = 0;
    const int total = __reduce_add_sync(peers, nt);
    if (lane == leader) slot = A.bucketStart[gcode] + atomicAdd(&A.bucketCursor[gcode], total);
    slot = __shfl_sync(peers, slot, leader) + before;
    const int off = myOff;
    if (np_pillar_code(gcode)) {
      for (int t = 0; t < nt; t++) A.bucket[slot + t] = off + t;
    } else {
      A.taskPair[off] = k;
      A.taskInfo[off] = gcode | (jt << 5);
      A.bucket[slot] = off;
    }
  }
}

// ---- raw contact emission ----------------------------------------------------------------------------
struct RawOut {
  NpArrays A;
  int task;
  int start;
  int n;
};
__device__ __forceinline__ bool raw_alloc(RawOut& o, int m) {
  o.n = 0;
  o.start = 0;
  if (o.A.taskHit) {  // justTest task: remember whether the resolver would have created a contact, create none
    const int jt = (o.A.taskInfo[o.task] >> 5) & 1;
    o.A.taskHit[o.task] = jt ? (m > 0) : 0;
    if (jt) m = 0;
  }
  if (m > 0) {
    o.start = atomicAdd(o.A.rawCount, m);
    if (o.start + m > o.A.contactCap) { atomicMax(o.A.overflowContacts, o.start + m); m = 0; o.start = 0; }
  }
  o.A.taskCnt[o.task] = m;
  o.A.taskRaw[o.task] = o.start;
  return m > 0;
}
// "Make relative to bodies": r.add2(x, r); r.sub2(body.position, r)
__device__ __forceinline__ f3 rel_to_body(const f3& r, const f3& x, const f3& bodyPos) { return vsub(vadd(r, x), bodyPos); }
__device__ __forceinline__ void raw_put(RawOut& o, const f3& ri, const f3& rj, const f3& ni) {
  const int k = o.start + o.n++;
  o.A.rawRi[k] = st3(ri);
  o.A.rawRj[k] = st3(rj);
  o.A.rawNi[k] = st3(ni);
}

struct TaskCtx {
  int task, pair, first, second, info;
  f3 xi, xj;   // world positions of the two shapes (narrow_phase.dart:672-680)
  q4 qi, qj;
  f3 bxi, bxj; // positions of the bodies that own them ("make relative to bodies")
  ShapeDev si, sj;
};
__device__ __forceinline__ void load_task(const BodyArrays& B, const ShapeTables& T, const NpArrays& A, int task, TaskCtx& c) {
  c.task = task;
  c.pair = A.taskPair[task];
  c.info = A.taskInfo[task];
  const int a = A.p1[c.pair], b = A.p2[c.pair];
  const ShapeDev sa = T.shapes[B.shape[a]], sb = T.shapes[B.shape[b]];
  if (sa.type < sb.type) { c.first = a; c.second = b; c.si = sa; c.sj = sb; }
  else { c.first = b; c.second = a; c.si = sb; c.sj = sa; }
  c.xi = ld3(B.pos[c.first]); c.xj = ld3(B.pos[c.second]);
  c.qi = ldq(B.quat[c.first]); c.qj = ldq(B.quat[c.second]);
  c.bxi = ld3(B.bpos[np_owner(B, c.first)]); c.bxj = ld3(B.bpos[np_owner(B, c.second)]);
}

#define NP_BUCKET_LOOP(TYPE)                                                     \
  const int nb__ = (*A.nTasks <= A.taskCap) ? A.bucketCount[TYPE] : 0;          \
  for (int u__ = blockIdx.x * blockDim.x + threadIdx.x; u__ < nb__; u__ += gridDim.x * blockDim.x)
#define NP_TASK(TYPE) A.bucket[A.bucketStart[TYPE] + u__]

// sphereSphere, narrow_phase.dart:723-765: always one contact (the prologue's bounding test is the only test)
__global__ void __launch_bounds__(256) k_np_sphere_sphere(BodyArrays B, ShapeTables T, NpArrays A) {
  NP_BUCKET_LOOP(NP_SS) {
    TaskCtx c; load_task(B, T, A, NP_TASK(NP_SS), c);
    RawOut o; o.A = A; o.task = c.task;
    int m = 1;
    if ((c.info >> 5) & 1) {  // justTest has a test of its own (:735-737): Vector3.distanceSquared (vec3.dart:88-93) against (ri + rj)^2
      const double dx = W(c.xj.x) - W(c.xi.x), dy = W(c.xj.y) - W(c.xi.y), dz = W(c.xj.z) - W(c.xi.z);
      const double rs = c.si.radius + c.sj.radius;
      m = dx * dx + dy * dy + dz * dz < rs * rs ? 1 : 0;
    }
    if (!raw_alloc(o, m)) continue;
    f3 ni = vsub(c.xj, c.xi);
    vnormalize(ni);
    f3 ri = vscale(c.si.radius, ni);
    f3 rj = vscale(-c.sj.radius, ni);
    raw_put(o, rel_to_body(ri, c.xi, c.bxi), rel_to_body(rj, c.xj, c.bxj), ni);
  }
}

// spherePlane, narrow_phase.dart:766-815
__global__ void __launch_bounds__(256) k_np_sphere_plane(BodyArrays B, ShapeTables T, NpArrays A) {
  NP_BUCKET_LOOP(NP_SP) {
    TaskCtx c; load_task(B, T, A, NP_TASK(NP_SP), c);
    RawOut o; o.A = A; o.task = c.task;
    f3 z; z.x = 0.f; z.y = 0.f; z.z = 1.f;
    f3 ni = vneg(qrot(c.qj, z));
    vnormalize(ni);
    const double R = c.si.radius;
    f3 ri = vscale(R, ni);
    const f3 p2s = vsub(c.xi, c.xj);
    const f3 ortho = vscale(vdot(ni, p2s), ni);
    f3 rj = vsub(p2s, ortho);
    const bool hit = -vdot(p2s, ni) <= R;
    if (!raw_alloc(o, hit ? 1 : 0)) continue;
    raw_put(o, rel_to_body(ri, c.xi, c.bxi), rel_to_body(rj, c.xj, c.bxj), ni);
  }
}

// sphereBox, narrow_phase.dart:816-1038
__global__ void __launch_bounds__(128) k_np_sphere_box(BodyArrays B, ShapeTables T, NpArrays A) {
  NP_BUCKET_LOOP(NP_SB) {
    TaskCtx c; load_task(B, T, A, NP_TASK(NP_SB), c);
    RawOut o; o.A = A; o.task = c.task;
    const f3 xi = c.xi, xj = c.xj;
    f3 sides[6];  // Box.getSideNormals, box.dart:99-115
    {
      const float hx = c.sj.hx, hy = c.sj.hy, hz = c.sj.hz;
      f3 t;
      t.x = hx; t.y = 0.f; t.z = 0.f; sides[0] = qrot(c.qj, t);
      t.x = 0.f; t.y = hy; t.z = 0.f; sides[1] = qrot(c.qj, t);
      t.x = 0.f; t.y = 0.f; t.z = hz; sides[2] = qrot(c.qj, t);
      t.x = -hx; t.y = 0.f; t.z = 0.f; sides[3] = qrot(c.qj, t);
      t.x = 0.f; t.y = -hy; t.z = 0.f; sides[4] = qrot(c.qj, t);
      t.x = 0.f; t.y = 0.f; t.z = -hz; sides[5] = qrot(c.qj, t);
    }
    const f3 boxToSphere = vsub(xi, xj);
    const double R = c.si.radius;
    bool found = false;
    f3 outRi, outRj, outNi;
    {  // side (plane) intersections: keep the strictly smallest |dot - h - R|
      f3 sideNs, sideNs1, sideNs2;
      double sideH = 0, sideDot1 = 0, sideDot2 = 0, sideDistance = 0;
      bool have = false;
      for (int idx = 0; idx != 6; idx++) {
        f3 ns = sides[idx];
        const double h = vlen(ns);
        vnormalize(ns);
        const double dt = vdot(boxToSphere, ns);
        if (dt < h + R && dt > 0) {
          f3 ns1 = sides[(idx + 1) % 3], ns2 = sides[(idx + 2) % 3];
          const double h1 = vlen(ns1), h2 = vlen(ns2);
          vnormalize(ns1);
          vnormalize(ns2);
          const double dot1 = vdot(boxToSphere, ns1), dot2 = vdot(boxToSphere, ns2);
          if (dot1 < h1 && dot1 > -h1 && dot2 < h2 && dot2 > -h2) {
            const double dist = fabs(dt - h - R);
            if (!have || dist < sideDistance) {
              have = true; sideDistance = dist; sideDot1 = dot1; sideDot2 = dot2; sideH = h;
              sideNs = ns; sideNs1 = ns1; sideNs2 = ns2;
            }
          }
        }
      }
      if (have) {
        found = true;
        outRi = vscale(-R, sideNs);
        outNi = vneg(sideNs);
        sideNs = vscale(sideH, sideNs);
        sideNs1 = vscale(sideDot1, sideNs1);
        sideNs = vadd(sideNs, sideNs1);
        sideNs2 = vscale(sideDot2, sideNs2);
        outRj = vadd(sideNs, sideNs2);
      }
    }
    // corners
    for (int j = 0; j != 2 && !found; j++)
      for (int k = 0; k != 2 && !found; k++)
        for (int l = 0; l != 2 && !found; l++) {
          f3 rj; rj.x = rj.y = rj.z = 0.f;
          rj = j ? vadd(sides[0], rj) : vsub(rj, sides[0]);
          rj = k ? vadd(sides[1], rj) : vsub(rj, sides[1]);
          rj = l ? vadd(sides[2], rj) : vsub(rj, sides[2]);
          f3 s2c = vsub(vadd(xj, rj), xi);
          if (vlen2(s2c) < R * R) {
            found = true;
            outRi = s2c;
            vnormalize(outRi);
            outNi = outRi;
            outRi = vscale(R, outRi);
            outRj = rj;
          }
        }
    // edges
    for (int j = 0; j != 6 && !found; j++)
      for (int k = 0; k != 6 && !found; k++) {
        if (j % 3 == k % 3) continue;
        f3 edgeTangent = vcross(sides[k], sides[j]);
        vnormalize(edgeTangent);
        const f3 edgeCenter = vadd(sides[j], sides[k]);
        f3 r = vsub(vsub(xi, edgeCenter), xj);
        const double orthonorm = vdot(r, edgeTangent);
        const f3 orthogonal = vscale(orthonorm, edgeTangent);
        int l = 0;
        while (l == j % 3 || l == k % 3) l++;
        f3 dist = vsub(vsub(vsub(xi, orthogonal), edgeCenter), xj);
        const double tdist = fabs(orthonorm);
        const double ndist = vlen(dist);
        if (tdist < vlen(sides[l]) && ndist < R) {
          found = true;
          outRj = vadd(edgeCenter, orthogonal);
          outNi = vneg(dist);
          vnormalize(outNi);
          outRi = vsub(vadd(outRj, xj), xi);
          vnormalize(outRi);
          outRi = vscale(R, outRi);
        }
      }
    if (!raw_alloc(o, found ? 1 : 0)) continue;
    raw_put(o, rel_to_body(outRi, xi, c.bxi), rel_to_body(outRj, xj, c.bxj), outNi);
  }
}

// _pointInPolygon, narrow_phase.dart:2584-2617 over the world-space vertices of face f
__device__ inline bool point_in_face(const HullView& H, int f, const q4& q, const f3& x, const f3& normal, const f3& p) {
  int positive = -1;
  const int o = H.fvOff[f], N = H.fvOff[f + 1] - o;
  for (int i = 0; i != N; i++) {
    const f3 v = vadd(x, qrot(q, ld3(H.v[H.fvIdx[o + i]])));
    const f3 vn = vadd(x, qrot(q, ld3(H.v[H.fvIdx[o + (i + 1) % N]])));
    const f3 edge = vsub(vn, v);
    const f3 exn = vcross(edge, normal);
    const f3 v2p = vsub(p, v);
    const double r = vdot(exn, v2p);
    if (positive == -1 || (r > 0 && positive == 1) || (r <= 0 && positive == 0)) {
      if (positive == -1) positive = r > 0 ? 1 : 0;
      continue;
    }
    return false;
  }
  return true;
}

// sphereConvex, narrow_phase.dart:1039-1257: first hit wins (vertex, then per face: inside polygon, else its edges)
__device__ inline bool sphere_convex(const HullView& H, double R, const f3& xi, const f3& xj, const q4& qj, f3& ri, f3& rj, f3& ni) {
  for (int i = 0; i != H.nV; i++) {
    const f3 worldCorner = vadd(xj, qrot(qj, ld3(H.v[i])));
    const f3 s2c = vsub(worldCorner, xi);
    if (vlen2(s2c) < R * R) {
      ri = s2c;
      vnormalize(ri);
      ni = ri;
      ri = vscale(R, ri);
      rj = vsub(worldCorner, xj);
      return true;
    }
  }
  for (int f = 0; f != H.nF; f++) {
    const f3 worldNormal = qrot(qj, ld3(H.n[f]));
    const int o = H.fvOff[f], L = H.fvOff[f + 1] - o;
    const f3 worldPoint = vadd(qrot(qj, ld3(H.v[H.fvIdx[o]])), xj);
    const f3 closest = vadd(xi, vscale(-R, worldNormal));
    const f3 penVec = vsub(closest, worldPoint);
    const double penetration = vdot(penVec, worldNormal);
    const f3 wp2s = vsub(xi, worldPoint);
    if (penetration < 0 && vdot(wp2s, worldNormal) > 0) {
      if (point_in_face(H, f, qj, xj, worldNormal, xi)) {
        ri = vscale(-R, worldNormal);
        ni = vneg(worldNormal);
        const f3 penVec2 = vscale(-penetration, worldNormal);
        const f3 penSpherePoint = vscale(-R, worldNormal);
        rj = vsub(xi, xj);
        rj = vadd(rj, penSpherePoint);
        rj = vadd(rj, penVec2);
        return true;
      }
      for (int j = 0; j != L; j++) {
        const f3 v1 = vadd(xj, qrot(qj, ld3(H.v[H.fvIdx[o + (j + 1) % L]])));
        const f3 v2 = vadd(xj, qrot(qj, ld3(H.v[H.fvIdx[o + (j + 2) % L]])));
        const f3 edge = vsub(v2, v1);
        const f3 edgeUnit = vunit(edge);
        const f3 v1ToXi = vsub(xi, v1);
        const double dt = vdot(v1ToXi, edgeUnit);
        f3 p = vadd(vscale(dt, edgeUnit), v1);
        const f3 xiToP = vsub(p, xi);
        if (dt > 0 && dt * dt < vlen2(edge) && vlen2(xiToP) < R * R) {
          rj = vsub(p, xj);
          ni = vsub(p, xi);
          vnormalize(ni);
          ri = vscale(R, ni);
          return true;
        }
      }
    }
  }
  return false;
}

__global__ void __launch_bounds__(128) k_np_sphere_hull(BodyArrays B, ShapeTables T, NpArrays A) {
  NP_BUCKET_LOOP(NP_SH) {
    TaskCtx c; load_task(B, T, A, NP_TASK(NP_SH), c);
    RawOut o; o.A = A; o.task = c.task;
    const HullView H = hull_view(T, c.sj.hull);
    f3 ri, rj, ni;
    const bool hit = sphere_convex(H, c.si.radius, c.xi, c.xj, c.qj, ri, rj, ni);
    if (!raw_alloc(o, hit ? 1 : 0)) continue;
    raw_put(o, rel_to_body(ri, c.xi, c.bxi), rel_to_body(rj, c.xj, c.bxj), ni);
  }
}

// planeConvex / planeBox, narrow_phase.dart:1847-1915: every hull vertex on or behind the plane
__global__ void __launch_bounds__(128) k_np_plane_hull(BodyArrays B, ShapeTables T, NpArrays A) {
  NP_BUCKET_LOOP(NP_PH) {
    TaskCtx c; load_task(B, T, A, NP_TASK(NP_PH), c);
    RawOut o; o.A = A; o.task = c.task;
    const HullView H = hull_view(T, c.sj.hull);
    f3 z; z.x = 0.f; z.y = 0.f; z.z = 1.f;
    const f3 worldNormal = qrot(c.qi, z);
    int m = 0;
    for (int i = 0; i != H.nV; i++) {
      const f3 wv = vadd(c.xj, qrot(c.qj, ld3(H.v[i])));
      if (vdot(worldNormal, vsub(wv, c.xi)) <= 0.0) m++;
    }
    if (!raw_alloc(o, m)) continue;
    for (int i = 0; i != H.nV; i++) {
      const f3 wv = vadd(c.xj, qrot(c.qj, ld3(H.v[i])));
      const f3 relpos = vsub(wv, c.xi);
      const double dt = vdot(worldNormal, relpos);
      if (dt <= 0.0) {
        f3 projected = vscale(vdot(worldNormal, relpos), worldNormal);
        projected = vsub(wv, projected);
        const f3 ri = vsub(projected, c.xi);
        const f3 rj = vsub(wv, c.xj);
        raw_put(o, rel_to_body(ri, c.xi, c.bxi), rel_to_body(rj, c.xj, c.bxj), worldNormal);
      }
    }
  }
}

// ConvexPolyhedron.project, convex_polyhedron.dart:843-883
__device__ __forceinline__ void hull_project(const HullView& H, const f3& axis, const f3& pos, const q4& quat, double& mx, double& mn) {
  const f3 localAxis = qrot(qnegw(quat), axis);
  f3 zero; zero.x = zero.y = zero.z = 0.f;
  const f3 localOrigin = to_local_point(pos, quat, zero);
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

// testSepAxis, convex_polyhedron.dart:360-384
__device__ __forceinline__ bool test_sep_axis(const HullView& HA, const HullView& HB, const f3& axis, const f3& posA, const q4& quatA,
                                              const f3& posB, const q4& quatB, double& depth) {
  double maxA, minA, maxB, minB;
  hull_project(HA, axis, posA, quatA, maxA, minA);
  hull_project(HB, axis, posB, quatB, maxB, minB);
  if (maxA < minB || maxB < minA) return false;
  const double d0 = maxA - minB, d1 = maxB - minA;
  depth = d0 < d1 ? d0 : d1;
  return true;
}

// findSeparatingAxis, convex_polyhedron.dart:232-356 (face normals only for hulls with `uniqueAxes`, §5.9-9)
__device__ inline bool find_sep_axis(const HullView& HA, const HullView& HB, const f3& posA, const q4& quatA, const f3& posB, const q4& quatB,
                                     bool onlyFace0OfA, f3& target) {
  double dmin = INFINITY;
  target.x = target.y = target.z = 0.f;
  if (HA.hasAxes) {
    const int nfa = onlyFace0OfA ? 1 : HA.nF;
    for (int i = 0; i < nfa; i++) {
      const f3 n = qrot(quatA, ld3(HA.n[i]));
      double d;
      if (!test_sep_axis(HA, HB, n, posA, quatA, posB, quatB, d)) return false;
      if (d < dmin) { dmin = d; target = n; }
    }
  }
  if (HB.hasAxes) {
    for (int i = 0; i < HB.nF; i++) {
      const f3 n = qrot(quatB, ld3(HB.n[i]));
      double d;
      if (!test_sep_axis(HA, HB, n, posA, quatA, posB, quatB, d)) return false;
      if (d < dmin) { dmin = d; target = n; }
    }
  }
  for (int e0 = 0; e0 != HA.nE; e0++) {
    const f3 we0 = qrot(quatA, ld3(HA.e[e0]));
    for (int e1 = 0; e1 != HB.nE; e1++) {
      const f3 we1 = qrot(quatB, ld3(HB.e[e1]));
      f3 c = vcross(we0, we1);
      if (!valmost_zero(c)) {
        vnormalize(c);
        double d;
        if (!test_sep_axis(HA, HB, c, posA, quatA, posB, quatB, d)) return false;
        if (d < dmin) { dmin = d; target = c; }
      }
    }
  }
  const f3 deltaC = vsub(posB, posA);
  if (vdot(deltaC, target) > 0.0) target = vneg(target);
  return true;
}

// clipAgainstHull + clipFaceAgainstHull + clipFaceAgainstPlane, convex_polyhedron.dart:189-227,417-587.
// Returns the number of kept points; pts/depth hold them; nrm = world normal of A's reference face.
__device__ inline int clip_hulls(const HullView& HA, const f3& posA, const q4& quatA, const HullView& HB, const f3& posB, const q4& quatB,
                                 const f3& sep, f3* pa, f3* pb, double* depthOut, f3& nrm, bool& overflow) {
  int closestB = -1;
  double dmax = -INFINITY;
  for (int f = 0; f < HB.nF; f++) {
    const double d = vdot(qrot(quatB, ld3(HB.n[f])), sep);
    if (d > dmax) { dmax = d; closestB = f; }
  }
  if (closestB < 0) return 0;
  int nIn = 0;
  {
    const int o = HB.fvOff[closestB], L = HB.fvOff[closestB + 1] - o;
    for (int i = 0; i < L && nIn < NP_MAXPOLY; i++) pa[nIn++] = vadd(posB, qrot(quatB, ld3(HB.v[HB.fvIdx[o + i]])));
    if (L > NP_MAXPOLY) overflow = true;
  }
  int closestA = -1;
  double dmin = INFINITY;
  for (int f = 0; f < HA.nF; f++) {
    const double d = vdot(qrot(quatA, ld3(HA.n[f])), sep);
    if (d < dmin) { dmin = d; closestA = f; }
  }
  if (closestA < 0) return 0;
  const int numVerticesA = HA.fvOff[closestA + 1] - HA.fvOff[closestA];
  const int co = HA.fcOff[closestA], nConn = HA.fcOff[closestA + 1] - co;
  f3* in = pa;
  f3* out = pb;
  for (int i = 0; i < numVerticesA; i++) {
    const int otherFace = (nConn > i) ? HA.fcIdx[co + i] : 0;
    const f3 pn = qrot(quatA, ld3(HA.n[otherFace]));
    const double pc = HA.pc[otherFace] - vdot(pn, posA);
    int nOut = 0;
    if (nIn >= 2) {  // clipFaceAgainstPlane
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
  nrm = qrot(quatA, ld3(HA.n[closestA]));
  const double planeEq = HA.pc[closestA] - vdot(nrm, posA);
  int kept = 0;
  for (int i = 0; i < nIn; i++) {
    double depth = vdot(nrm, in[i]) + planeEq;
    if (depth <= -100.0) depth = -100.0;
    if (depth <= 100.0 && depth <= 1e-6) {
      const f3 p = in[i];
      out[kept] = p;  // `out` is free now; kept <= i
      depthOut[kept] = depth;
      kept++;
    }
  }
  // results live in `out`; make the caller's pa hold them
  if (out != pa)
    for (int i = 0; i < kept; i++) pa[i] = out[i];
  return kept;
}

// convexConvex, narrow_phase.dart:1981-2043 (A = first, B = second)
__device__ inline void convex_convex_emit(RawOut& o, const HullView& HA, const HullView& HB, const f3& xi, const f3& xj, const q4& qi,
                                          const q4& qj, const f3& bodyPosI, const f3& bodyPosJ, bool onlyFace0OfA, int* clipOverflow) {
  f3 sep;
  int kept = 0;
  f3 pa[NP_MAXPOLY], pb[NP_MAXPOLY];
  double depth[NP_MAXPOLY];
  f3 nrm;
  nrm.x = nrm.y = nrm.z = 0.f;
  if (!(vdist(xi, xj) > HA.bsr + HB.bsr) && find_sep_axis(HA, HB, xi, qi, xj, qj, onlyFace0OfA, sep)) {
    bool ovf = false;
    kept = clip_hulls(HA, xi, qi, HB, xj, qj, sep, pa, pb, depth, nrm, ovf);
    if (ovf) atomicExch(clipOverflow, 1);
  }
  if (!raw_alloc(o, kept)) return;
  const f3 ni = vneg(sep);
  for (int j = 0; j < kept; j++) {
    f3 q = vscale(depth[j], vneg(nrm));
    f3 ri = vadd(pa[j], q);
    f3 rj = pa[j];
    ri = vsub(ri, xi);
    rj = vsub(rj, xj);
    ri = vsub(vadd(ri, xi), bodyPosI);
    rj = vsub(vadd(rj, xj), bodyPosJ);
    raw_put(o, ri, rj, ni);
  }
}

// hulls too large for the shared-memory scratch of the tile kernel (k_sat_warp.cuh)
#define SAT_MAXF 32
#define SAT_MAXE 32
#define SAT_MAXV 24  // vertices per hull staged as doubles for the projections (box 8, 8-segment cylinder 16, pillar 6)
__device__ __forceinline__ bool sat_oversize(const HullView& HA, const HullView& HB) {
  return HA.nF > SAT_MAXF || HB.nF > SAT_MAXF || HA.nE > SAT_MAXE || HB.nE > SAT_MAXE || HA.nV > SAT_MAXV || HB.nV > SAT_MAXV;
}

__global__ void __launch_bounds__(64) k_np_hull_hull(BodyArrays B, ShapeTables T, NpArrays A, int* clipOverflow, int oversizeOnly) {
  NP_BUCKET_LOOP(NP_HH) {
    TaskCtx c; load_task(B, T, A, NP_TASK(NP_HH), c);
    RawOut o; o.A = A; o.task = c.task;
    const HullView HA = hull_view(T, c.si.hull), HB = hull_view(T, c.sj.hull);
    if (oversizeOnly && !sat_oversize(HA, HB)) continue;
    convex_convex_emit(o, HA, HB, c.xi, c.xj, c.qi, c.qj, c.bxi, c.bxj, false, clipOverflow);
  }
}

// Heightfield.getConvexTrianglePillar, heightfield.dart:330-487, into thread-local storage
struct PillarStore {
  float4 v[6];
  float4 n[5];
  double pc[5];
  float4 e[18];
  int nE;
  double bsr;
};
__constant__ int c_pillarFvOff[6] = {0, 3, 6, 10, 14, 18};
__constant__ int c_pillarLower[18] = {0, 1, 2, 5, 4, 3, 0, 2, 5, 3, 1, 0, 3, 4, 4, 5, 2, 1};
__constant__ int c_pillarUpper[18] = {0, 1, 2, 5, 4, 3, 2, 5, 3, 0, 3, 4, 1, 0, 1, 4, 5, 2};

__device__ inline void build_pillar(const ShapeTables& T, const HfDev& hf, int xi, int yi, bool upper, PillarStore& S, f3& offset,
                                    bool needEdges = true) {
  const double es = (double)hf.esize;
  const double* d = T.hfdata + hf.dataOff;
  const double h00 = d[(size_t)xi * hf.ny + yi], h10 = d[(size_t)(xi + 1) * hf.ny + yi], h01 = d[(size_t)xi * hf.ny + yi + 1],
               h11 = d[(size_t)(xi + 1) * hf.ny + yi + 1];
  const double h = (fmin(fmin(h00, h10), fmin(h01, h11)) - hf.minV) / 2 + hf.minV;
  f3 v[6];
  if (!upper) {
    offset = mk3((xi + 0.25) * es, (yi + 0.25) * es, h);
    v[0] = mk3(-0.25 * es, -0.25 * es, h00 - h);
    v[1] = mk3(0.75 * es, -0.25 * es, h10 - h);
    v[2] = mk3(-0.25 * es, 0.75 * es, h01 - h);
    v[3] = mk3(-0.25 * es, -0.25 * es, -h - 1);
    v[4] = mk3(0.75 * es, -0.25 * es, -h - 1);
    v[5] = mk3(-0.25 * es, 0.75 * es, -h - 1);
  } else {
    offset = mk3((xi + 0.75) * es, (yi + 0.75) * es, h);
    v[0] = mk3(0.25 * es, 0.25 * es, h11 - h);
    v[1] = mk3(-0.75 * es, 0.25 * es, h01 - h);
    v[2] = mk3(0.25 * es, -0.75 * es, h10 - h);
    v[3] = mk3(0.25 * es, 0.25 * es, -h - 1);
    v[4] = mk3(-0.75 * es, 0.25 * es, -h - 1);
    v[5] = mk3(0.25 * es, -0.75 * es, -h - 1);
  }
  const int* fv = upper ? c_pillarUpper : c_pillarLower;
  double max2 = 0;
  for (int i = 0; i < 6; i++) {
    S.v[i] = st3(v[i]);
    const double n2 = vlen2(v[i]);
    if (n2 > max2) max2 = n2;
  }
  S.bsr = sqrt(max2);
  S.nE = 0;
  for (int f = 0; f < 5; f++) {
    const int o = c_pillarFvOff[f], L = c_pillarFvOff[f + 1] - o;
    const f3 va = v[fv[o]], vb = v[fv[o + 1]], vc = v[fv[o + 2]];
    f3 nn = vcross(vsub(vc, vb), vsub(vb, va));
    if (!(nn.x == 0.f && nn.y == 0.f && nn.z == 0.f)) vnormalize(nn);
    nn = vneg(nn);
    S.n[f] = st3(nn);
    S.pc[f] = -vdot(nn, va);
    if (!needEdges) continue;  // sphereConvex never looks at uniqueEdges
    for (int j = 0; j < L; j++) {
      f3 e = vsub(v[fv[o + j]], v[fv[o + (j + 1) % L]]);
      vnormalize(e);
      bool found = false;
      for (int p = 0; p < S.nE; p++)
        if (valmost_eq(ld3(S.e[p]), e)) { found = true; break; }
      if (!found) S.e[S.nE++] = st3(e);
    }
  }
}
__device__ __forceinline__ HullView pillar_view(const PillarStore& S, bool upper) {
  HullView H;
  H.v = S.v; H.nV = 6;
  H.n = S.n; H.pc = S.pc; H.nF = 5;
  H.fvOff = c_pillarFvOff; H.fvIdx = upper ? c_pillarUpper : c_pillarLower;
  H.fcOff = nullptr; H.fcIdx = nullptr;  // a pillar is never the reference hull (A) of a clip
  H.e = S.e; H.nE = S.nE;
  H.ek = nullptr; H.nEk = 0; H.fk = nullptr; H.nFk = 0;
  H.hasAxes = 0;  // plain ConvexPolyhedron(): contributes no face-normal axes (§5.9-9)
  H.bsr = S.bsr;
  return H;
}

// ---- precomputed pillar table (PillarRec, world.cuh) ---------------------------------------------------
__device__ __forceinline__ bool vpm_eq(const f3& a, const f3& b) {  // numerically a == b or a == -b
  return (a.x == b.x && a.y == b.y && a.z == b.z) || (a.x == -b.x && a.y == -b.y && a.z == -b.z);
}
// one thread per (cell, lower/upper): build_pillar exactly as the per-task kernels used to, then drop the unique
// edges that are +-copies of an earlier one (each geometric edge of the prism shows up once per adjacent face, in
// opposite directions: 14 "unique" edges are 7 directions)
__global__ void __launch_bounds__(128) k_pillars_build(ShapeTables T, HfDev hf, PillarRec* __restrict__ out) {
  const long long np = (long long)(hf.nx - 1) * (hf.ny - 1) * 2;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < np; k += (long long)gridDim.x * blockDim.x) {
    const int upper = (int)(k & 1);
    const long long cell = k >> 1;
    const int xi = (int)(cell / (hf.ny - 1)), yi = (int)(cell % (hf.ny - 1));
    PillarStore S;
    f3 off;
    build_pillar(T, hf, xi, yi, upper != 0, S, off);
    PillarRec& R = out[hf.pilOff + k];
    R.off = st3(off);
    for (int i = 0; i < 6; i++) R.v[i] = S.v[i];
    for (int i = 0; i < 5; i++) { R.n[i] = S.n[i]; R.pc[i] = S.pc[i]; }
    R.bsr = S.bsr;
    int nE = 0;
    for (int i = 0; i < S.nE; i++) {
      const f3 e = ld3(S.e[i]);
      bool dup = false;
      for (int p = 0; p < i && !dup; p++) dup = vpm_eq(ld3(S.e[p]), e);
      if (!dup && nE < 9) R.e[nE++] = S.e[i];
    }
    for (int i = nE; i < 9; i++) R.e[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    R.nE = nE;
    R.pad = 0;
  }
}
__device__ __forceinline__ const PillarRec* pillar_rec(const ShapeTables& T, const HfDev& hf, int xi, int yi, bool upper) {
  return T.pillars + hf.pilOff + (((long long)xi * (hf.ny - 1) + yi) * 2 + (upper ? 1 : 0));
}
// view of a precomputed pillar; e / nE are the pruned edges (only the tile SAT kernel may use them as the edge list)
__device__ __forceinline__ HullView pillar_view_rec(const PillarRec* R, bool upper) {
  HullView H;
  H.v = R->v; H.nV = 6;
  H.n = R->n; H.pc = R->pc; H.nF = 5;
  H.fvOff = c_pillarFvOff; H.fvIdx = upper ? c_pillarUpper : c_pillarLower;
  H.fcOff = nullptr; H.fcIdx = nullptr;
  H.e = R->e; H.nE = R->nE;
  H.ek = R->e; H.nEk = R->nE; H.fk = nullptr; H.nFk = 0;
  H.hasAxes = 0;
  H.bsr = R->bsr;
  return H;
}

// sphereHeightfield, narrow_phase.dart:1365-1435: one task per (cell, lower/upper) pillar
__global__ void __launch_bounds__(64) k_np_sphere_pillar(BodyArrays B, ShapeTables T, NpArrays A) {
  NP_BUCKET_LOOP(NP_SPIL) {
    TaskCtx c; load_task(B, T, A, NP_TASK(NP_SPIL), c);
    RawOut o; o.A = A; o.task = c.task;
    const HfDev hf = T.hfs[c.sj.hf];
    const int2 cell = A.taskCell[c.task];
    const bool upper = (c.info >> 4) & 1;
    const PillarRec* R = pillar_rec(T, hf, cell.x, cell.y, upper);
    const f3 wpo = to_world_point(c.xj, c.qj, ld3(R->off));
    bool hit = false;
    f3 ri, rj, ni;
    if (vdist(c.xi, wpo) < R->bsr + c.si.bsr) {
      const HullView H = pillar_view_rec(R, upper);  // sphereConvex never looks at the edge list
      hit = sphere_convex(H, c.si.radius, c.xi, wpo, c.qj, ri, rj, ni);
    }
    if (!raw_alloc(o, hit ? 1 : 0)) continue;
    raw_put(o, rel_to_body(ri, c.xi, c.bxi), rel_to_body(rj, wpo, c.bxj), ni);
  }
}

// heightfieldConvex / boxHeightfield, narrow_phase.dart:2116-2175: convexConvex(hull, pillar, faceListA=[0])
__global__ void __launch_bounds__(64) k_np_hull_pillar(BodyArrays B, ShapeTables T, NpArrays A, int* clipOverflow, int oversizeOnly) {
  NP_BUCKET_LOOP(NP_HPIL) {
    TaskCtx c; load_task(B, T, A, NP_TASK(NP_HPIL), c);
    RawOut o; o.A = A; o.task = c.task;
    if (oversizeOnly) {
      const HullDev hd = T.hulls[c.si.hull];
      if (!(hd.nF > 32 || hd.nE > 32)) continue;
    }
    const HfDev hf = T.hfs[c.sj.hf];
    const int2 cell = A.taskCell[c.task];
    const bool upper = (c.info >> 4) & 1;
    PillarStore S;
    f3 off;
    build_pillar(T, hf, cell.x, cell.y, upper, S, off);
    const f3 wpo = to_world_point(c.xj, c.qj, off);
    const HullView HA = hull_view(T, c.si.hull);
    if (vdist(c.xi, wpo) < S.bsr + HA.bsr) {
      const HullView HB = pillar_view(S, upper);
      convex_convex_emit(o, HA, HB, c.xi, wpo, c.qi, c.qj, c.bxi, c.bxj, true, clipOverflow);
    } else {
      raw_alloc(o, 0);
    }
  }
}

// Ray.pointInTriangle, ray_class.dart:697-709
__device__ inline bool np_point_in_triangle(const f3& p, const f3& a, const f3& b, const f3& c) {
  const f3 v0 = vsub(c, a), v1 = vsub(b, a), v2 = vsub(p, a);
  const double dot00 = vdot(v0, v0), dot01 = vdot(v0, v1), dot02 = vdot(v0, v2), dot11 = vdot(v1, v1), dot12 = vdot(v1, v2);
  const double u = dot11 * dot02 - dot01 * dot12;
  const double v = dot00 * dot12 - dot01 * dot02;
  return u >= 0 && v >= 0 && (u + v) < (dot00 * dot11 - dot01 * dot01);
}

// ---- Particle resolvers -------------------------------------------------------------------
// load_task order: `first` = the sphere / plane / hull / heightfield (lower ShapeType), `second` = the particle.
// sphereParticle :1258-1294 and planeParticle :1805-1846: one thread per task.
__global__ void __launch_bounds__(128) k_np_particle_simple(BodyArrays B, ShapeTables T, NpArrays A) {
  {
    NP_BUCKET_LOOP(NP_SPT) {
      TaskCtx c; load_task(B, T, A, NP_TASK(NP_SPT), c);
      RawOut o; o.A = A; o.task = c.task;
      f3 normal = vsub(c.xj, c.xi);  // particle - sphere
      const double lengthSquared = vlen2(normal);
      const bool hit = lengthSquared <= c.si.radius * c.si.radius;
      if (!raw_alloc(o, hit ? 1 : 0)) continue;
      vnormalize(normal);
      f3 zero; zero.x = zero.y = zero.z = 0.f;
      raw_put(o, zero, vscale(c.si.radius, normal), vneg(normal));  // ri = particle centre, rj on the sphere, ni = -normal
    }
  }
  {
    NP_BUCKET_LOOP(NP_PPT) {
      TaskCtx c; load_task(B, T, A, NP_TASK(NP_PPT), c);
      RawOut o; o.A = A; o.task = c.task;
      f3 up; up.x = 0.f; up.y = 0.f; up.z = 1.f;
      const f3 normal = qrot(ldq(B.bquat[np_owner(B, c.first)]), up);  // bj.quaternion / bj.position: the BODY's pose (:1821-1823)
      const f3 relpos = vsub(c.xj, c.bxi);
      const bool hit = vdot(normal, relpos) <= 0.0;
      if (!raw_alloc(o, hit ? 1 : 0)) continue;
      f3 projected = vscale(vdot(normal, c.xj), normal);
      projected = vsub(c.xj, projected);
      f3 zero; zero.x = zero.y = zero.z = 0.f;
      raw_put(o, zero, projected, vneg(normal));
    }
  }
}

// ConvexPolyhedron.pointIsInside, convex_polyhedron.dart:760-787 (getAveragePointLocal :712-720)
__device__ inline bool hull_point_inside(const HullView& H, const f3& p) {
  f3 pointInside; pointInside.x = pointInside.y = pointInside.z = 0.f;
  for (int i = 0; i < H.nV; i++) pointInside = vadd(pointInside, ld3(H.v[i]));
  pointInside = vscale(1.0 / (double)H.nV, pointInside);
  for (int i = 0; i < H.nF; i++) {
    const f3 n = ld3(H.n[i]);
    const f3 v = ld3(H.v[H.fvIdx[H.fvOff[i]]]);
    const double r1 = vdot(n, vsub(p, v));
    const double r2 = vdot(n, vsub(pointInside, v));
    if ((r1 < 0 && r2 > 0) || (r1 > 0 && r2 < 0)) return false;
  }
  return true;
}

// particleConvex :2179-2264 / boxParticle :1731 / heightfieldParticle :2343-2444 in three launches over the NP_HPT and
// NP_PILPT buckets. The reference measures the penetration against worldVertices / worldFaceNormals that it computes
// on the FIRST penetration a hull object ever sees and never refreshes (:2207-2212), and tasks are resolved in pair
// order, so: PHASE 0 tests pointIsInside (stateless, local frame) and lets the lowest penetrating task of a not yet
// frozen target claim it; PHASE 1 freezes the claimed targets at their claimant's pose; PHASE 2 emits the contacts
// against the frozen poses and clears the claims.
template <int PHASE>
__global__ void __launch_bounds__(128) k_np_particle_hull(BodyArrays B, ShapeTables T, NpArrays A) {
#pragma unroll 1
  for (int pil = 0; pil < 2; pil++) {
    const int TYPE = pil ? NP_PILPT : NP_HPT;
    NP_BUCKET_LOOP(TYPE) {
      TaskCtx c; load_task(B, T, A, NP_TASK(TYPE), c);
      HullView H;
      f3 xh;       // hull position: the body's, or the pillar offset in the world frame
      int target;
      if (pil) {
        const HfDev hf = T.hfs[c.si.hf];
        const int2 cell = A.taskCell[c.task];
        const bool upper = (c.info >> 4) & 1;
        const PillarRec* R = pillar_rec(T, hf, cell.x, cell.y, upper);
        xh = to_world_point(c.xi, c.qi, ld3(R->off));
        H = pillar_view_rec(R, upper);
        target = T.nShapes + (int)(R - T.pillars);
      } else {
        H = hull_view(T, c.si.hull);
        xh = c.xi;
        target = B.shape[c.first];
      }
      const f3 xp = c.xj;  // the particle
      RawOut o; o.A = A; o.task = c.task;
      if (PHASE == 0) {
        // the pillar bounding gate (:2421) was applied when the task was created
        const bool inside = hull_point_inside(H, qrot(qconj(c.qi), vsub(xp, xh)));
        A.taskCnt[c.task] = inside ? 1 : 0;  // provisional; PHASE 2 writes the final count
        if (inside && !T.pcFrozen[target]) atomicMin(&T.pcFreezeTask[target], c.task);
        continue;
      }
      if (A.taskCnt[c.task] == 0) { if (PHASE == 2) raw_alloc(o, 0); continue; }
      if (PHASE == 1) {
        if (!T.pcFrozen[target] && T.pcFreezeTask[target] == c.task) {
          T.pcPos[target] = st3(xh);
          const q4 q = c.qi;
          T.pcQuat[target] = make_float4(q.x, q.y, q.z, q.w);
          __threadfence();
          T.pcFrozen[target] = 1;
        }
        continue;
      }
      // PHASE 2
      if (T.pcFreezeTask[target] == c.task) T.pcFreezeTask[target] = 0x7f7f7f7f;  // the memset pattern of cannon_world_set_shapes
      const f3 fpos = ld3(T.pcPos[target]);
      const q4 fq = ldq(T.pcQuat[target]);
      int penetratedFaceIndex = -1;
      double minPenetration = 0.0;
      f3 penetratedFaceNormal; penetratedFaceNormal.x = penetratedFaceNormal.y = penetratedFaceNormal.z = 0.f;
      for (int i = 0; i < H.nF; i++) {
        const f3 verts = vadd(fpos, qrot(fq, ld3(H.v[H.fvIdx[H.fvOff[i]]])));  // computeWorldVertices :599-600
        const f3 normal = qrot(fq, ld3(H.n[i]));                                // computeWorldFaceNormals :642
        const double penetration = -vdot(normal, vsub(xp, verts));
        if (penetratedFaceIndex < 0 || fabs(penetration) < fabs(minPenetration)) {
          minPenetration = penetration;
          penetratedFaceIndex = i;
          penetratedFaceNormal = normal;
        }
      }
      if (!raw_alloc(o, penetratedFaceIndex >= 0 ? 1 : 0)) continue;
      f3 wpv = vscale(minPenetration, penetratedFaceNormal);
      wpv = vadd(wpv, xp);
      wpv = vsub(wpv, xh);
      f3 rj = qrot(c.qi, wpv);  // :2246 rotates the world-frame vector once more
      f3 ri; ri.x = ri.y = ri.z = 0.f;
      ri = rel_to_body(ri, xp, c.bxj);
      rj = rel_to_body(rj, xh, c.bxi);
      raw_put(o, ri, rj, vneg(penetratedFaceNormal));
    }
  }
}

// raw pool -> canonical order + createContactEquation / createFrictionEquationsFromContact parameters
struct NpWorld {
  double dt;
  double gnorm;   // (frictionGravity ?? gravity).length
  cannon_contact_material defaultCm;
};

__global__ void __launch_bounds__(256) k_np_finalize(BodyArrays B, ShapeTables T, NpArrays A, ContactArrays C, NpWorld Wd) {
  const int nt = min(*A.nTasks, A.taskCap);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
    const int m = A.taskCnt[t];
    if (m == 0) continue;
    const int dst = A.taskOff[t], src = A.taskRaw[t];
    if (dst + m > A.contactCap) continue;
    const int pair = A.taskPair[t];
    const int a = A.p1[pair], b = A.p2[pair];
    const ShapeDev sa = T.shapes[B.shape[a]], sb = T.shapes[B.shape[b]];
    int first, second;
    np_order(a, b, sa.type, sb.type, first, second);
    const ShapeDev s1 = (first == a) ? sa : sb, s2 = (first == a) ? sb : sa;
    int matA = B.material[first], matB = B.material[second];
    const cannon_contact_material* cm = &Wd.defaultCm;
    if (matA >= 0 && matB >= 0) {
      const int idx = T.cmTable[matA * T.nMat + matB];
      if (idx >= 0) cm = &T.cms[idx];
    }
    int fricA = matA, fricB = matB;
    if (sa.material >= 0 || sb.material >= 0) {
      // Shape.material (see cannon_shape_desc.material): the pair's contact material prefers the shapes' (narrow_phase.dart:
      // 692-696); createContactEquation takes shape ?? body material in resolver order, a pillar convex has none (:517-518);
      // the friction takes the shapes in PAIR order (c.si = rsi) with the bodies in RESOLVER order (c.bi) (:533-542)
      if (sa.material >= 0 && sb.material >= 0) {
        const int idx = T.cmTable[sa.material * T.nMat + sb.material];
        if (idx >= 0) cm = &T.cms[idx];
      }
      const int m1 = s1.type == CANNON_SHAPE_HEIGHTFIELD ? -1 : s1.material, m2 = s2.type == CANNON_SHAPE_HEIGHTFIELD ? -1 : s2.material;
      const bool pf = s2.type == CANNON_SHAPE_PARTICLE;              // c.bi = the particle's body
      const int cbi = pf ? second : first, cbj = pf ? first : second;
      fricA = sa.material >= 0 ? sa.material : B.material[cbi];
      fricB = sb.material >= 0 ? sb.material : B.material[cbj];
      if (m1 >= 0) matA = m1;
      if (m2 >= 0) matB = m2;
    }
    const bool cr2 = (s2.type == CANNON_SHAPE_HEIGHTFIELD) ? true : (s2.collisionResponse != 0);  // pillar hulls default to true
    const int enabled = ((B.flags[first] & BF_COLLISION_RESPONSE) && (B.flags[second] & BF_COLLISION_RESPONSE) && s1.collisionResponse && cr2) ? 1 : 0;
    double restitution = cm->restitution;
    double friction = cm->friction;
    if (matA >= 0 && matB >= 0) {
      const double ra = T.matRestitution[matA], rb = T.matRestitution[matB];
      if (ra >= 0 && rb >= 0) restitution = ra * rb;
    }
    if (fricA >= 0 && fricB >= 0) {
      const double fa = T.matFriction[fricA], fbb = T.matFriction[fricB];
      if (fa >= 0 && fbb >= 0) friction = fa * fbb;
    }
    const double h = Wd.dt;
    double k = cm->contact_equation_stiffness, d = cm->contact_equation_relaxation;
    const double ca = 4.0 / (h * (1 + 4 * d)), cb = 4.0 * d / (1 + 4 * d), ceps = 4.0 / (h * h * k * (1 + 4 * d));
    k = cm->friction_equation_stiffness; d = cm->friction_equation_relaxation;
    const double fb = 4.0 * d / (1 + 4 * d), feps = 4.0 / (h * h * k * (1 + 4 * d));
    double slip = 0.0;
    if (friction > 0) {
      const double mug = friction * Wd.gnorm;
      double reducedMass = B.invMass[first] + B.invMass[second];
      if (reducedMass > 0) reducedMass = 1 / reducedMass;
      slip = mug * reducedMass;
    }
    // the particle resolvers hand createContactEquation the particle's body first (narrow_phase.dart:1280,1829,2239)
    const bool particleFirst = s2.type == CANNON_SHAPE_PARTICLE;
    for (int q = 0; q < m; q++) {
      const int o = dst + q;
      C.bi[o] = np_owner(B, particleFirst ? second : first); C.bj[o] = np_owner(B, particleFirst ? first : second);
      C.ri[o] = A.rawRi[src + q]; C.rj[o] = A.rawRj[src + q]; C.ni[o] = A.rawNi[src + q];
      C.rest[o] = restitution; C.mu[o] = friction; C.slip[o] = slip;
      C.ca[o] = ca; C.cb[o] = cb; C.ceps[o] = ceps; C.fb[o] = fb; C.feps[o] = feps;
      C.enabled[o] = enabled;
      C.row[o] = -1;
      C.task[o] = t;
    }
  }
}

__global__ void __launch_bounds__(256) k_np_per_pair(NpArrays A, int* __restrict__ perPair) {
  const int np = *A.nPairs;
  const int nt = *A.nTasks;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < np; k += gridDim.x * blockDim.x) {
    const int t0 = A.pairTaskOff[k];
    const int t1 = t0 + A.pairTasks[k];
    int s = 0;
    for (int t = t0; t < t1 && t < nt; t++) s += A.taskCnt[t];
    perPair[k] = s;
  }
}

// ---- Trimesh -------------------------------------------------------------------------------
// sphereTrimesh, narrow_phase.dart:1438-1691 as the Dart port runs it: every triangle in index order (the octree query of
// :1480 is not used) and per corner j a vertex test, an edge test and the triangle-face test (:1573-1603 sits inside the
// j loop: a face contact is reported three times). EMIT = false counts, EMIT = true writes (same walk, same order).
template <bool EMIT>
__device__ inline int sphere_trimesh(const ShapeTables& T, const TrimeshDev& m, const TaskCtx& c, RawOut& o) {
  const f3 xi = c.xi, xj = c.xj;
  const q4 qj = c.qj;
  const double R = c.si.radius;
  const f3 local = to_local_point(xj, qj, xi);
  const double radiusSquared = R * R;
  const float4* V = T.tmVerts + m.vOff;
  const int* I = T.tmIdx + m.iOff;
  int n = 0;
  auto emit_local = [&](f3 tmp) {  // :1549-1566 / :1586-1601
    if (EMIT) {
      f3 ni = vsub(tmp, local);
      vnormalize(ni);
      f3 ri = vscale(R, ni);
      ri = vadd(ri, xi);
      ri = vsub(ri, c.bxi);
      tmp = to_world_point(xj, qj, tmp);
      const f3 rj = vsub(tmp, c.bxj);
      ni = qrot(qj, ni);
      ri = qrot(qj, ri);
      raw_put(o, ri, rj, ni);
    }
    n++;
  };
  for (int i = 0; i < m.nT; i++) {
    const f3 va = ld3(V[I[i * 3]]), vb = ld3(V[I[i * 3 + 1]]), vc = ld3(V[I[i * 3 + 2]]);
    const f3 normal = ld3(T.tmNormals[m.iOff / 3 + i]);
#pragma unroll 1
    for (int j = 0; j < 3; j++) {
      const f3 A = j == 0 ? va : (j == 1 ? vb : vc), Bv = j == 0 ? vb : (j == 1 ? vc : va);
      {
        f3 relpos = vsub(A, local);
        if (vlen2(relpos) <= radiusSquared) {
          if (EMIT) {
            const f3 v = to_world_point(xj, qj, A);
            relpos = vsub(v, xi);
            f3 ni = relpos;
            vnormalize(ni);
            f3 ri = vscale(R, ni);
            ri = vadd(ri, xi);
            ri = vsub(ri, c.bxi);
            raw_put(o, ri, vsub(v, c.bxj), ni);
          }
          n++;
        }
        const f3 edgeVector = vsub(Bv, A);
        f3 tmp = vsub(local, Bv);
        const double positionAlongEdgeB = vdot(tmp, edgeVector);
        tmp = vsub(local, A);
        double positionAlongEdgeA = vdot(tmp, edgeVector);
        if (positionAlongEdgeA > 0 && positionAlongEdgeB < 0) {
          f3 unit = edgeVector;
          vnormalize(unit);
          positionAlongEdgeA = vdot(tmp, unit);
          tmp = vscale(positionAlongEdgeA, unit);
          tmp = vadd(tmp, A);
          if (vdist(tmp, local) < R) emit_local(tmp);
        }
      }
      {
        f3 tmp = vsub(local, va);
        double dist = vdot(tmp, normal);
        tmp = vscale(dist, normal);
        tmp = vsub(local, tmp);
        dist = vdist(tmp, local);
        if (np_point_in_triangle(tmp, va, vb, vc) && dist < R) emit_local(tmp);
      }
    }
  }
  return n;
}
// sphereTrimesh / planeTrimesh (narrow_phase.dart:1438,1916): one thread per task, count then emit
__global__ void __launch_bounds__(128) k_np_trimesh(BodyArrays B, ShapeTables T, NpArrays A) {
  {
    NP_BUCKET_LOOP(NP_STM) {
      TaskCtx c; load_task(B, T, A, NP_TASK(NP_STM), c);
      RawOut o; o.A = A; o.task = c.task;
      const TrimeshDev m = T.tms[c.sj.tm];
      const int n = sphere_trimesh<false>(T, m, c, o);
      if (!raw_alloc(o, n)) continue;
      sphere_trimesh<true>(T, m, c, o);
    }
  }
  {
    NP_BUCKET_LOOP(NP_PTM) {
      TaskCtx c; load_task(B, T, A, NP_TASK(NP_PTM), c);
      RawOut o; o.A = A; o.task = c.task;
      const TrimeshDev m = T.tms[c.sj.tm];
      f3 up; up.x = 0.f; up.y = 0.f; up.z = 1.f;
      const f3 normal = qrot(c.qi, up);
      int n = 0;
      for (int i = 0; i < m.nV; i++) {
        const f3 v = to_world_point(c.xj, c.qj, ld3(T.tmVerts[m.vOff + i]));
        if (vdot(normal, vsub(v, c.xi)) <= 0.0) n++;
      }
      if (!raw_alloc(o, n)) continue;
      for (int i = 0; i < m.nV; i++) {
        const f3 v = to_world_point(c.xj, c.qj, ld3(T.tmVerts[m.vOff + i]));
        const f3 relpos = vsub(v, c.xi);
        if (vdot(normal, relpos) <= 0.0) {
          f3 projected = vscale(vdot(relpos, normal), normal);
          projected = vsub(v, projected);
          raw_put(o, vsub(projected, c.bxi), vsub(v, c.bxj), normal);
        }
      }
    }
  }
}

// ---- compound bodies (cannon_world_set_body_shapes) ------------------------------------------
// The narrowphase runs on shape instances ("proxies"): k_proxies gives every instance its world pose
// (bi.quaternion.vmult(shapeOffsets[i]) + bi.position, bi.quaternion * shapeOrientations[i]; narrow_phase.dart:671-680) and
// a copy of the body attributes the resolvers read; k_pp_count / k_pp_fill expand the body pairs into instance pairs in
// the reference's loop order (pair, i over bi.shapes, j over bj.shapes; :669-676). Everything downstream is unchanged:
// tasks, resolvers and the canonical contact order work on the instance pairs, k_np_finalize maps back to bodies.
struct ProxyArrays {
  const int* instBody;
  float4 *pos, *quat;
  int *type, *flags, *material;
  double* invMass;
  int nInst;
};
__global__ void __launch_bounds__(256) k_proxies(BodyArrays B, ShapeTables T, ProxyArrays X) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < X.nInst; k += gridDim.x * blockDim.x) {
    const int b = X.instBody[k];
    const q4 q = ldq(B.quat[b]);
    X.pos[k] = st3(vadd(qrot(q, ld3(T.instOff[k])), ld3(B.pos[b])));
    const q4 o = qmul(q, ldq(T.instQuat[k]));
    X.quat[k] = make_float4(o.x, o.y, o.z, o.w);
    X.type[k] = B.type[b]; X.flags[k] = B.flags[b]; X.material[k] = B.material[b]; X.invMass[k] = B.invMass[b];
  }
}
__global__ void __launch_bounds__(256) k_pp_count(const int* __restrict__ p1, const int* __restrict__ p2, const int* __restrict__ nPairs, int pairCap,
                                                  const int* __restrict__ instFirst, int* __restrict__ cnt) {
  const int np = min(*nPairs, pairCap);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < pairCap; k += gridDim.x * blockDim.x) {
    int c = 0;
    if (k < np) { const int a = p1[k], b = p2[k]; c = (instFirst[a + 1] - instFirst[a]) * (instFirst[b + 1] - instFirst[b]); }
    cnt[k] = c;
  }
}
__global__ void __launch_bounds__(256) k_pp_fill(const int* __restrict__ p1, const int* __restrict__ p2, const int* __restrict__ nPairs, int pairCap,
                                                 const int* __restrict__ instFirst, const int* __restrict__ off, int* __restrict__ pp1, int* __restrict__ pp2,
                                                 int ppCap, int* __restrict__ overflow) {
  const int np = min(*nPairs, pairCap);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < np; k += gridDim.x * blockDim.x) {
    const int a = p1[k], b = p2[k];
    const int a0 = instFirst[a], a1 = instFirst[a + 1], b0 = instFirst[b], b1 = instFirst[b + 1];
    int o = off[k];
    if (o + (a1 - a0) * (b1 - b0) > ppCap) { atomicMax(overflow, o + (a1 - a0) * (b1 - b0)); continue; }
    for (int i = a0; i < a1; i++)
      for (int j = b0; j < b1; j++) { pp1[o] = i; pp2[o] = j; o++; }
  }
}
// contacts per BODY pair from the per-instance-pair counts (cannon_narrowphase_contacts' per_pair_count)
__global__ void __launch_bounds__(256) k_pp_per_pair(const int* __restrict__ perProxy, const int* __restrict__ off, const int* __restrict__ cnt, int nBodyPairs,
                                                     int* __restrict__ out) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nBodyPairs; k += gridDim.x * blockDim.x) {
    int s = 0;
    for (int t = off[k]; t < off[k] + cnt[k]; t++) s += perProxy[t];
    out[k] = s;
  }
}

// ---- contact events ---------------------------------------------------------------------
// bodyOverlapKeeper (overlap_keeper.dart) as two open-addressing hash sets of 64-bit body-pair keys: the pairs that own
// at least one ContactEquation this step (world_class.dart:606) are inserted into the current set, the difference
// against the previous step's set in both directions gives beginContact / endContact (getDiff, overlap_keeper.dart:48-82).
// The lists leave the device unordered; cannon_world_get_contact_events sorts the few events by key on the host.
#define EV_EMPTY 0xffffffffffffffffull
struct EvArrays {
  unsigned long long *keysCur, *keysPrev, *tabCur, *tabPrev, *begin, *end;
  int* cnt;  // [0] pairs in contact now, [1] in the previous step, [2] begin events, [3] end events
  unsigned mask;
  int cap;
};
__device__ __forceinline__ unsigned ev_hash(unsigned long long k, unsigned mask) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33;
  return (unsigned)k & mask;
}
__device__ __forceinline__ bool ev_find(const unsigned long long* __restrict__ tab, unsigned mask, unsigned long long key) {
  for (unsigned h = ev_hash(key, mask);; h = (h + 1) & mask) {
    const unsigned long long v = tab[h];
    if (v == key) return true;
    if (v == EV_EMPTY) return false;
  }
}
__global__ void k_ev_begin(EvArrays E) { if (threadIdx.x == 0) { E.cnt[0] = 0; E.cnt[2] = 0; E.cnt[3] = 0; } }
__global__ void __launch_bounds__(256) k_ev_collect(NpArrays A, EvArrays E, const int* __restrict__ owner) {
  const int np = *A.nPairs, nt = min(*A.nTasks, A.taskCap);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < np; k += gridDim.x * blockDim.x) {
    const int t0 = A.pairTaskOff[k], t1 = min(t0 + A.pairTasks[k], nt);
    bool any = false;
    for (int t = t0; t < t1 && !any; t++) any = A.taskCnt[t] > 0 || (A.taskHit && A.taskHit[t] != 0);
    if (!any) continue;
    const int a = owner ? owner[A.p1[k]] : A.p1[k], b = owner ? owner[A.p2[k]] : A.p2[k];  // bodyOverlapKeeper: body ids
    const unsigned long long key = ((unsigned long long)(unsigned)min(a, b) << 32) | (unsigned long long)(unsigned)max(a, b);
    for (unsigned h = ev_hash(key, E.mask);; h = (h + 1) & E.mask) {
      const unsigned long long prev = atomicCAS(&E.tabCur[h], EV_EMPTY, key);
      if (prev == EV_EMPTY) {  // first insertion of this pair (OverlapKeeper.set ignores duplicates)
        const int idx = atomicAdd(&E.cnt[0], 1);
        if (idx < E.cap) E.keysCur[idx] = key;
        break;
      }
      if (prev == key) break;
    }
  }
}
__global__ void __launch_bounds__(256) k_ev_diff(EvArrays E) {
  const int nCur = min(E.cnt[0], E.cap), nPrev = min(E.cnt[1], E.cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nCur + nPrev; i += gridDim.x * blockDim.x) {
    if (i < nCur) {
      const unsigned long long key = E.keysCur[i];
      if (!ev_find(E.tabPrev, E.mask, key)) { const int o = atomicAdd(&E.cnt[2], 1); if (o < E.cap) E.begin[o] = key; }
    } else {
      const unsigned long long key = E.keysPrev[i - nCur];
      if (!ev_find(E.tabCur, E.mask, key)) { const int o = atomicAdd(&E.cnt[3], 1); if (o < E.cap) E.end[o] = key; }
    }
  }
}
// OverlapKeeper.tick for the next step: current becomes previous (the tables are copied by the host code)
__global__ void __launch_bounds__(256) k_ev_roll(EvArrays E) {
  const int nCur = min(E.cnt[0], E.cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nCur; i += gridDim.x * blockDim.x) E.keysPrev[i] = E.keysCur[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) E.cnt[1] = nCur;
}
