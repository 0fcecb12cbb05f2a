// k_raycast.cuh — World.raycastClosest / raycastAny / raycastAll and Broadphase.aabbQuery on device-resident bodies
// (lib/world/world_class.dart:248-277, lib/collision/ray_class.dart:175-283).
//
// One warp per ray. The lanes walk the bodies (lane, lane + 32, ...): Body.updateAABB on the current pose, the ray-AABB
// overlap of NaiveBroadphase.aabbQuery (naive_broadphase.dart:39-56, aabb.dart:131-147), the filters of Ray.intersectBody
// (ray_class.dart:201-225), the bounding-sphere rejection of _intersectShape (:270-283) and the shape's intersection routine
// (_intersectSphere :411-458, _intersectPlane :289-326, _intersectBox :285-287, _intersectConvex :460-553) - every
// expression evaluated like the Dart VM does (f64 on f32-stored vectors, every Vector3 temporary rounded to float).
// The reference visits candidates one after the other; here every intersection carries its place in that sequence
// (key = body index << 16 | ordinal of the report inside the body) and the warp reduces:
//   closest: smallest distance, the earliest report winning ties (`distance < result.distance || !hasHit`, :672);
//   any:     the earliest report (the reference stops there);
//   all:     every report, appended to a list the host sorts by (ray, key) = the callback sequence;
//   RaycastResult.hitFaceIndex is written by EVERY report before the mode is looked at (:664): the latest report's face.
// Candidate order is NaiveBroadphase's (body index) for every broadphase kind; heightfield / trimesh rays are refused by the
// entry point (documented in include/cannon_cuda.h).
#pragma once
#include "k_narrowphase.cuh"

struct RayArgs {
  int nRays, nBodies;
  const float* from;  // 3 floats per ray
  const float* to;
  int mode, skipBackfaces, mask, group, checkCollisionResponse;
  // per ray (closest / any; also the final RaycastResult of `all`)
  unsigned char* hasHit;
  int *body, *face;
  double* dist;
  float4 *point, *normal;
  // RayMode.all: every report
  int* allCount;
  int allCap;
  int *allRay, *allBody, *allFace, *allInst;
  int* inst;  // per ray: position of the hit shape in Body.shapes
  unsigned long long* allKey;
  double* allDist;
  float4 *allPoint, *allNormal;
};

struct RayLane {
  f3 from, to, dir;
  bool has;
  double dist;
  unsigned long long key, lastKey;
  f3 point, normal;
  int body, face, lastFace;
  int inst, curInst;  // shape of the held report / shape being intersected (position in Body.shapes)
  bool any;  // lastKey is valid
};

// Quaternion.multiply2, quaternion.dart:65-83
__device__ __forceinline__ q4 ray_qmul(const q4& a, const q4& b) {
  const double ax = W(a.x), ay = W(a.y), az = W(a.z), aw = W(a.w), bx = W(b.x), by = W(b.y), bz = W(b.z), bw = W(b.w);
  q4 t;
  t.x = (float)(ax * bw + aw * bx + ay * bz - az * by);
  t.y = (float)(ay * bw + aw * by + az * bx - ax * bz);
  t.z = (float)(az * bw + aw * bz + ax * by - ay * bx);
  t.w = (float)(aw * bw - ax * bx - ay * by - az * bz);
  return t;
}

// Ray._reportIntersection, ray_class.dart:655-693
__device__ inline void ray_report(const RayArgs& A, RayLane& L, int ray, const f3& normal, const f3& hit, int body, int face, int& ordinal) {
  const double distance = vdist(L.from, hit);
  if (A.skipBackfaces && vdot(normal, L.dir) > 0) return;
  const unsigned long long key = ((unsigned long long)(unsigned)body << 16) | (unsigned)min(ordinal, 0xffff);
  ordinal++;
  L.lastKey = key; L.lastFace = face; L.any = true;  // a lane sees its bodies in ascending order: the latest report so far
  if (A.mode == CANNON_RAY_ALL) {
    const int k = atomicAdd(A.allCount, 1);
    if (k < A.allCap) {
      A.allRay[k] = ray; A.allBody[k] = body; A.allFace[k] = face; A.allKey[k] = key; A.allDist[k] = distance; A.allInst[k] = L.curInst;
      A.allPoint[k] = st3(hit); A.allNormal[k] = st3(normal);
    }
  }
  bool take;
  if (A.mode == CANNON_RAY_CLOSEST) take = !L.has || distance < L.dist;
  else if (A.mode == CANNON_RAY_ANY) take = !L.has;  // the earliest report of this lane
  else take = true;                                  // all: RaycastResult holds the latest report
  if (take) { L.has = true; L.dist = distance; L.key = key; L.point = hit; L.normal = normal; L.body = body; L.face = face; L.inst = L.curInst; }
}

// Ray.pointInTriangle, ray_class.dart:696-708
__device__ inline bool ray_point_in_triangle(const f3& p, const f3& a, const f3& b, const f3& c) {
  const f3 v0 = vsub(c, a), v1 = vsub(b, a), v2 = vsub(p, a);
  const double dot00 = vdot(v0, v0), dot01 = vdot(v0, v1), dot02 = vdot(v0, v2), dot11 = vdot(v1, v1), dot12 = vdot(v1, v2);
  const double u = dot11 * dot02 - dot01 * dot12;
  const double v = dot00 * dot12 - dot01 * dot02;
  return u >= 0 && v >= 0 && (u + v) < (dot00 * dot11 - dot01 * dot01);
}

__device__ inline void ray_sphere(const RayArgs& A, RayLane& L, int ray, double rad, const f3& position, int body, int& ord) {
  const f3 &from = L.from, &to = L.to;
  const double dx = W(to.x) - W(from.x), dy = W(to.y) - W(from.y), dz = W(to.z) - W(from.z);
  const double fx = W(from.x) - W(position.x), fy = W(from.y) - W(position.y), fz = W(from.z) - W(position.z);
  const double a = dx * dx + dy * dy + dz * dz;
  const double b = 2 * (dx * fx + dy * fy + dz * fz);
  const double c = fx * fx + fy * fy + fz * fz - rad * rad;
  const double delta = b * b - 4 * a * c;
  if (delta < 0) return;
  if (delta == 0) {
    const f3 p = vlerp(from, to, delta);
    f3 normal = vsub(p, position);
    vnormalize(normal);
    ray_report(A, L, ray, normal, p, body, -1, ord);
  } else {
    const double d1 = (-b - sqrt(delta)) / (2 * a), d2 = (-b + sqrt(delta)) / (2 * a);
    if (d1 >= 0 && d1 <= 1) {
      const f3 p = vlerp(from, to, d1);
      f3 normal = vsub(p, position);
      vnormalize(normal);
      ray_report(A, L, ray, normal, p, body, -1, ord);
    }
    if (A.mode == CANNON_RAY_ANY && L.has) return;
    if (d2 >= 0 && d2 <= 1) {
      const f3 p = vlerp(from, to, d2);
      f3 normal = vsub(p, position);
      vnormalize(normal);
      ray_report(A, L, ray, normal, p, body, -1, ord);
    }
  }
}

__device__ inline void ray_plane(const RayArgs& A, RayLane& L, int ray, const q4& quat, const f3& position, int body, int& ord) {
  const f3 &from = L.from, &to = L.to;
  const f3 worldNormal = qrot(quat, mk3(0.0, 0.0, 1.0));
  f3 len = vsub(from, position);
  const double planeToFrom = vdot(len, worldNormal);
  len = vsub(to, position);
  const double planeToTo = vdot(len, worldNormal);
  if (planeToFrom * planeToTo > 0) return;
  if (vdist(from, to) < planeToFrom) return;
  const double nDotDir = vdot(worldNormal, L.dir);
  if (fabs(nDotDir) < 0.0001) return;
  const f3 planePointToFrom = vsub(from, position);
  const double t = -vdot(worldNormal, planePointToFrom) / nDotDir;
  const f3 hit = vadd(from, vscale(t, L.dir));
  ray_report(A, L, ray, worldNormal, hit, body, -1, ord);
}

__device__ inline void ray_convex(const RayArgs& A, RayLane& L, int ray, const HullView& H, const q4& q, const f3& x, int body, int& ord) {
  const f3 &from = L.from, &to = L.to;
  const double fromToDistance = vdist(from, to);
  for (int fi = 0; fi < H.nF; fi++) {
    if (A.mode == CANNON_RAY_ANY && L.has) return;
    const int o = H.fvOff[fi], nv = H.fvOff[fi + 1] - o;
    f3 vector = ld3(H.v[H.fvIdx[o]]);
    vector = qrot(q, vector);
    vector = vadd(vector, x);
    vector = vsub(vector, from);
    const f3 normal = qrot(q, ld3(H.n[fi]));
    const double d = vdot(L.dir, normal);
    const double scalar = vdot(normal, vector) / d;
    if (scalar < 0) continue;
    f3 ip = vscale(scalar, L.dir);
    ip = vadd(ip, from);
    f3 a = qrot(q, ld3(H.v[H.fvIdx[o]]));
    a = vadd(x, a);
    for (int i = 1; i < nv - 1; i++) {
      if (A.mode == CANNON_RAY_ANY && L.has) return;
      f3 b = qrot(q, ld3(H.v[H.fvIdx[o + i]])), c = qrot(q, ld3(H.v[H.fvIdx[o + i + 1]]));
      b = vadd(x, b);
      c = vadd(x, c);
      const double distance = vdist(ip, from);
      if (!(ray_point_in_triangle(ip, a, b, c) || ray_point_in_triangle(ip, b, a, c)) || distance > fromToDistance) continue;
      ray_report(A, L, ray, normal, ip, body, fi, ord);
    }
  }
}

// AABB.overlaps, aabb.dart:131-147
__device__ __forceinline__ bool ray_aabb_overlaps(const f3& l1, const f3& u1, const f3& l2, const f3& u2) {
  const bool ox = (l2.x <= u1.x && u1.x <= u2.x) || (l1.x <= u2.x && u2.x <= u1.x);
  const bool oy = (l2.y <= u1.y && u1.y <= u2.y) || (l1.y <= u2.y && u2.y <= u1.y);
  const bool oz = (l2.z <= u1.z && u1.z <= u2.z) || (l1.z <= u2.z && u2.z <= u1.z);
  return ox && oy && oz;
}
__device__ __forceinline__ void ray_body_aabb(const BodyArrays& B, const ShapeTables& T, int b, f3& mn, f3& mx) {
  const int sh = B.shape[b];
  const f3 pos = ld3(B.pos[b]);
  if (sh < 0) { mn = pos; mx = pos; } else body_aabb(B, T, b, mn, mx);
}

__global__ void __launch_bounds__(128) k_raycast(BodyArrays B, ShapeTables T, RayArgs A) {
  const int lane = threadIdx.x & 31;
  for (int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ray < A.nRays; ray += (gridDim.x * blockDim.x) >> 5) {
    RayLane L;
    L.from = mk3(A.from[3 * ray], A.from[3 * ray + 1], A.from[3 * ray + 2]);
    L.to = mk3(A.to[3 * ray], A.to[3 * ray + 1], A.to[3 * ray + 2]);
    L.dir = vsub(L.to, L.from);  // Ray._updateDirection, ray_class.dart:258-261
    vnormalize(L.dir);
    L.has = false; L.any = false; L.dist = -1.0; L.key = ~0ull; L.lastKey = 0ull; L.body = -1; L.face = -1; L.lastFace = -1; L.inst = -1; L.curInst = 0;
    L.point = mk3(0.0, 0.0, 0.0); L.normal = L.point;
    f3 lo, hi;  // Ray.getAABB, :329-342
    lo.x = fminf(L.to.x, L.from.x); lo.y = fminf(L.to.y, L.from.y); lo.z = fminf(L.to.z, L.from.z);
    hi.x = fmaxf(L.to.x, L.from.x); hi.y = fmaxf(L.to.y, L.from.y); hi.z = fmaxf(L.to.z, L.from.z);
    for (int b = lane; b < A.nBodies; b += 32) {
      if (A.mode == CANNON_RAY_ANY && L.has) break;  // later bodies of this lane come later in the sequence
      f3 mn, mx;
      ray_body_aabb(B, T, b, mn, mx);
      if (!ray_aabb_overlaps(mn, mx, lo, hi)) continue;
      const int fl = B.flags[b];
      if (A.checkCollisionResponse && !(fl & BF_COLLISION_RESPONSE)) continue;
      if ((A.group & B.mask[b]) == 0 || (B.group[b] & A.mask) == 0) continue;
      if (B.shape[b] < 0) continue;
      const q4 bq = ldq(B.quat[b]);
      const f3 bp = ld3(B.pos[b]);
      const int k0 = T.instFirst ? T.instFirst[b] : 0, k1 = T.instFirst ? T.instFirst[b + 1] : 1;
      int ord = 0;
      for (int k = k0; k < k1; k++) {  // Ray.intersectBody over body.shapes, :226-243
        if (A.mode == CANNON_RAY_ANY && L.has) break;  // result.shouldStop
        L.curInst = k - k0;
        const ShapeDev s = T.shapes[T.instFirst ? T.instShape[k] : B.shape[b]];
        if (A.checkCollisionResponse && !s.collisionResponse) continue;
        q4 so; so.x = so.y = so.z = 0.f; so.w = 1.f;
        f3 off = mk3(0.0, 0.0, 0.0);
        if (T.instFirst) { so = ldq(T.instQuat[k]); off = ld3(T.instOff[k]); }
        const q4 qi = ray_qmul(bq, so);
        const f3 xi = vadd(qrot(bq, off), bp);
        {  // Ray.distanceFromIntersection, :709-722
          const f3 v0 = vsub(xi, L.from);
          const double d = vdot(v0, L.dir);
          f3 ip = vscale(d, L.dir);
          ip = vadd(ip, L.from);
          if (vdist(xi, ip) > s.bsr) continue;
        }
        if (s.type == CANNON_SHAPE_SPHERE) ray_sphere(A, L, ray, s.radius, xi, b, ord);
        else if (s.type == CANNON_SHAPE_PLANE) ray_plane(A, L, ray, qi, xi, b, ord);
        else if (s.type == CANNON_SHAPE_BOX || s.type == CANNON_SHAPE_CONVEX || s.type == CANNON_SHAPE_CYLINDER)  // ray_class.dart:101-123: no other handler
          ray_convex(A, L, ray, hull_view(T, s.hull), qi, xi, b, ord);
      }
    }
    // the warp's answer: (distance, key) minimum for closest, key minimum for any, key maximum for all / hitFaceIndex
    double dist = L.has ? L.dist : INFINITY;
    unsigned long long key = L.has ? L.key : ~0ull;
    int src = lane;
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, dist, o);
      const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o);
      const int os = __shfl_xor_sync(0xffffffffu, src, o);
      bool better;
      if (A.mode == CANNON_RAY_CLOSEST) better = ok != ~0ull && (key == ~0ull || od < dist || (od == dist && ok < key));
      else if (A.mode == CANNON_RAY_ANY) better = ok < key;
      else better = ok != ~0ull && (key == ~0ull || ok > key);
      if (better) { dist = od; key = ok; src = os; }
    }
    unsigned long long lk = L.any ? L.lastKey + 1 : 0ull;  // 0 = no report at all
    int lf = L.lastFace;
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ok = __shfl_xor_sync(0xffffffffu, lk, o);
      const int of = __shfl_xor_sync(0xffffffffu, lf, o);
      if (ok > lk) { lk = ok; lf = of; }
    }
    const bool anyHit = key != ~0ull;
    if (lane == src) {
      A.hasHit[ray] = anyHit ? 1 : 0;
      A.body[ray] = anyHit ? L.body : -1;
      A.inst[ray] = anyHit ? L.inst : -1;
      A.dist[ray] = anyHit ? L.dist : -1.0;
      A.point[ray] = st3(anyHit ? L.point : mk3(0.0, 0.0, 0.0));
      A.normal[ray] = st3(anyHit ? L.normal : mk3(0.0, 0.0, 0.0));
      // any: the reference stopped at its report; closest / all: the latest report of the whole sequence
      A.face[ray] = A.mode == CANNON_RAY_ANY ? (anyHit ? L.face : -1) : lf;
    }
  }
}

// NaiveBroadphase.aabbQuery (naive_broadphase.dart:39-56): flag[b] = body b's AABB overlaps the query box
__global__ void __launch_bounds__(256) k_aabb_query(BodyArrays B, ShapeTables T, int n, f3 lo, f3 hi, int* __restrict__ flag) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += gridDim.x * blockDim.x) {
    f3 mn, mx;
    ray_body_aabb(B, T, b, mn, mx);
    flag[b] = ray_aabb_overlaps(mn, mx, lo, hi) ? 1 : 0;
  }
}
