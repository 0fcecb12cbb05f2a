/*
 * oracle_raycast.cpp — TEST INFRASTRUCTURE ONLY (PARITY UNPINNED, see oracle_math.h). CPU restatement of the reference's
 * ray casts and AABB query:
 *   World.raycastClosest / raycastAny / raycastAll  lib/world/world_class.dart:248-277
 *   Ray.intersectWorld / intersectBodies / intersectBody / _intersectShape  lib/collision/ray_class.dart:175-199,201-283
 *   _intersectSphere :411-458, _intersectPlane :289-326, _intersectBox :285-287, _intersectConvex :460-553,
 *   _reportIntersection :655-693, Ray.pointInTriangle :696-708, Ray.distanceFromIntersection :709-722
 *   NaiveBroadphase.aabbQuery  lib/collision/naive_broadphase.dart:39-56, AABB.overlaps  lib/collision/aabb.dart:131-147
 * Every Vector3 temporary of the Dart code is a float32 store here (V3), every scalar expression a double.
 * Candidate order: body index order (NaiveBroadphase.aabbQuery). SAPBroadphase.aabbQuery walks its axis list instead and
 * GridBroadphase has no aabbQuery at all (Broadphase.aabbQuery returns [] with a log line, broadphase.dart:151-154): both
 * are served with Naive's order here - a documented choice shared with the CUDA library.
 * Heightfield rays are refused: the reference reads their cell range from a list getIndexOfPosition only appends to
 * (heightfield.dart:173, ray_class.dart:367-372), so its answer depends on the first ray ever cast. Trimesh is out of scope.
 */
#include <algorithm>

#include "oracle_world.h"

namespace orc {

void shape_world_aabb(const Shape& s, const V3& pos, const Q4& q, V3& mn, V3& mx);

namespace {

struct RayCtx {
  V3 from, to, direction;
  int mode, skipBackfaces, mask, group, checkCollisionResponse;
  // RaycastResult (raycast_result.dart)
  bool hasHit = false, shouldStop = false;
  int body = -1, hitFaceIndex = -1;
  int shapeOrdinal = -1, curShape = 0;  // result.shape (set with the result) / the shape being intersected
  double distance = -1;
  V3 hitNormalWorld{0, 0, 0}, hitPointWorld{0, 0, 0};
  std::vector<RayHit>* all = nullptr;
  int rayIndex = 0;
};

// ray_class.dart:655-693
void report(RayCtx& r, const V3& normal, const V3& hitPointWorld, int body, int hitFaceIndex) {
  const double distance = distance_to(r.from, hitPointWorld);
  if (r.skipBackfaces && dot(normal, r.direction) > 0) return;
  r.hitFaceIndex = hitFaceIndex;  // written for every reported intersection, whatever the mode does with it
  auto set = [&]() { r.hitNormalWorld = normal; r.hitPointWorld = hitPointWorld; r.body = body; r.distance = distance; r.shapeOrdinal = r.curShape; };
  switch (r.mode) {
    case CANNON_RAY_ALL:
      r.hasHit = true;
      set();
      r.all->push_back(RayHit{r.rayIndex, body, hitFaceIndex, distance, hitPointWorld, normal, r.curShape});
      break;
    case CANNON_RAY_CLOSEST:
      if (distance < r.distance || !r.hasHit) { r.hasHit = true; set(); }
      break;
    default:  // any
      r.hasHit = true;
      set();
      r.shouldStop = true;
  }
}

// ray_class.dart:696-708
bool pointInTriangle(const V3& p, const V3& a, const V3& b, const V3& c) {
  const V3 v0 = sub(c, a), v1 = sub(b, a), v2 = sub(p, a);
  const double dot00 = dot(v0, v0), dot01 = dot(v0, v1), dot02 = dot(v0, v2), dot11 = dot(v1, v1), dot12 = dot(v1, v2);
  const double u = dot11 * dot02 - dot01 * dot12;
  const double v = dot00 * dot12 - dot01 * dot02;
  return u >= 0 && v >= 0 && (u + v) < (dot00 * dot11 - dot01 * dot01);
}

// ray_class.dart:709-722
double distanceFromIntersection(const V3& from, const V3& direction, const V3& position) {
  const V3 v0 = sub(position, from);
  const double d = dot(v0, direction);
  V3 intersect = scale(d, direction);
  intersect = add(intersect, from);
  return distance_to(position, intersect);
}

// ray_class.dart:411-458
void intersectSphere(RayCtx& r, const Shape& sphere, const V3& position, int body) {
  const V3 &from = r.from, &to = r.to;
  const double rad = sphere.radius;
  const double dx = D(to.x) - D(from.x), dy = D(to.y) - D(from.y), dz = D(to.z) - D(from.z);
  const double fx = D(from.x) - D(position.x), fy = D(from.y) - D(position.y), fz = D(from.z) - D(position.z);
  const double a = dx * dx + dy * dy + dz * dz;  // math.pow(x, 2) is x * x on the VM
  const double b = 2 * (dx * fx + dy * fy + dz * fz);
  const double c = fx * fx + fy * fy + fz * fz - rad * rad;
  const double delta = b * b - 4 * a * c;
  if (delta < 0) return;
  if (delta == 0) {
    const V3 p = lerp(from, to, delta);
    V3 normal = sub(p, position);
    normalize(normal);
    report(r, normal, p, body, -1);
  } else {
    const double d1 = (-b - std::sqrt(delta)) / (2 * a);
    const double d2 = (-b + std::sqrt(delta)) / (2 * a);
    if (d1 >= 0 && d1 <= 1) {
      const V3 p = lerp(from, to, d1);
      V3 normal = sub(p, position);
      normalize(normal);
      report(r, normal, p, body, -1);
    }
    if (r.shouldStop) return;
    if (d2 >= 0 && d2 <= 1) {
      const V3 p = lerp(from, to, d2);
      V3 normal = sub(p, position);
      normalize(normal);
      report(r, normal, p, body, -1);
    }
  }
}

// ray_class.dart:289-326
void intersectPlane(RayCtx& r, const Q4& quat, const V3& position, int body) {
  const V3 &from = r.from, &to = r.to, &direction = r.direction;
  const V3 worldNormal = qvmult(quat, V3{0, 0, 1});
  V3 len = sub(from, position);
  const double planeToFrom = dot(len, worldNormal);
  len = sub(to, position);
  const double planeToTo = dot(len, worldNormal);
  if (planeToFrom * planeToTo > 0) return;
  if (distance_to(from, to) < planeToFrom) return;
  const double nDotDir = dot(worldNormal, direction);
  if (std::fabs(nDotDir) < 0.0001) return;  // Ray.precision
  const V3 planePointToFrom = sub(from, position);
  const double t = -dot(worldNormal, planePointToFrom) / nDotDir;
  const V3 dirScaledWithT = scale(t, direction);
  const V3 hitPointWorld = add(from, dirScaledWithT);
  report(r, worldNormal, hitPointWorld, body, -1);
}

// ray_class.dart:460-553 (faceList == null: every face)
void intersectConvex(RayCtx& r, const Hull& shape, const Q4& q, const V3& x, int body) {
  const V3 &from = r.from, &to = r.to, &direction = r.direction;
  const double fromToDistance = distance_to(from, to);
  const int nFaces = (int)shape.faces.size();
  for (int fi = 0; !r.shouldStop && fi < nFaces; fi++) {
    const std::vector<int>& face = shape.faces[fi];
    V3 vector = shape.vertices[face[0]];
    vector = qvmult(q, vector);
    vector = add(vector, x);
    vector = sub(vector, from);
    const V3 normal = qvmult(q, shape.faceNormals[fi]);
    const double d = dot(direction, normal);
    const double scalar = dot(normal, vector) / d;
    if (scalar < 0) continue;
    V3 ip = scale(scalar, direction);
    ip = add(ip, from);
    V3 a = qvmult(q, shape.vertices[face[0]]);
    a = add(x, a);
    for (int i = 1; !r.shouldStop && i < (int)face.size() - 1; i++) {
      V3 b = qvmult(q, shape.vertices[face[i]]), c = qvmult(q, shape.vertices[face[i + 1]]);
      b = add(x, b);
      c = add(x, c);
      const double distance = distance_to(ip, from);
      if (!(pointInTriangle(ip, a, b, c) || pointInTriangle(ip, b, a, c)) || distance > fromToDistance) continue;
      report(r, normal, ip, body, fi);
    }
  }
}

// aabb.dart:131-147
bool aabbOverlaps(const V3& l1, const V3& u1, const V3& l2, const V3& u2) {
  const bool ox = (l2.x <= u1.x && u1.x <= u2.x) || (l1.x <= u2.x && u2.x <= u1.x);
  const bool oy = (l2.y <= u1.y && u1.y <= u2.y) || (l1.y <= u2.y && u2.y <= u1.y);
  const bool oz = (l2.z <= u1.z && u1.z <= u2.z) || (l1.z <= u2.z && u2.z <= u1.z);
  return ox && oy && oz;
}

}  // namespace

// NaiveBroadphase.aabbQuery, naive_broadphase.dart:39-56 (Body.updateAABB on the current pose)
void World::aabbQuery(const V3& lower, const V3& upper, std::vector<int>& result) {
  for (int i = 0; i < (int)bodies.size(); i++) {
    Body& b = bodies[i];
    if (b.shapes.empty()) { b.aabbLower = b.position; b.aabbUpper = b.position; } else updateAABB(b);
    if (aabbOverlaps(b.aabbLower, b.aabbUpper, lower, upper)) result.push_back(i);
  }
}

// Ray.intersectWorld, ray_class.dart:175-199; returns hasHit. `all` receives the RayMode.all callback sequence.
bool World::raycast(int rayIndex, const V3& from, const V3& to, const cannon_ray_options& opt, RayHit& out, std::vector<RayHit>* all) {
  RayCtx r;
  r.from = from; r.to = to;
  r.mode = opt.mode; r.skipBackfaces = opt.skip_backfaces; r.mask = opt.collision_filter_mask; r.group = opt.collision_filter_group;
  r.checkCollisionResponse = opt.check_collision_response;
  r.all = all; r.rayIndex = rayIndex;
  r.direction = sub(to, from);  // _updateDirection
  normalize(r.direction);
  // getAABB, ray_class.dart:329-342
  const V3 lower{std::min(to.x, from.x), std::min(to.y, from.y), std::min(to.z, from.z)};
  const V3 upper{std::max(to.x, from.x), std::max(to.y, from.y), std::max(to.z, from.z)};
  std::vector<int> cand;
  aabbQuery(lower, upper, cand);
  for (size_t k = 0; !r.shouldStop && k < cand.size(); k++) {  // intersectBodies / intersectBody, :201-268
    const Body& body = bodies[cand[k]];
    if (r.checkCollisionResponse && !body.collisionResponse) continue;
    if ((r.group & body.mask) == 0 || (body.group & r.mask) == 0) continue;
    for (size_t si = 0; si < body.shapes.size(); si++) {  // intersectBody, :226-243
      const Shape& shape = shapes[body.shapes[si]];
      r.curShape = (int)si;
      if (r.checkCollisionResponse && !shape.collisionResponse) continue;
      const Q4 qi = qmul(body.quaternion, body.shapeOrientations[si]);
      const V3 xi = add(qvmult(body.quaternion, body.shapeOffsets[si]), body.position);
      // _intersectShape, :270-283
      if (!(distanceFromIntersection(r.from, r.direction, xi) > shape.boundingSphereRadius)) {
        switch (shape.type) {
          case CANNON_SHAPE_SPHERE: intersectSphere(r, shape, xi, cand[k]); break;
          case CANNON_SHAPE_PLANE: intersectPlane(r, qi, xi, cand[k]); break;
          case CANNON_SHAPE_BOX: case CANNON_SHAPE_CYLINDER: case CANNON_SHAPE_CONVEX: intersectConvex(r, shape.hull, qi, xi, cand[k]); break;
          default: break;  // heightfield rays: outside the scope (the API refuses such worlds); no handler for the other types
        }
      }
      if (r.shouldStop) break;
    }
  }
  out = RayHit{rayIndex, r.body, r.hitFaceIndex, r.distance, r.hitPointWorld, r.hitNormalWorld, r.hasHit ? r.shapeOrdinal : -1};
  return r.hasHit;
}

}  // namespace orc
