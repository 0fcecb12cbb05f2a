/*
 * oracle_shapes.cpp — TEST INFRASTRUCTURE ONLY (PARITY UNPINNED, see oracle_math.h).
 * Shape geometry, hull preprocessing, world AABBs and body mass properties, restated from
 *   lib/rigid_body_shapes/{box,cylinder,convex_polyhedron,sphere,plane,heightfield}.dart
 *   lib/objects/rigid_body.dart:395-466,587-609
 */
#include <cmath>
#include <limits>

#include "oracle_world.h"

namespace orc {

// ConvexPolyhedron.computeNormal / getFaceNormal / computeNormals, convex_polyhedron.dart:143-185
void Hull::computeNormals() {
  faceNormals.assign(faces.size(), V3{0, 0, 0});
  for (size_t i = 0; i < faces.size(); i++) {
    const std::vector<int>& f = faces[i];
    const V3& va = vertices[f[0]];
    const V3& vb = vertices[f[1]];
    const V3& vc = vertices[f[2]];
    V3 ab = sub(vb, va);
    V3 cb = sub(vc, vb);
    V3 n = cross(cb, ab);
    if (!(n.x == 0 && n.y == 0 && n.z == 0)) normalize(n);
    faceNormals[i] = neg(n);
  }
}

// ConvexPolyhedron.computeEdges, convex_polyhedron.dart:110-139 (anti-parallel edges are NOT merged)
void Hull::computeEdges() {
  uniqueEdges.clear();
  for (size_t i = 0; i < faces.size(); i++) {
    const std::vector<int>& face = faces[i];
    int nv = (int)face.size();
    for (int j = 0; j < nv; j++) {
      int k = (j + 1) % nv;
      V3 edge = sub(vertices[face[j]], vertices[face[k]]);
      normalize(edge);
      bool found = false;
      for (size_t p = 0; p < uniqueEdges.size(); p++) {
        if (almost_equals(uniqueEdges[p], edge)) {
          found = true;
          break;
        }
      }
      if (!found) uniqueEdges.push_back(edge);
    }
  }
}

// ConvexPolyhedron.updateBoundingSphereRadius, convex_polyhedron.dart:649-660
void Hull::updateBoundingSphereRadius() {
  double max2 = 0;
  for (const V3& v : vertices) {
    double n2 = length2(v);
    if (n2 > max2) max2 = n2;
  }
  boundingSphereRadius = std::sqrt(max2);
}

// ConvexPolyhedron.getPlaneConstantOfFace, convex_polyhedron.dart:405-411
double Hull::planeConstantOfFace(int f) const { return -dot(faceNormals[f], vertices[faces[f][0]]); }

// Box.updateConvexPolyhedronRepresentation, box.dart:41-86
void make_box_hull(const V3& he, Hull& h) {
  double sx = D(he.x), sy = D(he.y), sz = D(he.z);
  h.vertices = {v3(-sx, -sy, -sz), v3(sx, -sy, -sz), v3(sx, sy, -sz), v3(-sx, sy, -sz),
                v3(-sx, -sy, sz),  v3(sx, -sy, sz),  v3(sx, sy, sz),  v3(-sx, sy, sz)};
  h.faces = {{3, 2, 1, 0}, {4, 5, 6, 7}, {5, 4, 0, 1}, {2, 3, 7, 6}, {0, 4, 7, 3}, {1, 2, 6, 5}};
  h.hasUniqueAxes = true;  // `axes` is passed, so face normals are tested (convex_polyhedron.dart:253)
  h.computeNormals();
  h.updateBoundingSphereRadius();
  h.computeEdges();
}

// Cylinder constructor, cylinder.dart:21-101 (axis along y)
void make_cylinder_hull(double radiusTop, double radiusBottom, double height, int N, Hull& h) {
  std::vector<int> bottomface, topface;
  h.vertices.clear();
  h.faces.clear();
  h.vertices.push_back(v3(-radiusBottom * std::sin(0.0), -height * 0.5, radiusBottom * std::cos(0.0)));
  bottomface.push_back(0);
  h.vertices.push_back(v3(-radiusTop * std::sin(0.0), height * 0.5, radiusTop * std::cos(0.0)));
  topface.push_back(1);
  for (int i = 0; i < N; i++) {
    double theta = ((2 * M_PI) / N) * (i + 1);
    if (i < N - 1) {
      h.vertices.push_back(v3(-radiusBottom * std::sin(theta), -height * 0.5, radiusBottom * std::cos(theta)));
      bottomface.push_back(2 * i + 2);
      h.vertices.push_back(v3(-radiusTop * std::sin(theta), height * 0.5, radiusTop * std::cos(theta)));
      topface.push_back(2 * i + 3);
      h.faces.push_back({2 * i, 2 * i + 1, 2 * i + 3, 2 * i + 2});
    } else {
      h.faces.push_back({2 * i, 2 * i + 1, 1, 0});
    }
  }
  h.faces.push_back(bottomface);
  std::vector<int> temp;
  for (size_t i = 0; i < topface.size(); i++) temp.push_back(topface[topface.size() - i - 1]);
  h.faces.push_back(temp);
  h.hasUniqueAxes = true;  // `axes` is passed
  h.computeNormals();
  // the constructor leaves boundingSphereRadius = 0 (cylinder.dart:100); Body.addShape ->
  // updateBoundingRadius -> shape.updateBoundingSphereRadius() recomputes it (rigid_body.dart:403)
  h.updateBoundingSphereRadius();
  h.computeEdges();
}

// Shape.calculateWorldAABB for every in-scope shape
void shape_world_aabb(const Shape& s, const V3& pos, const Q4& q, V3& mn, V3& mx) {
  const float inf = std::numeric_limits<float>::infinity();
  switch (s.type) {
    case CANNON_SHAPE_SPHERE: {  // sphere.dart:43-53
      double r = s.radius;
      mn = v3(D(pos.x) - r, D(pos.y) - r, D(pos.z) - r);
      mx = v3(D(pos.x) + r, D(pos.y) + r, D(pos.z) + r);
      break;
    }
    case CANNON_SHAPE_PLANE: {  // plane.dart:44-69
      V3 n = qvmult(q, V3{0, 0, 1});
      mn = V3{-inf, -inf, -inf};
      mx = V3{inf, inf, inf};
      if (n.x == 1) mx.x = pos.x; else if (n.x == -1) mn.x = pos.x;
      if (n.y == 1) mx.y = pos.y; else if (n.y == -1) mn.y = pos.y;
      if (n.z == 1) mx.z = pos.z; else if (n.z == -1) mn.z = pos.z;
      break;
    }
    case CANNON_SHAPE_BOX: {  // box.dart:149-192
      const V3& e = s.halfExtents;
      V3 c[8] = {{e.x, e.y, e.z},   {-e.x, e.y, e.z},  {-e.x, -e.y, e.z}, {-e.x, -e.y, -e.z},
                 {e.x, -e.y, -e.z}, {e.x, e.y, -e.z},  {-e.x, e.y, -e.z}, {e.x, -e.y, e.z}};
      V3 wc = add(qvmult(q, c[0]), pos);
      mx = wc;
      mn = wc;
      for (int i = 1; i < 8; i++) {
        V3 w = add(qvmult(q, c[i]), pos);
        if (w.x > mx.x) mx.x = w.x;
        if (w.y > mx.y) mx.y = w.y;
        if (w.z > mx.z) mx.z = w.z;
        if (w.x < mn.x) mn.x = w.x;
        if (w.y < mn.y) mn.y = w.y;
        if (w.z < mn.z) mn.z = w.z;
      }
      break;
    }
    case CANNON_SHAPE_CONVEX:
    case CANNON_SHAPE_CYLINDER:
    case CANNON_SHAPE_CAPSULE:
    case CANNON_SHAPE_CONE:
    case CANNON_SHAPE_SIZED_PLANE: {  // convex_polyhedron.dart:663-703
      bool first = true;
      for (const V3& v : s.hull.vertices) {
        V3 w = add(qvmult(q, v), pos);
        if (first) { mn = w; mx = w; first = false; continue; }
        if (w.x < mn.x) mn.x = w.x;
        if (w.x > mx.x) mx.x = w.x;
        if (w.y < mn.y) mn.y = w.y;
        if (w.y > mx.y) mx.y = w.y;
        if (w.z < mn.z) mn.z = w.z;
        if (w.z > mx.z) mx.z = w.z;
      }
      break;
    }
    case CANNON_SHAPE_TRIMESH: {  // trimesh.dart:366-375: AABB.toWorldFrame of the local AABB (aabb.dart:175-187,216-237, setFromPoints :45-82)
      const V3 &l = s.tmLo, &u = s.tmHi;
      const V3 c[8] = {l, V3{u.x, l.y, l.z}, V3{u.x, u.y, l.z}, V3{l.x, u.y, u.z}, V3{u.x, l.y, u.z}, V3{l.x, u.y, l.z}, V3{l.x, l.y, u.z}, u};
      for (int i = 0; i < 8; i++) {
        const V3 p = point_to_world_frame(pos, q, c[i]);
        if (i == 0) { mn = p; mx = p; continue; }
        if (p.x > mx.x) mx.x = p.x;
        if (p.x < mn.x) mn.x = p.x;
        if (p.y > mx.y) mx.y = p.y;
        if (p.y < mn.y) mn.y = p.y;
        if (p.z > mx.z) mx.z = p.z;
        if (p.z < mn.z) mn.z = p.z;
      }
      break;
    }
    case CANNON_SHAPE_PARTICLE:  // particle.dart:29-33
      mn = pos;
      mx = pos;
      break;
    case CANNON_SHAPE_HEIGHTFIELD:  // heightfield.dart:499-503
    default:
      mn = V3{-inf, -inf, -inf};
      mx = V3{inf, inf, inf};
      break;
  }
}

// Body.updateAABB, rigid_body.dart:415-447; AABB.extend, aabb.dart:121-128
void World::updateAABB(Body& b) const {
  for (size_t i = 0; i < b.shapes.size(); i++) {
    V3 offset = qvmult(b.quaternion, b.shapeOffsets[i]);
    offset = add(offset, b.position);
    const Q4 orientation = qmul(b.quaternion, b.shapeOrientations[i]);
    V3 lo, hi;
    shape_world_aabb(shapes[b.shapes[i]], offset, orientation, lo, hi);
    if (i == 0) { b.aabbLower = lo; b.aabbUpper = hi; continue; }
    b.aabbLower = V3{std::fmin(b.aabbLower.x, lo.x), std::fmin(b.aabbLower.y, lo.y), std::fmin(b.aabbLower.z, lo.z)};
    b.aabbUpper = V3{std::fmax(b.aabbUpper.x, hi.x), std::fmax(b.aabbUpper.y, hi.y), std::fmax(b.aabbUpper.z, hi.z)};
  }
}

// Body.updateInertiaWorld, rigid_body.dart:450-466
void World::updateInertiaWorld(Body& b, bool force) const {
  const V3& I = b.invInertia;
  if (I.x == I.y && I.y == I.z && !force) return;
  M3 m1 = m_from_quat(b.quaternion);
  M3 m2 = m_transpose(m1);
  m1 = m_vscale(m1, I);
  b.invInertiaWorld = m_mul(m1, m2);
}

// Body.updateMassProperties, rigid_body.dart:587-609 (+ Box.calculateInertia, box.dart:88-94)
void World::updateMassProperties(Body& b) const {
  b.invMass = b.mass > 0 ? 1.0 / b.mass : 0;
  updateAABB(b);
  V3 he = v3((D(b.aabbUpper.x) - D(b.aabbLower.x)) / 2, (D(b.aabbUpper.y) - D(b.aabbLower.y)) / 2,
             (D(b.aabbUpper.z) - D(b.aabbLower.z)) / 2);
  if (b.shape < 0) he = V3{0, 0, 0};
  double m = b.mass;
  double ex = D(he.x), ey = D(he.y), ez = D(he.z);
  b.inertia.x = (float)(1.0 / 12.0 * m * (2 * ey * 2 * ey + 2 * ez * 2 * ez));
  b.inertia.y = (float)(1.0 / 12.0 * m * (2 * ex * 2 * ex + 2 * ez * 2 * ez));
  b.inertia.z = (float)(1.0 / 12.0 * m * (2 * ey * 2 * ey + 2 * ex * 2 * ex));
  bool fixed = b.fixedRotation;
  b.invInertia = v3(b.inertia.x > 0 && !fixed ? 1.0 / D(b.inertia.x) : 0, b.inertia.y > 0 && !fixed ? 1.0 / D(b.inertia.y) : 0,
                    b.inertia.z > 0 && !fixed ? 1.0 / D(b.inertia.z) : 0);
  updateInertiaWorld(b, true);
}

// Body.updateBoundingRadius, rigid_body.dart:395-412
void World::updateBoundingRadius(Body& b) const {
  double radius = 0;
  for (size_t i = 0; i < b.shapes.size(); i++) {
    const double offset = length(b.shapeOffsets[i]);
    const double r = shapes[b.shapes[i]].boundingSphereRadius;
    if (offset + r > radius) radius = offset + r;
  }
  b.boundingRadius = radius;
}

}  // namespace orc
