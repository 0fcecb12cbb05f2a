/*
 * oracle_narrowphase.cpp — TEST INFRASTRUCTURE ONLY (PARITY UNPINNED, see oracle_math.h).
 * Contact generation restated from lib/world/narrow_phase.dart and
 * lib/rigid_body_shapes/{convex_polyhedron,heightfield}.dart:
 *   getContacts :634-721, createContactEquation :492-528, createFrictionEquationsFromContact :530-586,
 *   sphereSphere :723, spherePlane :766, sphereBox :816, sphereConvex :1039, sphereHeightfield :1295,
 *   planeConvex/planeBox :1847/:1786, convexConvex/boxBox/boxConvex :1981/:1693/:1713,
 *   heightfieldConvex/boxHeightfield :2044/:1749, _pointInPolygon :2584,
 *   ConvexPolyhedron.findSeparatingAxis :232, testSepAxis :360, project :843, clipAgainstHull :189,
 *   clipFaceAgainstHull :417, clipFaceAgainstPlane :545, Heightfield.getConvexTrianglePillar :330,
 *   sphereParticle :1258, planeParticle :1805, boxParticle :1731, particleConvex :2179, heightfieldParticle :2343,
 *   ConvexPolyhedron.pointIsInside :760, getAveragePointLocal :712, computeWorldVertices :590, computeWorldFaceNormals :633.
 */
#include <algorithm>
#include <cmath>
#include <limits>

#include "oracle_world.h"

namespace orc {

namespace {

struct NP {
  World& w;
  const cannon_contact_material* cm = nullptr;  // currentContactMaterial
  // Shape.material of the shapes handed to the current resolver, by body (createContactEquation: `si.material ?? bi.material`;
  // a heightfield pillar is a plain ConvexPolyhedron without material) and of the pair's shapes in pair order (c.si / c.sj = rsi / rsj)
  int resBody[2] = {-1, -1}, resMat[2] = {-1, -1};
  int rsiMat = -1, rsjMat = -1;
  int shapeMatOf(int body) const { return body == resBody[0] ? resMat[0] : (body == resBody[1] ? resMat[1] : -1); }
  int manifold = 0;  // ordinal of the current resolver call (one sphereConvex / convexConvex call per heightfield pillar)
  explicit NP(World& w_) : w(w_) {}

  // createContactEquation, narrow_phase.dart:492-528. `crA/crB` are the collisionResponse flags of the
  // shapes handed to the resolver (for boxes: the hull representation, which copies the box's flag).
  Eq createContactEquation(int bi, int bj, bool crA, bool crB) {
    Eq c;
    c.kind = EQ_CONTACT;
    c.bi = bi;
    c.bj = bj;
    c.minForce = 0;
    c.maxForce = 1e6;
    const Body& A = w.bodies[bi];
    const Body& B = w.bodies[bj];
    c.enabled = A.collisionResponse && B.collisionResponse && crA && crB;
    c.restitution = cm->restitution;
    c.setSpookParams(cm->contact_equation_stiffness, cm->contact_equation_relaxation, w.dt);
    int matA = shapeMatOf(bi) >= 0 ? shapeMatOf(bi) : A.material, matB = shapeMatOf(bj) >= 0 ? shapeMatOf(bj) : B.material;  // :517-518
    if (matA >= 0 && matB >= 0 && w.matRestitution[matA] >= 0 && w.matRestitution[matB] >= 0)
      c.restitution = w.matRestitution[matA] * w.matRestitution[matB];
    return c;
  }

  // createFrictionEquationsFromContact, narrow_phase.dart:530-586
  void addContact(Eq& c) {
    const Body& A = w.bodies[c.bi];
    const Body& B = w.bodies[c.bj];
    double friction = cm->friction;
    int matA = rsiMat >= 0 ? rsiMat : A.material, matB = rsjMat >= 0 ? rsjMat : B.material;  // shapeA = c.si = rsi, bodyA = c.bi (:533-542)
    if (matA >= 0 && matB >= 0 && w.matFriction[matA] >= 0 && w.matFriction[matB] >= 0)
      friction = w.matFriction[matA] * w.matFriction[matB];
    c.friction = friction;
    w.contacts.push_back(c);
    w.contactManifold.push_back(manifold);
    if (friction > 0) {
      V3 g = w.desc.has_friction_gravity ? V3{w.desc.friction_gravity[0], w.desc.friction_gravity[1], w.desc.friction_gravity[2]}
                                         : V3{w.desc.gravity[0], w.desc.gravity[1], w.desc.gravity[2]};
      double mug = friction * length(g);
      double reducedMass = A.invMass + B.invMass;
      if (reducedMass > 0) reducedMass = 1 / reducedMass;
      Eq c1;
      c1.kind = EQ_FRICTION;
      c1.bi = c.bi;
      c1.bj = c.bj;
      c1.minForce = -mug * reducedMass;
      c1.maxForce = mug * reducedMass;
      c1.ri = c.ri;
      c1.rj = c.rj;
      c1.setSpookParams(cm->friction_equation_stiffness, cm->friction_equation_relaxation, w.dt);
      c1.enabled = c.enabled;
      Eq c2 = c1;
      tangents(c.ni, c1.ni, c2.ni);
      w.frictions.push_back(c1);
      w.frictions.push_back(c2);
    }
  }

  // "Make relative to bodies": r.add2(x, r); r.sub2(body.position, r)
  static V3 rel(const V3& r, const V3& x, const V3& bodyPos) { return sub(add(r, x), bodyPos); }

  // sphereSphere, narrow_phase.dart:723-765 (no distance test of its own)
  void sphereSphere(const Shape& si, const Shape& sj, V3 xi, V3 xj, int bi, int bj) {
    Eq c = createContactEquation(bi, bj, si.collisionResponse, sj.collisionResponse);
    c.ni = sub(xj, xi);
    normalize(c.ni);
    c.ri = scale(si.radius, c.ni);
    c.rj = scale(-sj.radius, c.ni);
    c.ri = rel(c.ri, xi, w.bodies[bi].position);
    c.rj = rel(c.rj, xj, w.bodies[bj].position);
    addContact(c);
  }

  // spherePlane, narrow_phase.dart:766-815
  void spherePlane(const Shape& si, const Shape& sj, V3 xi, V3 xj, Q4 qj, int bi, int bj) {
    Eq r = createContactEquation(bi, bj, si.collisionResponse, sj.collisionResponse);
    r.ni = qvmult(qj, V3{0, 0, 1});
    r.ni = neg(r.ni);
    normalize(r.ni);
    r.ri = scale(si.radius, r.ni);
    V3 pointOnPlaneToSphere = sub(xi, xj);
    V3 planeToSphereOrtho = scale(dot(r.ni, pointOnPlaneToSphere), r.ni);
    r.rj = sub(pointOnPlaneToSphere, planeToSphereOrtho);
    if (-dot(pointOnPlaneToSphere, r.ni) <= si.radius) {
      r.ri = rel(r.ri, xi, w.bodies[bi].position);
      r.rj = rel(r.rj, xj, w.bodies[bj].position);
      addContact(r);
    }
  }

  // sphereBox, narrow_phase.dart:816-1038
  void sphereBox(const Shape& si, const Shape& sj, V3 xi, V3 xj, Q4 qj, int bi, int bj) {
    V3 sides[6];
    {  // Box.getSideNormals, box.dart:99-115
      const V3& ex = sj.halfExtents;
      sides[0] = V3{ex.x, 0, 0};
      sides[1] = V3{0, ex.y, 0};
      sides[2] = V3{0, 0, ex.z};
      sides[3] = V3{-ex.x, 0, 0};
      sides[4] = V3{0, -ex.y, 0};
      sides[5] = V3{0, 0, -ex.z};
      for (int i = 0; i < 6; i++) sides[i] = qvmult(qj, sides[i]);
    }
    V3 boxToSphere = sub(xi, xj);
    const double R = si.radius;
    bool found = false;
    V3 sideNs{0, 0, 0}, sideNs1{0, 0, 0}, sideNs2{0, 0, 0};
    double sideH = 0, sideDot1 = 0, sideDot2 = 0, sideDistance = 0;
    bool haveSideDistance = false;
    int sidePenetrations = 0;
    for (int idx = 0; idx != 6 && !found; idx++) {
      V3 ns = sides[idx];
      double h = length(ns);
      normalize(ns);
      double dt = dot(boxToSphere, ns);
      if (dt < h + R && dt > 0) {
        V3 ns1 = sides[(idx + 1) % 3];
        V3 ns2 = sides[(idx + 2) % 3];
        double h1 = length(ns1), h2 = length(ns2);
        normalize(ns1);
        normalize(ns2);
        double dot1 = dot(boxToSphere, ns1);
        double dot2 = dot(boxToSphere, ns2);
        if (dot1 < h1 && dot1 > -h1 && dot2 < h2 && dot2 > -h2) {
          double dist = std::fabs(dt - h - R);
          if (!haveSideDistance || dist < sideDistance) {
            haveSideDistance = true;
            sideDistance = dist;
            sideDot1 = dot1;
            sideDot2 = dot2;
            sideH = h;
            sideNs = ns;
            sideNs1 = ns1;
            sideNs2 = ns2;
            sidePenetrations++;
          }
        }
      }
    }
    if (sidePenetrations != 0) {
      found = true;
      Eq r = createContactEquation(bi, bj, si.collisionResponse, sj.collisionResponse);
      r.ri = scale(-R, sideNs);
      r.ni = neg(sideNs);
      sideNs = scale(sideH, sideNs);
      sideNs1 = scale(sideDot1, sideNs1);
      sideNs = add(sideNs, sideNs1);
      sideNs2 = scale(sideDot2, sideNs2);
      r.rj = add(sideNs, sideNs2);
      r.ri = rel(r.ri, xi, w.bodies[bi].position);
      r.rj = rel(r.rj, xj, w.bodies[bj].position);
      addContact(r);
    }
    // corners
    for (int j = 0; j != 2 && !found; j++)
      for (int k = 0; k != 2 && !found; k++)
        for (int l = 0; l != 2 && !found; l++) {
          V3 rj{0, 0, 0};
          rj = (j != 0) ? add(sides[0], rj) : sub(rj, sides[0]);
          rj = (k != 0) ? add(sides[1], rj) : sub(rj, sides[1]);
          rj = (l != 0) ? add(sides[2], rj) : sub(rj, sides[2]);
          V3 sphereToCorner = add(xj, rj);
          sphereToCorner = sub(sphereToCorner, xi);
          if (length2(sphereToCorner) < R * R) {
            found = true;
            Eq r = createContactEquation(bi, bj, si.collisionResponse, sj.collisionResponse);
            r.ri = sphereToCorner;
            normalize(r.ri);
            r.ni = r.ri;
            r.ri = scale(R, r.ri);
            r.rj = rj;
            r.ri = rel(r.ri, xi, w.bodies[bi].position);
            r.rj = rel(r.rj, xj, w.bodies[bj].position);
            addContact(r);
          }
        }
    // edges
    for (int j = 0; j != 6 && !found; j++)
      for (int k = 0; k != 6 && !found; k++) {
        if (j % 3 == k % 3) continue;
        V3 edgeTangent = cross(sides[k], sides[j]);
        normalize(edgeTangent);
        V3 edgeCenter = add(sides[j], sides[k]);
        V3 r = xi;
        r = sub(r, edgeCenter);
        r = sub(r, xj);
        double orthonorm = dot(r, edgeTangent);
        V3 orthogonal = scale(orthonorm, edgeTangent);
        int l = 0;
        while (l == j % 3 || l == k % 3) l++;
        V3 dist = xi;
        dist = sub(dist, orthogonal);
        dist = sub(dist, edgeCenter);
        dist = sub(dist, xj);
        double tdist = std::fabs(orthonorm);
        double ndist = length(dist);
        if (tdist < length(sides[l]) && ndist < R) {
          found = true;
          Eq res = createContactEquation(bi, bj, si.collisionResponse, sj.collisionResponse);
          res.rj = add(edgeCenter, orthogonal);
          res.ni = neg(dist);
          normalize(res.ni);
          res.ri = res.rj;
          res.ri = add(res.ri, xj);
          res.ri = sub(res.ri, xi);
          normalize(res.ri);
          res.ri = scale(R, res.ri);
          res.ri = rel(res.ri, xi, w.bodies[bi].position);
          res.rj = rel(res.rj, xj, w.bodies[bj].position);
          addContact(res);
        }
      }
  }

  // _pointInPolygon, narrow_phase.dart:2584-2617
  static bool pointInPolygon(const std::vector<V3>& verts, const V3& normal, const V3& p) {
    int positiveResult = -1;  // null
    const int N = (int)verts.size();
    for (int i = 0; i != N; i++) {
      const V3& v = verts[i];
      V3 edge = sub(verts[(i + 1) % N], v);
      V3 edgeXNormal = cross(edge, normal);
      V3 vertexToP = sub(p, v);
      double r = dot(edgeXNormal, vertexToP);
      if (positiveResult == -1 || (r > 0 && positiveResult == 1) || (r <= 0 && positiveResult == 0)) {
        if (positiveResult == -1) positiveResult = r > 0 ? 1 : 0;
        continue;
      } else {
        return false;
      }
    }
    return true;
  }

  // sphereConvex, narrow_phase.dart:1039-1257. (xj,qj) is the hull frame; crB = hull.collisionResponse.
  void sphereConvex(const Shape& si, const Hull& sj, bool crB, V3 xi, V3 xj, Q4 qj, int bi, int bj) {
    const double R = si.radius;
    const std::vector<V3>& verts = sj.vertices;
    for (size_t i = 0; i != verts.size(); i++) {
      V3 worldCorner = qvmult(qj, verts[i]);
      worldCorner = add(xj, worldCorner);
      V3 sphereToCorner = sub(worldCorner, xi);
      if (length2(sphereToCorner) < R * R) {
        Eq r = createContactEquation(bi, bj, si.collisionResponse, crB);
        r.ri = sphereToCorner;
        normalize(r.ri);
        r.ni = r.ri;
        r.ri = scale(R, r.ri);
        r.rj = sub(worldCorner, xj);
        r.ri = rel(r.ri, xi, w.bodies[bi].position);
        r.rj = rel(r.rj, xj, w.bodies[bj].position);
        addContact(r);
        return;
      }
    }
    for (size_t i = 0; i != sj.faces.size(); i++) {
      const V3& normal = sj.faceNormals[i];
      const std::vector<int>& face = sj.faces[i];
      V3 worldNormal = qvmult(qj, normal);
      V3 worldPoint = qvmult(qj, verts[face[0]]);
      worldPoint = add(worldPoint, xj);
      V3 worldSpherePointClosestToPlane = scale(-R, worldNormal);
      worldSpherePointClosestToPlane = add(xi, worldSpherePointClosestToPlane);
      V3 penetrationVec = sub(worldSpherePointClosestToPlane, worldPoint);
      double penetration = dot(penetrationVec, worldNormal);
      V3 worldPointToSphere = sub(xi, worldPoint);
      if (penetration < 0 && dot(worldPointToSphere, worldNormal) > 0) {
        std::vector<V3> faceVerts;
        for (size_t j = 0; j != face.size(); j++) {
          V3 worldVertex = qvmult(qj, verts[face[j]]);
          worldVertex = add(xj, worldVertex);
          faceVerts.push_back(worldVertex);
        }
        if (pointInPolygon(faceVerts, worldNormal, xi)) {
          Eq r = createContactEquation(bi, bj, si.collisionResponse, crB);
          r.ri = scale(-R, worldNormal);
          r.ni = neg(worldNormal);
          V3 penetrationVec2 = scale(-penetration, worldNormal);
          V3 penetrationSpherePoint = scale(-R, worldNormal);
          r.rj = sub(xi, xj);
          r.rj = add(r.rj, penetrationSpherePoint);
          r.rj = add(r.rj, penetrationVec2);
          r.rj = rel(r.rj, xj, w.bodies[bj].position);
          r.ri = rel(r.ri, xi, w.bodies[bi].position);
          addContact(r);
          return;
        } else {
          const int L = (int)face.size();
          for (int j = 0; j != L; j++) {
            V3 v1 = qvmult(qj, verts[face[(j + 1) % L]]);
            V3 v2 = qvmult(qj, verts[face[(j + 2) % L]]);
            v1 = add(xj, v1);
            v2 = add(xj, v2);
            V3 edge = sub(v2, v1);
            V3 edgeUnit = unit(edge);
            V3 v1ToXi = sub(xi, v1);
            double dt = dot(v1ToXi, edgeUnit);
            V3 p = scale(dt, edgeUnit);
            p = add(p, v1);
            V3 xiToP = sub(p, xi);
            if (dt > 0 && dt * dt < length2(edge) && length2(xiToP) < R * R) {
              Eq r = createContactEquation(bi, bj, si.collisionResponse, crB);
              r.rj = sub(p, xj);
              r.ni = sub(p, xi);
              normalize(r.ni);
              r.ri = scale(R, r.ni);
              r.rj = rel(r.rj, xj, w.bodies[bj].position);
              r.ri = rel(r.ri, xi, w.bodies[bi].position);
              addContact(r);
              return;
            }
          }
        }
      }
    }
  }

  // planeConvex / planeBox, narrow_phase.dart:1847-1915 / 1786-1804
  void planeConvex(const Shape& si, const Hull& sj, bool crB, V3 xi, V3 xj, Q4 qi, Q4 qj, int bi, int bj) {
    V3 worldNormal = qvmult(qi, V3{0, 0, 1});
    for (size_t i = 0; i != sj.vertices.size(); i++) {
      V3 worldVertex = qvmult(qj, sj.vertices[i]);
      worldVertex = add(xj, worldVertex);
      V3 relpos = sub(worldVertex, xi);
      double dt = dot(worldNormal, relpos);
      if (dt <= 0.0) {
        Eq r = createContactEquation(bi, bj, si.collisionResponse, crB);
        V3 projected = scale(dot(worldNormal, relpos), worldNormal);
        projected = sub(worldVertex, projected);
        r.ri = sub(projected, xi);
        r.ni = worldNormal;
        r.rj = sub(worldVertex, xj);
        r.ri = rel(r.ri, xi, w.bodies[bi].position);
        r.rj = rel(r.rj, xj, w.bodies[bj].position);
        addContact(r);
      }
    }
  }

  // ConvexPolyhedron.project, convex_polyhedron.dart:843-883
  static void project(const Hull& shape, const V3& axis, const V3& pos, const Q4& quat, double& mx, double& mn) {
    V3 localAxis = vector_to_local_frame(quat, axis);
    V3 localOrigin = point_to_local_frame(pos, quat, V3{0, 0, 0});
    double add_ = dot(localOrigin, localAxis);
    const std::vector<V3>& vs = shape.vertices;
    mn = mx = dot(vs[0], localAxis);
    for (size_t i = 1; i < vs.size(); i++) {
      double val = dot(vs[i], localAxis);
      if (val > mx) mx = val;
      if (val < mn) mn = val;
    }
    mn -= add_;
    mx -= add_;
    if (mn > mx) std::swap(mn, mx);
  }

  // testSepAxis, convex_polyhedron.dart:360-384 ; returns false when separated
  static bool testSepAxis(const Hull& A, const V3& axis, const Hull& B, const V3& posA, const Q4& quatA, const V3& posB,
                          const Q4& quatB, double& depth) {
    double maxA, minA, maxB, minB;
    project(A, axis, posA, quatA, maxA, minA);
    project(B, axis, posB, quatB, maxB, minB);
    if (maxA < minB || maxB < minA) return false;
    double d0 = maxA - minB, d1 = maxB - minA;
    depth = d0 < d1 ? d0 : d1;
    return true;
  }

  // findSeparatingAxis, convex_polyhedron.dart:232-356 (inverted uniqueAxes logic reproduced, §5.9-9)
  static bool findSeparatingAxis(const Hull& A, const Hull& B, const V3& posA, const Q4& quatA, const V3& posB,
                                 const Q4& quatB, V3& target, const int* faceListA, int nFaceListA) {
    double dmin = std::numeric_limits<double>::infinity();
    if (A.hasUniqueAxes) {
      int numFacesA = faceListA ? nFaceListA : (int)A.faces.size();
      for (int i = 0; i < numFacesA; i++) {
        int fi = faceListA ? faceListA[i] : i;
        V3 n = qvmult(quatA, A.faceNormals[fi]);
        double d;
        if (!testSepAxis(A, n, B, posA, quatA, posB, quatB, d)) return false;
        if (d < dmin) { dmin = d; target = n; }
      }
    }
    if (B.hasUniqueAxes) {
      for (size_t i = 0; i < B.faces.size(); i++) {
        V3 n = qvmult(quatB, B.faceNormals[i]);
        double d;
        if (!testSepAxis(A, n, B, posA, quatA, posB, quatB, d)) return false;
        if (d < dmin) { dmin = d; target = n; }
      }
    }
    for (size_t e0 = 0; e0 != A.uniqueEdges.size(); e0++) {
      V3 worldEdge0 = qvmult(quatA, A.uniqueEdges[e0]);
      for (size_t e1 = 0; e1 != B.uniqueEdges.size(); e1++) {
        V3 worldEdge1 = qvmult(quatB, B.uniqueEdges[e1]);
        V3 c = cross(worldEdge0, worldEdge1);
        if (!almost_zero(c)) {
          normalize(c);
          double d;
          if (!testSepAxis(A, c, B, posA, quatA, posB, quatB, d)) return false;
          if (d < dmin) { dmin = d; target = c; }
        }
      }
    }
    V3 deltaC = sub(posB, posA);
    if (dot(deltaC, target) > 0.0) target = neg(target);
    return true;
  }

  struct ClipPoint { V3 point, normal; double depth; };

  // clipFaceAgainstPlane, convex_polyhedron.dart:545-587
  static void clipFaceAgainstPlane(const std::vector<V3>& in, std::vector<V3>& out, const V3& n, double c) {
    const int numVerts = (int)in.size();
    if (numVerts < 2) return;
    V3 firstVertex = in[numVerts - 1];
    double nDotFirst = dot(n, firstVertex) + c;
    for (int vi = 0; vi < numVerts; vi++) {
      V3 lastVertex = in[vi];
      double nDotLast = dot(n, lastVertex) + c;
      if (nDotFirst < 0) {
        if (nDotLast < 0) out.push_back(lastVertex);
        else out.push_back(lerp(firstVertex, lastVertex, nDotFirst / (nDotFirst - nDotLast)));
      } else {
        if (nDotLast < 0) {
          out.push_back(lerp(firstVertex, lastVertex, nDotFirst / (nDotFirst - nDotLast)));
          out.push_back(lastVertex);
        }
      }
      firstVertex = lastVertex;
      nDotFirst = nDotLast;
    }
  }

  // clipFaceAgainstHull, convex_polyhedron.dart:417-541
  static void clipFaceAgainstHull(const Hull& A, const V3& sepNormal, const V3& posA, const Q4& quatA,
                                  std::vector<V3> worldVertsB1, double minDist, double maxDist, std::vector<ClipPoint>& result) {
    int closestFaceA = -1;
    double dmin = std::numeric_limits<double>::infinity();
    for (size_t face = 0; face < A.faces.size(); face++) {
      V3 n = qvmult(quatA, A.faceNormals[face]);
      double d = dot(n, sepNormal);
      if (d < dmin) { dmin = d; closestFaceA = (int)face; }
    }
    if (closestFaceA < 0) return;
    const std::vector<int>& polyA = A.faces[closestFaceA];
    std::vector<int> connectedFaces;
    for (size_t i = 0; i < A.faces.size(); i++)
      for (size_t j = 0; j < A.faces[i].size(); j++) {
        bool shares = std::find(polyA.begin(), polyA.end(), A.faces[i][j]) != polyA.end();
        if (shares && (int)i != closestFaceA &&
            std::find(connectedFaces.begin(), connectedFaces.end(), (int)i) == connectedFaces.end())
          connectedFaces.push_back((int)i);
      }
    std::vector<V3> pVtxIn = worldVertsB1, pVtxOut;
    const int numVerticesA = (int)polyA.size();
    for (int i = 0; i < numVerticesA; i++) {
      int otherFace = (!connectedFaces.empty() && (int)connectedFaces.size() > i) ? connectedFaces[i] : 0;
      double localPlaneEq = A.planeConstantOfFace(otherFace);
      V3 planeNormalWS = qvmult(quatA, A.faceNormals[otherFace]);
      double planeEqWS = localPlaneEq - dot(planeNormalWS, posA);
      pVtxOut.clear();
      clipFaceAgainstPlane(pVtxIn, pVtxOut, planeNormalWS, planeEqWS);
      pVtxIn = pVtxOut;
    }
    double localPlaneEq = A.planeConstantOfFace(closestFaceA);
    V3 planeNormalWS = qvmult(quatA, A.faceNormals[closestFaceA]);
    double planeEqWS = localPlaneEq - dot(planeNormalWS, posA);
    for (size_t i = 0; i < pVtxIn.size(); i++) {
      double depth = dot(planeNormalWS, pVtxIn[i]) + planeEqWS;
      if (depth <= minDist) depth = minDist;
      if (depth <= maxDist) {
        if (depth <= 1e-6) result.push_back(ClipPoint{pVtxIn[i], planeNormalWS, depth});
      }
    }
  }

  // clipAgainstHull, convex_polyhedron.dart:189-227
  static void clipAgainstHull(const Hull& A, const V3& posA, const Q4& quatA, const Hull& B, const V3& posB, const Q4& quatB,
                              const V3& sepNormal, double minDist, double maxDist, std::vector<ClipPoint>& result) {
    int closestFaceB = -1;
    double dmax = -std::numeric_limits<double>::infinity();
    for (size_t face = 0; face < B.faces.size(); face++) {
      V3 n = qvmult(quatB, B.faceNormals[face]);
      double d = dot(n, sepNormal);
      if (d > dmax) { dmax = d; closestFaceB = (int)face; }
    }
    if (closestFaceB < 0) return;
    std::vector<V3> worldVertsB1;
    for (size_t i = 0; i < B.faces[closestFaceB].size(); i++) {
      V3 wb = qvmult(quatB, B.vertices[B.faces[closestFaceB][i]]);
      wb = add(posB, wb);
      worldVertsB1.push_back(wb);
    }
    clipFaceAgainstHull(A, sepNormal, posA, quatA, worldVertsB1, minDist, maxDist, result);
  }

  // convexConvex, narrow_phase.dart:1981-2043
  void convexConvex(const Hull& si, const Hull& sj, bool crA, bool crB, V3 xi, V3 xj, Q4 qi, Q4 qj, int bi, int bj,
                    const int* faceListA, int nFaceListA) {
    if (distance_to(xi, xj) > si.boundingSphereRadius + sj.boundingSphereRadius) return;
    V3 sepAxis{0, 0, 0};
    if (findSeparatingAxis(si, sj, xi, qi, xj, qj, sepAxis, faceListA, nFaceListA)) {
      std::vector<ClipPoint> res;
      clipAgainstHull(si, xi, qi, sj, xj, qj, sepAxis, -100, 100, res);
      for (size_t j = 0; j != res.size(); j++) {
        Eq r = createContactEquation(bi, bj, crA, crB);
        r.ni = neg(sepAxis);
        V3 q = neg(res[j].normal);
        q = scale(res[j].depth, q);
        V3 ri = add(res[j].point, q);
        V3 rj = res[j].point;
        ri = sub(ri, xi);
        rj = sub(rj, xj);
        ri = add(ri, xi);
        ri = sub(ri, w.bodies[bi].position);
        rj = add(rj, xj);
        rj = sub(rj, w.bodies[bj].position);
        r.ri = ri;
        r.rj = rj;
        addContact(r);
      }
    }
  }

  // sphereParticle, narrow_phase.dart:1258-1294. (sj, xj, bj) = the sphere, (xi, bi) = the particle; the contact's
  // first body is the PARTICLE's (createContactEquation(bi, bj, ...), :1280).
  void sphereParticle(const Shape& sj, const Shape& si, V3 xj, V3 xi, int bj, int bi) {
    V3 normal = sub(xi, xj);
    const double lengthSquared = length2(normal);
    if (lengthSquared <= sj.radius * sj.radius) {
      Eq r = createContactEquation(bi, bj, si.collisionResponse, sj.collisionResponse);
      normalize(normal);
      r.rj = normal;
      r.rj = scale(sj.radius, r.rj);
      r.ni = neg(normal);
      r.ri = V3{0, 0, 0};
      addContact(r);
    }
  }

  // planeParticle, narrow_phase.dart:1805-1846
  void planeParticle(const Shape& sj, const Shape& si, V3 xj, V3 xi, int bj, int bi) {
    V3 normal = qvmult(w.bodies[bj].quaternion, V3{0, 0, 1});
    V3 relpos = sub(xi, w.bodies[bj].position);
    const double d = dot(normal, relpos);
    if (d <= 0.0) {
      Eq r = createContactEquation(bi, bj, si.collisionResponse, sj.collisionResponse);
      r.ni = neg(normal);
      r.ri = V3{0, 0, 0};
      V3 projected = scale(dot(normal, xi), normal);
      projected = sub(xi, projected);
      r.rj = projected;  // the reference does not subtract the plane position here
      addContact(r);
    }
  }

  // ConvexPolyhedron.pointIsInside, convex_polyhedron.dart:760-787 (with getAveragePointLocal :712-720)
  static bool pointIsInside(const Hull& h, const V3& p) {
    V3 pointInside{0, 0, 0};
    for (const V3& v : h.vertices) pointInside = add(pointInside, v);
    pointInside = scale(1.0 / (double)h.vertices.size(), pointInside);
    for (size_t i = 0; i < h.faces.size(); i++) {
      const V3& n = h.faceNormals[i];
      const V3& v = h.vertices[h.faces[i][0]];
      const V3 vToP = sub(p, v);
      const double r1 = dot(n, vToP);
      const V3 vToPointInside = sub(pointInside, v);
      const double r2 = dot(n, vToPointInside);
      if ((r1 < 0 && r2 > 0) || (r1 > 0 && r2 < 0)) return false;
    }
    return true;
  }

  // particleConvex, narrow_phase.dart:2179-2264. (sj, xj, qj, bj) = the hull, (xi, bi) = the particle. The world
  // vertices / normals are those frozen in `wv` / `wn` (computed on the first penetration of this hull object).
  void particleConvex(const Hull& sj, bool crHull, bool crParticle, std::vector<V3>& wv, std::vector<V3>& wn, bool& needsUpdate, V3 xj, V3 xi, Q4 qj,
                      int bj, int bi) {
    V3 local = sub(xi, xj);
    local = qvmult(qconj(qj), local);
    if (!pointIsInside(sj, local)) return;
    if (needsUpdate) {  // computeWorldVertices :590-604 + computeWorldFaceNormals :633-646; never invalidated afterwards
      wv.resize(sj.vertices.size());
      for (size_t i = 0; i < sj.vertices.size(); i++) wv[i] = add(xj, qvmult(qj, sj.vertices[i]));
      wn.resize(sj.faceNormals.size());
      for (size_t i = 0; i < sj.faceNormals.size(); i++) wn[i] = qvmult(qj, sj.faceNormals[i]);
      needsUpdate = false;
    }
    int penetratedFaceIndex = -1;
    V3 penetratedFaceNormal{0, 0, 0};
    double minPenetration = 0;
    bool have = false;
    for (size_t i = 0; i < sj.faces.size(); i++) {
      const V3& verts = wv[sj.faces[i][0]];
      const V3& normal = wn[i];
      const V3 v2p = sub(xi, verts);
      const double penetration = -dot(normal, v2p);
      if (!have || std::fabs(penetration) < std::fabs(minPenetration)) {
        have = true;
        minPenetration = penetration;
        penetratedFaceIndex = (int)i;
        penetratedFaceNormal = normal;
      }
    }
    if (penetratedFaceIndex != -1) {
      Eq r = createContactEquation(bi, bj, crParticle, crHull);
      V3 worldPenetrationVec = scale(minPenetration, penetratedFaceNormal);
      worldPenetrationVec = add(worldPenetrationVec, xi);
      worldPenetrationVec = sub(worldPenetrationVec, xj);
      r.rj = worldPenetrationVec;
      r.rj = qvmult(qj, r.rj);  // :2246 rotates the world-frame vector once more
      r.ni = neg(penetratedFaceNormal);
      r.ri = V3{0, 0, 0};
      r.ri = rel(r.ri, xi, w.bodies[bi].position);
      r.rj = rel(r.rj, xj, w.bodies[bj].position);
      addContact(r);
    }
  }

  // Heightfield.getConvexTrianglePillar, heightfield.dart:330-487
  static void pillar(const Shape& hf, int xi, int yi, bool upper, Hull& result, V3& offset) {
    const double es = hf.elementSize;
    const double minValue = hf.minValue;
    double hmin = std::fmin(std::fmin(hf.h(xi, yi), hf.h(xi + 1, yi)), std::fmin(hf.h(xi, yi + 1), hf.h(xi + 1, yi + 1)));
    const double h = (hmin - minValue) / 2 + minValue;
    result.vertices.resize(6);
    std::vector<V3>& v = result.vertices;
    if (!upper) {
      offset = v3((xi + 0.25) * es, (yi + 0.25) * es, h);
      v[0] = v3(-0.25 * es, -0.25 * es, hf.h(xi, yi) - h);
      v[1] = v3(0.75 * es, -0.25 * es, hf.h(xi + 1, yi) - h);
      v[2] = v3(-0.25 * es, 0.75 * es, hf.h(xi, yi + 1) - h);
      v[3] = v3(-0.25 * es, -0.25 * es, -h - 1);
      v[4] = v3(0.75 * es, -0.25 * es, -h - 1);
      v[5] = v3(-0.25 * es, 0.75 * es, -h - 1);
      result.faces = {{0, 1, 2}, {5, 4, 3}, {0, 2, 5, 3}, {1, 0, 3, 4}, {4, 5, 2, 1}};
    } else {
      offset = v3((xi + 0.75) * es, (yi + 0.75) * es, h);
      v[0] = v3(0.25 * es, 0.25 * es, hf.h(xi + 1, yi + 1) - h);
      v[1] = v3(-0.75 * es, 0.25 * es, hf.h(xi, yi + 1) - h);
      v[2] = v3(0.25 * es, -0.75 * es, hf.h(xi + 1, yi) - h);
      v[3] = v3(0.25 * es, 0.25 * es, -h - 1);
      v[4] = v3(-0.75 * es, 0.25 * es, -h - 1);
      v[5] = v3(0.25 * es, -0.75 * es, -h - 1);
      result.faces = {{0, 1, 2}, {5, 4, 3}, {2, 5, 3, 0}, {3, 4, 1, 0}, {1, 4, 5, 2}};
    }
    result.hasUniqueAxes = false;  // a plain ConvexPolyhedron() (heightfield.dart:49,343)
    result.computeNormals();
    result.computeEdges();
    result.updateBoundingSphereRadius();
  }

  // common index-window prologue of sphereHeightfield :1318-1362 / heightfieldConvex :2070-2113
  static bool hfWindow(const Shape& hf, const V3& local, double radius, int& iMinX, int& iMaxX, int& iMinY, int& iMaxY) {
    const double wd = hf.elementSize;
    const int nx = hf.nx, ny = hf.ny;
    iMinX = (int)std::floor((D(local.x) - radius) / wd) - 1;
    iMaxX = (int)std::ceil((D(local.x) + radius) / wd) + 1;
    iMinY = (int)std::floor((D(local.y) - radius) / wd) - 1;
    iMaxY = (int)std::ceil((D(local.y) + radius) / wd) + 1;
    if (iMaxX < 0 || iMaxY < 0 || iMinX > nx || iMinY > ny) return false;
    if (iMinX < 0) iMinX = 0;
    if (iMaxX < 0) iMaxX = 0;
    if (iMinY < 0) iMinY = 0;
    if (iMaxY < 0) iMaxY = 0;
    if (iMinX >= nx) iMinX = nx - 1;
    if (iMaxX >= nx) iMaxX = nx - 1;
    if (iMaxY >= ny) iMaxY = ny - 1;
    if (iMinY >= ny) iMinY = ny - 1;
    // getRectMinMax, heightfield.dart:146-162: min is the global minValue
    double mx = hf.minValue;
    for (int i = iMinX; i <= iMaxX; i++)
      for (int j = iMinY; j <= iMaxY; j++)
        if (hf.h(i, j) > mx) mx = hf.h(i, j);
    double mn = hf.minValue;
    if (D(local.z) - radius > mx || D(local.z) + radius < mn) return false;
    return true;
  }

  // sphereHeightfield, narrow_phase.dart:1295-1437
  void sphereHeightfield(const Shape& si, const Shape& sj, V3 xi, V3 xj, Q4 qj, int bi, int bj) {
    V3 local = point_to_local_frame(xj, qj, xi);
    int iMinX, iMaxX, iMinY, iMaxY;
    if (!hfWindow(sj, local, si.radius, iMinX, iMaxX, iMinY, iMaxY)) return;
    Hull pc;
    V3 po;
    for (int i = iMinX; i < iMaxX; i++)
      for (int j = iMinY; j < iMaxY; j++) {
        size_t numContactsBefore = w.contacts.size();
        for (int up = 0; up < 2; up++) {
          pillar(sj, i, j, up != 0, pc, po);
          V3 worldPillarOffset = point_to_world_frame(xj, qj, po);
          if (distance_to(xi, worldPillarOffset) < pc.boundingSphereRadius + si.boundingSphereRadius) {
            manifold++;
            sphereConvex(si, pc, true, xi, worldPillarOffset, qj, bi, bj);
          }
        }
        if (w.contacts.size() - numContactsBefore > 2) return;
      }
  }

  // heightfieldConvex / boxHeightfield, narrow_phase.dart:2044-2178 / 1749-1766
  void heightfieldConvex(const Hull& si, bool crA, const Shape& sj, V3 xi, V3 xj, Q4 qi, Q4 qj, int bi, int bj) {
    const double radius = si.boundingSphereRadius;
    V3 local = point_to_local_frame(xj, qj, xi);
    int iMinX, iMaxX, iMinY, iMaxY;
    if (!hfWindow(sj, local, radius, iMinX, iMaxX, iMinY, iMaxY)) return;
    static const int faceList[1] = {0};
    Hull pc;
    V3 po;
    for (int i = iMinX; i < iMaxX; i++)
      for (int j = iMinY; j < iMaxY; j++)
        for (int up = 0; up < 2; up++) {
          pillar(sj, i, j, up != 0, pc, po);
          V3 worldPillarOffset = point_to_world_frame(xj, qj, po);
          if (distance_to(xi, worldPillarOffset) < pc.boundingSphereRadius + si.boundingSphereRadius) {
            manifold++;
            convexConvex(si, pc, crA, true, xi, worldPillarOffset, qi, qj, bi, bj, faceList, 1);
          }
        }
  }

  // heightfieldParticle, narrow_phase.dart:2343-2444. (si, xi, qi, bi) = the heightfield, (xj, bj) = the particle.
  void heightfieldParticle(Shape& si, const Shape& sj, V3 xi, V3 xj, Q4 qi, int bi, int bj) {
    const double radius = sj.boundingSphereRadius;
    V3 local = point_to_local_frame(xi, qi, xj);
    int iMinX, iMaxX, iMinY, iMaxY;
    if (!hfWindow(si, local, radius, iMinX, iMaxX, iMinY, iMaxY)) return;
    Hull pc;
    V3 po;
    for (int i = iMinX; i < iMaxX; i++)
      for (int j = iMinY; j < iMaxY; j++)
        for (int up = 0; up < 2; up++) {
          pillar(si, i, j, up != 0, pc, po);
          V3 worldPillarOffset = point_to_world_frame(xi, qi, po);
          if (distance_to(xj, worldPillarOffset) < pc.boundingSphereRadius) {
            manifold++;
            const long long key = (((long long)i * si.ny + j) << 1) | up;
            auto it = si.pillarWorld.find(key);
            bool needsUpdate = it == si.pillarWorld.end();
            Shape::PillarWorld fresh;
            Shape::PillarWorld& pw = needsUpdate ? fresh : it->second;
            particleConvex(pc, true, sj.collisionResponse, pw.worldVertices, pw.worldFaceNormals, needsUpdate, worldPillarOffset, xj, qi, bi, bj);
            if (it == si.pillarWorld.end() && !needsUpdate) si.pillarWorld[key] = fresh;
          }
        }
  }

  // Ray.pointInTriangle, ray_class.dart:697-709
  static bool pointInTriangle(const V3& p, const V3& a, const V3& b, const V3& c) {
    const V3 v0 = sub(c, a), v1 = sub(b, a), v2 = sub(p, a);
    const double dot00 = dot(v0, v0), dot01 = dot(v0, v1), dot02 = dot(v0, v2), dot11 = dot(v1, v1), dot12 = dot(v1, v2);
    const double u = dot11 * dot02 - dot01 * dot12;
    const double v = dot00 * dot12 - dot01 * dot02;
    return u >= 0 && v >= 0 && (u + v) < (dot00 * dot11 - dot01 * dot01);
  }

  // sphereTrimesh, narrow_phase.dart:1438-1691 as the Dart port runs it: every triangle in index order (the octree query
  // of :1480 is not used), and per corner j a vertex test, an edge test and the triangle-face test (:1573-1603 sits inside
  // the j loop, so a face contact appears three times)
  void sphereTrimesh(const Shape& si, const Shape& sj, V3 xi, V3 xj, Q4 qj, int bi, int bj) {
    const V3 local = point_to_local_frame(xj, qj, xi);
    const double radiusSquared = si.radius * si.radius;
    const int nT = (int)sj.tmIdx.size() / 3;
    auto emitLocal = [&](V3 tmp) {  // :1549-1566 / :1586-1601
      Eq r = createContactEquation(bi, bj, si.collisionResponse, sj.collisionResponse);
      r.ni = sub(tmp, local);
      normalize(r.ni);
      r.ri = scale(si.radius, r.ni);
      r.ri = add(r.ri, xi);
      r.ri = sub(r.ri, w.bodies[bi].position);
      tmp = point_to_world_frame(xj, qj, tmp);
      r.rj = sub(tmp, w.bodies[bj].position);
      r.ni = qvmult(qj, r.ni);
      r.ri = qvmult(qj, r.ri);
      addContact(r);
    };
    for (int i = 0; i < nT; i++)
      for (int j = 0; j < 3; j++) {
        {
          V3 v = sj.tmVerts[sj.tmIdx[i * 3 + j]];
          V3 relpos = sub(v, local);
          if (length2(relpos) <= radiusSquared) {
            v = point_to_world_frame(xj, qj, v);
            relpos = sub(v, xi);
            Eq r = createContactEquation(bi, bj, si.collisionResponse, sj.collisionResponse);
            r.ni = relpos;
            normalize(r.ni);
            r.ri = scale(si.radius, r.ni);
            r.ri = add(r.ri, xi);
            r.ri = sub(r.ri, w.bodies[bi].position);
            r.rj = sub(v, w.bodies[bj].position);
            addContact(r);
          }
          const V3 edgeVertexA = sj.tmVerts[sj.tmIdx[i * 3 + j]], edgeVertexB = sj.tmVerts[sj.tmIdx[i * 3 + ((j + 1) % 3)]];
          const V3 edgeVector = sub(edgeVertexB, edgeVertexA);
          V3 tmp = sub(local, edgeVertexB);
          const double positionAlongEdgeB = dot(tmp, edgeVector);
          tmp = sub(local, edgeVertexA);
          double positionAlongEdgeA = dot(tmp, edgeVector);
          if (positionAlongEdgeA > 0 && positionAlongEdgeB < 0) {
            tmp = sub(local, edgeVertexA);
            V3 edgeVectorUnit = edgeVector;
            normalize(edgeVectorUnit);
            positionAlongEdgeA = dot(tmp, edgeVectorUnit);
            tmp = scale(positionAlongEdgeA, edgeVectorUnit);
            tmp = add(tmp, edgeVertexA);
            const double dist = distance_to(tmp, local);
            if (dist < si.radius) emitLocal(tmp);
          }
        }
        {
          const V3 &va = sj.tmVerts[sj.tmIdx[i * 3]], &vb = sj.tmVerts[sj.tmIdx[i * 3 + 1]], &vc = sj.tmVerts[sj.tmIdx[i * 3 + 2]];
          const V3& normal = sj.tmNormals[i];
          V3 tmp = sub(local, va);
          double dist = dot(tmp, normal);
          tmp = scale(dist, normal);
          tmp = sub(local, tmp);
          dist = distance_to(tmp, local);
          if (pointInTriangle(tmp, va, vb, vc) && dist < si.radius) emitLocal(tmp);
        }
      }
  }

  // planeTrimesh, narrow_phase.dart:1916-1980
  void planeTrimesh(const Shape& planeShape, const Shape& sj, V3 planePos, V3 xj, Q4 planeQuat, Q4 qj, int planeBody, int bj) {
    const V3 normal = qvmult(planeQuat, V3{0, 0, 1});
    for (size_t i = 0; i < sj.tmVerts.size(); i++) {
      const V3 v = point_to_world_frame(xj, qj, sj.tmVerts[i]);
      const V3 relpos = sub(v, planePos);
      if (dot(normal, relpos) <= 0.0) {
        Eq r = createContactEquation(planeBody, bj, planeShape.collisionResponse, sj.collisionResponse);
        r.ni = normal;
        V3 projected = scale(dot(relpos, normal), normal);
        projected = sub(v, projected);
        r.ri = sub(projected, w.bodies[planeBody].position);
        r.rj = sub(v, w.bodies[bj].position);
        addContact(r);
      }
    }
  }

  // box, convex, cylinder, capsule, cone, sizedPlane: every one of them reaches the convex resolvers (narrow_phase.dart:131-184)
  static bool isHullType(int t) { return t >= CANNON_SHAPE_BOX && t <= CANNON_SHAPE_SIZED_PLANE; }

  // dispatch: getCollisionType + operator[] (narrow_phase.dart:116-238,336-489) for the in-scope types.
  // (sa,xa,qa,ba) has the lower ShapeType index; equal types arrive swapped (narrow_phase.dart:706-710).
  void resolve(Shape& sa, Shape& sb, V3 xa, V3 xb, Q4 qa, Q4 qb, int ba, int bb) {
    const int ta = sa.type, tb = sb.type;
    manifold++;
    resBody[0] = ba; resBody[1] = bb;
    resMat[0] = ta == CANNON_SHAPE_HEIGHTFIELD ? -1 : sa.material;  // the pillar convex carries no material
    resMat[1] = tb == CANNON_SHAPE_HEIGHTFIELD ? -1 : sb.material;
    if (tb == CANNON_SHAPE_TRIMESH) {  // narrow_phase.dart:212-238; heightfield-trimesh has no key
      if (ta == CANNON_SHAPE_SPHERE) sphereTrimesh(sa, sb, xa, xb, qb, ba, bb);
      else if (ta == CANNON_SHAPE_PLANE) planeTrimesh(sa, sb, xa, xb, qa, qb, ba, bb);
      else if (ta != CANNON_SHAPE_HEIGHTFIELD) w.unsupportedPair = true;  // boxTrimesh / trimeshConvex / particleTrimesh / trimeshTrimesh: unfinished in the reference
      return;
    }
    if (tb == CANNON_SHAPE_PARTICLE) {  // narrow_phase.dart:196-211; particle-particle has no key
      if (ta == CANNON_SHAPE_SPHERE) sphereParticle(sa, sb, xa, xb, ba, bb);
      else if (ta == CANNON_SHAPE_PLANE) planeParticle(sa, sb, xa, xb, ba, bb);
      else if (isHullType(ta)) particleConvex(sa.hull, sa.collisionResponse, sb.collisionResponse, sa.hull.worldVertices, sa.hull.worldFaceNormals, sa.hull.worldNeedsUpdate, xa, xb, qa, ba, bb);
      else if (ta == CANNON_SHAPE_HEIGHTFIELD) heightfieldParticle(sa, sb, xa, xb, qa, ba, bb);
      return;
    }
    if (ta == CANNON_SHAPE_SPHERE) {
      if (tb == CANNON_SHAPE_SPHERE) sphereSphere(sa, sb, xa, xb, ba, bb);
      else if (tb == CANNON_SHAPE_PLANE) spherePlane(sa, sb, xa, xb, qb, ba, bb);
      else if (tb == CANNON_SHAPE_BOX) sphereBox(sa, sb, xa, xb, qb, ba, bb);
      else if (isHullType(tb)) sphereConvex(sa, sb.hull, sb.collisionResponse, xa, xb, qb, ba, bb);
      else if (tb == CANNON_SHAPE_HEIGHTFIELD) sphereHeightfield(sa, sb, xa, xb, qb, ba, bb);
    } else if (ta == CANNON_SHAPE_PLANE) {
      if (isHullType(tb)) planeConvex(sa, sb.hull, sb.collisionResponse, xa, xb, qa, qb, ba, bb);
      // plane-plane, plane-heightfield: no resolver
    } else if (isHullType(ta)) {
      // narrow_phase.dart:448 spells the key "convexSizedPlane"; getCollisionType (:476-479) compares lower-cased names,
      // so a plain convex against a sized plane finds no resolver
      if (ta == CANNON_SHAPE_CONVEX && tb == CANNON_SHAPE_SIZED_PLANE) return;
      if (isHullType(tb)) convexConvex(sa.hull, sb.hull, sa.collisionResponse, sb.collisionResponse, xa, xb, qa, qb, ba, bb, nullptr, 0);
      else if (tb == CANNON_SHAPE_HEIGHTFIELD) heightfieldConvex(sa.hull, sa.collisionResponse, sb, xa, xb, qa, qb, ba, bb);
    }
    // heightfield-heightfield: no resolver
  }
};

}  // namespace

const cannon_contact_material* World::contactMaterial(int ma, int mb) const {
  if (ma < 0 || mb < 0) return nullptr;
  int n = (int)matFriction.size();
  if (ma >= n || mb >= n) return nullptr;
  int idx = cmTable[(size_t)ma * n + mb];
  return idx >= 0 ? &cms[idx] : nullptr;
}

// Narrowphase.getContacts, narrow_phase.dart:634-721
void World::getContacts() {
  contacts.clear();
  frictions.clear();
  contactManifold.clear();
  justTestOverlaps.clear();
  perPairCount.assign(p1.size(), 0);
  NP np(*this);
  for (size_t k = 0; k != p1.size(); k++) {
    const int bi = p1[k], bj = p2[k];
    const Body& A = bodies[bi];
    const Body& B = bodies[bj];
    size_t before = contacts.size();
    const cannon_contact_material* bodyCm = contactMaterial(A.material, B.material);
    const bool justTest = (A.type == CANNON_BODY_KINEMATIC && B.type == CANNON_BODY_STATIC) ||
                          (A.type == CANNON_BODY_STATIC && B.type == CANNON_BODY_KINEMATIC) ||
                          (A.type == CANNON_BODY_KINEMATIC && B.type == CANNON_BODY_KINEMATIC);
    for (size_t i = 0; i < A.shapes.size(); i++) {      // narrow_phase.dart:670-721
      Shape& si = shapes[A.shapes[i]];
      const Q4 qi = qmul(A.quaternion, A.shapeOrientations[i]);
      const V3 xi = add(qvmult(A.quaternion, A.shapeOffsets[i]), A.position);
      for (size_t j = 0; j < B.shapes.size(); j++) {
        Shape& sj = shapes[B.shapes[j]];
        const Q4 qj = qmul(B.quaternion, B.shapeOrientations[j]);
        const V3 xj = add(qvmult(B.quaternion, B.shapeOffsets[j]), B.position);
        if (!((si.mask & sj.group) != 0 && (sj.mask & si.group) != 0)) continue;
        if (distance_to(xi, xj) > si.boundingSphereRadius + sj.boundingSphereRadius) continue;
        const cannon_contact_material* shapeCm = (si.material >= 0 && sj.material >= 0) ? contactMaterial(si.material, sj.material) : nullptr;  // :692-696
        np.cm = shapeCm ? shapeCm : (bodyCm ? bodyCm : &desc.default_contact_material);
        np.rsiMat = si.material; np.rsjMat = sj.material;
        if (justTest) {
          // justTest mode (:706-716): no equations; the resolver returns true where it would have created its first contact
          // (every `if (justTest) return true` sits right before a createContactEquation), and the world only uses that to
          // feed the overlap keepers. sphereSphere has its own test (:735-737).
          if (!trackOverlaps) continue;
          bool hit;
          if (si.type == CANNON_SHAPE_SPHERE && sj.type == CANNON_SHAPE_SPHERE) {
            const double dx = D(xj.x) - D(xi.x), dy = D(xj.y) - D(xi.y), dz = D(xj.z) - D(xi.z);  // Vector3.distanceSquared, vec3.dart:88-93
            const double rs = si.radius + sj.radius;
            hit = dx * dx + dy * dy + dz * dz < rs * rs;
          } else {
            const size_t nc = contacts.size(), nf = frictions.size(), nm = contactManifold.size();
            if (si.type < sj.type) np.resolve(si, sj, xi, xj, qi, qj, bi, bj);
            else np.resolve(sj, si, xj, xi, qj, qi, bj, bi);
            hit = contacts.size() > nc;
            contacts.resize(nc); frictions.resize(nf); contactManifold.resize(nm);
          }
          if (hit) { justTestOverlaps.push_back(bi); justTestOverlaps.push_back(bj); }
          continue;
        }
        if (si.type < sj.type) np.resolve(si, sj, xi, xj, qi, qj, bi, bj);
        else np.resolve(sj, si, xj, xi, qj, qi, bj, bi);
      }
    }
    perPairCount[k] = (int)(contacts.size() - before);
  }
}

}  // namespace orc
