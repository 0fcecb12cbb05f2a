/*
 * oracle_broadphase.cpp — TEST INFRASTRUCTURE ONLY (PARITY UNPINNED, see oracle_math.h).
 * Broadphase pair finding restated from
 *   lib/collision/broadphase.dart:44-102        needBroadphaseCollision / intersection tests
 *   lib/collision/naive_broadphase.dart:14-33   NaiveBroadphase.collisionPairs
 *   lib/collision/sap_broadphase.dart:41-189    SAPBroadphase (persistent axisList, insertion sort, sweep)
 *   lib/collision/grid_broadphase.dart:59-239   GridBroadphase (INTENDED semantics: the code throws as
 *                                               written; documented fix below)
 *   lib/collision/aabb.dart:131-147             AABB.overlaps
 *   lib/world/world_class.dart:488-499          constraint-pair filter
 */
#include <algorithm>
#include <cmath>
#include <set>
#include <utility>

#include "oracle_world.h"

namespace orc {

// broadphase.dart:44-63
static bool needBroadphaseCollision(const Body& a, const Body& b) {
  if ((a.group & b.mask) == 0 || (b.group & a.mask) == 0) return false;
  if ((a.type == CANNON_BODY_STATIC || a.sleepState == CANNON_SLEEPING) &&
      (b.type == CANNON_BODY_STATIC || b.sleepState == CANNON_SLEEPING))
    return false;
  return true;
}

// aabb.dart:131-147
static bool aabbOverlaps(const Body& A, const Body& B) {
  const V3 &l1 = A.aabbLower, &u1 = A.aabbUpper, &l2 = B.aabbLower, &u2 = B.aabbUpper;
  bool ox = (l2.x <= u1.x && u1.x <= u2.x) || (l1.x <= u2.x && u2.x <= u1.x);
  bool oy = (l2.y <= u1.y && u1.y <= u2.y) || (l1.y <= u2.y && u2.y <= u1.y);
  bool oz = (l2.z <= u1.z && u1.z <= u2.z) || (l1.z <= u2.z && u2.z <= u1.z);
  return ox && oy && oz;
}

// broadphase.dart:67-102
static void intersectionTest(const World& w, int ia, int ib, std::vector<int>& p1, std::vector<int>& p2) {
  const Body& A = w.bodies[ia];
  const Body& B = w.bodies[ib];
  if (w.desc.use_bounding_boxes) {
    if (aabbOverlaps(A, B)) {
      p1.push_back(ia);
      p2.push_back(ib);
    }
  } else {
    V3 r = sub(B.position, A.position);
    double s = A.boundingRadius + B.boundingRadius;
    double boundingRadiusSum2 = s * s;  // math.pow(x, 2)
    double norm2 = length2(r);
    if (norm2 < boundingRadiusSum2) {
      p1.push_back(ia);
      p2.push_back(ib);
    }
  }
}

static long safe_floor(double v) {
  if (!(v > -1e15)) return -(long)1e15;
  if (!(v < 1e15)) return (long)1e15;
  return (long)std::floor(v);
}
static long safe_ceil(double v) {
  if (!(v > -1e15)) return -(long)1e15;
  if (!(v < 1e15)) return (long)1e15;
  return (long)std::ceil(v);
}

void World::collisionPairs() {
  p1.clear();
  p2.clear();
  const int N = (int)bodies.size();
  const bool needAABB = desc.use_bounding_boxes || desc.broadphase_kind != CANNON_BP_NAIVE;
  if (needAABB)
    for (Body& b : bodies) updateAABB(b);

  if (desc.broadphase_kind == CANNON_BP_NAIVE) {
    // naive_broadphase.dart:21-32: for i, for j<i -> (bodies[i], bodies[j])
    if (desc.n_worlds > 1) {
      // batch of independent worlds: each world is its own World object in the reference, so pairs
      // are only formed among bodies of one world (bodies of a world are contiguous)
      for (int i = 0; i < N; i++)
        for (int j = 0; j < i; j++) {
          if (bodies[i].worldId != bodies[j].worldId) continue;
          if (!needBroadphaseCollision(bodies[i], bodies[j])) continue;
          intersectionTest(*this, i, j, p1, p2);
        }
    } else {
      for (int i = 0; i < N; i++)
        for (int j = 0; j < i; j++) {
          if (!needBroadphaseCollision(bodies[i], bodies[j])) continue;
          intersectionTest(*this, i, j, p1, p2);
        }
    }
  } else if (desc.broadphase_kind == CANNON_BP_SAP) {
    // sap_broadphase.dart:138-189
    if ((int)sapAxisList.size() != N) {
      sapAxisList.resize(N);
      for (int i = 0; i < N; i++) sapAxisList[i] = i;
    }
    const int axis = desc.sap_axis;
    auto lower = [&](int b) -> float {
      const V3& l = bodies[b].aabbLower;
      return axis == 0 ? l.x : (axis == 1 ? l.y : l.z);
    };
    // insertionSortX/Y/Z, sap_broadphase.dart:66-111
    std::vector<int>& a = sapAxisList;
    for (int i = 1; i < N; i++) {
      int v = a[i];
      int j;
      for (j = i - 1; j >= 0; j--) {
        if (lower(a[j]) <= lower(v)) break;
        a[j + 1] = a[j];
      }
      a[j + 1] = v;
    }
    auto pos = [&](int b) -> double {
      const V3& p = bodies[b].position;
      return axis == 0 ? D(p.x) : (axis == 1 ? D(p.y) : D(p.z));
    };
    for (int i = 0; i < N; i++) {
      int bi = a[i];
      for (int j = i + 1; j < N; j++) {
        int bj = a[j];
        if (!needBroadphaseCollision(bodies[bi], bodies[bj])) continue;
        // checkBounds, sap_broadphase.dart:41-62
        double boundA2 = pos(bi) + bodies[bi].boundingRadius;
        double boundB1 = pos(bj) - bodies[bj].boundingRadius;
        if (!(boundB1 < boundA2)) break;
        intersectionTest(*this, bi, bj, p1, p2);
      }
    }
  } else {
    // GridBroadphase, grid_broadphase.dart:59-239, intended semantics:
    //  - bins really are per-bin lists; +-inf AABB bounds clamp to the grid;
    //  - a pair is reported once iff the bodies share >= 1 bin AND pass needBroadphaseCollision AND
    //    the bounding-sphere / AABB test;
    //  - DOCUMENTED DEVIATION: makePairsUnique cannot run as written, so the order of the unique
    //    pairs is defined as the NaiveBroadphase order (i descending major... i.e. (i, j<i), i ascending).
    const int nx = desc.grid_nx, ny = desc.grid_ny, nz = desc.grid_nz;
    const int xstep = ny * nz, ystep = nz, zstep = 1;
    const double xmax = D(desc.grid_max[0]), ymax = D(desc.grid_max[1]), zmax = D(desc.grid_max[2]);
    const double xmin = D(desc.grid_min[0]), ymin = D(desc.grid_min[1]), zmin = D(desc.grid_min[2]);
    const double xmult = nx / (xmax - xmin), ymult = ny / (ymax - ymin), zmult = nz / (zmax - zmin);
    const double binsizeX = (xmax - xmin) / nx, binsizeY = (ymax - ymin) / ny, binsizeZ = (zmax - zmin) / nz;
    const double binRadius = std::sqrt(binsizeX * binsizeX + binsizeY * binsizeY + binsizeZ * binsizeZ) * 0.5;
    const int nBins = nx * ny * nz;
    std::vector<std::vector<int>> bins(nBins);

    auto clampi = [](long v, int n) -> int { return v < 0 ? 0 : (v >= n ? n - 1 : (int)v); };
    auto addBoxToBins = [&](double x0, double y0, double z0, double x1, double y1, double z1, int bi) {
      int xoff0 = clampi(safe_floor((x0 - xmin) * xmult), nx) * xstep;
      int yoff0 = clampi(safe_floor((y0 - ymin) * ymult), ny) * ystep;
      int zoff0 = clampi(safe_floor((z0 - zmin) * zmult), nz) * zstep;
      int xoff1 = clampi(safe_ceil((x1 - xmin) * xmult), nx) * xstep;
      int yoff1 = clampi(safe_ceil((y1 - ymin) * ymult), ny) * ystep;
      int zoff1 = clampi(safe_ceil((z1 - zmin) * zmult), nz) * zstep;
      for (int xoff = xoff0; xoff <= xoff1; xoff += xstep)
        for (int yoff = yoff0; yoff <= yoff1; yoff += ystep)
          for (int zoff = zoff0; zoff <= zoff1; zoff += zstep) bins[xoff + yoff + zoff].push_back(bi);
    };

    for (int i = 0; i < N; i++) {
      const Body& bi = bodies[i];
      if (bi.shape < 0) continue;
      const Shape& si = shapes[bi.shape];
      if (si.type == CANNON_SHAPE_SPHERE) {
        double x = D(bi.position.x), y = D(bi.position.y), z = D(bi.position.z), r = si.radius;
        addBoxToBins(x - r, y - r, z - r, x + r, y + r, z + r, i);
      } else if (si.type == CANNON_SHAPE_PLANE) {
        V3 planeNormal = qvmult(bi.quaternion, V3{0, 0, 1});
        double xreset = xmin + binsizeX * 0.5 - D(bi.position.x);
        double yreset = ymin + binsizeY * 0.5 - D(bi.position.y);
        double zreset = zmin + binsizeZ * 0.5 - D(bi.position.z);
        V3 d = v3(xreset, yreset, zreset);  // a Vector3: float storage, increments round each time
        for (int xi = 0, xoff = 0; xi != nx; xi++, xoff += xstep, d.y = (float)yreset, d.x = (float)(D(d.x) + binsizeX))
          for (int yi = 0, yoff = 0; yi != ny; yi++, yoff += ystep, d.z = (float)zreset, d.y = (float)(D(d.y) + binsizeY))
            for (int zi = 0, zoff = 0; zi != nz; zi++, zoff += zstep, d.z = (float)(D(d.z) + binsizeZ))
              if (dot(d, planeNormal) < binRadius) bins[xoff + yoff + zoff].push_back(i);
      } else {
        addBoxToBins(D(bi.aabbLower.x), D(bi.aabbLower.y), D(bi.aabbLower.z), D(bi.aabbUpper.x), D(bi.aabbUpper.y),
                     D(bi.aabbUpper.z), i);
      }
    }
    std::set<std::pair<int, int>> seen;
    std::vector<int> q1, q2;
    for (int b = 0; b < nBins; b++) {
      const std::vector<int>& bin = bins[b];
      for (size_t xi = 0; xi < bin.size(); xi++)
        for (size_t yi = 0; yi < xi; yi++) {
          int bi = bin[xi], bj = bin[yi];
          if (!needBroadphaseCollision(bodies[bi], bodies[bj])) continue;
          if (seen.count({bi, bj})) continue;
          q1.clear();
          q2.clear();
          intersectionTest(*this, bi, bj, q1, q2);
          if (!q1.empty()) seen.insert({bi, bj});
        }
    }
    for (const auto& pr : seen) {  // std::set order == (i ascending, j ascending) == Naive order
      p1.push_back(pr.first);
      p2.push_back(pr.second);
    }
  }

  // world_class.dart:488-499
  for (const Constraint& c : constraints) {
    if (c.collideConnected) continue;
    for (int j = (int)p1.size() - 1; j >= 0; j--) {
      if ((c.bodyA == p1[j] && c.bodyB == p2[j]) || (c.bodyB == p1[j] && c.bodyA == p2[j])) {
        p1.erase(p1.begin() + j);
        p2.erase(p2.begin() + j);
      }
    }
  }
}

}  // namespace orc
