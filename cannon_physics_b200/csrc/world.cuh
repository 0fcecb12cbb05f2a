// world.cuh — device-resident data model of one cannon_world and the host-side
// shape flattening (Shape / ConvexPolyhedron objects -> hull tables).
//
// HBM layout: one array per attribute ("SoA of float4"): every 3-/4-vector attribute is a float4 per
// body so both the streaming kernels (one 128-bit load per thread, fully coalesced) and the gather
// kernels (narrowphase / solver: one 128-bit transaction per body attribute) are vectorised.
#pragma once
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "dmath.cuh"

// body flag bits
#define BF_ALLOW_SLEEP 1
#define BF_COLLISION_RESPONSE 2
#define BF_IS_TRIGGER 4
#define BF_FIXED_ROTATION 8
#define BF_BIG 16        // bounding radius too large for the uniform grid: handled by the big-body path
#define BF_WAKE 32       // Body.wakeUpAfterNarrowphase

struct BodyArrays {
  float4 *pos, *quat, *vel, *angvel, *force, *torque, *vlam, *wlam;
  float4 *iiw0, *iiw1, *iiw2;  // invInertiaWorld rows
  float4 *invI, *linF, *angF;  // local inverse inertia diagonal, linearFactor, angularFactor
  float4 *aabbLo, *aabbHi;
  double *mass, *invMass, *brad, *ldamp, *adamp, *ldpow, *adpow, *sleepSpeed, *sleepTime, *tLastSleepy;
  int *type, *sleep, *shape, *material, *group, *mask, *world, *flags;
  // Narrowphase view of a world with compound bodies (cannon_world_set_body_shapes): the arrays above that the
  // narrowphase reads (pos, quat, shape, type, flags, material, invMass) are then per SHAPE INSTANCE ("proxy"), owner[]
  // maps a proxy to its body and bpos / bquat are the real per-body poses (indexed by body). Otherwise owner == nullptr
  // and bpos / bquat alias pos / quat.
  const float4 *bpos, *bquat;
  const int* owner;
};
__device__ __forceinline__ int np_owner(const BodyArrays& B, int i) { return B.owner ? B.owner[i] : i; }

struct ShapeDev {
  int type, collisionResponse, group, mask;
  double radius, bsr;
  float hx, hy, hz;
  int hull;  // index into hull table (box / convex / cylinder), -1 otherwise
  int hf;    // index into heightfield table, -1 otherwise
  int tm;    // index into trimesh table, -1 otherwise
  int material, pad;  // Shape.material (index into the material table) or -1
};

struct HullDev {
  int vOff, nV;       // vertices
  int fOff, nF;       // faces: CSR over fvOff[fOff + f .. fOff + f + 1]
  int eOff, nE;       // unique edges
  int hasAxes;        // `uniqueAxes != null` (convex_polyhedron.dart:253,290)
  int pad;
  double bsr;         // boundingSphereRadius
  int ekOff, nEk;     // unique edges that are not +-copies of an earlier one (SAT axis pruning, k_sat_warp.cuh)
  int fkOff, nFk;     // likewise for the face normals: list of face indices
};

// Trimesh (trimesh.dart): getVertex results, triangle indices, face normals, local AABB
struct TrimeshDev {
  int vOff, nV, iOff, nT;  // vertices in tmVerts, 3 indices per triangle in tmIdx (local to the mesh), normals at tmNormals[iOff / 3 + t]
  float4 lo, hi;           // computeLocalAABB, trimesh.dart:315-343
};

struct HfDev {
  int nx, ny, esize, dataOff;
  double minV, maxV;
  long long pilOff;   // first PillarRec of this heightfield: ((xi * (ny - 1) + yi) * 2 + upper)
};

// Heightfield.getConvexTrianglePillar (heightfield.dart:330-487) evaluated once per (cell, lower/upper) when the
// shapes are set - the reference caches pillars too (getCachedConvexTrianglePillar, heightfield.dart:300-328); the
// height samples are immutable through this ABI. e[] holds the unique edges minus +-copies, in their original order.
struct PillarRec {
  // first 128 bytes = one cache line: everything the task expansion (bounding gate, quick separation) reads per pillar
  float4 off;      // pillar offset in the heightfield frame
  float4 v[6];
  double bsr;
  int nE, pad;
  // the resolvers' part
  float4 n[5];
  float4 e[9];
  double pc[5];
};
static_assert(sizeof(PillarRec) == 400 && offsetof(PillarRec, n) == 128, "PillarRec: the gate fields must fill the first cache line");

struct ShapeTables {
  const ShapeDev* shapes;
  const HullDev* hulls;
  const float4* verts;      // hull vertices
  const float4* fnormals;   // face normals (indexed fOff + f)
  const double* fplanec;    // -n . v0 per face (getPlaneConstantOfFace)
  const int* fvOff;         // per face: offset into fvIdx (size totalFaces + nHulls, CSR per hull)
  const int* fvIdx;         // face vertex indices (local to the hull)
  const int* fcOff;         // per face: offset into fcIdx (connected faces, precomputed clipFaceAgainstHull :459-473)
  const int* fcIdx;
  const float4* edges;      // unique edges
  const float4* edgesK;     // unique edges without +-copies (HullDev.ekOff)
  const int* facesK;        // faces whose normal is not a +-copy of an earlier one (HullDev.fkOff)
  const PillarRec* pillars; // precomputed triangle pillars of every heightfield (HfDev.pilOff)
  const HfDev* hfs;
  const double* hfdata;
  const int* cmTable;       // nMat*nMat -> contact material index or -1
  const cannon_contact_material* cms;
  const double* matFriction;
  const double* matRestitution;
  int nMat;
  const TrimeshDev* tms;
  const float4* tmVerts;
  const float4* tmNormals;
  const int* tmIdx;
  // compound bodies: body b owns the instances [instFirst[b], instFirst[b+1]); nullptr = one shape per body at its origin
  const int* instFirst;
  const int* instShape;
  const float4* instOff;
  const float4* instQuat;
  // particleConvex state (k_narrowphase.cuh, k_np_particle_hull): the pose a hull shape / heightfield pillar was frozen at.
  // Target t < nShapes is shape t, nShapes + p is pillar p. nullptr unless the shape table holds a Particle.
  int nShapes;
  int* pcFrozen;       // 0 until the target's first penetration
  int* pcFreezeTask;   // lowest task of this step that penetrates a not yet frozen target (0x7f7f7f7f between steps)
  float4* pcPos;
  float4* pcQuat;
};

// ---------------------------------------------------------------------------------------------------
// host-side hull construction (setup path, mirrors the reference constructors; runs once per shape)
// ---------------------------------------------------------------------------------------------------
struct HostHull {
  std::vector<f3> v;
  std::vector<std::vector<int>> faces;
  std::vector<f3> n;
  std::vector<double> planec;
  std::vector<f3> edges;
  std::vector<std::vector<int>> connected;
  bool hasAxes = false;
  double bsr = 0;

  void finish() {
    // ConvexPolyhedron.computeNormals, convex_polyhedron.dart:143-185
    n.resize(faces.size());
    for (size_t i = 0; i < faces.size(); i++) {
      const f3 &va = v[faces[i][0]], &vb = v[faces[i][1]], &vc = v[faces[i][2]];
      f3 ab = vsub(vb, va), cb = vsub(vc, vb);
      f3 nn = vcross(cb, ab);
      if (!(nn.x == 0 && nn.y == 0 && nn.z == 0)) vnormalize(nn);
      n[i] = vneg(nn);
    }
    // updateBoundingSphereRadius, :649-660
    double max2 = 0;
    for (const f3& p : v) max2 = fmax(max2, vlen2(p));
    bsr = sqrt(max2);
    // computeEdges, :110-139
    edges.clear();
    for (size_t i = 0; i < faces.size(); i++) {
      int nv = (int)faces[i].size();
      for (int j = 0; j < nv; j++) {
        f3 e = vsub(v[faces[i][j]], v[faces[i][(j + 1) % nv]]);
        vnormalize(e);
        bool found = false;
        for (const f3& u : edges)
          if (valmost_eq(u, e)) { found = true; break; }
        if (!found) edges.push_back(e);
      }
    }
    // plane constants (:405-411) and the connected-face lists clipFaceAgainstHull rebuilds per call (:459-473)
    planec.resize(faces.size());
    connected.assign(faces.size(), {});
    for (size_t f = 0; f < faces.size(); f++) {
      planec[f] = -vdot(n[f], v[faces[f][0]]);
      for (size_t i = 0; i < faces.size(); i++)
        for (size_t j = 0; j < faces[i].size(); j++) {
          bool shares = std::find(faces[f].begin(), faces[f].end(), faces[i][j]) != faces[f].end();
          if (shares && i != f && std::find(connected[f].begin(), connected[f].end(), (int)i) == connected[f].end())
            connected[f].push_back((int)i);
        }
    }
  }
};

static inline void host_box_hull(const float he[3], HostHull& h) {  // box.dart:41-86
  double sx = he[0], sy = he[1], sz = he[2];
  h.v = {mk3(-sx, -sy, -sz), mk3(sx, -sy, -sz), mk3(sx, sy, -sz), mk3(-sx, sy, -sz),
         mk3(-sx, -sy, sz),  mk3(sx, -sy, sz),  mk3(sx, sy, sz),  mk3(-sx, sy, sz)};
  h.faces = {{3, 2, 1, 0}, {4, 5, 6, 7}, {5, 4, 0, 1}, {2, 3, 7, 6}, {0, 4, 7, 3}, {1, 2, 6, 5}};
  h.hasAxes = true;
  h.finish();
}

static inline void host_cylinder_hull(double rt, double rb, double height, int N, HostHull& h) {  // cylinder.dart:21-101
  std::vector<int> bottom, top;
  h.v.clear();
  h.faces.clear();
  h.v.push_back(mk3(-rb * sin(0.0), -height * 0.5, rb * cos(0.0)));
  bottom.push_back(0);
  h.v.push_back(mk3(-rt * sin(0.0), height * 0.5, rt * cos(0.0)));
  top.push_back(1);
  for (int i = 0; i < N; i++) {
    double theta = ((2 * M_PI) / N) * (i + 1);
    if (i < N - 1) {
      h.v.push_back(mk3(-rb * sin(theta), -height * 0.5, rb * cos(theta)));
      bottom.push_back(2 * i + 2);
      h.v.push_back(mk3(-rt * sin(theta), height * 0.5, rt * cos(theta)));
      top.push_back(2 * i + 3);
      h.faces.push_back({2 * i, 2 * i + 1, 2 * i + 3, 2 * i + 2});
    } else {
      h.faces.push_back({2 * i, 2 * i + 1, 1, 0});
    }
  }
  h.faces.push_back(bottom);
  std::vector<int> rev(top.rbegin(), top.rend());
  h.faces.push_back(rev);
  h.hasAxes = true;
  h.finish();
}
