/*
 * oracle_world.h — TEST INFRASTRUCTURE ONLY. Data model of the CPU restatement (PARITY UNPINNED, see
 * oracle_math.h). Mirrors the reference's object graph (Body / Shape / ConvexPolyhedron / Equation /
 * World) closely on purpose: the product uses a completely different (SoA, device) layout.
 */
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../include/cannon_cuda.h"
#include "oracle_math.h"

namespace orc {

// ConvexPolyhedron, lib/rigid_body_shapes/convex_polyhedron.dart:51
struct Hull {
  std::vector<V3> vertices;
  std::vector<std::vector<int>> faces;
  std::vector<V3> faceNormals;
  std::vector<V3> uniqueEdges;
  bool hasUniqueAxes = false;  // `uniqueAxes != null` (convex_polyhedron.dart:253,290)
  double boundingSphereRadius = 0;
  void computeNormals();
  void computeEdges();
  void updateBoundingSphereRadius();
  double planeConstantOfFace(int f) const;
  // worldVertices / worldFaceNormals with their NeedsUpdate flags (convex_polyhedron.dart:55-58,101-103): only
  // particleConvex reads them (narrow_phase.dart:2207-2218) and nothing ever sets the flags back to true, so they hold
  // the pose of the FIRST penetration this ConvexPolyhedron object ever saw
  bool worldNeedsUpdate = true;
  std::vector<V3> worldVertices, worldFaceNormals;
};

struct Shape {
  int type = CANNON_SHAPE_SPHERE;
  bool collisionResponse = true;
  int group = -1, mask = -1;
  int material = -1;  // Shape.material (index into the material table) or -1 = null
  double boundingSphereRadius = 0;
  double radius = 1;  // sphere
  V3 halfExtents{0, 0, 0};
  Hull hull;  // box / cylinder / convex
  // heightfield
  int nx = 0, ny = 0, elementSize = 1;
  std::vector<double> data;
  double minValue = 0, maxValue = 0;
  double h(int i, int j) const { return data[(size_t)i * ny + j]; }
  // Heightfield._cachedPillars (heightfield.dart:52,285-301; cacheEnabled = true): one ConvexPolyhedron object per
  // (xi, yi, upper), kept for the life of the shape. Only the frozen world geometry of particleConvex is state.
  // trimesh (trimesh.dart): getVertex results (scale applied), triangle indices, face normals, local AABB
  std::vector<V3> tmVerts, tmNormals;
  std::vector<int> tmIdx;
  V3 tmLo{0, 0, 0}, tmHi{0, 0, 0};
  struct PillarWorld { std::vector<V3> worldVertices, worldFaceNormals; };
  std::map<long long, PillarWorld> pillarWorld;
};

struct Body {
  V3 position{0, 0, 0}, velocity{0, 0, 0}, angularVelocity{0, 0, 0}, force{0, 0, 0}, torque{0, 0, 0};
  Q4 quaternion{0, 0, 0, 1};
  V3 vlambda{0, 0, 0}, wlambda{0, 0, 0};
  double mass = 0, invMass = 0;
  int type = CANNON_BODY_STATIC;
  int sleepState = CANNON_AWAKE;
  double timeLastSleepy = 0;
  bool allowSleep = true;
  double sleepSpeedLimit = 0.1, sleepTimeLimit = 1;
  double linearDamping = 0.01, angularDamping = 0.01;
  V3 linearFactor{1, 1, 1}, angularFactor{1, 1, 1};
  bool fixedRotation = false;
  int group = 1, mask = -1;
  bool collisionResponse = true, isTrigger = false;
  int material = -1;
  int shape = -1;  // shapes[0] (or -1): what GridBroadphase looks at (grid_broadphase.dart:136)
  // Body.shapes / shapeOffsets / shapeOrientations, rigid_body.dart:96-104 (indices into World::shapes)
  std::vector<int> shapes;
  std::vector<V3> shapeOffsets;
  std::vector<Q4> shapeOrientations;
  int worldId = 0;
  bool wakeUpAfterNarrowphase = false;
  V3 inertia{0, 0, 0}, invInertia{0, 0, 0};
  M3 invInertiaWorld{{0, 0, 0, 0, 0, 0, 0, 0, 0}};
  double invMassSolve = 0;
  M3 invInertiaWorldSolve{{0, 0, 0, 0, 0, 0, 0, 0, 0}};
  double boundingRadius = 0;
  V3 aabbLower{0, 0, 0}, aabbUpper{0, 0, 0};
};

enum { EQ_CONTACT = 0, EQ_FRICTION = 1, EQ_ROTATIONAL = 2, EQ_MOTOR = 3 };

// Equation + subclasses, lib/equations/*.dart
struct Eq {
  int kind = EQ_CONTACT;
  int bi = -1, bj = -1;
  double minForce = -1e6, maxForce = 1e6;
  double a = 0, b = 0, eps = 0;
  bool enabled = true;
  double multiplier = 0;
  V3 ri{0, 0, 0}, rj{0, 0, 0}, ni{0, 0, 0};  // ni doubles as the friction tangent t
  double restitution = 0;
  V3 axisA{1, 0, 0}, axisB{0, 1, 0};
  double maxAngle = M_PI / 2;
  double targetVelocity = 0;
  double friction = 0;  // contact only: mu used for its friction equations (<=0: none)
  // jacobian elements
  V3 sA{0, 0, 0}, rA{0, 0, 0}, sB{0, 0, 0}, rB{0, 0, 0};
  void setSpookParams(double k, double d, double h) {  // equation_class.dart:53-60
    a = 4.0 / (h * (1 + 4 * d));
    b = 4.0 * d / (1 + 4 * d);
    eps = 4.0 / (h * h * k * (1 + 4 * d));
  }
};

struct Constraint {
  int type = 0, bodyA = -1, bodyB = -1;
  V3 pivotA{0, 0, 0}, pivotB{0, 0, 0}, axisA{1, 0, 0}, axisB{1, 0, 0};
  bool collideConnected = true;
  double distance = 0;  // DistanceConstraint.distance
  // P2P: x,y,z ; hinge: x,y,z,rot1,rot2,motor ; distance: d ; lock: x,y,z,r1,r2,r3 ; cone-twist: x,y,z,cone,twist
  std::vector<Eq> eqs;
  std::vector<V3> locA, locB;  // body-local axes of the rotational equations that are re-oriented in update() (lock, cone-twist)
};

struct Spring {  // lib/objects/spring.dart
  int bodyA = -1, bodyB = -1;
  double restLength = 1, stiffness = 100, damping = 1;
  V3 localAnchorA{0, 0, 0}, localAnchorB{0, 0, 0};
};

struct RayHit {  // one reported intersection (RaycastResult, lib/collision/raycast_result.dart)
  int ray, body, hitFaceIndex;
  double distance;
  V3 hitPointWorld, hitNormalWorld;
  int shapeOrdinal = -1;  // RaycastResult.shape as its position in Body.shapes
};

struct RowDebug {
  int bi, bj;
  double B, invC, lambda;
  int level = 0;  // COLORED: colour of the row's unit
};

struct World {
  cannon_world_desc desc;
  std::vector<Shape> shapes;
  std::vector<Body> bodies;
  std::vector<double> matFriction, matRestitution;
  std::vector<cannon_contact_material> cms;
  std::vector<int> cmTable;  // nmat*nmat -> cm index or -1
  std::vector<Constraint> constraints;
  std::vector<Spring> springs;  // applied in the postStep slot, in order
  double time = 0;
  double dt = -1;
  int64_t stepnumber = 0;
  std::vector<int> sapAxisList;
  // outputs of the last stages
  std::vector<int> p1, p2;
  bool unsupportedPair = false;
  std::vector<int> justTestOverlaps;  // (bi, bj) of the kinematic / static pairs whose resolver returned true (narrow_phase.dart:712-716)
  // SPHSystem, sph_system.dart (World.subsystems)
  struct Sph { std::vector<int> particles; double density = 1, smoothingRadius = 1, speedOfSound = 1, viscosity = 0.01, eps = 0.00001; };
  std::vector<Sph> sphSystems;
  void sphUpdate();  // a pair only the reference's unfinished trimesh resolvers would handle reached getContacts
  // cannon_world_set_body_shapes: the table the next set_bodies consumes
  std::vector<int> pendFirst, pendShape;
  std::vector<V3> pendOffset;
  std::vector<Q4> pendOrient;
  std::vector<Eq> contacts;   // ContactEquations (World.contacts)
  std::vector<Eq> frictions;  // FrictionEquations (World.frictionEquations)
  std::vector<int> perPairCount;
  std::vector<int> contactManifold;  // per ContactEquation: ordinal of the resolver call that created it (COLORED solver units)
  std::vector<RowDebug> rows;
  cannon_profile prof{};
  std::string err;
  // OverlapKeeper of body ids (overlap_keeper.dart): sorted key lists of this and the previous step + the last diff
  bool trackOverlaps = false;
  std::vector<int64_t> overlapCurrent, overlapPrevious;
  std::vector<int> additions, removals;  // flat (a, b) pairs, a < b, like OverlapKeeper._unpackAndPush
  void overlapSet(int i, int j);         // OverlapKeeper.set, overlap_keeper.dart:20-37
  void emitContactEvents();              // OverlapKeeper.getDiff, overlap_keeper.dart:48-82 (world_class.dart:703-708)

  const cannon_contact_material* contactMaterial(int ma, int mb) const;
  void updateAABB(Body& b) const;
  void updateMassProperties(Body& b) const;
  void updateInertiaWorld(Body& b, bool force) const;
  void updateBoundingRadius(Body& b) const;

  void aabbQuery(const V3& lower, const V3& upper, std::vector<int>& result);  // NaiveBroadphase.aabbQuery
  bool raycast(int rayIndex, const V3& from, const V3& to, const cannon_ray_options& opt, RayHit& out, std::vector<RayHit>* all);
  void collisionPairs();            // broadphase + constraint-pair filter
  void getContacts();               // narrowphase over p1/p2
  void makeContactConstraints();    // restitution override + wake-up flags
  int solve(double dt);             // GSSolver.solve over frictions ++ contacts ++ constraint rows
  void integrateAll(double dt);     // damping, integrate, clearForces, postStep springs, sleepTick
  void applySprings();              // Spring.applyForce for every spring, spring.dart:108-157
  void internalStep(double dt);
};

// shape constructors
void make_box_hull(const V3& he, Hull& h);
void make_cylinder_hull(double rTop, double rBottom, double height, int nSeg, Hull& h);
void shape_world_aabb(const Shape& s, const V3& pos, const Q4& q, V3& mn, V3& mx);

}  // namespace orc
