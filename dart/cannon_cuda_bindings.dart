// cannon_cuda_bindings.dart — dart:ffi binding of libcannon_cuda.so: EVERY struct and symbol of include/cannon_cuda.h.
//
// NOT compiled in the build container (no Dart SDK there; SURVEY.md 8b caveat). It is kept mechanical: one Struct per
// C struct with the fields in declaration order, one lookupFunction per symbol with the C signature next to it. The same
// table exists as ctypes prototypes in cannon_physics_b200/_ffi.py, where tests/test_host_logic.py checks it against the
// header symbol by symbol and against both libraries' export tables; a maintainer who changes the header changes the
// three places together. Drop the file into lib/cuda/ of cannon_physics next to cuda_world.dart.
// ignore_for_file: non_constant_identifier_names, camel_case_types
import 'dart:ffi';
import 'package:ffi/ffi.dart';

// ---- enums of the header -------------------------------------------------------------------------------------------
const cannonOk = 0, cannonEInvalid = -1, cannonECuda = -2, cannonECapacity = -3, cannonEUnsupported = -4, cannonENoGpu = -5;
const shapeSphere = 0, shapePlane = 1, shapeBox = 2, shapeConvex = 3, shapeCylinder = 4, shapeCapsule = 5, shapeCone = 6, shapeSizedPlane = 7,
    shapeHeightfield = 8, shapeParticle = 9, shapeTrimesh = 10; // ShapeType.index
const bpNaive = 0, bpSap = 1, bpGrid = 2;
const solverReferenceOrder = 0, solverColored = 1, solverSplit = 2, solverColoredF32 = 3;
const constraintPointToPoint = 0, constraintHinge = 1, constraintDistance = 2, constraintLock = 3, constraintConeTwist = 4;
const cannonBatchMaxGpus = 16;
const rayClosest = 1, rayAny = 2, rayAll = 4; // RayMode

// ---- structs ---------------------------------------------------------------------------------------------------------
final class CannonContactMaterial extends Struct {
  @Int32() external int materialA;
  @Int32() external int materialB;
  @Double() external double friction;
  @Double() external double restitution;
  @Double() external double contactEquationStiffness;
  @Double() external double contactEquationRelaxation;
  @Double() external double frictionEquationStiffness;
  @Double() external double frictionEquationRelaxation;
}

final class CannonWorldDesc extends Struct {
  @Array(3) external Array<Float> gravity;
  @Array(3) external Array<Float> frictionGravity;
  @Int32() external int hasFrictionGravity;
  @Int32() external int allowSleep;
  @Int32() external int quatNormalizeSkip;
  @Int32() external int quatNormalizeFast;
  @Int32() external int solverKind;
  @Int32() external int solverIterations;
  @Double() external double solverTolerance;
  @Int32() external int broadphaseKind;
  @Int32() external int useBoundingBoxes;
  @Int32() external int sapAxis;
  @Int32() external int gridNx;
  @Int32() external int gridNy;
  @Int32() external int gridNz;
  @Array(3) external Array<Float> gridMin;
  @Array(3) external Array<Float> gridMax;
  external CannonContactMaterial defaultContactMaterial;
  @Int32() external int nWorlds;
  @Int32() external int maxPairs;
  @Int32() external int maxContacts;
}

final class CannonShapeDesc extends Struct {
  @Int32() external int type;
  @Int32() external int collisionResponse;
  @Int32() external int collisionFilterGroup;
  @Int32() external int collisionFilterMask;
  @Double() external double radius;
  @Array(3) external Array<Float> halfExtents;
  @Double() external double radiusTop;
  @Double() external double radiusBottom;
  @Double() external double height;
  @Int32() external int numSegments;
  @Int32() external int nVertices;
  external Pointer<Float> vertices;
  @Int32() external int nFaces;
  external Pointer<Int32> faceOffsets;
  external Pointer<Int32> faceIndices;
  @Int32() external int hfNx;
  @Int32() external int hfNy;
  external Pointer<Double> hfData;
  @Int32() external int hfElementSize;
  @Int32() external int convexHasAxes;
  @Int32() external int nTriangles;
  external Pointer<Int32> tmIndices;
  @Array(3) external Array<Float> tmScale;
  @Int32() external int material;
}

final class CannonBodiesSoa extends Struct {
  @Int32() external int n;
  external Pointer<Float> position;
  external Pointer<Float> quaternion;
  external Pointer<Float> velocity;
  external Pointer<Float> angularVelocity;
  external Pointer<Float> force;
  external Pointer<Float> torque;
  external Pointer<Double> mass;
  external Pointer<Int32> type;
  external Pointer<Int32> sleepState;
  external Pointer<Double> timeLastSleepy;
  external Pointer<Uint8> allowSleep;
  external Pointer<Double> sleepSpeedLimit;
  external Pointer<Double> sleepTimeLimit;
  external Pointer<Double> linearDamping;
  external Pointer<Double> angularDamping;
  external Pointer<Float> linearFactor;
  external Pointer<Float> angularFactor;
  external Pointer<Uint8> fixedRotation;
  external Pointer<Int32> collisionFilterGroup;
  external Pointer<Int32> collisionFilterMask;
  external Pointer<Uint8> collisionResponse;
  external Pointer<Uint8> isTrigger;
  external Pointer<Int32> material;
  external Pointer<Int32> shape;
  external Pointer<Int32> worldId;
  external Pointer<Double> invMass;
  external Pointer<Float> invInertia;
  external Pointer<Float> invInertiaWorld;
  external Pointer<Double> boundingRadius;
  external Pointer<Float> aabb;
}

final class CannonConstraintDesc extends Struct {
  @Int32() external int type;
  @Int32() external int bodyA;
  @Int32() external int bodyB;
  @Array(3) external Array<Float> pivotA;
  @Array(3) external Array<Float> pivotB;
  @Array(3) external Array<Float> axisA;
  @Array(3) external Array<Float> axisB;
  @Double() external double maxForce;
  @Int32() external int collideConnected;
  @Int32() external int motorEnabled;
  @Double() external double motorTargetVelocity;
  @Double() external double motorMaxForce;
  @Double() external double distance;
  @Double() external double angle;
  @Double() external double twistAngle;
  @Int32() external int hasCtorPose;
  @Array(3) external Array<Float> ctorPosA;
  @Array(4) external Array<Float> ctorQuatA;
  @Array(3) external Array<Float> ctorPosB;
  @Array(4) external Array<Float> ctorQuatB;
}

final class CannonSpringDesc extends Struct {
  @Int32() external int bodyA;
  @Int32() external int bodyB;
  @Double() external double restLength;
  @Double() external double stiffness;
  @Double() external double damping;
  @Array(3) external Array<Float> localAnchorA;
  @Array(3) external Array<Float> localAnchorB;
}

final class CannonSphDesc extends Struct {
  @Int32() external int nParticles;
  external Pointer<Int32> particles;
  @Double() external double density;
  @Double() external double smoothingRadius;
  @Double() external double speedOfSound;
  @Double() external double viscosity;
  @Double() external double eps;
}

final class CannonContactsSoa extends Struct {
  @Int32() external int capacity;
  external Pointer<Int32> bodyI;
  external Pointer<Int32> bodyJ;
  external Pointer<Float> ri;
  external Pointer<Float> rj;
  external Pointer<Float> ni;
  external Pointer<Double> restitution;
  external Pointer<Double> friction;
  external Pointer<Uint8> enabled;
  external Pointer<Double> multiplier;
}

final class CannonProfile extends Struct {
  @Double() external double solve;
  @Double() external double makeContactConstraints;
  @Double() external double broadphase;
  @Double() external double integrate;
  @Double() external double narrowphase;
  @Int64() external int nPairs;
  @Int64() external int nContacts;
  @Int64() external int nRows;
  @Int64() external int nLevels;
  @Int64() external int iterationsDone;
  @Int64() external int steps;
  @Int64() external int contactItersTotal;
  @Double() external double stepCallMs;
  @Double() external double scheduleMs;
  @Double() external double gsMs;
  @Int64() external int kernelLaunches;
  @Int64() external int nTasks;
  @Int64() external int nIslands;
  @Array(8) external Array<Int64> nTasksByType;
  @Int64() external int sumSteps;
  @Double() external double sumStepMs;
  @Double() external double sumBroadphase;
  @Double() external double sumNarrowphase;
  @Double() external double sumSolve;
  @Double() external double sumIntegrate;
  @Double() external double sumSchedule;
  @Double() external double sumGs;
}

final class CannonRayOptions extends Struct {
  @Int32() external int mode;
  @Int32() external int skipBackfaces;
  @Int32() external int collisionFilterMask;
  @Int32() external int collisionFilterGroup;
  @Int32() external int checkCollisionResponse;
}

final class CannonRayHitsSoa extends Struct {
  @Int32() external int capacity;
  external Pointer<Int32> ray;
  external Pointer<Int32> body;
  external Pointer<Int32> hitFaceIndex;
  external Pointer<Double> distance;
  external Pointer<Float> hitPointWorld;
  external Pointer<Float> hitNormalWorld;
  external Pointer<Int32> shapeOrdinal;
}

final class CannonBatchStatistics extends Struct {
  @Int32() external int nGpus;
  @Int32() external int nWorlds;
  @Int32() external int bodiesPerWorld;
  @Int32() external int pad0;
  @Int64() external int nPairs;
  @Int64() external int nContacts;
  @Int64() external int nRows;
  @Int64() external int iterationsDone;
  @Int64() external int steps;
  @Int64() external int contactItersTotal;
  @Double() external double stepCallMsMax;
  @Array(16) external Array<Double> gpuStepCallMs;
  @Array(16) external Array<Int32> gpuWorlds;
}

typedef H = Pointer<Void>; // opaque cannon_ctx* / cannon_world* / cannon_batch*

// ---- symbols ---------------------------------------------------------------------------------------------------------
class CannonCuda {
  final DynamicLibrary lib;
  CannonCuda([String path = 'libcannon_cuda.so']) : lib = DynamicLibrary.open(path);

  // lifecycle
  late final int Function() version = lib.lookupFunction<Int32 Function(), int Function()>('cannon_version');
  late final Pointer<Utf8> Function() backend = lib.lookupFunction<Pointer<Utf8> Function(), Pointer<Utf8> Function()>('cannon_backend');
  late final int Function(int, Pointer<H>) ctxCreate =
      lib.lookupFunction<Int32 Function(Int32, Pointer<H>), int Function(int, Pointer<H>)>('cannon_ctx_create');
  late final void Function(H) ctxDestroy = lib.lookupFunction<Void Function(H), void Function(H)>('cannon_ctx_destroy');
  late final Pointer<Utf8> Function(H) lastError = lib.lookupFunction<Pointer<Utf8> Function(H), Pointer<Utf8> Function(H)>('cannon_last_error');
  late final void Function(Pointer<CannonWorldDesc>) worldDescDefault =
      lib.lookupFunction<Void Function(Pointer<CannonWorldDesc>), void Function(Pointer<CannonWorldDesc>)>('cannon_world_desc_default');
  late final void Function(Pointer<CannonShapeDesc>) shapeDescDefault =
      lib.lookupFunction<Void Function(Pointer<CannonShapeDesc>), void Function(Pointer<CannonShapeDesc>)>('cannon_shape_desc_default');

  // world
  late final int Function(H, Pointer<CannonWorldDesc>, Pointer<H>) worldCreate = lib.lookupFunction<
      Int32 Function(H, Pointer<CannonWorldDesc>, Pointer<H>), int Function(H, Pointer<CannonWorldDesc>, Pointer<H>)>('cannon_world_create');
  late final void Function(H) worldDestroy = lib.lookupFunction<Void Function(H), void Function(H)>('cannon_world_destroy');
  late final int Function(H, int, Pointer<Double>, Pointer<Double>, int, Pointer<CannonContactMaterial>) worldSetMaterials = lib.lookupFunction<
      Int32 Function(H, Int32, Pointer<Double>, Pointer<Double>, Int32, Pointer<CannonContactMaterial>),
      int Function(H, int, Pointer<Double>, Pointer<Double>, int, Pointer<CannonContactMaterial>)>('cannon_world_set_materials');
  late final int Function(H, int, Pointer<CannonShapeDesc>) worldSetShapes =
      lib.lookupFunction<Int32 Function(H, Int32, Pointer<CannonShapeDesc>), int Function(H, int, Pointer<CannonShapeDesc>)>('cannon_world_set_shapes');
  late final int Function(H, int, Pointer<Int32>, Pointer<Int32>, Pointer<Float>, Pointer<Float>) worldSetBodyShapes = lib.lookupFunction<
      Int32 Function(H, Int32, Pointer<Int32>, Pointer<Int32>, Pointer<Float>, Pointer<Float>),
      int Function(H, int, Pointer<Int32>, Pointer<Int32>, Pointer<Float>, Pointer<Float>)>('cannon_world_set_body_shapes');
  late final int Function(H, Pointer<CannonBodiesSoa>) worldSetBodies =
      lib.lookupFunction<Int32 Function(H, Pointer<CannonBodiesSoa>), int Function(H, Pointer<CannonBodiesSoa>)>('cannon_world_set_bodies');
  late final int Function(H, Pointer<CannonBodiesSoa>) worldGetBodies =
      lib.lookupFunction<Int32 Function(H, Pointer<CannonBodiesSoa>), int Function(H, Pointer<CannonBodiesSoa>)>('cannon_world_get_bodies');
  late final int Function(H, int, Pointer<CannonConstraintDesc>) worldSetConstraints = lib.lookupFunction<
      Int32 Function(H, Int32, Pointer<CannonConstraintDesc>), int Function(H, int, Pointer<CannonConstraintDesc>)>('cannon_world_set_constraints');
  late final int Function(H, int, Pointer<CannonSpringDesc>) worldSetSprings =
      lib.lookupFunction<Int32 Function(H, Int32, Pointer<CannonSpringDesc>), int Function(H, int, Pointer<CannonSpringDesc>)>('cannon_world_set_springs');
  late final void Function(Pointer<CannonSphDesc>) sphDescDefault =
      lib.lookupFunction<Void Function(Pointer<CannonSphDesc>), void Function(Pointer<CannonSphDesc>)>('cannon_sph_desc_default');
  late final int Function(H, int, Pointer<CannonSphDesc>) worldSetSphSystems =
      lib.lookupFunction<Int32 Function(H, Int32, Pointer<CannonSphDesc>), int Function(H, int, Pointer<CannonSphDesc>)>('cannon_world_set_sph_systems');
  late final int Function(H, double) worldSetTime = lib.lookupFunction<Int32 Function(H, Double), int Function(H, double)>('cannon_world_set_time');
  late final int Function(H, Pointer<Double>, Pointer<Int64>) worldGetTime =
      lib.lookupFunction<Int32 Function(H, Pointer<Double>, Pointer<Int64>), int Function(H, Pointer<Double>, Pointer<Int64>)>('cannon_world_get_time');
  late final int Function(H, int) worldSetStepnumber = lib.lookupFunction<Int32 Function(H, Int64), int Function(H, int)>('cannon_world_set_stepnumber');
  late final int Function(H, double) worldSetDt = lib.lookupFunction<Int32 Function(H, Double), int Function(H, double)>('cannon_world_set_dt');

  // staged entry points
  late final int Function(H) applyGravity = lib.lookupFunction<Int32 Function(H), int Function(H)>('cannon_apply_gravity');
  late final int Function(H, Pointer<Int32>, Pointer<Int32>, int, Pointer<Int32>) broadphasePairs = lib.lookupFunction<
      Int32 Function(H, Pointer<Int32>, Pointer<Int32>, Int32, Pointer<Int32>),
      int Function(H, Pointer<Int32>, Pointer<Int32>, int, Pointer<Int32>)>('cannon_broadphase_pairs');
  late final int Function(H, Pointer<Int32>, Pointer<Int32>, int, Pointer<CannonContactsSoa>, Pointer<Int32>, Pointer<Int32>) narrowphaseContacts =
      lib.lookupFunction<Int32 Function(H, Pointer<Int32>, Pointer<Int32>, Int32, Pointer<CannonContactsSoa>, Pointer<Int32>, Pointer<Int32>),
          int Function(H, Pointer<Int32>, Pointer<Int32>, int, Pointer<CannonContactsSoa>, Pointer<Int32>, Pointer<Int32>)>('cannon_narrowphase_contacts');
  late final int Function(H, double, Pointer<Int32>) solverSolve =
      lib.lookupFunction<Int32 Function(H, Double, Pointer<Int32>), int Function(H, double, Pointer<Int32>)>('cannon_solver_solve');
  late final int Function(H, double) integrate = lib.lookupFunction<Int32 Function(H, Double), int Function(H, double)>('cannon_integrate');

  // fused
  late final int Function(H, double, int) worldStep = lib.lookupFunction<Int32 Function(H, Double, Int32), int Function(H, double, int)>('cannon_world_step');
  late final int Function(H, double, int) worldStepProfiled =
      lib.lookupFunction<Int32 Function(H, Double, Int32), int Function(H, double, int)>('cannon_world_step_profiled');
  late final int Function(H, double, int) worldStepAsync =
      lib.lookupFunction<Int32 Function(H, Double, Int32), int Function(H, double, int)>('cannon_world_step_async');
  late final int Function(H) ctxSync = lib.lookupFunction<Int32 Function(H), int Function(H)>('cannon_ctx_sync');
  late final int Function(H, Pointer<CannonProfile>) worldProfile =
      lib.lookupFunction<Int32 Function(H, Pointer<CannonProfile>), int Function(H, Pointer<CannonProfile>)>('cannon_world_profile');
  late final int Function(H, Pointer<CannonContactsSoa>, Pointer<Int32>) worldGetContacts = lib.lookupFunction<
      Int32 Function(H, Pointer<CannonContactsSoa>, Pointer<Int32>), int Function(H, Pointer<CannonContactsSoa>, Pointer<Int32>)>('cannon_world_get_contacts');
  late final int Function(H, int) enableContactEvents =
      lib.lookupFunction<Int32 Function(H, Int32), int Function(H, int)>('cannon_world_enable_contact_events');
  late final int Function(H, int, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>) getContactEvents =
      lib.lookupFunction<Int32 Function(H, Int32, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>),
          int Function(H, int, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>)>(
          'cannon_world_get_contact_events');
  late final int Function(H, int, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>, Pointer<Double>, Pointer<Double>, Pointer<Double>, Pointer<Int32>) worldGetRows =
      lib.lookupFunction<
          Int32 Function(H, Int32, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>, Pointer<Double>, Pointer<Double>, Pointer<Double>, Pointer<Int32>),
          int Function(H, int, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>, Pointer<Double>, Pointer<Double>, Pointer<Double>, Pointer<Int32>)>(
          'cannon_world_get_rows');

  // user mutations between steps
  late final int Function(H, int, int, Pointer<Float>, Pointer<Float>, Pointer<Float>, Pointer<Float>, Pointer<Float>, Pointer<Float>) worldUpdateBodies =
      lib.lookupFunction<Int32 Function(H, Int32, Int32, Pointer<Float>, Pointer<Float>, Pointer<Float>, Pointer<Float>, Pointer<Float>, Pointer<Float>),
          int Function(H, int, int, Pointer<Float>, Pointer<Float>, Pointer<Float>, Pointer<Float>, Pointer<Float>, Pointer<Float>)>(
          'cannon_world_update_bodies');
  late final int Function(H, int, int, Pointer<Float>) worldSetInvInertia =
      lib.lookupFunction<Int32 Function(H, Int32, Int32, Pointer<Float>), int Function(H, int, int, Pointer<Float>)>('cannon_world_set_inv_inertia');
  late final int Function(H, int, int, Pointer<Int32>) worldUpdateSleepStates =
      lib.lookupFunction<Int32 Function(H, Int32, Int32, Pointer<Int32>), int Function(H, int, int, Pointer<Int32>)>('cannon_world_update_sleep_states');
  late final int Function(H, int, int, double, double) worldSetHingeMotor =
      lib.lookupFunction<Int32 Function(H, Int32, Int32, Double, Double), int Function(H, int, int, double, double)>('cannon_world_set_hinge_motor');

  // ray casts and AABB queries
  late final void Function(Pointer<CannonRayOptions>) rayOptionsDefault =
      lib.lookupFunction<Void Function(Pointer<CannonRayOptions>), void Function(Pointer<CannonRayOptions>)>('cannon_ray_options_default');
  late final int Function(H, int, Pointer<Float>, Pointer<Float>, Pointer<CannonRayOptions>, Pointer<Uint8>, Pointer<CannonRayHitsSoa>, Pointer<Int32>) worldRaycast =
      lib.lookupFunction<
          Int32 Function(H, Int32, Pointer<Float>, Pointer<Float>, Pointer<CannonRayOptions>, Pointer<Uint8>, Pointer<CannonRayHitsSoa>, Pointer<Int32>),
          int Function(H, int, Pointer<Float>, Pointer<Float>, Pointer<CannonRayOptions>, Pointer<Uint8>, Pointer<CannonRayHitsSoa>, Pointer<Int32>)>(
          'cannon_world_raycast');
  late final int Function(H, Pointer<Float>, Pointer<Float>, Pointer<Int32>, int, Pointer<Int32>) worldAabbQuery = lib.lookupFunction<
      Int32 Function(H, Pointer<Float>, Pointer<Float>, Pointer<Int32>, Int32, Pointer<Int32>),
      int Function(H, Pointer<Float>, Pointer<Float>, Pointer<Int32>, int, Pointer<Int32>)>('cannon_world_aabb_query');

  // batches of independent worlds over the GPUs of one box
  late final int Function(Pointer<Int32>, int, Pointer<CannonWorldDesc>, int, int, Pointer<H>) batchCreate = lib.lookupFunction<
      Int32 Function(Pointer<Int32>, Int32, Pointer<CannonWorldDesc>, Int32, Int32, Pointer<H>),
      int Function(Pointer<Int32>, int, Pointer<CannonWorldDesc>, int, int, Pointer<H>)>('cannon_batch_create');
  late final void Function(H) batchDestroy = lib.lookupFunction<Void Function(H), void Function(H)>('cannon_batch_destroy');
  late final Pointer<Utf8> Function(H) batchLastError =
      lib.lookupFunction<Pointer<Utf8> Function(H), Pointer<Utf8> Function(H)>('cannon_batch_last_error');
  late final int Function(H, int, Pointer<Double>, Pointer<Double>, int, Pointer<CannonContactMaterial>) batchSetMaterials = lib.lookupFunction<
      Int32 Function(H, Int32, Pointer<Double>, Pointer<Double>, Int32, Pointer<CannonContactMaterial>),
      int Function(H, int, Pointer<Double>, Pointer<Double>, int, Pointer<CannonContactMaterial>)>('cannon_batch_set_materials');
  late final int Function(H, int, Pointer<CannonShapeDesc>) batchSetShapes =
      lib.lookupFunction<Int32 Function(H, Int32, Pointer<CannonShapeDesc>), int Function(H, int, Pointer<CannonShapeDesc>)>('cannon_batch_set_shapes');
  late final int Function(H, Pointer<CannonBodiesSoa>) batchSetBodies =
      lib.lookupFunction<Int32 Function(H, Pointer<CannonBodiesSoa>), int Function(H, Pointer<CannonBodiesSoa>)>('cannon_batch_set_bodies');
  late final int Function(H, int, Pointer<Int32>, Pointer<Int32>, Pointer<Float>, Pointer<Float>) batchSetBodyShapes = lib.lookupFunction<
      Int32 Function(H, Int32, Pointer<Int32>, Pointer<Int32>, Pointer<Float>, Pointer<Float>),
      int Function(H, int, Pointer<Int32>, Pointer<Int32>, Pointer<Float>, Pointer<Float>)>('cannon_batch_set_body_shapes');
  late final int Function(H, int, Pointer<CannonConstraintDesc>) batchSetConstraints = lib.lookupFunction<
      Int32 Function(H, Int32, Pointer<CannonConstraintDesc>), int Function(H, int, Pointer<CannonConstraintDesc>)>('cannon_batch_set_constraints');
  late final int Function(H, double, int) batchStep = lib.lookupFunction<Int32 Function(H, Double, Int32), int Function(H, double, int)>('cannon_batch_step');
  late final int Function(H, Pointer<CannonBatchStatistics>) batchStats =
      lib.lookupFunction<Int32 Function(H, Pointer<CannonBatchStatistics>), int Function(H, Pointer<CannonBatchStatistics>)>('cannon_batch_stats');
  late final int Function(H, Pointer<CannonBodiesSoa>) batchGetBodies =
      lib.lookupFunction<Int32 Function(H, Pointer<CannonBodiesSoa>), int Function(H, Pointer<CannonBodiesSoa>)>('cannon_batch_get_bodies');
  late final int Function(H, int, Pointer<Int32>, Pointer<Int32>, Pointer<H>) batchShard = lib.lookupFunction<
      Int32 Function(H, Int32, Pointer<Int32>, Pointer<Int32>, Pointer<H>), int Function(H, int, Pointer<Int32>, Pointer<Int32>, Pointer<H>)>('cannon_batch_shard');
}
