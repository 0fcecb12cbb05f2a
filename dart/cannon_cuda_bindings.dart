// cannon_cuda_bindings.dart — dart:ffi binding of libcannon_cuda.so (include/cannon_cuda.h).
//
// NOT compiled in the build container (no Dart SDK there); kept thin and mechanical so a maintainer of
// cannon_physics can drop it into lib/cuda/ next to the reference's Broadphase / Narrowphase / Solver classes.
// Every native symbol below is exported by the library and exercised through the identical ctypes binding in
// cannon_physics_b200/_ffi.py.
import 'dart:ffi';
import 'package:ffi/ffi.dart';

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
  @Int32() external int solverKind;        // 0 reference order (bit-reproducible), 1 colored (throughput)
  @Int32() external int solverIterations;
  @Double() external double solverTolerance;
  @Int32() external int broadphaseKind;    // 0 Naive, 1 SAP, 2 Grid
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

final class CannonBodiesSoa extends Struct {
  @Int32() external int n;
  external Pointer<Float> position, quaternion, velocity, angularVelocity, force, torque;
  external Pointer<Double> mass;
  external Pointer<Int32> type, sleepState;
  external Pointer<Double> timeLastSleepy;
  external Pointer<Uint8> allowSleep;
  external Pointer<Double> sleepSpeedLimit, sleepTimeLimit, linearDamping, angularDamping;
  external Pointer<Float> linearFactor, angularFactor;
  external Pointer<Uint8> fixedRotation;
  external Pointer<Int32> collisionFilterGroup, collisionFilterMask;
  external Pointer<Uint8> collisionResponse, isTrigger;
  external Pointer<Int32> material, shape, worldId;
  external Pointer<Double> invMass;
  external Pointer<Float> invInertia, invInertiaWorld;
  external Pointer<Double> boundingRadius;
  external Pointer<Float> aabb;
}

final class CannonContactsSoa extends Struct {
  @Int32() external int capacity;
  external Pointer<Int32> bodyI, bodyJ;
  external Pointer<Float> ri, rj, ni;
  external Pointer<Double> restitution, friction;
  external Pointer<Uint8> enabled;
  external Pointer<Double> multiplier;
}

typedef _CtxCreateC = Int32 Function(Int32, Pointer<Pointer<Void>>);
typedef _CtxCreateD = int Function(int, Pointer<Pointer<Void>>);
typedef _WorldCreateC = Int32 Function(Pointer<Void>, Pointer<CannonWorldDesc>, Pointer<Pointer<Void>>);
typedef _WorldCreateD = int Function(Pointer<Void>, Pointer<CannonWorldDesc>, Pointer<Pointer<Void>>);
typedef _SetBodiesC = Int32 Function(Pointer<Void>, Pointer<CannonBodiesSoa>);
typedef _SetBodiesD = int Function(Pointer<Void>, Pointer<CannonBodiesSoa>);
typedef _StepC = Int32 Function(Pointer<Void>, Double, Int32);
typedef _StepD = int Function(Pointer<Void>, double, int);
typedef _PairsC = Int32 Function(Pointer<Void>, Pointer<Int32>, Pointer<Int32>, Int32, Pointer<Int32>);
typedef _PairsD = int Function(Pointer<Void>, Pointer<Int32>, Pointer<Int32>, int, Pointer<Int32>);
typedef _ContactsC = Int32 Function(Pointer<Void>, Pointer<Int32>, Pointer<Int32>, Int32, Pointer<CannonContactsSoa>, Pointer<Int32>, Pointer<Int32>);
typedef _ContactsD = int Function(Pointer<Void>, Pointer<Int32>, Pointer<Int32>, int, Pointer<CannonContactsSoa>, Pointer<Int32>, Pointer<Int32>);
typedef _EventsEnableC = Int32 Function(Pointer<Void>, Int32);
typedef _EventsEnableD = int Function(Pointer<Void>, int);
typedef _EventsGetC = Int32 Function(Pointer<Void>, Int32, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>);
typedef _EventsGetD = int Function(Pointer<Void>, int, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>, Pointer<Int32>);
typedef _SolveC = Int32 Function(Pointer<Void>, Double, Pointer<Int32>);
typedef _SolveD = int Function(Pointer<Void>, double, Pointer<Int32>);

class CannonCuda {
  final DynamicLibrary lib;
  late final _CtxCreateD ctxCreate = lib.lookupFunction<_CtxCreateC, _CtxCreateD>('cannon_ctx_create');
  late final void Function(Pointer<CannonWorldDesc>) worldDescDefault =
      lib.lookupFunction<Void Function(Pointer<CannonWorldDesc>), void Function(Pointer<CannonWorldDesc>)>('cannon_world_desc_default');
  late final _WorldCreateD worldCreate = lib.lookupFunction<_WorldCreateC, _WorldCreateD>('cannon_world_create');
  late final _SetBodiesD worldSetBodies = lib.lookupFunction<_SetBodiesC, _SetBodiesD>('cannon_world_set_bodies');
  late final _SetBodiesD worldGetBodies = lib.lookupFunction<_SetBodiesC, _SetBodiesD>('cannon_world_get_bodies');
  late final _StepD worldStep = lib.lookupFunction<_StepC, _StepD>('cannon_world_step');
  late final _PairsD broadphasePairs = lib.lookupFunction<_PairsC, _PairsD>('cannon_broadphase_pairs');
  late final _ContactsD narrowphaseContacts = lib.lookupFunction<_ContactsC, _ContactsD>('cannon_narrowphase_contacts');
  late final _SolveD solverSolve = lib.lookupFunction<_SolveC, _SolveD>('cannon_solver_solve');
  late final _EventsEnableD enableContactEvents = lib.lookupFunction<_EventsEnableC, _EventsEnableD>('cannon_world_enable_contact_events');
  late final _EventsGetD getContactEvents = lib.lookupFunction<_EventsGetC, _EventsGetD>('cannon_world_get_contact_events');
  late final Pointer<Utf8> Function(Pointer<Void>) lastError =
      lib.lookupFunction<Pointer<Utf8> Function(Pointer<Void>), Pointer<Utf8> Function(Pointer<Void>)>('cannon_last_error');
  CannonCuda([String path = 'libcannon_cuda.so']) : lib = DynamicLibrary.open(path);
}
