// cuda_world.dart — the drop-in classes of the north_star for the reference's pluggable step stages.
//
//   CudaSession      one cannon_ctx + cannon_world for a reference World: flattens World / Body / Shape / Material /
//                    ContactMaterial / Constraint objects into the SoA structs of include/cannon_cuda.h (the same mapping
//                    as cannon_physics_b200/api.py World._spec, which the CPU tests exercise), uploads, gathers and
//                    scatters body state
//   CudaWorld        extends World, overrides internalStep(dt) (lib/world/world_class.dart:433-701): state device-resident,
//                    ONE FFI call per step, poses downloaded after the step (or lazily with syncEveryStep = false)
//   CudaBroadphase   extends Broadphase: collisionPairs(world, p1, p2) (lib/collision/broadphase.dart:39)
//   CudaNarrowphase  extends Narrowphase: getContacts(...) (lib/world/narrow_phase.dart:634-642); contact geometry from the
//                    device, equation objects / materials / friction equations through the reference's own
//                    createContactEquation / createFrictionEquationsFromContact (:492-586)
//   CudaGSSolver     extends Solver: solve(dt, world) (lib/solver/solver.dart:26-46, gs_solver.dart:27-133) over the contacts
//                    of the last CudaNarrowphase call plus the world's constraints; writes velocities and multipliers back
//
// Drop-in (staged) mode re-gathers the body state before every stage, so user code that edits bodies between stages stays
// coherent; it is for API compatibility and parity tests (O(N) Dart marshalling per stage). Throughput mode is CudaWorld.
// Errors are thrown as strings, like the reference does (e.g. lib/collision/broadphase.dart:40).
//
// NOT compiled in the build container (no Dart SDK there). Every native call below is exercised through the identical
// ctypes binding (cannon_physics_b200/_ffi.py, engine.py, api.py) by tests/.
import 'dart:ffi';
import 'package:ffi/ffi.dart';
import 'package:vector_math/vector_math.dart';
import 'package:cannon_physics/cannon_physics.dart';
import 'cannon_cuda_bindings.dart';

int _shapeTypeCode(Shape s) {
  switch (s.type) {
    case ShapeType.sphere: return shapeSphere;
    case ShapeType.plane: return shapePlane;
    case ShapeType.box: return shapeBox;
    case ShapeType.convex: return shapeConvex;
    case ShapeType.cylinder: return shapeCylinder;
    case ShapeType.capsule: return shapeCapsule;        // Capsule / CapsuleLathe: the hull their constructor built is passed as is
    case ShapeType.cone: return shapeCone;
    case ShapeType.sizedPlane: return shapeSizedPlane;
    case ShapeType.heightfield: return shapeHeightfield;
    case ShapeType.particle: return shapeParticle;
    case ShapeType.trimesh: return shapeTrimesh;
    default: throw 'CudaSession: shape type ${s.type} is outside the hot-path scope (SURVEY.md 8f)';
  }
}

void _put3(Array<Float> a, Vector3 v) { a[0] = v.x; a[1] = v.y; a[2] = v.z; }

class CudaSession {
  final CannonCuda cuda;
  final World world;
  final int device;
  H ctx = nullptr, handle = nullptr;
  final Arena _arena = Arena(); // staging buffers that live as long as the session
  Pointer<CannonBodiesSoa> _soa = nullptr;
  int _n = 0, _nConstraints = -1;
  final Map<Shape, int> _shapeIndex = {};
  final Map<Material, int> _materialIndex = {};

  CudaSession(this.cuda, this.world, {this.device = 0, int solverKind = solverReferenceOrder});

  void check(int rc, String what) {
    if (rc != cannonOk) throw '$what failed ($rc): ${ctx == nullptr ? '' : cuda.lastError(ctx).toDartString()}';
  }

  void dispose() {
    if (handle != nullptr) cuda.worldDestroy(handle);
    if (ctx != nullptr) cuda.ctxDestroy(ctx);
    handle = nullptr; ctx = nullptr;
    _arena.releaseAll();
  }

  // ---- upload: World -> cannon_world (rebuilds when the body / constraint lists changed) -----------------------------
  void ensureUploaded({int solverKind = solverReferenceOrder}) {
    if (handle != nullptr && _n == world.bodies.length && _nConstraints == world.constraints.length) return;
    if (handle != nullptr) { cuda.worldDestroy(handle); handle = nullptr; }
    if (ctx == nullptr) {
      final pc = calloc<H>();
      try { check(cuda.ctxCreate(device, pc), 'cannon_ctx_create'); ctx = pc.value; } finally { calloc.free(pc); }
    }
    using((Arena a) {
      // World constructor parameters + pluggable objects flattened to POD (world_class.dart:135-162)
      final d = a<CannonWorldDesc>();
      cuda.worldDescDefault(d);
      _put3(d.ref.gravity, world.gravity);
      if (world.frictionGravity != null) { _put3(d.ref.frictionGravity, world.frictionGravity!); d.ref.hasFrictionGravity = 1; }
      d.ref.allowSleep = world.allowSleep ? 1 : 0;
      d.ref.quatNormalizeSkip = world.quatNormalizeSkip;
      d.ref.quatNormalizeFast = world.quatNormalizeFast ? 1 : 0;
      d.ref.solverKind = world.solver is SplitSolver ? solverSplit : solverKind;
      d.ref.solverIterations = world.solver is SplitSolver ? (world.solver as SplitSolver).subsolver.iterations : world.solver.iterations;
      d.ref.solverTolerance = world.solver is SplitSolver ? (world.solver as SplitSolver).subsolver.tolerance : world.solver.tolerance;
      final bp = world.broadphase is CudaBroadphase ? (world.broadphase as CudaBroadphase).model : world.broadphase;
      d.ref.useBoundingBoxes = bp.useBoundingBoxes ? 1 : 0;
      if (bp is SAPBroadphase) { d.ref.broadphaseKind = bpSap; d.ref.sapAxis = bp.axisIndex.index; }
      else if (bp is GridBroadphase) {
        d.ref.broadphaseKind = bpGrid; d.ref.gridNx = bp.nx; d.ref.gridNy = bp.ny; d.ref.gridNz = bp.nz;
        _put3(d.ref.gridMin, bp.aabbMin); _put3(d.ref.gridMax, bp.aabbMax);
      } else { d.ref.broadphaseKind = bpNaive; }
      _putCm(d.ref.defaultContactMaterial, world.defaultContactMaterial, 0, 0);
      final ph = a<H>();
      check(cuda.worldCreate(ctx, d, ph), 'cannon_world_create');
      handle = ph.value;

      // materials and contact materials (lib/material/*.dart; World.addContactMaterial, world_class.dart:343-348)
      _materialIndex.clear();
      int mat(Material? m) => m == null ? -1 : _materialIndex.putIfAbsent(m, () => _materialIndex.length);
      for (final cm in world.contactmaterials) { mat(cm.materials[0]); mat(cm.materials[1]); }
      for (final b in world.bodies) { mat(b.material); for (final sh in b.shapes) { mat(sh.material); } }
      final nm = _materialIndex.length;
      final fr = a<Double>(nm + 1), re = a<Double>(nm + 1);
      _materialIndex.forEach((m, i) { fr[i] = m.friction; re[i] = m.restitution; });
      final cms = a<CannonContactMaterial>(world.contactmaterials.length + 1);
      for (var i = 0; i < world.contactmaterials.length; i++) {
        final cm = world.contactmaterials[i];
        _putCm(cms[i], cm, mat(cm.materials[0]), mat(cm.materials[1]));
      }
      check(cuda.worldSetMaterials(handle, nm, fr, re, world.contactmaterials.length, cms), 'cannon_world_set_materials');

      // shape table: one entry per distinct Shape object (the demos share shapes, examples/lib/examples/container.dart:105)
      _shapeIndex.clear();
      final shapes = <Shape>[];
      for (final b in world.bodies) {
        for (final sh in b.shapes) {
          if (_shapeIndex.putIfAbsent(sh, () => shapes.length) == shapes.length) shapes.add(sh);
        }
      }
      final sd = a<CannonShapeDesc>(shapes.length + 1);
      for (var i = 0; i < shapes.length; i++) {
        final s = shapes[i];
        cuda.shapeDescDefault(sd + i);
        final r = sd[i];
        r.type = _shapeTypeCode(s);
        r.collisionResponse = s.collisionResponse ? 1 : 0;
        r.collisionFilterGroup = s.collisionFilterGroup;
        r.collisionFilterMask = s.collisionFilterMask;
        r.material = mat(s.material);  // Shape.material (shape.dart:48), -1 = null
        if (s is Sphere) { r.radius = s.radius; }
        else if (s is Box) { _put3(r.halfExtents, s.halfExtents); }
        else if (s is Cylinder) { r.radiusTop = s.radiusTop; r.radiusBottom = s.radiusBottom; r.height = s.height; r.numSegments = s.numSegments; }
        else if (s is Heightfield) {
          final nx = s.data.length, ny = s.data[0].length;
          final hd = a<Double>(nx * ny);
          for (var x = 0; x < nx; x++) { for (var y = 0; y < ny; y++) { hd[x * ny + y] = s.data[x][y]; } }
          r.hfNx = nx; r.hfNy = ny; r.hfData = hd; r.hfElementSize = s.elementSize.toInt();
        } else if (s is Trimesh) {
          final nv = s.vertices.length ~/ 3;
          final v = a<Float>(3 * nv + 3);
          for (var k = 0; k < 3 * nv; k++) { v[k] = s.vertices[k]; }  // getVertex rounds to float the same way (trimesh.dart:270-275)
          final idx = a<Int32>(s.indices.length + 1);
          for (var k = 0; k < s.indices.length; k++) { idx[k] = s.indices[k]; }
          r.nVertices = nv; r.vertices = v; r.nTriangles = s.indices.length ~/ 3; r.tmIndices = idx;
          _put3(r.tmScale, s.scale);
        } else if (s is ConvexPolyhedron) {
          final v = a<Float>(3 * s.vertices.length);
          for (var k = 0; k < s.vertices.length; k++) { v[3 * k] = s.vertices[k].x; v[3 * k + 1] = s.vertices[k].y; v[3 * k + 2] = s.vertices[k].z; }
          var total = 0;
          for (final f in s.faces) { total += f.length; }
          final off = a<Int32>(s.faces.length + 1), idx = a<Int32>(total + 1);
          var o = 0;
          for (var f = 0; f < s.faces.length; f++) { off[f] = o; for (final vi in s.faces[f]) { idx[o++] = vi; } }
          off[s.faces.length] = o;
          r.nVertices = s.vertices.length; r.vertices = v; r.nFaces = s.faces.length; r.faceOffsets = off; r.faceIndices = idx;
          r.convexHasAxes = s.uniqueAxes != null ? 1 : 0;  // findSeparatingAxis only asks whether axes were given
        }
      }
      check(cuda.worldSetShapes(handle, shapes.length, sd), 'cannon_world_set_shapes');

      // bodies (lib/objects/rigid_body.dart:27-86): the static attributes once, the dynamic state through _gather()
      _n = world.bodies.length;
      final n = _n + 1;
      _arena.releaseAll();
      _soa = _arena<CannonBodiesSoa>();
      final s = _soa.ref;
      s.n = _n;
      s.position = _arena<Float>(3 * n); s.quaternion = _arena<Float>(4 * n); s.velocity = _arena<Float>(3 * n);
      s.angularVelocity = _arena<Float>(3 * n); s.force = _arena<Float>(3 * n); s.torque = _arena<Float>(3 * n);
      s.mass = _arena<Double>(n); s.type = _arena<Int32>(n); s.sleepState = _arena<Int32>(n); s.timeLastSleepy = _arena<Double>(n);
      s.allowSleep = _arena<Uint8>(n); s.sleepSpeedLimit = _arena<Double>(n); s.sleepTimeLimit = _arena<Double>(n);
      s.linearDamping = _arena<Double>(n); s.angularDamping = _arena<Double>(n);
      s.linearFactor = _arena<Float>(3 * n); s.angularFactor = _arena<Float>(3 * n); s.fixedRotation = _arena<Uint8>(n);
      s.collisionFilterGroup = _arena<Int32>(n); s.collisionFilterMask = _arena<Int32>(n);
      s.collisionResponse = _arena<Uint8>(n); s.isTrigger = _arena<Uint8>(n); s.material = _arena<Int32>(n); s.shape = _arena<Int32>(n);
      s.worldId = nullptr; s.invMass = nullptr; s.invInertia = nullptr; s.invInertiaWorld = nullptr; s.boundingRadius = nullptr; s.aabb = nullptr;
      for (var i = 0; i < _n; i++) {
        final b = world.bodies[i];
        s.mass[i] = b.mass; s.type[i] = b.type.index; s.timeLastSleepy[i] = b.timeLastSleepy.toDouble();
        s.allowSleep[i] = b.allowSleep ? 1 : 0; s.sleepSpeedLimit[i] = b.sleepSpeedLimit; s.sleepTimeLimit[i] = b.sleepTimeLimit;
        s.linearDamping[i] = b.linearDamping; s.angularDamping[i] = b.angularDamping;
        for (var k = 0; k < 3; k++) { s.linearFactor[3 * i + k] = b.linearFactor[k]; s.angularFactor[3 * i + k] = b.angularFactor[k]; }
        s.fixedRotation[i] = b.fixedRotation ? 1 : 0;
        s.collisionFilterGroup[i] = b.collisionFilterGroup; s.collisionFilterMask[i] = b.collisionFilterMask;
        s.collisionResponse[i] = b.collisionResponse ? 1 : 0; s.isTrigger[i] = b.isTrigger ? 1 : 0;
        s.material[i] = mat(b.material);
        s.shape[i] = b.shapes.isEmpty ? -1 : _shapeIndex[b.shapes[0]]!;
      }
      _gather();
      // Body.shapes / shapeOffsets / shapeOrientations (rigid_body.dart:96-104) as the instance table of the next set_bodies
      var nInst = 0;
      for (final b in world.bodies) { nInst += b.shapes.length; }
      final first = a<Int32>(_n + 2), ish = a<Int32>(nInst + 1);
      final ioff = a<Float>(3 * nInst + 3), iori = a<Float>(4 * nInst + 4);
      var k = 0;
      for (var i = 0; i < _n; i++) {
        final b = world.bodies[i];
        first[i] = k;
        for (var j = 0; j < b.shapes.length; j++, k++) {
          ish[k] = _shapeIndex[b.shapes[j]]!;
          final o = b.shapeOffsets[j], q = b.shapeOrientations[j];
          ioff[3 * k] = o.x; ioff[3 * k + 1] = o.y; ioff[3 * k + 2] = o.z;
          iori[4 * k] = q.x; iori[4 * k + 1] = q.y; iori[4 * k + 2] = q.z; iori[4 * k + 3] = q.w;
        }
      }
      first[_n] = k;
      check(cuda.worldSetBodyShapes(handle, _n, first, ish, ioff, iori), 'cannon_world_set_body_shapes');
      check(cuda.worldSetBodies(handle, _soa), 'cannon_world_set_bodies');

      // constraints (lib/constraints/*.dart). LockConstraint / default-distance parameters are evaluated by the library on
      // the uploaded poses, like the reference constructors (include/cannon_cuda.h, cannon_constraint_desc)
      _nConstraints = world.constraints.length;
      final cd = a<CannonConstraintDesc>(_nConstraints + 1);
      for (var i = 0; i < _nConstraints; i++) {
        final c = world.constraints[i];
        final r = cd[i];
        r.bodyA = c.bodyA.index; r.bodyB = c.bodyB.index; r.collideConnected = c.collideConnected ? 1 : 0;
        r.maxForce = 1e6; r.distance = -1; r.axisA[0] = 1; r.axisB[0] = 1;
        if (c is HingeConstraint) {
          r.type = constraintHinge; _put3(r.pivotA, c.pivotA); _put3(r.pivotB, c.pivotB); _put3(r.axisA, c.axisA); _put3(r.axisB, c.axisB);
          r.maxForce = c.equationX.maxForce;
          r.motorEnabled = c.motorEquation.enabled ? 1 : 0; r.motorTargetVelocity = c.motorEquation.targetVelocity; r.motorMaxForce = c.motorEquation.maxForce;
        } else if (c is ConeTwistConstraint) {
          r.type = constraintConeTwist; _put3(r.pivotA, c.pivotA); _put3(r.pivotB, c.pivotB); _put3(r.axisA, c.axisA); _put3(r.axisB, c.axisB);
          r.maxForce = c.equationX.maxForce; r.angle = c.angle; r.twistAngle = c.twistAngle;
        } else if (c is LockConstraint) {
          r.type = constraintLock; r.maxForce = c.equationX.maxForce;
        } else if (c is PointToPointConstraint) {
          r.type = constraintPointToPoint; _put3(r.pivotA, c.pivotA); _put3(r.pivotB, c.pivotB); r.maxForce = c.equationX.maxForce;
        } else if (c is DistanceConstraint) {
          r.type = constraintDistance; r.distance = c.distance; r.maxForce = c.distanceEquation.maxForce;
        } else if (c is SpringConstraint) {  // a distance row with |force| <= stiffness (spring_constraint.dart:34-41)
          r.type = constraintDistance; r.distance = -1; r.maxForce = c.stiffness;
        } else {
          throw 'CudaSession: ${c.runtimeType} is outside the hot-path scope (SURVEY.md 8f)';
        }
      }
      check(cuda.worldSetConstraints(handle, _nConstraints, cd), 'cannon_world_set_constraints');
      // World.subsystems (world_class.dart:121,472-475): the SPHSystem instances, particles as body indices in add() order
      final sphs = world.subsystems.whereType<SPHSystem>().toList();
      final sphd = a<CannonSphDesc>(sphs.length + 1);
      for (var i = 0; i < sphs.length; i++) {
        final sp = sphs[i];
        cuda.sphDescDefault(sphd + i);
        final pl = a<Int32>(sp.particles.length + 1);
        for (var k = 0; k < sp.particles.length; k++) { pl[k] = sp.particles[k].index; }
        sphd[i].nParticles = sp.particles.length; sphd[i].particles = pl;
        sphd[i].density = sp.density; sphd[i].smoothingRadius = sp.smoothingRadius; sphd[i].speedOfSound = sp.speedOfSound;
        sphd[i].viscosity = sp.viscosity; sphd[i].eps = sp.eps;
      }
      check(cuda.worldSetSphSystems(handle, sphs.length, sphd), 'cannon_world_set_sph_systems');
      check(cuda.worldSetTime(handle, world.time), 'cannon_world_set_time');
      check(cuda.worldSetStepnumber(handle, world.stepnumber), 'cannon_world_set_stepnumber');
    });
  }

  void _putCm(CannonContactMaterial r, ContactMaterial cm, int a, int b) {
    r.materialA = a; r.materialB = b; r.friction = cm.friction; r.restitution = cm.restitution;
    r.contactEquationStiffness = cm.contactEquationStiffness; r.contactEquationRelaxation = cm.contactEquationRelaxation;
    r.frictionEquationStiffness = cm.frictionEquationStiffness; r.frictionEquationRelaxation = cm.frictionEquationRelaxation;
  }

  // ---- dynamic state: Body objects -> staging (gather), staging -> Body objects (scatter) ----------------------------
  void _gather() {
    final s = _soa.ref;
    for (var i = 0; i < _n; i++) {
      final b = world.bodies[i];
      for (var k = 0; k < 3; k++) {
        s.position[3 * i + k] = b.position[k]; s.velocity[3 * i + k] = b.velocity[k]; s.angularVelocity[3 * i + k] = b.angularVelocity[k];
        s.force[3 * i + k] = b.force[k]; s.torque[3 * i + k] = b.torque[k];
      }
      s.quaternion[4 * i] = b.quaternion.x; s.quaternion[4 * i + 1] = b.quaternion.y; s.quaternion[4 * i + 2] = b.quaternion.z; s.quaternion[4 * i + 3] = b.quaternion.w;
      s.sleepState[i] = b.sleepState.index;
    }
  }

  /// Host edits (body.position / velocity / force ..., applyForce, sleep, wakeUp) -> device, before a stage or a step.
  void pushState() {
    _gather();
    final s = _soa.ref;
    check(cuda.worldUpdateBodies(handle, 0, _n, s.position, s.quaternion, s.velocity, s.angularVelocity, s.force, s.torque), 'cannon_world_update_bodies');
    check(cuda.worldUpdateSleepStates(handle, 0, _n, s.sleepState), 'cannon_world_update_sleep_states');
  }

  /// Device state -> Body objects (poses, velocities, sleep state; forces are cleared by the step like clearForces does).
  void pullState({bool forces = false}) {
    check(cuda.worldGetBodies(handle, _soa), 'cannon_world_get_bodies');
    final s = _soa.ref;
    for (var i = 0; i < _n; i++) {
      final b = world.bodies[i];
      b.position.setValues(s.position[3 * i], s.position[3 * i + 1], s.position[3 * i + 2]);
      b.quaternion.setValues(s.quaternion[4 * i], s.quaternion[4 * i + 1], s.quaternion[4 * i + 2], s.quaternion[4 * i + 3]);
      b.velocity.setValues(s.velocity[3 * i], s.velocity[3 * i + 1], s.velocity[3 * i + 2]);
      b.angularVelocity.setValues(s.angularVelocity[3 * i], s.angularVelocity[3 * i + 1], s.angularVelocity[3 * i + 2]);
      b.sleepState = BodySleepStates.values[s.sleepState[i]];
      b.timeLastSleepy = s.timeLastSleepy[i];
      if (forces) {
        b.force.setValues(s.force[3 * i], s.force[3 * i + 1], s.force[3 * i + 2]);
        b.torque.setValues(s.torque[3 * i], s.torque[3 * i + 1], s.torque[3 * i + 2]);
      }
    }
  }
}

/// Throughput mode: the whole of World.internalStep in one native call, state resident on the device.
class CudaWorld extends World {
  final CannonCuda cuda;
  late final CudaSession session;
  final int solverKind;
  /// true: Body objects are refreshed after every step (drop-in behaviour). false: call [syncBodies] when poses are needed.
  bool syncEveryStep;
  bool hostDirty = true; // user code edited bodies since the last step: set it (or call markDirty) before stepping
  bool _eventsOn = false;

  CudaWorld(this.cuda, {super.gravity, super.frictionGravity, super.allowSleep, super.broadphase, super.solver, super.quatNormalizeFast,
      super.quatNormalizeSkip, int device = 0, this.solverKind = solverColored, this.syncEveryStep = true}) {
    session = CudaSession(cuda, this, device: device);
  }

  void markDirty() { hostDirty = true; }
  void syncBodies() { session.pullState(); }

  @override
  void internalStep(double dt) {
    final rebuilt = session.handle == nullptr || session._n != bodies.length || session._nConstraints != constraints.length;
    session.ensureUploaded(solverKind: solverKind);
    if (rebuilt && _eventsOn) session.check(cuda.enableContactEvents(session.handle, 1), 'cannon_world_enable_contact_events');
    if (hostDirty && !rebuilt) session.pushState();
    hostDirty = false;
    this.dt = dt;
    session.check(cuda.worldStep(session.handle, dt, 1), 'cannon_world_step');
    stepnumber += 1;
    time += dt;
    if (syncEveryStep) {
      session.pullState();
      for (final b in bodies) { b.force.setZero(); b.torque.setZero(); } // clearForces, world_class.dart:682
    }
    if (hasAnyEventListener('beginContact') || hasAnyEventListener('endContact')) _emitContactEvents();
  }

  /// World.raycastClosest (world_class.dart:269-277) against the device-resident poses.
  @override
  bool raycastClosest([Vector3? from, Vector3? to, RayOptions? options, RaycastResult? result]) {
    session.ensureUploaded(solverKind: solverKind);
    if (hostDirty) { session.pushState(); hostDirty = false; }
    final f = from ?? Vector3.zero(), t = to ?? Vector3.zero();
    return using((Arena a) {
      final pf = a<Float>(3), pt = a<Float>(3);
      for (var k = 0; k < 3; k++) { pf[k] = f[k]; pt[k] = t[k]; }
      final opt = a<CannonRayOptions>();
      cuda.rayOptionsDefault(opt);
      opt.ref.mode = rayClosest;
      opt.ref.skipBackfaces = (options?.skipBackfaces ?? true) ? 1 : 0;
      opt.ref.collisionFilterMask = options?.collisionFilterMask ?? -1;
      opt.ref.collisionFilterGroup = options?.collisionFilterGroup ?? -1;
      opt.ref.checkCollisionResponse = (options?.checkCollisionResponse ?? true) ? 1 : 0;
      final hits = a<CannonRayHitsSoa>();
      hits.ref.capacity = 1;
      hits.ref.ray = a<Int32>(); hits.ref.body = a<Int32>(); hits.ref.hitFaceIndex = a<Int32>(); hits.ref.distance = a<Double>();
      hits.ref.hitPointWorld = a<Float>(3); hits.ref.hitNormalWorld = a<Float>(3); hits.ref.shapeOrdinal = a<Int32>();
      final has = a<Uint8>(), n = a<Int32>();
      session.check(cuda.worldRaycast(session.handle, 1, pf, pt, opt, has, hits, n), 'cannon_world_raycast');
      final r = result ?? RaycastResult();
      r.reset();
      if (has.value != 0) {
        final b = bodies[hits.ref.body.value];
        final hp = hits.ref.hitPointWorld, hn = hits.ref.hitNormalWorld;
        r.set(f, t, Vector3(hn[0], hn[1], hn[2]), Vector3(hp[0], hp[1], hp[2]), b.shapes[hits.ref.shapeOrdinal.value], b, hits.ref.distance.value);
        r.hasHit = true;
        r.hitFaceIndex = hits.ref.hitFaceIndex.value;
      }
      return has.value != 0;
    });
  }

  // World.emitContactEvents (world_class.dart:703-730) from the device-side pair-set difference instead of OverlapKeeper
  void _emitContactEvents() {
    if (!_eventsOn) { session.check(cuda.enableContactEvents(session.handle, 1), 'cannon_world_enable_contact_events'); _eventsOn = true; return; }
    var cap = 4096;
    while (true) {
      final nb = calloc<Int32>(), ne = calloc<Int32>();
      final ba = calloc<Int32>(cap), bb = calloc<Int32>(cap), ea = calloc<Int32>(cap), eb = calloc<Int32>(cap);
      try {
        final rc = cuda.getContactEvents(session.handle, cap, nb, ba, bb, ne, ea, eb);
        if (rc == cannonECapacity) { cap = 2 * (nb.value > ne.value ? nb.value : ne.value); continue; }
        session.check(rc, 'cannon_world_get_contact_events');
        for (var k = 0; k < nb.value; k++) {
          beginContactEvent.bodyA = bodies[ba[k]];
          beginContactEvent.bodyB = bodies[bb[k]];
          dispatchEvent(beginContactEvent);
        }
        for (var k = 0; k < ne.value; k++) {
          endContactEvent.bodyA = bodies[ea[k]];
          endContactEvent.bodyB = bodies[eb[k]];
          dispatchEvent(endContactEvent);
        }
        return;
      } finally {
        for (final p in [nb, ne, ba, bb, ea, eb]) { calloc.free(p); }
      }
    }
  }
}

/// Drop-in for World.broadphase. [model] carries the broadphase parameters (kind, axis, grid bounds, useBoundingBoxes).
class CudaBroadphase extends Broadphase {
  final CudaSession session;
  final Broadphase model;
  CudaBroadphase(this.session, this.model) { useBoundingBoxes = model.useBoundingBoxes; }

  @override
  void collisionPairs(World world, List<Body> p1, List<Body> p2) {
    session.ensureUploaded();
    session.pushState();
    final cuda = session.cuda;
    var cap = 16 * world.bodies.length + 1024;
    while (true) {
      final a = calloc<Int32>(cap), b = calloc<Int32>(cap), n = calloc<Int32>();
      try {
        final rc = cuda.broadphasePairs(session.handle, a, b, cap, n);
        if (rc == cannonECapacity) { cap = n.value + 1024; continue; }
        session.check(rc, 'cannon_broadphase_pairs');
        for (var k = 0; k < n.value; k++) {
          p1.add(world.bodies[a[k]]);
          p2.add(world.bodies[b[k]]);
        }
        return;
      } finally {
        calloc.free(a); calloc.free(b); calloc.free(n);
      }
    }
  }
}

/// Drop-in for World.narrowphase (assign `world.narrowphase = CudaNarrowphase(world, session)`).
class CudaNarrowphase extends Narrowphase {
  final CudaSession session;
  CudaNarrowphase(World world, this.session) : super(world);

  @override
  void getContacts(List<Body> p1, List<Body> p2, World world, List<ContactEquation> result, List<ContactEquation> oldcontacts,
      List<FrictionEquation> frictionResult, List<FrictionEquation> frictionPool) {
    contactPointPool = oldcontacts;
    frictionEquationPool = frictionPool;
    this.result = result;
    this.frictionResult = frictionResult;
    session.ensureUploaded();
    session.pushState();
    final cuda = session.cuda;
    final np = p1.length;
    session.check(cuda.worldSetDt(session.handle, world.dt), 'cannon_world_set_dt');
    var cap = 8 * np + 1024;
    while (true) {
      final a = calloc<Int32>(np + 1), b = calloc<Int32>(np + 1), nc = calloc<Int32>();
      final out = calloc<CannonContactsSoa>();
      final bi = calloc<Int32>(cap), bj = calloc<Int32>(cap);
      final ri = calloc<Float>(3 * cap), rj = calloc<Float>(3 * cap), ni = calloc<Float>(3 * cap);
      try {
        for (var k = 0; k < np; k++) { a[k] = p1[k].index; b[k] = p2[k].index; }
        out.ref.capacity = cap; out.ref.bodyI = bi; out.ref.bodyJ = bj; out.ref.ri = ri; out.ref.rj = rj; out.ref.ni = ni;
        final rc = cuda.narrowphaseContacts(session.handle, a, b, np, out, nc, nullptr);
        if (rc == cannonECapacity) { cap = nc.value + 1024; continue; }
        session.check(rc, 'cannon_narrowphase_contacts');
        for (var k = 0; k < nc.value; k++) {
          final bodyI = world.bodies[bi[k]], bodyJ = world.bodies[bj[k]];
          final si = bodyI.shapes[0], sj = bodyJ.shapes[0];
          // the material choice of narrow_phase.dart:660-695, then the reference's own equation factories
          ContactMaterial? cm;
          if (si.material != null && sj.material != null) cm = world.getContactMaterial(si.material!, sj.material!);
          if (cm == null && bodyI.material != null && bodyJ.material != null) cm = world.getContactMaterial(bodyI.material!, bodyJ.material!);
          currentContactMaterial = cm ?? world.defaultContactMaterial;
          final c = createContactEquation(bodyI, bodyJ, si, sj);
          c.ri.setValues(ri[3 * k], ri[3 * k + 1], ri[3 * k + 2]);
          c.rj.setValues(rj[3 * k], rj[3 * k + 1], rj[3 * k + 2]);
          c.ni.setValues(ni[3 * k], ni[3 * k + 1], ni[3 * k + 2]);
          result.add(c);
          createFrictionEquationsFromContact(c, frictionResult);
        }
        return;
      } finally {
        for (final p in <Pointer>[a, b, nc, out, bi, bj, ri, rj, ni]) { calloc.free(p); }
      }
    }
  }
}

/// Drop-in for World.solver. Solves the contacts of the last CudaNarrowphase.getContacts call of the same session plus the
/// world's constraints on the device (the equation list the World filled through addEquation is what the device assembles
/// itself, in the same order: world_class.dart:539-541,562,627-635), then writes velocities and multipliers back.
class CudaGSSolver extends Solver {
  final CudaSession session;
  CudaGSSolver(this.session, {super.iterations, super.tolerance});

  @override
  int solve(double dt, World world) {
    final cuda = session.cuda;
    final it = calloc<Int32>();
    try {
      session.check(cuda.solverSolve(session.handle, dt, it), 'cannon_solver_solve');
      session.pullState(); // velocity / angularVelocity updated like gs_solver.dart:111-121
      // Equation.multiplier of the contact rows (gs_solver.dart:123-129)
      final n = world.contacts.length;
      if (n > 0) {
        final out = calloc<CannonContactsSoa>();
        final bi = calloc<Int32>(n), bj = calloc<Int32>(n), mult = calloc<Double>(n), nc = calloc<Int32>();
        final ri = calloc<Float>(3 * n), rj = calloc<Float>(3 * n), ni = calloc<Float>(3 * n);
        try {
          out.ref.capacity = n; out.ref.bodyI = bi; out.ref.bodyJ = bj; out.ref.ri = ri; out.ref.rj = rj; out.ref.ni = ni; out.ref.multiplier = mult;
          session.check(cuda.worldGetContacts(session.handle, out, nc), 'cannon_world_get_contacts');
          for (var k = 0; k < nc.value && k < n; k++) { world.contacts[k].multiplier = mult[k]; }
        } finally {
          for (final p in <Pointer>[out, bi, bj, mult, nc, ri, rj, ni]) { calloc.free(p); }
        }
      }
      return it.value;
    } finally {
      calloc.free(it);
    }
  }
}
