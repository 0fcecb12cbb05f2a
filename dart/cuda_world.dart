// cuda_world.dart — drop-in classes for the reference's pluggable step stages (sketch, see INTEGRATION.md).
//
//   CudaWorld        extends World, overrides internalStep(dt): state device-resident, one FFI call per step
//   CudaBroadphase   extends Broadphase: collisionPairs(world, p1, p2) through cannon_broadphase_pairs
//   CudaGSSolver     extends Solver: solve(dt, world) through cannon_solver_solve (after CudaNarrowphase)
import 'dart:ffi';
import 'package:ffi/ffi.dart';
import 'package:cannon_physics/cannon_physics.dart';
import 'cannon_cuda_bindings.dart';

class CudaWorld extends World {
  final CannonCuda cuda;
  Pointer<Void> _ctx = nullptr, _world = nullptr;
  bool _uploaded = false;
  CudaWorld(this.cuda, {super.gravity, super.allowSleep, super.broadphase, super.solver});

  void _upload() {
    // flatten bodies / shapes / materials / constraints into the SoA structs of cannon_cuda.h
    // (one Float32List view per attribute; see cannon_physics_b200/api.py World._spec for the exact mapping)
    _uploaded = true;
  }

  @override
  void internalStep(double dt) {
    if (!_uploaded) _upload();
    final rc = cuda.worldStep(_world, dt, 1);
    if (rc != 0) throw cuda.lastError(_ctx).toDartString(); // the reference throws strings
    // poses are downloaded lazily: cannon_world_get_bodies fills position/quaternion views on demand
    stepnumber += 1;
    if (hasAnyEventListener('beginContact') || hasAnyEventListener('endContact')) _emitContactEvents();
  }

  // World.emitContactEvents (world_class.dart:703-730) from the device-side pair-set difference instead of OverlapKeeper
  bool _eventsOn = false;
  void _emitContactEvents() {
    if (!_eventsOn) { cuda.enableContactEvents(_world, 1); _eventsOn = true; return; } // tracking starts with the next step
    final cap = 4096;
    final nb = calloc<Int32>(), ne = calloc<Int32>();
    final ba = calloc<Int32>(cap), bb = calloc<Int32>(cap), ea = calloc<Int32>(cap), eb = calloc<Int32>(cap);
    try {
      final rc = cuda.getContactEvents(_world, cap, nb, ba, bb, ne, ea, eb);
      if (rc != 0) throw cuda.lastError(_ctx).toDartString();
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
    } finally {
      for (final p in [nb, ne, ba, bb, ea, eb]) { calloc.free(p); }
    }
  }
}

class CudaBroadphase extends Broadphase {
  final CannonCuda cuda;
  final Pointer<Void> handle;
  CudaBroadphase(this.cuda, this.handle);
  @override
  void collisionPairs(World world, List<Body> p1, List<Body> p2) {
    final cap = 16 * world.bodies.length + 1024;
    final a = calloc<Int32>(cap), b = calloc<Int32>(cap), n = calloc<Int32>();
    try {
      final rc = cuda.broadphasePairs(handle, a, b, cap, n);
      if (rc != 0) throw 'cannon_broadphase_pairs failed ($rc)';
      for (var k = 0; k < n.value; k++) {
        p1.add(world.bodies[a[k]]);
        p2.add(world.bodies[b[k]]);
      }
    } finally {
      calloc.free(a); calloc.free(b); calloc.free(n);
    }
  }
}
