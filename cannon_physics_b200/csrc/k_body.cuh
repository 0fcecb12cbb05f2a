// k_body.cuh — per-body streaming kernels (HBM-bound, one float4 transaction per attribute).
//   k_prestep    gravity accumulation + AABB refit           world_class.dart:460-471, rigid_body.dart:415-447
//   k_presolve   wake flagged bodies, zero vlambda/wlambda    world_class.dart:618-624, gs_solver.dart:67-73
//   k_integrate  apply lambda, damping, Body.integrate, inertia refit, clearForces, sleepTick
//                gs_solver.dart:111-121, world_class.dart:648-700, rigid_body.dart:627-680,450-466,282-300
#pragma once
#include "world.cuh"

struct StepParams {
  double dt;
  const long long* clk;  // device clock: [0] bits of World.time before this step's increment (sleepTick sees this,
                         // world_class.dart:693), [1] World.stepnumber; kept on the device so a captured step can be replayed
  int quatSkip;          // quatNormalizeSkip
  double gx, gy, gz;  // gravity widened from float
  int n;              // bodies
  int allowSleep, quatNormalizeFast;
  int needAABB;
  int nWorlds;
  int deferSleepTick;  // springs in the world: sleepTick runs after the postStep slot (k_sleep_tick)
};

// Shape.calculateWorldAABB for the in-scope shapes
__device__ inline void shape_aabb(const ShapeTables& T, int shapeIdx, const f3& pos, const q4& q, f3& mn, f3& mx) {
  const float inf = __int_as_float(0x7f800000);
  if (shapeIdx < 0) { mn = mk3(0.0, 0.0, 0.0); mx = mn; return; }  // Body.updateAABB has nothing to loop over: the AABB stays AABB() (aabb.dart:13-16)
  const ShapeDev s = T.shapes[shapeIdx];
  switch (s.type) {
    case CANNON_SHAPE_SPHERE: {  // sphere.dart:43-53
      double r = s.radius;
      mn = mk3(W(pos.x) - r, W(pos.y) - r, W(pos.z) - r);
      mx = mk3(W(pos.x) + r, W(pos.y) + r, W(pos.z) + r);
      break;
    }
    case CANNON_SHAPE_PLANE: {  // plane.dart:44-69
      f3 z; z.x = 0.f; z.y = 0.f; z.z = 1.f;
      f3 n = qrot(q, z);
      mn.x = mn.y = mn.z = -inf;
      mx.x = mx.y = mx.z = inf;
      if (n.x == 1.f) mx.x = pos.x; else if (n.x == -1.f) mn.x = pos.x;
      if (n.y == 1.f) mx.y = pos.y; else if (n.y == -1.f) mn.y = pos.y;
      if (n.z == 1.f) mx.z = pos.z; else if (n.z == -1.f) mn.z = pos.z;
      break;
    }
    case CANNON_SHAPE_BOX: {  // box.dart:149-192 (corner order of _worldCornersTemp)
      const float ex = s.hx, ey = s.hy, ez = s.hz;
      const float sx[8] = {1, -1, -1, -1, 1, 1, -1, 1}, sy[8] = {1, 1, -1, -1, -1, 1, 1, -1}, sz[8] = {1, 1, 1, -1, -1, -1, -1, 1};
      for (int i = 0; i < 8; i++) {
        f3 c; c.x = sx[i] * ex; c.y = sy[i] * ey; c.z = sz[i] * ez;
        f3 w = vadd(qrot(q, c), pos);
        if (i == 0) { mn = w; mx = w; continue; }
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
      const HullDev h = T.hulls[s.hull];
      for (int i = 0; i < h.nV; i++) {
        f3 w = vadd(qrot(q, ld3(T.verts[h.vOff + i])), pos);
        if (i == 0) { mn = w; mx = w; continue; }
        if (w.x < mn.x) mn.x = w.x;
        if (w.x > mx.x) mx.x = w.x;
        if (w.y < mn.y) mn.y = w.y;
        if (w.y > mx.y) mx.y = w.y;
        if (w.z < mn.z) mn.z = w.z;
        if (w.z > mx.z) mx.z = w.z;
      }
      break;
    }
    case CANNON_SHAPE_PARTICLE:  // particle.dart:29-33
      mn = pos;
      mx = pos;
      break;
    case CANNON_SHAPE_TRIMESH: {  // trimesh.dart:366-375: AABB.toWorldFrame of the local AABB (aabb.dart:175-187,216-237)
      const TrimeshDev m = T.tms[s.tm];
      const f3 l = ld3(m.lo), u = ld3(m.hi);
      for (int i = 0; i < 8; i++) {
        f3 c;
        c.x = (i == 0 || i == 3 || i == 5 || i == 6) ? l.x : u.x;
        c.y = (i == 0 || i == 1 || i == 4 || i == 6) ? l.y : u.y;
        c.z = (i == 0 || i == 1 || i == 2 || i == 5) ? l.z : u.z;
        const f3 w = to_world_point(pos, q, c);
        if (i == 0) { mn = w; mx = w; continue; }
        if (w.x > mx.x) mx.x = w.x;
        if (w.x < mn.x) mn.x = w.x;
        if (w.y > mx.y) mx.y = w.y;
        if (w.y < mn.y) mn.y = w.y;
        if (w.z > mx.z) mx.z = w.z;
        if (w.z < mn.z) mn.z = w.z;
      }
      break;
    }
    default:  // heightfield.dart:499-503
      mn.x = mn.y = mn.z = -inf;
      mx.x = mx.y = mx.z = inf;
      break;
  }
}

// Body.updateAABB, rigid_body.dart:415-447: the union (AABB.extend, aabb.dart:121-128) of the shape AABBs at their
// world poses; a body without a shape table entry is the single-shape case
__device__ inline void body_aabb(const BodyArrays& B, const ShapeTables& T, int i, f3& mn, f3& mx) {
  const f3 pos = ld3(B.pos[i]);
  const q4 q = ldq(B.quat[i]);
  if (!T.instFirst) { shape_aabb(T, B.shape[i], pos, q, mn, mx); return; }
  const int k0 = T.instFirst[i], k1 = T.instFirst[i + 1];
  if (k0 == k1) { mn = mk3(0.0, 0.0, 0.0); mx = mn; return; }
  for (int k = k0; k < k1; k++) {
    const f3 offset = vadd(qrot(q, ld3(T.instOff[k])), pos);
    const q4 orientation = qmul(q, ldq(T.instQuat[k]));
    f3 lo, hi;
    shape_aabb(T, T.instShape[k], offset, orientation, lo, hi);
    if (k == k0) { mn = lo; mx = hi; continue; }
    mn.x = fminf(mn.x, lo.x); mn.y = fminf(mn.y, lo.y); mn.z = fminf(mn.z, lo.z);
    mx.x = fmaxf(mx.x, hi.x); mx.y = fmaxf(mx.y, hi.y); mx.z = fmaxf(mx.z, hi.z);
  }
}

__global__ void __launch_bounds__(256) k_prestep(BodyArrays B, ShapeTables T, StepParams P, int doGravity) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += gridDim.x * blockDim.x) {
    if (doGravity && B.type[i] == CANNON_BODY_DYNAMIC) {
      float4 f = B.force[i];
      const double m = B.mass[i];
      f.x = (float)(W(f.x) + m * P.gx);
      f.y = (float)(W(f.y) + m * P.gy);
      f.z = (float)(W(f.z) + m * P.gz);
      B.force[i] = f;
    }
    if (P.needAABB) {
      f3 mn, mx;
      body_aabb(B, T, i, mn, mx);
      B.aabbLo[i] = st3(mn);
      B.aabbHi[i] = st3(mx);
    }
  }
}

// wake-up after narrowphase + lambda reset. worldRows[w] > 0 <=> world w has equations this step.
__global__ void __launch_bounds__(256) k_presolve(BodyArrays B, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int fl = B.flags[i];
    if (fl & BF_WAKE) {
      B.sleep[i] = CANNON_AWAKE;  // Body.wakeUp, rigid_body.dart:263-270
      B.flags[i] = fl & ~BF_WAKE;
    }
    B.vlam[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
    B.wlam[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Algorithmic HBM bytes per body: see the K1 table; every array is touched once.
// Body.sleepTick (rigid_body.dart:282-300); true when the velocities were zeroed
__device__ __forceinline__ bool sleep_tick(const BodyArrays& B, const StepParams& P, int i, int sleep, f3& v, f3& w) {
  const double speedSquared = vlen2(v) + vlen2(w);
  const double lim = B.sleepSpeed[i];
  const double speedLimitSquared = lim * lim;
  if (sleep == CANNON_AWAKE && speedSquared < speedLimitSquared) {
    B.tLastSleepy[i] = __longlong_as_double(P.clk[0]);
    B.sleep[i] = CANNON_SLEEPY;
  } else if (sleep == CANNON_SLEEPY && speedSquared > speedLimitSquared) {
    B.sleep[i] = CANNON_AWAKE;
  } else if (sleep == CANNON_SLEEPY && __longlong_as_double(P.clk[0]) - B.tLastSleepy[i] > B.sleepTime[i]) {
    B.sleep[i] = CANNON_SLEEPING;
    v.x = v.y = v.z = 0.f;
    w.x = w.y = w.z = 0.f;
    return true;
  }
  return false;
}

__global__ void __launch_bounds__(256) k_integrate(BodyArrays B, StepParams P, const int* __restrict__ worldRows, int applyLambda) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += gridDim.x * blockDim.x) {
    const int type = B.type[i];
    int sleep = B.sleep[i];
    const int flags = B.flags[i];
    float4 v4 = B.vel[i], w4 = B.angvel[i];
    f3 v = ld3(v4), w = ld3(w4);
    const f3 angF = ld3(B.angF[i]);
    const f3 linF = ld3(B.linF[i]);
    bool dirtyVel = false;

    // gs_solver.dart:111-121 — only when the (body's) world had equations this step
    if (applyLambda && worldRows[P.nWorlds > 1 ? B.world[i] : 0] > 0) {
      f3 vl = vmulc(ld3(B.vlam[2 * i]), linF);
      v = vadd(vl, v);
      f3 wl = vmulc(ld3(B.wlam[2 * i]), angF);
      w = vadd(wl, w);
      dirtyVel = true;
    }
    // damping, world_class.dart:648-659 (pow(1-d, dt) is evaluated on the host with libm)
    if (type == CANNON_BODY_DYNAMIC) {
      v = vscale(B.ldpow[i], v);
      w = vscale(B.adpow[i], w);
      dirtyVel = true;
    }
    // Body.integrate, rigid_body.dart:627-680
    const bool moves = (type == CANNON_BODY_DYNAMIC || type == CANNON_BODY_KINEMATIC) && sleep != CANNON_SLEEPING;
    if (moves) {
      const float4 f4 = B.force[i], t4 = B.torque[i];
      const double iMdt = B.invMass[i] * P.dt;
      v.x = (float)(W(v.x) + W(f4.x) * iMdt * W(linF.x));
      v.y = (float)(W(v.y) + W(f4.y) * iMdt * W(linF.y));
      v.z = (float)(W(v.z) + W(f4.z) * iMdt * W(linF.z));
      float4 r0 = B.iiw0[i], r1 = B.iiw1[i], r2 = B.iiw2[i];
      const double tx = W(t4.x) * W(angF.x), ty = W(t4.y) * W(angF.y), tz = W(t4.z) * W(angF.z);
      w.x = (float)(W(w.x) + P.dt * (W(r0.x) * tx + W(r0.y) * ty + W(r0.z) * tz));
      w.y = (float)(W(w.y) + P.dt * (W(r1.x) * tx + W(r1.y) * ty + W(r1.z) * tz));
      w.z = (float)(W(w.z) + P.dt * (W(r2.x) * tx + W(r2.y) * ty + W(r2.z) * tz));
      float4 p4 = B.pos[i];
      p4.x = (float)(W(p4.x) + W(v.x) * P.dt);
      p4.y = (float)(W(p4.y) + W(v.y) * P.dt);
      p4.z = (float)(W(p4.z) + W(v.z) * P.dt);
      B.pos[i] = p4;
      // Quat.integrate, quaternion.dart:93-111
      q4 q = ldq(B.quat[i]);
      const double ax = W(w.x) * W(angF.x), ay = W(w.y) * W(angF.y), az = W(w.z) * W(angF.z);
      const double bx = W(q.x), by = W(q.y), bz = W(q.z), bw = W(q.w);
      const double halfDt = P.dt * 0.5;
      q.x = (float)(bx + halfDt * (ax * bw + ay * bz - az * by));
      q.y = (float)(by + halfDt * (ay * bw + az * bx - ax * bz));
      q.z = (float)(bz + halfDt * (az * bw + ax * by - ay * bx));
      q.w = (float)(bw + halfDt * (-ax * bx - ay * by - az * bz));
      if (P.clk[1] % (long long)(P.quatSkip + 1) == 0) {  // world_class.dart:668
        if (P.quatNormalizeFast) {  // quaternion.dart:171-185
          const double f = (3.0 - (W(q.x) * W(q.x) + W(q.y) * W(q.y) + W(q.z) * W(q.z) + W(q.w) * W(q.w))) / 2.0;
          if (f == 0) { q.x = q.y = q.z = q.w = 0.f; }
          else { q.x = (float)(W(q.x) * f); q.y = (float)(W(q.y) * f); q.z = (float)(W(q.z) * f); q.w = (float)(W(q.w) * f); }
        } else {  // Quaternion.normalize()
          const double l = sqrt((W(q.x) * W(q.x)) + (W(q.y) * W(q.y)) + (W(q.z) * W(q.z)) + (W(q.w) * W(q.w)));
          if (l != 0.0) {
            const double d = 1.0 / l;
            q.x = (float)(W(q.x) * d); q.y = (float)(W(q.y) * d); q.z = (float)(W(q.z) * d); q.w = (float)(W(q.w) * d);
          }
        }
      }
      B.quat[i] = stq(q);
      // updateInertiaWorld(), skipped for isotropic inverse inertia (rigid_body.dart:452)
      const f3 I = ld3(B.invI[i]);
      if (!(I.x == I.y && I.y == I.z)) {
        inertia_world(q, I, r0, r1, r2);
        B.iiw0[i] = r0; B.iiw1[i] = r1; B.iiw2[i] = r2;
      }
      dirtyVel = true;
    }
    // clearForces, world_class.dart:773-781
    B.force[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    B.torque[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    // sleepTick, rigid_body.dart:282-300 - after the postStep slot (world_class.dart:685-699): with springs in the world it
    // runs in k_sleep_tick, behind k_springs, so Spring.applyForce still sees the velocities of a body that falls asleep now
    if (P.allowSleep && !P.deferSleepTick && (flags & BF_ALLOW_SLEEP)) dirtyVel |= sleep_tick(B, P, i, sleep, v, w);
    if (dirtyVel) {
      B.vel[i] = st3(v);
      B.angvel[i] = st3(w);
    }
  }
}

// Spring.applyForce (lib/objects/spring.dart:108-157) in the postStep slot of the step. The reference walks the springs
// in listener order and every body's force / torque is a float accumulated in that order; here one thread owns a body
// and walks THAT body's springs in the same order (CSR built at cannon_world_set_springs), recomputing the spring force
// from the (unchanged) poses and velocities, so every accumulator sees the reference's sequence of float additions.
struct SpringArrays {
  int n;
  const int *bodyA, *bodyB;
  const double *rest, *stiffness, *damping;
  const float4 *anchorA, *anchorB;
  const int *off, *idx;  // body -> springs touching it, ascending
};

// the deferred sleepTick of a world with springs (world_class.dart:691-699 comes after the postStep event of :685)
__global__ void __launch_bounds__(256) k_sleep_tick(BodyArrays B, StepParams P) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += gridDim.x * blockDim.x) {
    if (!(B.flags[i] & BF_ALLOW_SLEEP)) continue;
    f3 v = ld3(B.vel[i]), w = ld3(B.angvel[i]);
    if (sleep_tick(B, P, i, B.sleep[i], v, w)) { B.vel[i] = st3(v); B.angvel[i] = st3(w); }
  }
}

__global__ void __launch_bounds__(256) k_springs(BodyArrays B, SpringArrays S, int nBodies) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nBodies; b += gridDim.x * blockDim.x) {
    const int s0 = S.off[b], s1 = S.off[b + 1];
    if (s0 == s1) continue;
    f3 force = ld3(B.force[b]), torque = ld3(B.torque[b]);
    for (int k = s0; k < s1; k++) {
      const int sp = S.idx[k];
      const int a = S.bodyA[sp], c = S.bodyB[sp];
      const f3 xa = ld3(B.pos[a]), xc = ld3(B.pos[c]);
      const f3 worldAnchorA = to_world_point(xa, ldq(B.quat[a]), ld3(S.anchorA[sp]));  // pointToWorldFrame, rigid_body.dart:332-337
      const f3 worldAnchorB = to_world_point(xc, ldq(B.quat[c]), ld3(S.anchorB[sp]));
      const f3 ri = vsub(worldAnchorA, xa), rj = vsub(worldAnchorB, xc);
      const f3 r = vsub(worldAnchorB, worldAnchorA);
      const double rlen = vlen(r);
      f3 rUnit = r;
      vnormalize(rUnit);
      f3 u = vsub(ld3(B.vel[c]), ld3(B.vel[a]));
      u = vadd(u, vcross(ld3(B.angvel[c]), rj));
      u = vsub(u, vcross(ld3(B.angvel[a]), ri));
      const f3 f = vscale(-S.stiffness[sp] * (rlen - S.rest[sp]) - S.damping[sp] * vdot(u, rUnit), rUnit);
      if (a == b) { force = vsub(force, f); torque = vsub(torque, vcross(ri, f)); }
      if (c == b) { force = vadd(force, f); torque = vadd(torque, vcross(rj, f)); }
    }
    B.force[b] = st3(force);
    B.torque[b] = st3(torque);
  }
}
