// k_solver.cuh — equation rows (K4), dependency-level scheduling and the Gauss-Seidel sweeps (K5).
//
// Reference: GSSolver.solve (lib/solver/gs_solver.dart:27-133) iterates over the equations in insertion
// order: all FrictionEquations, then all ContactEquations, then the constraints' equations
// (lib/world/world_class.dart:539-541,562,627-635). Gauss-Seidel is order dependent, so:
//
//   REFERENCE_ORDER  rows keep the reference's order and are executed by *dependency levels*: row r gets
//                    level 1 + max(level of the previous row touching either of its movable bodies). Rows of
//                    one level touch disjoint movable bodies, so running them concurrently performs exactly
//                    the same floating-point operations on exactly the same operands as the sequential
//                    loop: results are bit-identical to the reference order while still parallel.
//   COLORED          the same machinery with units = contact manifolds (all contacts of one resolver task,
//                    rows [f1,f2,n] per contact) and hashed priorities, i.e. a randomised greedy colouring
//                    with O(max degree) colours (statistical agreement only).
//
// Both run as persistent cooperative kernels (one launch per solve) with a hand-rolled grid barrier between
// levels: launch latency would otherwise dominate (iterations x levels dependent launches).
#pragma once
#include "k_narrowphase.cuh"
#include "world.cuh"

enum { ROW_CONTACT = 0, ROW_FRICTION = 1, ROW_ROT = 2, ROW_MOTOR = 3, ROW_OFF = -1 };

struct RowArrays {
  int* nRows;        // rows in storage (device)
  int *bi, *bj, *kind;
  float4 *n;         // spatial Jacobian of body j (sB); body i gets -n (zero for rotational rows)
  float4 *rA, *rB;   // rotational Jacobians
  float4 *iA, *iB;   // invInertiaWorldSolve * rA / rB, rounded to float like the reference's temp vector
  double *B, *invC, *eps, *minF, *maxF, *imA, *imB, *lambda;
  int* flags;        // bit0: body i movable, bit1: body j movable
  int rowCap;
};

struct JointArrays {  // one entry per constraint equation (P2P: 3, hinge: 6), uploaded at set_constraints
  int n;
  const int *bodyA, *bodyB, *kind, *enabled, *rowSlot;  // rowSlot: index among accepted joint rows or -1
  const float4 *pivotA, *pivotB, *axisA, *axisB;        // local frame
  const float4 *ni;                                     // P2P rows: world x / y / z
  const double *minF, *maxF, *a, *b, *eps, *targetVel;
  const int *first;                                     // index of the constraint's first equation (for hinge tangents)
  int nAccepted;
  double cosMaxAngle;  // cos(RotationalEquation.maxAngle = pi/2), evaluated on the host with libm
};

struct SolveParams {
  double dt;
  double tol2;
  int maxIter;
  int nBodies;
  int nWorlds;
  int colored;
};

// rigid_body.dart:303-314
__device__ __forceinline__ bool body_frozen(const BodyArrays& B, int b) { return B.sleep[b] == CANNON_SLEEPING || B.type[b] == CANNON_BODY_KINEMATIC; }

struct RowBody {
  f3 pos, vel, angvel, force, torque;
  float4 r0, r1, r2;  // invInertiaWorldSolve rows
  double im;          // invMassSolve
  bool movable;
};
__device__ __forceinline__ void load_row_body(const BodyArrays& B, int b, RowBody& r) {
  r.pos = ld3(B.pos[b]); r.vel = ld3(B.vel[b]); r.angvel = ld3(B.angvel[b]);
  r.force = ld3(B.force[b]); r.torque = ld3(B.torque[b]);
  if (body_frozen(B, b)) {
    r.im = 0.0;
    r.r0 = r.r1 = r.r2 = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    r.im = B.invMass[b];
    r.r0 = B.iiw0[b]; r.r1 = B.iiw1[b]; r.r2 = B.iiw2[b];
  }
  r.movable = r.im != 0.0 || r.r0.x != 0.f || r.r0.y != 0.f || r.r0.z != 0.f || r.r1.x != 0.f || r.r1.y != 0.f || r.r1.z != 0.f ||
              r.r2.x != 0.f || r.r2.y != 0.f || r.r2.z != 0.f;
}

// computeGiMf (equation_class.dart:108-127), computeC (:130-148,172-174) and row store
__device__ inline void finish_row(const RowArrays& R, int row, int kind, int bi, int bj, const RowBody& A, const RowBody& Bd, const f3& sA,
                                  const f3& rA, const f3& sB, const f3& rB, double gterm /* -g*a or 0 */, double gw, double b, double eps,
                                  double minF, double maxF, double h) {
  const f3 iMfi = vscale(A.im, A.force), iMfj = vscale(Bd.im, Bd.force);
  const f3 iTi = mrow_mul(A.r0, A.r1, A.r2, A.torque), iTj = mrow_mul(Bd.r0, Bd.r1, Bd.r2, Bd.torque);
  const double giMf = (vdot(iMfi, sA) + vdot(iTi, rA)) + (vdot(iMfj, sB) + vdot(iTj, rB));
  const double Bv = gterm - gw * b - h * giMf;
  const f3 iA = mrow_mul(A.r0, A.r1, A.r2, rA), iB = mrow_mul(Bd.r0, Bd.r1, Bd.r2, rB);
  double c = A.im + Bd.im;
  c += vdot(iA, rA);
  c += vdot(iB, rB);
  c += eps;
  R.bi[row] = bi; R.bj[row] = bj; R.kind[row] = kind;
  R.n[row] = st3(sB); R.rA[row] = st3(rA); R.rB[row] = st3(rB); R.iA[row] = st3(iA); R.iB[row] = st3(iB);
  R.B[row] = Bv; R.invC[row] = 1.0 / c; R.eps[row] = eps; R.minF[row] = minF; R.maxF[row] = maxF;
  R.imA[row] = A.im; R.imB[row] = Bd.im; R.lambda[row] = 0.0;
  R.flags[row] = (A.movable ? 1 : 0) | (Bd.movable ? 2 : 0);
}

// per contact: acceptance flags (Solver.addEquation filter, solver.dart:30-34) + wake-up flags (world_class.dart:564-590)
__global__ void __launch_bounds__(256) k_contact_flags(BodyArrays B, ContactArrays C, int contactCap, int* __restrict__ fricFlag,
                                                       int* __restrict__ contFlag, int allowSleepWorld) {
  const int nc = min(*C.nContacts, contactCap);
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
    const int bi = C.bi[c], bj = C.bj[c];
    const int fi = B.flags[bi], fj = B.flags[bj];
    const bool ok = C.enabled[c] && !(fi & BF_IS_TRIGGER) && !(fj & BF_IS_TRIGGER);
    contFlag[c] = ok ? 1 : 0;
    fricFlag[c] = (ok && C.mu[c] > 0) ? 1 : 0;
    const int si = B.sleep[bi], sj = B.sleep[bj];
    const int ti = B.type[bi], tj = B.type[bj];
    if ((fi & BF_ALLOW_SLEEP) && ti == CANNON_BODY_DYNAMIC && si == CANNON_SLEEPING && sj == CANNON_AWAKE && tj != CANNON_BODY_STATIC) {
      const double s2 = vlen2(ld3(B.vel[bj])) + vlen2(ld3(B.angvel[bj]));
      const double lim = B.sleepSpeed[bj];
      if (s2 >= lim * lim * 2) atomicOr(&B.flags[bi], BF_WAKE);
    }
    if ((fj & BF_ALLOW_SLEEP) && tj == CANNON_BODY_DYNAMIC && sj == CANNON_SLEEPING && si == CANNON_AWAKE && ti != CANNON_BODY_STATIC) {
      const double s2 = vlen2(ld3(B.vel[bi])) + vlen2(ld3(B.angvel[bi]));
      const double lim = B.sleepSpeed[bi];
      if (s2 >= lim * lim * 2) atomicOr(&B.flags[bj], BF_WAKE);
    }
  }
}

// rows of the contacts: ContactEquation.computeB (contact_equation.dart:34-77), FrictionEquation.computeB
// (friction_equation.dart:19-47). Row index: reference order [2*fricRank | nF2 + contRank], coloured [3c, 3c+1, 3c+2].
__global__ void __launch_bounds__(128) k_rows_contacts(BodyArrays B, ContactArrays C, RowArrays R, SolveParams S, int contactCap,
                                                       const int* __restrict__ fricOff, const int* __restrict__ contOff,
                                                       const int* __restrict__ fricTotal, const int* __restrict__ contTotal,
                                                       int* __restrict__ worldRows, int* __restrict__ rowOverflow, int nJointRows,
                                                       int* __restrict__ nContactRows) {
  const int nc = min(*C.nContacts, contactCap);
  const int nF2 = 2 * (*fricTotal);
  const int nCR = S.colored ? 3 * nc : nF2 + *contTotal;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *nContactRows = nCR;            // joint rows are appended behind by k_rows_joints
    *R.nRows = nCR + nJointRows;
    if (nCR + nJointRows > R.rowCap) atomicMax(rowOverflow, nCR + nJointRows);
  }
  if (nCR + nJointRows > R.rowCap) return;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
    const int f0 = fricOff[c], c0 = contOff[c];
    const bool hasF = (c + 1 < nc ? fricOff[c + 1] : *fricTotal) > f0;
    const bool hasC = (c + 1 < nc ? contOff[c + 1] : *contTotal) > c0;
    int rowF1, rowF2, rowN;
    if (S.colored) { rowF1 = 3 * c; rowF2 = 3 * c + 1; rowN = 3 * c + 2; }
    else { rowF1 = 2 * f0; rowF2 = 2 * f0 + 1; rowN = nF2 + c0; }
    if (S.colored) {
      if (!hasF) { R.kind[rowF1] = ROW_OFF; R.kind[rowF2] = ROW_OFF; }
      if (!hasC) R.kind[rowN] = ROW_OFF;
    }
    C.row[c] = hasC ? rowN : -1;
    if (!hasC && !hasF) continue;
    const int bi = C.bi[c], bj = C.bj[c];
    RowBody A, Bd;
    load_row_body(B, bi, A);
    load_row_body(B, bj, Bd);
    const f3 ri = ld3(C.ri[c]), rj = ld3(C.rj[c]), ni = ld3(C.ni[c]);
    const double h = S.dt;
    if (hasC) {
      const f3 rixn = vcross(ri, ni), rjxn = vcross(rj, ni);
      f3 pen = vadd(Bd.pos, rj);
      pen = vsub(pen, A.pos);
      pen = vsub(pen, ri);
      const double g = vdot(ni, pen);
      const double ePlusOne = C.rest[c] + 1;
      const double gw = ePlusOne * vdot(Bd.vel, ni) - ePlusOne * vdot(A.vel, ni) + vdot(Bd.angvel, rjxn) - vdot(A.angvel, rixn);
      finish_row(R, rowN, ROW_CONTACT, bi, bj, A, Bd, vneg(ni), vneg(rixn), ni, rjxn, -g * C.ca[c], gw, C.cb[c], C.ceps[c], 0.0, 1e6, h);
      atomicAdd(&worldRows[S.nWorlds > 1 ? B.world[bi] : 0], 1);
    }
    if (hasF) {
      f3 t1, t2;
      vtangents(ni, t1, t2);
      const double slip = C.slip[c];
      for (int k = 0; k < 2; k++) {
        const f3 t = k ? t2 : t1;
        const f3 rixt = vcross(ri, t), rjxt = vcross(rj, t);
        const f3 sA = vneg(t), rA = vneg(rixt);
        const double gw = (vdot(A.vel, sA) + vdot(A.angvel, rA)) + (vdot(Bd.vel, t) + vdot(Bd.angvel, rjxt));
        finish_row(R, k ? rowF2 : rowF1, ROW_FRICTION, bi, bj, A, Bd, sA, rA, t, rjxt, 0.0, gw, C.fb[c], C.feps[c], -slip, slip, h);
      }
      atomicAdd(&worldRows[S.nWorlds > 1 ? B.world[bi] : 0], 2);
    }
  }
}

// Constraint.update() + equation rows of the joints (point_to_point_constraint.dart:68-83, hinge_constraint.dart:79-104,
// rotational_equation.dart:34-58, rotational_motor_equation.dart:17-33). One thread per constraint equation.
__global__ void __launch_bounds__(128) k_rows_joints(BodyArrays B, JointArrays J, RowArrays R, SolveParams S, const int* __restrict__ nContactRows, int* __restrict__ worldRows,
                                                     int* __restrict__ rowOverflow) {
  const int base = *nContactRows;
  if (base + J.nAccepted > R.rowCap) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(rowOverflow, base + J.nAccepted);
    return;
  }
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < J.n; e += gridDim.x * blockDim.x) {
    const int slot = J.rowSlot[e];
    if (slot < 0) continue;
    const int row = base + slot;
    const int bi = J.bodyA[e], bj = J.bodyB[e];
    RowBody A, Bd;
    load_row_body(B, bi, A);
    load_row_body(B, bj, Bd);
    const q4 qa = ldq(B.quat[bi]), qb = ldq(B.quat[bj]);
    const double h = S.dt;
    const int kind = J.kind[e];
    f3 zero; zero.x = zero.y = zero.z = 0.f;
    if (kind == ROW_CONTACT) {
      const f3 ri = qrot(qa, ld3(J.pivotA[e])), rj = qrot(qb, ld3(J.pivotB[e]));
      const f3 ni = ld3(J.ni[e]);
      const f3 rixn = vcross(ri, ni), rjxn = vcross(rj, ni);
      f3 pen = vadd(Bd.pos, rj);
      pen = vsub(pen, A.pos);
      pen = vsub(pen, ri);
      const double g = vdot(ni, pen);
      const double ePlusOne = 0.0 + 1;
      const double gw = ePlusOne * vdot(Bd.vel, ni) - ePlusOne * vdot(A.vel, ni) + vdot(Bd.angvel, rjxn) - vdot(A.angvel, rixn);
      finish_row(R, row, ROW_CONTACT, bi, bj, A, Bd, vneg(ni), vneg(rixn), ni, rjxn, -g * J.a[e], gw, J.b[e], J.eps[e], J.minF[e], J.maxF[e], h);
    } else if (kind == ROW_ROT) {
      const f3 worldAxisA = qrot(qa, ld3(J.axisA[e])), worldAxisB = qrot(qb, ld3(J.axisB[e]));
      f3 t1, t2;
      vtangents(worldAxisA, t1, t2);
      const f3 axA = (e - J.first[e] == 3) ? t1 : t2;  // rotationalEquation1 / rotationalEquation2
      const f3 nixnj = vcross(axA, worldAxisB), njxni = vcross(worldAxisB, axA);
      const double g = J.cosMaxAngle - vdot(axA, worldAxisB);
      const double gw = (vdot(A.vel, zero) + vdot(A.angvel, njxni)) + (vdot(Bd.vel, zero) + vdot(Bd.angvel, nixnj));
      finish_row(R, row, ROW_ROT, bi, bj, A, Bd, zero, njxni, zero, nixnj, -g * J.a[e], gw, J.b[e], J.eps[e], J.minF[e], J.maxF[e], h);
    } else {
      const f3 axA = qrot(qa, ld3(J.axisA[e])), axB = qrot(qb, ld3(J.axisB[e]));
      const f3 rB = vneg(axB);
      const double gw = ((vdot(A.vel, zero) + vdot(A.angvel, axA)) + (vdot(Bd.vel, zero) + vdot(Bd.angvel, rB))) - J.targetVel[e];
      finish_row(R, row, ROW_MOTOR, bi, bj, A, Bd, zero, axA, zero, rB, 0.0, gw, J.b[e], J.eps[e], J.minF[e], J.maxF[e], h);
    }
    atomicAdd(&worldRows[S.nWorlds > 1 ? B.world[bi] : 0], 1);
  }

}

// ---- grid barrier -----------------------------------------------------------------------------------
// Monotonic-counter barrier for cooperative (co-resident) launches. `bar` is zeroed before the launch.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    __threadfence();
    atomicAdd(bar, 1u);
    while (*(volatile unsigned*)bar < epoch) {}
    __threadfence();
  }
  __syncthreads();
}

struct SchedArrays {
  unsigned long long* claim;  // per body
  int* unitLevel;             // per unit, -1 = unassigned
  int* order;                 // units sorted by level
  int* levelStart;            // [maxLevels+1]
  int* nLevels;
  int *act0, *act1;           // active (unassigned) unit lists
  int* actCount;              // [2]
  int* cursor;                // append cursor into order
  unsigned* bar;
  int maxLevels;
  int* levelOverflow;
  // units
  int nUnitsFixed;            // <0: read from nUnitsPtr
  const int* nUnitsPtr;
};

// unit -> its rows. reference mode: unit == row. coloured: unit == resolver task (rows 3*taskOff .. 3*(taskOff+cnt)) or a joint row.
struct UnitMap {
  int colored;
  const int* taskOff;  // per task: first contact
  const int* taskCnt;
  const int* nTasks;
  int taskCap;
  const int* nContacts;
  int contactCap;
};
__device__ __forceinline__ void unit_rows(const UnitMap& U, int u, int nRows, int& r0, int& r1) {
  if (!U.colored) { r0 = u; r1 = u + 1; return; }
  const int nt = min(*U.nTasks, U.taskCap);
  if (u < nt) { r0 = 3 * U.taskOff[u]; r1 = r0 + 3 * U.taskCnt[u]; }
  else { r0 = 3 * min(*U.nContacts, U.contactCap) + (u - nt); r1 = r0 + 1; }
  if (r1 > nRows) r1 = nRows;
}
__device__ __forceinline__ int unit_count(const UnitMap& U, int nRows) {
  if (!U.colored) return nRows;
  const int nt = min(*U.nTasks, U.taskCap);
  return nt + (nRows - 3 * min(*U.nContacts, U.contactCap));
}

// Dependency levels by repeated "claim the bodies with the smallest pending priority": a unit is released in
// the round in which it holds the minimum on all of its movable bodies, which is exactly
// level(u) = 1 + max(level of earlier units sharing a movable body). Keys carry the round in the high bits so
// stale claims of earlier rounds always lose the atomicMin (no clearing pass).
__global__ void __launch_bounds__(256) k_schedule(RowArrays R, SchedArrays S, UnitMap U) {
  unsigned epoch = 0;
  const int nRows = min(*R.nRows, R.rowCap);
  const int nUnits = unit_count(U, nRows);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int u = tid; u < nUnits; u += nth) { S.act0[u] = u; S.unitLevel[u] = -1; }
  if (tid == 0) { S.actCount[0] = nUnits; S.actCount[1] = 0; *S.cursor = 0; S.levelStart[0] = 0; }
  grid_barrier(S.bar, epoch);
  int round = 0;
  int cur = 0;
  while (true) {
    const int nAct = __ldcg(&S.actCount[cur]);
    if (nAct == 0) break;
    if (round >= S.maxLevels) { if (tid == 0) *S.levelOverflow = 1; break; }
    const int* act = cur ? S.act1 : S.act0;
    int* nxt = cur ? S.act0 : S.act1;
    const unsigned long long hi = (unsigned long long)(0x7fffffffu - (unsigned)round) << 32;
    for (int a = tid; a < nAct; a += nth) {
      const int u = __ldcg(&act[a]);
      const unsigned pri = U.colored ? (unsigned)u * 2654435761u : (unsigned)u;
      const unsigned long long key = hi | pri;
      int r0, r1;
      unit_rows(U, u, nRows, r0, r1);
      for (int r = r0; r < r1; r++) {
        if (R.kind[r] == ROW_OFF) continue;
        const int fl = R.flags[r];
        if (fl & 1) atomicMin(&S.claim[R.bi[r]], key);
        if (fl & 2) atomicMin(&S.claim[R.bj[r]], key);
        if (U.colored) break;  // all rows of a manifold share the same two bodies
      }
    }
    grid_barrier(S.bar, epoch);
    for (int a = tid; a < nAct; a += nth) {
      const int u = __ldcg(&act[a]);
      const unsigned pri = U.colored ? (unsigned)u * 2654435761u : (unsigned)u;
      const unsigned long long key = hi | pri;
      int r0, r1;
      unit_rows(U, u, nRows, r0, r1);
      bool win = true;
      for (int r = r0; r < r1; r++) {
        if (R.kind[r] == ROW_OFF) continue;
        const int fl = R.flags[r];
        if ((fl & 1) && __ldcg(&S.claim[R.bi[r]]) != key) win = false;
        if ((fl & 2) && __ldcg(&S.claim[R.bj[r]]) != key) win = false;
        if (U.colored) break;
      }
      if (win) {
        S.unitLevel[u] = round;
        S.order[atomicAdd(S.cursor, 1)] = u;
      } else {
        nxt[atomicAdd(&S.actCount[cur ^ 1], 1)] = u;
      }
    }
    grid_barrier(S.bar, epoch);
    if (tid == 0) { S.levelStart[round + 1] = *(volatile int*)S.cursor; S.actCount[cur] = 0; }
    round++;
    cur ^= 1;
    grid_barrier(S.bar, epoch);
  }
  if (tid == 0) *S.nLevels = round;
}

struct GsStats {
  double* worldTot;     // per world: sum |delta lambda| of the current iteration
  int* worldDone;       // per world: 1 once the tolerance test passed (gs_solver.dart:105)
  int* worldIters;      // per world: value of `iter` when its loop ended
  int* itersDone;       // max over worlds
};

// one GS update of row r (gs_solver.dart:80-102 + equation_class.dart:95-105,151-169)
__device__ __forceinline__ double gs_row(const RowArrays& R, const BodyArrays& B, int r) {
  const int kind = R.kind[r];
  const int bi = R.bi[r], bj = R.bj[r];
  const int fl = R.flags[r];
  const f3 n = ld3(R.n[r]), rA = ld3(R.rA[r]), rB = ld3(R.rB[r]);
  f3 sA;
  if (kind == ROW_ROT || kind == ROW_MOTOR) { sA.x = sA.y = sA.z = 0.f; } else sA = vneg(n);
  f3 vA = ld3(__ldcg(&B.vlam[bi])), wA = ld3(__ldcg(&B.wlam[bi]));
  f3 vB = ld3(__ldcg(&B.vlam[bj])), wB = ld3(__ldcg(&B.wlam[bj]));
  const double gwlambda = (vdot(vA, sA) + vdot(wA, rA)) + (vdot(vB, n) + vdot(wB, rB));
  const double lambdaj = R.lambda[r];
  double dl = R.invC[r] * (R.B[r] - gwlambda - R.eps[r] * lambdaj);
  const double minF = R.minF[r], maxF = R.maxF[r];
  if (lambdaj + dl < minF) dl = minF - lambdaj;
  else if (lambdaj + dl > maxF) dl = maxF - lambdaj;
  R.lambda[r] = lambdaj + dl;
  if (fl & 1) {
    vA = vaddscaled(vA, R.imA[r] * dl, sA);
    wA = vaddscaled(wA, dl, ld3(R.iA[r]));
    B.vlam[bi] = st3(vA);
    B.wlam[bi] = st3(wA);
  }
  if (fl & 2) {
    vB = vaddscaled(vB, R.imB[r] * dl, n);
    wB = vaddscaled(wB, dl, ld3(R.iB[r]));
    B.vlam[bj] = st3(vB);
    B.wlam[bj] = st3(wB);
  }
  return dl > 0.0 ? dl : -dl;
}

__global__ void __launch_bounds__(256) k_gs(RowArrays R, BodyArrays B, SchedArrays S, UnitMap U, SolveParams P, GsStats G) {
  __shared__ double s_red[8];
  unsigned epoch = 0;
  const int nRows = min(*R.nRows, R.rowCap);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  if (nRows == 0) { if (tid == 0) *G.itersDone = 0; return; }
  const int nLevels = *S.nLevels;
  const bool batch = P.nWorlds > 1;
  int iter = 0;
  for (; iter != P.maxIter; iter++) {
    double local = 0.0;
    for (int lvl = 0; lvl < nLevels; lvl++) {
      const int a0 = S.levelStart[lvl], a1 = S.levelStart[lvl + 1];
      for (int a = a0 + tid; a < a1; a += nth) {
        const int u = S.order[a];
        int r0, r1;
        unit_rows(U, u, nRows, r0, r1);
        for (int r = r0; r < r1; r++) {
          if (R.kind[r] == ROW_OFF) continue;
          if (batch) {
            const int w = B.world[R.bi[r]];
            if (__ldcg(&G.worldDone[w])) continue;
            atomicAdd(&G.worldTot[w], gs_row(R, B, r));
          } else {
            local += gs_row(R, B, r);
          }
        }
      }
      grid_barrier(S.bar, epoch);
    }
    // tolerance test (gs_solver.dart:105): the sum is order-insensitive for the comparison against tol^2
    bool allDone;
    if (!batch) {
      for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
      if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local;
      __syncthreads();
      if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) t += s_red[k];
        atomicAdd(&G.worldTot[0], t);
      }
      grid_barrier(S.bar, epoch);
      const double tot = __ldcg(&G.worldTot[0]);
      allDone = tot * tot < P.tol2;
      grid_barrier(S.bar, epoch);
      if (tid == 0) G.worldTot[0] = 0.0;
    } else {
      __shared__ int s_any;
      if (threadIdx.x == 0) s_any = 0;
      __syncthreads();
      int anyLive = 0;
      for (int w = tid; w < P.nWorlds; w += nth) {
        if (G.worldDone[w]) continue;
        const double tot = __ldcg(&G.worldTot[w]);
        if (tot * tot < P.tol2) { G.worldDone[w] = 1; G.worldIters[w] = iter; }
        else { anyLive = 1; G.worldTot[w] = 0.0; }
      }
      if (anyLive) atomicOr(&s_any, 1);
      __syncthreads();
      if (threadIdx.x == 0 && s_any) atomicOr(&G.worldDone[P.nWorlds], 1);  // slot nWorlds: "some world still iterating"
      grid_barrier(S.bar, epoch);
      allDone = __ldcg(&G.worldDone[P.nWorlds]) == 0;
      grid_barrier(S.bar, epoch);
      if (tid == 0) G.worldDone[P.nWorlds] = 0;
    }
    if (allDone) break;
  }
  if (tid == 0) *G.itersDone = iter;
}
