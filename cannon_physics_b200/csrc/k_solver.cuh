// k_solver.cuh — equation rows (K4), dependency-level scheduling and the Gauss-Seidel sweeps (K5).
//
// Reference: GSSolver.solve (lib/solver/gs_solver.dart:27-133) iterates over the equations in insertion
// order: all FrictionEquations, then all ContactEquations, then the constraints' equations
// (lib/world/world_class.dart:539-541,562,627-635). Gauss-Seidel is order dependent, so:
//
//   REFERENCE_ORDER  units = single rows in the reference's order, executed by *dependency levels*: a row gets
//                    level 1 + max(level of the previous row touching either of its movable bodies). Rows of
//                    one level touch disjoint movable bodies, so running them concurrently performs exactly
//                    the same floating-point operations on exactly the same operands as the sequential
//                    loop: results are bit-identical to the reference order while still parallel.
//   COLORED          the same machinery with units = contact manifolds (all contacts of one resolver task,
//                    rows [f1,f2,n] per contact) and hashed priorities, i.e. a randomised greedy colouring
//                    with O(max degree) colours (statistical agreement only).
//
// Pipeline: k_units_build (who touches which bodies) -> k_schedule (levels, execution order) ->
// scan(rows per unit in execution order) -> k_rows_build (rows written *in execution order*, so every sweep
// streams them with coalesced 128-bit loads) -> k_gs (persistent cooperative kernel: one thread per unit keeps
// the two bodies' lambda vectors in registers across the unit's rows; grid barrier between levels).
#pragma once
#include "k_narrowphase.cuh"
#include "world.cuh"

enum { ROW_CONTACT = 0, ROW_FRICTION = 1, ROW_ROT = 2, ROW_MOTOR = 3 };
// Constraint.update() flavours of a joint equation
enum { JM_P2P = 0, JM_HINGE_ROT = 1, JM_MOTOR = 2, JM_DISTANCE = 3, JM_DIRECT_ROT = 4 };
enum { SRC_NORMAL = 0, SRC_FRIC1 = 1, SRC_FRIC2 = 2, SRC_JOINT = 3, SRC_TASK = 4, SRC_JOINTS = 5 };

// rows in execution order (SoA; float4 vectors so a row is five 128-bit + six 64-bit transactions)
struct RowArrays {
  int* nRows;        // device count
  int* kind;
  float4 *n;         // spatial Jacobian of body j (sB); body i gets -n (zero for rotational rows)
  float4 *rA, *rB;   // rotational Jacobians
  float4 *iA, *iB;   // invInertiaWorldSolve * rA / rB, rounded to float like the reference's temp vector
  double *B, *invC, *eps, *minF, *maxF, *lambda;
  int rowCap;
  // COLORED (throughput) mode: the same row packed into one 80-byte record + one float = 84 B, solved in f32 with FMA.
  // The row order already differs from the reference there, so only statistical agreement is claimed and the
  // f64 emulation of the Dart VM buys nothing; B / invC are still evaluated in f64 and rounded once.
  int fast;
  float4* rec;     // 5 float4 per row, contiguous (80 B): (n,B) (rA,invC) (rB,eps) (iA,minF) (iB,maxF)
  float* flambda;
  // COLORED (exact) single-world mode: rows packed for k_gs_exact - the f32-stored Jacobian exactly as the reference keeps
  // it, the f64 scalars as f64 - in BLOCKS of 32 row slots, chunk-interleaved (see GxRow / gx_store_row): a window of the
  // sweep (32 units, one per lane) owns consecutive blocks, block r holds row r of every unit of the window, so a warp's
  // row step reads 3 KB of contiguous memory with fully coalesced 16-byte accesses. `lambda`, `minF`, `maxF` use the same
  // slot index (block * 32 + lane).
  float4* xblk;
};

// bound codes of a packed exact row: [0, bound] (contact), [-bound, bound] (friction, joints), [-bound, 0] (cone / twist),
// anything else keeps its bounds in RowArrays.minF / maxF
enum { GXB_POS = 0, GXB_SYM = 1, GXB_NEG = 2, GXB_GENERAL = 3 };
struct __align__(16) GxRow {
  float nx, ny, nz; int code;      // code bit 0: rotational row (sA = 0 instead of -n); bits 2-3: GXB_*
  float rAx, rAy, rAz, iAx;
  float rBx, rBy, rBz, iAy;
  float iBx, iBy, iBz, iAz;
  double B, invC, eps, bound;
};
static_assert(sizeof(GxRow) == 96, "a GxRow is six 16-byte chunks");
#define GX_CHUNKS 6

// units: by unit id (u*) before scheduling, by execution position (e*) after
struct UnitArrays {
  int* nUnits;       // device count
  int* nExec;        // device count of scheduled units (units without rows never enter the execution order)
  int *uBi, *uBj, *uFlags, *uRows, *uSrc;   // flags bit0/1: body i/j movable
  int* uKey;                                 // COLORED: canonical unit key (first contact index / nContacts + joint slot)
  int* uPri;                                 // COLORED: the key counted inside the unit's world (== uKey for a single world): what the colouring hashes
  int *eBi, *eBj, *eFlags, *eRowBase;        // eRowBase has nUnits+1 entries (exclusive scan of rows in exec order)
  double *eImA, *eImB;                       // invMassSolve of the two bodies
  int* eRows;                                // rows per unit in exec order (scan input)
  int* unitRow;                              // unit id -> first execution row (debug / multipliers)
  int unitCap;
  struct GsUnitRec* rec;                     // COLORED_F32 mode: the execution record the staged sweep copies to shared memory
  int* eLevel;                               // colour of each execution position (per-world sweep of a batch)
  struct GxUnit* xrec;                       // COLORED (exact) mode: execution record of k_gs_exact
  const int* unitSeq;                        // per unit id: rank of the unit among the units of body i / body j (colour order)
  const int* bodyCnt;                        // per body: number of scheduled units that move it
  // exact staged sweep: windows of 32 units per colour; winBase[w] = first row slot of window w (multiple of 32),
  // lvlWin[l] = first window of colour l
  const int* winBase;
  const int* lvlWin;
  const int* levelStart;
  const int* padTotal;  // row slots in use including the padding of the windows
};

// one 64-byte record per unit in execution order (exact staged sweep). seqA / degA: the unit is the seqA-th of the degA
// units that move body i (in colour order), so in iteration `it` it may run once done[bi] == it * degA + seqA.
struct __align__(16) GxUnit {
  int bi, bj, fl, r0;
  int r1, seqA, degA, seqB;
  int degB, pad0, pad1, pad2;
  double imA, imB;
};
static_assert(sizeof(GxUnit) == 64, "GxUnit is bulk-copied as 64-byte records");

// one 32-byte record per unit in execution order (COLORED mode), bulk-copied to shared memory by k_gs_fast
struct __align__(16) GsUnitRec {
  int bi, bj, fl, r0;
  int r1;
  float imA, imB;
  int grp;  // world id in a batch (early-exit group), 0 otherwise, -1 = no movable body
};

// warp tasks of the staged sweep: task t owns units [tab[t].x, tab[t+1].x) whose rows [tab[t].y, tab[t+1].y) are contiguous
struct GsTasks {
  int2* tab;
  int* lvlTask;  // [nLevels + 1] first task of each colour
  int* lvlWin;   // [nLevels] rows per task window of the colour
  int* nTasks;
  int taskCap;
  int winMin, winMax;  // rows per task window (GS_WIN_* for the f32 sweep, GX_WIN_* for the exact one)
};

struct JointArrays {  // one entry per constraint equation (P2P: 3, hinge: 6), uploaded at set_constraints
  int n;
  const int *bodyA, *bodyB, *kind, *enabled, *rowSlot;  // rowSlot: index among accepted joint rows or -1
  const int *slotEq;                                    // accepted slot -> equation index
  const float4 *pivotA, *pivotB, *axisA, *axisB;        // local frame
  const float4 *ni;                                     // P2P rows: world x / y / z
  const double *minF, *maxF, *a, *b, *eps, *targetVel;
  const int *first;                                     // index of the constraint's first equation (for hinge tangents)
  int nAccepted;
  const int* mode;      // JM_*: how Constraint.update() fills the equation
  const double* cosv;   // rotational rows: cos(maxAngle) / cos(ConeEquation.angle), evaluated on the host with libm
  const double* param;  // distance rows: DistanceConstraint.distance
};

struct SolveParams {
  double dt;
  double tol2;
  int maxIter;
  int nBodies;
  int nWorlds;
  int colored;
  int debugSkipWork;  // profiling aid (CANNON_DEBUG_SKIP_GS_WORK=1): run only the barrier skeleton of the sweeps
  long long* trace;   // profiling aid (CANNON_GS_TRACE=<file>): per CTA and phase, cycles of work and of barrier wait
};

// rigid_body.dart:303-314
__device__ __forceinline__ bool body_frozen(const BodyArrays& B, int b) { return B.sleep[b] == CANNON_SLEEPING || B.type[b] == CANNON_BODY_KINEMATIC; }

// a body whose solve mass and solve inertia are all zero never changes its vlambda/wlambda: no dependency
__device__ __forceinline__ bool body_movable(const BodyArrays& B, int b) {
  if (body_frozen(B, b)) return false;
  if (B.invMass[b] != 0.0) return true;
  const float4 r0 = B.iiw0[b], r1 = B.iiw1[b], r2 = B.iiw2[b];
  return r0.x != 0.f || r0.y != 0.f || r0.z != 0.f || r1.x != 0.f || r1.y != 0.f || r1.z != 0.f || r2.x != 0.f || r2.y != 0.f || r2.z != 0.f;
}

struct RowBody {
  f3 pos, vel, angvel, force, torque;
  float4 r0, r1, r2;  // invInertiaWorldSolve rows
  double im;          // invMassSolve
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
}

// computeGiMf (equation_class.dart:108-127), computeC (:130-148,172-174) and row store
__device__ inline void finish_row(const RowArrays& R, int row, int kind, const RowBody& A, const RowBody& Bd, const f3& sA, const f3& rA,
                                  const f3& sB, const f3& rB, double gterm /* -g*a or 0 */, double gw, double b, double eps, double minF,
                                  double maxF, double h) {
  const f3 iMfi = vscale(A.im, A.force), iMfj = vscale(Bd.im, Bd.force);
  const f3 iTi = mrow_mul(A.r0, A.r1, A.r2, A.torque), iTj = mrow_mul(Bd.r0, Bd.r1, Bd.r2, Bd.torque);
  const double giMf = (vdot(iMfi, sA) + vdot(iTi, rA)) + (vdot(iMfj, sB) + vdot(iTj, rB));
  const double Bv = gterm - gw * b - h * giMf;
  const f3 iA = mrow_mul(A.r0, A.r1, A.r2, rA), iB = mrow_mul(Bd.r0, Bd.r1, Bd.r2, rB);
  double c = A.im + Bd.im;
  c += vdot(iA, rA);
  c += vdot(iB, rB);
  c += eps;
  if (R.fast) {
    float4* q = R.rec + (size_t)row * 5;
    q[0] = st3(sB, (float)Bv); q[1] = st3(rA, (float)(1.0 / c)); q[2] = st3(rB, (float)eps);
    q[3] = st3(iA, (float)minF); q[4] = st3(iB, (float)maxF); R.flambda[row] = 0.f;
    return;
  }
  if (R.xblk) {
    GxRow q;
    q.nx = sB.x; q.ny = sB.y; q.nz = sB.z;
    q.rAx = rA.x; q.rAy = rA.y; q.rAz = rA.z; q.rBx = rB.x; q.rBy = rB.y; q.rBz = rB.z;
    q.iAx = iA.x; q.iAy = iA.y; q.iAz = iA.z; q.iBx = iB.x; q.iBy = iB.y; q.iBz = iB.z;
    q.B = Bv; q.invC = 1.0 / c; q.eps = eps;
    int bc = GXB_GENERAL;
    q.bound = maxF;
    const long long mn = __double_as_longlong(minF), mx = __double_as_longlong(maxF);
    if (mn == 0LL) bc = GXB_POS;
    else if (mn == __double_as_longlong(-maxF)) bc = GXB_SYM;
    else if (mx == 0LL) { bc = GXB_NEG; q.bound = -minF; }
    else { R.minF[row] = minF; R.maxF[row] = maxF; }
    q.code = ((kind == ROW_ROT || kind == ROW_MOTOR) ? 1 : 0) | (bc << 2);
    // slot `row` = block * 32 + lane; chunk c of the block sits at float4 index (block * 6 + c) * 32 + lane
    float4* d = R.xblk + (size_t)(row >> 5) * (GX_CHUNKS * 32) + (row & 31);
    const float4* src = (const float4*)&q;
#pragma unroll
    for (int c = 0; c < GX_CHUNKS; c++) d[c * 32] = src[c];
    R.lambda[row] = 0.0;
    return;
  }
  R.kind[row] = kind;
  R.n[row] = st3(sB); R.rA[row] = st3(rA); R.rB[row] = st3(rB); R.iA[row] = st3(iA); R.iB[row] = st3(iB);
  R.B[row] = Bv; R.invC[row] = 1.0 / c; R.eps[row] = eps; R.minF[row] = minF; R.maxF[row] = maxF; R.lambda[row] = 0.0;
}

// per contact: acceptance flags (Solver.addEquation filter, solver.dart:30-34) + wake-up flags (world_class.dart:564-590)
__global__ void __launch_bounds__(256) k_contact_flags(BodyArrays B, ContactArrays C, int contactCap, int* __restrict__ fricFlag,
                                                       int* __restrict__ contFlag, const double* __restrict__ matRestitution) {
  const int nc = min(*C.nContacts, contactCap);
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
    const int bi = C.bi[c], bj = C.bj[c];
    if (matRestitution) {  // worlds with shape materials: World.internalStep overrides the restitution by the BODY materials (world_class.dart:556-560)
      const int ma = B.material[bi], mb = B.material[bj];
      if (ma >= 0 && mb >= 0 && matRestitution[ma] >= 0 && matRestitution[mb] >= 0) C.rest[c] = matRestitution[ma] * matRestitution[mb];
    }
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

struct UnitSrc {  // what the units are made from
  int colored;
  int split;   // SplitSolver order: descending creation id (joints first created, then per contact: contact, friction 1, friction 2)
  const int* fricFlag; const int* contFlag;
  const int* fricOff; const int* contOff;   // exclusive scans over contacts
  const int* fricTotal; const int* contTotal;
  const int* taskOff; const int* taskCnt; const int* nTasks; int taskCap;   // resolver tasks = manifolds
  int contactCap;
};

__device__ __forceinline__ void put_unit(const UnitArrays& U, const BodyArrays& B, int u, int bi, int bj, int rows, int src, int nWorlds,
                                         int* __restrict__ worldRows, int key = 0, int priKey = 0) {
  U.uBi[u] = bi; U.uBj[u] = bj; U.uRows[u] = rows; U.uSrc[u] = src; U.uKey[u] = key; U.uPri[u] = priKey;
  U.uFlags[u] = rows > 0 ? ((body_movable(B, bi) ? 1 : 0) | (body_movable(B, bj) ? 2 : 0)) : 0;
  if (rows > 0) {
    // worldRows[w] > 0 <=> world w has equations this step: one plain store per (warp, world) instead of millions of
    // atomics on the same word
    const int wIdx = nWorlds > 1 ? B.world[bi] : 0;
    const unsigned peers = __match_any_sync(__activemask(), wIdx);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) worldRows[wIdx] = 1;
  }
}

// COLORED batches: the colouring hashes a unit's key, and a world's colours must not depend on which other worlds share
// the handle (a batch may be sharded over GPUs in any way). So the hashed key is counted inside the world:
// contacts from the world's first ContactEquation, constraints from the world's contact count. wk[w] = first contact
// index, wk[nW + w] = contacts, wk[2 nW + w] = first accepted joint slot of world w. For the per-world colouring
// (k_schedule_worlds) also the world's ranges of unit ids: wk[3 nW + w] / wk[4 nW + w] = first / last contact (a contact
// unit's id is the index of its first contact, k_units_build), wk[5 nW + w] = last accepted joint slot.
#define WK_ARRAYS 6
__global__ void __launch_bounds__(256) k_world_keys_init(int* __restrict__ wk, int nWorlds) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < WK_ARRAYS * nWorlds; i += gridDim.x * blockDim.x) {
    const int a = i / nWorlds;
    wk[i] = (a == 0 || a == 2 || a == 3) ? 0x7fffffff : (a == 1 ? 0 : -1);
  }
}
__global__ void __launch_bounds__(256) k_world_keys(BodyArrays B, ContactArrays C, UnitSrc S, JointArrays J, int nWorlds, int* __restrict__ wk) {
  const int nc = min(*C.nContacts, S.contactCap);
  const int nt = min(*S.nTasks, S.taskCap);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  // neighbouring tasks belong to the same world: one atomic per (warp, world) instead of one per task
  for (int t0 = tid - (int)(threadIdx.x & 31); t0 < nt; t0 += nth) {
    const int t = t0 + (int)(threadIdx.x & 31);
    int w = -1, c0 = 0x7fffffff, m = 0;
    if (t < nt) {
      m = S.taskCnt[t];
      const int c = S.taskOff[t];
      if (m > 0 && c + m <= nc) { w = B.world[C.bi[c]]; c0 = c; } else m = 0;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, w);
    const int mn = __reduce_min_sync(peers, c0), sum = __reduce_add_sync(peers, m);
    const int tmn = mn, tmx = __reduce_max_sync(peers, w >= 0 ? c0 + m - 1 : -1);  // the world's contacts are contiguous
    if (w >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) {
      atomicMin(&wk[w], mn);
      atomicAdd(&wk[nWorlds + w], sum);
      atomicMin(&wk[3 * nWorlds + w], tmn);
      atomicMax(&wk[4 * nWorlds + w], tmx);
    }
  }
  // neighbouring joint slots belong to the same world too: one atomic pair per (warp, world)
  for (int s0 = tid - (int)(threadIdx.x & 31); s0 < J.nAccepted; s0 += nth) {
    const int s2 = s0 + (int)(threadIdx.x & 31);
    const int w = s2 < J.nAccepted ? B.world[J.bodyA[J.slotEq[s2]]] : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, w);
    const int smn = __reduce_min_sync(peers, s2), smx = __reduce_max_sync(peers, s2);
    if (w >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) {
      atomicMin(&wk[2 * nWorlds + w], smn);
      atomicMax(&wk[5 * nWorlds + w], smx);
    }
  }
}

// unit table. Reference order: unit id == reference row index [2*fricRank | nF2 + contRank | joints].
// Coloured: a manifold's contacts are cut into runs of CANNON_COLORED_UNIT_CONTACTS; unit id == index of the run's first
// contact (the ids in between stay empty), followed by the joint rows.
__global__ void __launch_bounds__(256) k_units_build(BodyArrays B, ContactArrays C, UnitSrc S, JointArrays J, UnitArrays U, int nWorlds,
                                                     int* __restrict__ worldRows, int* __restrict__ unitOverflow, const int* __restrict__ wk) {
  const int nc = min(*C.nContacts, S.contactCap);
  const int nF2 = 2 * (*S.fricTotal), nC = *S.contTotal;
  const int nt = min(*S.nTasks, S.taskCap);
  const int jointBase = S.colored ? nc : nF2 + nC;
  const int nUnits = jointBase + J.nAccepted;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  if (tid == 0) { *U.nUnits = nUnits; if (nUnits > U.unitCap) atomicMax(unitOverflow, nUnits); }
  if (nUnits > U.unitCap) return;
  if (S.colored) {
    for (int c0 = tid; c0 < nc; c0 += nth) {
      const int t = C.task[c0];
      const int first = S.taskOff[t], end = min(first + S.taskCnt[t], nc);
      int rows = 0, bi = 0, bj = 0;
      if ((c0 - first) % CANNON_COLORED_UNIT_CONTACTS == 0) {  // head of a run
        const int c1 = min(c0 + CANNON_COLORED_UNIT_CONTACTS, end);
        bi = C.bi[c0]; bj = C.bj[c0];
        for (int c = c0; c < c1; c++) rows += S.contFlag[c] + 2 * S.fricFlag[c];
      }
      put_unit(U, B, c0, bi, bj, rows, c0 * 8 + SRC_TASK, nWorlds, worldRows, c0, (nWorlds > 1 && rows > 0) ? c0 - wk[B.world[bi]] : c0);
    }
  } else if (S.split) {
    // unit id = position in descending creation-id order (split_solver.dart:108,167-169)
    const int last = nUnits - 1;
    for (int c = tid; c < nc; c += nth) {
      const int bi = C.bi[c], bj = C.bj[c];
      const int asc = J.nAccepted + S.contOff[c] + 2 * S.fricOff[c];
      const int cf = S.contFlag[c];
      if (cf) put_unit(U, B, last - asc, bi, bj, 1, c * 8 + SRC_NORMAL, nWorlds, worldRows);
      if (S.fricFlag[c]) {
        put_unit(U, B, last - (asc + cf), bi, bj, 1, c * 8 + SRC_FRIC1, nWorlds, worldRows);
        put_unit(U, B, last - (asc + cf + 1), bi, bj, 1, c * 8 + SRC_FRIC2, nWorlds, worldRows);
      }
    }
    for (int s = tid; s < J.nAccepted; s += nth) {
      const int e = J.slotEq[s];
      put_unit(U, B, last - s, J.bodyA[e], J.bodyB[e], 1, e * 8 + SRC_JOINT, nWorlds, worldRows);
    }
    return;
  } else {
    for (int c = tid; c < nc; c += nth) {
      const int bi = C.bi[c], bj = C.bj[c];
      if (S.fricFlag[c]) {
        const int f0 = S.fricOff[c];
        put_unit(U, B, 2 * f0, bi, bj, 1, c * 8 + SRC_FRIC1, nWorlds, worldRows);
        put_unit(U, B, 2 * f0 + 1, bi, bj, 1, c * 8 + SRC_FRIC2, nWorlds, worldRows);
      }
      if (S.contFlag[c]) put_unit(U, B, nF2 + S.contOff[c], bi, bj, 1, c * 8 + SRC_NORMAL, nWorlds, worldRows);
    }
  }
  for (int s = tid; s < J.nAccepted; s += nth) {
    const int e = J.slotEq[s];
    if (S.colored) {
      // all accepted equations of one constraint (3 for a point-to-point joint, 5-6 for a hinge ...) act on the same two
      // bodies: one unit at the first of them, swept row after row like a contact manifold - a joint costs one colour
      // instead of one per equation. The other slots keep their unit ids but own no rows (they leave the schedule).
      const int f = J.first[e];
      const bool head = s == 0 || J.first[J.slotEq[s - 1]] != f;
      int rows = 0;
      if (head) { rows = 1; while (s + rows < J.nAccepted && J.first[J.slotEq[s + rows]] == f) rows++; }
      int pk = nc + s;
      if (nWorlds > 1) { const int w = B.world[J.bodyA[e]]; pk = wk[nWorlds + w] + (s - wk[2 * nWorlds + w]); }
      put_unit(U, B, jointBase + s, J.bodyA[e], J.bodyB[e], rows, s * 8 + SRC_JOINTS, nWorlds, worldRows, nc + s, pk);
    } else {
      put_unit(U, B, jointBase + s, J.bodyA[e], J.bodyB[e], 1, e * 8 + SRC_JOINT, nWorlds, worldRows);
    }
  }
}

// Islands of SplitSolver (split_solver.dart:76-117): connected components of the non-static bodies under the
// accepted equations; label = smallest body index of the component (min-label propagation + pointer jumping).
__global__ void __launch_bounds__(256) k_islands(BodyArrays B, UnitArrays U, int nBodies, int* __restrict__ label, int* __restrict__ changed,
                                                 int* __restrict__ nIslands, unsigned* bar);

// ---- grid barrier -----------------------------------------------------------------------------------
// Barrier for cooperative (co-resident) launches; `bar` (64 words, zeroed before the launch) holds a monotonic
// arrival counter in word 0 and the published generation in word 32 (its own 128-byte line). Only the arrivals
// touch the counter (one atomic per CTA); the last arriver publishes the generation and everybody else polls that
// second line with a short sleep, so the pollers do not fight the atomics for the same L2 line.
// how many CTAs of a cooperative launch take part: enough for `items` units of work per phase at `perCta` each, rounded
// up to a power of two, never more than were launched. The rest return at once, so a small world pays for a small
// barrier although the launch (and a captured graph of it) always has the full co-resident grid.
__device__ __forceinline__ int coop_ctas(long long items, int perCta) {
  long long b = (items + perCta - 1) / perCta;
  long long q = 1;
  while (q < b) q <<= 1;
  return (int)(q < (long long)gridDim.x ? q : (long long)gridDim.x);
}

// Monotonic-counter grid barrier: one release atomic to arrive, acquire loads of the same counter to wait (no separate
// generation word, no stand-alone fences: the release / acquire pair at gpu scope is cumulative over the CTA's writes
// ordered before it by the block barrier). The counter is cleared by the host before every cooperative launch.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& epoch, int nCtas) {
  __syncthreads();
  if (nCtas == 1) return;  // single-CTA launches (small worlds) only need the block barrier
  if (threadIdx.x == 0) {
    epoch += 1;
    const unsigned target = epoch * (unsigned)nCtas;
    unsigned v;
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(v) : "l"(bar) : "memory");
    v += 1u;
    while (v < target) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) k_islands(BodyArrays B, UnitArrays U, int nBodies, int* __restrict__ label, int* __restrict__ changed,
                                                 int* __restrict__ nIslands, unsigned* bar) {
  unsigned epoch = 0;
  const int nUnits = min(*U.nUnits, U.unitCap);
  const int nCtas = coop_ctas(max(nUnits / 2 + 1, nBodies / 4 + 1), 256);
  if ((int)blockIdx.x >= nCtas) return;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = nCtas * blockDim.x;
  for (int b = tid; b < nBodies; b += nth) label[b] = B.type[b] == CANNON_BODY_STATIC ? -1 : b;
  if (tid == 0) { changed[0] = 0; changed[1] = 0; *nIslands = 0; }
  grid_barrier(bar, epoch, nCtas);
  for (int round = 0;; round++) {
    int* flag = &changed[round & 1];
    for (int u = tid; u < nUnits; u += nth) {
      if (U.uRows[u] == 0) continue;
      const int bi = U.uBi[u], bj = U.uBj[u];
      const int la = __ldcg(&label[bi]), lb = __ldcg(&label[bj]);
      if (la < 0 || lb < 0 || la == lb) continue;  // static bodies do not merge islands
      const int m = min(la, lb);
      // hook the larger root label onto the smaller one
      atomicMin(&label[max(la, lb)], m);
      atomicMin(&label[bi], m);
      atomicMin(&label[bj], m);
      *flag = 1;
    }
    grid_barrier(bar, epoch, nCtas);
    for (int b = tid; b < nBodies; b += nth) {  // pointer jumping
      int l = __ldcg(&label[b]);
      if (l < 0) continue;
      int r = __ldcg(&label[l]);
      while (r != l) { l = r; r = __ldcg(&label[l]); }
      label[b] = l;
    }
    if (tid == 0) changed[(round + 1) & 1] = 0;
    grid_barrier(bar, epoch, nCtas);
    if (__ldcg(flag) == 0) break;
  }
  for (int b = tid; b < nBodies; b += nth)
    if (label[b] == b) atomicAdd(nIslands, 1);
}

struct SchedArrays {
  unsigned long long* claim;  // per body
  int* unitLevel;             // per unit, -1 = unassigned
  int* order;                 // execution position -> unit id
  int* levelStart;            // [maxLevels+1] execution positions
  int* nLevels;
  int *act0, *act1;           // active (unassigned) unit lists
  int* actCount;              // [2]
  int* cursor;                // append cursor into order
  unsigned* bar;
  int maxLevels;
  int* levelOverflow;
  int* unitSeq;   // exact staged sweep: [2 * unit] rank of the unit among the units moving body i / body j, or null
  int* bodyCnt;   // per body: units scheduled so far that move it
};

// Dependency levels by repeated "claim the bodies with the smallest pending priority": a unit is released in
// the round in which it holds the minimum on all of its movable bodies, which is exactly
// level(u) = 1 + max(level of earlier units sharing a movable body). Keys carry the round in the high bits so
// stale claims of earlier rounds always lose the atomicMin (no clearing pass).
__global__ void __launch_bounds__(256) k_schedule(UnitArrays U, SchedArrays S, int colored) {
  unsigned epoch = 0;
  const int nUnits = min(*U.nUnits, U.unitCap);
  const int nCtas = coop_ctas(nUnits / 2 + 1, 256);
  if ((int)blockIdx.x >= nCtas) return;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = nCtas * blockDim.x;
  for (int u = tid; u < nUnits; u += nth) { S.act0[u] = u; S.unitLevel[u] = -1; }
  if (tid == 0) { S.actCount[0] = nUnits; S.actCount[1] = 0; *S.cursor = 0; S.levelStart[0] = 0; }
  grid_barrier(S.bar, epoch, nCtas);
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
      const unsigned pri = colored ? (unsigned)U.uPri[u] * 2654435761u : (unsigned)u;
      const unsigned long long key = hi | pri;
      const int fl = U.uFlags[u];
      if (fl & 1) atomicMin(&S.claim[U.uBi[u]], key);
      if (fl & 2) atomicMin(&S.claim[U.uBj[u]], key);
    }
    grid_barrier(S.bar, epoch, nCtas);
    // winners go to the execution order, losers to the next round's list: one cursor atomic per warp and list
    for (int a0 = tid - (int)(threadIdx.x & 31); a0 < nAct; a0 += nth) {
      const int a = a0 + (int)(threadIdx.x & 31);
      const bool active = a < nAct;
      int u = 0;
      bool win = false, emit = false;
      if (active) {
        u = __ldcg(&act[a]);
        const unsigned pri = colored ? (unsigned)U.uPri[u] * 2654435761u : (unsigned)u;
        const unsigned long long key = hi | pri;
        const int fl = U.uFlags[u];
        win = true;
        if ((fl & 1) && __ldcg(&S.claim[U.uBi[u]]) != key) win = false;
        if ((fl & 2) && __ldcg(&S.claim[U.uBj[u]]) != key) win = false;
        if (win) {
          S.unitLevel[u] = round;
          // a unit without rows (a resolver task that produced no contact) is done here: it never enters the
          // execution order, so the sweeps do not have to step over it in every iteration
          emit = U.uRows[u] > 0;
          // at most one unit wins a body per round, so the rank of a unit on its body follows the colour order
          if (emit && S.unitSeq) {
            S.unitSeq[2 * u] = (fl & 1) ? atomicAdd(&S.bodyCnt[U.uBi[u]], 1) : 0;
            S.unitSeq[2 * u + 1] = (fl & 2) ? atomicAdd(&S.bodyCnt[U.uBj[u]], 1) : 0;
          }
        }
      }
      const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
      const unsigned mw = __ballot_sync(0xffffffffu, emit), ml = __ballot_sync(0xffffffffu, active && !win);
      int bw = 0, bl = 0;
      if ((threadIdx.x & 31) == 0) {
        if (mw) bw = atomicAdd(S.cursor, __popc(mw));
        if (ml) bl = atomicAdd(&S.actCount[cur ^ 1], __popc(ml));
      }
      bw = __shfl_sync(0xffffffffu, bw, 0);
      bl = __shfl_sync(0xffffffffu, bl, 0);
      if (emit) S.order[bw + __popc(mw & lt)] = u;
      else if (active && !win) nxt[bl + __popc(ml & lt)] = u;
    }
    grid_barrier(S.bar, epoch, nCtas);
    if (tid == 0) { S.levelStart[round + 1] = *(volatile int*)S.cursor; S.actCount[cur] = 0; }
    round++;
    cur ^= 1;
    grid_barrier(S.bar, epoch, nCtas);
  }
  if (tid == 0) { *S.nLevels = round; *U.nExec = *(volatile int*)S.cursor; }
}

// rows per unit in execution order (input of the row-base scan) + per-unit execution data
// The colouring of a COLORED batch, one warp per world (no grid barrier: worlds share no body). Same rule as k_schedule -
// colour(u) = round in which u holds the smallest pending priority on all of its movable bodies - over the world's own
// units: the contact runs with ids in [wk[3 nW + w], wk[4 nW + w]] and the joint units of slots [wk[2 nW + w], wk[5 nW + w]].
// The claims live in shared memory and are cleared every round; a round visits all candidates and skips the coloured ones
// (~100 units x ~13 rounds per world). Scheduled units are appended to `order` (k_world_count / k_world_fill regroup them
// by world and colour anyway); nLevels is the maximum over the worlds.
#define SW_REG 8     // units a lane keeps in registers (256 per world; larger worlds re-read theirs from global memory)
#define SW_MAXB 512  // bodies of a world (the host picks this kernel only for batches of worlds up to this size)
__global__ void __launch_bounds__(32) k_schedule_worlds(BodyArrays B, UnitArrays U, SchedArrays S, const int* __restrict__ wk, const int* __restrict__ worldBody,
                                                        int nWorlds, int jointBase0 /* unused */, const int* __restrict__ nTasksPtr, int taskCap) {
  __shared__ unsigned s_claim[SW_MAXB];
  const int lane = threadIdx.x, wd = blockIdx.x;
  const int b0 = worldBody[wd], nB = worldBody[wd + 1] - b0;
  const int jointBase = min(*nTasksPtr, taskCap);  // unit id of joint slot 0 = number of contacts (k_units_build)
  const int t0 = wk[3 * nWorlds + wd], t1 = wk[4 * nWorlds + wd];
  const int j0 = wk[2 * nWorlds + wd], j1 = wk[5 * nWorlds + wd];
  const int nT = t1 >= t0 ? t1 - t0 + 1 : 0, nJ = j1 >= j0 ? j1 - j0 + 1 : 0;
  const int nCand = nT + nJ;
  if (nCand == 0) return;
  auto unit_of = [&](int k) { return k < nT ? t0 + k : jointBase + j0 + (k - nT); };
  int round = 0;
  // the candidates that own rows, compacted in candidate order (most joint slots and a third of the tasks own none)
  __shared__ int s_u[32 * SW_REG];
  int nAct = 0;
  for (int kb = 0; kb < nCand; kb += 32) {
    const int k = kb + lane;
    const bool has = k < nCand && U.uRows[unit_of(k)] > 0;
    const unsigned bits = __ballot_sync(0xffffffffu, has);
    if (has) { const int pos = nAct + __popc(bits & ((1u << lane) - 1u)); if (pos < 32 * SW_REG) s_u[pos] = unit_of(k); }
    nAct += __popc(bits);
  }
  __syncwarp();
  if (nAct <= 32 * SW_REG) {
    // the usual case: every lane keeps its units (priority, body slots, level) in registers for all rounds
    unsigned pri[SW_REG];
    int ba[SW_REG], bb[SW_REG], lv[SW_REG];  // body slots (-1: not movable / no unit), level (-1 pending, -2 none)
#pragma unroll
    for (int i = 0; i < SW_REG; i++) {
      const int k = lane + 32 * i;
      pri[i] = 0u; ba[i] = bb[i] = -1; lv[i] = -2;
      if (k < nAct) {
        const int u = s_u[k];
        const int fl = U.uFlags[u];
        pri[i] = (unsigned)U.uPri[u] * 2654435761u;
        ba[i] = (fl & 1) ? U.uBi[u] - b0 : -1;
        bb[i] = (fl & 2) ? U.uBj[u] - b0 : -1;
        lv[i] = -1;
      }
    }
    while (true) {
      if (round >= S.maxLevels) { if (lane == 0) *S.levelOverflow = 1; break; }
      for (int i = lane; i < nB; i += 32) s_claim[i] = 0xffffffffu;
      __syncwarp();
      bool pending = false;
#pragma unroll
      for (int i = 0; i < SW_REG; i++)
        if (lv[i] == -1) {
          pending = true;
          if (ba[i] >= 0) atomicMin(&s_claim[ba[i]], pri[i]);
          if (bb[i] >= 0) atomicMin(&s_claim[bb[i]], pri[i]);
        }
      if (!__any_sync(0xffffffffu, pending)) break;
      __syncwarp();
#pragma unroll
      for (int i = 0; i < SW_REG; i++)
        if (lv[i] == -1 && (ba[i] < 0 || s_claim[ba[i]] == pri[i]) && (bb[i] < 0 || s_claim[bb[i]] == pri[i])) lv[i] = round;
      round++;
      __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < SW_REG; i++) {
      const int k = lane + 32 * i;
      if (k < nAct && lv[i] >= 0) S.unitLevel[s_u[k]] = lv[i];
    }
  } else {
  for (int k = lane; k < nCand; k += 32) S.unitLevel[unit_of(k)] = -1;
  __syncwarp();
  while (true) {
    if (round >= S.maxLevels) { if (lane == 0) *S.levelOverflow = 1; break; }
    for (int i = lane; i < nB; i += 32) s_claim[i] = 0xffffffffu;
    __syncwarp();
    int pending = 0;
    for (int k = lane; k < nCand; k += 32) {
      const int u = unit_of(k);
      if (U.uRows[u] <= 0 || S.unitLevel[u] >= 0) continue;
      pending++;
      const unsigned pri = (unsigned)U.uPri[u] * 2654435761u;
      const int fl = U.uFlags[u];
      if (fl & 1) atomicMin(&s_claim[U.uBi[u] - b0], pri);
      if (fl & 2) atomicMin(&s_claim[U.uBj[u] - b0], pri);
    }
    if (!__any_sync(0xffffffffu, pending > 0)) break;
    __syncwarp();
    for (int k = lane; k < nCand; k += 32) {
      const int u = unit_of(k);
      if (U.uRows[u] <= 0 || S.unitLevel[u] >= 0) continue;
      const unsigned pri = (unsigned)U.uPri[u] * 2654435761u;
      const int fl = U.uFlags[u];
      bool win = true;
      if ((fl & 1) && s_claim[U.uBi[u] - b0] != pri) win = false;
      if ((fl & 2) && s_claim[U.uBj[u] - b0] != pri) win = false;
      if (win) S.unitLevel[u] = round;
    }
    round++;
    __syncwarp();
  }
  }
  // the world's scheduled units, in candidate order, at a block of `order` reserved with one atomic
  int mine = 0;
  for (int k = lane; k < nCand; k += 32) mine += (U.uRows[unit_of(k)] > 0) ? 1 : 0;
  int incl = mine;
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  int base = 0;
  if (lane == 0 && total > 0) base = atomicAdd(U.nExec, total);
  base = __shfl_sync(0xffffffffu, base, 0) + incl - mine;
  for (int k = lane; k < nCand; k += 32) {
    const int u = unit_of(k);
    if (U.uRows[u] > 0) S.order[base++] = u;
  }
  if (lane == 0) atomicMax(S.nLevels, round);
}

__global__ void __launch_bounds__(256) k_exec_units(BodyArrays B, UnitArrays U, const int* __restrict__ order, const int* __restrict__ unitLevel) {
  const int nUnits = min(*U.nExec, U.unitCap);
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < nUnits; a += gridDim.x * blockDim.x) {
    const int u = order[a];
    const int bi = U.uBi[u], bj = U.uBj[u];
    U.eRows[a] = U.uRows[u];
    U.eLevel[a] = unitLevel[u];
    U.eBi[a] = bi; U.eBj[a] = bj; U.eFlags[a] = U.uFlags[u];
    U.eImA[a] = body_frozen(B, bi) ? 0.0 : B.invMass[bi];
    U.eImB[a] = body_frozen(B, bj) ? 0.0 : B.invMass[bj];
  }
}

// ContactEquation.computeB (contact_equation.dart:34-77)
__device__ __forceinline__ void build_normal_row(const RowArrays& R, int row, const RowBody& A, const RowBody& Bd, const f3& ri, const f3& rj,
                                                 const f3& ni, double restitution, double a, double b, double eps, double minF, double maxF, double h) {
  const f3 rixn = vcross(ri, ni), rjxn = vcross(rj, ni);
  f3 pen = vadd(Bd.pos, rj);
  pen = vsub(pen, A.pos);
  pen = vsub(pen, ri);
  const double g = vdot(ni, pen);
  const double ePlusOne = restitution + 1;
  const double gw = ePlusOne * vdot(Bd.vel, ni) - ePlusOne * vdot(A.vel, ni) + vdot(Bd.angvel, rjxn) - vdot(A.angvel, rixn);
  finish_row(R, row, ROW_CONTACT, A, Bd, vneg(ni), vneg(rixn), ni, rjxn, -g * a, gw, b, eps, minF, maxF, h);
}
// FrictionEquation.computeB (friction_equation.dart:19-47)
__device__ __forceinline__ void build_friction_row(const RowArrays& R, int row, const RowBody& A, const RowBody& Bd, const f3& ri, const f3& rj,
                                                   const f3& t, double b, double eps, double slip, double h) {
  const f3 rixt = vcross(ri, t), rjxt = vcross(rj, t);
  const f3 sA = vneg(t), rA = vneg(rixt);
  const double gw = (vdot(A.vel, sA) + vdot(A.angvel, rA)) + (vdot(Bd.vel, t) + vdot(Bd.angvel, rjxt));
  finish_row(R, row, ROW_FRICTION, A, Bd, sA, rA, t, rjxt, 0.0, gw, b, eps, -slip, slip, h);
}

// Constraint.update() + joint equation rows (point_to_point_constraint.dart:68-83, hinge_constraint.dart:79-104,
// distance_constraint.dart:27-38, lock_constraint.dart:70-88, cone_twist_constraint.dart:77-96, rotational_equation.dart:34-58,
// cone_equation.dart:34-56, rotational_motor_equation.dart:17-33)
__device__ inline void build_joint_row(const RowArrays& R, int row, const BodyArrays& B, const JointArrays& J, int e, const RowBody& A,
                                       const RowBody& Bd, double h) {
  const int bi = J.bodyA[e], bj = J.bodyB[e];
  const q4 qa = ldq(B.quat[bi]), qb = ldq(B.quat[bj]);
  const int kind = J.kind[e];
  f3 zero; zero.x = zero.y = zero.z = 0.f;
  const int mode = J.mode[e];
  if (kind == ROW_CONTACT && mode == JM_DISTANCE) {
    // DistanceConstraint.update (distance_constraint.dart:27-38)
    const double halfDist = J.param[e] * 0.5;
    f3 normal = vsub(Bd.pos, A.pos);
    vnormalize(normal);
    const f3 ri = vscale(halfDist, normal), rj = vscale(-halfDist, normal);
    build_normal_row(R, row, A, Bd, ri, rj, normal, 0.0, J.a[e], J.b[e], J.eps[e], J.minF[e], J.maxF[e], h);
  } else if (kind == ROW_CONTACT) {
    const f3 ri = qrot(qa, ld3(J.pivotA[e])), rj = qrot(qb, ld3(J.pivotB[e]));
    build_normal_row(R, row, A, Bd, ri, rj, ld3(J.ni[e]), 0.0, J.a[e], J.b[e], J.eps[e], J.minF[e], J.maxF[e], h);
  } else if (kind == ROW_ROT) {
    f3 axA, worldAxisB;
    if (mode == JM_DIRECT_ROT) {
      // lock_constraint.dart:79-87 / cone_twist_constraint.dart:84-94: body-local axes of the equation into the world frame
      axA = qrot(qa, ld3(J.axisA[e]));
      worldAxisB = qrot(qb, ld3(J.axisB[e]));
    } else {
      const f3 worldAxisA = qrot(qa, ld3(J.axisA[e]));
      worldAxisB = qrot(qb, ld3(J.axisB[e]));
      f3 t1, t2;
      vtangents(worldAxisA, t1, t2);
      axA = (e - J.first[e] == 3) ? t1 : t2;  // rotationalEquation1 / rotationalEquation2
    }
    const f3 nixnj = vcross(axA, worldAxisB), njxni = vcross(worldAxisB, axA);
    const double g = J.cosv[e] - vdot(axA, worldAxisB);
    const double gw = (vdot(A.vel, zero) + vdot(A.angvel, njxni)) + (vdot(Bd.vel, zero) + vdot(Bd.angvel, nixnj));
    finish_row(R, row, ROW_ROT, A, Bd, zero, njxni, zero, nixnj, -g * J.a[e], gw, J.b[e], J.eps[e], J.minF[e], J.maxF[e], h);
  } else {
    const f3 axA = qrot(qa, ld3(J.axisA[e])), axB = qrot(qb, ld3(J.axisB[e]));
    const f3 rB = vneg(axB);
    const double gw = ((vdot(A.vel, zero) + vdot(A.angvel, axA)) + (vdot(Bd.vel, zero) + vdot(Bd.angvel, rB))) - J.targetVel[e];
    finish_row(R, row, ROW_MOTOR, A, Bd, zero, axA, zero, rB, 0.0, gw, J.b[e], J.eps[e], J.minF[e], J.maxF[e], h);
  }
}

// rows, written at their execution positions: one thread per unit in execution order
__global__ void __launch_bounds__(128) k_rows_build(BodyArrays B, ContactArrays C, UnitSrc S, JointArrays J, UnitArrays U, RowArrays R,
                                                    SolveParams P, const int* __restrict__ order, int* __restrict__ rowOverflow,
                                                    const int* __restrict__ bodyGroup, int nGroups) {
  const int nUnits = min(*U.nExec, U.unitCap);
  const int nRows = U.eRowBase[nUnits];
  const int nSlots = R.xblk ? max(nRows, *U.padTotal) : nRows;
  if (blockIdx.x == 0 && threadIdx.x == 0) { *R.nRows = nRows; if (nSlots > R.rowCap) atomicMax(rowOverflow, nSlots); }
  if (nSlots > R.rowCap) return;
  const double h = P.dt;
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < nUnits; a += gridDim.x * blockDim.x) {
    const int u = order[a];
    int row = U.eRowBase[a];
    int rs = 1;  // distance between consecutive rows of a unit
    if (R.xblk) {
      // window-interleaved slots: unit = lane (a - a0) % 32 of window (a - a0) / 32 of its colour, row r in block r of the window
      const int l = U.eLevel[a], a0 = U.levelStart[l];
      row = U.winBase[U.lvlWin[l] + ((a - a0) >> 5)] + ((a - a0) & 31);
      rs = 32;
    }
    U.unitRow[u] = row;
    if (R.fast) {
      GsUnitRec rec;
      rec.bi = U.eBi[a]; rec.bj = U.eBj[a]; rec.fl = U.eFlags[a]; rec.r0 = row; rec.r1 = U.eRowBase[a + 1];
      rec.imA = (float)U.eImA[a]; rec.imB = (float)U.eImB[a];
      rec.grp = 0;
      if (nGroups > 1) { rec.grp = bodyGroup[rec.bi]; if (rec.grp < 0) rec.grp = bodyGroup[rec.bj]; }
      U.rec[a] = rec;
    }
    if (U.xrec) {
      GxUnit x;
      x.bi = U.eBi[a]; x.bj = U.eBj[a]; x.fl = U.eFlags[a]; x.r0 = row; x.r1 = row + 32 * U.eRows[a];  // row r of the unit = slot r0 + 32 r
      x.seqA = x.seqB = x.degA = x.degB = 0;
      if (U.unitSeq) {  // dataflow bookkeeping (see the history note in k_gs_exact.cuh): rank of the unit on its two bodies
        x.seqA = U.unitSeq[2 * u]; x.seqB = U.unitSeq[2 * u + 1];
        x.degA = (x.fl & 1) ? U.bodyCnt[x.bi] : 0; x.degB = (x.fl & 2) ? U.bodyCnt[x.bj] : 0;
      }
      x.pad0 = x.pad1 = x.pad2 = 0;
      x.imA = U.eImA[a]; x.imB = U.eImB[a];
      U.xrec[a] = x;
    }
    if (U.eRows[a] == 0) continue;
    const int src = U.uSrc[u], kind = src & 7, idx = src >> 3;
    RowBody A, Bd;
    load_row_body(B, U.eBi[a], A);
    load_row_body(B, U.eBj[a], Bd);
    if (kind == SRC_JOINT) {
      build_joint_row(R, row, B, J, idx, A, Bd, h);
    } else if (kind == SRC_JOINTS) {  // COLORED: the accepted equations of one constraint, slots idx .. idx + rows - 1
      for (int k = 0; k < U.eRows[a]; k++) build_joint_row(R, row + k * rs, B, J, J.slotEq[idx + k], A, Bd, h);
    } else if (kind == SRC_TASK) {
      const int tk = C.task[idx];  // idx = first contact of the run
      const int c0 = idx, c1 = min(c0 + CANNON_COLORED_UNIT_CONTACTS, S.taskOff[tk] + S.taskCnt[tk]);
      for (int c = c0; c < c1; c++) {
        const f3 ri = ld3(C.ri[c]), rj = ld3(C.rj[c]), ni = ld3(C.ni[c]);
        if (S.fricFlag[c]) {
          f3 t1, t2;
          vtangents(ni, t1, t2);
          build_friction_row(R, row, A, Bd, ri, rj, t1, C.fb[c], C.feps[c], C.slip[c], h);
          build_friction_row(R, row + rs, A, Bd, ri, rj, t2, C.fb[c], C.feps[c], C.slip[c], h);
          row += 2 * rs;
        }
        if (S.contFlag[c]) {
          C.row[c] = row;
          build_normal_row(R, row, A, Bd, ri, rj, ni, C.rest[c], C.ca[c], C.cb[c], C.ceps[c], 0.0, 1e6, h);
          row += rs;
        }
      }
    } else {
      const int c = idx;
      const f3 ri = ld3(C.ri[c]), rj = ld3(C.rj[c]), ni = ld3(C.ni[c]);
      if (kind == SRC_NORMAL) {
        C.row[c] = row;
        build_normal_row(R, row, A, Bd, ri, rj, ni, C.rest[c], C.ca[c], C.cb[c], C.ceps[c], 0.0, 1e6, h);
      } else {
        f3 t1, t2;
        vtangents(ni, t1, t2);
        build_friction_row(R, row, A, Bd, ri, rj, kind == SRC_FRIC1 ? t1 : t2, C.fb[c], C.feps[c], C.slip[c], h);
      }
    }
  }
}

// exact staged sweep: rows per window = 32 x the longest unit of the window (units of a colour are sorted by row count, so
// the padding is small); one warp per window
__global__ void __launch_bounds__(256) k_gx_windows(UnitArrays U, const int* __restrict__ levelStart, const int* __restrict__ nLevels,
                                                    const int* __restrict__ lvlWin, int* __restrict__ winRows, int winCap, int* __restrict__ overflow) {
  const int nl = *nLevels, nWin = lvlWin[nl];
  if (nWin > winCap) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(overflow, nWin); return; }
  const int lane = threadIdx.x & 31;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < winCap; w += (gridDim.x * blockDim.x) >> 5) {
    int m = 0;
    if (w < nWin) {
      int lo = 0, hi = nl;  // last colour whose first window is <= w
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (lvlWin[mid] <= w) lo = mid; else hi = mid;
      }
      const int a = levelStart[lo] + 32 * (w - lvlWin[lo]) + lane;
      if (a < levelStart[lo + 1]) m = U.eRows[a];
      for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    if (lane == 0) winRows[w] = 32 * m;  // windows beyond nWin contribute nothing to the scan
  }
}

struct GsStats {
  const int* bodyGroup; // grouped early exit: world id (batches) or island label (SplitSolver); -1 = static body
  int nGroups;          // <= 1: one group (plain GSSolver)
  double* worldTot;     // per world: sum |delta lambda| of the current iteration
  int* worldDone;       // per world: 1 once the tolerance test passed (gs_solver.dart:105)
  int* worldIters;      // per world: value of `iter` when its loop ended
  int* itersDone;       // max over worlds
};

struct RowData {
  float4 n, rA, rB, iA, iB;
  double B, invC, eps, minF, maxF, lambda;
  int kind;
};
__device__ __forceinline__ void load_row(const RowArrays& R, int r, RowData& d) {
  d.n = R.n[r]; d.rA = R.rA[r]; d.rB = R.rB[r]; d.iA = R.iA[r]; d.iB = R.iB[r];
  d.B = R.B[r]; d.invC = R.invC[r]; d.eps = R.eps[r]; d.minF = R.minF[r]; d.maxF = R.maxF[r]; d.lambda = R.lambda[r];
  d.kind = R.kind[r];
}

// persistent cooperative kernel: all iterations x levels of GSSolver.solve (gs_solver.dart:76-108)
__global__ void __launch_bounds__(256) k_gs(RowArrays R, BodyArrays B, UnitArrays U, SchedArrays S, SolveParams P, GsStats G) {
  __shared__ double s_red[8];
  __shared__ int s_any;
  unsigned epoch = 0;
  const int nRows = min(*R.nRows, R.rowCap);
  const int nLevels = *S.nLevels;
  // CTAs for about twice the mean colour width (one unit per thread)
  const int nCtas = coop_ctas(2LL * S.levelStart[nLevels] / max(nLevels, 1) + 1, 256);
  if ((int)blockIdx.x >= nCtas) return;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = nCtas * blockDim.x;
  // units of a level are dealt to the CTAs warp by warp (32 consecutive units per warp, consecutive warps on different
  // CTAs), so a narrow level still spreads over every SM instead of filling the first few CTAs
  const int itid = (((threadIdx.x >> 5) * nCtas + blockIdx.x) << 5) | (threadIdx.x & 31);
  if (nRows == 0) { if (tid == 0) *G.itersDone = 0; return; }
  const bool batch = G.nGroups > 1;
  int iter = 0;
  for (; iter != P.maxIter; iter++) {
    double local = 0.0;
    for (int lvl = 0; lvl < nLevels; lvl++) {
      const int a0 = S.levelStart[lvl], a1 = S.levelStart[lvl + 1];
      for (int a = a0 + itid; a < a1; a += nth) {
        const int r0 = U.eRowBase[a], r1 = U.eRowBase[a + 1];
        if (r0 == r1) continue;
        const int bi = U.eBi[a], bj = U.eBj[a], fl = U.eFlags[a];
        int w = 0;
        if (batch) {
          w = G.bodyGroup[bi];
          if (w < 0) w = G.bodyGroup[bj];
          if (w < 0 || __ldcg(&G.worldDone[w])) continue;
        }
        const double imA = U.eImA[a], imB = U.eImB[a];
        // the two bodies' lambda vectors stay in registers across the unit's rows (gs_solver.dart:88-102,
        // equation_class.dart:95-105,151-169); every update rounds to float exactly like the reference's stores
        // a body that is not movable keeps vlambda = wlambda = 0 for the whole solve: do not fetch it (thousands of
        // units resting on the same static body would otherwise all hit one L2 sector)
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        f3 vA = ld3((fl & 1) ? __ldcg(&B.vlam[2 * bi]) : z4), wA = ld3((fl & 1) ? __ldcg(&B.wlam[2 * bi]) : z4);
        f3 vB = ld3((fl & 2) ? __ldcg(&B.vlam[2 * bj]) : z4), wB = ld3((fl & 2) ? __ldcg(&B.wlam[2 * bj]) : z4);
        double acc = 0.0;
        RowData d, nx;
        load_row(R, r0, d);
        nx = d;
        for (int r = r0; r < r1; r++) {
          if (r + 1 < r1) load_row(R, r + 1, nx);  // software prefetch of the next row
          const f3 n = ld3(d.n), rA = ld3(d.rA), rB = ld3(d.rB);
          f3 sA;
          if (d.kind == ROW_ROT || d.kind == ROW_MOTOR) { sA.x = sA.y = sA.z = 0.f; } else sA = vneg(n);
          const double gwlambda = (vdot(vA, sA) + vdot(wA, rA)) + (vdot(vB, n) + vdot(wB, rB));
          double dl = d.invC * (d.B - gwlambda - d.eps * d.lambda);
          if (d.lambda + dl < d.minF) dl = d.minF - d.lambda;
          else if (d.lambda + dl > d.maxF) dl = d.maxF - d.lambda;
          R.lambda[r] = d.lambda + dl;
          if (fl & 1) {
            vA = vaddscaled(vA, imA * dl, sA);
            wA = vaddscaled(wA, dl, ld3(d.iA));
          }
          if (fl & 2) {
            vB = vaddscaled(vB, imB * dl, n);
            wB = vaddscaled(wB, dl, ld3(d.iB));
          }
          acc += dl > 0.0 ? dl : -dl;
          d = nx;
        }
        if (fl & 1) { B.vlam[2 * bi] = st3(vA); B.wlam[2 * bi] = st3(wA); }
        if (fl & 2) { B.vlam[2 * bj] = st3(vB); B.wlam[2 * bj] = st3(wB); }
        if (batch) atomicAdd(&G.worldTot[w], acc);
        else local += acc;
      }
      grid_barrier(S.bar, epoch, nCtas);
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
      grid_barrier(S.bar, epoch, nCtas);
      const double tot = __ldcg(&G.worldTot[0]);
      allDone = tot * tot < P.tol2;
      grid_barrier(S.bar, epoch, nCtas);
      if (tid == 0) G.worldTot[0] = 0.0;
    } else {
      if (threadIdx.x == 0) s_any = 0;
      __syncthreads();
      int anyLive = 0;
      for (int w = tid; w < G.nGroups; w += nth) {
        if (G.worldDone[w]) continue;
        const double tot = __ldcg(&G.worldTot[w]);
        if (tot * tot < P.tol2) { G.worldDone[w] = 1; G.worldIters[w] = iter; }
        else { anyLive = 1; G.worldTot[w] = 0.0; }
      }
      if (anyLive) atomicOr(&s_any, 1);
      __syncthreads();
      if (threadIdx.x == 0 && s_any) atomicOr(&G.worldDone[G.nGroups], 1);  // slot nWorlds: "some world still iterating"
      grid_barrier(S.bar, epoch, nCtas);
      allDone = __ldcg(&G.worldDone[G.nGroups]) == 0;
      grid_barrier(S.bar, epoch, nCtas);
      if (tid == 0) G.worldDone[G.nGroups] = 0;
    }
    if (allDone) break;
  }
  if (tid == 0) *G.itersDone = iter;
}


// ---- COLORED mode sweep: f32 + FMA, one thread per manifold, lambdas in registers ---------------------------
#define GS_CHUNK 4
__device__ __forceinline__ float dot3f(const float4& a, float bx, float by, float bz) { return fmaf(a.z, bz, fmaf(a.y, by, a.x * bx)); }

#define GS_TRACE_PHASES 64
#define GS_TRACE_BEGIN() long long trS = 0, trW = 0; if (P.trace && threadIdx.x == 0) trS = clock64();
#define GS_TRACE_WORK() if (P.trace) { __syncthreads(); if (threadIdx.x == 0) trW = clock64(); }
#define GS_TRACE_END()                                                                          \
  if (P.trace && threadIdx.x == 0) {                                                            \
    const int ph = iter * nLevels + lvl;                                                        \
    if (ph < GS_TRACE_PHASES) {                                                                 \
      P.trace[(blockIdx.x * GS_TRACE_PHASES + ph) * 2] = trW - trS;                             \
      P.trace[(blockIdx.x * GS_TRACE_PHASES + ph) * 2 + 1] = clock64() - trW;                   \
    }                                                                                           \
  }

__global__ void __launch_bounds__(256, 2) k_gs_fast_v1(RowArrays R, BodyArrays B, UnitArrays U, SchedArrays S, SolveParams P, GsStats G) {
  __shared__ double s_red[32];
  __shared__ int s_any;
  unsigned epoch = 0;
  const int nRows = min(*R.nRows, R.rowCap);
  const int nLevels = *S.nLevels;
  // CTAs for about twice the mean colour width (one unit per thread)
  const int nCtas = coop_ctas(2LL * S.levelStart[nLevels] / max(nLevels, 1) + 1, 256);
  if ((int)blockIdx.x >= nCtas) return;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = nCtas * blockDim.x;
  // units of a level are dealt to the CTAs warp by warp (32 consecutive units per warp, consecutive warps on different
  // CTAs), so a narrow level still spreads over every SM instead of filling the first few CTAs
  const int itid = (((threadIdx.x >> 5) * nCtas + blockIdx.x) << 5) | (threadIdx.x & 31);
  if (nRows == 0) { if (tid == 0) *G.itersDone = 0; return; }
  const bool batch = G.nGroups > 1;
  int iter = 0;
  for (; iter != P.maxIter; iter++) {
    double local = 0.0;
    for (int lvl = 0; lvl < nLevels; lvl++) {
      GS_TRACE_BEGIN();
      const int a0 = S.levelStart[lvl], a1 = S.levelStart[lvl + 1];
      for (int a = a0 + itid; a < (P.debugSkipWork ? a0 : a1); a += nth) {
        const int r0 = U.eRowBase[a], r1 = U.eRowBase[a + 1];
        if (r0 == r1) continue;
        const int bi = U.eBi[a], bj = U.eBj[a], fl = U.eFlags[a];
        int w = 0;
        if (batch) {
          w = G.bodyGroup[bi];
          if (w < 0) w = G.bodyGroup[bj];
          if (w < 0 || __ldcg(&G.worldDone[w])) continue;
        }
        const float imA = (float)U.eImA[a], imB = (float)U.eImB[a];
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 vA = (fl & 1) ? __ldcg(&B.vlam[2 * bi]) : z4, wA = (fl & 1) ? __ldcg(&B.wlam[2 * bi]) : z4;
        float4 vB = (fl & 2) ? __ldcg(&B.vlam[2 * bj]) : z4, wB = (fl & 2) ? __ldcg(&B.wlam[2 * bj]) : z4;
        float acc = 0.f;
        // rows are fetched in chunks of GS_CHUNK through the read-only path before any of them is solved: the
        // manifold's sequential chain then pays one memory round trip per chunk instead of one per row
        for (int rc = r0; rc < r1; rc += GS_CHUNK) {
          float4 c0[GS_CHUNK], c1[GS_CHUNK], c2[GS_CHUNK], c3[GS_CHUNK], c4[GS_CHUNK];
          float cl[GS_CHUNK];
#pragma unroll
          for (int k = 0; k < GS_CHUNK; k++) {
            const int r = min(rc + k, r1 - 1);
            const float4* q = R.rec + (size_t)r * 5;
            c0[k] = __ldg(q); c1[k] = __ldg(q + 1); c2[k] = __ldg(q + 2); c3[k] = __ldg(q + 3); c4[k] = __ldg(q + 4);
            cl[k] = R.flambda[r];
          }
#pragma unroll
          for (int k = 0; k < GS_CHUNK; k++) {
            const int r = rc + k;
            if (r >= r1) break;
            const float4 q0 = c0[k], q1 = c1[k], q2 = c2[k], q3 = c3[k], q4 = c4[k];
            const float lam = cl[k];
            // G*W_lambda with G = [-n, rA, n, rB]
            float gw = dot3f(q0, vB.x - vA.x, vB.y - vA.y, vB.z - vA.z);
            gw += dot3f(q1, wA.x, wA.y, wA.z);
            gw += dot3f(q2, wB.x, wB.y, wB.z);
            float dl = q1.w * (q0.w - gw - q2.w * lam);
            if (lam + dl < q3.w) dl = q3.w - lam;
            else if (lam + dl > q4.w) dl = q4.w - lam;
            cl[k] = lam + dl;
            if (fl & 1) {
              const float s = -imA * dl;
              vA.x = fmaf(s, q0.x, vA.x); vA.y = fmaf(s, q0.y, vA.y); vA.z = fmaf(s, q0.z, vA.z);
              wA.x = fmaf(dl, q3.x, wA.x); wA.y = fmaf(dl, q3.y, wA.y); wA.z = fmaf(dl, q3.z, wA.z);
            }
            if (fl & 2) {
              const float s = imB * dl;
              vB.x = fmaf(s, q0.x, vB.x); vB.y = fmaf(s, q0.y, vB.y); vB.z = fmaf(s, q0.z, vB.z);
              wB.x = fmaf(dl, q4.x, wB.x); wB.y = fmaf(dl, q4.y, wB.y); wB.z = fmaf(dl, q4.z, wB.z);
            }
            acc += fabsf(dl);
          }
#pragma unroll
          for (int k = 0; k < GS_CHUNK; k++)
            if (rc + k < r1) R.flambda[rc + k] = cl[k];
        }
        if (fl & 1) { B.vlam[2 * bi] = vA; B.wlam[2 * bi] = wA; }
        if (fl & 2) { B.vlam[2 * bj] = vB; B.wlam[2 * bj] = wB; }
        if (batch) atomicAdd(&G.worldTot[w], (double)acc);
        else local += (double)acc;
      }
      GS_TRACE_WORK();
      grid_barrier(S.bar, epoch, nCtas);
      GS_TRACE_END();
    }
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
      grid_barrier(S.bar, epoch, nCtas);
      const double tot = __ldcg(&G.worldTot[0]);
      allDone = tot * tot < P.tol2;
      grid_barrier(S.bar, epoch, nCtas);
      if (tid == 0) G.worldTot[0] = 0.0;
    } else {
      if (threadIdx.x == 0) s_any = 0;
      __syncthreads();
      int anyLive = 0;
      for (int w = tid; w < G.nGroups; w += nth) {
        if (G.worldDone[w]) continue;
        const double tot = __ldcg(&G.worldTot[w]);
        if (tot * tot < P.tol2) { G.worldDone[w] = 1; G.worldIters[w] = iter; }
        else { anyLive = 1; G.worldTot[w] = 0.0; }
      }
      if (anyLive) atomicOr(&s_any, 1);
      __syncthreads();
      if (threadIdx.x == 0 && s_any) atomicOr(&G.worldDone[G.nGroups], 1);
      grid_barrier(S.bar, epoch, nCtas);
      allDone = __ldcg(&G.worldDone[G.nGroups]) == 0;
      grid_barrier(S.bar, epoch, nCtas);
      if (tid == 0) G.worldDone[G.nGroups] = 0;
    }
    if (allDone && !P.debugSkipWork) break;
  }
  if (tid == 0) *G.itersDone = iter;
}

// ---- staged sweep: the colored solver's production kernel ---------------------------------------------------------
// k_gs_fast_v1 gives every thread one unit and lets it fetch its own rows: 32 lanes then read 32 different places of the
// row arrays (sector waste, one L1 wavefront per lane) and every phase pays the chain  barrier -> unit record -> rows
// (DRAM) -> solve -> store.  Here a WARP owns a window of consecutive rows (and the 24..40 units they belong to). Unit
// records and row records never change during a solve, so one lane copies both contiguous ranges global -> shared with
// cp.async.bulk (TMA, completion on an mbarrier) one task ahead of the warp, across colour and iteration boundaries;
// lambdas are prefetched and written back as coalesced lines. After a grid barrier only the body-lambda gather (L2)
// stands before the arithmetic, and DRAM sees purely sequential streams.
#define GS_WARPS 8
#define GS_THREADS (GS_WARPS * 32)
#define GS_CAP_ROWS 144
#define GS_CAP_UNITS 48
#define GS_WIN_MIN 96
#define GS_WIN_MAX 112
#define GS_LAM_REGS ((GS_CAP_ROWS + 31) / 32)
#define GS_LAM_OFF (GS_CAP_ROWS * 80 + GS_CAP_UNITS * 32)
#define GS_LAM_BYTES ((GS_CAP_ROWS + 8) * 4)  // the lambda range is copied from a 16-byte aligned start: up to 3 + 3 floats of slack
#define GS_BUF_BYTES (GS_LAM_OFF + GS_LAM_BYTES)
#define GS_WARP_BYTES (2 * GS_BUF_BYTES)
#define GS_SMEM_BYTES (GS_WARPS * GS_WARP_BYTES)
#define GS_LS_MAX 1024

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// 256-bit L2 access (LDG.E.ENL2.256 / STG.E.ENL2.256 on sm_100a): both halves of a 32-byte body record at once
__device__ __forceinline__ void ldcg_f8(const float4* p, float4& a, float4& b) {
  asm volatile("ld.global.cg.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p)
               : "memory");
}
__device__ __forceinline__ void st_f8(float4* p, const float4& a, const float4& b) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z),
               "f"(b.w)
               : "memory");
}
__device__ __forceinline__ float4 lds_f4(unsigned a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float lds_f1(unsigned a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f1(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(b))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  unsigned done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(smem_u32(b)), "r"(parity)
                 : "memory");
  } while (!done);
}

// task windows per colour: about 32 units per window, at most GS_WIN_MAX rows
__global__ void __launch_bounds__(256) k_gs_task_levels(UnitArrays U, SchedArrays S, GsTasks T, int* __restrict__ taskOverflow) {
  __shared__ int s_scan[256];
  __shared__ int s_base;
  const int nLevels = *S.nLevels;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int l0 = 0; l0 < nLevels; l0 += 256) {
    const int l = l0 + threadIdx.x;
    int n = 0;
    if (l < nLevels) {
      const int u0 = S.levelStart[l], u1 = S.levelStart[l + 1];
      const int rows = U.eRowBase[u1] - U.eRowBase[u0], units = u1 - u0;
      if (T.winMax == 0) {  // k_gs_exact: windows of 32 units, whatever their rows
        n = (units + 31) / 32;
      } else {
        int win = T.winMin;
        if (rows > 0 && units > 0) win = min(T.winMax, max(T.winMin, (int)((32LL * rows + units - 1) / units)));
        T.lvlWin[l] = win;
        n = (rows + win - 1) / win;
      }
    }
    s_scan[threadIdx.x] = n;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
      const int v = threadIdx.x >= o ? s_scan[threadIdx.x - o] : 0;
      __syncthreads();
      s_scan[threadIdx.x] += v;
      __syncthreads();
    }
    if (l < nLevels) T.lvlTask[l] = s_base + s_scan[threadIdx.x] - n;
    __syncthreads();
    if (threadIdx.x == 255) s_base += s_scan[255];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (T.winMax == 0) {  // k_gs_exact: lvlWin holds the first 8-unit task of each colour
      int q = 0;
      for (int l = 0; l < nLevels; l++) { T.lvlWin[l] = q; q += (S.levelStart[l + 1] - S.levelStart[l] + 7) / 8; }
      T.lvlWin[nLevels] = q;
    }
    T.lvlTask[nLevels] = s_base;
    *T.nTasks = s_base;
    if (s_base > T.taskCap) atomicMax(taskOverflow, s_base);
  }
}

__global__ void __launch_bounds__(256) k_gs_task_fill(UnitArrays U, SchedArrays S, GsTasks T) {
  const int nLevels = *S.nLevels;
  const int nTasks = min(*T.nTasks, T.taskCap);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t <= nTasks; t += gridDim.x * blockDim.x) {
    if (t == nTasks) {
      const int nu = S.levelStart[nLevels];
      T.tab[t] = make_int2(nu, U.eRowBase[nu]);
      continue;
    }
    int lo = 0, hi = nLevels;  // last colour whose first task is <= t
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (T.lvlTask[mid] <= t) lo = mid; else hi = mid;
    }
    const int u0 = S.levelStart[lo], u1 = S.levelStart[lo + 1];
    const int ws = U.eRowBase[u0] + (t - T.lvlTask[lo]) * T.lvlWin[lo];
    int a = u0, b = u1;  // first unit of the colour whose first row is >= ws
    while (a < b) {
      const int mid = (a + b) >> 1;
      if (U.eRowBase[mid] < ws) a = mid + 1; else b = mid;
    }
    T.tab[t] = make_int2(a, U.eRowBase[a]);
  }
}

struct GsTask { int a, lvl, it; };  // a: global task index, -1 = none

__device__ __forceinline__ void gs_seek(const int* lt, int nLevels, int maxIter, int gw, GsTask& t) {
  int hops = 0;
  while (t.a >= lt[t.lvl + 1]) {
    if (++hops > nLevels) { t.a = -1; return; }  // the warp owns no task in any colour
    if (++t.lvl == nLevels) { t.lvl = 0; if (++t.it >= maxIter) { t.a = -1; return; } }
    t.a = lt[t.lvl] + gw;
  }
}

__global__ void __launch_bounds__(GS_THREADS, 1) k_gs_fast(RowArrays R, BodyArrays B, UnitArrays U, SchedArrays S, GsTasks T, SolveParams P, GsStats G) {
  extern __shared__ __align__(128) unsigned char s_dyn[];
  __shared__ unsigned long long s_mbar[GS_WARPS][2];
  __shared__ double s_red[32];
  __shared__ int s_any;
  __shared__ int s_lt[GS_LS_MAX + 2];
  unsigned epoch = 0;
  const int nRows = min(*R.nRows, R.rowCap);
  const int nLevels = *S.nLevels;
  // CTAs for about twice the mean number of warp tasks per colour
  if (*T.nTasks > T.taskCap) return;  // reported through the row-overflow counter by k_gs_task_levels
  const int nCtas = coop_ctas(2LL * min(*T.nTasks, T.taskCap) / max(nLevels, 1) + 1, GS_WARPS);
  if ((int)blockIdx.x >= nCtas) return;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = nCtas * blockDim.x;
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  if (nRows == 0) { if (tid == 0) *G.itersDone = 0; return; }
  const bool batch = G.nGroups > 1;
  const int* lt = T.lvlTask;
  if (nLevels <= GS_LS_MAX) {
    for (int k = threadIdx.x; k <= nLevels; k += blockDim.x) s_lt[k] = T.lvlTask[k];
    lt = s_lt;
  }
  if (lane == 0) { mbar_init(&s_mbar[wic][0], 1); mbar_init(&s_mbar[wic][1], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  unsigned char* wbase = s_dyn + (size_t)wic * GS_WARP_BYTES;
  // tasks of a colour are dealt to the warps CTA-interleaved so a narrow colour still spreads over every SM
  const int gw = wic * nCtas + blockIdx.x, nW = nCtas * GS_WARPS;

  // three tasks in flight per warp: t0 being solved (rows in buffer `buf`), t1 rows in flight, t2 table entry in flight
  GsTask t0, t1, t2;
  int2 x0, y0, x1, y1, x2, y2;  // table entries [a], [a+1] of the three tasks
  auto table = [&](const GsTask& t, int2& x, int2& y) {
    x = make_int2(0, 0); y = x;
    if (t.a >= 0) { x = __ldg(&T.tab[t.a]); y = __ldg(&T.tab[t.a + 1]); }
  };
  auto next = [&](const GsTask& t) {
    GsTask n = t;
    if (n.a >= 0) { n.a += nW; gs_seek(lt, nLevels, P.maxIter, gw, n); }
    return n;
  };
  // unit records, rows and (withLam) the rows' lambdas of a task: three bulk copies on one mbarrier. The lambda range
  // starts at the 16-byte boundary below the first row (bulk copies need 16-byte alignment); no register ever waits
  // for it, so nothing of the prefetch can stall the warp before the data is used.
  auto issue = [&](const int2& x, const int2& y, int b, bool withLam) -> bool {  // returns whether a copy was started
    const int nUs = min(y.x - x.x, GS_CAP_UNITS), nRs = min(y.y - x.y, GS_CAP_ROWS);
    if (nUs <= 0) return false;
    if (lane == 0) {
      unsigned char* dst = wbase + (size_t)b * GS_BUF_BYTES;
      // the buffer was last written with ordinary shared stores (lambdas of an earlier task): order them before the
      // async-proxy writes of the bulk copies
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const int a0 = x.y & ~3;
      const unsigned lamBytes = (withLam && nRs > 0) ? (unsigned)(((x.y + nRs - a0 + 3) & ~3) * 4) : 0u;
      mbar_expect_tx(&s_mbar[wic][b], (unsigned)(nUs * 32 + nRs * 80) + lamBytes);
      bulk_g2s(dst + GS_CAP_ROWS * 80, U.rec + x.x, (unsigned)(nUs * 32), &s_mbar[wic][b]);
      if (nRs > 0) bulk_g2s(dst, R.rec + (size_t)x.y * 5, (unsigned)(nRs * 80), &s_mbar[wic][b]);
      if (lamBytes) bulk_g2s(dst + GS_LAM_OFF, R.flambda + a0, lamBytes, &s_mbar[wic][b]);
    }
    return true;
  };

  t0.a = lt[0] + gw; t0.lvl = 0; t0.it = 0;
  gs_seek(lt, nLevels, P.maxIter, gw, t0);
  t1 = next(t0);
  table(t0, x0, y0);
  table(t1, x1, y1);
  int buf = 0;
  unsigned parity = 0;  // bit b: phase parity of buffer b's mbarrier
  bool pend0 = issue(x0, y0, 0, true), pend1 = false;

  int iter = 0;
  int trN = 0;
  for (; iter != P.maxIter; iter++) {
    double local = 0.0;
    for (int lvl = 0; lvl < nLevels; lvl++) {
      GS_TRACE_BEGIN();
      while (t0.a >= 0 && t0.lvl == lvl && t0.it == iter) {
        // one task per warp and one colour: the next task is this one again, its lambdas are not written back yet
        // (they are handed over shared -> shared after the solve instead of being fetched)
        const bool sameNext = t1.a == t0.a;
        pend1 = issue(x1, y1, buf ^ 1, !sameNext);
        t2 = next(t1);
        table(t2, x2, y2);
        const int u0 = x0.x, nU = y0.x - x0.x, rBase = x0.y;
        const int nUs = min(nU, GS_CAP_UNITS), nRs = min(y0.y - x0.y, GS_CAP_ROWS);
        long long tk0 = 0, tk1 = 0, tk2 = 0, tkg = 0;
        const bool tr = P.trace && wic == 0 && iter == 1 && lvl == 0;
        if (tr) tk0 = clock64();
        if (pend0) { mbar_wait(&s_mbar[wic][buf], (parity >> buf) & 1u); parity ^= 1u << buf; }
        if (tr) tk1 = clock64();
        const unsigned char* sb = wbase + (size_t)buf * GS_BUF_BYTES;
        const float4* srows = (const float4*)sb;
        const GsUnitRec* sunits = (const GsUnitRec*)(sb + GS_CAP_ROWS * 80);
        float* const slam = (float*)(wbase + (size_t)buf * GS_BUF_BYTES + GS_LAM_OFF) + (rBase & 3);
        int flushEnd = nRs;
        if (!P.debugSkipWork) {
          for (int u = lane; u < nU; u += 32) {
            const GsUnitRec m = u < nUs ? sunits[u] : U.rec[u0 + u];
            if (m.r1 <= m.r0) continue;
            const bool staged = m.r1 - rBase <= nRs;
            if (!staged) flushEnd = min(flushEnd, m.r0 - rBase);  // that unit keeps its lambdas in global memory
            if (batch && (m.grp < 0 || __ldcg(&G.worldDone[m.grp]))) continue;
            // a body that is not movable keeps vlambda = wlambda = 0 for the whole solve: do not fetch it (thousands
            // of units resting on the same static body would otherwise all hit one L2 sector)
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            // vlambda and wlambda of a body are one 32-byte record: one 256-bit L2 access (one sector) per body
            float4 vA = z4, wA = z4, vB = z4, wB = z4;
            if (m.fl & 1) ldcg_f8(&B.vlam[2 * m.bi], vA, wA);
            if (m.fl & 2) ldcg_f8(&B.vlam[2 * m.bj], vB, wB);
            float acc = 0.f;
            if (tr && tkg == 0) {  // trace only: when did the body lambdas of this lane arrive?
              asm volatile("" ::"f"(vA.x), "f"(wA.x), "f"(vB.x), "f"(wB.x));
              tkg = clock64();
            }
            // one projected Gauss-Seidel row update: G*W_lambda with G = [-n, rA, n, rB], clamp, body deltas
#define GS_ROW_UPDATE(q0, q1, q2, q3, q4, lam, dl)                                                                   \
            {                                                                                                        \
              float gw_ = dot3f(q0, vB.x - vA.x, vB.y - vA.y, vB.z - vA.z);                                          \
              gw_ += dot3f(q1, wA.x, wA.y, wA.z);                                                                    \
              gw_ += dot3f(q2, wB.x, wB.y, wB.z);                                                                    \
              dl = q1.w * (q0.w - gw_ - q2.w * lam);                                                                 \
              if (lam + dl < q3.w) dl = q3.w - lam;                                                                  \
              else if (lam + dl > q4.w) dl = q4.w - lam;                                                             \
              if (m.fl & 1) {                                                                                        \
                const float s_ = -m.imA * dl;                                                                        \
                vA.x = fmaf(s_, q0.x, vA.x); vA.y = fmaf(s_, q0.y, vA.y); vA.z = fmaf(s_, q0.z, vA.z);               \
                wA.x = fmaf(dl, q3.x, wA.x); wA.y = fmaf(dl, q3.y, wA.y); wA.z = fmaf(dl, q3.z, wA.z);               \
              }                                                                                                      \
              if (m.fl & 2) {                                                                                        \
                const float s_ = m.imB * dl;                                                                         \
                vB.x = fmaf(s_, q0.x, vB.x); vB.y = fmaf(s_, q0.y, vB.y); vB.z = fmaf(s_, q0.z, vB.z);               \
                wB.x = fmaf(dl, q4.x, wB.x); wB.y = fmaf(dl, q4.y, wB.y); wB.z = fmaf(dl, q4.z, wB.z);               \
              }                                                                                                      \
              acc += fabsf(dl);                                                                                      \
            }
            if (staged) {
              // rows and lambdas of the window live in shared memory: explicit ld.shared / st.shared (a pointer that may
              // also be global compiles to generic loads) and the next row is fetched before the current one is solved
              unsigned qa = smem_u32(srows) + (unsigned)(m.r0 - rBase) * 80u, la = smem_u32(slam) + (unsigned)(m.r0 - rBase) * 4u;
              const int nr = m.r1 - m.r0;
              float4 n0 = lds_f4(qa), n1 = lds_f4(qa + 16), n2 = lds_f4(qa + 32), n3 = lds_f4(qa + 48), n4 = lds_f4(qa + 64);
              float nl = lds_f1(la);
              for (int r = 0; r < nr; r++) {
                const float4 q0 = n0, q1 = n1, q2 = n2, q3 = n3, q4 = n4;
                const float lam = nl;
                if (r + 1 < nr) {
                  qa += 80u;
                  n0 = lds_f4(qa); n1 = lds_f4(qa + 16); n2 = lds_f4(qa + 32); n3 = lds_f4(qa + 48); n4 = lds_f4(qa + 64);
                  nl = lds_f1(la + 4u);
                }
                float dl;
                GS_ROW_UPDATE(q0, q1, q2, q3, q4, lam, dl);
                sts_f1(la, lam + dl);
                la += 4u;
              }
            } else {
              const float4* q = R.rec + (size_t)m.r0 * 5;
              float* lp = R.flambda + m.r0;
              for (int r = m.r0; r < m.r1; r++, q += 5, lp++) {
                const float4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4];
                const float lam = *lp;
                float dl;
                GS_ROW_UPDATE(q0, q1, q2, q3, q4, lam, dl);
                *lp = lam + dl;
              }
            }
#undef GS_ROW_UPDATE
            if (m.fl & 1) st_f8(&B.vlam[2 * m.bi], vA, wA);
            if (m.fl & 2) st_f8(&B.vlam[2 * m.bj], vB, wB);
            if (batch) atomicAdd(&G.worldTot[m.grp], (double)acc);
            else local += (double)acc;
          }
        }
        if (tr) tk2 = clock64();
        flushEnd = __reduce_min_sync(0xffffffffu, flushEnd);
        // lambdas of the staged units back to global as full lines
#pragma unroll
        for (int j = 0; j < GS_LAM_REGS; j++) {
          const int idx = j * 32 + lane;
          if (idx < flushEnd) R.flambda[rBase + idx] = slam[idx];
        }
        asm volatile("fence.proxy.async.global;" ::: "memory");  // a later bulk copy (async proxy) reads these lambdas back
        if (sameNext) {
          float* const nlam = (float*)(wbase + (size_t)(buf ^ 1) * GS_BUF_BYTES + GS_LAM_OFF) + (rBase & 3);
#pragma unroll
          for (int j = 0; j < GS_LAM_REGS; j++) {
            const int idx = j * 32 + lane;
            if (idx < nRs) nlam[idx] = slam[idx];
          }
        }
        __syncwarp();
        if (tr && lane == 0 && trN < 4) {
          long long* o = P.trace + (size_t)gridDim.x * GS_TRACE_PHASES * 2 + (blockIdx.x * 4 + trN) * 4;
          o[0] = tk1 - tk0; o[1] = tk2 - tk1; o[2] = tkg - tk1; o[3] = nU * 1000 + nRs;
          trN++;
        }
        t0 = t1; x0 = x1; y0 = y1; pend0 = pend1;
        t1 = t2; x1 = x2; y1 = y2; pend1 = false;
        buf ^= 1;
      }
      if (!batch && lvl == nLevels - 1) {
        // the iteration's |delta lambda| total rides on the last colour barrier; totals rotate through three slots
        // so a slot can be cleared a full iteration before it is used again, without an extra barrier
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if (lane == 0) s_red[wic] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
          double t = 0.0;
          for (int k = 0; k < GS_WARPS; k++) t += s_red[k];
          atomicAdd(&G.worldTot[iter % 3], t);
        }
      }
      GS_TRACE_WORK();
      grid_barrier(S.bar, epoch, nCtas);
      GS_TRACE_END();
      if (!batch && lvl == 0 && tid == 0) G.worldTot[(iter + 2) % 3] = 0.0;  // last read before this barrier, next used in iter+2
    }
    bool allDone;
    if (!batch) {
      const double tot = __ldcg(&G.worldTot[iter % 3]);
      allDone = tot * tot < P.tol2;
    } else {
      if (threadIdx.x == 0) s_any = 0;
      __syncthreads();
      int anyLive = 0;
      for (int w = tid; w < G.nGroups; w += nth) {
        if (G.worldDone[w]) continue;
        const double tot = __ldcg(&G.worldTot[w]);
        if (tot * tot < P.tol2) { G.worldDone[w] = 1; G.worldIters[w] = iter; }
        else { anyLive = 1; G.worldTot[w] = 0.0; }
      }
      if (anyLive) atomicOr(&s_any, 1);
      __syncthreads();
      if (threadIdx.x == 0 && s_any) atomicOr(&G.worldDone[G.nGroups], 1);
      grid_barrier(S.bar, epoch, nCtas);
      allDone = __ldcg(&G.worldDone[G.nGroups]) == 0;
      grid_barrier(S.bar, epoch, nCtas);
      if (tid == 0) G.worldDone[G.nGroups] = 0;
    }
    if (allDone && !P.debugSkipWork) break;
  }
  // a copy started for a task that will never run must land before the CTA may retire
  if (pend0) mbar_wait(&s_mbar[wic][buf], (parity >> buf) & 1u);
  if (tid == 0) *G.itersDone = iter;
}

// ---- batches of small worlds: one CTA per world -------------------------------------------------------------------
// A batch (n_worlds > 1) has no coupling between worlds, so the colored sweep does not need the grid at all: the
// execution order is regrouped by world (k_world_count / k_world_fill), a CTA takes one world, keeps that world's
// vlambda / wlambda in shared memory, walks the world's units colour by colour with __syncthreads() between colours
// and leaves as soon as ITS world meets the tolerance (gs_solver.dart:105 per world, like separate World objects).
#define GW_THREADS 128
#define GW_MAXB 256    // bodies of a world whose lambdas fit the shared arrays (larger worlds use the global arrays)
#define GW_MAXU 2048   // units of a world sorted by colour in shared memory (larger worlds scan their unit range per colour)
#define GW_MAXL 256

// Execution order of a colored batch: grouped by world, by colour inside a world (counting sort over nW * GR_LV bins), so
// a world's units - and, because rows are built in execution order, its rows - are contiguous colour by colour.
#define GR_LV 32  // colours with their own bin; a batch with more colours takes the staged kernel, which re-sorts by colour
__global__ void __launch_bounds__(256) k_world_count(UnitArrays U, const int* __restrict__ order, const int* __restrict__ unitLevel,
                                                     const int* __restrict__ bodyWorld, int* __restrict__ binCount) {
  const int n = min(*U.nExec, U.unitCap);
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < n; a += gridDim.x * blockDim.x) {
    const int u = order[a];
    atomicAdd(&binCount[bodyWorld[U.uBi[u]] * GR_LV + min(unitLevel[u], GR_LV - 1)], 1);
  }
}

__global__ void __launch_bounds__(256) k_world_fill(UnitArrays U, const int* __restrict__ order, const int* __restrict__ unitLevel,
                                                    const int* __restrict__ bodyWorld, const int* __restrict__ binStart, int* __restrict__ binCursor,
                                                    int* __restrict__ orderW) {
  const int n = min(*U.nExec, U.unitCap);
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < n; a += gridDim.x * blockDim.x) {
    const int u = order[a], bin = bodyWorld[U.uBi[u]] * GR_LV + min(unitLevel[u], GR_LV - 1);
    orderW[binStart[bin] + atomicAdd(&binCursor[bin], 1)] = u;
  }
}

// ---- colored sweep: units of a colour ordered by row count ----------------------------------------------------------
// A warp of k_gs_fast owns a window of consecutive units, one unit per lane, and its time is the longest unit of the
// window (units have 3, 6, ... 24+ rows). Units of one colour are independent, so their order is free: a counting sort
// by (colour, row-count class descending) makes the windows homogeneous - a few windows of long units, many windows of
// 3-row units that finish after three rows. Colours beyond LEN_LEVELS keep their order.
#define LEN_LEVELS 64
#define LEN_CLASSES 8
#define LEN_BINS (LEN_LEVELS * LEN_CLASSES)
__device__ __forceinline__ int len_bin(int level, int rows) {
  const int cls = min(LEN_CLASSES - 1, max(rows - 1, 0) / 3);
  return level * LEN_CLASSES + (LEN_CLASSES - 1 - cls);
}
__global__ void __launch_bounds__(256) k_len_count(UnitArrays U, const int* __restrict__ order, const int* __restrict__ unitLevel,
                                                   int* __restrict__ bins) {
  const int n = min(*U.nExec, U.unitCap);
  const int lane = threadIdx.x & 31;
  for (int a0 = blockIdx.x * blockDim.x + threadIdx.x - lane; a0 < n; a0 += gridDim.x * blockDim.x) {
    const int a = a0 + lane;
    int bin = -1;
    if (a < n) {
      const int u = order[a], l = unitLevel[u];
      if (l < LEN_LEVELS) bin = len_bin(l, U.uRows[u]);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (bin >= 0 && lane == __ffs(peers) - 1) atomicAdd(&bins[bin], __popc(peers));
  }
}
// bins[0..LEN_BINS): counts -> bins[LEN_BINS..2*LEN_BINS): first execution position of the bin; cursors cleared
__global__ void __launch_bounds__(LEN_BINS) k_len_starts(const int* __restrict__ levelStart, const int* __restrict__ nLevels, int* __restrict__ bins) {
  const int b = threadIdx.x, l = b / LEN_CLASSES;
  int start = 0;
  if (l < *nLevels) {
    start = levelStart[l];
    for (int k = l * LEN_CLASSES; k < b; k++) start += bins[k];
  }
  bins[LEN_BINS + b] = start;
  bins[2 * LEN_BINS + b] = 0;
}
__global__ void __launch_bounds__(256) k_len_fill(UnitArrays U, const int* __restrict__ order, const int* __restrict__ unitLevel,
                                                  int* __restrict__ bins, int* __restrict__ orderL) {
  const int n = min(*U.nExec, U.unitCap);
  const int lane = threadIdx.x & 31;
  for (int a0 = blockIdx.x * blockDim.x + threadIdx.x - lane; a0 < n; a0 += gridDim.x * blockDim.x) {
    const int a = a0 + lane;
    int bin = -1, u = 0;
    if (a < n) {
      u = order[a];
      const int l = unitLevel[u];
      if (l < LEN_LEVELS) bin = len_bin(l, U.uRows[u]);
      else orderL[a] = u;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (bin >= 0 && lane == leader) base = atomicAdd(&bins[2 * LEN_BINS + bin], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    if (bin >= 0) orderL[bins[LEN_BINS + bin] + base + __popc(peers & ((1u << lane) - 1u))] = u;
  }
}

#define GW_ROWS_CAP 576   // rows of a world staged in shared memory (80 B each)
#define GW_UNITS_CAP 512  // unit records of a world staged in shared memory (32 B each)
#define GW_SMEM_BYTES (GW_ROWS_CAP * 80 + GW_UNITS_CAP * 32 + GW_ROWS_CAP * 4)

__global__ void __launch_bounds__(GW_THREADS) k_gs_world(RowArrays R, BodyArrays B, UnitArrays U, SchedArrays S, SolveParams P, GsStats G,
                                                         const int* __restrict__ worldStart, const int* __restrict__ worldBody, const int* __restrict__ ringOk) {
  extern __shared__ __align__(128) unsigned char s_dyn[];  // rows | unit records | lambdas of this world
  __shared__ float4 s_v[GW_MAXB], s_w[GW_MAXB];
  __shared__ unsigned short s_idx[GW_MAXU];
  __shared__ int s_start[GW_MAXL + 2], s_cur[GW_MAXL + 1];
  __shared__ double s_red[GW_THREADS / 32];
  __shared__ unsigned long long s_mbar;
  __shared__ int s_flag;
  if (ringOk && *ringOk) return;  // every world fits k_gs_world_ring: that kernel does the batch
  const int wd = blockIdx.x, tid = threadIdx.x;
  const int a0 = worldStart[(size_t)wd * GR_LV], nU = worldStart[(size_t)(wd + 1) * GR_LV] - a0;
  if (nU <= 0) return;
  const int b0 = worldBody[wd], nB = worldBody[wd + 1] - b0;
  const int r0w = U.eRowBase[a0], nR = U.eRowBase[a0 + nU] - r0w;
  // the whole world (rows, unit records) is copied to shared memory once with two bulk copies when it fits
  const bool staged = nR <= GW_ROWS_CAP && nU <= GW_UNITS_CAP && nR > 0;
  float4* const sRows = (float4*)s_dyn;
  GsUnitRec* const sUnits = (GsUnitRec*)(s_dyn + GW_ROWS_CAP * 80);
  float* const sLam = (float*)(s_dyn + GW_ROWS_CAP * 80 + GW_UNITS_CAP * 32);
  if (tid == 0) { mbar_init(&s_mbar, 1); s_flag = 0; }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (staged && tid == 0) {
    mbar_expect_tx(&s_mbar, (unsigned)(nR * 80 + nU * 32));
    bulk_g2s(sRows, R.rec + (size_t)r0w * 5, (unsigned)(nR * 80), &s_mbar);
    bulk_g2s(sUnits, U.rec + a0, (unsigned)(nU * 32), &s_mbar);
  }
  const bool sb = nB <= GW_MAXB;  // this world's vlambda / wlambda live in shared memory
  float4* const vl = sb ? s_v : B.vlam;  // global records are interleaved (stride 2), the shared arrays are not
  float4* const wl = sb ? s_w : B.wlam;
  const int ls = sb ? 1 : 2;
  const int boff = sb ? b0 : 0;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (sb) for (int i = tid; i < nB; i += GW_THREADS) { s_v[i] = z4; s_w[i] = z4; }
  if (staged) for (int i = tid; i < nR; i += GW_THREADS) sLam[i] = 0.f;  // k_rows_build left every lambda at zero
  for (int i = tid; i < GW_MAXL + 2; i += GW_THREADS) s_start[i] = 0;
  __syncthreads();
  bool sorted = nU <= GW_MAXU;
  if (sorted)
    for (int a = tid; a < nU; a += GW_THREADS) {
      const int l = U.eLevel[a0 + a];
      if (l >= GW_MAXL) s_flag = 1;
      else atomicAdd(&s_start[l + 2], 1);
    }
  __syncthreads();
  sorted = sorted && !s_flag;
  int nLv = *S.nLevels;
  if (sorted) {
    if (tid == 0) {  // s_start[l + 1] = first slot of colour l, s_start[0] = number of colours of this world
      int run = 0, last = 0;
      for (int l = 0; l < GW_MAXL; l++) {
        const int c = s_start[l + 2];
        s_start[l + 1] = run;
        s_cur[l] = run;
        run += c;
        if (c) last = l + 1;
      }
      s_start[GW_MAXL + 1] = run;
      s_start[0] = last;
    }
    __syncthreads();
    nLv = s_start[0];
    for (int a = tid; a < nU; a += GW_THREADS) s_idx[atomicAdd(&s_cur[U.eLevel[a0 + a]], 1)] = (unsigned short)a;
    __syncthreads();
  }
  if (staged) mbar_wait(&s_mbar, 0);
  const float4* const rowBase = staged ? sRows - (size_t)r0w * 5 : R.rec;
  float* const lamBase = staged ? sLam - r0w : R.flambda;
  int iter = 0;
  for (; iter != P.maxIter; iter++) {
    double local = 0.0;
    for (int l = 0; l < nLv; l++) {
      const int k0 = sorted ? s_start[l + 1] : 0, k1 = sorted ? s_start[l + 2] : nU;
      for (int k = k0 + tid; k < k1; k += GW_THREADS) {
        int a;
        if (sorted) a = s_idx[k];
        else { a = k; if (U.eLevel[a0 + a] != l) continue; }
        const GsUnitRec m = staged ? sUnits[a] : U.rec[a0 + a];
        if (m.r1 <= m.r0) continue;
        float4 vA = (m.fl & 1) ? vl[ls * (m.bi - boff)] : z4, wA = (m.fl & 1) ? wl[ls * (m.bi - boff)] : z4;
        float4 vB = (m.fl & 2) ? vl[ls * (m.bj - boff)] : z4, wB = (m.fl & 2) ? wl[ls * (m.bj - boff)] : z4;
        float acc = 0.f;
        const float4* q = rowBase + (size_t)m.r0 * 5;
        float* lp = lamBase + m.r0;
        for (int r = m.r0; r < m.r1; r++, q += 5, lp++) {
          const float4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4];
          const float lam = *lp;
          float gw_ = dot3f(q0, vB.x - vA.x, vB.y - vA.y, vB.z - vA.z);
          gw_ += dot3f(q1, wA.x, wA.y, wA.z);
          gw_ += dot3f(q2, wB.x, wB.y, wB.z);
          float dl = q1.w * (q0.w - gw_ - q2.w * lam);
          if (lam + dl < q3.w) dl = q3.w - lam;
          else if (lam + dl > q4.w) dl = q4.w - lam;
          *lp = lam + dl;
          if (m.fl & 1) {
            const float sA = -m.imA * dl;
            vA.x = fmaf(sA, q0.x, vA.x); vA.y = fmaf(sA, q0.y, vA.y); vA.z = fmaf(sA, q0.z, vA.z);
            wA.x = fmaf(dl, q3.x, wA.x); wA.y = fmaf(dl, q3.y, wA.y); wA.z = fmaf(dl, q3.z, wA.z);
          }
          if (m.fl & 2) {
            const float sB = m.imB * dl;
            vB.x = fmaf(sB, q0.x, vB.x); vB.y = fmaf(sB, q0.y, vB.y); vB.z = fmaf(sB, q0.z, vB.z);
            wB.x = fmaf(dl, q4.x, wB.x); wB.y = fmaf(dl, q4.y, wB.y); wB.z = fmaf(dl, q4.z, wB.z);
          }
          acc += fabsf(dl);
        }
        if (m.fl & 1) { vl[ls * (m.bi - boff)] = vA; wl[ls * (m.bi - boff)] = wA; }
        if (m.fl & 2) { vl[ls * (m.bj - boff)] = vB; wl[ls * (m.bj - boff)] = wB; }
        local += (double)acc;
      }
      __syncthreads();
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = local;
    __syncthreads();
    double tot = 0.0;
    for (int k = 0; k < GW_THREADS / 32; k++) tot += s_red[k];
    __syncthreads();
    if (tot * tot < P.tol2) break;
  }
  if (sb)
    for (int i = tid; i < nB; i += GW_THREADS) { B.vlam[2 * (b0 + i)] = s_v[i]; B.wlam[2 * (b0 + i)] = s_w[i]; }
  if (staged)
    for (int i = tid; i < nR; i += GW_THREADS) R.flambda[r0w + i] = sLam[i];
  if (tid == 0) { G.worldIters[wd] = iter; atomicMax(G.itersDone, iter); }
}

// ---- batches of small worlds, ring variant: one WARP per world, rows streamed colour by colour ----------------------
// The staged kernel keeps a whole world's rows in shared memory (2 worlds per SM) and is bound by the serial chain of a
// world: 180 colour phases x the longest unit, 14 waves of worlds. Here only what is reused lives in shared memory for
// the whole solve (unit records, row lambdas, body lambdas: ~11 KB); the rows of one colour are contiguous (the order
// above), so the warp copies them global -> shared with cp.async three colours ahead of the sweep into a four-slot
// ring. 9 worlds share an SM, the copies hide the DRAM latency, a phase costs only its own arithmetic.
#define GR_MAXB 96
#define GR_MAXU 160
#define GR_MAXR 704
#define GR_SLOTS 4
#define GR_SLOT_ROWS 40
#define GR_MAXSEG 96  // segments (runs of units of one colour that fit a ring slot) per world
#define GR_AHEAD (GR_SLOTS - 1)
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// ringOk = every world of the batch fits the shared tables of k_gs_world_ring
__global__ void __launch_bounds__(256) k_world_ring_check(UnitArrays U, const int* __restrict__ binStart, const int* __restrict__ worldBody, int nW,
                                                          const int* __restrict__ nLevels, int* __restrict__ ringOk) {
  if (*nLevels > GR_LV && blockIdx.x == 0 && threadIdx.x == 0) atomicExch(ringOk, 0);
  for (int wd = blockIdx.x * blockDim.x + threadIdx.x; wd < nW; wd += gridDim.x * blockDim.x) {
    const int* bs = binStart + (size_t)wd * GR_LV;
    const int a0 = bs[0], nU = bs[GR_LV] - a0;
    bool ok = nU <= GR_MAXU && worldBody[wd + 1] - worldBody[wd] <= GR_MAXB && U.eRowBase[a0 + nU] - U.eRowBase[a0] <= GR_MAXR;
    // the segmentation k_gs_world_ring will build: runs of units of one colour with at most GR_SLOT_ROWS rows
    int ns = 0;
    for (int l = 0, u = a0; l < GR_LV && ok; l++) {
      const int ue = bs[l + 1];
      while (u < ue && ok) {
        const int rows0 = U.eRowBase[u];
        if (U.eRowBase[u + 1] - rows0 > GR_SLOT_ROWS) ok = false;  // a single unit larger than a ring slot
        while (u < ue && U.eRowBase[u + 1] - rows0 <= GR_SLOT_ROWS) u++;
        ns++;
      }
    }
    if (!ok || ns > GR_MAXSEG) atomicExch(ringOk, 0);
  }
}

__global__ void __launch_bounds__(32) k_gs_world_ring(RowArrays R, BodyArrays B, UnitArrays U, SolveParams P, GsStats G,
                                                      const int* __restrict__ binStart, const int* __restrict__ worldBody,
                                                      const int* __restrict__ ringOk) {
  if (!*ringOk) return;
  __shared__ __align__(16) float4 s_vw[GR_MAXB * 2];
  __shared__ __align__(16) GsUnitRec s_units[GR_MAXU];
  __shared__ __align__(16) float4 s_ring[GR_SLOTS][GR_SLOT_ROWS * 5];
  __shared__ float s_lam[GR_MAXR];
  __shared__ int s_cs[GR_LV + 1];
  __shared__ unsigned char s_seg[GR_MAXSEG + 1];  // first unit of every segment (GR_MAXU <= 255)
  const int wd = blockIdx.x, lane = threadIdx.x;
  const int* bs = binStart + (size_t)wd * GR_LV;
  const int a0 = bs[0], nU = bs[GR_LV] - a0;
  if (nU <= 0) return;
  const int b0 = worldBody[wd], nB = worldBody[wd + 1] - b0;
  const int r0w = U.eRowBase[a0], nR = U.eRowBase[a0 + nU] - r0w;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = lane; i <= GR_LV; i += 32) s_cs[i] = bs[i] - a0;
  for (int i = lane; i < 2 * nB; i += 32) s_vw[i] = z4;
  for (int i = lane; i < nR; i += 32) s_lam[i] = 0.f;  // k_rows_build left every lambda at zero
  for (int i = lane; i < nU; i += 32) s_units[i] = U.rec[a0 + i];
  __syncwarp();
  // segments: runs of consecutive units of one colour whose rows fit a ring slot (units of a colour are independent,
  // and a single warp walks the segments in order, so colour boundaries need no more than the per-segment __syncwarp)
  if (lane == 0) {
    int ns = 0;
    for (int l = 0, u = 0; l < GR_LV; l++) {
      const int ue = s_cs[l + 1];
      while (u < ue) {
        const int rows0 = s_units[u].r0;
        s_seg[ns++] = (unsigned char)u;
        while (u < ue && s_units[u].r1 - rows0 <= GR_SLOT_ROWS) u++;
      }
    }
    s_seg[ns] = (unsigned char)nU;
    s_cs[0] = ns;  // s_cs is not needed any more
  }
  __syncwarp();
  const int nSeg = s_cs[0];
  const int nPhases = nSeg * P.maxIter;
  auto issue = [&](int p) {  // rows of phase p's segment -> ring slot p % GR_SLOTS (an empty group keeps the count uniform)
    if (p < nPhases) {
      const int k = p % nSeg;
      const int rlo = s_units[s_seg[k]].r0, rhi = s_units[s_seg[k + 1] - 1].r1;
      const float4* src = R.rec + (size_t)rlo * 5;
      float4* dst = s_ring[p % GR_SLOTS];
      const int n16 = (rhi - rlo) * 5;
      for (int i = lane; i < n16; i += 32) cp_async16(dst + i, src + i);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int p = 0; p < GR_AHEAD; p++) issue(p);
  int iter = 0, p = 0;
  for (; iter != P.maxIter; iter++) {
    double local = 0.0;  // same accumulation as k_gs_world: float inside a unit, double across units
    for (int k = 0; k < nSeg; k++, p++) {
      issue(p + GR_AHEAD);
      asm volatile("cp.async.wait_group %0;" ::"n"(GR_AHEAD) : "memory");
      __syncwarp();
      const float4* rows = s_ring[p % GR_SLOTS];
      const int u0 = s_seg[k], u1 = s_seg[k + 1];
      const int rlo = s_units[u0].r0;
      for (int u = u0 + lane; u < u1; u += 32) {
        const GsUnitRec m = s_units[u];
        if (m.r1 <= m.r0) continue;
        const int ia = 2 * (m.bi - b0), ib = 2 * (m.bj - b0);
        float4 vA = (m.fl & 1) ? s_vw[ia] : z4, wA = (m.fl & 1) ? s_vw[ia + 1] : z4;
        float4 vB = (m.fl & 2) ? s_vw[ib] : z4, wB = (m.fl & 2) ? s_vw[ib + 1] : z4;
        const float4* q = rows + (size_t)(m.r0 - rlo) * 5;
        float* lp = s_lam + (m.r0 - r0w);
        float acc = 0.f;
        for (int r = m.r0; r < m.r1; r++, q += 5, lp++) {
          const float4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4];
          const float lam = *lp;
          float gw_ = dot3f(q0, vB.x - vA.x, vB.y - vA.y, vB.z - vA.z);
          gw_ += dot3f(q1, wA.x, wA.y, wA.z);
          gw_ += dot3f(q2, wB.x, wB.y, wB.z);
          float dl = q1.w * (q0.w - gw_ - q2.w * lam);
          if (lam + dl < q3.w) dl = q3.w - lam;
          else if (lam + dl > q4.w) dl = q4.w - lam;
          *lp = lam + dl;
          if (m.fl & 1) {
            const float sA = -m.imA * dl;
            vA.x = fmaf(sA, q0.x, vA.x); vA.y = fmaf(sA, q0.y, vA.y); vA.z = fmaf(sA, q0.z, vA.z);
            wA.x = fmaf(dl, q3.x, wA.x); wA.y = fmaf(dl, q3.y, wA.y); wA.z = fmaf(dl, q3.z, wA.z);
          }
          if (m.fl & 2) {
            const float sB = m.imB * dl;
            vB.x = fmaf(sB, q0.x, vB.x); vB.y = fmaf(sB, q0.y, vB.y); vB.z = fmaf(sB, q0.z, vB.z);
            wB.x = fmaf(dl, q4.x, wB.x); wB.y = fmaf(dl, q4.y, wB.y); wB.z = fmaf(dl, q4.z, wB.z);
          }
          acc += fabsf(dl);
        }
        if (m.fl & 1) { s_vw[ia] = vA; s_vw[ia + 1] = wA; }
        if (m.fl & 2) { s_vw[ib] = vB; s_vw[ib + 1] = wB; }
        local += (double)acc;
      }
      __syncwarp();
    }
    double tot = local;
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (tot * tot < P.tol2) break;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  for (int i = lane; i < 2 * nB; i += 32) B.vlam[2 * b0 + i] = s_vw[i];
  for (int i = lane; i < nR; i += 32) R.flambda[r0w + i] = s_lam[i];
  if (lane == 0) { G.worldIters[wd] = iter; atomicMax(G.itersDone, iter); }
}
