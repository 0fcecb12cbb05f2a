// cannon_cuda.cu — libcannon_cuda.so: C ABI (include/cannon_cuda.h) + step orchestration.
//
// One cannon_world owns device-resident SoA state; World.internalStep (lib/world/world_class.dart:433-701) is a
// fixed sequence of kernel launches on one stream with every data-dependent count kept on the device, so
// nsteps steps are enqueued back to back without a host round trip. There is no CPU fallback: without a CUDA
// device cannon_ctx_create fails with CANNON_E_NOGPU.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "dmath.cuh"
#include "k_body.cuh"
#include "k_broadphase.cuh"
#include "k_narrowphase.cuh"
#include "k_sat_warp.cuh"
#include "k_solver.cuh"
#include "k_gs_exact.cuh"
#include "k_gs_world_exact.cuh"
#include "k_raycast.cuh"
#include "k_sph.cuh"
#include "world.cuh"

// ---- device counters ---------------------------------------------------------------------------------
enum {
  CT_NPAIRS = 0, CT_NTASKS, CT_NCONTACTS, CT_NROWS, CT_RAWCOUNT, CT_FRICTOTAL, CT_CONTTOTAL, CT_NLEVELS, CT_ITERS,
  CT_OVF_PAIRS, CT_OVF_TASKS, CT_OVF_CONTACTS, CT_OVF_ROWS, CT_OVF_LEVELS, CT_OVF_CLIP, CT_CURSOR, CT_ACT0, CT_ACT1,
  CT_GS_NTASKS, CT_NPAIRS_RAW, CT_NUNITS, CT_NEXEC, CT_NUNITS1, CT_ISL_CHANGED0, CT_ISL_CHANGED1, CT_NISLANDS, CT_RING_OK, CT_GS_ABORT, CT_NROWS_PAD, CT_NCLIP0, CT_NCLIP1, CT_NPPAIRS, CT_UNSUPPORTED, CT_BUCKETCOUNT, CT_BUCKETSTART = CT_BUCKETCOUNT + NP_NTYPES,
  CT_BUCKETCURSOR = CT_BUCKETSTART + NP_NTYPES, CT_BAR = ((CT_BUCKETCURSOR + NP_NTYPES + 31) / 32) * 32, CT_COUNT = CT_BAR + 64
};

__global__ void k_bucket_starts(int* cnt) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int s = 0;
    for (int t = 0; t < NP_NTYPES; t++) { cnt[CT_BUCKETSTART + t] = s; s += cnt[CT_BUCKETCOUNT + t]; cnt[CT_BUCKETCURSOR + t] = 0; }
  }
}
__global__ void k_set_int(int* p, int v) { if (threadIdx.x == 0 && blockIdx.x == 0) *p = v; }

// persistent accumulators (never reset by the per-step counter memset): statistics and sticky overflow needs
enum { AC_CONTACT_ITERS = 0, AC_STEPS, AC_OVF_PAIRS, AC_OVF_TASKS, AC_OVF_CONTACTS, AC_OVF_ROWS, AC_OVF_LEVELS, AC_OVF_CLIP, AC_GS_ABORT, AC_UNSUPPORTED, AC_COUNT };
__global__ void k_step_epilogue(const int* __restrict__ cnt, long long* __restrict__ acc, int taskCap, int contactCap, long long* __restrict__ clk,
                                double dt) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  clk[0] = __double_as_longlong(__longlong_as_double(clk[0]) + dt);  // World.step: time += dt after internalStep (world_class.dart:396-399)
  clk[1] += 1;
  acc[AC_CONTACT_ITERS] += (long long)cnt[CT_NCONTACTS] * cnt[CT_ITERS];
  acc[AC_STEPS] += 1;
  auto mx = [&](int slot, long long v) { if (v > acc[slot]) acc[slot] = v; };
  mx(AC_OVF_PAIRS, cnt[CT_OVF_PAIRS]);
  mx(AC_OVF_TASKS, cnt[CT_NTASKS] > taskCap ? cnt[CT_NTASKS] : cnt[CT_OVF_TASKS]);
  mx(AC_OVF_CONTACTS, cnt[CT_NCONTACTS] > contactCap ? cnt[CT_NCONTACTS] : cnt[CT_OVF_CONTACTS]);
  mx(AC_OVF_ROWS, cnt[CT_OVF_ROWS]);
  mx(AC_OVF_LEVELS, cnt[CT_OVF_LEVELS]);
  mx(AC_OVF_CLIP, cnt[CT_OVF_CLIP]);
  mx(AC_GS_ABORT, cnt[CT_GS_ABORT]);
  mx(AC_UNSUPPORTED, cnt[CT_UNSUPPORTED]);
}

// constraint-pair filter, world_class.dart:488-499: drop pairs joined by a constraint with collideConnected == false
__global__ void __launch_bounds__(256) k_pair_filter_flags(const int* __restrict__ p1, const int* __restrict__ p2, const int* __restrict__ nPairs,
                                                           int cap, const unsigned long long* __restrict__ keys, int nKeys, int* __restrict__ keep) {
  const int np = min(*nPairs, cap);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < np; k += gridDim.x * blockDim.x) {
    const unsigned a = (unsigned)min(p1[k], p2[k]), b = (unsigned)max(p1[k], p2[k]);
    const unsigned long long key = ((unsigned long long)a << 32) | b;
    int lo = 0, hi = nKeys;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] < key) lo = mid + 1; else hi = mid; }
    keep[k] = (lo < nKeys && keys[lo] == key) ? 0 : 1;
  }
}
__global__ void __launch_bounds__(256) k_pair_filter_compact(const int* __restrict__ p1, const int* __restrict__ p2, const int* __restrict__ nPairs,
                                                             int cap, const int* __restrict__ keep, const int* __restrict__ off,
                                                             int* __restrict__ q1, int* __restrict__ q2) {
  const int np = min(*nPairs, cap);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < np; k += gridDim.x * blockDim.x)
    if (keep[k]) { q1[off[k]] = p1[k]; q2[off[k]] = p2[k]; }
}

// Solver epilogue when the solve is invoked on its own (gs_solver.dart:111-121); the fused step folds this into k_integrate
__global__ void __launch_bounds__(256) k_apply_lambda(BodyArrays B, int n, int nWorlds, const int* __restrict__ worldRows) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (worldRows[nWorlds > 1 ? B.world[i] : 0] <= 0) continue;
    const f3 vl = vmulc(ld3(B.vlam[2 * i]), ld3(B.linF[i]));
    B.vel[i] = st3(vadd(vl, ld3(B.vel[i])));
    const f3 wl = vmulc(ld3(B.wlam[2 * i]), ld3(B.angF[i]));
    B.angvel[i] = st3(vadd(wl, ld3(B.angvel[i])));
  }
}
__global__ void __launch_bounds__(256) k_multipliers(ContactArrays C, RowArrays R, int contactCap, double invDt, double* __restrict__ out) {
  const int nc = min(*C.nContacts, contactCap);
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
    const int r = C.row[c];
    out[c] = r >= 0 ? (R.fast ? (double)R.flambda[r] : R.lambda[r]) * invDt : 0.0;  // Equation.multiplier, gs_solver.dart:124-129
  }
}

// host <-> device marshalling of 3-/4-vectors: the ABI uses tightly packed float3/float4 host arrays, the device float4;
// packing runs on the GPU so each attribute costs exactly one cudaMemcpy of the caller's buffer
__global__ void __launch_bounds__(256) k_pack_to4(const float* __restrict__ src, float4* __restrict__ dst, int n, int comps) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    dst[i] = make_float4(src[comps * i], src[comps * i + 1], src[comps * i + 2], comps == 4 ? src[4 * i + 3] : 0.f);
}
__global__ void __launch_bounds__(256) k_unpack_from4(const float4* __restrict__ src, float* __restrict__ dst, int n, int comps) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 v = src[i];
    dst[comps * i] = v.x; dst[comps * i + 1] = v.y; dst[comps * i + 2] = v.z;
    if (comps == 4) dst[4 * i + 3] = v.w;
  }
}

// ---- world -------------------------------------------------------------------------------------------
struct HostShape {
  int type = 0, collisionResponse = 1, group = -1, mask = -1;
  double radius = 0, bsr = 0;
  float he[3] = {0, 0, 0};
  int hull = -1, hf = -1, tm = -1, material = -1;
};

struct cannon_world {
  cannon_ctx* ctx = nullptr;
  cannon_world_desc desc;
  int n = 0;
  double time = 0, dt = -1, powDt = -1;
  int64_t stepnumber = 0;
  cannon_profile prof{};

  // host mirrors needed for derived quantities
  std::vector<HostShape> hShapes;
  std::vector<HostHull> hHulls;
  std::vector<HfDev> hHfs;
  std::vector<double> hHfData;
  std::vector<double> hLdamp, hAdamp;
  std::vector<int> hBig, hBigWorldStart, hWorldStart;
  // compound bodies (cannon_world_set_body_shapes): the table as given, and what cannon_world_set_bodies made of it
  std::vector<int> hInstFirst, hInstShape, hInstBody;
  std::vector<float4> hInstOff, hInstQuat;
  bool compound = false;  // some body has != 1 shape, or a shape away from the body origin / rotated against it
  int nInst = 0, maxInst = 1, ppCap = 0;
  DBuf<int> dInstFirst, dInstShape, dInstBody, pxType, pxFlags, pxMaterial, pp1, pp2, ppCnt, ppOff, ppPer;
  DBuf<float4> dInstOff, dInstQuat, pxPos, pxQuat;
  DBuf<double> pxInvMass;
  int maxWorldBodies = 0;  // largest world of a batch (chooses the all-pairs broadphase for small worlds)
  bool hasOversizeHull = false;  // some hull exceeds the tile kernel's scratch: run the sequential SAT kernels for those tasks
  double cell = 1.0;
  int nBig = 0, hashSize = 1024;
  int nMat = 0;

  // device: bodies
  DBuf<float4> pos, quat, vel, angvel, force, torque, lam, iiw0, iiw1, iiw2, invI, linF, angF, aabbLo, aabbHi;
  DBuf<double> mass, invMass, brad, ldamp, adamp, ldpow, adpow, sleepSpeed, sleepTime, tLastSleepy;
  DBuf<int> type, sleep, shape, material, group, mask, world, flags;
  // device: shape tables
  DBuf<ShapeDev> dShapes;
  DBuf<HullDev> dHulls;
  DBuf<float4> dVerts, dFnormals, dEdges, dEdgesK;
  DBuf<int> dFacesK;
  DBuf<PillarRec> dPillars;
  // particleConvex state of the shape table (ShapeTables.pc*): only allocated when a Particle shape exists
  // SPHSystem subsystems (k_sph.cuh)
  struct HostSph { int n = 0; double density = 1, h = 1, cs = 1, viscosity = 0.01, eps = 0.00001; DBuf<int> particles; DBuf<double> densities, pressures; };
  std::vector<HostSph> sph;
  bool hasParticle = false, hasTrimesh = false, hasShapeMaterial = false;
  std::vector<TrimeshDev> hTms;
  std::vector<float4> hTmVerts, hTmNormals;
  std::vector<int> hTmIdx;
  DBuf<TrimeshDev> dTms;
  DBuf<float4> dTmVerts, dTmNormals;
  DBuf<int> dTmIdx;
  DBuf<int> dPcFrozen, dPcFreezeTask;
  DBuf<float4> dPcPos, dPcQuat;
  DBuf<double> dFplanec, dHfData, dMatFriction, dMatRestitution;
  DBuf<int> dFvOff, dFvIdx, dFcOff, dFcIdx, dCmTable;
  DBuf<HfDev> dHfs;
  DBuf<cannon_contact_material> dCms;
  // device: broadphase
  DBuf<int4> cellc, smeta, scell;
  DBuf<int> nbCache, clipList;
  DBuf<float4> taskSep;
  DBuf<int> binLo, binHi, cellStart, cellEnd, bigList, bigWorldStart, worldStart, bpCounts, bpOffs;
  DBuf<uint32_t> skey, sval, sapKey, sapList;
  DBuf<float4> spos;
  DBuf<double> srad;
  DBuf<int> p1, p2, q1, q2, keep, keepOff;
  DBuf<unsigned long long> filterKeys;
  int nFilterKeys = 0;
  bool sapInit = false;
  int pairCap = 0;
  // device: narrowphase
  DBuf<int> pairTasks, pairTaskOff, taskPair, taskInfo, bucket, taskCnt, taskRaw, taskOff, taskHit;
  DBuf<unsigned long long> pairMask;
  DBuf<int2> taskCell;
  DBuf<float4> rawRi, rawRj, rawNi;
  int taskCap = 0, contactCap = 0;
  // device: contacts
  DBuf<int> cBi, cBj, cTask, cEnabled, cRow, fricFlag, contFlag, fricOff, contOff;
  DBuf<float4> cRi, cRj, cNi;
  DBuf<double> cRest, cMu, cSlip, cCa, cCb, cCeps, cFb, cFeps, cMult;
  // device: rows
  DBuf<int> rKind;
  DBuf<float4> rN, rRA, rRB, rIA, rIB;
  DBuf<double> rB, rInvC, rEps, rMinF, rMaxF, rLambda;
  DBuf<float4> rRec;
  DBuf<GsUnitRec> uRec;
  DBuf<float4> rXblk;     // COLORED (exact), single world: window-interleaved row blocks / unit records of k_gs_exact
  DBuf<int> gxWinRows, gxWinBase;  // row slots per window (32 x longest unit) and their exclusive scan
  int gxWinCap = 0;
  DBuf<GxUnit> uXrec;
  DBuf<int> unitSeq, gxBody;  // per unit: ranks on its two bodies; per body: [0, n) scheduled units, [n, 2n) progress counters
  // contact events (opt-in)
  bool evEnabled = false;
  int evCap = 0;
  unsigned evMask = 0;
  DBuf<unsigned long long> evKeysCur, evKeysPrev, evTabCur, evTabPrev, evBegin, evEnd;
  DBuf<int> evCnt;
  DBuf<int> eLevel, orderW, worldCount, worldUnitStart, lenBins;
  DBuf<int2> gsTab;
  DBuf<int> gsLvlTask, gsLvlWin;
  int gsTaskCap = 0;
  DBuf<float> rFlambda;
  int rowCap = 0;
  // device: solver units
  DBuf<int> uBi, uBj, uFlags, uRows, uSrc, uKey, uPri, worldKeys, eBi, eBj, eFlags, eRowBase, eRows, unitRow;
  DBuf<double> eImA, eImB;
  int unitCap = 0;
  // device: joints
  DBuf<int> jBodyA, jBodyB, jKind, jEnabled, jRowSlot, jFirst, jSlotEq;
  DBuf<float4> jPivotA, jPivotB, jAxisA, jAxisB, jNi;
  DBuf<double> jMinF, jMaxF, jA, jB, jEps, jTargetVel, jCos, jParam;
  DBuf<int> jMode;
  int nJointEq = 0, nJointAccepted = 0;
  std::vector<int> hConFirst, hConType, hJEnabled, hJTrig;  // host mirror for cannon_world_set_hinge_motor: first equation / type per constraint, flags per equation
  // device: scheduler / gs
  DBuf<unsigned long long> claim;
  DBuf<int> unitLevel, order, levelStart, act0, act1, worldRows, worldDone, worldIters, islandLabel;
  DBuf<double> worldTot;
  // springs (cannon_world_set_springs)
  int nSprings = 0;
  DBuf<int> spBodyA, spBodyB, spOff, spIdx;
  DBuf<double> spRest, spK, spD;
  DBuf<float4> spAnchorA, spAnchorB;
  DBuf<long long> dClock;  // [0] bits of World.time, [1] World.stepnumber
  long long hClock[2] = {0, 0};
  DBuf<long long> gsTrace;
  int maxLevels = 0;
  // counters
  DBuf<int> cnt;
  int* hCnt = nullptr;  // pinned
  ScanTmp scanTmp;
  SortTmp sortTmp;
  DBuf<float> stage;  // marshalling buffer (4 floats per body)
  DBuf<long long> acc;
  long long* hAcc = nullptr;  // pinned
  cudaEvent_t ev[11] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  long long lastUnits = 0, lastLevels = 0;  // widths seen by the last synchronised call (sizes the cooperative grids)
  bool recordSolveEvents = false;
  bool stepPending = false;             // cannon_world_step_async enqueued work that cannon_ctx_sync has not collected yet
  std::vector<cudaEvent_t> profPool;    // cannon_world_step_profiled: PROF_EV events per step
  bool stageEventsValid = false;        // the call's last step ran eagerly: ev[0..7,10] were recorded outside a graph and can be read
  // one World.step captured as a CUDA graph (all counts live on the device, so the launch sequence of a step is
  // fixed for a given dt and capacity; the cooperative kernels size their own barrier on the device); replayed by
  // cannon_world_step
  cudaGraphExec_t stepGraph = nullptr;
  double graphDt = 0.0;
  long long graphLaunches = 0;     // kernels per replay (for cannon_profile.kernel_launches)
  int eagerSteps = 0;              // steps run eagerly since the last (re)allocation: lazily sized scratch exists after one
  bool graphBroken = false;        // capture failed once: stay eager
  int coopBlocksSched = 0, coopBlocksGs = 0, coopBlocksGsFast = 0, coopBlocksGsFastV1 = 0, coopBlocksGx = 0;
  bool gxOff = false;     // CANNON_GS_NO_DATAFLOW: COLORED falls back to the unstaged level-by-level sweep k_gs (A/B measurements)
  bool gsFastV1 = false;  // CANNON_GS_FAST_V1: the unstaged colored sweep, kept for A/B measurements
  bool gwNoRing = false;     // CANNON_GW_NO_RING: batches always take the staged CTA-per-world kernel (A/B measurements)
  bool gsNoLenSort = false;  // CANNON_GS_NO_LEN_SORT: leave the units of a colour in schedule order (A/B measurements)
  // resolver kernels of different types are independent: they run on side streams between two events
  cudaStream_t npStream[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t npFork = nullptr, npJoin[3] = {nullptr, nullptr, nullptr};

  ~cannon_world() {
    if (hCnt) cudaFreeHost(hCnt);
    if (hAcc) cudaFreeHost(hAcc);
    for (auto& st : npStream) if (st) cudaStreamDestroy(st);
    if (npFork) cudaEventDestroy(npFork);
    for (auto& e : npJoin) if (e) cudaEventDestroy(e);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    for (auto& e : profPool) if (e) cudaEventDestroy(e);
  }
};

// COLORED sweeps the colour order with the reference arithmetic (f64, bit-exact against the oracle's restatement of the same
// order); COLORED_F32 packs the rows to f32 and sweeps with FMA (reduced precision, opt-in)
static inline bool kind_colored(const cannon_world* w) { return w->desc.solver_kind == CANNON_SOLVER_COLORED || w->desc.solver_kind == CANNON_SOLVER_COLORED_F32; }
static inline bool kind_fast(const cannon_world* w) { return w->desc.solver_kind == CANNON_SOLVER_COLORED_F32; }
// the exact colour sweep of a single world runs as the staged dataflow kernel k_gs_exact over packed 96-byte rows
static inline bool kind_packed(const cannon_world* w) { return w->desc.solver_kind == CANNON_SOLVER_COLORED && w->desc.n_worlds <= 1 && !w->gxOff; }

static int32_t fail(cannon_ctx* ctx, int32_t code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

#define W_TRY(w, expr) CU_TRY((w)->ctx, expr)

template <class T>
static cudaError_t upload(DBuf<T>& d, const std::vector<T>& h, cudaStream_t s) {
  cudaError_t e = d.reserve(h.size() ? h.size() : 1);
  if (e != cudaSuccess) return e;
  if (h.empty()) return cudaSuccess;
  return cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s);
}

static BodyArrays body_arrays(cannon_world* w) {
  BodyArrays B;
  B.pos = w->pos.p; B.quat = w->quat.p; B.vel = w->vel.p; B.angvel = w->angvel.p; B.force = w->force.p; B.torque = w->torque.p;
  B.vlam = w->lam.p; B.wlam = w->lam.p ? w->lam.p + 1 : nullptr; /* interleaved: one 32-byte record per body */ B.iiw0 = w->iiw0.p; B.iiw1 = w->iiw1.p; B.iiw2 = w->iiw2.p;
  B.invI = w->invI.p; B.linF = w->linF.p; B.angF = w->angF.p; B.aabbLo = w->aabbLo.p; B.aabbHi = w->aabbHi.p;
  B.mass = w->mass.p; B.invMass = w->invMass.p; B.brad = w->brad.p; B.ldamp = w->ldamp.p; B.adamp = w->adamp.p;
  B.ldpow = w->ldpow.p; B.adpow = w->adpow.p; B.sleepSpeed = w->sleepSpeed.p; B.sleepTime = w->sleepTime.p; B.tLastSleepy = w->tLastSleepy.p;
  B.type = w->type.p; B.sleep = w->sleep.p; B.shape = w->shape.p; B.material = w->material.p; B.group = w->group.p; B.mask = w->mask.p;
  B.world = w->world.p; B.flags = w->flags.p;
  B.bpos = w->pos.p; B.bquat = w->quat.p; B.owner = nullptr;
  return B;
}
// what the narrowphase kernels see: the bodies, or the shape instances of a world with compound bodies
static BodyArrays np_body_arrays(cannon_world* w) {
  BodyArrays B = body_arrays(w);
  if (w->compound) {
    B.pos = w->pxPos.p; B.quat = w->pxQuat.p; B.shape = w->dInstShape.p; B.type = w->pxType.p; B.flags = w->pxFlags.p;
    B.material = w->pxMaterial.p; B.invMass = w->pxInvMass.p; B.owner = w->dInstBody.p;
  }
  return B;
}
static ShapeTables shape_tables(cannon_world* w) {
  ShapeTables T;
  T.shapes = w->dShapes.p; T.hulls = w->dHulls.p; T.verts = w->dVerts.p; T.fnormals = w->dFnormals.p; T.fplanec = w->dFplanec.p;
  T.fvOff = w->dFvOff.p; T.fvIdx = w->dFvIdx.p; T.fcOff = w->dFcOff.p; T.fcIdx = w->dFcIdx.p; T.edges = w->dEdges.p;
  T.edgesK = w->dEdgesK.p; T.facesK = w->dFacesK.p; T.pillars = w->dPillars.p;
  T.hfs = w->dHfs.p; T.hfdata = w->dHfData.p; T.cmTable = w->dCmTable.p; T.cms = w->dCms.p;
  T.matFriction = w->dMatFriction.p; T.matRestitution = w->dMatRestitution.p; T.nMat = w->nMat;
  T.tms = w->dTms.p; T.tmVerts = w->dTmVerts.p; T.tmNormals = w->dTmNormals.p; T.tmIdx = w->dTmIdx.p;
  T.instFirst = w->compound ? w->dInstFirst.p : nullptr; T.instShape = w->dInstShape.p; T.instOff = w->dInstOff.p; T.instQuat = w->dInstQuat.p;
  T.nShapes = (int)w->hShapes.size();
  T.pcFrozen = w->dPcFrozen.p; T.pcFreezeTask = w->dPcFreezeTask.p; T.pcPos = w->dPcPos.p; T.pcQuat = w->dPcQuat.p;
  return T;
}
static int grid_for(cannon_world* w, long long n, int threads) {
  int g = div_up(n > 0 ? n : 1, threads);
  int mx = w->ctx->sms * 16;
  return g < 1 ? 1 : (g > mx ? mx : g);
}

extern "C" {

int32_t cannon_version(void) { return CANNON_ABI_VERSION; }
const char* cannon_backend(void) { return "cuda"; }

int32_t cannon_ctx_create(int32_t device, cannon_ctx** out) {
  if (!out) return CANNON_E_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return CANNON_E_NOGPU;
  cannon_ctx* ctx = new (std::nothrow) cannon_ctx();
  if (!ctx) return CANNON_E_INVALID;
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return CANNON_E_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return CANNON_E_CUDA; }
  *out = ctx;
  return CANNON_OK;
}
void cannon_ctx_destroy(cannon_ctx* ctx) {
  if (!ctx) return;
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}
const char* cannon_last_error(const cannon_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

void cannon_world_desc_default(cannon_world_desc* d) {
  memset(d, 0, sizeof(*d));
  d->solver_kind = CANNON_SOLVER_REFERENCE_ORDER;
  d->solver_iterations = 10;   // lib/solver/solver.dart:17
  d->solver_tolerance = 1e-7;  // lib/solver/solver.dart:18
  d->broadphase_kind = CANNON_BP_NAIVE;  // world_class.dart:153
  d->grid_nx = d->grid_ny = d->grid_nz = 10;  // grid_broadphase.dart:37-41
  for (int k = 0; k < 3; k++) { d->grid_min[k] = 100; d->grid_max[k] = -100; }
  cannon_contact_material& cm = d->default_contact_material;  // world_class.dart:155-158, contact_material.dart:45-70
  cm.material_a = cm.material_b = -1;
  cm.friction = 0.3;
  cm.restitution = 0.0;
  cm.contact_equation_stiffness = 1e7;
  cm.contact_equation_relaxation = 3;
  cm.friction_equation_stiffness = 1e7;
  cm.friction_equation_relaxation = 3;
  d->n_worlds = 1;
}
void cannon_shape_desc_default(cannon_shape_desc* d) {
  memset(d, 0, sizeof(*d));
  d->type = CANNON_SHAPE_SPHERE;
  d->collision_response = 1;
  d->collision_filter_group = -1;
  d->collision_filter_mask = -1;
  d->radius = 1.0;
  d->radius_top = d->radius_bottom = d->height = 1.0;
  d->num_segments = 8;
  d->hf_element_size = 1;
  d->tm_scale[0] = d->tm_scale[1] = d->tm_scale[2] = 1.f;
  d->material = -1;
}

int32_t cannon_world_create(cannon_ctx* ctx, const cannon_world_desc* desc, cannon_world** out) {
  if (!ctx || !desc || !out) return CANNON_E_INVALID;
  cudaSetDevice(ctx->device);
  cannon_world* w = new (std::nothrow) cannon_world();
  if (!w) return CANNON_E_INVALID;
  w->ctx = ctx;
  w->desc = *desc;
  if (w->desc.n_worlds < 1) w->desc.n_worlds = 1;
  if (w->desc.broadphase_kind == CANNON_BP_GRID && (desc->grid_nx > 1024 || desc->grid_ny > 1024 || desc->grid_nz > 1024 ||
                                                    desc->grid_nx < 1 || desc->grid_ny < 1 || desc->grid_nz < 1)) {
    delete w;
    return fail(ctx, CANNON_E_INVALID, "GridBroadphase: each dimension's n must be in 1..1024");
  }
  if (w->desc.n_worlds > 1 && w->desc.broadphase_kind == CANNON_BP_SAP) {
    delete w;
    return fail(ctx, CANNON_E_UNSUPPORTED, "batched worlds support NaiveBroadphase / GridBroadphase only");
  }
  if (cudaMallocHost((void**)&w->hCnt, CT_COUNT * sizeof(int)) != cudaSuccess) { delete w; return fail(ctx, CANNON_E_CUDA, "cudaMallocHost failed"); }
  memset(w->hCnt, 0, CT_COUNT * sizeof(int));
  if (w->cnt.reserve(CT_COUNT) != cudaSuccess) { delete w; return fail(ctx, CANNON_E_CUDA, "cudaMalloc failed"); }
  cudaMemsetAsync(w->cnt.p, 0, CT_COUNT * sizeof(int), ctx->stream);
  if (cudaMallocHost((void**)&w->hAcc, AC_COUNT * sizeof(long long)) != cudaSuccess || w->acc.reserve(AC_COUNT) != cudaSuccess) {
    delete w;
    return fail(ctx, CANNON_E_CUDA, "allocation failed");
  }
  memset(w->hAcc, 0, AC_COUNT * sizeof(long long));
  cudaMemsetAsync(w->acc.p, 0, AC_COUNT * sizeof(long long), ctx->stream);
  for (auto& e : w->ev) cudaEventCreate(&e);
  {
    // the hull / heightfield-pillar SAT kernel (npStream[1]) is the longest resolver: its CTAs get the SMs first, the
    // shorter resolvers on the other streams fill in behind it
    int prLo = 0, prHi = 0;
    cudaDeviceGetStreamPriorityRange(&prLo, &prHi);
    for (int k = 0; k < 3; k++) cudaStreamCreateWithPriority(&w->npStream[k], cudaStreamNonBlocking, k == 1 ? prHi : prLo);
  }
  cudaEventCreateWithFlags(&w->npFork, cudaEventDisableTiming);
  for (auto& e : w->npJoin) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  // empty tables so kernels always get valid pointers
  std::vector<double> e0;
  upload(w->dMatFriction, e0, ctx->stream);
  upload(w->dMatRestitution, e0, ctx->stream);
  w->dCmTable.reserve(1);
  w->dCms.reserve(1);
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_schedule, 256, 0);
  w->coopBlocksSched = ctx->sms * std::max(1, std::min(occ, 4));
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gs, 256, 0);
  w->coopBlocksGs = ctx->sms * std::max(1, std::min(occ, 4));
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gs_fast_v1, 256, 0);
  w->coopBlocksGsFastV1 = ctx->sms * std::max(1, std::min(occ, 4));
  cudaFuncSetAttribute(k_gs_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_SMEM_BYTES);
  cudaFuncSetAttribute(k_np_tasks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * 32 * QS_HOIST_AXES * sizeof(QsAxis)));
  cudaFuncSetAttribute(k_gs_world, cudaFuncAttributeMaxDynamicSharedMemorySize, GW_SMEM_BYTES);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gs_fast, GS_THREADS, GS_SMEM_BYTES);
  w->coopBlocksGsFast = ctx->sms * std::max(1, std::min(occ, 4));
  cudaFuncSetAttribute(k_gs_world_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, GWX_SMEM_BYTES);
  cudaFuncSetAttribute(k_gs_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, GX_SMEM_BYTES);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gs_exact, GX_THREADS, GX_SMEM_BYTES);
  w->coopBlocksGx = ctx->sms * std::max(1, std::min(occ, 1));
  w->gxOff = getenv("CANNON_GS_NO_DATAFLOW") != nullptr;
  w->gsFastV1 = getenv("CANNON_GS_FAST_V1") != nullptr;
  w->gsNoLenSort = getenv("CANNON_GS_NO_LEN_SORT") != nullptr;
  w->gwNoRing = getenv("CANNON_GW_NO_RING") != nullptr;
  *out = w;
  return CANNON_OK;
}
static void drop_step_graph(cannon_world* w);

void cannon_world_destroy(cannon_world* w) {
  if (w && w->ctx) { auto& pv = w->ctx->pending; pv.erase(std::remove(pv.begin(), pv.end(), w), pv.end()); }
  if (!w) return;
  drop_step_graph(w);
  cudaSetDevice(w->ctx->device);
  cudaStreamSynchronize(w->ctx->stream);
  // DBuf members are released explicitly (plain structs, no destructors)
#define REL(x) w->x.release()
  for (auto& hs : w->sph) { hs.particles.release(); hs.densities.release(); hs.pressures.release(); }
  REL(pos); REL(quat); REL(vel); REL(angvel); REL(force); REL(torque); REL(lam); REL(iiw0); REL(iiw1); REL(iiw2);
  REL(invI); REL(linF); REL(angF); REL(aabbLo); REL(aabbHi); REL(mass); REL(invMass); REL(brad); REL(ldamp); REL(adamp); REL(ldpow);
  REL(adpow); REL(sleepSpeed); REL(sleepTime); REL(tLastSleepy); REL(type); REL(sleep); REL(shape); REL(material); REL(group); REL(mask);
  REL(world); REL(flags); REL(dShapes); REL(dHulls); REL(dVerts); REL(dFnormals); REL(dEdges); REL(dFplanec); REL(dHfData); REL(dEdgesK); REL(dFacesK); REL(dPillars); REL(dInstFirst); REL(dInstShape); REL(dInstBody); REL(pxType); REL(pxFlags); REL(pxMaterial); REL(pp1); REL(pp2); REL(ppCnt); REL(ppOff); REL(ppPer); REL(dInstOff); REL(dInstQuat); REL(pxPos); REL(pxQuat); REL(pxInvMass); REL(dTms); REL(dTmVerts); REL(dTmNormals); REL(dTmIdx); REL(dPcFrozen); REL(dPcFreezeTask); REL(dPcPos); REL(dPcQuat);
  REL(dMatFriction); REL(dMatRestitution); REL(dFvOff); REL(dFvIdx); REL(dFcOff); REL(dFcIdx); REL(dCmTable); REL(dHfs); REL(dCms);
  REL(clipList); REL(taskSep); REL(nbCache); REL(cellc); REL(smeta); REL(scell); REL(binLo); REL(binHi); REL(cellStart); REL(cellEnd); REL(bigList); REL(bigWorldStart);
  REL(worldStart); REL(bpCounts); REL(bpOffs); REL(skey); REL(sval); REL(sapKey); REL(sapList); REL(spos); REL(srad); REL(p1); REL(p2);
  REL(q1); REL(q2); REL(keep); REL(keepOff); REL(filterKeys); REL(pairMask); REL(pairTasks); REL(pairTaskOff); REL(taskPair); REL(taskInfo); REL(bucket);
  REL(taskCnt); REL(taskRaw); REL(taskOff); REL(taskHit); REL(taskCell); REL(rawRi); REL(rawRj); REL(rawNi); REL(cBi); REL(cBj); REL(cTask); REL(cEnabled); REL(cRow);
  REL(fricFlag); REL(contFlag); REL(fricOff); REL(contOff); REL(cRi); REL(cRj); REL(cNi); REL(cRest); REL(cMu); REL(cSlip); REL(cCa);
  REL(cCb); REL(cCeps); REL(cFb); REL(cFeps); REL(cMult); REL(rKind); REL(rN); REL(rRA); REL(rRB);
  REL(rIA); REL(rIB); REL(rB); REL(rInvC); REL(rEps); REL(rMinF); REL(rMaxF); REL(rLambda); REL(jBodyA); REL(jBodyB);
  REL(uBi); REL(uBj); REL(uFlags); REL(uRows); REL(uSrc); REL(uKey); REL(uPri); REL(worldKeys); REL(eBi); REL(eBj); REL(eFlags); REL(eRowBase); REL(eRows); REL(unitRow);
  REL(eImA); REL(eImB); REL(jSlotEq); REL(rRec); REL(uRec); REL(rXblk); REL(gxWinRows); REL(gxWinBase); REL(uXrec); REL(unitSeq); REL(gxBody); REL(eLevel); REL(evKeysCur); REL(evKeysPrev); REL(evTabCur); REL(evTabPrev); REL(evBegin); REL(evEnd); REL(evCnt); REL(orderW); REL(worldCount); REL(lenBins); REL(worldUnitStart); REL(gsTab); REL(gsLvlTask); REL(gsLvlWin); REL(rFlambda);
  REL(jKind); REL(jEnabled); REL(jRowSlot); REL(jFirst); REL(jPivotA); REL(jPivotB); REL(jAxisA); REL(jAxisB); REL(jNi); REL(jMinF);
  REL(jMaxF); REL(jA); REL(jB); REL(jEps); REL(jTargetVel); REL(jCos); REL(jParam); REL(jMode); REL(claim); REL(unitLevel); REL(order); REL(levelStart); REL(act0); REL(act1);
  REL(worldRows); REL(worldDone); REL(worldIters); REL(worldTot); REL(dClock); REL(spBodyA); REL(spBodyB); REL(spOff); REL(spIdx); REL(spRest); REL(spK); REL(spD); REL(spAnchorA); REL(spAnchorB); REL(gsTrace); REL(cnt); REL(acc); REL(stage); REL(islandLabel);
  w->scanTmp.tiles.release();
  w->sortTmp.k2.release(); w->sortTmp.v2.release(); w->sortTmp.hist.release(); w->sortTmp.scan.tiles.release();
#undef REL
  delete w;
}

int32_t cannon_world_set_materials(cannon_world* w, int32_t n, const double* friction, const double* restitution, int32_t ncm,
                                   const cannon_contact_material* cms) {
  if (w) drop_step_graph(w);  // buffers may move: the captured step is rebuilt on the next cannon_world_step
  if (!w || n < 0 || ncm < 0) return CANNON_E_INVALID;
  cudaSetDevice(w->ctx->device);
  cudaStream_t s = w->ctx->stream;
  std::vector<double> f(n, -1.0), r(n, -1.0);
  for (int i = 0; i < n; i++) { if (friction) f[i] = friction[i]; if (restitution) r[i] = restitution[i]; }
  std::vector<int> table((size_t)n * n, -1);
  std::vector<cannon_contact_material> v(cms, cms + ncm);
  for (int k = 0; k < ncm; k++) {
    const int a = cms[k].material_a, b = cms[k].material_b;
    if (a < 0 || b < 0 || a >= n || b >= n) return fail(w->ctx, CANNON_E_INVALID, "contact material references unknown material");
    table[(size_t)a * n + b] = k;  // unordered key, lib/utils/tuple_dictionary.dart:1-3
    table[(size_t)b * n + a] = k;
  }
  W_TRY(w, upload(w->dMatFriction, f, s));
  W_TRY(w, upload(w->dMatRestitution, r, s));
  W_TRY(w, upload(w->dCmTable, table, s));
  W_TRY(w, upload(w->dCms, v, s));
  W_TRY(w, cudaStreamSynchronize(s));
  w->nMat = n;
  return CANNON_OK;
}

int32_t cannon_world_set_shapes(cannon_world* w, int32_t n, const cannon_shape_desc* sd) {
  if (w) drop_step_graph(w);  // buffers may move: the captured step is rebuilt on the next cannon_world_step
  if (!w || n < 0 || (n > 0 && !sd)) return CANNON_E_INVALID;
  cudaSetDevice(w->ctx->device);
  cudaStream_t s = w->ctx->stream;
  w->hShapes.assign(n, HostShape());
  w->hHulls.clear();
  w->hHfs.clear();
  w->hHfData.clear();
  w->hTms.clear(); w->hTmVerts.clear(); w->hTmNormals.clear(); w->hTmIdx.clear();
  long long nPillars = 0;
  for (int i = 0; i < n; i++) {
    const cannon_shape_desc& d = sd[i];
    HostShape& h = w->hShapes[i];
    h.type = d.type;
    h.collisionResponse = d.collision_response != 0;
    h.group = d.collision_filter_group;
    h.mask = d.collision_filter_mask;
    h.material = d.material;
    if (h.material >= w->nMat) return fail(w->ctx, CANNON_E_INVALID, "shape references unknown material");
    switch (d.type) {
      case CANNON_SHAPE_SPHERE:
        if (d.radius < 0) return fail(w->ctx, CANNON_E_INVALID, "The sphere radius cannot be negative.");  // sphere.dart:16-18
        h.radius = d.radius;
        h.bsr = d.radius;
        break;
      case CANNON_SHAPE_PLANE:
        h.bsr = INFINITY;  // plane.dart:20
        break;
      case CANNON_SHAPE_PARTICLE:
        h.bsr = 0;  // particle.dart:24-26
        break;
      case CANNON_SHAPE_TRIMESH: {  // Trimesh constructor, trimesh.dart:60-73
        if (d.n_vertices <= 0 || d.n_triangles <= 0 || !d.vertices || !d.tm_indices) return fail(w->ctx, CANNON_E_INVALID, "trimesh needs vertices and indices");
        for (int k = 0; k < 3 * d.n_triangles; k++)
          if (d.tm_indices[k] < 0 || d.tm_indices[k] >= d.n_vertices) return fail(w->ctx, CANNON_E_INVALID, "trimesh index out of range");
        TrimeshDev m;
        m.vOff = (int)w->hTmVerts.size(); m.nV = d.n_vertices; m.iOff = (int)w->hTmIdx.size(); m.nT = d.n_triangles;
        std::vector<f3> raw(d.n_vertices), sc(d.n_vertices);
        for (int v = 0; v < d.n_vertices; v++) {  // getVertex :260-268: setValues, then *= scale component by component
          raw[v].x = d.vertices[3 * v]; raw[v].y = d.vertices[3 * v + 1]; raw[v].z = d.vertices[3 * v + 2];
          sc[v] = mk3(W(raw[v].x) * W(d.tm_scale[0]), W(raw[v].y) * W(d.tm_scale[1]), W(raw[v].z) * W(d.tm_scale[2]));
          w->hTmVerts.push_back(st3(sc[v]));
        }
        for (int t = 0; t < d.n_triangles; t++) {  // updateNormals :157-175 with computeNormal(vb, va, vc) :220-230, unit scale
          const f3 &va = raw[d.tm_indices[3 * t]], &vb = raw[d.tm_indices[3 * t + 1]], &vc = raw[d.tm_indices[3 * t + 2]];
          const f3 ab = vsub(va, vb), cb = vsub(vc, va);
          f3 nn = vcross(cb, ab);
          if (!(nn.x == 0 && nn.y == 0 && nn.z == 0)) vnormalize(nn);
          w->hTmNormals.push_back(st3(nn));
          for (int k = 0; k < 3; k++) w->hTmIdx.push_back(d.tm_indices[3 * t + k]);
        }
        f3 l = sc[0], u = sc[0];  // computeLocalAABB :315-343 (note the else-if), updateBoundingSphereRadius :350-363
        double max2 = 0;
        for (const f3& v : sc) {
          if (v.x < l.x) l.x = v.x; else if (v.x > u.x) u.x = v.x;
          if (v.y < l.y) l.y = v.y; else if (v.y > u.y) u.y = v.y;
          if (v.z < l.z) l.z = v.z; else if (v.z > u.z) u.z = v.z;
          max2 = fmax(max2, vlen2(v));
        }
        m.lo = st3(l); m.hi = st3(u);
        h.bsr = sqrt(max2);
        h.tm = (int)w->hTms.size();
        w->hTms.push_back(m);
        break;
      }
      case CANNON_SHAPE_BOX: {
        for (int k = 0; k < 3; k++) h.he[k] = d.half_extents[k];
        HostHull hull;
        host_box_hull(h.he, hull);
        h.hull = (int)w->hHulls.size();
        w->hHulls.push_back(hull);
        f3 e; e.x = h.he[0]; e.y = h.he[1]; e.z = h.he[2];
        h.bsr = vlen(e);  // box.dart:124-126
        break;
      }
      case CANNON_SHAPE_CYLINDER: {
        if (d.radius_top < 0 || d.radius_bottom < 0) return fail(w->ctx, CANNON_E_INVALID, "The cylinder radius cannot be negative.");
        if (d.num_segments < 3) return fail(w->ctx, CANNON_E_INVALID, "cylinder needs >= 3 segments");
        HostHull hull;
        host_cylinder_hull(d.radius_top, d.radius_bottom, d.height, d.num_segments, hull);
        h.hull = (int)w->hHulls.size();
        h.bsr = hull.bsr;  // after Body.addShape -> updateBoundingSphereRadius (rigid_body.dart:403)
        w->hHulls.push_back(hull);
        break;
      }
      case CANNON_SHAPE_CONVEX:
      case CANNON_SHAPE_CAPSULE:
      case CANNON_SHAPE_CONE:
      case CANNON_SHAPE_SIZED_PLANE: {
        if (d.n_vertices <= 0 || d.n_faces <= 0 || !d.vertices || !d.face_offsets || !d.face_indices)
          return fail(w->ctx, CANNON_E_INVALID, "convex shape needs vertices and faces");
        HostHull hull;
        for (int v = 0; v < d.n_vertices; v++) { f3 p; p.x = d.vertices[3 * v]; p.y = d.vertices[3 * v + 1]; p.z = d.vertices[3 * v + 2]; hull.v.push_back(p); }
        for (int f = 0; f < d.n_faces; f++) {
          std::vector<int> face(d.face_indices + d.face_offsets[f], d.face_indices + d.face_offsets[f + 1]);
          if (face.size() < 3) return fail(w->ctx, CANNON_E_INVALID, "convex face needs >= 3 vertices");
          for (int idx : face) if (idx < 0 || idx >= d.n_vertices) return fail(w->ctx, CANNON_E_INVALID, "convex face index out of range");
          hull.faces.push_back(face);
        }
        hull.hasAxes = d.convex_has_axes != 0;  // no `axes` => no face-normal axes; Cone passes them
        hull.finish();
        h.hull = (int)w->hHulls.size();
        h.bsr = hull.bsr;
        w->hHulls.push_back(hull);
        break;
      }
      case CANNON_SHAPE_HEIGHTFIELD: {
        if (d.hf_nx < 2 || d.hf_ny < 2 || !d.hf_data) return fail(w->ctx, CANNON_E_INVALID, "heightfield needs >= 2x2 samples");
        HfDev hf;
        hf.nx = d.hf_nx; hf.ny = d.hf_ny; hf.esize = d.hf_element_size; hf.dataOff = (int)w->hHfData.size();
        const size_t cntv = (size_t)d.hf_nx * d.hf_ny;
        double mn = d.hf_data[0], mx = d.hf_data[0];  // updateMinValue / updateMaxValue, heightfield.dart:87-113
        for (size_t k = 0; k < cntv; k++) { mn = d.hf_data[k] < mn ? d.hf_data[k] : mn; mx = d.hf_data[k] > mx ? d.hf_data[k] : mx; }
        hf.minV = mn; hf.maxV = mx;
        hf.pilOff = nPillars;
        nPillars += (long long)(d.hf_nx - 1) * (d.hf_ny - 1) * 2;
        w->hHfData.insert(w->hHfData.end(), d.hf_data, d.hf_data + cntv);
        h.hf = (int)w->hHfs.size();
        w->hHfs.push_back(hf);
        const double es = (double)hf.esize;  // updateBoundingSphereRadius, heightfield.dart:505-515
        f3 t = mk3(hf.nx * es, hf.ny * es, fmax(fabs(mx), fabs(mn)));
        h.bsr = vlen(t);
        break;
      }
      default:
        return fail(w->ctx, CANNON_E_UNSUPPORTED, "shape type outside the hot-path scope (SURVEY.md §8)");
    }
  }
  // flatten
  std::vector<ShapeDev> shapes(n);
  for (int i = 0; i < n; i++) {
    const HostShape& h = w->hShapes[i];
    ShapeDev& d = shapes[i];
    d.type = h.type; d.collisionResponse = h.collisionResponse; d.group = h.group; d.mask = h.mask;
    d.radius = h.radius; d.bsr = h.bsr; d.hx = h.he[0]; d.hy = h.he[1]; d.hz = h.he[2]; d.hull = h.hull; d.hf = h.hf; d.tm = h.tm; d.material = h.material; d.pad = 0;
  }
  std::vector<HullDev> hulls;
  w->hasOversizeHull = false;
  for (const HostHull& h : w->hHulls) if (h.faces.size() > 32 || h.edges.size() > 32) w->hasOversizeHull = true;
  std::vector<float4> verts, fnormals, edges, edgesK;
  std::vector<int> facesK;
  std::vector<double> fplanec;
  std::vector<int> fvOff, fvIdx, fcOff, fcIdx;
  int totalFaces = 0;
  for (size_t k = 0; k < w->hHulls.size(); k++) {
    const HostHull& h = w->hHulls[k];
    HullDev d;
    d.vOff = (int)verts.size(); d.nV = (int)h.v.size();
    d.fOff = totalFaces; d.nF = (int)h.faces.size();
    d.eOff = (int)edges.size(); d.nE = (int)h.edges.size();
    d.hasAxes = h.hasAxes ? 1 : 0; d.pad = 0; d.bsr = h.bsr;
    for (const f3& p : h.v) verts.push_back(st3(p));
    for (const f3& p : h.edges) edges.push_back(st3(p));
    // axes that are numerically +-equal in the local frame stay +-equal after Quaternion.vmult: the later one can
    // never win findSeparatingAxis' strict `d < dmin` (k_sat_warp.cuh), so the tile kernel only walks these lists
    auto pm_eq = [](const f3& a, const f3& b) {
      return (a.x == b.x && a.y == b.y && a.z == b.z) || (a.x == -b.x && a.y == -b.y && a.z == -b.z);
    };
    d.ekOff = (int)edgesK.size(); d.nEk = 0;
    for (size_t i = 0; i < h.edges.size(); i++) {
      bool dup = false;
      for (size_t p = 0; p < i && !dup; p++) dup = pm_eq(h.edges[p], h.edges[i]);
      if (!dup) { edgesK.push_back(st3(h.edges[i])); d.nEk++; }
    }
    d.fkOff = (int)facesK.size(); d.nFk = 0;
    for (size_t i = 0; i < h.faces.size(); i++) {
      bool dup = false;
      for (size_t p = 0; p < i && !dup; p++) dup = pm_eq(h.n[p], h.n[i]);
      if (!dup) { facesK.push_back((int)i); d.nFk++; }
    }
    for (size_t f = 0; f < h.faces.size(); f++) {
      fnormals.push_back(st3(h.n[f]));
      fplanec.push_back(h.planec[f]);
      fvOff.push_back((int)fvIdx.size());
      for (int idx : h.faces[f]) fvIdx.push_back(idx);
      fcOff.push_back((int)fcIdx.size());
      for (int idx : h.connected[f]) fcIdx.push_back(idx);
    }
    fvOff.push_back((int)fvIdx.size());  // CSR terminator per hull: segment of hull k starts at fOff + k
    fcOff.push_back((int)fcIdx.size());
    totalFaces += d.nF;
    hulls.push_back(d);
  }
  W_TRY(w, upload(w->dShapes, shapes, s));
  W_TRY(w, upload(w->dHulls, hulls, s));
  W_TRY(w, upload(w->dVerts, verts, s));
  W_TRY(w, upload(w->dFnormals, fnormals, s));
  W_TRY(w, upload(w->dEdges, edges, s));
  W_TRY(w, upload(w->dFplanec, fplanec, s));
  W_TRY(w, upload(w->dFvOff, fvOff, s));
  W_TRY(w, upload(w->dFvIdx, fvIdx, s));
  W_TRY(w, upload(w->dFcOff, fcOff, s));
  W_TRY(w, upload(w->dFcIdx, fcIdx, s));
  W_TRY(w, upload(w->dEdgesK, edgesK, s));
  W_TRY(w, upload(w->dFacesK, facesK, s));
  W_TRY(w, upload(w->dHfs, w->hHfs, s));
  W_TRY(w, upload(w->dHfData, w->hHfData, s));
  w->hasTrimesh = !w->hTms.empty();
  w->hasShapeMaterial = false;
  for (const HostShape& h : w->hShapes) if (h.material >= 0) w->hasShapeMaterial = true;
  if (w->hasTrimesh) {
    W_TRY(w, upload(w->dTms, w->hTms, s)); W_TRY(w, upload(w->dTmVerts, w->hTmVerts, s)); W_TRY(w, upload(w->dTmNormals, w->hTmNormals, s));
    W_TRY(w, upload(w->dTmIdx, w->hTmIdx, s));
  }
  if (nPillars > 0) {
    W_TRY(w, w->dPillars.reserve((size_t)nPillars));
    const ShapeTables T = shape_tables(w);
    for (size_t k = 0; k < w->hHfs.size(); k++) {
      const HfDev& hf = w->hHfs[k];
      const long long np = (long long)(hf.nx - 1) * (hf.ny - 1) * 2;
      g_kernel_launches++;
      k_pillars_build<<<grid_for(w, np, 128), 128, 0, s>>>(T, hf, w->dPillars.p);
    }
    W_TRY(w, cudaGetLastError());
  }
  // a new shape table starts with no ConvexPolyhedron.worldVertices computed (convex_polyhedron.dart:101-103)
  w->hasParticle = false;
  for (const HostShape& h : w->hShapes) if (h.type == CANNON_SHAPE_PARTICLE) w->hasParticle = true;
  if (w->hasParticle) {
    const size_t nt = (size_t)n + (size_t)nPillars;
    W_TRY(w, w->dPcFrozen.reserve(nt)); W_TRY(w, w->dPcFreezeTask.reserve(nt)); W_TRY(w, w->dPcPos.reserve(nt)); W_TRY(w, w->dPcQuat.reserve(nt));
    W_TRY(w, cudaMemsetAsync(w->dPcFrozen.p, 0, nt * sizeof(int), s));
    W_TRY(w, cudaMemsetAsync(w->dPcFreezeTask.p, 0x7f, nt * sizeof(int), s));  // 0x7f7f7f7f: above every task index
  }
  W_TRY(w, cudaStreamSynchronize(s));
  return CANNON_OK;
}

// host twin of shape_aabb for the upload path (Body.updateMassProperties needs the AABB once)
static void host_shape_aabb(const cannon_world* w, int shapeIdx, const f3& pos, const q4& q, f3& mn, f3& mx) {
  if (shapeIdx < 0) { mn = pos; mx = pos; return; }
  const HostShape& s = w->hShapes[shapeIdx];
  const float inf = INFINITY;
  switch (s.type) {
    case CANNON_SHAPE_SPHERE: {
      const double r = s.radius;
      mn = mk3(W(pos.x) - r, W(pos.y) - r, W(pos.z) - r);
      mx = mk3(W(pos.x) + r, W(pos.y) + r, W(pos.z) + r);
      break;
    }
    case CANNON_SHAPE_PLANE: {
      f3 z; z.x = 0.f; z.y = 0.f; z.z = 1.f;
      const f3 n = qrot(q, z);
      mn.x = mn.y = mn.z = -inf;
      mx.x = mx.y = mx.z = inf;
      if (n.x == 1.f) mx.x = pos.x; else if (n.x == -1.f) mn.x = pos.x;
      if (n.y == 1.f) mx.y = pos.y; else if (n.y == -1.f) mn.y = pos.y;
      if (n.z == 1.f) mx.z = pos.z; else if (n.z == -1.f) mn.z = pos.z;
      break;
    }
    case CANNON_SHAPE_BOX: {
      const float sx[8] = {1, -1, -1, -1, 1, 1, -1, 1}, sy[8] = {1, 1, -1, -1, -1, 1, 1, -1}, sz[8] = {1, 1, 1, -1, -1, -1, -1, 1};
      for (int i = 0; i < 8; i++) {
        f3 c; c.x = sx[i] * s.he[0]; c.y = sy[i] * s.he[1]; c.z = sz[i] * s.he[2];
        const f3 p = vadd(qrot(q, c), pos);
        if (i == 0) { mn = p; mx = p; continue; }
        if (p.x > mx.x) mx.x = p.x;
        if (p.y > mx.y) mx.y = p.y;
        if (p.z > mx.z) mx.z = p.z;
        if (p.x < mn.x) mn.x = p.x;
        if (p.y < mn.y) mn.y = p.y;
        if (p.z < mn.z) mn.z = p.z;
      }
      break;
    }
    case CANNON_SHAPE_CONVEX:
    case CANNON_SHAPE_CYLINDER:
    case CANNON_SHAPE_CAPSULE:
    case CANNON_SHAPE_CONE:
    case CANNON_SHAPE_SIZED_PLANE: {
      const HostHull& h = w->hHulls[s.hull];
      for (size_t i = 0; i < h.v.size(); i++) {
        const f3 p = vadd(qrot(q, h.v[i]), pos);
        if (i == 0) { mn = p; mx = p; continue; }
        if (p.x < mn.x) mn.x = p.x;
        if (p.x > mx.x) mx.x = p.x;
        if (p.y < mn.y) mn.y = p.y;
        if (p.y > mx.y) mx.y = p.y;
        if (p.z < mn.z) mn.z = p.z;
        if (p.z > mx.z) mx.z = p.z;
      }
      break;
    }
    case CANNON_SHAPE_PARTICLE:  // particle.dart:29-33
      mn = pos;
      mx = pos;
      break;
    case CANNON_SHAPE_TRIMESH: {  // trimesh.dart:366-375 (device twin: shape_aabb)
      const TrimeshDev& m = w->hTms[s.tm];
      const f3 l = ld3(m.lo), u = ld3(m.hi);
      for (int i = 0; i < 8; i++) {
        f3 c;
        c.x = (i == 0 || i == 3 || i == 5 || i == 6) ? l.x : u.x;
        c.y = (i == 0 || i == 1 || i == 4 || i == 6) ? l.y : u.y;
        c.z = (i == 0 || i == 1 || i == 2 || i == 5) ? l.z : u.z;
        const f3 p = to_world_point(pos, q, c);
        if (i == 0) { mn = p; mx = p; continue; }
        if (p.x > mx.x) mx.x = p.x;
        if (p.x < mn.x) mn.x = p.x;
        if (p.y > mx.y) mx.y = p.y;
        if (p.y < mn.y) mn.y = p.y;
        if (p.z > mx.z) mx.z = p.z;
        if (p.z < mn.z) mn.z = p.z;
      }
      break;
    }
    default:
      mn.x = mn.y = mn.z = -inf;
      mx.x = mx.y = mx.z = inf;
      break;
  }
}

static int32_t ensure_capacities(cannon_world* w) {
  const int n = w->n;
  const int pairCap = w->desc.max_pairs > 0 ? w->desc.max_pairs : std::max(4096, 24 * n);
  const int contactCap = w->desc.max_contacts > 0 ? w->desc.max_contacts : std::max(4096, 32 * n);
  // compound bodies: the narrowphase pairs are shape-instance pairs
  long long ppc = w->compound ? (long long)pairCap * w->maxInst : pairCap;
  const int npPairCap = (int)std::min<long long>(ppc, 1 << 28);
  w->ppCap = w->compound ? npPairCap : 0;
  const int taskCap = std::max(2 * npPairCap, 48 * std::max(n, w->nInst));
  const int rowCap = 3 * contactCap + w->nJointAccepted + 16;
  w->pairCap = pairCap; w->contactCap = contactCap; w->taskCap = taskCap; w->rowCap = rowCap;
  w->maxLevels = std::min(rowCap, 1 << 20);
#define RES(buf, cnt) W_TRY(w, w->buf.reserve((size_t)(cnt)))
  RES(p1, pairCap); RES(p2, pairCap); RES(q1, pairCap); RES(q2, pairCap); RES(keep, pairCap); RES(keepOff, pairCap);
  RES(pairTasks, npPairCap); RES(pairTaskOff, npPairCap); RES(pairMask, npPairCap);
  if (w->compound) { RES(pp1, npPairCap); RES(pp2, npPairCap); RES(ppPer, npPairCap); RES(ppCnt, pairCap); RES(ppOff, pairCap); }
  RES(taskPair, taskCap); RES(taskInfo, taskCap); RES(taskCell, taskCap); RES(bucket, taskCap); RES(taskCnt, taskCap); RES(taskRaw, taskCap);
  RES(taskOff, taskCap);
  if (w->evEnabled) RES(taskHit, taskCap);
  RES(rawRi, contactCap); RES(rawRj, contactCap); RES(rawNi, contactCap);
  RES(cBi, contactCap); RES(cBj, contactCap); RES(cTask, contactCap); RES(cEnabled, contactCap); RES(cRow, contactCap); RES(fricFlag, contactCap);
  RES(contFlag, contactCap); RES(fricOff, contactCap); RES(contOff, contactCap); RES(cRi, contactCap); RES(cRj, contactCap);
  RES(cNi, contactCap); RES(cRest, contactCap); RES(cMu, contactCap); RES(cSlip, contactCap); RES(cCa, contactCap); RES(cCb, contactCap);
  RES(cCeps, contactCap); RES(cFb, contactCap); RES(cFeps, contactCap); RES(cMult, contactCap);
  if (kind_fast(w)) {
    RES(rRec, (size_t)rowCap * 5); RES(rFlambda, rowCap + 32); RES(uRec, rowCap + 3);
    w->gsTaskCap = rowCap / GS_WIN_MIN + w->maxLevels + 2;
    RES(gsTab, w->gsTaskCap + 2); RES(gsLvlTask, w->maxLevels + 2); RES(gsLvlWin, w->maxLevels + 2);
  } else if (kind_packed(w)) {
    RES(rXblk, ((size_t)rowCap / 32 + 2) * GX_CHUNKS * 32); RES(uXrec, rowCap + 3); RES(rMinF, rowCap + 32); RES(rMaxF, rowCap + 32); RES(rLambda, rowCap + 64);
    w->gxWinCap = rowCap / 32 + w->maxLevels + 2;
    RES(gxWinRows, w->gxWinCap + 1); RES(gxWinBase, w->gxWinCap + 1);
    w->gsTaskCap = 0x7fffffff;
    RES(gsLvlTask, w->maxLevels + 2); RES(gsLvlWin, w->maxLevels + 2);
  } else {
    RES(rKind, rowCap); RES(rN, rowCap); RES(rRA, rowCap); RES(rRB, rowCap);
    RES(rIA, rowCap); RES(rIB, rowCap); RES(rB, rowCap); RES(rInvC, rowCap); RES(rEps, rowCap); RES(rMinF, rowCap); RES(rMaxF, rowCap);
    RES(rLambda, rowCap);
  }
  const int unitCap = rowCap + 2;
  w->unitCap = unitCap;
  RES(uBi, unitCap); RES(uBj, unitCap); RES(uFlags, unitCap); RES(uRows, unitCap); RES(uSrc, unitCap); RES(uKey, unitCap); RES(uPri, unitCap); RES(worldKeys, WK_ARRAYS * (size_t)std::max(1, w->desc.n_worlds) + 1); RES(eBi, unitCap); RES(eBj, unitCap);
  RES(eFlags, unitCap); RES(eRowBase, unitCap + 1); RES(eRows, unitCap + 1); RES(unitRow, unitCap); RES(eImA, unitCap); RES(eImB, unitCap);
  RES(lenBins, 3 * LEN_BINS); RES(eLevel, unitCap); RES(orderW, unitCap); RES(worldCount, 2 * ((size_t)w->desc.n_worlds * GR_LV + 4)); RES(worldUnitStart, (size_t)w->desc.n_worlds * GR_LV + 4);
  RES(unitLevel, unitCap); RES(order, unitCap); RES(act0, unitCap); RES(act1, unitCap); RES(levelStart, w->maxLevels + 2);
  RES(claim, n + 1);
  const int nW = w->desc.n_worlds;
  const int nGroupsMax = w->desc.solver_kind == CANNON_SOLVER_SPLIT ? std::max(nW, n) : nW;  // islands are labelled by body index
  RES(worldRows, nW + 1); RES(worldDone, nGroupsMax + 2); RES(worldIters, nGroupsMax + 1); RES(worldTot, nGroupsMax + 3);
  RES(islandLabel, n + 1);
#undef RES
  return CANNON_OK;
}

int32_t cannon_world_set_body_shapes(cannon_world* w, int32_t n_bodies, const int32_t* first, const int32_t* shape, const float* offset,
                                     const float* orientation) {
  if (!w || n_bodies < 0) return CANNON_E_INVALID;
  w->hInstFirst.clear(); w->hInstShape.clear(); w->hInstOff.clear(); w->hInstQuat.clear();
  if (n_bodies == 0) return CANNON_OK;
  if (!first || first[0] != 0) return fail(w->ctx, CANNON_E_INVALID, "cannon_world_set_body_shapes: first[0] must be 0");
  for (int b = 0; b < n_bodies; b++) if (first[b + 1] < first[b]) return fail(w->ctx, CANNON_E_INVALID, "cannon_world_set_body_shapes: first[] must ascend");
  const int ni = first[n_bodies];
  if (ni > 0 && !shape) return fail(w->ctx, CANNON_E_INVALID, "cannon_world_set_body_shapes: shape[] missing");
  for (int k = 0; k < ni; k++) {
    if (shape[k] < 0 || shape[k] >= (int)w->hShapes.size()) return fail(w->ctx, CANNON_E_INVALID, "body references unknown shape");
    w->hInstShape.push_back(shape[k]);
    w->hInstOff.push_back(offset ? make_float4(offset[3 * k], offset[3 * k + 1], offset[3 * k + 2], 0.f) : make_float4(0.f, 0.f, 0.f, 0.f));
    w->hInstQuat.push_back(orientation ? make_float4(orientation[4 * k], orientation[4 * k + 1], orientation[4 * k + 2], orientation[4 * k + 3])
                                       : make_float4(0.f, 0.f, 0.f, 1.f));
  }
  w->hInstFirst.assign(first, first + n_bodies + 1);
  return CANNON_OK;
}

int32_t cannon_world_set_bodies(cannon_world* w, const cannon_bodies_soa* sb) {
  if (w) drop_step_graph(w);  // buffers may move: the captured step is rebuilt on the next cannon_world_step
  if (!w || !sb || sb->n < 0) return CANNON_E_INVALID;
  cudaSetDevice(w->ctx->device);
  cudaStream_t s = w->ctx->stream;
  const int n = sb->n;
  const int nW = w->desc.n_worlds;
  std::vector<float4> pos(n), quat(n), vel(n), angvel(n), force(n), torque(n), zero4(n, make_float4(0, 0, 0, 0)), iiw0(n), iiw1(n), iiw2(n),
      invI(n), linF(n), angF(n);
  std::vector<double> mass(n), invMass(n), brad(n), ldamp(n), adamp(n), sleepSpeed(n), sleepTime(n), tLast(n);
  std::vector<int> type(n), sleep(n), shape(n), material(n), group(n), mask(n), world(n), flags(n);
  auto g3 = [](const float* a, int i, float dx, float dy, float dz) {
    return a ? make_float4(a[3 * i], a[3 * i + 1], a[3 * i + 2], 0.f) : make_float4(dx, dy, dz, 0.f);
  };
  double radSum = 0;
  int radCnt = 0;
  const bool haveTable = !w->hInstFirst.empty();
  if (haveTable && (int)w->hInstFirst.size() != n + 1) return fail(w->ctx, CANNON_E_INVALID, "cannon_world_set_body_shapes described another body count");
  w->compound = false;
  w->nInst = 0;
  w->maxInst = 1;
  if (haveTable) {
    for (int sh : w->hInstShape) if (sh >= (int)w->hShapes.size()) return fail(w->ctx, CANNON_E_INVALID, "the body shape table references a shape the current shape table does not have");
    w->nInst = w->hInstFirst[n];
    w->hInstBody.assign(w->nInst, 0);
    for (int i = 0; i < n; i++) {
      const int c = w->hInstFirst[i + 1] - w->hInstFirst[i];
      if (c != 1) w->compound = true;
      w->maxInst = std::max(w->maxInst, c);
      for (int k = w->hInstFirst[i]; k < w->hInstFirst[i + 1]; k++) w->hInstBody[k] = i;
    }
    for (int k = 0; k < w->nInst; k++) {
      const float4 o = w->hInstOff[k], q = w->hInstQuat[k];
      if (o.x != 0.f || o.y != 0.f || o.z != 0.f || q.x != 0.f || q.y != 0.f || q.z != 0.f || q.w != 1.f) w->compound = true;
    }
  }
  for (int i = 0; i < n; i++) {
    pos[i] = g3(sb->position, i, 0, 0, 0);
    quat[i] = sb->quaternion ? make_float4(sb->quaternion[4 * i], sb->quaternion[4 * i + 1], sb->quaternion[4 * i + 2], sb->quaternion[4 * i + 3])
                             : make_float4(0, 0, 0, 1);
    vel[i] = g3(sb->velocity, i, 0, 0, 0);
    angvel[i] = g3(sb->angular_velocity, i, 0, 0, 0);
    force[i] = g3(sb->force, i, 0, 0, 0);
    torque[i] = g3(sb->torque, i, 0, 0, 0);
    linF[i] = g3(sb->linear_factor, i, 1, 1, 1);
    angF[i] = g3(sb->angular_factor, i, 1, 1, 1);
    mass[i] = sb->mass ? sb->mass[i] : 0.0;
    type[i] = mass[i] <= 0.0 ? CANNON_BODY_STATIC : CANNON_BODY_DYNAMIC;  // rigid_body.dart:61
    if (sb->type && sb->type[i] >= 0) type[i] = sb->type[i];
    sleep[i] = sb->sleep_state ? sb->sleep_state[i] : CANNON_AWAKE;
    tLast[i] = sb->time_last_sleepy ? sb->time_last_sleepy[i] : w->time;  // world_class.dart:291
    sleepSpeed[i] = sb->sleep_speed_limit ? sb->sleep_speed_limit[i] : 0.1;
    sleepTime[i] = sb->sleep_time_limit ? sb->sleep_time_limit[i] : 1.0;
    ldamp[i] = sb->linear_damping ? sb->linear_damping[i] : 0.01;
    adamp[i] = sb->angular_damping ? sb->angular_damping[i] : 0.01;
    group[i] = sb->collision_filter_group ? sb->collision_filter_group[i] : 1;
    mask[i] = sb->collision_filter_mask ? sb->collision_filter_mask[i] : -1;
    material[i] = sb->material ? sb->material[i] : -1;
    shape[i] = sb->shape ? sb->shape[i] : -1;
    if (haveTable) shape[i] = w->hInstFirst[i + 1] > w->hInstFirst[i] ? w->hInstShape[w->hInstFirst[i]] : -1;  // shapes[0]
    world[i] = sb->world_id ? sb->world_id[i] : 0;
    if (shape[i] >= (int)w->hShapes.size()) return fail(w->ctx, CANNON_E_INVALID, "body references unknown shape");
    if (material[i] >= w->nMat) return fail(w->ctx, CANNON_E_INVALID, "body references unknown material");
    if (world[i] < 0 || world[i] >= nW) return fail(w->ctx, CANNON_E_INVALID, "world_id out of range");
    if (nW > 1 && i > 0 && world[i] < world[i - 1]) return fail(w->ctx, CANNON_E_INVALID, "bodies of a batched world must be contiguous (world_id ascending)");
    int fl = 0;
    if (!sb->allow_sleep || sb->allow_sleep[i]) fl |= BF_ALLOW_SLEEP;
    if (!sb->collision_response || sb->collision_response[i]) fl |= BF_COLLISION_RESPONSE;
    if (sb->is_trigger && sb->is_trigger[i]) fl |= BF_IS_TRIGGER;
    const bool fixedRot = sb->fixed_rotation && sb->fixed_rotation[i];
    if (fixedRot) fl |= BF_FIXED_ROTATION;
    // Body.updateMassProperties, rigid_body.dart:587-609: inertia of the current world AABB box
    invMass[i] = mass[i] > 0 ? 1.0 / mass[i] : 0;
    f3 mn, mx;
    const f3 p = ld3(pos[i]);
    const q4 q = ldq(quat[i]);
    if (w->compound) {  // Body.updateAABB over the instances, rigid_body.dart:415-447
      for (int k = w->hInstFirst[i]; k < w->hInstFirst[i + 1]; k++) {
        const f3 offset = vadd(qrot(q, ld3(w->hInstOff[k])), p);
        const q4 orientation = qmul(q, ldq(w->hInstQuat[k]));
        f3 lo, hi;
        host_shape_aabb(w, w->hInstShape[k], offset, orientation, lo, hi);
        if (k == w->hInstFirst[i]) { mn = lo; mx = hi; continue; }
        mn.x = fminf(mn.x, lo.x); mn.y = fminf(mn.y, lo.y); mn.z = fminf(mn.z, lo.z);
        mx.x = fmaxf(mx.x, hi.x); mx.y = fmaxf(mx.y, hi.y); mx.z = fmaxf(mx.z, hi.z);
      }
      if (shape[i] < 0) { mn = p; mx = p; }
    } else host_shape_aabb(w, shape[i], p, q, mn, mx);
    f3 he = mk3((W(mx.x) - W(mn.x)) / 2, (W(mx.y) - W(mn.y)) / 2, (W(mx.z) - W(mn.z)) / 2);
    if (shape[i] < 0) { he.x = he.y = he.z = 0.f; }
    const double m = mass[i], ex = W(he.x), ey = W(he.y), ez = W(he.z);
    f3 I;  // Box.calculateInertia, box.dart:88-94
    I.x = (float)(1.0 / 12.0 * m * (2 * ey * 2 * ey + 2 * ez * 2 * ez));
    I.y = (float)(1.0 / 12.0 * m * (2 * ex * 2 * ex + 2 * ez * 2 * ez));
    I.z = (float)(1.0 / 12.0 * m * (2 * ey * 2 * ey + 2 * ex * 2 * ex));
    const f3 iI = mk3(I.x > 0 && !fixedRot ? 1.0 / W(I.x) : 0, I.y > 0 && !fixedRot ? 1.0 / W(I.y) : 0, I.z > 0 && !fixedRot ? 1.0 / W(I.z) : 0);
    invI[i] = st3(iI);
    inertia_world(q, iI, iiw0[i], iiw1[i], iiw2[i]);  // updateInertiaWorld(true)
    // Body.updateBoundingRadius, rigid_body.dart:395-412 (zero shape offset)
    brad[i] = shape[i] >= 0 ? 0.0 + w->hShapes[shape[i]].bsr : 0.0;
    if (w->compound) {
      brad[i] = 0.0;
      for (int k = w->hInstFirst[i]; k < w->hInstFirst[i + 1]; k++) {
        const double r = vlen(ld3(w->hInstOff[k])) + w->hShapes[w->hInstShape[k]].bsr;
        if (r > brad[i]) brad[i] = r;
      }
    }
    if (brad[i] < 0) brad[i] = 0;
    if (std::isfinite(brad[i])) { radSum += brad[i]; radCnt++; }
    flags[i] = fl;
  }
  // uniform-grid parameters: bodies with a non-finite or outsized bounding radius take the "big body" path
  const double mean = radCnt ? radSum / radCnt : 0.5;
  const double bigThreshold = std::max(8.0 * mean, 1e-3);
  double rmax = 0;
  w->hBig.clear();
  for (int i = 0; i < n; i++) {
    if (!std::isfinite(brad[i]) || brad[i] > bigThreshold) { flags[i] |= BF_BIG; w->hBig.push_back(i); }
    else rmax = std::max(rmax, brad[i]);
  }
  w->nBig = (int)w->hBig.size();
  w->cell = std::max(2.0 * rmax * 1.0001, 1e-3);
  int H = 1024;
  while (H < 2 * n) H <<= 1;
  w->hashSize = H;
  w->hBigWorldStart.assign(nW + 1, 0);
  w->hWorldStart.assign(nW + 1, n);
  for (int b : w->hBig) w->hBigWorldStart[world[b] + 1]++;
  for (int k = 0; k < nW; k++) w->hBigWorldStart[k + 1] += w->hBigWorldStart[k];
  for (int i = n - 1; i >= 0; i--) w->hWorldStart[world[i]] = i;
  for (int k = nW - 1; k >= 0; k--) if (w->hWorldStart[k] > w->hWorldStart[k + 1]) w->hWorldStart[k] = w->hWorldStart[k + 1];
  w->maxWorldBodies = 0;
  for (int k = 0; k < nW; k++) w->maxWorldBodies = std::max(w->maxWorldBodies, w->hWorldStart[k + 1] - w->hWorldStart[k]);
  w->hLdamp = ldamp;
  w->hAdamp = adamp;
  w->powDt = -1;
  w->n = n;
  w->sapInit = false;

  W_TRY(w, upload(w->pos, pos, s)); W_TRY(w, upload(w->quat, quat, s)); W_TRY(w, upload(w->vel, vel, s));
  W_TRY(w, upload(w->angvel, angvel, s)); W_TRY(w, upload(w->force, force, s)); W_TRY(w, upload(w->torque, torque, s));
  { std::vector<float4> zero8(2 * zero4.size(), make_float4(0.f, 0.f, 0.f, 0.f)); W_TRY(w, upload(w->lam, zero8, s)); W_TRY(w, cudaStreamSynchronize(s)); } W_TRY(w, upload(w->aabbLo, zero4, s));
  W_TRY(w, upload(w->aabbHi, zero4, s));
  W_TRY(w, upload(w->iiw0, iiw0, s)); W_TRY(w, upload(w->iiw1, iiw1, s)); W_TRY(w, upload(w->iiw2, iiw2, s));
  W_TRY(w, upload(w->invI, invI, s)); W_TRY(w, upload(w->linF, linF, s)); W_TRY(w, upload(w->angF, angF, s));
  W_TRY(w, upload(w->mass, mass, s)); W_TRY(w, upload(w->invMass, invMass, s)); W_TRY(w, upload(w->brad, brad, s));
  W_TRY(w, upload(w->ldamp, ldamp, s)); W_TRY(w, upload(w->adamp, adamp, s)); W_TRY(w, upload(w->ldpow, ldamp, s));
  W_TRY(w, upload(w->adpow, adamp, s)); W_TRY(w, upload(w->sleepSpeed, sleepSpeed, s)); W_TRY(w, upload(w->sleepTime, sleepTime, s));
  W_TRY(w, upload(w->tLastSleepy, tLast, s));
  W_TRY(w, upload(w->type, type, s)); W_TRY(w, upload(w->sleep, sleep, s)); W_TRY(w, upload(w->shape, shape, s));
  W_TRY(w, upload(w->material, material, s)); W_TRY(w, upload(w->group, group, s)); W_TRY(w, upload(w->mask, mask, s));
  W_TRY(w, upload(w->world, world, s)); W_TRY(w, upload(w->flags, flags, s));
  W_TRY(w, upload(w->bigList, w->hBig, s)); W_TRY(w, upload(w->bigWorldStart, w->hBigWorldStart, s));
  W_TRY(w, upload(w->worldStart, w->hWorldStart, s));
  if (w->compound) {
    W_TRY(w, upload(w->dInstFirst, w->hInstFirst, s)); W_TRY(w, upload(w->dInstShape, w->hInstShape, s)); W_TRY(w, upload(w->dInstBody, w->hInstBody, s));
    W_TRY(w, upload(w->dInstOff, w->hInstOff, s)); W_TRY(w, upload(w->dInstQuat, w->hInstQuat, s));
    const size_t ni = (size_t)std::max(w->nInst, 1);
    W_TRY(w, w->pxPos.reserve(ni)); W_TRY(w, w->pxQuat.reserve(ni)); W_TRY(w, w->pxType.reserve(ni)); W_TRY(w, w->pxFlags.reserve(ni));
    W_TRY(w, w->pxMaterial.reserve(ni)); W_TRY(w, w->pxInvMass.reserve(ni));
  }
  const size_t nn = (size_t)std::max(n, 1);
  W_TRY(w, w->nbCache.reserve((size_t)nn * BP_CACHE + 4)); W_TRY(w, w->cellc.reserve(nn)); W_TRY(w, w->smeta.reserve(nn)); W_TRY(w, w->scell.reserve(nn)); W_TRY(w, w->binLo.reserve(nn));
  W_TRY(w, w->binHi.reserve(nn)); W_TRY(w, w->skey.reserve(nn)); W_TRY(w, w->sval.reserve(nn)); W_TRY(w, w->sapKey.reserve(nn));
  W_TRY(w, w->sapList.reserve(nn)); W_TRY(w, w->spos.reserve(nn)); W_TRY(w, w->srad.reserve(nn)); W_TRY(w, w->bpCounts.reserve(nn));
  W_TRY(w, w->bpOffs.reserve(nn)); W_TRY(w, w->cellStart.reserve((size_t)H + 2)); W_TRY(w, w->cellEnd.reserve((size_t)H + 2));
  int32_t rc = ensure_capacities(w);
  if (rc != CANNON_OK) return rc;
  W_TRY(w, cudaStreamSynchronize(s));
  return CANNON_OK;
}

void cannon_sph_desc_default(cannon_sph_desc* d) {
  if (!d) return;
  memset(d, 0, sizeof(*d));
  d->density = 1; d->smoothing_radius = 1; d->speed_of_sound = 1; d->viscosity = 0.01; d->eps = 0.00001;  // sph_system.dart:10-17
}

int32_t cannon_world_set_sph_systems(cannon_world* w, int32_t n, const cannon_sph_desc* sd) {
  if (w) drop_step_graph(w);
  if (!w || n < 0 || (n > 0 && !sd)) return CANNON_E_INVALID;
  cudaSetDevice(w->ctx->device);
  cudaStream_t s = w->ctx->stream;
  W_TRY(w, cudaStreamSynchronize(s));
  for (auto& hs : w->sph) { hs.particles.release(); hs.densities.release(); hs.pressures.release(); }
  w->sph.clear();
  for (int k = 0; k < n; k++) {
    if (sd[k].n_particles < 0 || (sd[k].n_particles > 0 && !sd[k].particles)) return fail(w->ctx, CANNON_E_INVALID, "SPH system needs its particle list");
    for (int i = 0; i < sd[k].n_particles; i++)
      if (sd[k].particles[i] < 0 || sd[k].particles[i] >= w->n) return fail(w->ctx, CANNON_E_INVALID, "SPH particle references unknown body");
    w->sph.emplace_back();
    auto& hs = w->sph.back();
    hs.n = sd[k].n_particles; hs.density = sd[k].density; hs.h = sd[k].smoothing_radius; hs.cs = sd[k].speed_of_sound; hs.viscosity = sd[k].viscosity; hs.eps = sd[k].eps;
    std::vector<int> pl(sd[k].particles, sd[k].particles + hs.n);
    W_TRY(w, upload(hs.particles, pl, s));
    W_TRY(w, hs.densities.reserve((size_t)std::max(hs.n, 1))); W_TRY(w, hs.pressures.reserve((size_t)std::max(hs.n, 1)));
  }
  W_TRY(w, cudaStreamSynchronize(s));
  return CANNON_OK;
}

int32_t cannon_world_set_constraints(cannon_world* w, int32_t n, const cannon_constraint_desc* cs) {
  if (w) drop_step_graph(w);  // buffers may move: the captured step is rebuilt on the next cannon_world_step
  if (!w || n < 0 || (n > 0 && !cs)) return CANNON_E_INVALID;
  cudaSetDevice(w->ctx->device);
  cudaStream_t s = w->ctx->stream;
  std::vector<int> bodyA, bodyB, kind, enabled, rowSlot, first, slotEq, mode;
  std::vector<float4> pivotA, pivotB, axisA, axisB, ni;
  std::vector<double> minF, maxF, a, b, eps, targetVel, cosv, param;
  std::vector<unsigned long long> keys;
  // trigger flags are needed for the Solver.addEquation filter (solver.dart:30-34); constructors that read body state
  // (LockConstraint, DistanceConstraint) see the bodies as they are now
  std::vector<int> flags(w->n);
  std::vector<float4> hpos(w->n), hquat(w->n);
  if (w->n) {
    W_TRY(w, cudaMemcpy(flags.data(), w->flags.p, w->n * sizeof(int), cudaMemcpyDeviceToHost));
    W_TRY(w, cudaMemcpy(hpos.data(), w->pos.p, w->n * sizeof(float4), cudaMemcpyDeviceToHost));
    W_TRY(w, cudaMemcpy(hquat.data(), w->quat.p, w->n * sizeof(float4), cudaMemcpyDeviceToHost));
  }
  std::vector<int> wake;
  w->hConFirst.clear(); w->hConType.clear(); w->hJTrig.clear();
  // Equation ctor SPOOK parameters (equation_class.dart:38): k=1e7, d=4, h=1/60 — never refreshed for joints
  const double k0 = 1e7, d0 = 4, h0 = 1.0 / 60;
  const double sa = 4.0 / (h0 * (1 + 4 * d0)), sbv = 4.0 * d0 / (1 + 4 * d0), se = 4.0 / (h0 * h0 * k0 * (1 + 4 * d0));
  const double cosHalfPi = cos(M_PI / 2);  // RotationalEquation.maxAngle default, rotational_equation.dart:17
  int slot = 0;
  for (int i = 0; i < n; i++) {
    const cannon_constraint_desc& d = cs[i];
    if (d.body_a < 0 || d.body_b < 0 || d.body_a >= w->n || d.body_b >= w->n) return fail(w->ctx, CANNON_E_INVALID, "constraint references unknown body");
    if (d.type < CANNON_CONSTRAINT_POINT_TO_POINT || d.type > CANNON_CONSTRAINT_CONE_TWIST)
      return fail(w->ctx, CANNON_E_UNSUPPORTED, "constraint type outside the hot-path scope (SURVEY.md §8f)");
    wake.push_back(d.body_a);
    wake.push_back(d.body_b);
    if (!d.collide_connected) {
      const unsigned lo = (unsigned)std::min(d.body_a, d.body_b), hi = (unsigned)std::max(d.body_a, d.body_b);
      keys.push_back(((unsigned long long)lo << 32) | hi);
    }
    const bool trig = (flags[d.body_a] & BF_IS_TRIGGER) || (flags[d.body_b] & BF_IS_TRIGGER);
    const int firstEq = (int)bodyA.size();
    // the poses the constraint's constructor saw: the current ones, or the recorded ones of a constraint made earlier
    f3 xA = ld3(hpos[d.body_a]), xB = ld3(hpos[d.body_b]);
    q4 qA = ldq(hquat[d.body_a]), qB = ldq(hquat[d.body_b]);
    if (d.has_ctor_pose) {
      xA.x = d.ctor_pos_a[0]; xA.y = d.ctor_pos_a[1]; xA.z = d.ctor_pos_a[2]; xB.x = d.ctor_pos_b[0]; xB.y = d.ctor_pos_b[1]; xB.z = d.ctor_pos_b[2];
      qA.x = d.ctor_quat_a[0]; qA.y = d.ctor_quat_a[1]; qA.z = d.ctor_quat_a[2]; qA.w = d.ctor_quat_a[3];
      qB.x = d.ctor_quat_b[0]; qB.y = d.ctor_quat_b[1]; qB.z = d.ctor_quat_b[2]; qB.w = d.ctor_quat_b[3];
    }
    f3 pvA, pvB;
    pvA.x = d.pivot_a[0]; pvA.y = d.pivot_a[1]; pvA.z = d.pivot_a[2];
    pvB.x = d.pivot_b[0]; pvB.y = d.pivot_b[1]; pvB.z = d.pivot_b[2];
    f3 axA; axA.x = d.axis_a[0]; axA.y = d.axis_a[1]; axA.z = d.axis_a[2];
    f3 axB; axB.x = d.axis_b[0]; axB.y = d.axis_b[1]; axB.z = d.axis_b[2];
    f3 zero3; zero3.x = zero3.y = zero3.z = 0.f;
    // one equation of this constraint
    auto push = [&](int kd, int md, const f3& eqAxisA, const f3& eqAxisB, int axisIdx, double lo, double hi, double cs_, double prm, int en, double tv) {
      bodyA.push_back(d.body_a);
      bodyB.push_back(d.body_b);
      first.push_back(firstEq);
      pivotA.push_back(st3(pvA));
      pivotB.push_back(st3(pvB));
      axisA.push_back(st3(eqAxisA));
      axisB.push_back(st3(eqAxisB));
      ni.push_back(make_float4(axisIdx == 0 ? 1.f : 0.f, axisIdx == 1 ? 1.f : 0.f, axisIdx == 2 ? 1.f : 0.f, 0.f));
      a.push_back(sa); b.push_back(sbv); eps.push_back(se);
      kind.push_back(kd);
      mode.push_back(md);
      enabled.push_back(en);
      minF.push_back(lo);
      maxF.push_back(hi);
      cosv.push_back(cs_);
      param.push_back(prm);
      targetVel.push_back(tv);
      if (en && !trig) { slotEq.push_back((int)rowSlot.size()); rowSlot.push_back(slot++); } else rowSlot.push_back(-1);
      w->hJTrig.push_back(trig ? 1 : 0);
    };
    w->hConFirst.push_back(firstEq);
    w->hConType.push_back(d.type);
    if (d.type == CANNON_CONSTRAINT_DISTANCE) {  // distance_constraint.dart:14-23
      double dist = d.distance;
      if (dist < 0) dist = vdist(xA, xB);
      push(ROW_CONTACT, JM_DISTANCE, zero3, zero3, -1, -d.max_force, d.max_force, 0.0, dist, 1, 0.0);
      continue;
    }
    if (d.type == CANNON_CONSTRAINT_LOCK) {
      // lock_constraint.dart:29-33: pivots = the halfway point in both local frames (Body.pointToLocalFrame)
      f3 halfWay = vadd(xA, xB);
      halfWay = vscale(0.5, halfWay);
      pvB = qrot(qconj(qB), vsub(halfWay, xB));
      pvA = qrot(qconj(qA), vsub(halfWay, xA));
    }
    for (int e = 0; e < 3; e++) push(ROW_CONTACT, JM_P2P, zero3, zero3, e, -d.max_force, d.max_force, 0.0, 0.0, 1, 0.0);
    if (d.type == CANNON_CONSTRAINT_HINGE) {
      vnormalize(axA);  // hinge_constraint.dart:34-37
      vnormalize(axB);
      push(ROW_ROT, JM_HINGE_ROT, axA, axB, -1, -d.max_force, d.max_force, cosHalfPi, 0.0, 1, 0.0);
      push(ROW_ROT, JM_HINGE_ROT, axA, axB, -1, -d.max_force, d.max_force, cosHalfPi, 0.0, 1, 0.0);
      const double mf = d.motor_max_force > 0 ? d.motor_max_force : d.max_force;
      push(ROW_MOTOR, JM_MOTOR, axA, axB, -1, -mf, mf, 0.0, 0.0, d.motor_enabled != 0, d.motor_target_velocity);
    } else if (d.type == CANNON_CONSTRAINT_LOCK) {
      // Body.vectorToLocalFrame conjugates the body's quaternion IN PLACE before rotating (rigid_body.dart:325-329 with
      // vector_math's mutating Quaternion.conjugate), so after the pivot call the frame vectors see q, q*, q in turn
      // (lock_constraint.dart:38-43); four calls per body leave the quaternion where it was.
      f3 X, Y, Z;
      X.x = 1.f; X.y = 0.f; X.z = 0.f; Y.x = 0.f; Y.y = 1.f; Y.z = 0.f; Z.x = 0.f; Z.y = 0.f; Z.z = 1.f;
      const f3 lxA = qrot(qA, X), lxB = qrot(qB, X), lyA = qrot(qconj(qA), Y), lyB = qrot(qconj(qB), Y), lzA = qrot(qA, Z), lzB = qrot(qB, Z);
      push(ROW_ROT, JM_DIRECT_ROT, lxA, lyB, -1, -d.max_force, d.max_force, cosHalfPi, 0.0, 1, 0.0);  // lock_constraint.dart:79-87
      push(ROW_ROT, JM_DIRECT_ROT, lyA, lzB, -1, -d.max_force, d.max_force, cosHalfPi, 0.0, 1, 0.0);
      push(ROW_ROT, JM_DIRECT_ROT, lzA, lxB, -1, -d.max_force, d.max_force, cosHalfPi, 0.0, 1, 0.0);
    } else if (d.type == CANNON_CONSTRAINT_CONE_TWIST) {
      // cone (cone_equation.dart:34-56 == RotationalEquation.computeB with cos(angle)) and twist; both are built with
      // maxForce 0 and then get minForce = -maxForce (cone_twist_constraint.dart:47-71)
      push(ROW_ROT, JM_DIRECT_ROT, axA, axB, -1, -d.max_force, 0.0, cos(d.angle), 0.0, 1, 0.0);
      f3 t1, t2a, t2b;  // axisA.tangents(twist.axisA, twist.axisA): the second tangent survives (:88-94)
      vtangents(axA, t1, t2a);
      vtangents(axB, t1, t2b);
      push(ROW_ROT, JM_DIRECT_ROT, t2a, t2b, -1, -d.max_force, 0.0, cos(d.twist_angle), 0.0, 1, 0.0);
    }
  }
  std::sort(keys.begin(), keys.end());
  keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
  w->nFilterKeys = (int)keys.size();
  w->nJointEq = (int)bodyA.size();
  w->nJointAccepted = slot;
  w->hJEnabled = enabled;
  W_TRY(w, upload(w->jBodyA, bodyA, s)); W_TRY(w, upload(w->jBodyB, bodyB, s)); W_TRY(w, upload(w->jKind, kind, s));
  W_TRY(w, upload(w->jEnabled, enabled, s)); W_TRY(w, upload(w->jRowSlot, rowSlot, s)); W_TRY(w, upload(w->jFirst, first, s));
  W_TRY(w, upload(w->jSlotEq, slotEq, s));
  W_TRY(w, upload(w->jPivotA, pivotA, s)); W_TRY(w, upload(w->jPivotB, pivotB, s)); W_TRY(w, upload(w->jAxisA, axisA, s));
  W_TRY(w, upload(w->jAxisB, axisB, s)); W_TRY(w, upload(w->jNi, ni, s)); W_TRY(w, upload(w->jMinF, minF, s));
  W_TRY(w, upload(w->jMaxF, maxF, s)); W_TRY(w, upload(w->jA, a, s)); W_TRY(w, upload(w->jB, b, s)); W_TRY(w, upload(w->jEps, eps, s));
  W_TRY(w, upload(w->jTargetVel, targetVel, s)); W_TRY(w, upload(w->filterKeys, keys, s));
  W_TRY(w, upload(w->jMode, mode, s)); W_TRY(w, upload(w->jCos, cosv, s)); W_TRY(w, upload(w->jParam, param, s));
  // Constraint ctor wakes both bodies (constraint_class.dart:26-29)
  if (!wake.empty()) {
    std::vector<int> sl(w->n);
    W_TRY(w, cudaMemcpy(sl.data(), w->sleep.p, w->n * sizeof(int), cudaMemcpyDeviceToHost));
    for (int bidx : wake) sl[bidx] = CANNON_AWAKE;
    W_TRY(w, cudaMemcpyAsync(w->sleep.p, sl.data(), w->n * sizeof(int), cudaMemcpyHostToDevice, s));
  }
  W_TRY(w, cudaStreamSynchronize(s));
  return ensure_capacities(w);
}

int32_t cannon_world_set_springs(cannon_world* w, int32_t n, const cannon_spring_desc* sp) {
  if (w) drop_step_graph(w);
  if (!w || n < 0 || (n > 0 && !sp)) return CANNON_E_INVALID;
  cudaSetDevice(w->ctx->device);
  cudaStream_t s = w->ctx->stream;
  std::vector<int> a(n), b(n), off(w->n + 1, 0), idx;
  std::vector<double> rest(n), k(n), d(n);
  std::vector<float4> aa(n), ab(n);
  for (int i = 0; i < n; i++) {
    if (sp[i].body_a < 0 || sp[i].body_b < 0 || sp[i].body_a >= w->n || sp[i].body_b >= w->n) return fail(w->ctx, CANNON_E_INVALID, "spring references unknown body");
    a[i] = sp[i].body_a; b[i] = sp[i].body_b;
    rest[i] = sp[i].rest_length; k[i] = sp[i].stiffness; d[i] = sp[i].damping;
    aa[i] = make_float4(sp[i].local_anchor_a[0], sp[i].local_anchor_a[1], sp[i].local_anchor_a[2], 0.f);
    ab[i] = make_float4(sp[i].local_anchor_b[0], sp[i].local_anchor_b[1], sp[i].local_anchor_b[2], 0.f);
    off[a[i] + 1]++;
    if (b[i] != a[i]) off[b[i] + 1]++;
  }
  for (int i = 0; i < w->n; i++) off[i + 1] += off[i];
  idx.resize(off[w->n]);
  std::vector<int> cur(off.begin(), off.end() - 1);
  for (int i = 0; i < n; i++) {  // ascending spring index per body
    idx[cur[a[i]]++] = i;
    if (b[i] != a[i]) idx[cur[b[i]]++] = i;
  }
  w->nSprings = n;
  W_TRY(w, upload(w->spBodyA, a, s)); W_TRY(w, upload(w->spBodyB, b, s)); W_TRY(w, upload(w->spOff, off, s)); W_TRY(w, upload(w->spIdx, idx, s));
  W_TRY(w, upload(w->spRest, rest, s)); W_TRY(w, upload(w->spK, k, s)); W_TRY(w, upload(w->spD, d, s));
  W_TRY(w, upload(w->spAnchorA, aa, s)); W_TRY(w, upload(w->spAnchorB, ab, s));
  W_TRY(w, cudaStreamSynchronize(s));
  return CANNON_OK;
}

int32_t cannon_world_set_time(cannon_world* w, double t) { if (!w) return CANNON_E_INVALID; w->time = t; return CANNON_OK; }
int32_t cannon_world_get_time(cannon_world* w, double* t, int64_t* stepnumber) {
  if (!w) return CANNON_E_INVALID;
  if (t) *t = w->time;
  if (stepnumber) *stepnumber = w->stepnumber;
  return CANNON_OK;
}
int32_t cannon_world_set_stepnumber(cannon_world* w, int64_t n) { if (!w || n < 0) return CANNON_E_INVALID; w->stepnumber = n; return CANNON_OK; }
int32_t cannon_world_set_dt(cannon_world* w, double dt) { if (!w) return CANNON_E_INVALID; w->dt = dt; return CANNON_OK; }

}  // extern "C"

// ---- stages -------------------------------------------------------------------------------------------
// the host owns World.time / stepnumber between calls; every entry point that runs step kernels pushes them first
static cudaError_t sync_clock(cannon_world* w) {
  cudaError_t e = w->dClock.reserve(2);
  if (e != cudaSuccess) return e;
  memcpy(&w->hClock[0], &w->time, sizeof(double));
  w->hClock[1] = w->stepnumber;
  return cudaMemcpyAsync(w->dClock.p, w->hClock, sizeof w->hClock, cudaMemcpyHostToDevice, w->ctx->stream);
}

static StepParams step_params(cannon_world* w, double dt) {
  StepParams P;
  P.dt = dt; P.clk = w->dClock.p; P.quatSkip = w->desc.quat_normalize_skip;
  P.gx = W(w->desc.gravity[0]); P.gy = W(w->desc.gravity[1]); P.gz = W(w->desc.gravity[2]);
  P.n = w->n;
  P.allowSleep = w->desc.allow_sleep;
  P.quatNormalizeFast = w->desc.quat_normalize_fast;
  P.needAABB = (w->desc.use_bounding_boxes || w->desc.broadphase_kind != CANNON_BP_NAIVE) ? 1 : 0;
  P.nWorlds = w->desc.n_worlds;
  P.deferSleepTick = w->nSprings > 0 ? 1 : 0;
  return P;
}

static int32_t st_reset_counters(cannon_world* w) {
  W_TRY(w, cudaMemsetAsync(w->cnt.p, 0, CT_COUNT * sizeof(int), w->ctx->stream));
  return CANNON_OK;
}

static int32_t st_prestep(cannon_world* w, double dt, int doGravity, int forceAABB) {
  StepParams P = step_params(w, dt);
  if (forceAABB) P.needAABB = 1;
  if (!doGravity && !P.needAABB) return CANNON_OK;
  { g_kernel_launches++; k_prestep<<<grid_for(w, w->n, 256), 256, 0, w->ctx->stream>>>(body_arrays(w), shape_tables(w), P, doGravity); }
  if (doGravity) {  // World.subsystems update between gravity and the broadphase (world_class.dart:472-475)
    for (auto& hs : w->sph) {
      if (hs.n == 0) continue;
      SphDev S;
      S.particles = hs.particles.p; S.densities = hs.densities.p; S.pressures = hs.pressures.p; S.n = hs.n;
      S.density = hs.density; S.h = hs.h; S.h9 = pow(hs.h, 9); S.cs = hs.cs; S.viscosity = hs.viscosity; S.eps = hs.eps;
      { g_kernel_launches++; k_sph_density<<<grid_for(w, hs.n, 128), 128, 0, w->ctx->stream>>>(body_arrays(w), S); }
      { g_kernel_launches++; k_sph_forces<<<grid_for(w, hs.n, 128), 128, 0, w->ctx->stream>>>(body_arrays(w), S); }
    }
  }
  W_TRY(w, cudaGetLastError());
  return CANNON_OK;
}

static BpParams bp_params(cannon_world* w) {
  BpParams P;
  memset(&P, 0, sizeof P);
  const cannon_world_desc& d = w->desc;
  P.n = w->n; P.kind = d.broadphase_kind; P.useBoxes = d.use_bounding_boxes; P.nWorlds = d.n_worlds;
  P.hashMask = w->hashSize - 1; P.cell = w->cell; P.nBig = w->nBig;
  P.gnx = d.grid_nx; P.gny = d.grid_ny; P.gnz = d.grid_nz;
  // grid_broadphase.dart:72-86
  const double xmax = W(d.grid_max[0]), ymax = W(d.grid_max[1]), zmax = W(d.grid_max[2]);
  const double xmin = W(d.grid_min[0]), ymin = W(d.grid_min[1]), zmin = W(d.grid_min[2]);
  P.gxmin = xmin; P.gymin = ymin; P.gzmin = zmin;
  P.gxmult = d.grid_nx / (xmax - xmin); P.gymult = d.grid_ny / (ymax - ymin); P.gzmult = d.grid_nz / (zmax - zmin);
  P.gbx = (xmax - xmin) / d.grid_nx; P.gby = (ymax - ymin) / d.grid_ny; P.gbz = (zmax - zmin) / d.grid_nz;
  P.gBinRadius = sqrt(P.gbx * P.gbx + P.gby * P.gby + P.gbz * P.gbz) * 0.5;
  P.sapAxis = d.sap_axis;
  return P;
}
static BpArrays bp_arrays(cannon_world* w) {
  BpArrays A;
  A.nbCache = getenv("CANNON_BP_NO_CACHE") ? nullptr : w->nbCache.p;
  A.cellc = w->cellc.p; A.binLo = w->binLo.p; A.binHi = w->binHi.p; A.skey = w->skey.p; A.sval = w->sval.p;
  A.cellStart = w->cellStart.p; A.cellEnd = w->cellEnd.p; A.spos = w->spos.p; A.srad = w->srad.p; A.smeta = w->smeta.p; A.scell = w->scell.p;
  A.bigList = w->bigList.p; A.bigWorldStart = w->bigWorldStart.p; A.worldStart = w->worldStart.p; A.counts = w->bpCounts.p; A.offs = w->bpOffs.p;
  A.sapKey = w->sapKey.p; A.sapList = w->sapList.p;
  return A;
}

__global__ void k_iota(uint32_t* p, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = (uint32_t)i;
}
__global__ void k_clamp_count(int* cntp, int cap) { if (threadIdx.x == 0 && blockIdx.x == 0 && *cntp > cap) *cntp = cap; }

// Broadphase.collisionPairs + constraint filter; leaves pairs in w->p1/p2 and the count in cnt[CT_NPAIRS]
static int32_t st_broadphase(cannon_world* w) {
  cudaStream_t s = w->ctx->stream;
  const int n = w->n;
  BodyArrays B = body_arrays(w);
  BpParams P = bp_params(w);
  BpArrays A = bp_arrays(w);
  int* cnt = w->cnt.p;
  if (n == 0) return CANNON_OK;
  if (P.kind == CANNON_BP_SAP) {
    if (!w->sapInit) {  // SAPBroadphase.setWorld: axisList = bodies in insertion order (sap_broadphase.dart:114-135)
      { g_kernel_launches++; k_iota<<<grid_for(w, n, 256), 256, 0, s>>>(A.sapList, n); }
      w->sapInit = true;
    }
    // sortList (:168-189): stable sort of the persistent list by aabb.lowerBound[axis] == insertion sort result
    { g_kernel_launches++; k_sap_keys<<<grid_for(w, n, 256), 256, 0, s>>>(B, P, A); }
    W_TRY(w, radix_sort_pairs(A.sapKey, A.sapList, n, 32, w->sortTmp, s));
    const int gw = grid_for(w, (long long)n * 32, 256);
    { g_kernel_launches++; k_sap_sweep<<<gw, 256, 0, s>>>(B, P, A, 0, nullptr, nullptr, 0, nullptr); }
    W_TRY(w, scan_exclusive(A.counts, A.offs, nullptr, n, n, cnt + CT_NPAIRS, w->scanTmp, s));
    { g_kernel_launches++; k_sap_sweep<<<gw, 256, 0, s>>>(B, P, A, 1, w->p1.p, w->p2.p, w->pairCap, cnt + CT_OVF_PAIRS); }
  } else if (P.kind == CANNON_BP_NAIVE && w->desc.n_worlds > 1 && w->maxWorldBodies <= 512 && !getenv("CANNON_BP_NO_WORLD_KERNEL")) {
    // batches of small worlds: all pairs inside every world, no grid (k_broadphase.cuh, k_bp_world_all)
    const int g = grid_for(w, 32LL * n, 128);
    { g_kernel_launches++; k_bp_world_all<<<g, 128, 0, s>>>(B, P, A, n, 0, nullptr, nullptr, 0, nullptr); }
    W_TRY(w, scan_exclusive(A.counts, A.offs, nullptr, n, n, cnt + CT_NPAIRS, w->scanTmp, s));
    { g_kernel_launches++; k_bp_world_all<<<g, 128, 0, s>>>(B, P, A, n, 1, w->p1.p, w->p2.p, w->pairCap, cnt + CT_OVF_PAIRS); }
  } else {
    { g_kernel_launches++; k_bp_cells<<<grid_for(w, n, 256), 256, 0, s>>>(B, shape_tables(w), P, A); }
    int bits = 1;
    while ((1 << bits) <= w->hashSize) bits++;
    W_TRY(w, radix_sort_pairs(A.skey, A.sval, n, bits, w->sortTmp, s));
    W_TRY(w, cudaMemsetAsync(A.cellStart, 0, ((size_t)w->hashSize + 2) * sizeof(int), s));
    W_TRY(w, cudaMemsetAsync(A.cellEnd, 0, ((size_t)w->hashSize + 2) * sizeof(int), s));
    const int nSmall = n - w->nBig;
    if (nSmall > 0) {
      { g_kernel_launches++; k_bp_ranges<<<grid_for(w, nSmall, 256), 256, 0, s>>>(P, A); }
      { g_kernel_launches++; k_bp_reorder<<<grid_for(w, nSmall, 256), 256, 0, s>>>(B, P, A); }
      { g_kernel_launches++; k_bp_small<<<grid_for(w, nSmall, 128), 128, 0, s>>>(B, P, A, 0, nullptr, nullptr, 0, nullptr); }
    }
    if (w->nBig > 0) { g_kernel_launches++; k_bp_big<<<std::min(w->nBig, w->ctx->sms * 8), 256, 0, s>>>(B, P, A, 0, nullptr, nullptr, 0, nullptr); }
    W_TRY(w, scan_exclusive(A.counts, A.offs, nullptr, n, n, cnt + CT_NPAIRS, w->scanTmp, s));
    if (nSmall > 0) { g_kernel_launches++; k_bp_small<<<grid_for(w, nSmall, 128), 128, 0, s>>>(B, P, A, 1, w->p1.p, w->p2.p, w->pairCap, cnt + CT_OVF_PAIRS); }
    if (w->nBig > 0) { g_kernel_launches++; k_bp_big<<<std::min(w->nBig, w->ctx->sms * 8), 256, 0, s>>>(B, P, A, 1, w->p1.p, w->p2.p, w->pairCap, cnt + CT_OVF_PAIRS); }
  }
  { g_kernel_launches++; k_clamp_count<<<1, 32, 0, s>>>(cnt + CT_NPAIRS, w->pairCap); }
  if (w->nFilterKeys > 0) {  // world_class.dart:488-499
    const int g = grid_for(w, w->pairCap, 256);
    { g_kernel_launches++; k_pair_filter_flags<<<g, 256, 0, s>>>(w->p1.p, w->p2.p, cnt + CT_NPAIRS, w->pairCap, w->filterKeys.p, w->nFilterKeys, w->keep.p); }
    W_TRY(w, scan_exclusive(w->keep.p, w->keepOff.p, cnt + CT_NPAIRS, 0, w->pairCap, cnt + CT_NPAIRS_RAW, w->scanTmp, s));
    { g_kernel_launches++; k_pair_filter_compact<<<g, 256, 0, s>>>(w->p1.p, w->p2.p, cnt + CT_NPAIRS, w->pairCap, w->keep.p, w->keepOff.p, w->q1.p, w->q2.p); }
    W_TRY(w, cudaMemcpyAsync(cnt + CT_NPAIRS, cnt + CT_NPAIRS_RAW, sizeof(int), cudaMemcpyDeviceToDevice, s));
    std::swap(w->p1, w->q1);
    std::swap(w->p2, w->q2);
  }
  W_TRY(w, cudaGetLastError());
  return CANNON_OK;
}

static NpArrays np_arrays(cannon_world* w) {
  NpArrays A;
  int* cnt = w->cnt.p;
  A.p1 = w->p1.p; A.p2 = w->p2.p; A.nPairs = cnt + CT_NPAIRS;
  if (w->compound) { A.p1 = w->pp1.p; A.p2 = w->pp2.p; A.nPairs = cnt + CT_NPPAIRS; }
  A.pairTasks = w->pairTasks.p; A.pairTaskOff = w->pairTaskOff.p; A.pairMask = w->pairMask.p; A.nTasks = cnt + CT_NTASKS;
  A.taskPair = w->taskPair.p; A.taskInfo = w->taskInfo.p; A.taskCell = w->taskCell.p;
  A.bucket = w->bucket.p; A.bucketCount = cnt + CT_BUCKETCOUNT; A.bucketStart = cnt + CT_BUCKETSTART; A.bucketCursor = cnt + CT_BUCKETCURSOR;
  A.taskCnt = w->taskCnt.p; A.taskRaw = w->taskRaw.p; A.taskOff = w->taskOff.p; A.rawCount = cnt + CT_RAWCOUNT;
  A.rawRi = w->rawRi.p; A.rawRj = w->rawRj.p; A.rawNi = w->rawNi.p;
  A.taskCap = w->taskCap; A.contactCap = w->contactCap;
  A.overflowTasks = cnt + CT_OVF_TASKS; A.overflowContacts = cnt + CT_OVF_CONTACTS; A.unsupported = cnt + CT_UNSUPPORTED;
  A.taskHit = (w->evEnabled && w->evCap > 0) ? w->taskHit.p : nullptr;
  { const char* e = getenv("CANNON_NP_DEBUG"); A.debug = e ? atoi(e) : 0; }
  A.clipList = w->clipList.p; A.nClip = cnt + CT_NCLIP0; A.taskSep = w->taskSep.p;
  return A;
}
static ContactArrays contact_arrays(cannon_world* w) {
  ContactArrays C;
  C.nContacts = w->cnt.p + CT_NCONTACTS;
  C.bi = w->cBi.p; C.bj = w->cBj.p; C.ri = w->cRi.p; C.rj = w->cRj.p; C.ni = w->cNi.p;
  C.rest = w->cRest.p; C.mu = w->cMu.p; C.slip = w->cSlip.p; C.ca = w->cCa.p; C.cb = w->cCb.p; C.ceps = w->cCeps.p;
  C.fb = w->cFb.p; C.feps = w->cFeps.p; C.enabled = w->cEnabled.p; C.row = w->cRow.p; C.task = w->cTask.p;
  return C;
}

// Narrowphase.getContacts over the device pair list
static int32_t st_narrowphase(cannon_world* w, double dt) {
  cudaStream_t s = w->ctx->stream;
  BodyArrays B = np_body_arrays(w);
  ShapeTables T = shape_tables(w);
  NpArrays A = np_arrays(w);
  ContactArrays C = contact_arrays(w);
  int* cnt = w->cnt.p;
  const int npPairCap = w->compound ? w->ppCap : w->pairCap;
  if (w->compound) {  // shape instances at their world poses, body pairs -> instance pairs (narrow_phase.dart:669-680)
    ProxyArrays X;
    X.instBody = w->dInstBody.p; X.pos = w->pxPos.p; X.quat = w->pxQuat.p; X.type = w->pxType.p; X.flags = w->pxFlags.p;
    X.material = w->pxMaterial.p; X.invMass = w->pxInvMass.p; X.nInst = w->nInst;
    { g_kernel_launches++; k_proxies<<<grid_for(w, w->nInst, 256), 256, 0, s>>>(body_arrays(w), T, X); }
    const int gb = grid_for(w, w->pairCap, 256);
    { g_kernel_launches++; k_pp_count<<<gb, 256, 0, s>>>(w->p1.p, w->p2.p, cnt + CT_NPAIRS, w->pairCap, w->dInstFirst.p, w->ppCnt.p); }
    W_TRY(w, scan_exclusive(w->ppCnt.p, w->ppOff.p, nullptr, w->pairCap, w->pairCap, cnt + CT_NPPAIRS, w->scanTmp, s));
    { g_kernel_launches++; k_pp_fill<<<gb, 256, 0, s>>>(w->p1.p, w->p2.p, cnt + CT_NPAIRS, w->pairCap, w->dInstFirst.p, w->ppOff.p, w->pp1.p, w->pp2.p, w->ppCap, cnt + CT_OVF_PAIRS); }
    { g_kernel_launches++; k_clamp_count<<<1, 32, 0, s>>>(cnt + CT_NPPAIRS, w->ppCap); }
  }
  const int gp = grid_for(w, npPairCap, 128);
  const int hoistBytes = (int)(4 * 32 * QS_HOIST_AXES * sizeof(QsAxis));  // pass 0 only
  { g_kernel_launches++; k_np_tasks<<<gp, 128, hoistBytes, s>>>(B, T, A, 0, hoistBytes); }
  W_TRY(w, scan_exclusive(A.pairTasks, A.pairTaskOff, A.nPairs, 0, npPairCap, cnt + CT_NTASKS, w->scanTmp, s));
  { g_kernel_launches++; k_bucket_starts<<<1, 32, 0, s>>>(cnt); }
  { g_kernel_launches++; k_np_tasks<<<gp, 128, 0, s>>>(B, T, A, 1, 0); }
  const int g = w->ctx->sms * 8;
  // fork: the heavy SAT kernels go to side streams, the cheap analytic resolvers stay on the main stream
  cudaStream_t s1 = w->npStream[0], s2 = w->npStream[1], s3 = w->npStream[2];
  W_TRY(w, cudaEventRecord(w->npFork, s));
  const bool hf = !w->hHfs.empty();
  if (hf) {
    W_TRY(w, cudaStreamWaitEvent(s2, w->npFork, 0));
    { g_kernel_launches++; k_np_hull_warp<true, 0><<<g * 2, SAT_TILES * SAT_GROUP, 0, s2>>>(B, T, A, cnt + CT_OVF_CLIP); }
    { g_kernel_launches++; k_np_hull_warp<true, 1><<<g * 2, SAT_TILES * SAT_GROUP, 0, s2>>>(B, T, A, cnt + CT_OVF_CLIP); }
    if (w->hasOversizeHull) { g_kernel_launches++; k_np_hull_pillar<<<g * 2, 64, 0, s2>>>(B, T, A, cnt + CT_OVF_CLIP, 1); }
    W_TRY(w, cudaEventRecord(w->npJoin[1], s2));
  }
  W_TRY(w, cudaStreamWaitEvent(s1, w->npFork, 0));
  { g_kernel_launches++; k_np_hull_warp<false, 0><<<g * 2, SAT_TILES * SAT_GROUP, 0, s1>>>(B, T, A, cnt + CT_OVF_CLIP); }
  { g_kernel_launches++; k_np_hull_warp<false, 1><<<g * 2, SAT_TILES * SAT_GROUP, 0, s1>>>(B, T, A, cnt + CT_OVF_CLIP); }
  if (w->hasOversizeHull) { g_kernel_launches++; k_np_hull_hull<<<g * 2, 64, 0, s1>>>(B, T, A, cnt + CT_OVF_CLIP, 1); }
  W_TRY(w, cudaEventRecord(w->npJoin[0], s1));
  if (hf) {
    W_TRY(w, cudaStreamWaitEvent(s3, w->npFork, 0));
    { g_kernel_launches++; k_np_sphere_pillar<<<g * 2, 64, 0, s3>>>(B, T, A); }
    W_TRY(w, cudaEventRecord(w->npJoin[2], s3));
  }
  { g_kernel_launches++; k_np_sphere_sphere<<<g, 256, 0, s>>>(B, T, A); }
  { g_kernel_launches++; k_np_sphere_plane<<<g, 256, 0, s>>>(B, T, A); }
  { g_kernel_launches++; k_np_sphere_box<<<g, 128, 0, s>>>(B, T, A); }
  { g_kernel_launches++; k_np_sphere_hull<<<g, 128, 0, s>>>(B, T, A); }
  { g_kernel_launches++; k_np_plane_hull<<<g, 128, 0, s>>>(B, T, A); }
  if (w->hasTrimesh) { g_kernel_launches++; k_np_trimesh<<<g, 128, 0, s>>>(B, T, A); }
  if (w->hasParticle) {
    { g_kernel_launches++; k_np_particle_simple<<<g, 128, 0, s>>>(B, T, A); }
    { g_kernel_launches++; k_np_particle_hull<0><<<g, 128, 0, s>>>(B, T, A); }
    { g_kernel_launches++; k_np_particle_hull<1><<<g, 128, 0, s>>>(B, T, A); }
    { g_kernel_launches++; k_np_particle_hull<2><<<g, 128, 0, s>>>(B, T, A); }
  }
  W_TRY(w, cudaStreamWaitEvent(s, w->npJoin[0], 0));
  if (hf) {
    W_TRY(w, cudaStreamWaitEvent(s, w->npJoin[1], 0));
    W_TRY(w, cudaStreamWaitEvent(s, w->npJoin[2], 0));
  }
  // a task count above taskCap stays visible to the host as an overflow
  { g_kernel_launches++; k_clamp_count<<<1, 32, 0, s>>>(cnt + CT_NTASKS, w->taskCap + 1); }
  W_TRY(w, scan_exclusive(A.taskCnt, A.taskOff, cnt + CT_NTASKS, 0, w->taskCap, cnt + CT_NCONTACTS, w->scanTmp, s));
  NpWorld Wd;
  Wd.dt = dt;
  {
    f3 g3;
    if (w->desc.has_friction_gravity) { g3.x = w->desc.friction_gravity[0]; g3.y = w->desc.friction_gravity[1]; g3.z = w->desc.friction_gravity[2]; }
    else { g3.x = w->desc.gravity[0]; g3.y = w->desc.gravity[1]; g3.z = w->desc.gravity[2]; }
    Wd.gnorm = vlen(g3);  // narrow_phase.dart:550
  }
  Wd.defaultCm = w->desc.default_contact_material;
  { g_kernel_launches++; k_np_finalize<<<grid_for(w, w->taskCap, 256), 256, 0, s>>>(B, T, A, C, Wd); }
  W_TRY(w, cudaGetLastError());
  return CANNON_OK;
}

static RowArrays row_arrays(cannon_world* w) {
  RowArrays R;
  R.nRows = w->cnt.p + CT_NROWS;
  R.kind = w->rKind.p; R.n = w->rN.p; R.rA = w->rRA.p; R.rB = w->rRB.p; R.iA = w->rIA.p; R.iB = w->rIB.p;
  R.B = w->rB.p; R.invC = w->rInvC.p; R.eps = w->rEps.p; R.minF = w->rMinF.p; R.maxF = w->rMaxF.p; R.lambda = w->rLambda.p;
  R.rowCap = w->rowCap;
  R.fast = kind_fast(w) ? 1 : 0;
  R.rec = w->rRec.p; R.flambda = w->rFlambda.p;
  R.xblk = kind_packed(w) ? w->rXblk.p : nullptr;
  return R;
}
static UnitArrays unit_arrays(cannon_world* w) {
  UnitArrays U;
  U.nUnits = w->cnt.p + CT_NUNITS; U.nExec = w->cnt.p + CT_NEXEC;
  U.uBi = w->uBi.p; U.uBj = w->uBj.p; U.uFlags = w->uFlags.p; U.uRows = w->uRows.p; U.uSrc = w->uSrc.p; U.uKey = w->uKey.p; U.uPri = w->uPri.p;
  U.eBi = w->eBi.p; U.eBj = w->eBj.p; U.eFlags = w->eFlags.p; U.eRowBase = w->eRowBase.p; U.eImA = w->eImA.p; U.eImB = w->eImB.p; U.rec = w->uRec.p; U.eLevel = w->eLevel.p;
  U.eRows = w->eRows.p; U.unitRow = w->unitRow.p; U.unitCap = w->unitCap;
  const bool packed = kind_packed(w);
  U.xrec = packed ? w->uXrec.p : nullptr; U.unitSeq = nullptr; U.bodyCnt = nullptr;
  U.winBase = w->gxWinBase.p; U.lvlWin = w->gsLvlTask.p; U.levelStart = w->levelStart.p; U.padTotal = w->cnt.p + CT_NROWS_PAD;
  return U;
}
static JointArrays joint_arrays(cannon_world* w) {
  JointArrays J;
  J.n = w->nJointEq;
  J.bodyA = w->jBodyA.p; J.bodyB = w->jBodyB.p; J.kind = w->jKind.p; J.enabled = w->jEnabled.p; J.rowSlot = w->jRowSlot.p;
  J.slotEq = w->jSlotEq.p;
  J.pivotA = w->jPivotA.p; J.pivotB = w->jPivotB.p; J.axisA = w->jAxisA.p; J.axisB = w->jAxisB.p; J.ni = w->jNi.p;
  J.minF = w->jMinF.p; J.maxF = w->jMaxF.p; J.a = w->jA.p; J.b = w->jB.p; J.eps = w->jEps.p; J.targetVel = w->jTargetVel.p;
  J.first = w->jFirst.p; J.nAccepted = w->nJointAccepted;
  J.mode = w->jMode.p; J.cosv = w->jCos.p; J.param = w->jParam.p;
  return J;
}

__global__ void k_units_plus_one(int* cnt) { if (threadIdx.x == 0 && blockIdx.x == 0) cnt[CT_NUNITS1] = cnt[CT_NEXEC] + 1; }
__global__ void __launch_bounds__(256) k_zero_tail(int* eRows, const int* nUnits, int cap) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { const int n = *nUnits; if (n <= cap) eRows[n] = 0; }
}

static EvArrays ev_arrays(cannon_world* w) {
  EvArrays E;
  E.keysCur = w->evKeysCur.p; E.keysPrev = w->evKeysPrev.p; E.tabCur = w->evTabCur.p; E.tabPrev = w->evTabPrev.p;
  E.begin = w->evBegin.p; E.end = w->evEnd.p; E.cnt = w->evCnt.p; E.mask = w->evMask; E.cap = w->evCap;
  return E;
}
// queue of the tile SAT kernel's clipping launch: only worlds with hull shapes (box / convex / cylinder) need it
static int32_t ensure_clip_buffers(cannon_world* w) {
  if (w->hHulls.empty() || w->taskCap <= 0) return CANNON_OK;
  if (w->clipList.cap >= 2 * (size_t)w->taskCap && w->taskSep.cap >= (size_t)w->taskCap) return CANNON_OK;
  drop_step_graph(w);
  W_TRY(w, w->clipList.reserve(2 * (size_t)w->taskCap));
  W_TRY(w, w->taskSep.reserve((size_t)w->taskCap));
  return CANNON_OK;
}
// (re)allocates the event buffers for the current pair capacity and starts from an empty previous set
static int32_t ensure_events(cannon_world* w) {
  if (!w->evEnabled || (w->evCap == w->pairCap && w->evCap > 0)) return CANNON_OK;
  if (w->pairCap <= 0) return CANNON_OK;  // no bodies yet: allocated on the first step
  drop_step_graph(w);
  cudaStream_t s = w->ctx->stream;
  const size_t cap = (size_t)w->pairCap;
  size_t tab = 1024;
  while (tab < 2 * cap) tab <<= 1;
  W_TRY(w, w->evKeysCur.reserve(cap)); W_TRY(w, w->evKeysPrev.reserve(cap)); W_TRY(w, w->evBegin.reserve(cap)); W_TRY(w, w->evEnd.reserve(cap));
  W_TRY(w, w->evTabCur.reserve(tab)); W_TRY(w, w->evTabPrev.reserve(tab)); W_TRY(w, w->evCnt.reserve(8));
  W_TRY(w, w->taskHit.reserve((size_t)std::max(w->taskCap, 1)));
  w->evCap = (int)cap; w->evMask = (unsigned)(tab - 1);
  W_TRY(w, cudaMemsetAsync(w->evTabPrev.p, 0xff, tab * sizeof(unsigned long long), s));
  W_TRY(w, cudaMemsetAsync(w->evCnt.p, 0, 8 * sizeof(int), s));
  return CANNON_OK;
}
// World.emitContactEvents (world_class.dart:610,703-730) from the pairs that own contacts after the narrowphase
static int32_t st_contact_events(cannon_world* w) {
  if (!w->evEnabled || w->evCap <= 0) return CANNON_OK;
  cudaStream_t s = w->ctx->stream;
  const EvArrays E = ev_arrays(w);
  const size_t tabBytes = ((size_t)w->evMask + 1) * sizeof(unsigned long long);
  W_TRY(w, cudaMemsetAsync(E.tabCur, 0xff, tabBytes, s));
  { g_kernel_launches++; k_ev_begin<<<1, 32, 0, s>>>(E); }
  { g_kernel_launches++; k_ev_collect<<<grid_for(w, w->compound ? w->ppCap : w->pairCap, 256), 256, 0, s>>>(np_arrays(w), E, w->compound ? w->dInstBody.p : nullptr); }
  { g_kernel_launches++; k_ev_diff<<<grid_for(w, 2LL * w->pairCap, 256), 256, 0, s>>>(E); }
  { g_kernel_launches++; k_ev_roll<<<grid_for(w, w->pairCap, 256), 256, 0, s>>>(E); }
  W_TRY(w, cudaMemcpyAsync(E.tabPrev, E.tabCur, tabBytes, cudaMemcpyDeviceToDevice, s));
  return CANNON_OK;
}

// world_class.dart:539-645 without the final velocity update (k_integrate / k_apply_lambda do that)
static int32_t st_solve(cannon_world* w, double dt) {
  cudaStream_t s = w->ctx->stream;
  BodyArrays B = body_arrays(w);
  ContactArrays C = contact_arrays(w);
  RowArrays R = row_arrays(w);
  UnitArrays U = unit_arrays(w);
  JointArrays J = joint_arrays(w);
  int* cnt = w->cnt.p;
  const int nW = w->desc.n_worlds;
  { const int32_t rcEv = st_contact_events(w); if (rcEv != CANNON_OK) return rcEv; }
  SolveParams P;
  P.dt = dt; P.tol2 = w->desc.solver_tolerance * w->desc.solver_tolerance; P.maxIter = w->desc.solver_iterations;
  P.nBodies = w->n; P.nWorlds = nW; P.colored = kind_colored(w);
  P.debugSkipWork = getenv("CANNON_DEBUG_SKIP_GS_WORK") ? 1 : 0;
  P.trace = nullptr;
  const char* tracePath = getenv("CANNON_GS_TRACE");
  const size_t traceWords = (size_t)w->ctx->sms * 4 * GS_TRACE_PHASES * 2 + (size_t)w->ctx->sms * 4 * 16;
  if (tracePath) {
    if (w->gsTrace.reserve(traceWords) != cudaSuccess) return fail(w->ctx, CANNON_E_CUDA, "trace buffer");
    W_TRY(w, cudaMemsetAsync(w->gsTrace.p, 0, traceWords * sizeof(long long), s));
    P.trace = w->gsTrace.p;
  }
  const int gc = grid_for(w, w->contactCap, 256);
  { g_kernel_launches++; k_contact_flags<<<gc, 256, 0, s>>>(B, C, w->contactCap, w->fricFlag.p, w->contFlag.p, w->hasShapeMaterial ? w->dMatRestitution.p : nullptr); }
  W_TRY(w, scan_exclusive(w->fricFlag.p, w->fricOff.p, cnt + CT_NCONTACTS, 0, w->contactCap, cnt + CT_FRICTOTAL, w->scanTmp, s));
  W_TRY(w, scan_exclusive(w->contFlag.p, w->contOff.p, cnt + CT_NCONTACTS, 0, w->contactCap, cnt + CT_CONTTOTAL, w->scanTmp, s));
  { g_kernel_launches++; k_presolve<<<grid_for(w, w->n, 256), 256, 0, s>>>(B, w->n); }
  const bool split = w->desc.solver_kind == CANNON_SOLVER_SPLIT;
  const int nGroups = split ? w->n : nW;
  W_TRY(w, cudaMemsetAsync(w->worldRows.p, 0, (nW + 1) * sizeof(int), s));
  W_TRY(w, cudaMemsetAsync(w->worldDone.p, 0, (nGroups + 2) * sizeof(int), s));
  W_TRY(w, cudaMemsetAsync(w->worldIters.p, 0, (nGroups + 1) * sizeof(int), s));
  W_TRY(w, cudaMemsetAsync(w->worldTot.p, 0, (nGroups + 3) * sizeof(double), s));
  UnitSrc Us;
  Us.colored = P.colored; Us.split = split ? 1 : 0; Us.fricFlag = w->fricFlag.p; Us.contFlag = w->contFlag.p; Us.fricOff = w->fricOff.p; Us.contOff = w->contOff.p;
  Us.fricTotal = cnt + CT_FRICTOTAL; Us.contTotal = cnt + CT_CONTTOTAL; Us.taskOff = w->taskOff.p; Us.taskCnt = w->taskCnt.p;
  Us.nTasks = cnt + CT_NTASKS; Us.taskCap = w->taskCap; Us.contactCap = w->contactCap;
  if (P.colored && nW > 1) {  // world-local keys for the colouring of a batch
    { g_kernel_launches++; k_world_keys_init<<<grid_for(w, WK_ARRAYS * nW, 256), 256, 0, s>>>(w->worldKeys.p, nW); }
    { g_kernel_launches++; k_world_keys<<<grid_for(w, w->taskCap, 256), 256, 0, s>>>(B, C, Us, J, nW, w->worldKeys.p); }
  }
  { g_kernel_launches++; k_units_build<<<grid_for(w, w->unitCap, 256), 256, 0, s>>>(B, C, Us, J, U, nW, w->worldRows.p, cnt + CT_OVF_ROWS, w->worldKeys.p); }
  // dependency levels + sweeps: persistent cooperative kernels
  SchedArrays S;
  S.claim = w->claim.p; S.unitLevel = w->unitLevel.p; S.order = w->order.p; S.levelStart = w->levelStart.p; S.nLevels = cnt + CT_NLEVELS;
  S.act0 = w->act0.p; S.act1 = w->act1.p; S.actCount = cnt + CT_ACT0; S.cursor = cnt + CT_CURSOR; S.bar = (unsigned*)(cnt + CT_BAR);
  S.maxLevels = w->maxLevels; S.levelOverflow = cnt + CT_OVF_LEVELS;
  const bool packed = kind_packed(w);
  S.unitSeq = nullptr; S.bodyCnt = nullptr;
  W_TRY(w, cudaMemsetAsync(w->claim.p, 0xff, ((size_t)w->n + 1) * sizeof(unsigned long long), s));
  if (w->recordSolveEvents) cudaEventRecord(w->ev[5], s);
  // a COLORED batch of small worlds is coloured world by world (one warp each, no grid barrier)
  const bool worldSched = P.colored && nW > 1 && w->maxWorldBodies <= SW_MAXB && !split && !kind_fast(w) && !w->gxOff && !getenv("CANNON_NO_WORLD_SCHEDULE");
  if (worldSched) {
    W_TRY(w, cudaMemsetAsync(cnt + CT_NEXEC, 0, sizeof(int), s));
    W_TRY(w, cudaMemsetAsync(cnt + CT_NLEVELS, 0, sizeof(int), s));
    g_kernel_launches++;
    k_schedule_worlds<<<nW, 32, 0, s>>>(B, U, S, w->worldKeys.p, w->worldStart.p, nW, 0, cnt + CT_NCONTACTS, w->contactCap);
  } else {
    int colored = P.colored;
    void* args[] = {&U, &S, &colored};
    g_kernel_launches++;
    W_TRY(w, cudaLaunchCooperativeKernel((void*)k_schedule, dim3(w->coopBlocksSched), dim3(256), args, 0, s));
  }
  if (w->recordSolveEvents) cudaEventRecord(w->ev[6], s);
  W_TRY(w, cudaMemsetAsync(cnt + CT_BAR, 0, 64 * sizeof(int), s));
  if (split) {  // island labels for the per-island tolerance exits of SplitSolver
    int nb = w->n;
    int* label = w->islandLabel.p;
    int* changed = cnt + CT_ISL_CHANGED0;
    int* nIsl = cnt + CT_NISLANDS;
    unsigned* bar = S.bar;
    void* args[] = {&B, &U, &nb, &label, &changed, &nIsl, &bar};
    g_kernel_launches++;
    W_TRY(w, cudaLaunchCooperativeKernel((void*)k_islands, dim3(w->coopBlocksSched), dim3(256), args, 0, s));
    W_TRY(w, cudaMemsetAsync(cnt + CT_BAR, 0, 64 * sizeof(int), s));
  }
  // a colored batch is swept world by world (k_gs_world): regroup the execution order by world first
  // (measured: the same per-world scheme for the f64 reference-order rows is slower than the grid-wide k_gs, 4.7 vs 4.2 ms
  // on c4 — its phases are two dependent global loads long and only 15 worlds fit an SM — so it is not in the tree)
  const bool fast = kind_fast(w);
  const bool perWorld = fast && nW > 1 && !w->gsFastV1 && !getenv("CANNON_GS_NO_WORLD_KERNEL");
  // exact COLORED batches: eight lanes per world (k_gs_world_exact) over the same (world, colour) grouping
  const bool perWorldX = P.colored && !fast && !split && nW > 1 && !w->gxOff;
  const int* order = w->order.p;
  if (perWorld || perWorldX) {
    // (world, colour) bins: units - hence rows - of a world are contiguous colour by colour
    const int nBins = nW * GR_LV + 1;
    int* wc = w->worldCount.p;
    W_TRY(w, cudaMemsetAsync(wc, 0, 2 * (size_t)(nBins + 1) * sizeof(int), s));
    { g_kernel_launches++; k_world_count<<<grid_for(w, w->unitCap, 256), 256, 0, s>>>(U, w->order.p, w->unitLevel.p, w->world.p, wc); }
    W_TRY(w, scan_exclusive(wc, w->worldUnitStart.p, nullptr, nBins, nBins, nullptr, w->scanTmp, s));
    { g_kernel_launches++; k_world_fill<<<grid_for(w, w->unitCap, 256), 256, 0, s>>>(U, w->order.p, w->unitLevel.p, w->world.p, w->worldUnitStart.p, wc + nBins + 1, w->orderW.p); }
    order = w->orderW.p;
  } else if (((fast && !w->gsFastV1) || packed) && !w->gsNoLenSort) {
    // homogeneous windows for k_gs_fast: units of a colour ordered by row count (k_solver.cuh, k_len_*)
    W_TRY(w, cudaMemsetAsync(w->lenBins.p, 0, LEN_BINS * sizeof(int), s));
    { g_kernel_launches++; k_len_count<<<grid_for(w, w->unitCap, 256), 256, 0, s>>>(U, w->order.p, w->unitLevel.p, w->lenBins.p); }
    { g_kernel_launches++; k_len_starts<<<1, LEN_BINS, 0, s>>>(w->levelStart.p, cnt + CT_NLEVELS, w->lenBins.p); }
    { g_kernel_launches++; k_len_fill<<<grid_for(w, w->unitCap, 256), 256, 0, s>>>(U, w->order.p, w->unitLevel.p, w->lenBins.p, w->orderW.p); }
    order = w->orderW.p;
  }
  { g_kernel_launches++; k_exec_units<<<grid_for(w, w->unitCap, 256), 256, 0, s>>>(B, U, order, w->unitLevel.p); }
  { g_kernel_launches++; k_zero_tail<<<1, 32, 0, s>>>(w->eRows.p, cnt + CT_NEXEC, w->unitCap); }
  { g_kernel_launches++; k_units_plus_one<<<1, 32, 0, s>>>(cnt); }
  W_TRY(w, scan_exclusive(w->eRows.p, w->eRowBase.p, cnt + CT_NUNITS1, 0, w->unitCap + 1, nullptr, w->scanTmp, s));
  GsTasks T;
  T.tab = w->gsTab.p; T.lvlTask = w->gsLvlTask.p; T.lvlWin = w->gsLvlWin.p; T.nTasks = cnt + CT_GS_NTASKS; T.taskCap = w->gsTaskCap;
  T.winMin = packed ? 0 : GS_WIN_MIN; T.winMax = packed ? 0 : GS_WIN_MAX;
  if (packed) {
    // windows of 32 units per colour, their row slots (32 x longest unit of the window), then the rows at those slots
    { g_kernel_launches++; k_gs_task_levels<<<1, 256, 0, s>>>(U, S, T, cnt + CT_OVF_ROWS); }
    { g_kernel_launches++; k_gx_windows<<<grid_for(w, 32LL * w->gxWinCap, 256), 256, 0, s>>>(U, w->levelStart.p, cnt + CT_NLEVELS, w->gsLvlTask.p, w->gxWinRows.p, w->gxWinCap, cnt + CT_OVF_ROWS); }
    W_TRY(w, scan_exclusive(w->gxWinRows.p, w->gxWinBase.p, nullptr, w->gxWinCap, w->gxWinCap, cnt + CT_NROWS_PAD, w->scanTmp, s));
  }
  { g_kernel_launches++; k_rows_build<<<grid_for(w, w->unitCap, 128), 128, 0, s>>>(B, C, Us, J, U, R, P, order, cnt + CT_OVF_ROWS, split ? w->islandLabel.p : w->world.p, nGroups); }
  if (packed) {
  } else if (fast && !w->gsFastV1 && !perWorld) {
    { g_kernel_launches++; k_gs_task_levels<<<1, 256, 0, s>>>(U, S, T, cnt + CT_OVF_ROWS); }
    { g_kernel_launches++; k_gs_task_fill<<<grid_for(w, w->gsTaskCap + 1, 256), 256, 0, s>>>(U, S, T); }
  }
  if (w->recordSolveEvents) cudaEventRecord(w->ev[7], s);
  GsStats G;
  G.worldTot = w->worldTot.p; G.worldDone = w->worldDone.p; G.worldIters = w->worldIters.p; G.itersDone = cnt + CT_ITERS;
  G.bodyGroup = split ? w->islandLabel.p : w->world.p; G.nGroups = nGroups;
  {
    void* args[] = {&R, &B, &U, &S, &P, &G};
    void* argsT[] = {&R, &B, &U, &S, &T, &P, &G};
    g_kernel_launches++;
    if (perWorldX) {
      k_gs_world_exact<<<nW, 32, GWX_SMEM_BYTES, s>>>(R, B, U, P, G, w->worldUnitStart.p, w->worldStart.p, cnt + CT_NLEVELS);
    } else if (perWorld) {
      const int* wus = w->worldUnitStart.p;
      const int* wbs = w->worldStart.p;
      // small worlds (the RL / parameter-sweep case) take the warp-per-world ring kernel; if one world of the batch
      // does not fit its shared tables the staged CTA-per-world kernel does the whole batch (decided on the device)
      int* ringOk = cnt + CT_RING_OK;
      { g_kernel_launches++; k_set_int<<<1, 32, 0, s>>>(ringOk, w->gwNoRing ? 0 : 1); }
      if (!w->gwNoRing) {
        { g_kernel_launches++; k_world_ring_check<<<grid_for(w, nW, 256), 256, 0, s>>>(U, wus, wbs, nW, cnt + CT_NLEVELS, ringOk); }
        { g_kernel_launches++; k_gs_world_ring<<<nW, 32, 0, s>>>(R, B, U, P, G, wus, wbs, ringOk); }
      }
      k_gs_world<<<nW, GW_THREADS, GW_SMEM_BYTES, s>>>(R, B, U, S, P, G, wus, wbs, ringOk);
    } else if (fast && w->gsFastV1) W_TRY(w, cudaLaunchCooperativeKernel((void*)k_gs_fast_v1, dim3(w->coopBlocksGsFastV1), dim3(256), args, 0, s));
    else if (packed) {
      GxState X;
      X.lvlTask = w->gsLvlTask.p;
      void* argsX[] = {&R, &B, &U, &S, &P, &G, &X};
      W_TRY(w, cudaLaunchCooperativeKernel((void*)k_gs_exact, dim3(w->coopBlocksGx), dim3(GX_THREADS), argsX, GX_SMEM_BYTES, s));
    }
    else if (fast) W_TRY(w, cudaLaunchCooperativeKernel((void*)k_gs_fast, dim3(w->coopBlocksGsFast), dim3(GS_THREADS), argsT, GS_SMEM_BYTES, s));
    else W_TRY(w, cudaLaunchCooperativeKernel((void*)k_gs, dim3(w->coopBlocksGs), dim3(256), args, 0, s));
  }
  if (w->recordSolveEvents) cudaEventRecord(w->ev[10], s);
  W_TRY(w, cudaGetLastError());
  if (tracePath) {
    std::vector<long long> h(traceWords);
    W_TRY(w, cudaStreamSynchronize(s));
    W_TRY(w, cudaMemcpy(h.data(), w->gsTrace.p, traceWords * sizeof(long long), cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(tracePath, "wb")) { fwrite(h.data(), sizeof(long long), traceWords, f); fclose(f); }
  }
  return CANNON_OK;
}

static int32_t refresh_damping(cannon_world* w, double dt) {
  if (w->powDt == dt) return CANNON_OK;
  // world_class.dart:652,656: pow(1 - damping, dt) with the host libm (bodies mostly share a few damping values)
  std::vector<double> lp(w->n), ap(w->n);
  std::map<double, double> cache;
  auto pw = [&](double d) {
    auto it = cache.find(d);
    if (it != cache.end()) return it->second;
    const double v = pow(1.0 - d, dt);
    cache[d] = v;
    return v;
  };
  for (int i = 0; i < w->n; i++) { lp[i] = pw(w->hLdamp[i]); ap[i] = pw(w->hAdamp[i]); }
  W_TRY(w, upload(w->ldpow, lp, w->ctx->stream));
  W_TRY(w, upload(w->adpow, ap, w->ctx->stream));
  W_TRY(w, cudaStreamSynchronize(w->ctx->stream));
  w->powDt = dt;
  return CANNON_OK;
}

static int32_t st_integrate(cannon_world* w, double dt, int applyLambda) {
  StepParams P = step_params(w, dt);
  { g_kernel_launches++; k_integrate<<<grid_for(w, w->n, 256), 256, 0, w->ctx->stream>>>(body_arrays(w), P, w->worldRows.p, applyLambda); }
  if (w->nSprings > 0) {  // the postStep slot (world_class.dart:685): forces for the next step
    SpringArrays S;
    S.n = w->nSprings; S.bodyA = w->spBodyA.p; S.bodyB = w->spBodyB.p; S.rest = w->spRest.p; S.stiffness = w->spK.p; S.damping = w->spD.p;
    S.anchorA = w->spAnchorA.p; S.anchorB = w->spAnchorB.p; S.off = w->spOff.p; S.idx = w->spIdx.p;
    g_kernel_launches++;
    k_springs<<<grid_for(w, w->n, 256), 256, 0, w->ctx->stream>>>(body_arrays(w), S, w->n);
    if (P.allowSleep) { g_kernel_launches++; k_sleep_tick<<<grid_for(w, w->n, 256), 256, 0, w->ctx->stream>>>(body_arrays(w), P); }
  }
  W_TRY(w, cudaGetLastError());
  return CANNON_OK;
}

static int32_t check_overflow_acc(cannon_world* w) {
  const long long* a = w->hAcc;
  char buf[256];
  if (a[AC_OVF_PAIRS] > 0) { snprintf(buf, sizeof buf, "pair capacity exceeded: need %lld, have %d (set cannon_world_desc.max_pairs)", a[AC_OVF_PAIRS], w->pairCap); return fail(w->ctx, CANNON_E_CAPACITY, buf); }
  if (a[AC_OVF_TASKS] > 0) { snprintf(buf, sizeof buf, "narrowphase task capacity exceeded: need %lld, have %d (raise max_pairs)", a[AC_OVF_TASKS], w->taskCap); return fail(w->ctx, CANNON_E_CAPACITY, buf); }
  if (a[AC_OVF_CONTACTS] > 0) { snprintf(buf, sizeof buf, "contact capacity exceeded: need %lld, have %d (set cannon_world_desc.max_contacts)", a[AC_OVF_CONTACTS], w->contactCap); return fail(w->ctx, CANNON_E_CAPACITY, buf); }
  if (a[AC_OVF_ROWS] > 0) { snprintf(buf, sizeof buf, "row capacity exceeded: need %lld, have %d", a[AC_OVF_ROWS], w->rowCap); return fail(w->ctx, CANNON_E_CAPACITY, buf); }
  if (a[AC_OVF_LEVELS] > 0) return fail(w->ctx, CANNON_E_CAPACITY, "solver dependency levels exceeded the level table");
  if (a[AC_OVF_CLIP] > 0) return fail(w->ctx, CANNON_E_CAPACITY, "clipped contact polygon exceeded NP_MAXPOLY vertices");
  if (a[AC_GS_ABORT] > 0) return fail(w->ctx, CANNON_E_CUDA, "k_gs_exact: a dependency wait ran out (solver aborted instead of hanging)");
  if (a[AC_UNSUPPORTED] > 0) return fail(w->ctx, CANNON_E_UNSUPPORTED, "a trimesh met a box / convex / particle / trimesh: the reference's resolvers for these pairs are unfinished (narrow_phase.dart:2265-2341)");
  return CANNON_OK;
}

static int32_t check_overflow(cannon_world* w) {
  const int* c = w->hCnt;
  char buf[256];
  if (c[CT_OVF_PAIRS] > 0) { snprintf(buf, sizeof buf, "pair capacity exceeded: need %d, have %d (set cannon_world_desc.max_pairs)", c[CT_OVF_PAIRS], w->pairCap); return fail(w->ctx, CANNON_E_CAPACITY, buf); }
  if (c[CT_OVF_TASKS] > 0 || c[CT_NTASKS] > w->taskCap) { snprintf(buf, sizeof buf, "narrowphase task capacity exceeded: need %d, have %d (raise max_pairs)", std::max(c[CT_OVF_TASKS], c[CT_NTASKS]), w->taskCap); return fail(w->ctx, CANNON_E_CAPACITY, buf); }
  if (c[CT_OVF_CONTACTS] > 0 || c[CT_NCONTACTS] > w->contactCap) { snprintf(buf, sizeof buf, "contact capacity exceeded: need %d, have %d (set cannon_world_desc.max_contacts)", std::max(c[CT_OVF_CONTACTS], c[CT_NCONTACTS]), w->contactCap); return fail(w->ctx, CANNON_E_CAPACITY, buf); }
  if (c[CT_OVF_ROWS] > 0) { snprintf(buf, sizeof buf, "row capacity exceeded: need %d, have %d", c[CT_OVF_ROWS], w->rowCap); return fail(w->ctx, CANNON_E_CAPACITY, buf); }
  if (c[CT_OVF_LEVELS] > 0) return fail(w->ctx, CANNON_E_CAPACITY, "solver dependency levels exceeded the level table");
  if (c[CT_OVF_CLIP] > 0) return fail(w->ctx, CANNON_E_CAPACITY, "clipped contact polygon exceeded NP_MAXPOLY vertices");
  if (c[CT_GS_ABORT] > 0) return fail(w->ctx, CANNON_E_CUDA, "k_gs_exact: a dependency wait ran out (solver aborted instead of hanging)");
  if (c[CT_UNSUPPORTED] > 0) return fail(w->ctx, CANNON_E_UNSUPPORTED, "a trimesh met a box / convex / particle / trimesh: the reference's resolvers for these pairs are unfinished (narrow_phase.dart:2265-2341)");
  return CANNON_OK;
}

static int32_t sync_counters(cannon_world* w) {
  cudaStream_t s = w->ctx->stream;
  W_TRY(w, cudaMemcpyAsync(w->hCnt, w->cnt.p, CT_COUNT * sizeof(int), cudaMemcpyDeviceToHost, s));
  W_TRY(w, cudaStreamSynchronize(s));
  return check_overflow(w);
}

extern "C" {

int32_t cannon_apply_gravity(cannon_world* w) {
  if (!w) return CANNON_E_INVALID;
  cudaSetDevice(w->ctx->device);
  W_TRY(w, sync_clock(w));
  int32_t rc = st_prestep(w, w->dt > 0 ? w->dt : 1.0 / 60, 1, 0);
  if (rc != CANNON_OK) return rc;
  W_TRY(w, cudaStreamSynchronize(w->ctx->stream));
  return CANNON_OK;
}

int32_t cannon_broadphase_pairs(cannon_world* w, int32_t* p1, int32_t* p2, int32_t cap, int32_t* n_pairs) {
  if (!w || !n_pairs) return CANNON_E_INVALID;
  cudaSetDevice(w->ctx->device);
  W_TRY(w, sync_clock(w));
  int32_t rc;
  if ((rc = st_reset_counters(w)) != CANNON_OK) return rc;
  if ((rc = st_prestep(w, w->dt > 0 ? w->dt : 1.0 / 60, 0, 0)) != CANNON_OK) return rc;
  if ((rc = st_broadphase(w)) != CANNON_OK) return rc;
  if ((rc = sync_counters(w)) != CANNON_OK) { *n_pairs = w->hCnt[CT_OVF_PAIRS]; return rc; }
  const int np = w->hCnt[CT_NPAIRS];
  *n_pairs = np;
  w->prof.n_pairs = np;
  if (np > cap) return fail(w->ctx, CANNON_E_CAPACITY, "pair buffer too small");
  if (np > 0) {
    if (p1) W_TRY(w, cudaMemcpy(p1, w->p1.p, np * sizeof(int), cudaMemcpyDeviceToHost));
    if (p2) W_TRY(w, cudaMemcpy(p2, w->p2.p, np * sizeof(int), cudaMemcpyDeviceToHost));
  }
  return CANNON_OK;
}

static int32_t export_contacts(cannon_world* w, cannon_contacts_soa* out, int32_t* n_contacts) {
  const int nc = w->hCnt[CT_NCONTACTS];
  if (n_contacts) *n_contacts = nc;
  if (!out) return CANNON_OK;
  if (out->capacity < nc) return fail(w->ctx, CANNON_E_CAPACITY, "contact buffer too small");
  if (nc == 0) return CANNON_OK;
  std::vector<float4> tmp(nc);
  auto get3 = [&](float* dst, const float4* src) -> cudaError_t {
    cudaError_t e = cudaMemcpy(tmp.data(), src, nc * sizeof(float4), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return e;
    for (int k = 0; k < nc; k++) { dst[3 * k] = tmp[k].x; dst[3 * k + 1] = tmp[k].y; dst[3 * k + 2] = tmp[k].z; }
    return cudaSuccess;
  };
  if (out->body_i) W_TRY(w, cudaMemcpy(out->body_i, w->cBi.p, nc * sizeof(int), cudaMemcpyDeviceToHost));
  if (out->body_j) W_TRY(w, cudaMemcpy(out->body_j, w->cBj.p, nc * sizeof(int), cudaMemcpyDeviceToHost));
  if (out->ri) W_TRY(w, get3(out->ri, w->cRi.p));
  if (out->rj) W_TRY(w, get3(out->rj, w->cRj.p));
  if (out->ni) W_TRY(w, get3(out->ni, w->cNi.p));
  if (out->restitution) W_TRY(w, cudaMemcpy(out->restitution, w->cRest.p, nc * sizeof(double), cudaMemcpyDeviceToHost));
  if (out->friction) W_TRY(w, cudaMemcpy(out->friction, w->cMu.p, nc * sizeof(double), cudaMemcpyDeviceToHost));
  if (out->enabled) {
    std::vector<int> en(nc);
    W_TRY(w, cudaMemcpy(en.data(), w->cEnabled.p, nc * sizeof(int), cudaMemcpyDeviceToHost));
    for (int k = 0; k < nc; k++) out->enabled[k] = (uint8_t)en[k];
  }
  if (out->multiplier) {
    const double h = w->dt > 0 ? w->dt : 1.0 / 60;
    { g_kernel_launches++; k_multipliers<<<grid_for(w, nc, 256), 256, 0, w->ctx->stream>>>(contact_arrays(w), row_arrays(w), w->contactCap, 1 / h, w->cMult.p); }
    W_TRY(w, cudaStreamSynchronize(w->ctx->stream));
    W_TRY(w, cudaMemcpy(out->multiplier, w->cMult.p, nc * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return CANNON_OK;
}

int32_t cannon_narrowphase_contacts(cannon_world* w, const int32_t* p1, const int32_t* p2, int32_t np, cannon_contacts_soa* out,
                                    int32_t* n_contacts, int32_t* per_pair_count) {
  if (!w || np < 0 || (np > 0 && (!p1 || !p2))) return CANNON_E_INVALID;
  cudaSetDevice(w->ctx->device);
  W_TRY(w, sync_clock(w));
  cudaStream_t s = w->ctx->stream;
  if (np > w->pairCap) return fail(w->ctx, CANNON_E_CAPACITY, "more pairs than cannon_world_desc.max_pairs");
  for (int k = 0; k < np; k++)
    if (p1[k] < 0 || p2[k] < 0 || p1[k] >= w->n || p2[k] >= w->n) return fail(w->ctx, CANNON_E_INVALID, "pair references unknown body");
  int32_t rc;
  if ((rc = ensure_clip_buffers(w)) != CANNON_OK) return rc;
  if ((rc = st_reset_counters(w)) != CANNON_OK) return rc;
  if (np > 0) {
    W_TRY(w, cudaMemcpyAsync(w->p1.p, p1, np * sizeof(int), cudaMemcpyHostToDevice, s));
    W_TRY(w, cudaMemcpyAsync(w->p2.p, p2, np * sizeof(int), cudaMemcpyHostToDevice, s));
  }
  { g_kernel_launches++; k_set_int<<<1, 32, 0, s>>>(w->cnt.p + CT_NPAIRS, np); }
  if (w->dt < 0) w->dt = 1.0 / 60;  // World.defaultDt
  if ((rc = st_narrowphase(w, w->dt)) != CANNON_OK) return rc;
  if (per_pair_count && np > 0) {
    if (w->compound) {
      { g_kernel_launches++; k_np_per_pair<<<grid_for(w, w->ppCap, 256), 256, 0, s>>>(np_arrays(w), w->ppPer.p); }
      { g_kernel_launches++; k_pp_per_pair<<<grid_for(w, np, 256), 256, 0, s>>>(w->ppPer.p, w->ppOff.p, w->ppCnt.p, np, w->keep.p); }
    } else {
      { g_kernel_launches++; k_np_per_pair<<<grid_for(w, np, 256), 256, 0, s>>>(np_arrays(w), w->keep.p); }
    }
    W_TRY(w, cudaGetLastError());
  }
  if ((rc = sync_counters(w)) != CANNON_OK) return rc;
  w->prof.n_pairs = np;
  w->prof.n_contacts = w->hCnt[CT_NCONTACTS];
  if (per_pair_count && np > 0) W_TRY(w, cudaMemcpy(per_pair_count, w->keep.p, np * sizeof(int), cudaMemcpyDeviceToHost));
  return export_contacts(w, out, n_contacts);
}

int32_t cannon_solver_solve(cannon_world* w, double dt, int32_t* iterations_done) {
  if (!w) return CANNON_E_INVALID;
  cudaSetDevice(w->ctx->device);
  cudaStream_t s = w->ctx->stream;
  int32_t rc;
  if ((rc = ensure_events(w)) != CANNON_OK) return rc;
  // keep the pair/task/contact counts of the preceding narrowphase call, clear the solver's
  W_TRY(w, cudaMemsetAsync(w->cnt.p + CT_NROWS, 0, sizeof(int), s));
  W_TRY(w, cudaMemsetAsync(w->cnt.p + CT_FRICTOTAL, 0, (CT_GS_NTASKS + 1 - CT_FRICTOTAL) * sizeof(int), s));
  W_TRY(w, cudaMemsetAsync(w->cnt.p + CT_NUNITS, 0, (CT_NISLANDS + 1 - CT_NUNITS) * sizeof(int), s));
  W_TRY(w, cudaMemsetAsync(w->cnt.p + CT_BAR, 0, 64 * sizeof(int), s));
  if ((rc = st_solve(w, dt)) != CANNON_OK) return rc;
  { g_kernel_launches++; k_apply_lambda<<<grid_for(w, w->n, 256), 256, 0, s>>>(body_arrays(w), w->n, w->desc.n_worlds, w->worldRows.p); }
  W_TRY(w, cudaGetLastError());
  if ((rc = sync_counters(w)) != CANNON_OK) return rc;
  w->dt = dt;
  w->prof.n_rows = w->hCnt[CT_NROWS];
  w->prof.n_levels = w->hCnt[CT_NLEVELS];
  w->prof.iterations_done = w->hCnt[CT_ITERS];
  w->lastUnits = w->hCnt[CT_NUNITS]; w->lastLevels = w->hCnt[CT_NLEVELS];
  w->prof.n_islands = w->hCnt[CT_NISLANDS];
  if (iterations_done) *iterations_done = w->hCnt[CT_ITERS];
  return CANNON_OK;
}

int32_t cannon_integrate(cannon_world* w, double dt) {
  if (!w) return CANNON_E_INVALID;
  cudaSetDevice(w->ctx->device);
  W_TRY(w, sync_clock(w));
  int32_t rc;
  if ((rc = refresh_damping(w, dt)) != CANNON_OK) return rc;
  if ((rc = st_integrate(w, dt, 0)) != CANNON_OK) return rc;
  W_TRY(w, cudaStreamSynchronize(w->ctx->stream));
  w->stepnumber += 1;
  w->time += dt;
  return CANNON_OK;
}

// everything one World.step enqueues on the library's stream(s), with the stage events of cannon_profile
static int32_t enqueue_step(cannon_world* w, double dt) {
  cudaStream_t s = w->ctx->stream;
  int32_t rc;
  w->recordSolveEvents = true;
  if ((rc = st_reset_counters(w)) != CANNON_OK) return rc;
  cudaEventRecord(w->ev[0], s);
  if ((rc = st_prestep(w, dt, 1, 0)) != CANNON_OK) return rc;
  if ((rc = st_broadphase(w)) != CANNON_OK) return rc;
  cudaEventRecord(w->ev[1], s);
  if ((rc = st_narrowphase(w, dt)) != CANNON_OK) return rc;
  cudaEventRecord(w->ev[2], s);
  if ((rc = st_solve(w, dt)) != CANNON_OK) return rc;
  cudaEventRecord(w->ev[3], s);
  if ((rc = st_integrate(w, dt, 1)) != CANNON_OK) return rc;
  cudaEventRecord(w->ev[4], s);
  // statistics + sticky overflow needs are folded on the device: no host round trip between steps
  { g_kernel_launches++; k_step_epilogue<<<1, 32, 0, s>>>(w->cnt.p, w->acc.p, w->taskCap, w->contactCap, w->dClock.p, dt); }
  w->recordSolveEvents = false;
  return CANNON_OK;
}

static void drop_step_graph(cannon_world* w) {
  if (w->stepGraph) { cudaGraphExecDestroy(w->stepGraph); w->stepGraph = nullptr; }
  w->eagerSteps = 0;
}

// stage events of one eager step, in enqueue_step's order: start, after broadphase / narrowphase / solve / integrate,
// schedule begin / end, sweep begin, -, -, sweep end (indices into cannon_world::ev)
#define PROF_EV 11

// enqueues nsteps steps and the asynchronous read-back of the counters; no host synchronisation.
// profiled: every step runs eagerly with its own set of stage events (cannon_world_step_profiled)
static int32_t step_enqueue(cannon_world* w, double dt, int32_t nsteps, bool profiled) {
  cudaSetDevice(w->ctx->device);
  W_TRY(w, sync_clock(w));
  cudaStream_t s = w->ctx->stream;
  int32_t rc;
  if ((rc = refresh_damping(w, dt)) != CANNON_OK) return rc;
  if ((rc = ensure_events(w)) != CANNON_OK) return rc;
  if ((rc = ensure_clip_buffers(w)) != CANNON_OK) return rc;
  w->dt = dt;
  const bool wantGraph = !profiled && !w->graphBroken && !getenv("CANNON_NO_GRAPH") && !getenv("CANNON_GS_TRACE");
  if (profiled) {
    const size_t need = (size_t)nsteps * PROF_EV;
    while (w->profPool.size() < need) {
      cudaEvent_t e;
      W_TRY(w, cudaEventCreate(&e));
      w->profPool.push_back(e);
    }
  }
  cudaEvent_t keep[PROF_EV];
  for (int k = 0; k < PROF_EV; k++) keep[k] = w->ev[k];
  cudaEventRecord(keep[8], s);
  for (int it = 0; it < nsteps; it++) {
    bool done = false;
    // the stage events of cannon_profile only time eager launches, so a multi-step call runs its last step eagerly
    const bool lastEager = it == nsteps - 1 && nsteps > 1;
    if (wantGraph && w->eagerSteps >= 1 && !lastEager) {
      if (w->stepGraph && w->graphDt != dt) { cudaGraphExecDestroy(w->stepGraph); w->stepGraph = nullptr; }
      if (!w->stepGraph) {
        const long long l0 = g_kernel_launches;
        cudaGraph_t g = nullptr;
        {
          const cudaError_t stale = cudaGetLastError();
          if (stale != cudaSuccess && getenv("CANNON_GRAPH_DEBUG")) fprintf(stderr, "[cannon] stale error before capture: %s\n", cudaGetErrorString(stale));
        }
        cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
          rc = enqueue_step(w, dt);
          e = cudaStreamEndCapture(s, &g);
          if (rc == CANNON_OK && e == cudaSuccess && g) e = cudaGraphInstantiate(&w->stepGraph, g, 0);
          else if (e == cudaSuccess) e = cudaErrorUnknown;
          if (g) cudaGraphDestroy(g);
        }
        w->graphLaunches = g_kernel_launches - l0;
        g_kernel_launches = l0;  // nothing ran yet
        if (e != cudaSuccess || !w->stepGraph) {
          if (getenv("CANNON_GRAPH_DEBUG")) fprintf(stderr, "[cannon] step graph capture failed: rc=%d cuda=%s last=%s\n", rc, cudaGetErrorString(e), cannon_last_error(w->ctx));
          cudaGetLastError();
          w->stepGraph = nullptr;
          w->graphBroken = true;  // fall back to eager launches for good
        } else {
          w->graphDt = dt;
        }
      }
      if (w->stepGraph) {
        W_TRY(w, cudaGraphLaunch(w->stepGraph, s));
        g_kernel_launches += w->graphLaunches;
        done = true;
      }
    }
    if (!done) {
      if (profiled)
        for (int k = 0; k < PROF_EV; k++)
          if (k != 8 && k != 9) w->ev[k] = w->profPool[(size_t)it * PROF_EV + k];
      rc = enqueue_step(w, dt);
      if (profiled) for (int k = 0; k < PROF_EV; k++) w->ev[k] = keep[k];
      if (rc != CANNON_OK) return rc;
      w->eagerSteps++;
    }
    w->stageEventsValid = !done;
    w->stepnumber += 1;
    w->time += dt;  // World.step: time += dt after internalStep (world_class.dart:396-399)
  }
  w->recordSolveEvents = false;
  cudaEventRecord(w->ev[9], s);
  if (nsteps > 0) {
    W_TRY(w, cudaMemcpyAsync(w->hCnt, w->cnt.p, CT_COUNT * sizeof(int), cudaMemcpyDeviceToHost, s));
    W_TRY(w, cudaMemcpyAsync(w->hAcc, w->acc.p, AC_COUNT * sizeof(long long), cudaMemcpyDeviceToHost, s));
  }
  return CANNON_OK;
}

// after the stream has been synchronised: overflow check + cannon_profile of the call
static int32_t step_finish(cannon_world* w, int32_t nsteps, bool profiled) {
  w->stepPending = false;
  if (nsteps <= 0) return CANNON_OK;
  int32_t rc;
  if ((rc = check_overflow_acc(w)) != CANNON_OK) return rc;
  float ms;
  cannon_profile& p = w->prof;
  p.steps = w->hAcc[AC_STEPS];
  p.contact_iters_total = w->hAcc[AC_CONTACT_ITERS];
  auto stage = [&](cudaEvent_t* ev) {
    if (cudaEventElapsedTime(&ms, ev[0], ev[1]) == cudaSuccess) p.broadphase = ms;
    if (cudaEventElapsedTime(&ms, ev[1], ev[2]) == cudaSuccess) p.narrowphase = ms;
    if (cudaEventElapsedTime(&ms, ev[2], ev[3]) == cudaSuccess) { p.solve = ms; p.make_contact_constraints = 0; }
    if (cudaEventElapsedTime(&ms, ev[3], ev[4]) == cudaSuccess) p.integrate = ms;
    if (cudaEventElapsedTime(&ms, ev[5], ev[6]) == cudaSuccess) p.schedule_ms = ms;
    if (cudaEventElapsedTime(&ms, ev[7], ev[10]) == cudaSuccess) p.gs_ms = ms;
  };
  if (profiled) {
    p.sum_steps = 0; p.sum_step_ms = p.sum_broadphase = p.sum_narrowphase = p.sum_solve = p.sum_integrate = p.sum_schedule = p.sum_gs = 0;
    for (int it = 0; it < nsteps; it++) {
      cudaEvent_t* ev = &w->profPool[(size_t)it * PROF_EV];
      stage(ev);
      p.sum_broadphase += p.broadphase; p.sum_narrowphase += p.narrowphase; p.sum_solve += p.solve; p.sum_integrate += p.integrate;
      p.sum_schedule += p.schedule_ms; p.sum_gs += p.gs_ms;
      if (cudaEventElapsedTime(&ms, ev[0], ev[4]) == cudaSuccess) p.sum_step_ms += ms;
      p.sum_steps += 1;
    }
  } else if (w->stageEventsValid) {  // events recorded inside a graph replay cannot be read: the previous values stay
    stage(w->ev);
  }
  if (cudaEventElapsedTime(&ms, w->ev[8], w->ev[9]) == cudaSuccess) p.step_call_ms = ms;
  cudaGetLastError();  // an event pair that was not recorded in this call is not an error of the step
  p.n_pairs = w->hCnt[CT_NPAIRS]; p.n_contacts = w->hCnt[CT_NCONTACTS]; p.n_rows = w->hCnt[CT_NROWS];
  p.n_levels = w->hCnt[CT_NLEVELS]; p.iterations_done = w->hCnt[CT_ITERS];
  w->lastUnits = w->hCnt[CT_NUNITS]; w->lastLevels = w->hCnt[CT_NLEVELS];
  p.n_tasks = w->hCnt[CT_NTASKS];
  p.n_islands = w->hCnt[CT_NISLANDS];
  for (int t = 0; t < 8; t++) p.n_tasks_by_type[t] = w->hCnt[CT_BUCKETCOUNT + t];  // the four Particle task types are only in n_tasks
  return CANNON_OK;
}

int32_t cannon_world_step(cannon_world* w, double dt, int32_t nsteps) {
  if (!w || nsteps < 0) return CANNON_E_INVALID;
  int32_t rc;
  if ((rc = step_enqueue(w, dt, nsteps, false)) != CANNON_OK) return rc;
  W_TRY(w, cudaStreamSynchronize(w->ctx->stream));
  return step_finish(w, nsteps, false);
}

// every step eager and bracketed by its own stage events: cannon_profile.sum_* hold the per-stage device time summed over
// the nsteps steps of this call (what bench.py reports as the sweep kernel's average launch duration)
int32_t cannon_world_step_profiled(cannon_world* w, double dt, int32_t nsteps) {
  if (!w || nsteps < 0 || nsteps > 4096) return CANNON_E_INVALID;
  int32_t rc;
  if ((rc = step_enqueue(w, dt, nsteps, true)) != CANNON_OK) return rc;
  W_TRY(w, cudaStreamSynchronize(w->ctx->stream));
  return step_finish(w, nsteps, true);
}

// World.step without waiting for the device: the call returns once the work is enqueued on the ctx's stream, so one host
// thread (one Dart isolate) can keep several GPUs busy; cannon_ctx_sync collects the result
int32_t cannon_world_step_async(cannon_world* w, double dt, int32_t nsteps) {
  if (!w || nsteps < 0) return CANNON_E_INVALID;
  int32_t rc;
  if ((rc = step_enqueue(w, dt, nsteps, false)) != CANNON_OK) return rc;
  if (nsteps > 0 && !w->stepPending) { w->stepPending = true; w->ctx->pending.push_back(w); }
  return CANNON_OK;
}

int32_t cannon_ctx_sync(cannon_ctx* ctx) {
  if (!ctx) return CANNON_E_INVALID;
  cudaSetDevice(ctx->device);
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  int32_t first = CANNON_OK;
  std::vector<cannon_world*> pend;
  pend.swap(ctx->pending);
  for (cannon_world* w : pend) {
    const int32_t rc = step_finish(w, 1, false);
    if (rc != CANNON_OK && first == CANNON_OK) first = rc;
  }
  return first;
}

int32_t cannon_world_profile(cannon_world* w, cannon_profile* out) {
  if (!w || !out) return CANNON_E_INVALID;
  w->prof.kernel_launches = g_kernel_launches;
  *out = w->prof;
  return CANNON_OK;
}

int32_t cannon_world_enable_contact_events(cannon_world* w, int32_t enable) {
  if (!w) return CANNON_E_INVALID;
  cudaSetDevice(w->ctx->device);
  drop_step_graph(w);  // the captured step does or does not contain the event kernels
  w->evEnabled = enable != 0;
  w->evCap = 0;        // (re)start from an empty previous set
  return ensure_events(w);
}

int32_t cannon_world_get_contact_events(cannon_world* w, int32_t cap, int32_t* n_begin, int32_t* begin_a, int32_t* begin_b, int32_t* n_end,
                                        int32_t* end_a, int32_t* end_b) {
  if (!w || cap < 0 || !n_begin || !n_end) return CANNON_E_INVALID;
  if (!w->evEnabled) return fail(w->ctx, CANNON_E_INVALID, "contact events are not enabled");
  cudaSetDevice(w->ctx->device);
  *n_begin = 0; *n_end = 0;
  if (w->evCap <= 0) return CANNON_OK;
  int c[8];
  W_TRY(w, cudaStreamSynchronize(w->ctx->stream));
  W_TRY(w, cudaMemcpy(c, w->evCnt.p, sizeof c, cudaMemcpyDeviceToHost));
  if (c[0] > w->evCap || c[2] > w->evCap || c[3] > w->evCap) return fail(w->ctx, CANNON_E_CAPACITY, "contact pair capacity exceeded (set cannon_world_desc.max_pairs)");
  *n_begin = c[2]; *n_end = c[3];
  if (c[2] > cap || c[3] > cap) return fail(w->ctx, CANNON_E_CAPACITY, "contact event arrays too small");
  std::vector<unsigned long long> kb(c[2]), ke(c[3]);
  if (c[2]) W_TRY(w, cudaMemcpy(kb.data(), w->evBegin.p, kb.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (c[3]) W_TRY(w, cudaMemcpy(ke.data(), w->evEnd.p, ke.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  std::sort(kb.begin(), kb.end());  // OverlapKeeper.getDiff walks its sorted key lists
  std::sort(ke.begin(), ke.end());
  for (int k = 0; k < c[2]; k++) { if (begin_a) begin_a[k] = (int32_t)(kb[k] >> 32); if (begin_b) begin_b[k] = (int32_t)(kb[k] & 0xffffffffull); }
  for (int k = 0; k < c[3]; k++) { if (end_a) end_a[k] = (int32_t)(ke[k] >> 32); if (end_b) end_b[k] = (int32_t)(ke[k] & 0xffffffffull); }
  return CANNON_OK;
}

int32_t cannon_world_get_contacts(cannon_world* w, cannon_contacts_soa* out, int32_t* n_contacts) {
  if (!w) return CANNON_E_INVALID;
  cudaSetDevice(w->ctx->device);
  return export_contacts(w, out, n_contacts);
}

int32_t cannon_world_get_rows(cannon_world* w, int32_t cap, int32_t* n_rows, int32_t* body_i, int32_t* body_j, double* B, double* invC,
                              double* lambda, int32_t* level) {
  if (!w) return CANNON_E_INVALID;
  cudaSetDevice(w->ctx->device);
  const int n = w->hCnt[CT_NROWS];
  const int nu = w->hCnt[CT_NUNITS];
  if (n_rows) *n_rows = n;
  if (cap < n) return fail(w->ctx, CANNON_E_CAPACITY, "row buffer too small");
  if (n == 0) return CANNON_OK;
  // rows live in execution order on the device. REFERENCE_ORDER: unit id == reference row index, so the rows are
  // returned in the reference's solve order; COLORED: returned in execution order.
  std::vector<double> hB(n), hC(n), hL(n);
  std::vector<int> uBi(nu), uBj(nu), uRow(nu), uLvl(nu), uRows(nu);
  if (kind_fast(w)) {
    std::vector<float4> q((size_t)n * 5);
    std::vector<float> fl(n);
    W_TRY(w, cudaMemcpy(q.data(), w->rRec.p, (size_t)n * 5 * sizeof(float4), cudaMemcpyDeviceToHost));
    W_TRY(w, cudaMemcpy(fl.data(), w->rFlambda.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    for (int k = 0; k < n; k++) { hB[k] = q[(size_t)k * 5].w; hC[k] = q[(size_t)k * 5 + 1].w; hL[k] = fl[k]; }
  } else if (kind_packed(w)) {
    // slot index = block * 32 + lane; B / invC are chunk 4 of the block (k_solver.cuh, GxRow)
    const int nSlots = std::max(n, w->hCnt[CT_NROWS_PAD]);
    std::vector<float4> q(((size_t)nSlots / 32 + 1) * GX_CHUNKS * 32);
    W_TRY(w, cudaMemcpy(q.data(), w->rXblk.p, q.size() * sizeof(float4), cudaMemcpyDeviceToHost));
    hB.assign(nSlots + 32, 0.0); hC.assign(nSlots + 32, 0.0); hL.assign(nSlots + 32, 0.0);
    W_TRY(w, cudaMemcpy(hL.data(), w->rLambda.p, nSlots * sizeof(double), cudaMemcpyDeviceToHost));
    for (int k = 0; k < nSlots; k++) {
      const double* d = (const double*)&q[((size_t)(k >> 5) * GX_CHUNKS + 4) * 32 + (k & 31)];
      hB[k] = d[0]; hC[k] = d[1];
    }
  } else {
    W_TRY(w, cudaMemcpy(hB.data(), w->rB.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    W_TRY(w, cudaMemcpy(hC.data(), w->rInvC.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    W_TRY(w, cudaMemcpy(hL.data(), w->rLambda.p, n * sizeof(double), cudaMemcpyDeviceToHost));
  }
  W_TRY(w, cudaMemcpy(uBi.data(), w->uBi.p, nu * sizeof(int), cudaMemcpyDeviceToHost));
  W_TRY(w, cudaMemcpy(uBj.data(), w->uBj.p, nu * sizeof(int), cudaMemcpyDeviceToHost));
  W_TRY(w, cudaMemcpy(uRow.data(), w->unitRow.p, nu * sizeof(int), cudaMemcpyDeviceToHost));
  W_TRY(w, cudaMemcpy(uLvl.data(), w->unitLevel.p, nu * sizeof(int), cudaMemcpyDeviceToHost));
  W_TRY(w, cudaMemcpy(uRows.data(), w->uRows.p, nu * sizeof(int), cudaMemcpyDeviceToHost));
  int out = 0;
  const bool refOrder = !kind_colored(w);
  const int ne = refOrder ? nu : std::min(nu, w->hCnt[CT_NEXEC]);  // units without rows are not in the execution order
  std::vector<int> unitsInOrder(nu);
  if (refOrder) { for (int u = 0; u < nu; u++) unitsInOrder[u] = u; }
  else if (ne > 0) {
    // the schedule appends the units of a colour in arrival order: report them by (colour, unit key), the canonical order
    // of include/cannon_cuda.h (units of one colour are independent, so any order inside a colour is the same solve)
    W_TRY(w, cudaMemcpy(unitsInOrder.data(), w->order.p, ne * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<int> uKey(nu);
    W_TRY(w, cudaMemcpy(uKey.data(), w->uKey.p, nu * sizeof(int), cudaMemcpyDeviceToHost));
    std::sort(unitsInOrder.begin(), unitsInOrder.begin() + ne, [&](int a, int b) { return uLvl[a] != uLvl[b] ? uLvl[a] < uLvl[b] : uKey[a] < uKey[b]; });
  }
  for (int k = 0; k < ne; k++) {
    const int u = unitsInOrder[k];
    const int rstride = kind_packed(w) ? 32 : 1;  // consecutive rows of a unit sit in consecutive blocks of its window
    for (int q = 0; q < uRows[u] && out < n; q++, out++) {
      const int r = uRow[u] + q * rstride;
      if (body_i) body_i[out] = uBi[u];
      if (body_j) body_j[out] = uBj[u];
      if (B) B[out] = hB[r];
      if (invC) invC[out] = hC[r];
      if (lambda) lambda[out] = hL[r];
      if (level) level[out] = uLvl[u];
    }
  }
  return CANNON_OK;
}

int32_t cannon_world_get_bodies(cannon_world* w, cannon_bodies_soa* o) {
  if (!w || !o) return CANNON_E_INVALID;
  cudaSetDevice(w->ctx->device);
  const int n = w->n;
  if (o->n < n) { o->n = n; return fail(w->ctx, CANNON_E_CAPACITY, "cannon_bodies_soa.n too small"); }
  o->n = n;
  if (n == 0) return CANNON_OK;
  W_TRY(w, cudaStreamSynchronize(w->ctx->stream));
  W_TRY(w, w->stage.reserve((size_t)4 * n));
  cudaStream_t gs = w->ctx->stream;
  auto get4 = [&](float* dst, const float4* src, int comps) -> cudaError_t {
    g_kernel_launches++;
    k_unpack_from4<<<grid_for(w, n, 256), 256, 0, gs>>>(src, w->stage.p, n, comps);
    cudaError_t e = cudaMemcpyAsync(dst, w->stage.p, (size_t)comps * n * sizeof(float), cudaMemcpyDeviceToHost, gs);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(gs);
  };
  if (o->position) W_TRY(w, get4(o->position, w->pos.p, 3));
  if (o->quaternion) W_TRY(w, get4(o->quaternion, w->quat.p, 4));
  if (o->velocity) W_TRY(w, get4(o->velocity, w->vel.p, 3));
  if (o->angular_velocity) W_TRY(w, get4(o->angular_velocity, w->angvel.p, 3));
  if (o->force) W_TRY(w, get4(o->force, w->force.p, 3));
  if (o->torque) W_TRY(w, get4(o->torque, w->torque.p, 3));
  if (o->linear_factor) W_TRY(w, get4(o->linear_factor, w->linF.p, 3));
  if (o->angular_factor) W_TRY(w, get4(o->angular_factor, w->angF.p, 3));
  if (o->inv_inertia) W_TRY(w, get4(o->inv_inertia, w->invI.p, 3));
  if (o->inv_inertia_world) {
    std::vector<float4> r0(n), r1(n), r2(n);
    W_TRY(w, cudaMemcpy(r0.data(), w->iiw0.p, n * sizeof(float4), cudaMemcpyDeviceToHost));
    W_TRY(w, cudaMemcpy(r1.data(), w->iiw1.p, n * sizeof(float4), cudaMemcpyDeviceToHost));
    W_TRY(w, cudaMemcpy(r2.data(), w->iiw2.p, n * sizeof(float4), cudaMemcpyDeviceToHost));
    for (int k = 0; k < n; k++) {
      float* d = o->inv_inertia_world + 9 * k;
      d[0] = r0[k].x; d[1] = r0[k].y; d[2] = r0[k].z; d[3] = r1[k].x; d[4] = r1[k].y; d[5] = r1[k].z; d[6] = r2[k].x; d[7] = r2[k].y; d[8] = r2[k].z;
    }
  }
  if (o->aabb) {
    int32_t rc = st_prestep(w, w->dt > 0 ? w->dt : 1.0 / 60, 0, 1);
    if (rc != CANNON_OK) return rc;
    W_TRY(w, cudaStreamSynchronize(w->ctx->stream));
    std::vector<float4> lo(n), hi(n);
    W_TRY(w, cudaMemcpy(lo.data(), w->aabbLo.p, n * sizeof(float4), cudaMemcpyDeviceToHost));
    W_TRY(w, cudaMemcpy(hi.data(), w->aabbHi.p, n * sizeof(float4), cudaMemcpyDeviceToHost));
    for (int k = 0; k < n; k++) {
      float* d = o->aabb + 6 * k;
      d[0] = lo[k].x; d[1] = lo[k].y; d[2] = lo[k].z; d[3] = hi[k].x; d[4] = hi[k].y; d[5] = hi[k].z;
    }
  }
#define GETA(field, buf, T) if (o->field) W_TRY(w, cudaMemcpy(o->field, w->buf.p, n * sizeof(T), cudaMemcpyDeviceToHost))
  GETA(mass, mass, double); GETA(type, type, int); GETA(sleep_state, sleep, int); GETA(time_last_sleepy, tLastSleepy, double);
  GETA(sleep_speed_limit, sleepSpeed, double); GETA(sleep_time_limit, sleepTime, double); GETA(linear_damping, ldamp, double);
  GETA(angular_damping, adamp, double); GETA(collision_filter_group, group, int); GETA(collision_filter_mask, mask, int);
  GETA(material, material, int); GETA(shape, shape, int); GETA(world_id, world, int); GETA(inv_mass, invMass, double);
  GETA(bounding_radius, brad, double);
#undef GETA
  if (o->allow_sleep || o->collision_response || o->is_trigger || o->fixed_rotation) {
    std::vector<int> fl(n);
    W_TRY(w, cudaMemcpy(fl.data(), w->flags.p, n * sizeof(int), cudaMemcpyDeviceToHost));
    for (int k = 0; k < n; k++) {
      if (o->allow_sleep) o->allow_sleep[k] = (fl[k] & BF_ALLOW_SLEEP) ? 1 : 0;
      if (o->collision_response) o->collision_response[k] = (fl[k] & BF_COLLISION_RESPONSE) ? 1 : 0;
      if (o->is_trigger) o->is_trigger[k] = (fl[k] & BF_IS_TRIGGER) ? 1 : 0;
      if (o->fixed_rotation) o->fixed_rotation[k] = (fl[k] & BF_FIXED_ROTATION) ? 1 : 0;
    }
  }
  return CANNON_OK;
}

__global__ void __launch_bounds__(256) k_refresh_inertia(BodyArrays B, int first, int count, int force) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
    const int i = first + k;
    const f3 I = ld3(B.invI[i]);
    if (I.x == I.y && I.y == I.z && !force) continue;  // updateInertiaWorld(force), rigid_body.dart:450-466
    float4 r0, r1, r2;
    inertia_world(ldq(B.quat[i]), I, r0, r1, r2);
    B.iiw0[i] = r0; B.iiw1[i] = r1; B.iiw2[i] = r2;
  }
}

int32_t cannon_world_update_bodies(cannon_world* w, int32_t first, int32_t count, const float* position, const float* quaternion,
                                   const float* velocity, const float* angular_velocity, const float* force, const float* torque) {
  if (!w || first < 0 || count < 0 || first + count > w->n) return CANNON_E_INVALID;
  cudaSetDevice(w->ctx->device);
  if (count == 0) return CANNON_OK;
  cudaStream_t s = w->ctx->stream;
  W_TRY(w, cudaStreamSynchronize(s));
  W_TRY(w, w->stage.reserve((size_t)4 * w->n));
  auto put = [&](const float* src, float4* dst, int comps) -> cudaError_t {
    cudaError_t e = cudaMemcpyAsync(w->stage.p, src, (size_t)comps * count * sizeof(float), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    g_kernel_launches++;
    k_pack_to4<<<grid_for(w, count, 256), 256, 0, s>>>(w->stage.p, dst + first, count, comps);
    return cudaStreamSynchronize(s);  // the caller's buffer may be reused as soon as we return
  };
  if (position) W_TRY(w, put(position, w->pos.p, 3));
  if (quaternion) W_TRY(w, put(quaternion, w->quat.p, 4));
  if (velocity) W_TRY(w, put(velocity, w->vel.p, 3));
  if (angular_velocity) W_TRY(w, put(angular_velocity, w->angvel.p, 3));
  if (force) W_TRY(w, put(force, w->force.p, 3));
  if (torque) W_TRY(w, put(torque, w->torque.p, 3));
  if (quaternion) {
    { g_kernel_launches++; k_refresh_inertia<<<grid_for(w, count, 256), 256, 0, s>>>(body_arrays(w), first, count, 0); }
    W_TRY(w, cudaStreamSynchronize(s));
  }
  return CANNON_OK;
}

int32_t cannon_world_set_inv_inertia(cannon_world* w, int32_t first, int32_t count, const float* inv_inertia) {
  if (!w || first < 0 || count < 0 || first + count > w->n || (count > 0 && !inv_inertia)) return CANNON_E_INVALID;
  cudaSetDevice(w->ctx->device);
  if (count == 0) return CANNON_OK;
  cudaStream_t s = w->ctx->stream;
  W_TRY(w, cudaStreamSynchronize(s));
  W_TRY(w, w->stage.reserve((size_t)4 * w->n));
  W_TRY(w, cudaMemcpyAsync(w->stage.p, inv_inertia, (size_t)3 * count * sizeof(float), cudaMemcpyHostToDevice, s));
  { g_kernel_launches++; k_pack_to4<<<grid_for(w, count, 256), 256, 0, s>>>(w->stage.p, w->invI.p + first, count, 3); }
  { g_kernel_launches++; k_refresh_inertia<<<grid_for(w, count, 256), 256, 0, s>>>(body_arrays(w), first, count, 1); }
  W_TRY(w, cudaStreamSynchronize(s));
  return CANNON_OK;
}

int32_t cannon_world_update_sleep_states(cannon_world* w, int32_t first, int32_t count, const int32_t* sleep_state) {
  if (!w || first < 0 || count < 0 || first + count > w->n || (count > 0 && !sleep_state)) return CANNON_E_INVALID;
  for (int k = 0; k < count; k++)
    if (sleep_state[k] < CANNON_AWAKE || sleep_state[k] > CANNON_SLEEPING) return fail(w->ctx, CANNON_E_INVALID, "sleep state out of range");
  cudaSetDevice(w->ctx->device);
  if (count == 0) return CANNON_OK;
  cudaStream_t s = w->ctx->stream;
  W_TRY(w, cudaStreamSynchronize(s));
  W_TRY(w, cudaMemcpyAsync(w->sleep.p + first, sleep_state, (size_t)count * sizeof(int), cudaMemcpyHostToDevice, s));
  W_TRY(w, cudaStreamSynchronize(s));
  return CANNON_OK;
}

int32_t cannon_world_set_hinge_motor(cannon_world* w, int32_t constraint, int32_t enabled, double target_velocity, double max_force) {
  if (!w || constraint < 0 || constraint >= (int)w->hConType.size()) return CANNON_E_INVALID;
  if (w->hConType[constraint] != CANNON_CONSTRAINT_HINGE) return fail(w->ctx, CANNON_E_INVALID, "constraint is not a HingeConstraint");
  cudaSetDevice(w->ctx->device);
  cudaStream_t s = w->ctx->stream;
  W_TRY(w, cudaStreamSynchronize(s));
  const int eq = w->hConFirst[constraint] + 5;  // x, y, z, rotational1, rotational2, motor (hinge_constraint.dart:50)
  const bool wasOn = w->hJEnabled[eq] != 0, on = enabled != 0;
  const double mn = -max_force;
  W_TRY(w, cudaMemcpyAsync(w->jTargetVel.p + eq, &target_velocity, sizeof(double), cudaMemcpyHostToDevice, s));
  W_TRY(w, cudaMemcpyAsync(w->jMaxF.p + eq, &max_force, sizeof(double), cudaMemcpyHostToDevice, s));
  W_TRY(w, cudaMemcpyAsync(w->jMinF.p + eq, &mn, sizeof(double), cudaMemcpyHostToDevice, s));
  W_TRY(w, cudaStreamSynchronize(s));
  if (wasOn == on) return CANNON_OK;
  // the set of accepted equations changes (Solver.addEquation skips disabled ones, world_class.dart:627-633): new row slots
  drop_step_graph(w);
  w->hJEnabled[eq] = on ? 1 : 0;
  std::vector<int> rowSlot(w->nJointEq), slotEq;
  int slot = 0;
  for (int e = 0; e < w->nJointEq; e++) {
    if (w->hJEnabled[e] && !w->hJTrig[e]) { slotEq.push_back(e); rowSlot[e] = slot++; } else rowSlot[e] = -1;
  }
  w->nJointAccepted = slot;
  W_TRY(w, upload(w->jEnabled, w->hJEnabled, s)); W_TRY(w, upload(w->jRowSlot, rowSlot, s)); W_TRY(w, upload(w->jSlotEq, slotEq, s));
  W_TRY(w, cudaStreamSynchronize(s));
  return ensure_capacities(w);
}

}  // extern "C"

// ---- ray casts / AABB query ---------------------------------------------------------------
extern "C" {

void cannon_ray_options_default(cannon_ray_options* o) {
  if (!o) return;
  o->mode = CANNON_RAY_CLOSEST; o->skip_backfaces = 1; o->collision_filter_mask = -1; o->collision_filter_group = -1; o->check_collision_response = 1;
}

int32_t cannon_world_raycast(cannon_world* w, int32_t n_rays, const float* from, const float* to, const cannon_ray_options* opt, uint8_t* has_hit,
                             cannon_ray_hits_soa* hits, int32_t* n_hits) {
  if (!w || n_rays < 0 || !opt || !hits || !n_hits || (n_rays > 0 && (!from || !to))) return CANNON_E_INVALID;
  if (opt->mode != CANNON_RAY_CLOSEST && opt->mode != CANNON_RAY_ANY && opt->mode != CANNON_RAY_ALL) return fail(w->ctx, CANNON_E_INVALID, "ray mode");
  if (!w->hHfs.empty() || w->hasTrimesh) return fail(w->ctx, CANNON_E_UNSUPPORTED, "heightfield / trimesh rays are outside the hot-path scope (SURVEY.md 8f)");
  const bool all = opt->mode == CANNON_RAY_ALL;
  if (!all && hits->capacity < n_rays) { *n_hits = n_rays; return fail(w->ctx, CANNON_E_CAPACITY, "hit arrays smaller than n_rays"); }
  *n_hits = 0;
  if (n_rays == 0) return CANNON_OK;
  cudaSetDevice(w->ctx->device);
  cudaStream_t s = w->ctx->stream;
  W_TRY(w, cudaStreamSynchronize(s));
  const size_t n = (size_t)n_rays;
  const int allCap = all ? std::max(hits->capacity, 1) : 1;
  DBuf<float> dFrom, dTo;
  DBuf<unsigned char> dHas;
  DBuf<int> dBody, dFace, dCount, aRay, aBody, aFace, dInst, aInst;
  DBuf<double> dDist, aDist;
  DBuf<float4> dPoint, dNormal, aPoint, aNormal;
  DBuf<unsigned long long> aKey;
  int32_t rc = CANNON_OK;
  auto done = [&](int32_t code) {
    dFrom.release(); dTo.release(); dHas.release(); dBody.release(); dFace.release(); dCount.release(); aRay.release(); aBody.release(); aFace.release(); dInst.release(); aInst.release();
    dDist.release(); aDist.release(); dPoint.release(); dNormal.release(); aPoint.release(); aNormal.release(); aKey.release();
    return code;
  };
#define R_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fail(w->ctx, CANNON_E_CUDA, cudaGetErrorString(e_)); return done(CANNON_E_CUDA); } } while (0)
  R_TRY(dFrom.reserve(3 * n)); R_TRY(dTo.reserve(3 * n)); R_TRY(dHas.reserve(n)); R_TRY(dBody.reserve(n)); R_TRY(dFace.reserve(n)); R_TRY(dCount.reserve(1));
  R_TRY(dDist.reserve(n)); R_TRY(dPoint.reserve(n)); R_TRY(dNormal.reserve(n)); R_TRY(dInst.reserve(n)); R_TRY(aInst.reserve(allCap));
  R_TRY(aRay.reserve(allCap)); R_TRY(aBody.reserve(allCap)); R_TRY(aFace.reserve(allCap)); R_TRY(aKey.reserve(allCap)); R_TRY(aDist.reserve(allCap));
  R_TRY(aPoint.reserve(allCap)); R_TRY(aNormal.reserve(allCap));
  R_TRY(cudaMemcpyAsync(dFrom.p, from, 3 * n * sizeof(float), cudaMemcpyHostToDevice, s));
  R_TRY(cudaMemcpyAsync(dTo.p, to, 3 * n * sizeof(float), cudaMemcpyHostToDevice, s));
  R_TRY(cudaMemsetAsync(dCount.p, 0, sizeof(int), s));
  RayArgs A;
  A.nRays = n_rays; A.nBodies = w->n; A.from = dFrom.p; A.to = dTo.p;
  A.mode = opt->mode; A.skipBackfaces = opt->skip_backfaces; A.mask = opt->collision_filter_mask; A.group = opt->collision_filter_group;
  A.checkCollisionResponse = opt->check_collision_response;
  A.hasHit = dHas.p; A.body = dBody.p; A.face = dFace.p; A.dist = dDist.p; A.point = dPoint.p; A.normal = dNormal.p;
  A.allCount = dCount.p; A.allCap = all ? hits->capacity : 0; A.allRay = aRay.p; A.allBody = aBody.p; A.allFace = aFace.p; A.allKey = aKey.p; A.allDist = aDist.p;
  A.allPoint = aPoint.p; A.allNormal = aNormal.p; A.inst = dInst.p; A.allInst = aInst.p;
  { g_kernel_launches++; k_raycast<<<grid_for(w, 32LL * n_rays, 128), 128, 0, s>>>(body_arrays(w), shape_tables(w), A); }
  R_TRY(cudaGetLastError());
  std::vector<unsigned char> hHas(n);
  std::vector<int> hBody(n), hFace(n), hInst(n);
  std::vector<double> hDist(n);
  std::vector<float4> hPoint(n), hNormal(n);
  int count = 0;
  R_TRY(cudaMemcpyAsync(hHas.data(), dHas.p, n, cudaMemcpyDeviceToHost, s));
  R_TRY(cudaMemcpyAsync(&count, dCount.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  if (!all) {
    R_TRY(cudaMemcpyAsync(hBody.data(), dBody.p, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    R_TRY(cudaMemcpyAsync(hFace.data(), dFace.p, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    R_TRY(cudaMemcpyAsync(hInst.data(), dInst.p, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    R_TRY(cudaMemcpyAsync(hDist.data(), dDist.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    R_TRY(cudaMemcpyAsync(hPoint.data(), dPoint.p, n * sizeof(float4), cudaMemcpyDeviceToHost, s));
    R_TRY(cudaMemcpyAsync(hNormal.data(), dNormal.p, n * sizeof(float4), cudaMemcpyDeviceToHost, s));
  }
  R_TRY(cudaStreamSynchronize(s));
  if (has_hit) memcpy(has_hit, hHas.data(), n);
  auto put = [&](int k, int ray, int body, int face, double dist, const float4& p, const float4& nn, int inst) {
    if (hits->shape_ordinal) hits->shape_ordinal[k] = inst;
    if (hits->ray) hits->ray[k] = ray;
    if (hits->body) hits->body[k] = body;
    if (hits->hit_face_index) hits->hit_face_index[k] = face;
    if (hits->distance) hits->distance[k] = dist;
    if (hits->hit_point_world) { hits->hit_point_world[3 * k] = p.x; hits->hit_point_world[3 * k + 1] = p.y; hits->hit_point_world[3 * k + 2] = p.z; }
    if (hits->hit_normal_world) { hits->hit_normal_world[3 * k] = nn.x; hits->hit_normal_world[3 * k + 1] = nn.y; hits->hit_normal_world[3 * k + 2] = nn.z; }
  };
  if (!all) {
    int nh = 0;
    for (int r = 0; r < n_rays; r++) { put(r, r, hBody[r], hFace[r], hDist[r], hPoint[r], hNormal[r], hInst[r]); nh += hHas[r] ? 1 : 0; }
    *n_hits = nh;
  } else {
    *n_hits = count;
    if (count > hits->capacity) { fail(w->ctx, CANNON_E_CAPACITY, "hit arrays too small"); return done(CANNON_E_CAPACITY); }
    const size_t m = (size_t)count;
    std::vector<int> r(m), bd(m), fc(m), in(m);
    std::vector<unsigned long long> key(m);
    std::vector<double> ds(m);
    std::vector<float4> pt(m), nm(m);
    if (m) {
      R_TRY(cudaMemcpy(r.data(), aRay.p, m * sizeof(int), cudaMemcpyDeviceToHost)); R_TRY(cudaMemcpy(bd.data(), aBody.p, m * sizeof(int), cudaMemcpyDeviceToHost));
      R_TRY(cudaMemcpy(fc.data(), aFace.p, m * sizeof(int), cudaMemcpyDeviceToHost)); R_TRY(cudaMemcpy(key.data(), aKey.p, m * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      R_TRY(cudaMemcpy(ds.data(), aDist.p, m * sizeof(double), cudaMemcpyDeviceToHost)); R_TRY(cudaMemcpy(pt.data(), aPoint.p, m * sizeof(float4), cudaMemcpyDeviceToHost));
      R_TRY(cudaMemcpy(nm.data(), aNormal.p, m * sizeof(float4), cudaMemcpyDeviceToHost));
      R_TRY(cudaMemcpy(in.data(), aInst.p, m * sizeof(int), cudaMemcpyDeviceToHost));
    }
    std::vector<int> ord(m);
    for (size_t k = 0; k < m; k++) ord[k] = (int)k;
    // the reference's callback sequence: ray by ray, bodies ascending, reports of a body in the order they were made
    std::sort(ord.begin(), ord.end(), [&](int a, int b) { return r[a] != r[b] ? r[a] < r[b] : key[a] < key[b]; });
    for (size_t k = 0; k < m; k++) { const int i = ord[k]; put((int)k, r[i], bd[i], fc[i], ds[i], pt[i], nm[i], in[i]); }
  }
#undef R_TRY
  return done(rc);
}

int32_t cannon_world_aabb_query(cannon_world* w, const float* lower, const float* upper, int32_t* bodies, int32_t cap, int32_t* n) {
  if (!w || !lower || !upper || !n || cap < 0 || (cap > 0 && !bodies)) return CANNON_E_INVALID;
  cudaSetDevice(w->ctx->device);
  cudaStream_t s = w->ctx->stream;
  W_TRY(w, cudaStreamSynchronize(s));
  *n = 0;
  if (w->n == 0) return CANNON_OK;
  DBuf<int> flag;
  if (flag.reserve(w->n) != cudaSuccess) return fail(w->ctx, CANNON_E_CUDA, "cudaMalloc failed");
  f3 lo, hi;
  lo.x = lower[0]; lo.y = lower[1]; lo.z = lower[2]; hi.x = upper[0]; hi.y = upper[1]; hi.z = upper[2];
  { g_kernel_launches++; k_aabb_query<<<grid_for(w, w->n, 256), 256, 0, s>>>(body_arrays(w), shape_tables(w), w->n, lo, hi, flag.p); }
  std::vector<int> h(w->n);
  cudaError_t e = cudaMemcpyAsync(h.data(), flag.p, (size_t)w->n * sizeof(int), cudaMemcpyDeviceToHost, s);  // the ctx stream does not
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);                                                              // synchronise with the legacy stream
  flag.release();
  if (e != cudaSuccess) return fail(w->ctx, CANNON_E_CUDA, cudaGetErrorString(e));
  int cnt = 0;
  for (int b = 0; b < w->n; b++) if (h[b]) { if (cnt < cap) bodies[cnt] = b; cnt++; }
  *n = cnt;
  return cnt > cap ? fail(w->ctx, CANNON_E_CAPACITY, "body array too small") : CANNON_OK;
}

}  // extern "C"

// cannon_batch_*: host glue over the entry points above, shared by both libraries
#include "batch_impl.inc"
