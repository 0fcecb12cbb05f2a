// k_gs_exact.cuh — the COLORED solver's production sweep: GSSolver's arithmetic (gs_solver.dart:76-108, f64 on f32-stored
// operands, no FMA, every Vector3 store rounds to float) over the colour order of include/cannon_cuda.h.
//
// Persistent cooperative kernel, one grid barrier per colour phase (units of a colour touch disjoint movable bodies, so a
// phase is embarrassingly parallel and the result does not depend on how its units are spread over lanes and warps).
//   * A window = 32 consecutive units of one colour, one unit per lane. Units of a colour are ordered by row count
//     (k_len_*), so the lanes of a window carry equal work. The unit of balance is the SM sub-partition (its warps share
//     one conversion pipe): the windows of a colour, longest first, are dealt to the sub-partitions in rounds, forwards
//     and backwards alternately, rotated per colour.
//   * Rows are stored window-interleaved (k_solver.cuh, GxRow): block r of a window holds row r of each of its 32 units,
//     chunk by chunk, so one row step of a warp reads 3 KB of contiguous memory with coalesced 16-byte accesses. Every lane
//     streams ITS rows global -> shared with cp.async into a private four-slot ring; the window's remaining blocks are
//     prefetched into L2 one phase ahead.
//   * The row loop is software-pipelined: the record of row r + 1 is read from the ring and widened while the dependent
//     chain of row r runs, and the loop body is kept one basic block (no movable-body branches: an immovable body's
//     deltas stay +0 by arithmetic) so the compiler can interleave the two.
//   * One CTA of 16 warps per SM. The grid barrier is split into arrive / wait and the requests for the next phase's rows
//     are issued in between (a release waits for the thread's outstanding loads); arrivals and the phase flag live on
//     different lines, pollers use relaxed loads + one fence. The sparse last colours run on CTA 0 with block barriers.
// What bounds it (tools/ubench/row_step.cu, cvt_rate.cu on a B200): a row update is 27 f32->f64 widenings, 12 f64->f32
// roundings and ~60 FP64 operations per lane. F2F.F64.F32 retires ~3 lanes / clk / sub-partition, so a warp row step costs
// ~350 clk of conversion pipe however many lanes are active (measured: 545 clk for a lone warp, 1412 clk with four warps
// per sub-partition; replacing the conversions by integer or add-magic forms is exact but not faster). The solve is
// therefore conversion-bound at ~0.13 ms for this scene if perfectly packed; the rest of the time is the longest unit of
// every colour phase and ~3 us of barrier + prologue per phase (100-120 phases). The colour order therefore cuts manifolds
// into units of at most CANNON_COLORED_UNIT_CONTACTS = 4 contacts (12 rows on one lane instead of up to 30).
// History on the settled 100k pile of config 3 (1.36e6 rows, 10 colours, 10 iterations; profiles/README.md):
//   unstaged level sweep k_gs 3.43 ms -> TMA-staged per-warp row windows with per-body dataflow counters 1.74 ms (8 warps
//   per SM, 12 of 32 lanes busy: instruction-latency bound) -> 32-unit windows, dynamic claiming, dataflow 2.32 ms (every
//   iteration drains at the tolerance barrier; a window waits for the slowest of its 64 predecessors, i.e. dataflow
//   degenerates to one barrier per colour plus polling traffic) -> barrier per colour, window-interleaved rows, cp.async
//   rings 1.55 ms -> four lanes per unit (quarter of the chain, but 3x the warp instructions: 2.4 ms, rejected) -> split
//   barrier, balanced dealing, tail colours on one CTA, pipelined row loop 1.28 ms -> units of at most 4 contacts 1.04 ms.
#pragma once
#include "k_solver.cuh"

#define GX_WARPS 16
#define GX_THREADS (GX_WARPS * 32)
#define GX_CTAS_PER_SM 1
#define GX_TAIL_WINS GX_WARPS            // colours from which on at most this many windows remain run on CTA 0 only
#define GX_LS_MAX 1024
#define GX_SLOTS 4                       // ring slots per lane: GX_SLOTS - 1 rows in flight
#define GX_SLOT_BYTES (GX_CHUNKS * 32 * 16 + 32 * 8)  // six 16-byte chunks + the row's multiplier (8 bytes) per lane
#define GX_WARP_BYTES (GX_SLOTS * GX_SLOT_BYTES)
#define GX_SMEM_BYTES (GX_WARPS * GX_WARP_BYTES)

__device__ __forceinline__ int4 ldnc_i4(const void* p) {
  int4 v;
  asm volatile("ld.global.nc.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void ldnc_d2(const void* p, double& a, double& b) { asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(p)); }
__device__ __forceinline__ double ldcg_f64(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stcg_f64(double* p, double v) { asm volatile("st.global.cg.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ double lds_f64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void lds_d2(unsigned a, double& x, double& y) { asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a) : "memory"); }
__device__ __forceinline__ void cp_async8_s(unsigned dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async16_s(unsigned dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }

// Grid barrier for one CTA per SM, split into ARRIVE and WAIT so the requests for the next phase's rows can be issued in
// between: a release has to wait for the thread's outstanding loads, and a warp that has just asked for 10 KB of rows from
// DRAM would sit on its arrival for microseconds (measured: 10 - 17 us per phase when all CTAs arrive together with their
// prefetches in flight, 1 us when they do not). Arrivals go to one counter; the last arriver publishes the phase number on a
// second line and everybody else polls that line with relaxed loads (nobody writes it until the phase is over, so the
// arrival atomics never queue behind the pollers), then fences once.
__device__ __forceinline__ bool gx_arrive(unsigned* bar, unsigned& epoch, int nCtas) {  // one thread per CTA; true: this CTA was the last
  epoch += 1;
  unsigned v;
  asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(v) : "l"(bar) : "memory");
  const bool last = v + 1u == epoch * (unsigned)nCtas;
  if (last) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 32), "r"(epoch) : "memory");
  return last;
}
__device__ __forceinline__ void gx_wait(unsigned* bar, unsigned epoch, bool last) {
  if (last) return;
  unsigned g;
  while (true) {
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(bar + 32) : "memory");
    if (g >= epoch) break;
  }
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

// a row of the ring, widened: everything the update of one row needs besides the body deltas
struct GxJ {
  double nx, ny, nz, sAx, sAy, sAz, rAx, rAy, rAz, rBx, rBy, rBz, iAx, iAy, iAz, iBx, iBy, iBz;
  double Bv, invC, eps, lam, mn, mx;
  bool general;  // bounds in RowArrays.minF / maxF
};
__device__ __forceinline__ void gx_read_row(unsigned sa, unsigned sl, GxJ& j) {
  const float4 q0 = lds_f4(sa), q1 = lds_f4(sa + 512), q2 = lds_f4(sa + 1024), q3 = lds_f4(sa + 1536);
  double bound;
  lds_d2(sa + 2048, j.Bv, j.invC);
  lds_d2(sa + 2560, j.eps, bound);
  j.lam = lds_f64(sl);
  const int code = __float_as_int(q0.w);
  j.nx = (double)q0.x; j.ny = (double)q0.y; j.nz = (double)q0.z;
  // sA = -n, or 0 for a rotational row (equation_class.dart:95-105 with the Jacobian of rotational_equation.dart)
  const bool rot = code & 1;
  j.sAx = rot ? 0.0 : -j.nx; j.sAy = rot ? 0.0 : -j.ny; j.sAz = rot ? 0.0 : -j.nz;
  j.rAx = (double)q1.x; j.rAy = (double)q1.y; j.rAz = (double)q1.z;
  j.rBx = (double)q2.x; j.rBy = (double)q2.y; j.rBz = (double)q2.z;
  j.iBx = (double)q3.x; j.iBy = (double)q3.y; j.iBz = (double)q3.z;
  j.iAx = (double)q1.w; j.iAy = (double)q2.w; j.iAz = (double)q3.w;
  const int bc = code >> 2;
  j.general = bc == GXB_GENERAL;
  j.mn = bc == GXB_POS ? 0.0 : -bound;
  j.mx = bc == GXB_NEG ? 0.0 : bound;
}
// Vector3.dot of an f32-stored vector with a widened one, left to right (vec3.dart:62)
__device__ __forceinline__ double gx_dot(const f3& a, double bx, double by, double bz) {
  double s = (double)a.x * bx;
  s += (double)a.y * by;
  s += (double)a.z * bz;
  return s;
}
// Vector3.addScaledVector into an f32-stored vector (vec3.dart:140): a + s * b, every component rounded to float
__device__ __forceinline__ f3 gx_axpy(const f3& a, double s, double bx, double by, double bz) {
  f3 r;
  r.x = (float)((double)a.x + s * bx); r.y = (float)((double)a.y + s * by); r.z = (float)((double)a.z + s * bz);
  return r;
}

struct GxState {
  const int* lvlTask;  // [nLevels + 1] first window of each colour (k_gs_task_levels, windows of 32 units)
};

struct GxWin { int a, lvl, it, k; };  // a: window index inside the iteration, -1 = none; k: the warp's k-th window of colour lvl

__global__ void __launch_bounds__(GX_THREADS, GX_CTAS_PER_SM) k_gs_exact(RowArrays R, BodyArrays B, UnitArrays U, SchedArrays S, SolveParams P, GsStats G,
                                                                         GxState X) {
  extern __shared__ __align__(128) unsigned char s_dyn[];
  __shared__ double s_red[GX_WARPS];
  __shared__ int s_lt[GX_LS_MAX + 2];
  __shared__ int s_trc[2];  // trace: windows and row steps this CTA has swept so far
  int trcW = 0, trcR = 0;
  if (threadIdx.x == 0) s_trc[0] = s_trc[1] = 0;
  unsigned epoch = 0;
  const int nRows = min(*R.nRows, R.rowCap);
  const int nLevels = *S.nLevels;
  const int nTasks = X.lvlTask[nLevels];  // windows per iteration
  // CTAs for about twice the mean number of windows per colour
  const int nCtas = coop_ctas(4LL * nTasks / max(nLevels, 1) + 1, GX_WARPS);
  if ((int)blockIdx.x >= nCtas) return;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  if (nRows == 0 || nTasks == 0) { if (tid == 0) *G.itersDone = 0; return; }
  const int* lt = X.lvlTask;
  if (nLevels <= GX_LS_MAX) {
    for (int k = threadIdx.x; k <= nLevels; k += blockDim.x) s_lt[k] = X.lvlTask[k];
    lt = s_lt;
  }
  __syncthreads();
  const float4* const xblk = R.xblk;
  // this lane's ring: slot s, chunk c at ring + s * GX_SLOT_BYTES + c * 512 (lanes interleaved by 16 B: conflict-free
  // LDS.128); the row's multiplier at ringL + s * GX_SLOT_BYTES (lanes interleaved by 8 B)
  const unsigned ring = smem_u32(s_dyn) + (unsigned)wic * GX_WARP_BYTES + (unsigned)lane * 16u;
  const unsigned ringL = smem_u32(s_dyn) + (unsigned)wic * GX_WARP_BYTES + GX_CHUNKS * 512u + (unsigned)lane * 8u;
  // colours [tailStart, nLevels) hold at most GX_TAIL_WINS windows together (the last colours of a greedy colouring are a
  // handful of units): CTA 0 sweeps them alone with block barriers in between - a grid barrier costs more than their work
  int tailStart = nLevels;
  while (tailStart > 0 && nTasks - lt[tailStart - 1] <= GX_TAIL_WINS) tailStart--;
  if (nCtas == 1) tailStart = 0;
  // Dealing the windows of a wide colour. Warps w, w + 4, ... of a CTA share an SM sub-partition (issue port, F2F
  // and FP64 pipes): the unit of balance is the sub-partition, nS of them. The windows of a colour are sorted longest
  // first and dealt in rounds of nS, forwards in even rounds and backwards in odd ones (sub-partition s gets windows s,
  // 2 nS - 1 - s, 2 nS + s, ...: every sub-partition receives nearly the same number of rows), the start rotated per colour.
  // Round rho of a sub-partition is taken by its warp rho % (GX_WARPS / 4). Consecutive positions lie on different SMs.
  const int nS = nCtas * 4, sid = (wic & 3) * nCtas + (int)blockIdx.x, sv = wic >> 2;
  auto win_of = [&](int lvl, int k) -> int {  // the warp's k-th window of colour lvl, -1 past the end
    const int n = lt[lvl + 1] - lt[lvl];
    int j;
    if (lvl >= tailStart) {
      if (blockIdx.x != 0) return -1;
      j = wic + GX_WARPS * k;
    } else {
      const int rho = sv + (GX_WARPS / 4) * k;
      const int s = (sid + nS - (lvl * 61) % nS) % nS;
      j = rho * nS + ((rho & 1) ? nS - 1 - s : s);
    }
    return j < n ? lt[lvl] + j : -1;
  };
  auto seek = [&](GxWin& t) {  // (t.lvl, t.it, t.k) is a candidate: move on to the warp's next existing window
    int hops = 0;
    while ((t.a = win_of(t.lvl, t.k)) < 0) {
      if (++hops > nLevels) return;  // the warp owns no window in any colour
      t.k = 0;
      if (++t.lvl == nLevels) { t.lvl = 0; if (++t.it >= P.maxIter) return; }
    }
  };

  GxUnit m;         // unit of this lane in the warp's current window
  bool act = false; // the lane has a unit with rows
  // unit record of window t into registers + the first GX_SLOTS - 1 rows into the ring (GX_SLOTS - 1 groups committed)
  auto prime = [&](const GxWin& t) {
    act = false;
    m.fl = 0; m.r0 = m.r1 = 0; m.bi = m.bj = 0; m.imA = m.imB = 0.0;
    if (t.a >= 0) {
      const int a0 = S.levelStart[t.lvl], a1 = S.levelStart[t.lvl + 1];
      const int u = a0 + 32 * (t.a - lt[t.lvl]) + lane;
      if (u < a1) {
        const int4* up = (const int4*)(U.xrec + u);
        const int4 a = ldnc_i4(up), b = ldnc_i4(up + 1);
        m.bi = a.x; m.bj = a.y; m.fl = a.z; m.r0 = a.w; m.r1 = b.x;
        ldnc_d2(up + 3, m.imA, m.imB);
        act = m.r1 > m.r0;
      }
    }
    const int nr = act ? (m.r1 - m.r0) >> 5 : 0;
    // The ring keeps three rows in flight per lane, which hides L2 latency but not DRAM latency (a lane walks its rows one
    // after the other). So the rest of the window's blocks are pulled into L2 right away - this runs before the grid
    // barrier of the previous phase, i.e. the DRAM time of a phase's rows overlaps the phase before it.
    int nrMax = nr, blk0 = act ? m.r0 >> 5 : 0;
    for (int o = 16; o > 0; o >>= 1) { nrMax = max(nrMax, __shfl_xor_sync(0xffffffffu, nrMax, o)); blk0 = max(blk0, __shfl_xor_sync(0xffffffffu, blk0, o)); }
    {
      const unsigned char* q = (const unsigned char*)(xblk + (size_t)blk0 * (GX_CHUNKS * 32));
      const int nb = nrMax * GX_CHUNKS * 512;
      for (int o = (GX_SLOTS - 1) * GX_CHUNKS * 512 + lane * 128; o < nb; o += 32 * 128) prefetch_l2(q + o);
      const unsigned char* ql = (const unsigned char*)(R.lambda + (size_t)blk0 * 32);  // the rows' multipliers: 256 B per block
      for (int o = lane * 128; o < nrMax * 256; o += 32 * 128) prefetch_l2(ql + o);
    }
#pragma unroll
    for (int r = 0; r < GX_SLOTS - 1; r++) {
      if (r < nr) {
        const float4* q = xblk + (size_t)((m.r0 >> 5) + r) * (GX_CHUNKS * 32) + lane;
        const unsigned d = ring + (unsigned)(r * GX_SLOT_BYTES);
#pragma unroll
        for (int c = 0; c < GX_CHUNKS; c++) cp_async16_s(d + c * 512, q + c * 32);
        cp_async8_s(ringL + (unsigned)(r * GX_SLOT_BYTES), R.lambda + m.r0 + 32 * r);
      }
      cp_async_commit();
    }
  };

  // row r of the lane's current unit (and its multiplier) into ring slot r % GX_SLOTS; commits a group either way
  // row r of the lane's current unit (and its multiplier) into ring slot r % GX_SLOTS; commits a group either way.
  // (Measured: a branch-free form - copies with a source size of zero past the last row - is 10 % slower than this branch;
  // three MORE uniform branches in the row loop doubled the sweep's time: the loop lives on the compiler interleaving the
  // next row's loads and widenings with the current row's dependent chain inside one basic block.)
  auto gx_request = [&](int r) {
    if (r < ((m.r1 - m.r0) >> 5)) {
      const float4* q = xblk + (size_t)((m.r0 >> 5) + r) * (GX_CHUNKS * 32) + lane;
      const unsigned dso = (unsigned)((r % GX_SLOTS) * GX_SLOT_BYTES), d = ring + dso;
#pragma unroll
      for (int c = 0; c < GX_CHUNKS; c++) cp_async16_s(d + c * 512, q + c * 32);
      cp_async8_s(ringL + dso, R.lambda + m.r0 + 32 * r);
    }
    cp_async_commit();
  };
  auto prime_all = [&](const GxWin& t) {
    prime(t);
    // unit records of the window after that one towards L2 (32 x 64 B = 16 lines)
    GxWin t2 = t;
    if (t2.a >= 0) { t2.k++; seek(t2); }
    if (t2.a >= 0 && lane < 16) prefetch_l2((const unsigned char*)(U.xrec + S.levelStart[t2.lvl] + 32 * (t2.a - lt[t2.lvl])) + lane * 128);
  };
  GxWin t;
  t.lvl = 0; t.it = 0; t.k = 0; t.a = -1;
  seek(t);
  prime(t);
  bool pend = false;  // the next window's rows are still to be requested

  int iter = 0;
  int trN = 0;
  for (; iter != P.maxIter; iter++) {
    double local = 0.0;
    for (int lvl = 0; lvl < nLevels; lvl++) {
      GS_TRACE_BEGIN();
      while (t.a >= 0 && t.lvl == lvl && t.it == iter) {
        const bool tr = P.trace && wic == 0 && iter == 1 && lvl == 0 && trN < 4;
        long long tk0 = 0, tk1 = 0, tk2 = 0;
        if (tr) tk0 = clock64();
        if (act) {
          // a body that is not movable keeps vlambda = wlambda = 0 for the whole solve: it is never fetched
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          float4 vA4 = z4, wA4 = z4, vB4 = z4, wB4 = z4;
          if (m.fl & 1) ldcg_f8(&B.vlam[2 * m.bi], vA4, wA4);
          if (m.fl & 2) ldcg_f8(&B.vlam[2 * m.bj], vB4, wB4);
          double* lp = R.lambda + m.r0;  // slot of row r: m.r0 + 32 r
          if (tr) { asm volatile("" ::"f"(vA4.x), "f"(wA4.x), "f"(vB4.x), "f"(wB4.x)); tk2 = clock64() - tk0; tk1 = 0; }
          f3 vA = ld3(vA4), wA = ld3(wA4), vB = ld3(vB4), wB = ld3(wB4);
          double acc = 0.0;
          const int nr = (m.r1 - m.r0) >> 5;
          const double imA = m.imA, imB = m.imB;
          // Software pipeline over the rows of the unit. The dependent chain of a row update (widen the body deltas, four dot
          // products, delta lambda, clamp, four axpys, round) is ~20 f64 / conversion operations long; the 15 widenings of
          // the row's own Jacobian and its shared-memory reads do not depend on the previous row, so the NEXT row's record
          // is read and widened in the shadow of this row's chain (measured with tools/ubench/row_step: a lone warp needs
          // 545 cycles per row when everything sits on the chain; what a colour phase costs is its longest unit).
          // An immovable body has invMassSolve = 0 and I^-1 r = 0, so its (never stored) deltas stay +0 without a branch.
          GxJ cur, nxt;
          asm volatile("cp.async.wait_group %0;" ::"n"(GX_SLOTS - 2) : "memory");  // row 0 has landed
          gx_read_row(ring, ringL, cur);
          gx_request(3);  // rows 0 .. GX_SLOTS - 2 were requested with the unit record (prime)
          for (int r = 0; r < nr; r++) {
            asm volatile("cp.async.wait_group %0;" ::"n"(GX_SLOTS - 2) : "memory");  // row r + 1 has landed (garbage past the end: never used)
            const unsigned so = (unsigned)(((r + 1) % GX_SLOTS) * GX_SLOT_BYTES);
            gx_read_row(ring + so, ringL + so, nxt);
            gx_request(r + GX_SLOTS);  // into the slot of row r, which was read one step ago
            // one projected Gauss-Seidel row update (gs_solver.dart:88-102, equation_class.dart:95-105,151-169)
            const double gwl = (gx_dot(vA, cur.sAx, cur.sAy, cur.sAz) + gx_dot(wA, cur.rAx, cur.rAy, cur.rAz)) +
                               (gx_dot(vB, cur.nx, cur.ny, cur.nz) + gx_dot(wB, cur.rBx, cur.rBy, cur.rBz));
            double dl = cur.invC * (cur.Bv - gwl - cur.eps * cur.lam);
            double mn = cur.mn, mx = cur.mx;
            if (cur.general) { mn = R.minF[m.r0 + 32 * r]; mx = R.maxF[m.r0 + 32 * r]; }
            if (cur.lam + dl < mn) dl = mn - cur.lam;
            else if (cur.lam + dl > mx) dl = mx - cur.lam;
            lp[32 * r] = cur.lam + dl;  // plain store: the same lane reads it back through L1 (cp.async.ca) one iteration later
            const double dA = imA * dl, dB = imB * dl;
            vA = gx_axpy(vA, dA, cur.sAx, cur.sAy, cur.sAz); wA = gx_axpy(wA, dl, cur.iAx, cur.iAy, cur.iAz);
            vB = gx_axpy(vB, dB, cur.nx, cur.ny, cur.nz); wB = gx_axpy(wB, dl, cur.iBx, cur.iBy, cur.iBz);
            acc += dl > 0.0 ? dl : -dl;
            cur = nxt;
          }
          if (m.fl & 1) st_f8(&B.vlam[2 * m.bi], st3(vA), st3(wA));
          if (m.fl & 2) st_f8(&B.vlam[2 * m.bj], st3(vB), st3(wB));
          local += acc;
        }
        __syncwarp();
        if (P.trace) {
          const long long tend = clock64();
          int nrm = act ? (m.r1 - m.r0) >> 5 : 0;
          for (int o = 16; o > 0; o >>= 1) {
            nrm = max(nrm, __shfl_xor_sync(0xffffffffu, nrm, o)); tk1 = max(tk1, __shfl_xor_sync(0xffffffffu, tk1, o));
            tk2 = max(tk2, __shfl_xor_sync(0xffffffffu, tk2, o));
          }
          if (lane == 0) { atomicAdd(&s_trc[0], 1); atomicAdd(&s_trc[1], nrm); }
          if (lane == 0 && tr) {  // (body gather, whole window, of which waiting for rows, 32000 + rows)
            long long* o = P.trace + (size_t)gridDim.x * GS_TRACE_PHASES * 2 + (blockIdx.x * 4 + trN) * 4;
            o[0] = tk2; o[1] = tend - tk0; o[2] = tk1; o[3] = 32 * 1000 + nrm;
          }
          if (tr) trN++;
        }
        // the warp's next window: requested right away if it belongs to this phase, else after the arrival at the barrier
        t.k++;
        seek(t);
        if (t.a >= 0 && t.lvl == lvl && t.it == iter) prime_all(t); else pend = true;
      }
      if (lvl == nLevels - 1) {
        // the iteration's |delta lambda| total rides on the last colour barrier; totals rotate through three slots so a slot
        // can be cleared a full iteration before it is used again, without an extra barrier
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if (lane == 0) s_red[wic] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
          double tt = 0.0;
          for (int w = 0; w < GX_WARPS; w++) tt += s_red[w];
          atomicAdd(&G.worldTot[iter % 3], tt);
        }
      }
      GS_TRACE_WORK();
      // wide colours end with a grid barrier; inside the tail only CTA 0 works, a block barrier separates its colours
      const bool grid = nCtas > 1 && (lvl < tailStart || lvl == nLevels - 1);
      __syncthreads();
      bool last = false;
      if (grid && threadIdx.x == 0) last = gx_arrive(S.bar, epoch, nCtas);
      if (pend) { prime_all(t); pend = false; }
      if (grid) {
        if (threadIdx.x == 0) gx_wait(S.bar, epoch, last);
        __syncthreads();
      }
      if (P.trace && threadIdx.x == 0) {
        const int ph = iter * nLevels + lvl;
        if (ph < GS_TRACE_PHASES) {  // low words: cycles of work / of barrier wait; high words: windows / row steps of this CTA in the phase
          const int cw = s_trc[0], cr = s_trc[1];
          P.trace[(blockIdx.x * GS_TRACE_PHASES + ph) * 2] = ((trW - trS) & 0xffffffffLL) | ((long long)(cw - trcW) << 32);
          P.trace[(blockIdx.x * GS_TRACE_PHASES + ph) * 2 + 1] = ((clock64() - trW) & 0xffffffffLL) | ((long long)(cr - trcR) << 32);
          trcW = cw; trcR = cr;
        }
      }
      if (lvl == 0 && tid == 0) G.worldTot[(iter + 2) % 3] = 0.0;  // last read before this barrier, next used in iter + 2
    }
    // tolerance test (gs_solver.dart:99-107)
    const double tot = __ldcg(&G.worldTot[iter % 3]);
    if (tot * tot < P.tol2) break;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");  // rows requested for a window that will never run must land before the CTA retires
  if (tid == 0) *G.itersDone = iter;
}
