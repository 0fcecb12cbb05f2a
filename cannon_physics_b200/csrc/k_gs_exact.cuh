// k_gs_exact.cuh — the COLORED solver's production sweep: GSSolver's arithmetic (gs_solver.dart:76-108, f64 on f32-stored
// operands, no FMA, every Vector3 store rounds to float) over the colour order of include/cannon_cuda.h, as a staged
// DATAFLOW kernel.
//
// Two units conflict only if they move the same body; the colouring puts the units of a body into distinct colours, so per
// body there is a fixed sequence of units (ascending colour). Executing every unit after its predecessor ON EACH OF ITS TWO
// BODIES performs exactly the floating-point operations of the colour-by-colour sweep on exactly the same operands - no
// matter how the units of different bodies interleave. The kernel therefore has no grid barrier between colours:
//   * k_schedule hands every unit its rank seq among the deg units of each of its movable bodies;
//   * done[b] counts the units that have updated body b in this solve; a unit of iteration `it` runs when
//     done[b] == it * deg + seq on both bodies, then publishes done[b] + 1 (fence + flag store / acquire load);
//   * a warp owns a window of consecutive units of one colour (one unit per lane) and walks its windows in (iteration,
//     colour) order. That order is a topological order of the dependencies and all warps are co-resident (cooperative
//     launch), so the globally first unfinished window can always run: no deadlock. Every spin is bounded anyway and
//     raises `abort` (reported as an error) instead of hanging the device.
// The only grid barrier left is the one per iteration that the tolerance test of gs_solver.dart:105 needs.
// Rows (96-byte GxRow records, built in execution order), unit records and the rows' multipliers are bulk-copied global ->
// shared with cp.async.bulk one window ahead of the warp, across colour and iteration boundaries (rows never change during
// a solve), exactly like the f32 sweep k_gs_fast.
#pragma once
#include "k_solver.cuh"

#define GX_WARPS 8
#define GX_THREADS (GX_WARPS * 32)
#define GX_CAP_ROWS 104
#define GX_CAP_UNITS 40
#define GX_WIN_MIN 72
#define GX_WIN_MAX 80
#define GX_ROW_BYTES 96
#define GX_UNIT_BYTES 64
#define GX_LAM_REGS ((GX_CAP_ROWS + 31) / 32)
#define GX_UNIT_OFF (GX_CAP_ROWS * GX_ROW_BYTES)
#define GX_LAM_OFF (GX_UNIT_OFF + GX_CAP_UNITS * GX_UNIT_BYTES)
#define GX_LAM_BYTES ((GX_CAP_ROWS + 2) * 8)  // the multiplier range is copied from a 16-byte aligned start: one double of slack each side
#define GX_BUF_BYTES (GX_LAM_OFF + GX_LAM_BYTES)
#define GX_WARP_BYTES (2 * GX_BUF_BYTES)
#define GX_SMEM_BYTES (GX_WARPS * GX_WARP_BYTES)
#define GX_LS_MAX 1024
#define GX_SPIN_LIMIT (1 << 20)

__device__ __forceinline__ double lds_f64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_i32(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_i32(int* p, int v) { asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// grid barrier that gives up when another CTA raised `abort` (a dependency wait timed out): nobody is left spinning
__device__ __forceinline__ bool grid_barrier_abortable(unsigned* bar, unsigned& epoch, int nCtas, const int* abortFlag) {
  __shared__ int s_ab;
  __syncthreads();
  if (threadIdx.x == 0) {
    int ab = 0;
    if (nCtas > 1) {
      epoch += 1;
      const unsigned target = epoch * (unsigned)nCtas;
      unsigned v;
      asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(v) : "l"(bar) : "memory");
      v += 1u;
      int spins = 0;
      while (v < target) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        if ((++spins & 255) == 0 && ld_acquire_i32(abortFlag)) { ab = 1; break; }
        if (spins > GX_SPIN_LIMIT) { ab = 1; break; }
      }
    }
    if (!ab) ab = ld_acquire_i32(abortFlag);
    s_ab = ab;
  }
  __syncthreads();
  return s_ab != 0;
}

struct GxState {
  int* done;    // [nBodies] units that have updated the body in this solve (zeroed before the launch)
  int* abort;   // raised when a bounded wait ran out (reported by the host as an error)
};

__global__ void __launch_bounds__(GX_THREADS, 1) k_gs_exact(RowArrays R, BodyArrays B, UnitArrays U, SchedArrays S, GsTasks T, SolveParams P, GsStats G,
                                                            GxState X) {
  extern __shared__ __align__(128) unsigned char s_dyn[];
  __shared__ unsigned long long s_mbar[GX_WARPS][2];
  __shared__ double s_red[GX_WARPS];
  __shared__ int s_lt[GX_LS_MAX + 2];
  unsigned epoch = 0;
  const int nRows = min(*R.nRows, R.rowCap);
  const int nLevels = *S.nLevels;
  if (*T.nTasks > T.taskCap) return;  // reported through the row-overflow counter by k_gs_task_levels
  // the dataflow needs parallel slack, not one CTA per colour width: every resident CTA takes part unless the solve is tiny
  const int nCtas = coop_ctas(4LL * min(*T.nTasks, T.taskCap) + 1, GX_WARPS);
  if ((int)blockIdx.x >= nCtas) return;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  if (nRows == 0) { if (tid == 0) *G.itersDone = 0; return; }
  const int* lt = T.lvlTask;
  if (nLevels <= GX_LS_MAX) {
    for (int k = threadIdx.x; k <= nLevels; k += blockDim.x) s_lt[k] = T.lvlTask[k];
    lt = s_lt;
  }
  if (lane == 0) { mbar_init(&s_mbar[wic][0], 1); mbar_init(&s_mbar[wic][1], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  unsigned char* wbase = s_dyn + (size_t)wic * GX_WARP_BYTES;
  // windows of a colour are dealt to the warps CTA-interleaved so a narrow colour still spreads over every SM
  const int gw = wic * nCtas + blockIdx.x, nW = nCtas * GX_WARPS;

  GsTask t0, t1, t2;
  int2 x0, y0, x1, y1, x2, y2;  // table entries [a], [a+1] of the three tasks: (first unit, first row)
  auto table = [&](const GsTask& t, int2& x, int2& y) {
    x = make_int2(0, 0); y = x;
    if (t.a >= 0) { x = __ldg(&T.tab[t.a]); y = __ldg(&T.tab[t.a + 1]); }
  };
  auto next = [&](const GsTask& t) {
    GsTask n = t;
    if (n.a >= 0) { n.a += nW; gs_seek(lt, nLevels, P.maxIter, gw, n); }
    return n;
  };
  // unit records, rows and (withLam) the rows' multipliers of a window: three bulk copies on one mbarrier
  auto issue = [&](const int2& x, const int2& y, int b, bool withLam) -> bool {
    const int nUs = min(y.x - x.x, GX_CAP_UNITS), nRs = min(y.y - x.y, GX_CAP_ROWS);
    if (nUs <= 0) return false;
    if (lane == 0) {
      unsigned char* dst = wbase + (size_t)b * GX_BUF_BYTES;
      // the buffer was last written with ordinary shared stores (multipliers of an earlier window): order them before the
      // async-proxy writes of the bulk copies
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const int a0 = x.y & ~1;
      const unsigned lamBytes = (withLam && nRs > 0) ? (unsigned)(((x.y + nRs - a0 + 1) & ~1) * 8) : 0u;
      mbar_expect_tx(&s_mbar[wic][b], (unsigned)(nUs * GX_UNIT_BYTES + nRs * GX_ROW_BYTES) + lamBytes);
      bulk_g2s(dst + GX_UNIT_OFF, U.xrec + x.x, (unsigned)(nUs * GX_UNIT_BYTES), &s_mbar[wic][b]);
      if (nRs > 0) bulk_g2s(dst, R.xrec + x.y, (unsigned)(nRs * GX_ROW_BYTES), &s_mbar[wic][b]);
      if (lamBytes) bulk_g2s(dst + GX_LAM_OFF, R.lambda + a0, lamBytes, &s_mbar[wic][b]);
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
  bool aborted = false;

  int iter = 0;
  for (; iter != P.maxIter && !aborted; iter++) {
    double local = 0.0;
    while (t0.a >= 0 && t0.it == iter) {
      // the warp's next window is this one again (it owns a single window per iteration): its multipliers are handed over
      // shared -> shared after the solve instead of being fetched
      const bool sameNext = t1.a == t0.a;
      pend1 = issue(x1, y1, buf ^ 1, !sameNext);
      t2 = next(t1);
      table(t2, x2, y2);
      const int u0 = x0.x, nU = y0.x - x0.x, rBase = x0.y;
      const int nUs = min(nU, GX_CAP_UNITS), nRs = min(y0.y - x0.y, GX_CAP_ROWS);
      if (pend0) { mbar_wait(&s_mbar[wic][buf], (parity >> buf) & 1u); parity ^= 1u << buf; pend0 = false; }
      const unsigned char* sb = wbase + (size_t)buf * GX_BUF_BYTES;
      const GxUnit* sunits = (const GxUnit*)(sb + GX_UNIT_OFF);
      double* const slam = (double*)(wbase + (size_t)buf * GX_BUF_BYTES + GX_LAM_OFF) + (rBase & 1);
      int flushEnd = nRs;
      for (int ub = 0; ub < nU && !aborted; ub += 32) {
        const int u = ub + lane;
        const bool have = u < nU;
        GxUnit m;
        m.fl = 0; m.r0 = m.r1 = 0; m.bi = m.bj = 0; m.seqA = m.seqB = m.degA = m.degB = 0; m.imA = m.imB = 0.0;
        if (have) m = u < nUs ? sunits[u] : U.xrec[u0 + u];
        const bool work = have && m.r1 > m.r0;
        const bool staged = work && m.r1 - rBase <= nRs;
        if (work && !staged) flushEnd = min(flushEnd, m.r0 - rBase);  // that unit keeps its multipliers in global memory
        // ---- wait for the predecessors of this unit on both of its movable bodies ----
        const int expA = iter * m.degA + m.seqA, expB = iter * m.degB + m.seqB;
        {
          bool rdy = !work;
          int spins = 0;
          while (true) {
            if (!rdy) {
              const int a = (m.fl & 1) ? ld_acquire_i32(X.done + m.bi) : expA;
              const int b = (m.fl & 2) ? ld_acquire_i32(X.done + m.bj) : expB;
              rdy = a == expA && b == expB;
            }
            if (__all_sync(0xffffffffu, rdy)) break;
            __nanosleep(40);
            if ((++spins & 63) == 0) {
              int ab = 0;
              if (lane == 0) ab = ld_acquire_i32(X.abort);
              if (spins > GX_SPIN_LIMIT) ab = 1;
              ab = __any_sync(0xffffffffu, ab);
              if (ab) { if (lane == 0) atomicExch(X.abort, 1); aborted = true; break; }
            }
          }
        }
        if (aborted) break;
        if (work) {
          // a body that is not movable keeps vlambda = wlambda = 0 for the whole solve: do not fetch it
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          float4 vA4 = z4, wA4 = z4, vB4 = z4, wB4 = z4;
          if (m.fl & 1) ldcg_f8(&B.vlam[2 * m.bi], vA4, wA4);
          if (m.fl & 2) ldcg_f8(&B.vlam[2 * m.bj], vB4, wB4);
          f3 vA = ld3(vA4), wA = ld3(wA4), vB = ld3(vB4), wB = ld3(wB4);
          double acc = 0.0;
          // one projected Gauss-Seidel row update (gs_solver.dart:88-102, equation_class.dart:95-105,151-169)
#define GX_ROW_UPDATE(q0, q1, q2, q3, Bv, invC, eps, bound, lam, lamNew, rowIdx)                                             \
          {                                                                                                                  \
            const int code_ = __float_as_int(q0.w);                                                                         \
            f3 n_, rA_, rB_, iA_, iB_, sA_;                                                                                  \
            n_.x = q0.x; n_.y = q0.y; n_.z = q0.z; rA_.x = q1.x; rA_.y = q1.y; rA_.z = q1.z;                                 \
            rB_.x = q2.x; rB_.y = q2.y; rB_.z = q2.z; iB_.x = q3.x; iB_.y = q3.y; iB_.z = q3.z;                              \
            iA_.x = q1.w; iA_.y = q2.w; iA_.z = q3.w;                                                                        \
            if (code_ & 1) { sA_.x = sA_.y = sA_.z = 0.f; } else sA_ = vneg(n_);                                             \
            const double gwl_ = (vdot(vA, sA_) + vdot(wA, rA_)) + (vdot(vB, n_) + vdot(wB, rB_));                            \
            double dl_ = invC * (Bv - gwl_ - eps * lam);                                                                     \
            double mn_, mx_;                                                                                                 \
            const int bc_ = code_ >> 2;                                                                                      \
            if (bc_ == GXB_POS) { mn_ = 0.0; mx_ = bound; }                                                                  \
            else if (bc_ == GXB_SYM) { mn_ = -bound; mx_ = bound; }                                                          \
            else if (bc_ == GXB_NEG) { mn_ = -bound; mx_ = 0.0; }                                                            \
            else { mn_ = R.minF[rowIdx]; mx_ = R.maxF[rowIdx]; }                                                             \
            if (lam + dl_ < mn_) dl_ = mn_ - lam;                                                                            \
            else if (lam + dl_ > mx_) dl_ = mx_ - lam;                                                                       \
            lamNew = lam + dl_;                                                                                              \
            if (m.fl & 1) { vA = vaddscaled(vA, m.imA * dl_, sA_); wA = vaddscaled(wA, dl_, iA_); }                          \
            if (m.fl & 2) { vB = vaddscaled(vB, m.imB * dl_, n_); wB = vaddscaled(wB, dl_, iB_); }                           \
            acc += dl_ > 0.0 ? dl_ : -dl_;                                                                                   \
          }
          if (staged) {
            // rows and multipliers of the window live in shared memory; the next row is fetched before the current one is solved
            unsigned qa = smem_u32(sb) + (unsigned)(m.r0 - rBase) * GX_ROW_BYTES, la = smem_u32(slam) + (unsigned)(m.r0 - rBase) * 8u;
            const int nr = m.r1 - m.r0;
            float4 n0 = lds_f4(qa), n1 = lds_f4(qa + 16), n2 = lds_f4(qa + 32), n3 = lds_f4(qa + 48);
            double nB = lds_f64(qa + 64), nC = lds_f64(qa + 72), nE = lds_f64(qa + 80), nBd = lds_f64(qa + 88);
            double nl = lds_f64(la);
            for (int r = 0; r < nr; r++) {
              const float4 q0 = n0, q1 = n1, q2 = n2, q3 = n3;
              const double Bv = nB, invC = nC, eps = nE, bound = nBd, lam = nl;
              if (r + 1 < nr) {
                qa += GX_ROW_BYTES;
                n0 = lds_f4(qa); n1 = lds_f4(qa + 16); n2 = lds_f4(qa + 32); n3 = lds_f4(qa + 48);
                nB = lds_f64(qa + 64); nC = lds_f64(qa + 72); nE = lds_f64(qa + 80); nBd = lds_f64(qa + 88);
                nl = lds_f64(la + 8u);
              }
              double lamNew;
              GX_ROW_UPDATE(q0, q1, q2, q3, Bv, invC, eps, bound, lam, lamNew, m.r0 + r);
              sts_f64(la, lamNew);
              la += 8u;
            }
          } else {
            for (int r = m.r0; r < m.r1; r++) {
              const float4* q = (const float4*)(R.xrec + r);
              const float4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
              const double* d = (const double*)(q + 4);
              const double Bv = d[0], invC = d[1], eps = d[2], bound = d[3], lam = R.lambda[r];
              double lamNew;
              GX_ROW_UPDATE(q0, q1, q2, q3, Bv, invC, eps, bound, lam, lamNew, r);
              R.lambda[r] = lamNew;
            }
          }
#undef GX_ROW_UPDATE
          if (m.fl & 1) st_f8(&B.vlam[2 * m.bi], st3(vA), st3(wA));
          if (m.fl & 2) st_f8(&B.vlam[2 * m.bj], st3(vB), st3(wB));
          // publish: the body records above become visible before the counters (fence + relaxed store = release)
          __threadfence();
          if (m.fl & 1) st_relaxed_i32(X.done + m.bi, expA + 1);
          if (m.fl & 2) st_relaxed_i32(X.done + m.bj, expB + 1);
          local += acc;
        }
      }
      if (aborted) break;
      flushEnd = __reduce_min_sync(0xffffffffu, flushEnd);
      __syncwarp();
      // multipliers of the staged units back to global as full lines
#pragma unroll
      for (int j = 0; j < GX_LAM_REGS; j++) {
        const int idx = j * 32 + lane;
        if (idx < flushEnd) R.lambda[rBase + idx] = slam[idx];
      }
      asm volatile("fence.proxy.async.global;" ::: "memory");  // a later bulk copy (async proxy) reads these multipliers back
      if (sameNext) {
        double* const nlam = (double*)(wbase + (size_t)(buf ^ 1) * GX_BUF_BYTES + GX_LAM_OFF) + (rBase & 1);
#pragma unroll
        for (int j = 0; j < GX_LAM_REGS; j++) {
          const int idx = j * 32 + lane;
          if (idx < nRs) nlam[idx] = slam[idx];
        }
      }
      __syncwarp();
      t0 = t1; x0 = x1; y0 = y1; pend0 = pend1;
      t1 = t2; x1 = x2; y1 = y2; pend1 = false;
      buf ^= 1;
    }
    if (aborted) break;
    // tolerance test (gs_solver.dart:99-107): the iteration's sum of |delta lambda| over all rows; totals rotate through
    // three slots so a slot can be cleared a full iteration before it is used again
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if (lane == 0) s_red[wic] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int k = 0; k < GX_WARPS; k++) t += s_red[k];
      atomicAdd(&G.worldTot[iter % 3], t);
      if (blockIdx.x == 0) G.worldTot[(iter + 1) % 3] = 0.0;  // last read two barriers ago, next used in iter + 1
    }
    if (grid_barrier_abortable(S.bar, epoch, nCtas, X.abort)) { aborted = true; break; }
    const double tot = __ldcg(&G.worldTot[iter % 3]);
    if (tot * tot < P.tol2) break;
  }
  // a copy started for a window that will never run must land before the CTA may retire
  if (pend0) mbar_wait(&s_mbar[wic][buf], (parity >> buf) & 1u);
  if (pend1) mbar_wait(&s_mbar[wic][buf ^ 1], (parity >> (buf ^ 1)) & 1u);
  if (tid == 0) *G.itersDone = iter;
}
