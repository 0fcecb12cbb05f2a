/*
 * oracle_solver.cpp — TEST INFRASTRUCTURE ONLY (PARITY UNPINNED, see oracle_math.h).
 * Equation assembly, Gauss-Seidel solve, integration and sleeping restated from
 *   lib/world/world_class.dart:433-701           World.internalStep
 *   lib/solver/gs_solver.dart:27-133             GSSolver.solve
 *   lib/solver/solver.dart:30-34                 Solver.addEquation filter
 *   lib/equations/equation_class.dart:53-174     SPOOK, computeGW/GiMf/GiMGt/addToWlambda
 *   lib/equations/{contact,friction,rotational,rotational_motor}_equation.dart  computeB
 *   lib/constraints/{point_to_point,hinge}_constraint.dart  update()
 *   lib/objects/rigid_body.dart:263-314,627-680  sleep FSM, solve mass, integrate
 */
#include <algorithm>
#include <cmath>
#include <unordered_map>
#include <utility>

#include "oracle_world.h"

namespace orc {

namespace {

// Equation.computeGW, equation_class.dart:80-92
double computeGW(const Eq& e, const Body& bi, const Body& bj) {
  return (dot(bi.velocity, e.sA) + dot(bi.angularVelocity, e.rA)) + (dot(bj.velocity, e.sB) + dot(bj.angularVelocity, e.rB));
}
// Equation.computeGWlambda, equation_class.dart:95-105
double computeGWlambda(const Eq& e, const Body& bi, const Body& bj) {
  return (dot(bi.vlambda, e.sA) + dot(bi.wlambda, e.rA)) + (dot(bj.vlambda, e.sB) + dot(bj.wlambda, e.rB));
}
// Equation.computeGiMf, equation_class.dart:108-127
double computeGiMf(const Eq& e, const Body& bi, const Body& bj) {
  V3 iMfi = scale(bi.invMassSolve, bi.force);
  V3 iMfj = scale(bj.invMassSolve, bj.force);
  V3 invIiVmultTaui = mvmult(bi.invInertiaWorldSolve, bi.torque);
  V3 invIjVmultTauj = mvmult(bj.invInertiaWorldSolve, bj.torque);
  return (dot(iMfi, e.sA) + dot(invIiVmultTaui, e.rA)) + (dot(iMfj, e.sB) + dot(invIjVmultTauj, e.rB));
}
// Equation.computeGiMGt + computeC, equation_class.dart:130-148,172-174
double computeC(const Eq& e, const Body& bi, const Body& bj) {
  double result = bi.invMassSolve + bj.invMassSolve;
  V3 tmp = mvmult(bi.invInertiaWorldSolve, e.rA);
  result += dot(tmp, e.rA);
  tmp = mvmult(bj.invInertiaWorldSolve, e.rB);
  result += dot(tmp, e.rB);
  return result + e.eps;
}
// Equation.addToWlambda, equation_class.dart:151-169
void addToWlambda(const Eq& e, Body& bi, Body& bj, double dl) {
  bi.vlambda = add_scaled(bi.vlambda, bi.invMassSolve * dl, e.sA);
  bj.vlambda = add_scaled(bj.vlambda, bj.invMassSolve * dl, e.sB);
  V3 temp = mvmult(bi.invInertiaWorldSolve, e.rA);
  bi.wlambda = add_scaled(bi.wlambda, dl, temp);
  temp = mvmult(bj.invInertiaWorldSolve, e.rB);
  bj.wlambda = add_scaled(bj.wlambda, dl, temp);
}

double computeB(Eq& e, const Body& bi, const Body& bj, double h) {
  switch (e.kind) {
    case EQ_CONTACT: {  // contact_equation.dart:34-77
      V3 rixn = cross(e.ri, e.ni);
      V3 rjxn = cross(e.rj, e.ni);
      e.sA = neg(e.ni);
      e.rA = neg(rixn);
      e.sB = e.ni;
      e.rB = rjxn;
      V3 pen = bj.position;
      pen = add(pen, e.rj);
      pen = sub(pen, bi.position);
      pen = sub(pen, e.ri);
      double g = dot(e.ni, pen);
      double ePlusOne = e.restitution + 1;
      double gw = ePlusOne * dot(bj.velocity, e.ni) - ePlusOne * dot(bi.velocity, e.ni) + dot(bj.angularVelocity, rjxn) -
                  dot(bi.angularVelocity, rixn);
      double giMf = computeGiMf(e, bi, bj);
      return -g * e.a - gw * e.b - h * giMf;
    }
    case EQ_FRICTION: {  // friction_equation.dart:19-47
      V3 rixt = cross(e.ri, e.ni);
      V3 rjxt = cross(e.rj, e.ni);
      e.sA = neg(e.ni);
      e.rA = neg(rixt);
      e.sB = e.ni;
      e.rB = rjxt;
      double gw = computeGW(e, bi, bj);
      double giMf = computeGiMf(e, bi, bj);
      return -gw * e.b - h * giMf;
    }
    case EQ_ROTATIONAL: {  // rotational_equation.dart:34-58
      V3 nixnj = cross(e.axisA, e.axisB);
      V3 njxni = cross(e.axisB, e.axisA);
      e.rA = njxni;
      e.rB = nixnj;
      double g = std::cos(e.maxAngle) - dot(e.axisA, e.axisB);
      double gW = computeGW(e, bi, bj);
      double giMf = computeGiMf(e, bi, bj);
      return -g * e.a - gW * e.b - h * giMf;
    }
    default: {  // rotational_motor_equation.dart:17-33
      e.rA = e.axisA;
      e.rB = neg(e.axisB);
      double gw = computeGW(e, bi, bj) - e.targetVelocity;
      double giMf = computeGiMf(e, bi, bj);
      return -gw * e.b - h * giMf;
    }
  }
}

// Body.updateSolveMassProperties, rigid_body.dart:303-314
void updateSolveMassProperties(Body& b) {
  if (b.sleepState == CANNON_SLEEPING || b.type == CANNON_BODY_KINEMATIC) {
    b.invMassSolve = 0;
    b.invInertiaWorldSolve = M3{{0, 0, 0, 0, 0, 0, 0, 0, 0}};
  } else {
    b.invMassSolve = b.invMass;
    b.invInertiaWorldSolve = b.invInertiaWorld;
  }
}

// GSSolver.solve, gs_solver.dart:27-133, over `eqs` and the bodies [b0,b1)
int gsSolve(World& w, std::vector<Eq*>& eqs, const std::vector<int>& bodyIdx, double h, std::vector<RowDebug>& dbg) {
  int iter = 0;
  const int maxIter = w.desc.solver_iterations;
  const double tolSquared = w.desc.solver_tolerance * w.desc.solver_tolerance;
  const int nEq = (int)eqs.size();
  std::vector<Body>& bodies = w.bodies;
  if (nEq != 0)
    for (int i : bodyIdx) updateSolveMassProperties(bodies[i]);
  std::vector<double> invCs(nEq), bs(nEq), lambda(nEq);
  for (int i = 0; i != nEq; i++) {
    Eq& c = *eqs[i];
    lambda[i] = 0.0;
    bs[i] = computeB(c, bodies[c.bi], bodies[c.bj], h);
    invCs[i] = 1.0 / computeC(c, bodies[c.bi], bodies[c.bj]);
  }
  if (nEq != 0) {
    for (int i : bodyIdx) {
      bodies[i].vlambda = V3{0, 0, 0};
      bodies[i].wlambda = V3{0, 0, 0};
    }
    for (iter = 0; iter != maxIter; iter++) {
      double deltalambdaTot = 0.0;
      for (int j = 0; j != nEq; j++) {
        Eq& c = *eqs[j];
        double B = bs[j], invC = invCs[j], lambdaj = lambda[j];
        double gwlambda = computeGWlambda(c, bodies[c.bi], bodies[c.bj]);
        double deltalambda = invC * (B - gwlambda - c.eps * lambdaj);
        if (lambdaj + deltalambda < c.minForce) deltalambda = c.minForce - lambdaj;
        else if (lambdaj + deltalambda > c.maxForce) deltalambda = c.maxForce - lambdaj;
        lambda[j] = lambda[j] + deltalambda;
        deltalambdaTot += deltalambda > 0.0 ? deltalambda : -deltalambda;
        addToWlambda(c, bodies[c.bi], bodies[c.bj], deltalambda);
      }
      if (deltalambdaTot * deltalambdaTot < tolSquared) break;
    }
    for (int i : bodyIdx) {
      Body& b = bodies[i];
      b.vlambda = mulc(b.vlambda, b.linearFactor);
      b.velocity = add(b.vlambda, b.velocity);
      b.wlambda = mulc(b.wlambda, b.angularFactor);
      b.angularVelocity = add(b.wlambda, b.angularVelocity);
    }
    double invDt = 1 / h;
    for (int l = nEq - 1; l > -1; l--) eqs[l]->multiplier = lambda[l] * invDt;
  }
  for (int i = 0; i != nEq; i++) dbg.push_back(RowDebug{eqs[i]->bi, eqs[i]->bj, bs[i], invCs[i], lambda[i]});
  return iter;
}

}  // namespace

// world_class.dart:543-624: per-contact restitution override (idempotent with createContactEquation's
// rule because shape materials are out of scope), wake-up flags, then wake flagged bodies
// OverlapKeeper.set (overlap_keeper.dart:20-37): insertion into the sorted key list, duplicates ignored. The reference
// key is (i << 16) | j with i < j; 32 bits per id keep it injective beyond 65535 bodies with the same order below.
void World::overlapSet(int i, int j) {
  if (j < i) { const int t = j; j = i; i = t; }
  const int64_t key = ((int64_t)i << 32) | (int64_t)j;
  size_t index = 0;
  while (index < overlapCurrent.size() && key > overlapCurrent[index]) index++;
  if (index < overlapCurrent.size() && key == overlapCurrent[index]) return;
  overlapCurrent.insert(overlapCurrent.begin() + (long)index, key);
}

// OverlapKeeper.getDiff (overlap_keeper.dart:48-82): keys of this step missing from the previous one, then the reverse
void World::emitContactEvents() {
  additions.clear();
  removals.clear();
  const std::vector<int64_t>&a = overlapCurrent, &b = overlapPrevious;
  size_t j = 0;
  for (size_t i = 0; i < a.size(); i++) {
    const int64_t keyA = a[i];
    while (j < b.size() && keyA > b[j]) j++;
    const bool found = j < b.size() && keyA == b[j];
    if (!found) { additions.push_back((int)(keyA >> 32)); additions.push_back((int)(keyA & 0xffffffffll)); }
  }
  j = 0;
  for (size_t i = 0; i < b.size(); i++) {
    const int64_t keyB = b[i];
    while (j < a.size() && keyB > a[j]) j++;
    const bool found = j < a.size() && a[j] == keyB;
    if (!found) { removals.push_back((int)(keyB >> 32)); removals.push_back((int)(keyB & 0xffffffffll)); }
  }
}

void World::makeContactConstraints() {
  if (trackOverlaps) {  // collisionMatrixTick, world_class.dart:219 (bodyOverlapKeeper.tick, overlap_keeper.dart:40-45)
    overlapCurrent.swap(overlapPrevious);
    overlapCurrent.clear();
    // bodyOverlapKeeper.set of the justTest pairs happened inside getContacts, after the tick (world_class.dart:500,519; narrow_phase.dart:715)
    for (size_t k = 0; k + 1 < justTestOverlaps.size(); k += 2) overlapSet(justTestOverlaps[k], justTestOverlaps[k + 1]);
  }
  for (Eq& c : contacts) {
    Body& bi = bodies[c.bi];
    Body& bj = bodies[c.bj];
    if (trackOverlaps) overlapSet(c.bi, c.bj);  // world_class.dart:606
    if (bi.material >= 0 && bj.material >= 0) {
      if (matRestitution[bi.material] >= 0 && matRestitution[bj.material] >= 0)
        c.restitution = matRestitution[bi.material] * matRestitution[bj.material];
    }
    if (bi.allowSleep && bi.type == CANNON_BODY_DYNAMIC && bi.sleepState == CANNON_SLEEPING && bj.sleepState == CANNON_AWAKE &&
        bj.type != CANNON_BODY_STATIC) {
      double speedSquaredB = length2(bj.velocity) + length2(bj.angularVelocity);
      double speedLimitSquaredB = bj.sleepSpeedLimit * bj.sleepSpeedLimit;
      if (speedSquaredB >= speedLimitSquaredB * 2) bi.wakeUpAfterNarrowphase = true;
    }
    if (bj.allowSleep && bj.type == CANNON_BODY_DYNAMIC && bj.sleepState == CANNON_SLEEPING && bi.sleepState == CANNON_AWAKE &&
        bi.type != CANNON_BODY_STATIC) {
      double speedSquaredA = length2(bi.velocity) + length2(bi.angularVelocity);
      double speedLimitSquaredA = bi.sleepSpeedLimit * bi.sleepSpeedLimit;
      if (speedSquaredA >= speedLimitSquaredA * 2) bj.wakeUpAfterNarrowphase = true;
    }
  }
  if (trackOverlaps) emitContactEvents();  // world_class.dart:610
  for (Body& b : bodies) {
    if (b.wakeUpAfterNarrowphase) {
      b.sleepState = CANNON_AWAKE;
      b.wakeUpAfterNarrowphase = false;
    }
  }
}

int World::solve(double h) {
  // constraints: c.update(); then their equations (world_class.dart:627-635)
  for (Constraint& c : constraints) {
    const Body& A = bodies[c.bodyA];
    const Body& B = bodies[c.bodyB];
    if (c.type == CANNON_CONSTRAINT_DISTANCE) {  // DistanceConstraint.update, distance_constraint.dart:27-38
      Eq& e = c.eqs[0];
      const double halfDist = c.distance * 0.5;
      V3 normal = sub(B.position, A.position);
      normalize(normal);
      e.ni = normal;
      e.ri = scale(halfDist, normal);
      e.rj = scale(-halfDist, normal);
      continue;
    }
    // PointToPointConstraint.update, point_to_point_constraint.dart:68-83
    V3 ri = qvmult(A.quaternion, c.pivotA);
    V3 rj = qvmult(B.quaternion, c.pivotB);
    for (int k = 0; k < 3; k++) {
      c.eqs[k].ri = ri;
      c.eqs[k].rj = rj;
    }
    if (c.type == CANNON_CONSTRAINT_HINGE) {  // hinge_constraint.dart:79-104
      V3 worldAxisA = qvmult(A.quaternion, c.axisA);
      V3 worldAxisB = qvmult(B.quaternion, c.axisB);
      tangents(worldAxisA, c.eqs[3].axisA, c.eqs[4].axisA);
      c.eqs[3].axisB = worldAxisB;
      c.eqs[4].axisB = worldAxisB;
      if (c.eqs[5].enabled) {
        c.eqs[5].axisA = qvmult(A.quaternion, c.axisA);
        c.eqs[5].axisB = qvmult(B.quaternion, c.axisB);
      }
    } else if (c.type == CANNON_CONSTRAINT_LOCK || c.type == CANNON_CONSTRAINT_CONE_TWIST) {
      // lock_constraint.dart:79-87 / cone_twist_constraint.dart:84-94: body-local axes into the world frame
      for (size_t k = 3; k < c.eqs.size(); k++) {
        c.eqs[k].axisA = qvmult(A.quaternion, c.locA[k]);
        c.eqs[k].axisB = qvmult(B.quaternion, c.locB[k]);
      }
    }
  }
  auto accept = [&](const Eq& e) { return e.enabled && !bodies[e.bi].isTrigger && !bodies[e.bj].isTrigger; };
  rows.clear();
  int itersMax = 0;
  const int nW = desc.n_worlds > 1 ? desc.n_worlds : 1;
  if (desc.solver_kind == CANNON_SOLVER_SPLIT) {
    // SplitSolver.solve, split_solver.dart:50-120. Equation ids = creation order of a pool-less step (constraint
    // equations are constructed with their constraint, then per contact: ContactEquation, FrictionEquation x2);
    // each island is solved in descending id order (sortById, :167-169).
    std::vector<Eq*> byId;
    for (Constraint& c : constraints)
      for (Eq& e : c.eqs) byId.push_back(&e);
    {
      size_t f = 0;
      for (Eq& c : contacts) {
        byId.push_back(&c);
        if (c.friction > 0) { byId.push_back(&frictions[f]); byId.push_back(&frictions[f + 1]); f += 2; }
      }
    }
    std::vector<Eq*> acc;
    for (Eq* e : byId) if (accept(*e)) acc.push_back(e);
    const int nB = (int)bodies.size();
    // islands: non-static bodies connected through equations (static bodies are never visited, :122-131)
    std::vector<int> parent(nB);
    for (int i = 0; i < nB; i++) parent[i] = i;
    auto find = [&](int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    for (Eq* e : acc) {
      if (bodies[e->bi].type == CANNON_BODY_STATIC || bodies[e->bj].type == CANNON_BODY_STATIC) continue;
      int a = find(e->bi), b = find(e->bj);
      if (a != b) parent[std::max(a, b)] = std::min(a, b);
    }
    std::vector<std::vector<int>> islandBodies(nB);
    std::vector<std::vector<Eq*>> islandEqs(nB);
    for (int i = 0; i < nB; i++) if (bodies[i].type != CANNON_BODY_STATIC) islandBodies[find(i)].push_back(i);
    for (int k = (int)acc.size() - 1; k >= 0; k--) {  // descending id
      Eq* e = acc[k];
      int root = bodies[e->bi].type != CANNON_BODY_STATIC ? find(e->bi) : (bodies[e->bj].type != CANNON_BODY_STATIC ? find(e->bj) : -1);
      if (root >= 0) islandEqs[root].push_back(e);
    }
    // debug rows are reported in the batch-wide descending-id order
    std::vector<int> pos(acc.size());
    int nIslands = 0;
    std::vector<RowDebug> all;
    std::unordered_map<Eq*, RowDebug> tagged;
    for (int r = 0; r < nB; r++) {
      if (islandBodies[r].empty()) continue;
      nIslands++;
      std::vector<RowDebug> dbg;
      int it = gsSolve(*this, islandEqs[r], islandBodies[r], h, dbg);
      if (it > itersMax) itersMax = it;
      for (size_t k = 0; k < dbg.size(); k++) tagged[islandEqs[r][k]] = dbg[k];
    }
    for (int k = (int)acc.size() - 1; k >= 0; k--) {
      auto it = tagged.find(acc[k]);
      if (it != tagged.end()) rows.push_back(it->second);
    }
    prof.n_islands = nIslands;
  } else if (desc.solver_kind == CANNON_SOLVER_COLORED || desc.solver_kind == CANNON_SOLVER_COLORED_F32) {
    // COLORED order (include/cannon_cuda.h): GSSolver's arithmetic over a different, fully specified equation order.
    // Units = contact manifolds (the ContactEquations of one resolver call; rows [f1, f2, n] per contact) in runs of at most
    // CANNON_COLORED_UNIT_CONTACTS contacts, and one unit per
    // constraint; colour(u) = round in which u holds the smallest pending priority on all of its movable bodies; colours
    // ascend, units of a colour are independent (ordered by key here). Sequential GS over that list is the coloured sweep.
    // In a batch the hashed key is counted inside the unit's world (contacts from the world's first ContactEquation,
    // constraints from the world's contact count), so a world's colours do not depend on the rest of the batch.
    struct CUnit { int bi, bj, key, level; unsigned pri; std::vector<Eq*> eqs; int priKey; };
    const int nWk = std::max(1, desc.n_worlds);
    std::vector<int> wFirstC(nWk, 0x7fffffff), wCountC(nWk, 0), wFirstSlot(nWk, 0x7fffffff);
    auto worldOf = [&](int b) { return nWk > 1 ? bodies[b].worldId : 0; };
    for (size_t c = 0; c < contacts.size(); c++) {
      const int wi = worldOf(contacts[c].bi);
      wFirstC[wi] = std::min(wFirstC[wi], (int)c);
      wCountC[wi]++;
    }
    std::vector<CUnit> units;
    auto movable = [&](int b) {
      const Body& B = bodies[b];
      if (B.sleepState == CANNON_SLEEPING || B.type == CANNON_BODY_KINEMATIC) return false;
      if (B.invMass != 0.0) return true;
      for (int k = 0; k < 9; k++) if (B.invInertiaWorld.e[k] != 0.f) return true;
      return false;
    };
    {
      size_t f = 0;
      for (size_t c = 0; c < contacts.size();) {
        size_t e = c;
        while (e < contacts.size() && contactManifold[e] == contactManifold[c] && e - c < (size_t)CANNON_COLORED_UNIT_CONTACTS) e++;
        CUnit u{contacts[c].bi, contacts[c].bj, (int)c, -1, 0u, {}, (int)c - (nWk > 1 ? wFirstC[worldOf(contacts[c].bi)] : 0)};
        for (size_t k = c; k < e; k++) {
          if (contacts[k].friction > 0) {
            if (accept(frictions[f])) { u.eqs.push_back(&frictions[f]); u.eqs.push_back(&frictions[f + 1]); }
            f += 2;
          }
          if (accept(contacts[k])) u.eqs.push_back(&contacts[k]);
        }
        if (!u.eqs.empty()) units.push_back(u);
        c = e;
      }
      int slot = 0;
      for (Constraint& c : constraints) {
        CUnit u{-1, -1, 0, -1, 0u, {}, 0};
        for (Eq& e : c.eqs)
          if (accept(e)) {
            if (u.eqs.empty()) {
              u.bi = e.bi; u.bj = e.bj; u.key = (int)contacts.size() + slot;
              const int wi = worldOf(e.bi);
              wFirstSlot[wi] = std::min(wFirstSlot[wi], slot);  // constraints are visited in slot order
              u.priKey = nWk > 1 ? wCountC[wi] + (slot - wFirstSlot[wi]) : u.key;
            }
            u.eqs.push_back(&e);
            slot++;
          }
        if (!u.eqs.empty()) units.push_back(u);
      }
    }
    int nLevels = 0;
    {
      std::vector<unsigned> claim(bodies.size(), 0xffffffffu);
      std::vector<int> active(units.size()), next;
      for (size_t k = 0; k < units.size(); k++) { active[k] = (int)k; units[k].pri = (unsigned)units[k].priKey * 2654435761u; }
      while (!active.empty()) {
        for (int k : active) {
          const CUnit& u = units[k];
          if (movable(u.bi) && u.pri < claim[u.bi]) claim[u.bi] = u.pri;
          if (movable(u.bj) && u.pri < claim[u.bj]) claim[u.bj] = u.pri;
        }
        next.clear();
        for (int k : active) {
          CUnit& u = units[k];
          const bool win = (!movable(u.bi) || claim[u.bi] == u.pri) && (!movable(u.bj) || claim[u.bj] == u.pri);
          if (win) u.level = nLevels; else next.push_back(k);
        }
        for (int k : active) { claim[units[k].bi] = 0xffffffffu; claim[units[k].bj] = 0xffffffffu; }
        active.swap(next);
        nLevels++;
      }
    }
    std::vector<int> ord(units.size());
    for (size_t k = 0; k < units.size(); k++) ord[k] = (int)k;
    std::sort(ord.begin(), ord.end(), [&](int a, int b) {
      return units[a].level != units[b].level ? units[a].level < units[b].level : units[a].key < units[b].key;
    });
    prof.n_levels = nLevels;
    std::vector<std::vector<Eq*>> per(nW);
    for (int k : ord) {
      const int wi = nW > 1 ? bodies[units[k].bi].worldId : 0;
      for (Eq* e : units[k].eqs) per[wi].push_back(e);
    }
    std::vector<std::vector<int>> wbodies(nW);
    for (int i = 0; i < (int)bodies.size(); i++) wbodies[nW > 1 ? bodies[i].worldId : 0].push_back(i);
    std::unordered_map<Eq*, RowDebug> tagged;
    for (int wi = 0; wi < nW; wi++) {
      if (wbodies[wi].empty()) continue;
      std::vector<RowDebug> dbg;
      int it = gsSolve(*this, per[wi], wbodies[wi], h, dbg);
      if (it > itersMax) itersMax = it;
      for (size_t k = 0; k < dbg.size(); k++) tagged[per[wi][k]] = dbg[k];
    }
    for (int k : ord)
      for (Eq* e : units[k].eqs) {
        RowDebug r = tagged[e];
        r.level = units[k].level;
        rows.push_back(r);
      }
  } else if (nW == 1) {
    std::vector<Eq*> eqs;
    for (Eq& e : frictions) if (accept(e)) eqs.push_back(&e);
    for (Eq& e : contacts) if (accept(e)) eqs.push_back(&e);
    for (Constraint& c : constraints)
      for (Eq& e : c.eqs) if (accept(e)) eqs.push_back(&e);
    std::vector<int> all(bodies.size());
    for (size_t i = 0; i < bodies.size(); i++) all[i] = (int)i;
    itersMax = gsSolve(*this, eqs, all, h, rows);
  } else {
    // a batch is nW separate World objects in the reference: each solves its own equation list
    // (own iteration loop and tolerance early-exit) over its own bodies
    std::vector<std::vector<Eq*>> per(nW);
    std::vector<std::vector<int>> perGlobal(nW);  // position of each equation in the batch-wide list (debug row order)
    int gidx = 0;
    auto put = [&](Eq& e) { per[bodies[e.bi].worldId].push_back(&e); perGlobal[bodies[e.bi].worldId].push_back(gidx++); };
    for (Eq& e : frictions) if (accept(e)) put(e);
    for (Eq& e : contacts) if (accept(e)) put(e);
    for (Constraint& c : constraints)
      for (Eq& e : c.eqs) if (accept(e)) put(e);
    std::vector<int> wb0(nW, -1), wb1(nW, 0);
    for (int i = 0; i < (int)bodies.size(); i++) {
      int wi = bodies[i].worldId;
      if (wb0[wi] < 0) wb0[wi] = i;
      wb1[wi] = i + 1;
    }
    for (int wi = 0; wi < nW; wi++) {
      if (wb0[wi] < 0) continue;
      std::vector<RowDebug> dbg;
      std::vector<int> idx;
      for (int i = wb0[wi]; i < wb1[wi]; i++) idx.push_back(i);
      int it = gsSolve(*this, per[wi], idx, h, dbg);
      if (it > itersMax) itersMax = it;
      if ((int)rows.size() < gidx) rows.resize(gidx);
      for (size_t k = 0; k < dbg.size(); k++) rows[perGlobal[wi][k]] = dbg[k];
    }
  }
  return itersMax;
}

void World::integrateAll(double h) {
  // damping, world_class.dart:648-659
  for (Body& bi : bodies) {
    if (bi.type == CANNON_BODY_DYNAMIC) {
      double ld = std::pow(1.0 - bi.linearDamping, h);
      bi.velocity = scale(ld, bi.velocity);
      double ad = std::pow(1.0 - bi.angularDamping, h);
      bi.angularVelocity = scale(ad, bi.angularVelocity);
    }
  }
  const bool quatNormalize = stepnumber % (desc.quat_normalize_skip + 1) == 0;
  for (Body& b : bodies) {
    // Body.integrate, rigid_body.dart:627-680
    if (!(b.type == CANNON_BODY_DYNAMIC || b.type == CANNON_BODY_KINEMATIC) || b.sleepState == CANNON_SLEEPING) continue;
    double iMdt = b.invMass * h;
    b.velocity.x = (float)(D(b.velocity.x) + D(b.force.x) * iMdt * D(b.linearFactor.x));
    b.velocity.y = (float)(D(b.velocity.y) + D(b.force.y) * iMdt * D(b.linearFactor.y));
    b.velocity.z = (float)(D(b.velocity.z) + D(b.force.z) * iMdt * D(b.linearFactor.z));
    const float* e = b.invInertiaWorld.e;
    double tx = D(b.torque.x) * D(b.angularFactor.x);
    double ty = D(b.torque.y) * D(b.angularFactor.y);
    double tz = D(b.torque.z) * D(b.angularFactor.z);
    b.angularVelocity.x = (float)(D(b.angularVelocity.x) + h * (D(e[0]) * tx + D(e[1]) * ty + D(e[2]) * tz));
    b.angularVelocity.y = (float)(D(b.angularVelocity.y) + h * (D(e[3]) * tx + D(e[4]) * ty + D(e[5]) * tz));
    b.angularVelocity.z = (float)(D(b.angularVelocity.z) + h * (D(e[6]) * tx + D(e[7]) * ty + D(e[8]) * tz));
    b.position.x = (float)(D(b.position.x) + D(b.velocity.x) * h);
    b.position.y = (float)(D(b.position.y) + D(b.velocity.y) * h);
    b.position.z = (float)(D(b.position.z) + D(b.velocity.z) * h);
    {  // Quat.integrate, quaternion.dart:93-111
      double ax = D(b.angularVelocity.x) * D(b.angularFactor.x), ay = D(b.angularVelocity.y) * D(b.angularFactor.y),
             az = D(b.angularVelocity.z) * D(b.angularFactor.z);
      Q4& q = b.quaternion;
      double bx = D(q.x), by = D(q.y), bz = D(q.z), bw = D(q.w);
      double halfDt = h * 0.5;
      q.x = (float)(D(q.x) + halfDt * (ax * bw + ay * bz - az * by));
      q.y = (float)(D(q.y) + halfDt * (ay * bw + az * bx - ax * bz));
      q.z = (float)(D(q.z) + halfDt * (az * bw + ax * by - ay * bx));
      q.w = (float)(D(q.w) + halfDt * (-ax * bx - ay * by - az * bz));
    }
    if (quatNormalize) {
      Q4& q = b.quaternion;
      if (desc.quat_normalize_fast) {  // quaternion.dart:171-185
        double f = (3.0 - (D(q.x) * D(q.x) + D(q.y) * D(q.y) + D(q.z) * D(q.z) + D(q.w) * D(q.w))) / 2.0;
        if (f == 0) {
          q = Q4{0, 0, 0, 0};
        } else {
          q.x = (float)(D(q.x) * f);
          q.y = (float)(D(q.y) * f);
          q.z = (float)(D(q.z) * f);
          q.w = (float)(D(q.w) * f);
        }
      } else {  // vector_math Quaternion.normalize()
        double l = std::sqrt((D(q.x) * D(q.x)) + (D(q.y) * D(q.y)) + (D(q.z) * D(q.z)) + (D(q.w) * D(q.w)));
        if (l != 0.0) {
          double d = 1.0 / l;
          q.x = (float)(D(q.x) * d);
          q.y = (float)(D(q.y) * d);
          q.z = (float)(D(q.z) * d);
          q.w = (float)(D(q.w) * d);
        }
      }
    }
    updateInertiaWorld(b, false);
  }
  // clearForces, world_class.dart:773-781
  for (Body& b : bodies) {
    b.force = V3{0, 0, 0};
    b.torque = V3{0, 0, 0};
  }
  stepnumber += 1;
  applySprings();  // the postStep event, world_class.dart:685
  // sleepTick, world_class.dart:688-700 + rigid_body.dart:282-300 (sees `time` before this step's increment)
  if (desc.allow_sleep) {
    for (Body& b : bodies) {
      if (!b.allowSleep) continue;
      double speedSquared = length2(b.velocity) + length2(b.angularVelocity);
      double speedLimitSquared = b.sleepSpeedLimit * b.sleepSpeedLimit;
      if (b.sleepState == CANNON_AWAKE && speedSquared < speedLimitSquared) {
        b.sleepState = CANNON_SLEEPY;
        b.timeLastSleepy = time;
      } else if (b.sleepState == CANNON_SLEEPY && speedSquared > speedLimitSquared) {
        b.sleepState = CANNON_AWAKE;
        b.wakeUpAfterNarrowphase = false;
      } else if (b.sleepState == CANNON_SLEEPY && time - b.timeLastSleepy > b.sleepTimeLimit) {
        b.sleepState = CANNON_SLEEPING;
        b.velocity = V3{0, 0, 0};
        b.angularVelocity = V3{0, 0, 0};
        b.wakeUpAfterNarrowphase = false;
      }
    }
  }
}

// Spring.applyForce, lib/objects/spring.dart:108-157, for every spring in order (every Vector3 store rounds to float)
void World::applySprings() {
  for (const Spring& sp : springs) {
    Body& A = bodies[sp.bodyA];
    Body& B = bodies[sp.bodyB];
    const double k = sp.stiffness, d = sp.damping, l = sp.restLength;
    const V3 worldAnchorA = add(qvmult(A.quaternion, sp.localAnchorA), A.position);  // pointToWorldFrame, rigid_body.dart:332-337
    const V3 worldAnchorB = add(qvmult(B.quaternion, sp.localAnchorB), B.position);
    const V3 ri = sub(worldAnchorA, A.position), rj = sub(worldAnchorB, B.position);
    const V3 r = sub(worldAnchorB, worldAnchorA);
    const double rlen = length(r);
    V3 rUnit = r;
    normalize(rUnit);
    V3 u = sub(B.velocity, A.velocity);
    V3 tmp = cross(B.angularVelocity, rj);
    u = add(u, tmp);
    tmp = cross(A.angularVelocity, ri);
    u = sub(u, tmp);
    const V3 f = scale(-k * (rlen - l) - d * dot(u, rUnit), rUnit);
    A.force = sub(A.force, f);
    B.force = add(B.force, f);
    const V3 rixf = cross(ri, f), rjxf = cross(rj, f);
    A.torque = sub(A.torque, rixf);
    B.torque = add(B.torque, rjxf);
  }
}

// math.pow(x, 3) correctly rounded: x*x exactly as hi + lo (fma), times x in double-double, rounded once. A libm pow with
// < 1 ulp error agrees except when the exact cube lies within its error of a rounding boundary.
static inline double pow3_cr(double x) {
  const double hi = x * x, lo = std::fma(x, x, -hi);
  const double p = hi * x, e = std::fma(hi, x, -p);
  return p + (e + lo * x);
}

// SPHSystem.update, sph_system.dart:62-163 (w :166-170, gradw :173-177, nablaw :180-184)
void World::sphUpdate() {
  for (const Sph& S : sphSystems) {
    const int N = (int)S.particles.size();
    const double h = S.smoothingRadius, r2 = h * h, cs = S.speedOfSound, eps = S.eps;
    const double h9 = std::pow(h, 9);
    std::vector<double> densities(N), pressures(N);
    std::vector<std::vector<int>> neighbors(N);
    for (int i = 0; i < N; i++) {
      const Body& p = bodies[S.particles[i]];
      std::vector<int>& nb = neighbors[i];
      for (int k = 0; k < N; k++) {  // getNeighbors :50-60
        const Body& q = bodies[S.particles[k]];
        const V3 dist = sub(q.position, p.position);
        if (S.particles[k] != S.particles[i] && length2(dist) < r2) nb.push_back(S.particles[k]);
      }
      nb.push_back(S.particles[i]);
      double sum = 0.0;
      for (int b : nb) {
        const V3 dist = sub(p.position, bodies[b].position);
        const double len = length(dist);
        const double weight = (315.0 / (64.0 * M_PI * h9)) * pow3_cr(h * h - len * len);
        sum += bodies[b].mass * weight;
      }
      densities[i] = sum;
      pressures[i] = cs * cs * (densities[i] - S.density);
    }
    for (int i = 0; i < N; i++) {
      Body& particle = bodies[S.particles[i]];
      V3 aPressure{0, 0, 0}, aVisc{0, 0, 0};
      const std::vector<int>& nb = neighbors[i];
      for (int j = 0; j < (int)nb.size(); j++) {
        const Body& neighbor = bodies[nb[j]];
        const V3 rVec = sub(particle.position, neighbor.position);
        const double r = length(rVec);
        // pressures[j] / densities[j]: indexed by the position in the neighbour list (:131-133), as written
        const double pij = -neighbor.mass * (pressures[i] / (densities[i] * densities[i] + eps) + pressures[j] / (densities[j] * densities[j] + eps));
        V3 gradW;
        {  // gradw: r is recomputed from rVec (:174)
          const double rr = length(rVec);
          const double q = h * h - rr * rr;
          gradW = scale(945.0 / (32.0 * M_PI * h9) * (q * q), rVec);
        }
        gradW = scale(pij, gradW);
        aPressure = add(aPressure, gradW);
        V3 u = sub(neighbor.velocity, particle.velocity);
        u = scale((1.0 / (0.0001 + densities[i] * densities[j])) * S.viscosity * neighbor.mass, u);
        const double nabla = (945.0 / (32.0 * M_PI * h9)) * (h * h - r * r) * (7 * r * r - 3 * h * h);
        u = scale(nabla, u);
        aVisc = add(aVisc, u);
      }
      aVisc = scale(particle.mass, aVisc);
      aPressure = scale(particle.mass, aPressure);
      particle.force = add(particle.force, aVisc);
      particle.force = add(particle.force, aPressure);
    }
  }
}

// World.internalStep + the `time += dt` of World.step, world_class.dart:392-399,433-701
void World::internalStep(double h) {
  dt = h;
  const double gx = D(desc.gravity[0]), gy = D(desc.gravity[1]), gz = D(desc.gravity[2]);
  for (Body& bi : bodies) {
    if (bi.type == CANNON_BODY_DYNAMIC) {
      double m = bi.mass;
      bi.force.x = (float)(D(bi.force.x) + m * gx);
      bi.force.y = (float)(D(bi.force.y) + m * gy);
      bi.force.z = (float)(D(bi.force.z) + m * gz);
    }
  }
  sphUpdate();  // world_class.dart:472-475
  collisionPairs();
  getContacts();
  makeContactConstraints();
  int iters = solve(h);
  prof.n_pairs = (int64_t)p1.size();
  prof.n_contacts = (int64_t)contacts.size();
  prof.n_rows = (int64_t)rows.size();
  prof.iterations_done = iters;
  prof.contact_iters_total += (int64_t)contacts.size() * iters;
  integrateAll(h);
  prof.steps += 1;
  time += h;
}

}  // namespace orc
