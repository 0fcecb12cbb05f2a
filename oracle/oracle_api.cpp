/*
 * oracle_api.cpp — TEST INFRASTRUCTURE ONLY (PARITY UNPINNED, see oracle_math.h).
 * Implements include/cannon_cuda.h on the CPU over the sequential restatement in this directory, so
 * parity tests can drive the product (libcannon_cuda.so) and the checker with identical calls.
 * Never linked or loaded by the product.
 */
#include <cmath>
#include <cstring>
#include <new>

#include "oracle_world.h"

using namespace orc;

struct cannon_ctx {
  std::string err;
};
struct cannon_world {
  cannon_ctx* ctx;
  World w;
};

static int32_t fail(cannon_ctx* ctx, int32_t code, const char* msg) {
  if (ctx) ctx->err = msg;
  return code;
}

extern "C" {

int32_t cannon_version(void) { return CANNON_ABI_VERSION; }
const char* cannon_backend(void) { return "oracle"; }

int32_t cannon_ctx_create(int32_t, cannon_ctx** out) {
  if (!out) return CANNON_E_INVALID;
  *out = new (std::nothrow) cannon_ctx();
  return *out ? CANNON_OK : CANNON_E_INVALID;
}
void cannon_ctx_destroy(cannon_ctx* ctx) { delete ctx; }
const char* cannon_last_error(const cannon_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

void cannon_world_desc_default(cannon_world_desc* d) {
  std::memset(d, 0, sizeof(*d));
  d->solver_kind = CANNON_SOLVER_REFERENCE_ORDER;
  d->solver_iterations = 10;   // solver.dart:17
  d->solver_tolerance = 1e-7;  // solver.dart:18
  d->broadphase_kind = CANNON_BP_NAIVE;
  d->grid_nx = d->grid_ny = d->grid_nz = 10;  // grid_broadphase.dart:37-41
  for (int k = 0; k < 3; k++) {
    d->grid_min[k] = 100;
    d->grid_max[k] = -100;
  }
  cannon_contact_material& cm = d->default_contact_material;  // world_class.dart:155-158
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
  std::memset(d, 0, sizeof(*d));
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
  cannon_world* w = new (std::nothrow) cannon_world();
  if (!w) return CANNON_E_INVALID;
  w->ctx = ctx;
  w->w.desc = *desc;
  if (w->w.desc.n_worlds < 1) w->w.desc.n_worlds = 1;
  *out = w;
  return CANNON_OK;
}
void cannon_world_destroy(cannon_world* w) { delete w; }

int32_t cannon_world_set_materials(cannon_world* cw, int32_t n, const double* friction, const double* restitution,
                                   int32_t ncm, const cannon_contact_material* cms) {
  if (!cw || n < 0 || ncm < 0) return CANNON_E_INVALID;
  World& w = cw->w;
  w.matFriction.assign(n, -1.0);
  w.matRestitution.assign(n, -1.0);
  for (int i = 0; i < n; i++) {
    if (friction) w.matFriction[i] = friction[i];
    if (restitution) w.matRestitution[i] = restitution[i];
  }
  w.cms.assign(cms, cms + ncm);
  w.cmTable.assign((size_t)n * n, -1);
  for (int k = 0; k < ncm; k++) {
    int a = cms[k].material_a, b = cms[k].material_b;
    if (a < 0 || b < 0 || a >= n || b >= n) return fail(cw->ctx, CANNON_E_INVALID, "contact material references unknown material");
    w.cmTable[(size_t)a * n + b] = k;  // TupleDictionary key is unordered (tuple_dictionary.dart:1-3)
    w.cmTable[(size_t)b * n + a] = k;
  }
  return CANNON_OK;
}

int32_t cannon_world_set_shapes(cannon_world* cw, int32_t n, const cannon_shape_desc* sd) {
  if (!cw || n < 0 || (n > 0 && !sd)) return CANNON_E_INVALID;
  World& w = cw->w;
  w.shapes.clear();
  w.shapes.resize(n);
  for (int i = 0; i < n; i++) {
    Shape& s = w.shapes[i];
    const cannon_shape_desc& d = sd[i];
    s.type = d.type;
    s.collisionResponse = d.collision_response != 0;
    s.group = d.collision_filter_group;
    s.mask = d.collision_filter_mask;
    s.material = d.material;
    if (s.material >= (int)w.matFriction.size()) return fail(cw->ctx, CANNON_E_INVALID, "shape references unknown material");
    switch (d.type) {
      case CANNON_SHAPE_SPHERE:
        if (d.radius < 0) return fail(cw->ctx, CANNON_E_INVALID, "The sphere radius cannot be negative.");
        s.radius = d.radius;
        s.boundingSphereRadius = d.radius;  // sphere.dart:39-41
        break;
      case CANNON_SHAPE_PLANE:
        s.boundingSphereRadius = INFINITY;  // plane.dart:20
        break;
      case CANNON_SHAPE_PARTICLE:
        s.boundingSphereRadius = 0;  // particle.dart:24-26
        break;
      case CANNON_SHAPE_TRIMESH: {  // Trimesh constructor, trimesh.dart:60-73
        if (d.n_vertices <= 0 || d.n_triangles <= 0 || !d.vertices || !d.tm_indices) return fail(cw->ctx, CANNON_E_INVALID, "trimesh needs vertices and indices");
        for (int k = 0; k < 3 * d.n_triangles; k++)
          if (d.tm_indices[k] < 0 || d.tm_indices[k] >= d.n_vertices) return fail(cw->ctx, CANNON_E_INVALID, "trimesh index out of range");
        s.tmIdx.assign(d.tm_indices, d.tm_indices + 3 * d.n_triangles);
        std::vector<V3> raw(d.n_vertices);
        s.tmVerts.resize(d.n_vertices);
        for (int v = 0; v < d.n_vertices; v++) {  // getVertex :260-268: setValues, then *= scale component by component
          raw[v] = V3{d.vertices[3 * v], d.vertices[3 * v + 1], d.vertices[3 * v + 2]};
          s.tmVerts[v] = v3(D(raw[v].x) * D(d.tm_scale[0]), D(raw[v].y) * D(d.tm_scale[1]), D(raw[v].z) * D(d.tm_scale[2]));
        }
        s.tmNormals.resize(d.n_triangles);
        for (int t = 0; t < d.n_triangles; t++) {  // updateNormals :157-175 with computeNormal(vb, va, vc) :220-230, unit scale
          const V3 &va = raw[s.tmIdx[3 * t]], &vb = raw[s.tmIdx[3 * t + 1]], &vc = raw[s.tmIdx[3 * t + 2]];
          const V3 ab = sub(va, vb), cb = sub(vc, va);
          V3 n = cross(cb, ab);
          if (!(n.x == 0 && n.y == 0 && n.z == 0)) normalize(n);
          s.tmNormals[t] = n;
        }
        // computeLocalAABB :315-343 (note the else-if) and updateBoundingSphereRadius :350-363
        V3 l = s.tmVerts[0], u = s.tmVerts[0];
        double max2 = 0;
        for (const V3& v : s.tmVerts) {
          if (v.x < l.x) l.x = v.x; else if (v.x > u.x) u.x = v.x;
          if (v.y < l.y) l.y = v.y; else if (v.y > u.y) u.y = v.y;
          if (v.z < l.z) l.z = v.z; else if (v.z > u.z) u.z = v.z;
          const double n2 = length2(v);
          if (n2 > max2) max2 = n2;
        }
        s.tmLo = l; s.tmHi = u;
        s.boundingSphereRadius = std::sqrt(max2);
        break;
      }
      case CANNON_SHAPE_BOX:
        s.halfExtents = V3{d.half_extents[0], d.half_extents[1], d.half_extents[2]};
        make_box_hull(s.halfExtents, s.hull);
        s.boundingSphereRadius = length(s.halfExtents);  // box.dart:124-126
        break;
      case CANNON_SHAPE_CYLINDER:
        if (d.radius_top < 0 || d.radius_bottom < 0) return fail(cw->ctx, CANNON_E_INVALID, "The cylinder radius cannot be negative.");
        make_cylinder_hull(d.radius_top, d.radius_bottom, d.height, d.num_segments, s.hull);
        s.boundingSphereRadius = s.hull.boundingSphereRadius;
        break;
      case CANNON_SHAPE_CONVEX:
      case CANNON_SHAPE_CAPSULE:      // ConvexPolyhedron subclasses (capsule.dart, capsule_lathe.dart, cone.dart, sized_plane.dart):
      case CANNON_SHAPE_CONE:         // the hull arrives as the reference's constructor built it
      case CANNON_SHAPE_SIZED_PLANE: {
        if (d.n_vertices <= 0 || d.n_faces <= 0 || !d.vertices || !d.face_offsets || !d.face_indices)
          return fail(cw->ctx, CANNON_E_INVALID, "convex shape needs vertices and faces");
        s.hull.vertices.resize(d.n_vertices);
        for (int v = 0; v < d.n_vertices; v++) s.hull.vertices[v] = V3{d.vertices[3 * v], d.vertices[3 * v + 1], d.vertices[3 * v + 2]};
        s.hull.faces.resize(d.n_faces);
        for (int f = 0; f < d.n_faces; f++)
          s.hull.faces[f].assign(d.face_indices + d.face_offsets[f], d.face_indices + d.face_offsets[f + 1]);
        s.hull.hasUniqueAxes = d.convex_has_axes != 0;  // no `axes` => no face-normal axes (§5.9-9); Cone passes them (cone.dart:58-61)
        s.hull.computeNormals();
        s.hull.updateBoundingSphereRadius();
        s.hull.computeEdges();
        s.boundingSphereRadius = s.hull.boundingSphereRadius;
        break;
      }
      case CANNON_SHAPE_HEIGHTFIELD: {
        if (d.hf_nx < 2 || d.hf_ny < 2 || !d.hf_data) return fail(cw->ctx, CANNON_E_INVALID, "heightfield needs >= 2x2 samples");
        s.nx = d.hf_nx;
        s.ny = d.hf_ny;
        s.elementSize = d.hf_element_size;
        s.data.assign(d.hf_data, d.hf_data + (size_t)d.hf_nx * d.hf_ny);
        // updateMinValue / updateMaxValue, heightfield.dart:87-113
        double mn = s.data[0], mx = s.data[0];
        for (double v : s.data) {
          if (v < mn) mn = v;
          if (v > mx) mx = v;
        }
        s.minValue = mn;
        s.maxValue = mx;
        // updateBoundingSphereRadius, heightfield.dart:505-515 (components rounded to float first)
        double es = (double)s.elementSize;
        V3 t = v3(s.nx * es, s.ny * es, std::fmax(std::fabs(mx), std::fabs(mn)));
        s.boundingSphereRadius = length(t);
        break;
      }
      default:
        return fail(cw->ctx, CANNON_E_UNSUPPORTED, "shape type outside the hot-path scope");
    }
  }
  return CANNON_OK;
}

#define GETF3(dst, src, i) dst = V3{src[3 * (i)], src[3 * (i) + 1], src[3 * (i) + 2]}

int32_t cannon_world_set_body_shapes(cannon_world* cw, int32_t n_bodies, const int32_t* first, const int32_t* shape, const float* offset,
                                     const float* orientation) {
  if (!cw || n_bodies < 0) return CANNON_E_INVALID;
  World& w = cw->w;
  w.pendFirst.clear(); w.pendShape.clear(); w.pendOffset.clear(); w.pendOrient.clear();
  if (n_bodies == 0) return CANNON_OK;
  if (!first || first[0] != 0) return fail(cw->ctx, CANNON_E_INVALID, "cannon_world_set_body_shapes: first[0] must be 0");
  for (int b = 0; b < n_bodies; b++) if (first[b + 1] < first[b]) return fail(cw->ctx, CANNON_E_INVALID, "cannon_world_set_body_shapes: first[] must ascend");
  const int ni = first[n_bodies];
  if (ni > 0 && !shape) return fail(cw->ctx, CANNON_E_INVALID, "cannon_world_set_body_shapes: shape[] missing");
  for (int k = 0; k < ni; k++) {
    if (shape[k] < 0 || shape[k] >= (int)w.shapes.size()) return fail(cw->ctx, CANNON_E_INVALID, "body references unknown shape");
    w.pendShape.push_back(shape[k]);
    w.pendOffset.push_back(offset ? V3{offset[3 * k], offset[3 * k + 1], offset[3 * k + 2]} : V3{0, 0, 0});
    w.pendOrient.push_back(orientation ? Q4{orientation[4 * k], orientation[4 * k + 1], orientation[4 * k + 2], orientation[4 * k + 3]} : Q4{0, 0, 0, 1});
  }
  w.pendFirst.assign(first, first + n_bodies + 1);
  return CANNON_OK;
}

int32_t cannon_world_set_bodies(cannon_world* cw, const cannon_bodies_soa* s) {
  if (!cw || !s || s->n < 0) return CANNON_E_INVALID;
  World& w = cw->w;
  const int n = s->n;
  if (!w.pendFirst.empty() && (int)w.pendFirst.size() != n + 1) return fail(cw->ctx, CANNON_E_INVALID, "cannon_world_set_body_shapes described another body count");
  for (int sh : w.pendShape) if (sh >= (int)w.shapes.size()) return fail(cw->ctx, CANNON_E_INVALID, "the body shape table references a shape the current shape table does not have");
  w.bodies.clear();
  w.bodies.resize(n);
  w.sapAxisList.clear();
  for (int i = 0; i < n; i++) {
    Body& b = w.bodies[i];
    if (s->position) GETF3(b.position, s->position, i);
    if (s->quaternion) b.quaternion = Q4{s->quaternion[4 * i], s->quaternion[4 * i + 1], s->quaternion[4 * i + 2], s->quaternion[4 * i + 3]};
    if (s->velocity) GETF3(b.velocity, s->velocity, i);
    if (s->angular_velocity) GETF3(b.angularVelocity, s->angular_velocity, i);
    if (s->force) GETF3(b.force, s->force, i);
    if (s->torque) GETF3(b.torque, s->torque, i);
    b.mass = s->mass ? s->mass[i] : 0.0;
    b.type = b.mass <= 0.0 ? CANNON_BODY_STATIC : CANNON_BODY_DYNAMIC;  // rigid_body.dart:61
    if (s->type && s->type[i] >= 0) b.type = s->type[i];
    if (s->sleep_state) b.sleepState = s->sleep_state[i];
    b.timeLastSleepy = s->time_last_sleepy ? s->time_last_sleepy[i] : w.time;  // world_class.dart:291
    if (s->allow_sleep) b.allowSleep = s->allow_sleep[i] != 0;
    if (s->sleep_speed_limit) b.sleepSpeedLimit = s->sleep_speed_limit[i];
    if (s->sleep_time_limit) b.sleepTimeLimit = s->sleep_time_limit[i];
    if (s->linear_damping) b.linearDamping = s->linear_damping[i];
    if (s->angular_damping) b.angularDamping = s->angular_damping[i];
    if (s->linear_factor) GETF3(b.linearFactor, s->linear_factor, i);
    if (s->angular_factor) GETF3(b.angularFactor, s->angular_factor, i);
    if (s->fixed_rotation) b.fixedRotation = s->fixed_rotation[i] != 0;
    if (s->collision_filter_group) b.group = s->collision_filter_group[i];
    if (s->collision_filter_mask) b.mask = s->collision_filter_mask[i];
    if (s->collision_response) b.collisionResponse = s->collision_response[i] != 0;
    if (s->is_trigger) b.isTrigger = s->is_trigger[i] != 0;
    b.material = s->material ? s->material[i] : -1;
    b.shape = s->shape ? s->shape[i] : -1;
    b.shapes.clear(); b.shapeOffsets.clear(); b.shapeOrientations.clear();
    if (!w.pendFirst.empty()) {  // Body.addShape for every instance of the table (rigid_body.dart:348-377)
      for (int k = w.pendFirst[i]; k < w.pendFirst[i + 1]; k++) {
        b.shapes.push_back(w.pendShape[k]);
        b.shapeOffsets.push_back(w.pendOffset[k]);
        b.shapeOrientations.push_back(w.pendOrient[k]);
      }
      b.shape = b.shapes.empty() ? -1 : b.shapes[0];
    } else if (b.shape >= 0) {
      b.shapes.push_back(b.shape);
      b.shapeOffsets.push_back(V3{0, 0, 0});
      b.shapeOrientations.push_back(Q4{0, 0, 0, 1});
    }
    b.worldId = s->world_id ? s->world_id[i] : 0;
    if (b.shape >= (int)w.shapes.size()) return fail(cw->ctx, CANNON_E_INVALID, "body references unknown shape");
    if (b.material >= (int)w.matFriction.size()) return fail(cw->ctx, CANNON_E_INVALID, "body references unknown material");
    if (b.worldId < 0 || b.worldId >= w.desc.n_worlds) return fail(cw->ctx, CANNON_E_INVALID, "world_id out of range");
    w.updateMassProperties(b);
    w.updateBoundingRadius(b);
  }
  return CANNON_OK;
}

#define PUTF3(dst, i, v) do { dst[3 * (i)] = (v).x; dst[3 * (i) + 1] = (v).y; dst[3 * (i) + 2] = (v).z; } while (0)

int32_t cannon_world_get_bodies(cannon_world* cw, cannon_bodies_soa* o) {
  if (!cw || !o) return CANNON_E_INVALID;
  World& w = cw->w;
  const int n = (int)w.bodies.size();
  if (o->n < n) {
    o->n = n;
    return fail(cw->ctx, CANNON_E_CAPACITY, "cannon_bodies_soa.n too small");
  }
  o->n = n;
  for (int i = 0; i < n; i++) {
    Body& b = w.bodies[i];
    if (o->position) PUTF3(o->position, i, b.position);
    if (o->quaternion) { o->quaternion[4 * i] = b.quaternion.x; o->quaternion[4 * i + 1] = b.quaternion.y; o->quaternion[4 * i + 2] = b.quaternion.z; o->quaternion[4 * i + 3] = b.quaternion.w; }
    if (o->velocity) PUTF3(o->velocity, i, b.velocity);
    if (o->angular_velocity) PUTF3(o->angular_velocity, i, b.angularVelocity);
    if (o->force) PUTF3(o->force, i, b.force);
    if (o->torque) PUTF3(o->torque, i, b.torque);
    if (o->mass) o->mass[i] = b.mass;
    if (o->type) o->type[i] = b.type;
    if (o->sleep_state) o->sleep_state[i] = b.sleepState;
    if (o->time_last_sleepy) o->time_last_sleepy[i] = b.timeLastSleepy;
    if (o->allow_sleep) o->allow_sleep[i] = b.allowSleep;
    if (o->sleep_speed_limit) o->sleep_speed_limit[i] = b.sleepSpeedLimit;
    if (o->sleep_time_limit) o->sleep_time_limit[i] = b.sleepTimeLimit;
    if (o->linear_damping) o->linear_damping[i] = b.linearDamping;
    if (o->angular_damping) o->angular_damping[i] = b.angularDamping;
    if (o->linear_factor) PUTF3(o->linear_factor, i, b.linearFactor);
    if (o->angular_factor) PUTF3(o->angular_factor, i, b.angularFactor);
    if (o->fixed_rotation) o->fixed_rotation[i] = b.fixedRotation;
    if (o->collision_filter_group) o->collision_filter_group[i] = b.group;
    if (o->collision_filter_mask) o->collision_filter_mask[i] = b.mask;
    if (o->collision_response) o->collision_response[i] = b.collisionResponse;
    if (o->is_trigger) o->is_trigger[i] = b.isTrigger;
    if (o->material) o->material[i] = b.material;
    if (o->shape) o->shape[i] = b.shape;
    if (o->world_id) o->world_id[i] = b.worldId;
    if (o->inv_mass) o->inv_mass[i] = b.invMass;
    if (o->inv_inertia) PUTF3(o->inv_inertia, i, b.invInertia);
    if (o->inv_inertia_world) std::memcpy(o->inv_inertia_world + 9 * i, b.invInertiaWorld.e, 9 * sizeof(float));
    if (o->bounding_radius) o->bounding_radius[i] = b.boundingRadius;
    if (o->aabb) {
      w.updateAABB(b);
      o->aabb[6 * i + 0] = b.aabbLower.x; o->aabb[6 * i + 1] = b.aabbLower.y; o->aabb[6 * i + 2] = b.aabbLower.z;
      o->aabb[6 * i + 3] = b.aabbUpper.x; o->aabb[6 * i + 4] = b.aabbUpper.y; o->aabb[6 * i + 5] = b.aabbUpper.z;
    }
  }
  return CANNON_OK;
}

int32_t cannon_world_update_bodies(cannon_world* cw, int32_t first, int32_t count, const float* position, const float* quaternion,
                                   const float* velocity, const float* angular_velocity, const float* force, const float* torque) {
  if (!cw || first < 0 || count < 0 || first + count > (int)cw->w.bodies.size()) return CANNON_E_INVALID;
  World& w = cw->w;
  for (int k = 0; k < count; k++) {
    Body& b = w.bodies[first + k];
    if (position) GETF3(b.position, position, k);
    if (quaternion) b.quaternion = Q4{quaternion[4 * k], quaternion[4 * k + 1], quaternion[4 * k + 2], quaternion[4 * k + 3]};
    if (velocity) GETF3(b.velocity, velocity, k);
    if (angular_velocity) GETF3(b.angularVelocity, angular_velocity, k);
    if (force) GETF3(b.force, force, k);
    if (torque) GETF3(b.torque, torque, k);
    if (quaternion) w.updateInertiaWorld(b, false);
  }
  return CANNON_OK;
}

int32_t cannon_world_set_inv_inertia(cannon_world* cw, int32_t first, int32_t count, const float* inv_inertia) {
  if (!cw || first < 0 || count < 0 || first + count > (int)cw->w.bodies.size() || (count > 0 && !inv_inertia)) return CANNON_E_INVALID;
  for (int k = 0; k < count; k++) {
    Body& b = cw->w.bodies[first + k];
    GETF3(b.invInertia, inv_inertia, k);
    cw->w.updateInertiaWorld(b, true);  // rigid_body.dart:450-466
  }
  return CANNON_OK;
}

int32_t cannon_world_update_sleep_states(cannon_world* cw, int32_t first, int32_t count, const int32_t* sleep_state) {
  if (!cw || first < 0 || count < 0 || first + count > (int)cw->w.bodies.size() || (count > 0 && !sleep_state)) return CANNON_E_INVALID;
  for (int k = 0; k < count; k++)
    if (sleep_state[k] < CANNON_AWAKE || sleep_state[k] > CANNON_SLEEPING) return fail(cw->ctx, CANNON_E_INVALID, "sleep state out of range");
  for (int k = 0; k < count; k++) cw->w.bodies[first + k].sleepState = sleep_state[k];  // rigid_body.dart:263-278
  return CANNON_OK;
}

int32_t cannon_world_set_hinge_motor(cannon_world* cw, int32_t constraint, int32_t enabled, double target_velocity, double max_force) {
  if (!cw || constraint < 0 || constraint >= (int)cw->w.constraints.size()) return CANNON_E_INVALID;
  Constraint& c = cw->w.constraints[constraint];
  if (c.type != CANNON_CONSTRAINT_HINGE) return fail(cw->ctx, CANNON_E_INVALID, "constraint is not a HingeConstraint");
  Eq& m = c.eqs[5];  // hinge_constraint.dart:50,56-76
  m.enabled = enabled != 0;
  m.targetVelocity = target_velocity;
  m.maxForce = max_force;
  m.minForce = -max_force;
  return CANNON_OK;
}

void cannon_sph_desc_default(cannon_sph_desc* d) {
  if (!d) return;
  memset(d, 0, sizeof(*d));
  d->density = 1; d->smoothing_radius = 1; d->speed_of_sound = 1; d->viscosity = 0.01; d->eps = 0.00001;  // sph_system.dart:10-17
}

int32_t cannon_world_set_sph_systems(cannon_world* cw, int32_t n, const cannon_sph_desc* sd) {
  if (!cw || n < 0 || (n > 0 && !sd)) return CANNON_E_INVALID;
  World& w = cw->w;
  w.sphSystems.clear();
  for (int k = 0; k < n; k++) {
    World::Sph S;
    if (sd[k].n_particles < 0 || (sd[k].n_particles > 0 && !sd[k].particles)) return fail(cw->ctx, CANNON_E_INVALID, "SPH system needs its particle list");
    for (int i = 0; i < sd[k].n_particles; i++) {
      if (sd[k].particles[i] < 0 || sd[k].particles[i] >= (int)w.bodies.size()) return fail(cw->ctx, CANNON_E_INVALID, "SPH particle references unknown body");
      S.particles.push_back(sd[k].particles[i]);
    }
    S.density = sd[k].density; S.smoothingRadius = sd[k].smoothing_radius; S.speedOfSound = sd[k].speed_of_sound; S.viscosity = sd[k].viscosity; S.eps = sd[k].eps;
    w.sphSystems.push_back(S);
  }
  return CANNON_OK;
}

int32_t cannon_world_set_constraints(cannon_world* cw, int32_t n, const cannon_constraint_desc* cs) {
  if (!cw || n < 0 || (n > 0 && !cs)) return CANNON_E_INVALID;
  World& w = cw->w;
  w.constraints.clear();
  const int nb = (int)w.bodies.size();
  for (int i = 0; i < n; i++) {
    const cannon_constraint_desc& d = cs[i];
    if (d.body_a < 0 || d.body_b < 0 || d.body_a >= nb || d.body_b >= nb) return fail(cw->ctx, CANNON_E_INVALID, "constraint references unknown body");
    Constraint c;
    c.type = d.type;
    c.bodyA = d.body_a;
    c.bodyB = d.body_b;
    c.pivotA = V3{d.pivot_a[0], d.pivot_a[1], d.pivot_a[2]};
    c.pivotB = V3{d.pivot_b[0], d.pivot_b[1], d.pivot_b[2]};
    c.collideConnected = d.collide_connected != 0;
    // Constraint ctor wakes both bodies up (constraint_class.dart:26-29)
    w.bodies[c.bodyA].sleepState = CANNON_AWAKE;
    w.bodies[c.bodyB].sleepState = CANNON_AWAKE;
    auto rotEq = [&](double maxAngle, double minF, double maxF) {
      Eq e;
      e.kind = EQ_ROTATIONAL;
      e.bi = c.bodyA;
      e.bj = c.bodyB;
      e.setSpookParams(1e7, 4, 1.0 / 60);
      e.minForce = minF;
      e.maxForce = maxF;
      e.maxAngle = maxAngle;
      return e;
    };
    Body Actor = w.bodies[c.bodyA], Bctor = w.bodies[c.bodyB];  // what the constraint's constructor saw
    if (d.has_ctor_pose) {
      Actor.position = V3{d.ctor_pos_a[0], d.ctor_pos_a[1], d.ctor_pos_a[2]}; Actor.quaternion = Q4{d.ctor_quat_a[0], d.ctor_quat_a[1], d.ctor_quat_a[2], d.ctor_quat_a[3]};
      Bctor.position = V3{d.ctor_pos_b[0], d.ctor_pos_b[1], d.ctor_pos_b[2]}; Bctor.quaternion = Q4{d.ctor_quat_b[0], d.ctor_quat_b[1], d.ctor_quat_b[2], d.ctor_quat_b[3]};
    }
    const Body &A = Actor, &B = Bctor;
    if (d.type == CANNON_CONSTRAINT_DISTANCE) {  // distance_constraint.dart:14-23: one bidirectional ContactEquation
      c.distance = d.distance >= 0 ? d.distance : distance_to(A.position, B.position);
      Eq e;
      e.kind = EQ_CONTACT;
      e.bi = c.bodyA;
      e.bj = c.bodyB;
      e.setSpookParams(1e7, 4, 1.0 / 60);
      e.minForce = -d.max_force;
      e.maxForce = d.max_force;
      c.eqs.push_back(e);
      w.constraints.push_back(c);
      continue;
    }
    // PointToPointConstraint ctor, point_to_point_constraint.dart:30-66: three bidirectional
    // ContactEquations with the Equation-ctor SPOOK parameters (1e7, 4, 1/60; equation_class.dart:38)
    for (int k = 0; k < 3; k++) {
      Eq e;
      e.kind = EQ_CONTACT;
      e.bi = c.bodyA;
      e.bj = c.bodyB;
      e.setSpookParams(1e7, 4, 1.0 / 60);
      e.minForce = -d.max_force;
      e.maxForce = d.max_force;
      e.ni = V3{k == 0 ? 1.f : 0.f, k == 1 ? 1.f : 0.f, k == 2 ? 1.f : 0.f};
      c.eqs.push_back(e);
    }
    if (d.type == CANNON_CONSTRAINT_HINGE) {  // hinge_constraint.dart:20-51
      c.axisA = V3{d.axis_a[0], d.axis_a[1], d.axis_a[2]};
      normalize(c.axisA);
      c.axisB = V3{d.axis_b[0], d.axis_b[1], d.axis_b[2]};
      normalize(c.axisB);
      for (int k = 0; k < 2; k++) {
        Eq e = rotEq(M_PI / 2, -d.max_force, d.max_force);
        e.axisA = c.axisA;
        e.axisB = c.axisB;
        c.eqs.push_back(e);
      }
      Eq m;
      m.kind = EQ_MOTOR;
      m.bi = c.bodyA;
      m.bj = c.bodyB;
      m.setSpookParams(1e7, 4, 1.0 / 60);
      double mf = d.motor_max_force > 0 ? d.motor_max_force : d.max_force;
      m.minForce = -mf;
      m.maxForce = mf;
      m.enabled = d.motor_enabled != 0;
      m.targetVelocity = d.motor_target_velocity;
      m.axisA = V3{0, 0, 0};
      m.axisB = V3{0, 0, 0};
      c.eqs.push_back(m);
    } else if (d.type == CANNON_CONSTRAINT_LOCK) {  // lock_constraint.dart:22-66
      // pivots: the halfway point in both local frames. Body.pointToLocalFrame / vectorToLocalFrame
      // (rigid_body.dart:317-329) conjugate the body's quaternion IN PLACE (vector_math Quaternion.conjugate mutates),
      // so successive calls alternate between q* and q: pivot (q*, correct), x (q), y (q*), z (q); after the four calls
      // of the constructor the body's quaternion is back where it was.
      V3 halfWay = add(A.position, B.position);
      halfWay = scale(0.5, halfWay);
      c.pivotB = qvmult(qconj(B.quaternion), sub(halfWay, B.position));
      c.pivotA = qvmult(qconj(A.quaternion), sub(halfWay, A.position));
      const V3 X{1, 0, 0}, Y{0, 1, 0}, Z{0, 0, 1};
      const V3 xA = qvmult(A.quaternion, X), xB = qvmult(B.quaternion, X);
      const V3 yA = qvmult(qconj(A.quaternion), Y), yB = qvmult(qconj(B.quaternion), Y);
      const V3 zA = qvmult(A.quaternion, Z), zB = qvmult(B.quaternion, Z);
      // update(): r1 (xA, yB), r2 (yA, zB), r3 (zA, xB), lock_constraint.dart:79-87
      const V3 la[3] = {xA, yA, zA}, lb[3] = {yB, zB, xB};
      c.locA.assign(c.eqs.size(), V3{0, 0, 0});
      c.locB.assign(c.eqs.size(), V3{0, 0, 0});
      for (int k = 0; k < 3; k++) {
        c.eqs.push_back(rotEq(M_PI / 2, -d.max_force, d.max_force));
        c.locA.push_back(la[k]);
        c.locB.push_back(lb[k]);
      }
    } else if (d.type == CANNON_CONSTRAINT_CONE_TWIST) {  // cone_twist_constraint.dart:26-74
      c.axisA = V3{d.axis_a[0], d.axis_a[1], d.axis_a[2]};
      c.axisB = V3{d.axis_b[0], d.axis_b[1], d.axis_b[2]};
      c.locA.assign(c.eqs.size(), V3{0, 0, 0});
      c.locB.assign(c.eqs.size(), V3{0, 0, 0});
      // ConeEquation(maxForce: 0) then minForce = -maxForce: pushes toward the cone axis only; same for the twist
      c.eqs.push_back(rotEq(d.angle, -d.max_force, 0.0));  // ConeEquation.computeB == RotationalEquation.computeB with cos(angle)
      c.locA.push_back(c.axisA);
      c.locB.push_back(c.axisB);
      // update(): axisA.tangents(twist.axisA, twist.axisA) leaves the SECOND tangent in twist.axisA (both outputs are
      // the same object, cross2 reads its argument before writing), then vectorToWorldFrame (:88-94)
      V3 t1, t2a, t2b;
      tangents(c.axisA, t1, t2a);
      tangents(c.axisB, t1, t2b);
      c.eqs.push_back(rotEq(d.twist_angle, -d.max_force, 0.0));
      c.locA.push_back(t2a);
      c.locB.push_back(t2b);
    } else if (d.type != CANNON_CONSTRAINT_POINT_TO_POINT) {
      return fail(cw->ctx, CANNON_E_UNSUPPORTED, "constraint type outside the hot-path scope");
    }
    w.constraints.push_back(c);
  }
  return CANNON_OK;
}

int32_t cannon_world_set_springs(cannon_world* cw, int32_t n, const cannon_spring_desc* sp) {
  if (!cw || n < 0 || (n > 0 && !sp)) return CANNON_E_INVALID;
  World& w = cw->w;
  const int nb = (int)w.bodies.size();
  w.springs.clear();
  for (int i = 0; i < n; i++) {
    if (sp[i].body_a < 0 || sp[i].body_b < 0 || sp[i].body_a >= nb || sp[i].body_b >= nb) return fail(cw->ctx, CANNON_E_INVALID, "spring references unknown body");
    Spring s;
    s.bodyA = sp[i].body_a;
    s.bodyB = sp[i].body_b;
    s.restLength = sp[i].rest_length;
    s.stiffness = sp[i].stiffness;
    s.damping = sp[i].damping;
    s.localAnchorA = V3{sp[i].local_anchor_a[0], sp[i].local_anchor_a[1], sp[i].local_anchor_a[2]};
    s.localAnchorB = V3{sp[i].local_anchor_b[0], sp[i].local_anchor_b[1], sp[i].local_anchor_b[2]};
    w.springs.push_back(s);
  }
  return CANNON_OK;
}

int32_t cannon_world_set_time(cannon_world* cw, double t) {
  if (!cw) return CANNON_E_INVALID;
  cw->w.time = t;
  return CANNON_OK;
}
int32_t cannon_world_get_time(cannon_world* cw, double* t, int64_t* stepnumber) {
  if (!cw) return CANNON_E_INVALID;
  if (t) *t = cw->w.time;
  if (stepnumber) *stepnumber = cw->w.stepnumber;
  return CANNON_OK;
}
int32_t cannon_world_set_stepnumber(cannon_world* cw, int64_t n) {
  if (!cw || n < 0) return CANNON_E_INVALID;
  cw->w.stepnumber = n;
  return CANNON_OK;
}
int32_t cannon_world_set_dt(cannon_world* cw, double dt) {
  if (!cw) return CANNON_E_INVALID;
  cw->w.dt = dt;
  return CANNON_OK;
}

int32_t cannon_apply_gravity(cannon_world* cw) {
  if (!cw) return CANNON_E_INVALID;
  World& w = cw->w;
  const double gx = D(w.desc.gravity[0]), gy = D(w.desc.gravity[1]), gz = D(w.desc.gravity[2]);
  for (Body& bi : w.bodies)
    if (bi.type == CANNON_BODY_DYNAMIC) {
      bi.force.x = (float)(D(bi.force.x) + bi.mass * gx);
      bi.force.y = (float)(D(bi.force.y) + bi.mass * gy);
      bi.force.z = (float)(D(bi.force.z) + bi.mass * gz);
    }
  w.sphUpdate();  // the subsystems run between gravity and the broadphase (world_class.dart:472-475)
  return CANNON_OK;
}

int32_t cannon_broadphase_pairs(cannon_world* cw, int32_t* p1, int32_t* p2, int32_t cap, int32_t* n_pairs) {
  if (!cw || !n_pairs) return CANNON_E_INVALID;
  World& w = cw->w;
  w.collisionPairs();
  *n_pairs = (int32_t)w.p1.size();
  if ((int)w.p1.size() > cap) return fail(cw->ctx, CANNON_E_CAPACITY, "pair buffer too small");
  for (size_t k = 0; k < w.p1.size(); k++) {
    if (p1) p1[k] = w.p1[k];
    if (p2) p2[k] = w.p2[k];
  }
  return CANNON_OK;
}

static int32_t export_contacts(cannon_world* cw, cannon_contacts_soa* out, int32_t* n_contacts) {
  World& w = cw->w;
  const int nc = (int)w.contacts.size();
  if (n_contacts) *n_contacts = nc;
  if (!out) return CANNON_OK;
  if (out->capacity < nc) return fail(cw->ctx, CANNON_E_CAPACITY, "contact buffer too small");
  for (int k = 0; k < nc; k++) {
    const Eq& c = w.contacts[k];
    if (out->body_i) out->body_i[k] = c.bi;
    if (out->body_j) out->body_j[k] = c.bj;
    if (out->ri) PUTF3(out->ri, k, c.ri);
    if (out->rj) PUTF3(out->rj, k, c.rj);
    if (out->ni) PUTF3(out->ni, k, c.ni);
    if (out->restitution) out->restitution[k] = c.restitution;
    if (out->friction) out->friction[k] = c.friction;
    if (out->enabled) out->enabled[k] = c.enabled;
    if (out->multiplier) out->multiplier[k] = c.multiplier;
  }
  return CANNON_OK;
}

int32_t cannon_narrowphase_contacts(cannon_world* cw, const int32_t* p1, const int32_t* p2, int32_t np, cannon_contacts_soa* out,
                                    int32_t* n_contacts, int32_t* per_pair_count) {
  if (!cw || np < 0 || (np > 0 && (!p1 || !p2))) return CANNON_E_INVALID;
  World& w = cw->w;
  const int nb = (int)w.bodies.size();
  w.p1.assign(p1, p1 + np);
  w.p2.assign(p2, p2 + np);
  for (int k = 0; k < np; k++)
    if (p1[k] < 0 || p2[k] < 0 || p1[k] >= nb || p2[k] >= nb) return fail(cw->ctx, CANNON_E_INVALID, "pair references unknown body");
  if (w.dt < 0) w.dt = 1.0 / 60;  // World.defaultDt
  w.getContacts();
  if (w.unsupportedPair) { w.unsupportedPair = false; return fail(cw->ctx, CANNON_E_UNSUPPORTED, "a trimesh met a box / convex / particle / trimesh: the reference's resolvers for these pairs are unfinished (narrow_phase.dart:2265-2341)"); }
  if (per_pair_count)
    for (int k = 0; k < np; k++) per_pair_count[k] = w.perPairCount[k];
  return export_contacts(cw, out, n_contacts);
}

int32_t cannon_solver_solve(cannon_world* cw, double dt, int32_t* iterations_done) {
  if (!cw) return CANNON_E_INVALID;
  World& w = cw->w;
  w.makeContactConstraints();
  int it = w.solve(dt);
  w.prof.n_rows = (int64_t)w.rows.size();
  w.prof.iterations_done = it;
  if (iterations_done) *iterations_done = it;
  return CANNON_OK;
}

int32_t cannon_integrate(cannon_world* cw, double dt) {
  if (!cw) return CANNON_E_INVALID;
  cw->w.integrateAll(dt);
  cw->w.time += dt;
  return CANNON_OK;
}

int32_t cannon_world_step(cannon_world* cw, double dt, int32_t nsteps) {
  if (!cw || nsteps < 0) return CANNON_E_INVALID;
  for (int s = 0; s < nsteps; s++) {
    cw->w.internalStep(dt);
    if (cw->w.unsupportedPair) { cw->w.unsupportedPair = false; return fail(cw->ctx, CANNON_E_UNSUPPORTED, "a trimesh met a box / convex / particle / trimesh: the reference's resolvers for these pairs are unfinished (narrow_phase.dart:2265-2341)"); }
  }
  return CANNON_OK;
}

// The checker has no device and no stages to time: the profiled / asynchronous variants are the plain step.
int32_t cannon_world_step_profiled(cannon_world* cw, double dt, int32_t nsteps) {
  const int32_t rc = cannon_world_step(cw, dt, nsteps);
  if (rc == CANNON_OK) cw->w.prof.sum_steps = nsteps;
  return rc;
}
int32_t cannon_world_step_async(cannon_world* cw, double dt, int32_t nsteps) { return cannon_world_step(cw, dt, nsteps); }
int32_t cannon_ctx_sync(cannon_ctx* ctx) { return ctx ? CANNON_OK : CANNON_E_INVALID; }

int32_t cannon_world_profile(cannon_world* cw, cannon_profile* out) {
  if (!cw || !out) return CANNON_E_INVALID;
  *out = cw->w.prof;
  return CANNON_OK;
}

int32_t cannon_world_get_contacts(cannon_world* cw, cannon_contacts_soa* out, int32_t* n_contacts) {
  if (!cw) return CANNON_E_INVALID;
  return export_contacts(cw, out, n_contacts);
}

int32_t cannon_world_enable_contact_events(cannon_world* cw, int32_t enable) {
  if (!cw) return CANNON_E_INVALID;
  World& w = cw->w;
  w.trackOverlaps = enable != 0;
  w.overlapCurrent.clear(); w.overlapPrevious.clear(); w.additions.clear(); w.removals.clear();
  return CANNON_OK;
}

int32_t cannon_world_get_contact_events(cannon_world* cw, int32_t cap, int32_t* n_begin, int32_t* begin_a, int32_t* begin_b, int32_t* n_end,
                                        int32_t* end_a, int32_t* end_b) {
  if (!cw || cap < 0 || !n_begin || !n_end) return CANNON_E_INVALID;
  const World& w = cw->w;
  if (!w.trackOverlaps) return fail(cw->ctx, CANNON_E_INVALID, "contact events are not enabled");
  const int nb = (int)w.additions.size() / 2, ne = (int)w.removals.size() / 2;
  *n_begin = nb; *n_end = ne;
  if (nb > cap || ne > cap) return fail(cw->ctx, CANNON_E_CAPACITY, "contact event arrays too small");
  for (int k = 0; k < nb; k++) { if (begin_a) begin_a[k] = w.additions[2 * k]; if (begin_b) begin_b[k] = w.additions[2 * k + 1]; }
  for (int k = 0; k < ne; k++) { if (end_a) end_a[k] = w.removals[2 * k]; if (end_b) end_b[k] = w.removals[2 * k + 1]; }
  return CANNON_OK;
}

int32_t cannon_world_get_rows(cannon_world* cw, int32_t cap, int32_t* n_rows, int32_t* body_i, int32_t* body_j, double* B, double* invC,
                              double* lambda, int32_t* level) {
  if (!cw) return CANNON_E_INVALID;
  World& w = cw->w;
  const int n = (int)w.rows.size();
  if (n_rows) *n_rows = n;
  if (cap < n) return fail(cw->ctx, CANNON_E_CAPACITY, "row buffer too small");
  for (int k = 0; k < n; k++) {
    if (body_i) body_i[k] = w.rows[k].bi;
    if (body_j) body_j[k] = w.rows[k].bj;
    if (B) B[k] = w.rows[k].B;
    if (invC) invC[k] = w.rows[k].invC;
    if (lambda) lambda[k] = w.rows[k].lambda;
    if (level) level[k] = w.rows[k].level;
  }
  return CANNON_OK;
}

}  // extern "C"

extern "C" {

void cannon_ray_options_default(cannon_ray_options* o) {
  if (!o) return;
  o->mode = CANNON_RAY_CLOSEST; o->skip_backfaces = 1; o->collision_filter_mask = -1; o->collision_filter_group = -1; o->check_collision_response = 1;
}

int32_t cannon_world_raycast(cannon_world* cw, int32_t n_rays, const float* from, const float* to, const cannon_ray_options* opt, uint8_t* has_hit,
                             cannon_ray_hits_soa* hits, int32_t* n_hits) {
  if (!cw || n_rays < 0 || !opt || !hits || !n_hits || (n_rays > 0 && (!from || !to))) return CANNON_E_INVALID;
  if (opt->mode != CANNON_RAY_CLOSEST && opt->mode != CANNON_RAY_ANY && opt->mode != CANNON_RAY_ALL) return fail(cw->ctx, CANNON_E_INVALID, "ray mode");
  World& w = cw->w;
  for (const Shape& s : w.shapes)
    if (s.type == CANNON_SHAPE_HEIGHTFIELD || s.type == CANNON_SHAPE_TRIMESH) return fail(cw->ctx, CANNON_E_UNSUPPORTED, "heightfield / trimesh rays are outside the hot-path scope (SURVEY.md 8f)");
  const bool all = opt->mode == CANNON_RAY_ALL;
  if (!all && hits->capacity < n_rays) { *n_hits = n_rays; return fail(cw->ctx, CANNON_E_CAPACITY, "hit arrays smaller than n_rays"); }
  std::vector<RayHit> seq;
  int count = 0;
  auto put = [&](int k, const RayHit& h) {
    if (hits->ray) hits->ray[k] = h.ray;
    if (hits->body) hits->body[k] = h.body;
    if (hits->hit_face_index) hits->hit_face_index[k] = h.hitFaceIndex;
    if (hits->distance) hits->distance[k] = h.distance;
    if (hits->hit_point_world) PUTF3(hits->hit_point_world, k, h.hitPointWorld);
    if (hits->hit_normal_world) PUTF3(hits->hit_normal_world, k, h.hitNormalWorld);
    if (hits->shape_ordinal) hits->shape_ordinal[k] = h.shapeOrdinal;
  };
  for (int r = 0; r < n_rays; r++) {
    RayHit res;
    const bool hit = w.raycast(r, V3{from[3 * r], from[3 * r + 1], from[3 * r + 2]}, V3{to[3 * r], to[3 * r + 1], to[3 * r + 2]}, *opt, res, all ? &seq : nullptr);
    if (has_hit) has_hit[r] = hit ? 1 : 0;
    if (!all) { put(r, res); count += hit ? 1 : 0; }
  }
  if (all) {
    *n_hits = (int)seq.size();
    if ((int)seq.size() > hits->capacity) return fail(cw->ctx, CANNON_E_CAPACITY, "hit arrays too small");
    for (size_t k = 0; k < seq.size(); k++) put((int)k, seq[k]);
  } else {
    *n_hits = count;
  }
  return CANNON_OK;
}

int32_t cannon_world_aabb_query(cannon_world* cw, const float* lower, const float* upper, int32_t* bodies, int32_t cap, int32_t* n) {
  if (!cw || !lower || !upper || !n || cap < 0 || (cap > 0 && !bodies)) return CANNON_E_INVALID;
  std::vector<int> res;
  cw->w.aabbQuery(V3{lower[0], lower[1], lower[2]}, V3{upper[0], upper[1], upper[2]}, res);
  *n = (int)res.size();
  if ((int)res.size() > cap) return fail(cw->ctx, CANNON_E_CAPACITY, "body array too small");
  for (size_t k = 0; k < res.size(); k++) bodies[k] = res[k];
  return CANNON_OK;
}

}  // extern "C"

// cannon_batch_*: host glue over the entry points above, shared by both libraries
#include "../cannon_physics_b200/csrc/batch_impl.inc"
