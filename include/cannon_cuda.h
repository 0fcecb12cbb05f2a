/*
 * cannon_cuda.h — C ABI of libcannon_cuda.so, the B200-native (sm_100a) implementation of the
 * per-step rigid-body hot path of Knightro63/cannon_physics:
 *
 *     World.internalStep(dt)            lib/world/world_class.dart:433-701
 *       broadphase.collisionPairs       lib/collision/{broadphase,naive_broadphase,sap_broadphase,grid_broadphase}.dart
 *       narrowphase.getContacts         lib/world/narrow_phase.dart:634-721 (+ resolvers)
 *       solver.solve                    lib/solver/gs_solver.dart:27-133
 *       Body.integrate / sleepTick      lib/objects/rigid_body.dart:627-680, 282-300
 *
 * The same header is implemented twice:
 *   - cannon_physics_b200/csrc  -> libcannon_cuda.so   (the product; CUDA kernels, no CPU fallback)
 *   - oracle/                   -> libcannon_oracle.so (TEST INFRASTRUCTURE ONLY: a sequential CPU
 *                                  restatement of the reference used as the parity checker)
 * so every parity test is "same inputs, two libraries".
 *
 * Conventions
 *   - every function returns an int32_t status (CANNON_OK == 0, < 0 error); nothing throws or aborts
 *     across the boundary; cannon_last_error() gives a human-readable message (replaces the
 *     reference's `throw '<string>'`, e.g. lib/collision/broadphase.dart:40).
 *   - host buffers are caller-owned; device memory is library-owned; handles are opaque.
 *   - body indices are int32_t and equal Body.index (lib/objects/rigid_body.dart:113).
 *   - vector state is float (the reference stores Float32List vectors, lib/math/vec3.dart:2) and is
 *     computed in double exactly like the Dart VM does; scalars the reference keeps as Dart `double`
 *     (mass, damping, radii, SPOOK parameters, dt) are double here.
 *   - quaternions are (x, y, z, w).
 *   - calls on one ctx are not thread-safe and are synchronous at return.
 *   - plain C layout, pointers + sizes only: directly bindable from dart:ffi.
 */
#ifndef CANNON_CUDA_H
#define CANNON_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CANNON_ABI_VERSION 1

/* ---- status codes ---- */
#define CANNON_OK            0
#define CANNON_E_INVALID    -1   /* bad argument                                           */
#define CANNON_E_CUDA       -2   /* CUDA runtime error (message in cannon_last_error)      */
#define CANNON_E_CAPACITY   -3   /* caller buffer too small; required size in the out-param */
#define CANNON_E_UNSUPPORTED -4  /* feature outside the hot-path scope       */
#define CANNON_E_NOGPU      -5   /* no CUDA device: there is deliberately no CPU fallback   */

/* ---- enums (all int32_t on the wire) ---- */
/* ShapeType order is the reference's, lib/rigid_body_shapes/shape.dart:6-18: it decides which
 * shape is passed first to a resolver (lib/world/narrow_phase.dart:706-710). */
enum {
  CANNON_SHAPE_SPHERE = 0, CANNON_SHAPE_PLANE = 1, CANNON_SHAPE_BOX = 2, CANNON_SHAPE_CONVEX = 3,
  CANNON_SHAPE_CYLINDER = 4, CANNON_SHAPE_CAPSULE = 5, CANNON_SHAPE_CONE = 6, CANNON_SHAPE_SIZED_PLANE = 7,
  CANNON_SHAPE_HEIGHTFIELD = 8, CANNON_SHAPE_PARTICLE = 9, CANNON_SHAPE_TRIMESH = 10
};
/* CAPSULE / CONE / SIZED_PLANE are ConvexPolyhedron subclasses in the reference
 * (lib/rigid_body_shapes/{capsule,capsule_lathe,cone,sized_plane}.dart): the binding passes the hull the reference's
 * own constructor built (vertices, faces, convex_has_axes) exactly like CONVEX; the type only selects the resolver and
 * which shape a resolver sees first. The table of narrow_phase.dart:336-473 is reproduced as written, including the
 * pair it cannot reach: its `convexSizedPlane` key is compared with a lower-cased name and never matches, so a plain
 * CONVEX and a SIZED_PLANE never collide. Rays ignore the three types (ray_class.dart:101-123 has no handler).
 *
 * PARTICLE (lib/rigid_body_shapes/particle.dart): a point; sphereParticle / planeParticle / boxParticle / particleConvex /
 * heightfieldParticle (narrow_phase.dart:1258,1805,1731,2179,2343). Its contacts list the particle's body first.
 * particleConvex measures the penetration against ConvexPolyhedron.worldVertices / worldFaceNormals, which the reference
 * computes when their NeedsUpdate flags are set and never invalidates (convex_polyhedron.dart:101-103,603,645;
 * narrow_phase.dart:2207-2212): every hull Shape (and every cached heightfield pillar, heightfield.dart:285-301) keeps
 * the pose of the first penetration it ever saw. That state lives with the shape table here as well: it is part of the
 * world, is cleared by cannon_world_set_shapes, and makes particle-in-hull contacts history dependent exactly as in
 * the reference.
 *
 * TRIMESH (lib/rigid_body_shapes/trimesh.dart): sphereTrimesh and planeTrimesh (narrow_phase.dart:1438,1916) as the Dart
 * port runs them - every triangle is visited in index order (the octree query result is not used, :1490), and the
 * triangle-face test sits inside the per-corner loop, so a face contact is reported three times. The reference's other
 * trimesh resolvers are unfinished (trimeshConvex indexes triangles by vertex index and collides a degenerate hull,
 * :2313-2330; boxTrimesh / particleTrimesh / trimeshTrimesh build on it): a step in which such a pair passes the
 * prologue of getContacts returns CANNON_E_UNSUPPORTED instead of inventing a result. Face normals are those of the
 * unscaled mesh (setScale calls updateNormals before it stores the new scale, trimesh.dart:142-152). Rays refuse worlds
 * with a trimesh like those with a heightfield. */
/* BodyTypes / BodySleepStates, lib/objects/rigid_body.dart:15-16 */
enum { CANNON_BODY_DYNAMIC = 0, CANNON_BODY_STATIC = 1, CANNON_BODY_KINEMATIC = 2 };
enum { CANNON_AWAKE = 0, CANNON_SLEEPY = 1, CANNON_SLEEPING = 2 };
/* World.broadphase choices: NaiveBroadphase / SAPBroadphase / GridBroadphase */
enum { CANNON_BP_NAIVE = 0, CANNON_BP_SAP = 1, CANNON_BP_GRID = 2 };
#ifndef CANNON_COLORED_UNIT_CONTACTS
#define CANNON_COLORED_UNIT_CONTACTS 4   /* part of the COLORED order: changing it changes results (and the fixtures) */
#endif
/* World.solver choices.
 *   REFERENCE_ORDER: GSSolver with the reference's exact equation order
 *       (lib/world/world_class.dart:539-541,562,627-635); bit-reproducible validation mode.
 *   COLORED: graph-coloured Gauss-Seidel, the throughput mode. Same per-row arithmetic as GSSolver (f64 on
 *       f32-stored operands, no FMA, every Vector3 store rounds to float), different - but fully specified -
 *       row order: units = contact manifolds (the ContactEquations of one resolver call, rows [f1,f2,n] per
 *       contact), cut into runs of at most CANNON_COLORED_UNIT_CONTACTS consecutive contacts (a colour lasts as
 *       long as its longest unit: 30-row manifolds made every colour of a pile a 30-step chain), followed by one
 *       unit per constraint; unit key = index of its first ContactEquation in
 *       World.contacts (constraints: n_contacts + ordinal of its first accepted equation), counted inside the
 *       unit's world for a batch (n_worlds > 1: contacts from the world's first ContactEquation, constraints
 *       from the world's contact count), so a world's result does not depend on the rest of the batch; priority =
 *       key * 2654435761 mod 2^32; colour(u) = round in which u holds the smallest pending priority on all of
 *       its movable bodies. Colours are swept in ascending order, units of a colour are independent. The oracle
 *       restates exactly this order sequentially, so COLORED is bit-exact against it; agreement
 *       with the reference's own insertion order is statistical (Gauss-Seidel is order dependent).
 *   COLORED_F32: the same colour order with rows packed to f32 and swept with FMA (reduced precision; opt-in,
 *       never used for a headline number).
 *   SPLIT: SplitSolver(GSSolver) (lib/solver/split_solver.dart:50-120): islands of non-static bodies
 *       connected by equations, one independent GSSolver pass per island (own tolerance early-exit),
 *       island equations in descending Equation.id order. The reference's ids depend on its object-pool
 *       history; here ids are the history-free creation order of a pool-less step
 *       (constraint equations first, then per contact: contact, friction 1, friction 2). */
enum { CANNON_SOLVER_REFERENCE_ORDER = 0, CANNON_SOLVER_COLORED = 1, CANNON_SOLVER_SPLIT = 2, CANNON_SOLVER_COLORED_F32 = 3 };
/* Constraint kinds, lib/constraints/{point_to_point,hinge}_constraint.dart */
enum {
  CANNON_CONSTRAINT_POINT_TO_POINT = 0, /* lib/constraints/point_to_point_constraint.dart */
  CANNON_CONSTRAINT_HINGE = 1,          /* hinge_constraint.dart */
  CANNON_CONSTRAINT_DISTANCE = 2,       /* distance_constraint.dart */
  CANNON_CONSTRAINT_LOCK = 3,           /* lock_constraint.dart */
  CANNON_CONSTRAINT_CONE_TWIST = 4      /* cone_twist_constraint.dart */
};

typedef struct cannon_ctx   cannon_ctx;
typedef struct cannon_world cannon_world;

/* ContactMaterial, lib/material/contact_material.dart:5-76 */
typedef struct cannon_contact_material {
  int32_t material_a;            /* Material index (ignored for the world default) */
  int32_t material_b;
  double  friction;                       /* default 0.3 */
  double  restitution;                    /* default 0.3 (0.0 for the world default, world_class.dart:155-158) */
  double  contact_equation_stiffness;     /* 1e7 */
  double  contact_equation_relaxation;    /* 3   */
  double  friction_equation_stiffness;    /* 1e7 */
  double  friction_equation_relaxation;   /* 3   */
} cannon_contact_material;

/* World constructor parameters, lib/world/world_class.dart:135-162, plus the pluggable
 * Broadphase / Solver objects flattened to POD. */
typedef struct cannon_world_desc {
  float   gravity[3];
  float   friction_gravity[3];
  int32_t has_friction_gravity;  /* World.frictionGravity != null */
  int32_t allow_sleep;           /* World.allowSleep */
  int32_t quat_normalize_skip;   /* World.quatNormalizeSkip */
  int32_t quat_normalize_fast;   /* World.quatNormalizeFast */
  int32_t solver_kind;           /* CANNON_SOLVER_* */
  int32_t solver_iterations;     /* Solver.iterations (10) */
  double  solver_tolerance;      /* Solver.tolerance (1e-7) */
  int32_t broadphase_kind;       /* CANNON_BP_* */
  int32_t use_bounding_boxes;    /* Broadphase.useBoundingBoxes */
  int32_t sap_axis;              /* SAPBroadphase.axisIndex: 0 x, 1 y, 2 z */
  int32_t grid_nx, grid_ny, grid_nz;      /* GridBroadphase.nx/ny/nz */
  float   grid_min[3], grid_max[3];       /* GridBroadphase.aabbMin/aabbMax */
  cannon_contact_material default_contact_material;  /* World.defaultContactMaterial */
  int32_t n_worlds;              /* >1: a batch of independent worlds in one handle; bodies carry world_id.
                                    Pairs are only formed inside a world; 0/1 = single world */
  int32_t max_pairs;             /* capacities, 0 = library picks (grows on demand) */
  int32_t max_contacts;
} cannon_world_desc;

/* Shape table entry (one per distinct Shape object; bodies reference it by index, the reference's
 * demos share one Shape between many bodies, examples/lib/examples/container.dart:105). */
typedef struct cannon_shape_desc {
  int32_t type;                  /* CANNON_SHAPE_* */
  int32_t collision_response;    /* Shape.collisionResponse (1) */
  int32_t collision_filter_group;/* Shape.collisionFilterGroup (-1) */
  int32_t collision_filter_mask; /* Shape.collisionFilterMask (-1) */
  double  radius;                /* Sphere.radius, lib/rigid_body_shapes/sphere.dart:11 */
  float   half_extents[3];       /* Box.halfExtents, lib/rigid_body_shapes/box.dart:14 */
  double  radius_top, radius_bottom, height; /* Cylinder, lib/rigid_body_shapes/cylinder.dart:15 */
  int32_t num_segments;
  /* ConvexPolyhedron, lib/rigid_body_shapes/convex_polyhedron.dart:51 (faces CCW, CSR layout) */
  int32_t n_vertices;
  const float*   vertices;       /* 3*n_vertices */
  int32_t n_faces;
  const int32_t* face_offsets;   /* n_faces+1 */
  const int32_t* face_indices;
  /* Heightfield, lib/rigid_body_shapes/heightfield.dart:34: data[ix][iy] row-major [nx][ny], f64 */
  int32_t hf_nx, hf_ny;
  const double*  hf_data;
  int32_t hf_element_size;       /* int in the reference (heightfield.dart:46) */
  /* CONVEX / CAPSULE / CONE / SIZED_PLANE: ConvexPolyhedron.uniqueAxes != null (convex_polyhedron.dart:105). Only its
   * presence matters: findSeparatingAxis (:255,:290) tests the hull's face normals when axes were given and none when
   * not (its `else if` on the same condition is dead code). Cone passes axes, Capsule / SizedPlane / Lathe do not. */
  int32_t convex_has_axes;
  /* Trimesh, lib/rigid_body_shapes/trimesh.dart:37: `vertices` / `n_vertices` above are the mesh vertices (the reference
   * keeps doubles and rounds them to float in getVertex, :260-275); tm_indices: 3 per triangle; tm_scale: Trimesh.scale */
  int32_t n_triangles;
  const int32_t* tm_indices;
  float   tm_scale[3];
  /* Shape.material (shape.dart:48) as an index into the material table, -1 = null (cannon_shape_desc_default sets -1).
   * Reproduced as the reference uses it: the contact material of a shape pair is getContactMaterial(si.material,
   * sj.material) when both shapes have one, else the bodies', else the default (narrow_phase.dart:692-696);
   * createContactEquation takes `shape.material ?? body.material` of the shapes and bodies in resolver order (:517-521; a
   * heightfield pillar has none), World.internalStep then overrides the restitution by the two BODY materials when both
   * exist (world_class.dart:556-560); createFrictionEquationsFromContact pairs the shapes in PAIR order (c.si = rsi) with
   * the bodies in RESOLVER order (c.bi), so a pair whose shapes arrive swapped mixes one body's shape material with the
   * other body's body material (:541-542). */
  int32_t material;
} cannon_shape_desc;

/* Body state, structure of arrays. In *_set_bodies a NULL pointer means "reference default"
 * (lib/objects/rigid_body.dart:27-48); in *_get_bodies a NULL pointer means "not wanted".
 * One shape per body, at the body origin (compound bodies: SURVEY.md §8f). */
typedef struct cannon_bodies_soa {
  int32_t  n;
  float*   position;          /* 3n */
  float*   quaternion;        /* 4n (x,y,z,w); default (0,0,0,1) */
  float*   velocity;          /* 3n */
  float*   angular_velocity;  /* 3n */
  float*   force;             /* 3n */
  float*   torque;            /* 3n */
  double*  mass;              /* n; default 0 (=> static) */
  int32_t* type;              /* n; default: mass<=0 ? STATIC : DYNAMIC */
  int32_t* sleep_state;       /* n; default AWAKE */
  double*  time_last_sleepy;  /* n; default 0 */
  uint8_t* allow_sleep;       /* n; default 1 */
  double*  sleep_speed_limit; /* n; default 0.1 */
  double*  sleep_time_limit;  /* n; default 1 */
  double*  linear_damping;    /* n; default 0.01 */
  double*  angular_damping;   /* n; default 0.01 */
  float*   linear_factor;     /* 3n; default 1,1,1 */
  float*   angular_factor;    /* 3n; default 1,1,1 */
  uint8_t* fixed_rotation;    /* n; default 0 */
  int32_t* collision_filter_group; /* n; default 1 */
  int32_t* collision_filter_mask;  /* n; default -1 */
  uint8_t* collision_response;     /* n; default 1 */
  uint8_t* is_trigger;        /* n; default 0 */
  int32_t* material;          /* n; Material index or -1 (default) */
  int32_t* shape;             /* n; index into the shape table, -1 = no shape */
  int32_t* world_id;          /* n; batch mode only, default 0 */
  /* derived by the library at set time (Body.updateMassProperties / updateBoundingRadius,
   * rigid_body.dart:587-609,395-412); only written by *_get_bodies */
  double*  inv_mass;          /* n */
  float*   inv_inertia;       /* 3n  local diagonal */
  float*   inv_inertia_world; /* 9n  row-major */
  double*  bounding_radius;   /* n */
  float*   aabb;              /* 6n  lower xyz, upper xyz */
} cannon_bodies_soa;

/* Constraint, lib/constraints/point_to_point_constraint.dart:20, hinge_constraint.dart:10, distance_constraint.dart:7,
 * lock_constraint.dart:9, cone_twist_constraint.dart:11. Constructors that read body state (LockConstraint's pivots and
 * frame vectors, DistanceConstraint's default distance) are evaluated by cannon_world_set_constraints on the bodies'
 * state at that moment, like `new XConstraint(bodyA, bodyB)` would. */
typedef struct cannon_constraint_desc {
  int32_t type;               /* CANNON_CONSTRAINT_* */
  int32_t body_a, body_b;
  float   pivot_a[3], pivot_b[3]; /* point-to-point, hinge, cone-twist (lock computes its own, distance has none) */
  float   axis_a[3], axis_b[3];   /* hinge: normalised by the library like hinge_constraint.dart:34-37; cone-twist: used as given */
  double  max_force;              /* 1e6 */
  int32_t collide_connected;      /* Constraint.collideConnected */
  int32_t motor_enabled;          /* hinge: RotationalMotorEquation.enabled */
  double  motor_target_velocity;
  double  motor_max_force;
  double  distance;               /* distance constraint: < 0 => bodyA.position.distanceTo(bodyB.position) at set time */
  double  angle;                  /* cone-twist: ConeEquation.angle */
  double  twist_angle;            /* cone-twist: maxAngle of the twist RotationalEquation */
  /* The poses the constraint's CONSTRUCTOR saw, for constraints that were made earlier than this upload (a world rebuilt
   * after bodies moved): has_ctor_pose != 0 makes LockConstraint's pivots / frame vectors (lock_constraint.dart:29-43) and
   * DistanceConstraint's default distance (distance_constraint.dart:16) come from these instead of the current poses. */
  int32_t has_ctor_pose;
  float   ctor_pos_a[3], ctor_quat_a[4], ctor_pos_b[3], ctor_quat_b[4];
} cannon_constraint_desc;

/* Spring, lib/objects/spring.dart:17. The reference applies springs from user code, canonically
 *   world.addEventListener('postStep', (e) { for (final s in springs) s.applyForce(); })
 * (examples/lib/examples/spring.dart:90,123); the library runs exactly that in the postStep slot of every step
 * (world_class.dart:685, after clearForces), in array order, so the forces act in the next step's integration. */
typedef struct cannon_spring_desc {
  int32_t body_a, body_b;
  double  rest_length;            /* 1 */
  double  stiffness;              /* 100 */
  double  damping;                /* 1 */
  float   local_anchor_a[3], local_anchor_b[3];
} cannon_spring_desc;

/* ContactEquation list produced by the narrowphase (lib/equations/contact_equation.dart). */
typedef struct cannon_contacts_soa {
  int32_t  capacity;          /* number of contacts the arrays can hold */
  int32_t* body_i;            /* ContactEquation.bi (index) */
  int32_t* body_j;            /* ContactEquation.bj */
  float*   ri;                /* 3*capacity */
  float*   rj;
  float*   ni;
  double*  restitution;       /* optional (may be NULL) */
  double*  friction;          /* optional: mu used for the two FrictionEquations, <=0 => none */
  uint8_t* enabled;           /* optional */
  double*  multiplier;        /* optional: Equation.multiplier of the contact row after the last solve */
} cannon_contacts_soa;

/* Profile, lib/world/world_class.dart:27-41, in milliseconds (float here, int in the reference),
 * plus counters of the last step. */
typedef struct cannon_profile {
  double solve, make_contact_constraints, broadphase, integrate, narrowphase;
  int64_t n_pairs, n_contacts, n_rows, n_levels, iterations_done;
  int64_t steps, contact_iters_total;  /* accumulated since world creation */
  /* device times (CUDA events on the library's stream), milliseconds */
  double step_call_ms;   /* whole last cannon_world_step call (all nsteps) */
  double schedule_ms;    /* dependency-level / colouring kernel of the last step */
  double gs_ms;          /* Gauss-Seidel sweep kernel of the last step */
  int64_t kernel_launches; /* kernels launched by the library since world creation */
  int64_t n_tasks;         /* narrowphase resolver tasks of the last step (pairs + heightfield pillars) */
  int64_t n_islands;       /* SPLIT solver: islands of the last solve (SplitSolver.solve's return value) */
  int64_t n_tasks_by_type[8]; /* sphere-sphere, sphere-plane, sphere-box, sphere-hull, plane-hull, hull-hull, sphere-pillar, hull-pillar */
  /* cannon_world_step_profiled: device time per stage summed over the sum_steps steps of that call, milliseconds */
  int64_t sum_steps;
  double sum_step_ms, sum_broadphase, sum_narrowphase, sum_solve, sum_integrate, sum_schedule, sum_gs;
} cannon_profile;

/* ---- lifecycle ---- */
int32_t     cannon_version(void);
/* "cuda" for the product, "oracle" for the CPU checker */
const char* cannon_backend(void);
int32_t     cannon_ctx_create(int32_t device, cannon_ctx** out);
void        cannon_ctx_destroy(cannon_ctx* ctx);
const char* cannon_last_error(const cannon_ctx* ctx);
/* fills the POD with the reference defaults (World(), GSSolver(), NaiveBroadphase()) */
void        cannon_world_desc_default(cannon_world_desc* d);
void        cannon_shape_desc_default(cannon_shape_desc* d);

/* ---- world ---- */
int32_t cannon_world_create(cannon_ctx* ctx, const cannon_world_desc* desc, cannon_world** out);
void    cannon_world_destroy(cannon_world* w);
/* Material table (friction / restitution, -1 = unset, lib/material/material.dart:20-21) and the
 * ContactMaterial table looked up by unordered material pair (World.addContactMaterial). */
int32_t cannon_world_set_materials(cannon_world* w, int32_t n_materials, const double* friction,
                                   const double* restitution, int32_t n_contact_materials,
                                   const cannon_contact_material* cms);
int32_t cannon_world_set_shapes(cannon_world* w, int32_t n_shapes, const cannon_shape_desc* shapes);
/* Compound bodies: Body.addShape(shape, offset, orientation), lib/objects/rigid_body.dart:348-377 (shapes, shapeOffsets,
 * shapeOrientations). The table replaces the `shape` column of the following cannon_world_set_bodies calls (until it is
 * replaced or dropped), which must describe the same n_bodies: body b owns the shape instances [first[b], first[b+1]) in addShape order (an empty
 * range = a body without shapes). offset: 3 floats per instance, NULL = zeros; orientation: 4 floats (x,y,z,w),
 * NULL = identity. n_bodies = 0 drops the table (one shape per body at the body origin again). What follows the
 * reference: Body.updateAABB / updateBoundingRadius / updateMassProperties over all instances (rigid_body.dart:395-447,
 * 587-609), the shape-pair loops of Narrowphase.getContacts (narrow_phase.dart:669-721: every resolver sees the shape's
 * world pose and makes ri / rj relative to the BODY position), Ray.intersectBody (ray_class.dart:226-243);
 * GridBroadphase bins a body by shapes[0] (grid_broadphase.dart:136), as the reference does. */
int32_t cannon_world_set_body_shapes(cannon_world* world, int32_t n_bodies, const int32_t* first, const int32_t* shape,
                                     const float* offset, const float* orientation);

/* World.addBody for all bodies at once (upload path); derives mass properties. */
int32_t cannon_world_set_bodies(cannon_world* w, const cannon_bodies_soa* bodies);
int32_t cannon_world_get_bodies(cannon_world* w, cannon_bodies_soa* out);
/* SPHSystem, lib/objects/sph_system.dart:6 (World.subsystems, updated after gravity and before the broadphase,
 * world_class.dart:472-475). particles: body indices in SPHSystem.add order. update() (:62-163) is reproduced as written:
 * neighbours are the particles within smoothingRadius in list order with the particle itself appended last; the
 * pressure / viscosity sums read pressures[j] / densities[j] with j = the POSITION in that neighbour list, not the
 * neighbour's own index (:131-133,142). math.pow(x, 2 | 3) of the kernel functions (:166-182) is evaluated correctly
 * rounded (x*x; x*x*x through an exact product), pow(h, 9) by the host's libm. */
typedef struct cannon_sph_desc {
  int32_t n_particles;
  const int32_t* particles;
  double density;            /* 1 */
  double smoothing_radius;   /* 1 */
  double speed_of_sound;     /* 1 */
  double viscosity;          /* 0.01 */
  double eps;                /* 0.00001 */
} cannon_sph_desc;
void    cannon_sph_desc_default(cannon_sph_desc* d);
int32_t cannon_world_set_sph_systems(cannon_world* world, int32_t n, const cannon_sph_desc* systems);

/* World.addConstraint for all constraints at once. */
int32_t cannon_world_set_constraints(cannon_world* w, int32_t n, const cannon_constraint_desc* cs);
/* Springs applied in every step's postStep slot (see cannon_spring_desc); n = 0 removes them. */
int32_t cannon_world_set_springs(cannon_world* w, int32_t n, const cannon_spring_desc* springs);
/* World.time (used by Body.sleepTick, world_class.dart:693) */
int32_t cannon_world_set_time(cannon_world* w, double time);
int32_t cannon_world_get_time(cannon_world* w, double* time, int64_t* stepnumber);
/* World.stepnumber (world_class.dart:696; the phase of quatNormalizeSkip, rigid_body.dart:637): restores it when a world
 * is rebuilt from a checkpoint (cannon_world_get_bodies / set_bodies round trip) */
int32_t cannon_world_set_stepnumber(cannon_world* w, int64_t stepnumber);

/* ---- staged entry points (drop-in for World.broadphase / narrowphase / solver) ---- */
/* Broadphase.collisionPairs(world,p1,p2), lib/collision/broadphase.dart:39, followed by the
 * constraint-pair filter of world_class.dart:488-499. Pairs come out in the reference's order. */
int32_t cannon_broadphase_pairs(cannon_world* w, int32_t* p1, int32_t* p2, int32_t cap, int32_t* n_pairs);
/* Narrowphase.getContacts(p1,p2,...), lib/world/narrow_phase.dart:634. per_pair_count (np entries,
 * may be NULL) receives the number of ContactEquations generated per pair. */
int32_t cannon_narrowphase_contacts(cannon_world* w, const int32_t* p1, const int32_t* p2, int32_t np,
                                    cannon_contacts_soa* out, int32_t* n_contacts, int32_t* per_pair_count);
/* World.dt used by the staged narrowphase for the SPOOK parameters (world_class.dart:434; -1 until
 * the first step, in which case World.defaultDt = 1/60 is used) */
int32_t cannon_world_set_dt(cannon_world* w, double dt);
/* gravity accumulation of world_class.dart:460-471 (first thing internalStep does) */
int32_t cannon_apply_gravity(cannon_world* w);
/* world_class.dart:539-645: wake-up flags, Constraint.update(), equation assembly in the reference
 * order and Solver.solve(dt, world) over the contacts of the last cannon_narrowphase_contacts call plus
 * the world's constraints; updates velocity / angularVelocity like gs_solver.dart:111-121. Returns the
 * iteration count like GSSolver.solve. */
int32_t cannon_solver_solve(cannon_world* w, double dt, int32_t* iterations_done);
/* damping + Body.integrate + clearForces + sleepTick (world_class.dart:648-700) */
int32_t cannon_integrate(cannon_world* w, double dt);

/* ---- fused ---- */
/* nsteps x World.step(dt) (fixed stepping, world_class.dart:393-399) with all state device-resident */
int32_t cannon_world_step(cannon_world* w, double dt, int32_t nsteps);
int32_t cannon_world_profile(cannon_world* w, cannon_profile* out);
/* the same nsteps steps, every one launched eagerly between its own stage events: cannon_profile.sum_* then hold the
 * device time of every stage summed over the call (World.profile of lib/world/world_class.dart:27-41 accumulated) */
int32_t cannon_world_step_profiled(cannon_world* w, double dt, int32_t nsteps);
/* World.step without waiting for the device (SURVEY.md 8b `_async` + cannon_ctx_sync): returns once the steps are
 * enqueued on the ctx's stream, so one host thread / Dart isolate can drive one ctx per GPU concurrently. Counters,
 * capacity errors and cannon_world_profile of the call become valid after cannon_ctx_sync(ctx), which returns the first
 * error of the collected calls. Any synchronous entry point on the same ctx also completes the pending work first. */
int32_t cannon_world_step_async(cannon_world* w, double dt, int32_t nsteps);
int32_t cannon_ctx_sync(cannon_ctx* ctx);
/* World.contacts of the last step */
int32_t cannon_world_get_contacts(cannon_world* w, cannon_contacts_soa* out, int32_t* n_contacts);
/* Contact events of the last step: World.emitContactEvents (lib/world/world_class.dart:703-730)
 * over bodyOverlapKeeper (lib/collision/overlap_keeper.dart:20-96, filled per contact at world_class.dart:606, ticked at
 * :219). Tracking is off until enabled; enabling starts from an empty "previous" set, like a new World. After a step,
 * begin_* lists the body pairs that have a ContactEquation now and had none in the previous step (`beginContact`),
 * end_* the pairs that lost theirs (`endContact`); a < b inside a pair and pairs ascend by (a, b) - the order of
 * OverlapKeeper.getDiff, whose key (i << 16) | j is injective below 65536 bodies (64-bit keys are used here).
 * With one shape per body `beginShapeContact` / `endShapeContact` (:732-769) are the same lists. The reference's
 * `collide` event fires for every contact of every step as written (collisionMatrixPrevious aliases collisionMatrix,
 * world_class.dart:213-217,595): it is cannon_world_get_contacts. A multi-step call keeps the events of its last step.
 * cap = capacity of each of the four arrays; CANNON_E_CAPACITY reports the needed size through n_begin / n_end. */
int32_t cannon_world_enable_contact_events(cannon_world* w, int32_t enable);
int32_t cannon_world_get_contact_events(cannon_world* w, int32_t cap, int32_t* n_begin, int32_t* begin_a, int32_t* begin_b,
                                        int32_t* n_end, int32_t* end_a, int32_t* end_b);
/* solver rows of the last solve, in solve order: debug / parity only. Arrays of `cap` entries (any may be NULL) */
int32_t cannon_world_get_rows(cannon_world* w, int32_t cap, int32_t* n_rows, int32_t* body_i, int32_t* body_j,
                              double* B, double* invC, double* lambda, int32_t* level);

/* user mutations between steps (Body.applyForce etc. reduce to writing these arrays on the host
 * and re-uploading the dirty range): partial update of dynamic state for bodies [first, first+count) */
int32_t cannon_world_update_bodies(cannon_world* w, int32_t first, int32_t count, const float* position,
                                   const float* quaternion, const float* velocity, const float* angular_velocity,
                                   const float* force, const float* torque);

/* ---- ray casts and AABB queries ----
 * World.raycastClosest / raycastAny / raycastAll (lib/world/world_class.dart:248-277) = Ray.intersectWorld
 * (lib/collision/ray_class.dart:175-199) for n_rays rays at once against the bodies' CURRENT poses: the ray's AABB against
 * every body's AABB (NaiveBroadphase.aabbQuery, naive_broadphase.dart:39-56), the collision filters of Ray.intersectBody
 * (:201-225), the bounding-sphere rejection of _intersectShape (:270-283), then _intersectSphere / _intersectPlane /
 * _intersectBox / _intersectConvex (:285-326,411-553) in the reference's arithmetic.
 *   CLOSEST / ANY: entry r of the hit arrays is ray r's RaycastResult (body -1 and distance -1 without a hit; hit_face_index
 *     is the face of the LAST reported intersection in CLOSEST mode, as ray_class.dart:664 writes it before looking at the
 *     mode); *n_hits = rays that hit; capacity >= n_rays.
 *   ALL: the hit arrays receive the callback sequence of ray 0, then ray 1, ... (`ray` tells which); *n_hits = their number;
 *     CANNON_E_CAPACITY reports the needed capacity through *n_hits.
 * has_hit (n_rays entries, may be NULL) is intersectWorld's return value per ray.
 * Candidates are visited in body-index order for every broadphase kind (SAPBroadphase.aabbQuery would walk its axis list,
 * GridBroadphase has none: broadphase.dart:151-154) - it decides ties between equidistant hits and the sequence of ALL.
 * Worlds with a Heightfield shape are refused (CANNON_E_UNSUPPORTED). The reference's own _intersectHeightfield
 * (ray_class.dart:344-409) takes its cell range from a list that Heightfield.getIndexOfPosition only ever appends to
 * (heightfield.dart:173: `result.addAll([xi, yi])`, read back as index[0] / index[1] at ray_class.dart:367-372), i.e. from the
 * FIRST heightfield ray a Ray object ever cast - there is no history-free behaviour to reproduce. */
enum { CANNON_RAY_CLOSEST = 1, CANNON_RAY_ANY = 2, CANNON_RAY_ALL = 4 };  /* RayMode, ray_class.dart:9-21 */
typedef struct cannon_ray_options {
  int32_t mode;                     /* CANNON_RAY_* */
  int32_t skip_backfaces;           /* 1: RayOptions.skipBackfaces ?? true (ray_class.dart:178) */
  int32_t collision_filter_mask;    /* -1 */
  int32_t collision_filter_group;   /* -1 */
  int32_t check_collision_response; /* 1 */
} cannon_ray_options;
typedef struct cannon_ray_hits_soa {
  int32_t  capacity;
  int32_t* ray;               /* ALL: index of the ray; CLOSEST / ANY: r */
  int32_t* body;              /* RaycastResult.body (index) or -1 */
  int32_t* hit_face_index;    /* RaycastResult.hitFaceIndex */
  double*  distance;          /* RaycastResult.distance */
  float*   hit_point_world;   /* 3 per entry */
  float*   hit_normal_world;  /* 3 per entry */
  int32_t* shape_ordinal;     /* RaycastResult.shape as its position in Body.shapes (0 for a single-shape body), -1 without a hit; may be NULL */
} cannon_ray_hits_soa;
void    cannon_ray_options_default(cannon_ray_options* o);
int32_t cannon_world_raycast(cannon_world* w, int32_t n_rays, const float* from, const float* to, const cannon_ray_options* opt,
                             uint8_t* has_hit, cannon_ray_hits_soa* hits, int32_t* n_hits);
/* Broadphase.aabbQuery (naive_broadphase.dart:39-56): indices of the bodies whose current AABB overlaps [lower, upper],
 * ascending; CANNON_E_CAPACITY reports the needed size through *n. */
int32_t cannon_world_aabb_query(cannon_world* w, const float* lower, const float* upper, int32_t* bodies, int32_t cap, int32_t* n);

/* ---- batches of independent worlds over the GPUs of one box ----
 * n_worlds worlds of bodies_per_world bodies each (config 4: the RL / parameter-sweep case; the reference equivalent is a
 * Dart program holding n_worlds World objects and calling World.step on each, lib/world/world_class.dart:392-431).
 * Worlds are split into ngpu contiguous shards, one cannon_world handle with desc.n_worlds = its share, on its own cannon_ctx per
 * device; body arrays are world-major (body b of world w = index w * bodies_per_world + b) and constraints use these
 * global indices. cannon_batch_step enqueues every shard's steps without waiting (cannon_world_step_async) and then
 * synchronises all of them (cannon_ctx_sync): one host thread / Dart isolate keeps all GPUs busy, with no cross-GPU
 * traffic on the step path. A world's results do not depend on ngpu. */
typedef struct cannon_batch cannon_batch;
#define CANNON_BATCH_MAX_GPUS 16
typedef struct cannon_batch_statistics {
  int32_t n_gpus, n_worlds, bodies_per_world, pad0;
  int64_t n_pairs, n_contacts, n_rows;   /* last step, summed over the shards */
  int64_t iterations_done;               /* last step, max over the shards */
  int64_t steps;                         /* steps taken since creation */
  int64_t contact_iters_total;           /* accumulated, summed over the shards */
  double  step_call_ms_max;              /* device time of the last cannon_batch_step call: max over the shards */
  double  gpu_step_call_ms[CANNON_BATCH_MAX_GPUS];
  int32_t gpu_worlds[CANNON_BATCH_MAX_GPUS];
} cannon_batch_statistics;
int32_t     cannon_batch_create(const int32_t* devices, int32_t ngpu, const cannon_world_desc* desc, int32_t n_worlds,
                                int32_t bodies_per_world, cannon_batch** out);
void        cannon_batch_destroy(cannon_batch* b);
const char* cannon_batch_last_error(const cannon_batch* b);
int32_t     cannon_batch_set_materials(cannon_batch* b, int32_t n_materials, const double* friction, const double* restitution,
                                       int32_t n_contact_materials, const cannon_contact_material* cms);
int32_t     cannon_batch_set_shapes(cannon_batch* b, int32_t n_shapes, const cannon_shape_desc* shapes);
int32_t     cannon_batch_set_body_shapes(cannon_batch* b, int32_t n_bodies, const int32_t* first, const int32_t* shape,
                                         const float* offset, const float* orientation);  /* cannon_world_set_body_shapes over the whole batch, before set_bodies */
int32_t     cannon_batch_set_bodies(cannon_batch* b, const cannon_bodies_soa* bodies);   /* world_id is derived, not read */
int32_t     cannon_batch_set_constraints(cannon_batch* b, int32_t n, const cannon_constraint_desc* cs);
int32_t     cannon_batch_step(cannon_batch* b, double dt, int32_t nsteps);
int32_t     cannon_batch_stats(cannon_batch* b, cannon_batch_statistics* out);
int32_t     cannon_batch_get_bodies(cannon_batch* b, cannon_bodies_soa* out);           /* out->n = n_worlds * bodies_per_world */
/* the shard of GPU `gpu`: its first world, its world count and the cannon_world handle that serves the per-world entry points */
int32_t     cannon_batch_shard(cannon_batch* b, int32_t gpu, int32_t* first_world, int32_t* n_worlds, cannon_world** world);

/* Body.sleep() / Body.wakeUp() (lib/objects/rigid_body.dart:263-278) for bodies [first, first+count): only sleepState is
 * written (Body.sleep also zeroes the velocities: send those through cannon_world_update_bodies). No other per-body
 * state - sleep timers, mass properties, the contact-event sets - is touched. */
int32_t cannon_world_update_sleep_states(cannon_world* w, int32_t first, int32_t count, const int32_t* sleep_state);
/* Body.invInertia (local diagonal, 3 floats per body) for bodies [first, first+count), replacing what
 * cannon_world_set_bodies derived from the pose at upload time: the reference computes it once, in the Body constructor /
 * addShape (rigid_body.dart:85,362,587-609), so a world rebuilt from a checkpoint carries the original values over.
 * invInertiaWorld is refreshed from the current orientation (rigid_body.dart:450-466). */
int32_t cannon_world_set_inv_inertia(cannon_world* w, int32_t first, int32_t count, const float* inv_inertia);
/* HingeConstraint.enableMotor / disableMotor / setMotorSpeed / setMotorMaxForce (lib/constraints/hinge_constraint.dart:
 * 56-76) for constraint `constraint` (its index in the array given to cannon_world_set_constraints): the motor equation's
 * enabled flag, targetVelocity and maxForce = -minForce are replaced; takes effect at the next step like in the
 * reference, and nothing else (body state, the other constraints' frozen parameters) changes. */
int32_t cannon_world_set_hinge_motor(cannon_world* w, int32_t constraint, int32_t enabled, double target_velocity, double max_force);

#ifdef __cplusplus
}
#endif
#endif /* CANNON_CUDA_H */
