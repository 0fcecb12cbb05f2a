// golden_dump.dart — replay a scene exported by export_scene.py through the UNMODIFIED reference (package:cannon_physics)
// and dump what compare_dump.py checks the CPU oracle against. See README.md. Needs a Dart SDK (none in the build container).
//
//   dart run tool/golden_dump.dart scene.json dump.json
//
// Only public reference API is used: World / Body / shapes / materials / constraints are constructed exactly as a user
// program would (lib/world/world_class.dart:135-162, lib/objects/rigid_body.dart:27-86, lib/constraints/*.dart), bodies are
// added in index order, constraints in list order, and the world is stepped with world.step(dt)
// (world_class.dart:392-399, fixed stepping). Field names of the JSON are those of include/cannon_cuda.h.
import 'dart:convert';
import 'dart:io';
import 'package:cannon_physics/cannon_physics.dart';
import 'package:vector_math/vector_math.dart';

Vector3 v3(List a, [int o = 0]) => Vector3((a[o] as num).toDouble(), (a[o + 1] as num).toDouble(), (a[o + 2] as num).toDouble());
double d(dynamic x, double dflt) => x == null ? dflt : (x as num).toDouble();

Shape makeShape(Map s) {
  switch (s['type'] as int) {
    case 0: return Sphere(d(s['radius'], 1.0));
    case 1: return Plane();
    case 2: return Box(v3(s['half_extents']));
    case 3:
    case 5:
    case 6:
    case 7:
      // hull subclasses: replay the reference constructor named by the exporter, so their vertex generation is pinned too
      final c = s['_ctor'] as Map?;
      if (c != null) {
        switch (c['kind'] as String) {
          case 'Cone': return Cone(radius: d(c['radius'], 1), height: d(c['height'], 1), numSegments: (c['numSegments'] ?? 8) as int);
          case 'Capsule': return Capsule(radiusTop: d(c['radiusTop'], 1), radiusBottom: d(c['radiusBottom'], 1), height: d(c['height'], 1),
              numSegments: (c['numSegments'] ?? 8) as int, numHeightSegments: (c['numHeightSegments'] ?? 4) as int);
          case 'CapsuleLathe': return CapsuleLathe(radiusTop: d(c['radiusTop'], 1), radiusBottom: d(c['radiusBottom'], 1), height: d(c['height'], 1),
              numSegments: (c['numSegments'] ?? 8) as int, numHeightSegments: (c['numHeightSegments'] ?? 4) as int);
          case 'SizedPlane': return SizedPlane(d(c['width'], 1), d(c['height'], 1));
          case 'LatheShape': return LatheShape(points: [for (final p in c['points'] as List) Vector2(((p as List)[0] as num).toDouble(), (p[1] as num).toDouble())],
              numSegments: (c['numSegments'] ?? 8) as int, phiStart: d(c['phiStart'], 0), phiLength: d(c['phiLength'], 6.283185307179586));
        }
      }
      final verts = <Vector3>[for (final p in s['vertices'] as List) v3(p as List)];
      final faces = <List<int>>[for (final f in s['faces'] as List) [for (final i in f as List) i as int]];
      return ConvexPolyhedron(vertices: verts, faces: faces, axes: (s['convex_has_axes'] ?? 0) != 0 ? [Vector3(0, 1, 0)] : null,
          type: ShapeType.values[s['type'] as int]);
    case 9: return Particle();
    case 10:
      final tv = <double>[for (final p in s['vertices'] as List) for (final x in p as List) (x as num).toDouble()];
      final ti = <int>[for (final i in s['tm_indices'] as List) i as int];
      final tm = Trimesh(tv, ti);
      if (s['tm_scale'] != null) tm.setScale(v3(s['tm_scale'] as List));
      return tm;
    case 4:
      return Cylinder(radiusTop: d(s['radius_top'], 1), radiusBottom: d(s['radius_bottom'], 1), height: d(s['height'], 1),
          numSegments: (s['num_segments'] ?? 8) as int);
    case 8:
      final data = <List<double>>[for (final row in s['hf_data'] as List) [for (final h in row as List) (h as num).toDouble()]];
      return Heightfield(data, elementSize: (s['hf_element_size'] ?? 1) as int);
    default: throw 'shape type ${s['type']} not handled';
  }
}

void main(List<String> args) {
  final scene = jsonDecode(File(args[0]).readAsStringSync()) as Map;
  final desc = scene['desc'] as Map;
  final n = scene['n_bodies'] as int;
  final bodies = scene['bodies'] as Map;

  // broadphase / solver objects from the flattened descriptor
  Broadphase bp;
  switch ((desc['broadphase_kind'] ?? 0) as int) {
    case 1:
      final sap = SAPBroadphase(null);
      sap.axisIndex = AxisIndex.values[(desc['sap_axis'] ?? 0) as int];
      bp = sap;
      break;
    case 2:
      bp = GridBroadphase(v3(desc['grid_min']), v3(desc['grid_max']), desc['grid_nx'] as int, desc['grid_ny'] as int, desc['grid_nz'] as int);
      break;
    default: bp = NaiveBroadphase();
  }
  bp.useBoundingBoxes = ((desc['use_bounding_boxes'] ?? 0) as int) != 0;
  final gs = GSSolver(iterations: (desc['solver_iterations'] ?? 10) as int, tolerance: d(desc['solver_tolerance'], 1e-7));
  final kind = (desc['solver_kind'] ?? 0) as int;
  if (kind == 1 || kind == 3) throw 'the COLORED order has no reference counterpart: export the scene with the reference-order solver';
  final Solver solver = kind == 2 ? SplitSolver(gs) : gs;

  final world = World(
    gravity: desc['gravity'] == null ? null : v3(desc['gravity']),
    frictionGravity: ((desc['has_friction_gravity'] ?? 0) as int) != 0 ? v3(desc['friction_gravity']) : null,
    allowSleep: ((desc['allow_sleep'] ?? 0) as int) != 0,
    broadphase: bp, solver: solver,
    quatNormalizeFast: ((desc['quat_normalize_fast'] ?? 0) as int) != 0,
    quatNormalizeSkip: (desc['quat_normalize_skip'] ?? 0) as int,
  );
  final dcm = desc['default_contact_material'] as Map?;
  if (dcm != null) {
    final c = world.defaultContactMaterial;
    c.friction = d(dcm['friction'], c.friction); c.restitution = d(dcm['restitution'], c.restitution);
    c.contactEquationStiffness = d(dcm['contact_equation_stiffness'], c.contactEquationStiffness);
    c.contactEquationRelaxation = d(dcm['contact_equation_relaxation'], c.contactEquationRelaxation);
    c.frictionEquationStiffness = d(dcm['friction_equation_stiffness'], c.frictionEquationStiffness);
    c.frictionEquationRelaxation = d(dcm['friction_equation_relaxation'], c.frictionEquationRelaxation);
  }

  // materials and contact materials
  final mf = scene['material_friction'] as List?, mr = scene['material_restitution'] as List?;
  final materials = <Material>[
    for (var i = 0; i < (mf?.length ?? 0); i++) Material(friction: (mf![i] as num).toDouble(), restitution: (mr![i] as num).toDouble(), name: 'm$i')
  ];
  for (final cm in (scene['contact_materials'] as List)) {
    final m = cm as Map;
    world.addContactMaterial(ContactMaterial(materials[m['material_a'] as int], materials[m['material_b'] as int],
        friction: d(m['friction'], 0.3), restitution: d(m['restitution'], 0.3),
        contactEquationStiffness: d(m['contact_equation_stiffness'], 1e7), contactEquationRelaxation: d(m['contact_equation_relaxation'], 3),
        frictionEquationStiffness: d(m['friction_equation_stiffness'], 1e7), frictionEquationRelaxation: d(m['friction_equation_relaxation'], 3)));
  }

  // shapes are shared between bodies like in the reference's demos (examples/lib/examples/container.dart:105)
  final shapes = <Shape>[for (final s in scene['shapes'] as List) makeShape(s as Map)];
  List? arr(String k) => bodies[k] as List?;
  final bodyShapes = scene['body_shapes'] as Map?;
  for (var i = 0; i < n; i++) {
    final mass = arr('mass') == null ? 0.0 : (arr('mass')![i] as num).toDouble();
    final q = arr('quaternion')?[i] as List?;
    final typeCode = arr('type') == null ? -1 : arr('type')![i] as int;
    final matIdx = arr('material') == null ? -1 : arr('material')![i] as int;
    final shapeIdx = arr('shape') == null ? -1 : arr('shape')![i] as int;
    final b = Body(
      mass: mass,
      position: arr('position') == null ? null : v3(arr('position')![i] as List),
      velocity: arr('velocity') == null ? null : v3(arr('velocity')![i] as List),
      angularVelocity: arr('angular_velocity') == null ? null : v3(arr('angular_velocity')![i] as List),
      quaternion: q == null ? null : Quaternion((q[0] as num).toDouble(), (q[1] as num).toDouble(), (q[2] as num).toDouble(), (q[3] as num).toDouble()),
      type: typeCode < 0 ? null : BodyTypes.values[typeCode],
      material: matIdx < 0 ? null : materials[matIdx],
      linearDamping: arr('linear_damping') == null ? 0.01 : (arr('linear_damping')![i] as num).toDouble(),
      angularDamping: arr('angular_damping') == null ? 0.01 : (arr('angular_damping')![i] as num).toDouble(),
      allowSleep: arr('allow_sleep') == null ? true : (arr('allow_sleep')![i] as int) != 0,
      sleepSpeedLimit: arr('sleep_speed_limit') == null ? 0.1 : (arr('sleep_speed_limit')![i] as num).toDouble(),
      sleepTimeLimit: arr('sleep_time_limit') == null ? 1 : (arr('sleep_time_limit')![i] as num).toDouble(),
      fixedRotation: arr('fixed_rotation') == null ? false : (arr('fixed_rotation')![i] as int) != 0,
      linearFactor: arr('linear_factor') == null ? null : v3(arr('linear_factor')![i] as List),
      angularFactor: arr('angular_factor') == null ? null : v3(arr('angular_factor')![i] as List),
      collisionFilterGroup: arr('collision_filter_group') == null ? 1 : arr('collision_filter_group')![i] as int,
      collisionFilterMask: arr('collision_filter_mask') == null ? -1 : arr('collision_filter_mask')![i] as int,
      collisionResponse: arr('collision_response') == null ? true : (arr('collision_response')![i] as int) != 0,
      isTrigger: arr('is_trigger') == null ? false : (arr('is_trigger')![i] as int) != 0,
      shape: (shapeIdx < 0 || bodyShapes != null) ? null : shapes[shapeIdx],
    );
    if (bodyShapes != null) {  // Body.addShape(shape, offset, orientation) in table order (rigid_body.dart:348-377)
      final first = bodyShapes['first'] as List, ish = bodyShapes['shape'] as List;
      final off = bodyShapes['offset'] as List?, ori = bodyShapes['orientation'] as List?;
      for (var k = first[i] as int; k < (first[i + 1] as int); k++) {
        final o = ori == null ? null : ori[k] as List;
        b.addShape(shapes[ish[k] as int], off == null ? null : v3(off[k] as List),
            o == null ? null : Quaternion((o[0] as num).toDouble(), (o[1] as num).toDouble(), (o[2] as num).toDouble(), (o[3] as num).toDouble()));
      }
    }
    if (arr('force') != null) b.force.setFrom(v3(arr('force')![i] as List));
    if (arr('torque') != null) b.torque.setFrom(v3(arr('torque')![i] as List));
    if (arr('sleep_state') != null) b.sleepState = BodySleepStates.values[arr('sleep_state')![i] as int];
    world.addBody(b);
  }

  // constraints, in list order (their equations enter the solver in this order, world_class.dart:627-635)
  for (final cj in scene['constraints'] as List) {
    final c = cj as Map;
    final a = world.bodies[c['body_a'] as int], b = world.bodies[c['body_b'] as int];
    final maxForce = d(c['max_force'], 1e6);
    Vector3? opt(String k) => c[k] == null ? null : v3(c[k] as List);
    Constraint k;
    switch (c['type'] as int) {
      case 0: k = PointToPointConstraint(a, b, opt('pivot_a'), opt('pivot_b'), maxForce); break;
      case 1:
        final h = HingeConstraint(a, b, pivotA: opt('pivot_a'), pivotB: opt('pivot_b'), axisA: opt('axis_a'), axisB: opt('axis_b'), maxForce: maxForce);
        if (((c['motor_enabled'] ?? 0) as int) != 0) h.enableMotor();
        if (c['motor_target_velocity'] != null) h.setMotorSpeed(d(c['motor_target_velocity'], 0));
        if (c['motor_max_force'] != null && d(c['motor_max_force'], 0) > 0) h.setMotorMaxForce(d(c['motor_max_force'], maxForce));
        k = h;
        break;
      case 2: k = DistanceConstraint(a, b, c['distance'] == null || d(c['distance'], -1) < 0 ? null : d(c['distance'], 0), maxForce); break;
      case 3: k = LockConstraint(a, b, maxForce: maxForce); break;
      case 4:
        k = ConeTwistConstraint(a, b, pivotA: opt('pivot_a'), pivotB: opt('pivot_b'), axisA: opt('axis_a'), axisB: opt('axis_b'),
            angle: d(c['angle'], 0), twistAngle: d(c['twist_angle'], 0), maxForce: maxForce);
        break;
      default: throw 'constraint type ${c['type']} not handled';
    }
    if (c['collide_connected'] != null) k.collideConnected = (c['collide_connected'] as int) != 0;
    world.addConstraint(k);
  }

  // springs: the canonical postStep listener (examples/lib/examples/spring.dart:90,123)
  final springs = <Spring>[
    for (final sj in scene['springs'] as List)
      Spring(world.bodies[(sj as Map)['body_a'] as int], world.bodies[sj['body_b'] as int],
          restLength: d(sj['rest_length'], 1), stiffness: d(sj['stiffness'], 100), damping: d(sj['damping'], 1),
          localAnchorA: sj['local_anchor_a'] == null ? null : v3(sj['local_anchor_a'] as List),
          localAnchorB: sj['local_anchor_b'] == null ? null : v3(sj['local_anchor_b'] as List))
  ];
  if (springs.isNotEmpty) {
    world.addEventListener('postStep', (e) { for (final s in springs) { s.applyForce(); } });
  }

  // SPH subsystems (lib/objects/sph_system.dart; World.subsystems, world_class.dart:121,472-475)
  for (final sj in (scene['sph_systems'] ?? const []) as List) {
    final m = sj as Map;
    final sph = SPHSystem();
    sph.density = d(m['density'], 1); sph.smoothingRadius = d(m['smoothing_radius'], 1); sph.speedOfSound = d(m['speed_of_sound'], 1);
    sph.viscosity = d(m['viscosity'], 0.01); sph.eps = d(m['eps'], 0.00001);
    for (final p in m['particles'] as List) { sph.add(world.bodies[p as int]); }
    world.subsystems.add(sph);
  }

  final dt = (scene['dt'] as num).toDouble();
  final steps = scene['steps'] as int;
  final checkpoints = {for (final s in scene['checkpoints'] as List) s as int};
  final contactsPerStep = <int>[];
  final out = <String, dynamic>{};
  for (var step = 1; step <= steps; step++) {
    world.step(dt);
    contactsPerStep.add(world.contacts.length);
    if (checkpoints.contains(step)) {
      out['$step'] = {
        'position': [for (final b in world.bodies) [b.position.x, b.position.y, b.position.z]],
        'quaternion': [for (final b in world.bodies) [b.quaternion.x, b.quaternion.y, b.quaternion.z, b.quaternion.w]],
        'velocity': [for (final b in world.bodies) [b.velocity.x, b.velocity.y, b.velocity.z]],
        'angular_velocity': [for (final b in world.bodies) [b.angularVelocity.x, b.angularVelocity.y, b.angularVelocity.z]],
        'sleep_state': [for (final b in world.bodies) b.sleepState.index],
      };
    }
  }
  // doubles print with shortest round-trip digits, and every value above is an f32 widened to double: the JSON is lossless
  File(args[1]).writeAsStringSync(jsonEncode({'name': scene['name'], 'contacts_per_step': contactsPerStep, 'checkpoints': out}));
  stdout.writeln('${scene['name']}: $steps steps, ${contactsPerStep.last} contacts in the last step -> ${args[1]}');
}
