"""ctypes binding of the C ABI declared in ``include/cannon_cuda.h``.

``bind(path)`` returns a library object with typed prototypes.  The product (``libcannon_cuda.so``)
and the test-only CPU checker (``oracle/libcannon_oracle.so``) export the same symbols, so the same
binder serves both; the package itself only ever loads the CUDA library (see ``__init__.py``).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

c_i32, c_i64, c_f32, c_f64, c_u8 = C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_uint8
P = C.POINTER

OK, E_INVALID, E_CUDA, E_CAPACITY, E_UNSUPPORTED, E_NOGPU = 0, -1, -2, -3, -4, -5

SHAPE_SPHERE, SHAPE_PLANE, SHAPE_BOX, SHAPE_CONVEX, SHAPE_CYLINDER, SHAPE_HEIGHTFIELD = 0, 1, 2, 3, 4, 8
SHAPE_CAPSULE, SHAPE_CONE, SHAPE_SIZED_PLANE, SHAPE_PARTICLE, SHAPE_TRIMESH = 5, 6, 7, 9, 10
BODY_DYNAMIC, BODY_STATIC, BODY_KINEMATIC = 0, 1, 2
AWAKE, SLEEPY, SLEEPING = 0, 1, 2
BP_NAIVE, BP_SAP, BP_GRID = 0, 1, 2
SOLVER_REFERENCE_ORDER, SOLVER_COLORED, SOLVER_SPLIT, SOLVER_COLORED_F32 = 0, 1, 2, 3
CONSTRAINT_POINT_TO_POINT, CONSTRAINT_HINGE, CONSTRAINT_DISTANCE, CONSTRAINT_LOCK, CONSTRAINT_CONE_TWIST = 0, 1, 2, 3, 4


class ContactMaterialPOD(C.Structure):
    _fields_ = [
        ("material_a", c_i32), ("material_b", c_i32),
        ("friction", c_f64), ("restitution", c_f64),
        ("contact_equation_stiffness", c_f64), ("contact_equation_relaxation", c_f64),
        ("friction_equation_stiffness", c_f64), ("friction_equation_relaxation", c_f64),
    ]


class WorldDesc(C.Structure):
    _fields_ = [
        ("gravity", c_f32 * 3), ("friction_gravity", c_f32 * 3), ("has_friction_gravity", c_i32),
        ("allow_sleep", c_i32), ("quat_normalize_skip", c_i32), ("quat_normalize_fast", c_i32),
        ("solver_kind", c_i32), ("solver_iterations", c_i32), ("solver_tolerance", c_f64),
        ("broadphase_kind", c_i32), ("use_bounding_boxes", c_i32), ("sap_axis", c_i32),
        ("grid_nx", c_i32), ("grid_ny", c_i32), ("grid_nz", c_i32),
        ("grid_min", c_f32 * 3), ("grid_max", c_f32 * 3),
        ("default_contact_material", ContactMaterialPOD),
        ("n_worlds", c_i32), ("max_pairs", c_i32), ("max_contacts", c_i32),
    ]


class ShapeDesc(C.Structure):
    _fields_ = [
        ("type", c_i32), ("collision_response", c_i32),
        ("collision_filter_group", c_i32), ("collision_filter_mask", c_i32),
        ("radius", c_f64), ("half_extents", c_f32 * 3),
        ("radius_top", c_f64), ("radius_bottom", c_f64), ("height", c_f64), ("num_segments", c_i32),
        ("n_vertices", c_i32), ("vertices", P(c_f32)),
        ("n_faces", c_i32), ("face_offsets", P(c_i32)), ("face_indices", P(c_i32)),
        ("hf_nx", c_i32), ("hf_ny", c_i32), ("hf_data", P(c_f64)), ("hf_element_size", c_i32),
        ("convex_has_axes", c_i32),
        ("n_triangles", c_i32), ("tm_indices", P(c_i32)), ("tm_scale", c_f32 * 3), ("material", c_i32),
    ]


class SphDesc(C.Structure):
    _fields_ = [("n_particles", c_i32), ("particles", P(c_i32)), ("density", c_f64), ("smoothing_radius", c_f64), ("speed_of_sound", c_f64),
                ("viscosity", c_f64), ("eps", c_f64)]


# (field, ctypes element type, numpy dtype, components per body)
BODY_FIELDS = [
    ("position", c_f32, np.float32, 3), ("quaternion", c_f32, np.float32, 4),
    ("velocity", c_f32, np.float32, 3), ("angular_velocity", c_f32, np.float32, 3),
    ("force", c_f32, np.float32, 3), ("torque", c_f32, np.float32, 3),
    ("mass", c_f64, np.float64, 1), ("type", c_i32, np.int32, 1), ("sleep_state", c_i32, np.int32, 1),
    ("time_last_sleepy", c_f64, np.float64, 1), ("allow_sleep", c_u8, np.uint8, 1),
    ("sleep_speed_limit", c_f64, np.float64, 1), ("sleep_time_limit", c_f64, np.float64, 1),
    ("linear_damping", c_f64, np.float64, 1), ("angular_damping", c_f64, np.float64, 1),
    ("linear_factor", c_f32, np.float32, 3), ("angular_factor", c_f32, np.float32, 3),
    ("fixed_rotation", c_u8, np.uint8, 1),
    ("collision_filter_group", c_i32, np.int32, 1), ("collision_filter_mask", c_i32, np.int32, 1),
    ("collision_response", c_u8, np.uint8, 1), ("is_trigger", c_u8, np.uint8, 1),
    ("material", c_i32, np.int32, 1), ("shape", c_i32, np.int32, 1), ("world_id", c_i32, np.int32, 1),
    ("inv_mass", c_f64, np.float64, 1), ("inv_inertia", c_f32, np.float32, 3),
    ("inv_inertia_world", c_f32, np.float32, 9), ("bounding_radius", c_f64, np.float64, 1),
    ("aabb", c_f32, np.float32, 6),
]
DERIVED_BODY_FIELDS = ("inv_mass", "inv_inertia", "inv_inertia_world", "bounding_radius", "aabb")


class BodiesSoA(C.Structure):
    _fields_ = [("n", c_i32)] + [(name, P(ct)) for name, ct, _, _ in BODY_FIELDS]


class ConstraintDesc(C.Structure):
    _fields_ = [
        ("type", c_i32), ("body_a", c_i32), ("body_b", c_i32),
        ("pivot_a", c_f32 * 3), ("pivot_b", c_f32 * 3), ("axis_a", c_f32 * 3), ("axis_b", c_f32 * 3),
        ("max_force", c_f64), ("collide_connected", c_i32), ("motor_enabled", c_i32),
        ("motor_target_velocity", c_f64), ("motor_max_force", c_f64),
        ("distance", c_f64), ("angle", c_f64), ("twist_angle", c_f64),
        ("has_ctor_pose", c_i32), ("ctor_pos_a", c_f32 * 3), ("ctor_quat_a", c_f32 * 4), ("ctor_pos_b", c_f32 * 3), ("ctor_quat_b", c_f32 * 4),
    ]


class SpringDesc(C.Structure):
    _fields_ = [
        ("body_a", c_i32), ("body_b", c_i32), ("rest_length", c_f64), ("stiffness", c_f64), ("damping", c_f64),
        ("local_anchor_a", c_f32 * 3), ("local_anchor_b", c_f32 * 3),
    ]


class ContactsSoA(C.Structure):
    _fields_ = [
        ("capacity", c_i32), ("body_i", P(c_i32)), ("body_j", P(c_i32)),
        ("ri", P(c_f32)), ("rj", P(c_f32)), ("ni", P(c_f32)),
        ("restitution", P(c_f64)), ("friction", P(c_f64)), ("enabled", P(c_u8)), ("multiplier", P(c_f64)),
    ]


class Profile(C.Structure):
    _fields_ = [
        ("solve", c_f64), ("make_contact_constraints", c_f64), ("broadphase", c_f64),
        ("integrate", c_f64), ("narrowphase", c_f64),
        ("n_pairs", c_i64), ("n_contacts", c_i64), ("n_rows", c_i64), ("n_levels", c_i64),
        ("iterations_done", c_i64), ("steps", c_i64), ("contact_iters_total", c_i64),
        ("step_call_ms", c_f64), ("schedule_ms", c_f64), ("gs_ms", c_f64), ("kernel_launches", c_i64), ("n_tasks", c_i64), ("n_islands", c_i64), ("n_tasks_by_type", c_i64 * 8),
        ("sum_steps", c_i64), ("sum_step_ms", c_f64), ("sum_broadphase", c_f64), ("sum_narrowphase", c_f64), ("sum_solve", c_f64),
        ("sum_integrate", c_f64), ("sum_schedule", c_f64), ("sum_gs", c_f64),
    ]


RAY_CLOSEST, RAY_ANY, RAY_ALL = 1, 2, 4


class RayOptions(C.Structure):
    _fields_ = [("mode", c_i32), ("skip_backfaces", c_i32), ("collision_filter_mask", c_i32), ("collision_filter_group", c_i32),
                ("check_collision_response", c_i32)]


class RayHitsSoA(C.Structure):
    _fields_ = [("capacity", c_i32), ("ray", P(c_i32)), ("body", P(c_i32)), ("hit_face_index", P(c_i32)), ("distance", P(c_f64)),
                ("hit_point_world", P(c_f32)), ("hit_normal_world", P(c_f32)), ("shape_ordinal", P(c_i32))]


BATCH_MAX_GPUS = 16


class BatchStats(C.Structure):
    _fields_ = [
        ("n_gpus", c_i32), ("n_worlds", c_i32), ("bodies_per_world", c_i32), ("pad0", c_i32),
        ("n_pairs", c_i64), ("n_contacts", c_i64), ("n_rows", c_i64), ("iterations_done", c_i64), ("steps", c_i64),
        ("contact_iters_total", c_i64), ("step_call_ms_max", c_f64), ("gpu_step_call_ms", c_f64 * BATCH_MAX_GPUS),
        ("gpu_worlds", c_i32 * BATCH_MAX_GPUS),
    ]


VP = C.c_void_p

# every symbol include/cannon_cuda.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "cannon_version": (c_i32, []),
    "cannon_backend": (C.c_char_p, []),
    "cannon_ctx_create": (c_i32, [c_i32, P(VP)]),
    "cannon_ctx_destroy": (None, [VP]),
    "cannon_last_error": (C.c_char_p, [VP]),
    "cannon_world_desc_default": (None, [P(WorldDesc)]),
    "cannon_shape_desc_default": (None, [P(ShapeDesc)]),
    "cannon_world_create": (c_i32, [VP, P(WorldDesc), P(VP)]),
    "cannon_world_destroy": (None, [VP]),
    "cannon_world_set_materials": (c_i32, [VP, c_i32, P(c_f64), P(c_f64), c_i32, P(ContactMaterialPOD)]),
    "cannon_world_set_shapes": (c_i32, [VP, c_i32, P(ShapeDesc)]),
    "cannon_sph_desc_default": (None, [P(SphDesc)]),
    "cannon_world_set_sph_systems": (c_i32, [VP, c_i32, P(SphDesc)]),
    "cannon_batch_set_body_shapes": (c_i32, [VP, c_i32, P(c_i32), P(c_i32), P(c_f32), P(c_f32)]),
    "cannon_world_set_body_shapes": (c_i32, [VP, c_i32, P(c_i32), P(c_i32), P(c_f32), P(c_f32)]),
    "cannon_world_set_bodies": (c_i32, [VP, P(BodiesSoA)]),
    "cannon_world_get_bodies": (c_i32, [VP, P(BodiesSoA)]),
    "cannon_world_set_constraints": (c_i32, [VP, c_i32, P(ConstraintDesc)]),
    "cannon_world_set_springs": (c_i32, [VP, c_i32, P(SpringDesc)]),
    "cannon_world_set_time": (c_i32, [VP, c_f64]),
    "cannon_world_get_time": (c_i32, [VP, P(c_f64), P(c_i64)]),
    "cannon_world_set_dt": (c_i32, [VP, c_f64]),
    "cannon_apply_gravity": (c_i32, [VP]),
    "cannon_broadphase_pairs": (c_i32, [VP, P(c_i32), P(c_i32), c_i32, P(c_i32)]),
    "cannon_narrowphase_contacts": (c_i32, [VP, P(c_i32), P(c_i32), c_i32, P(ContactsSoA), P(c_i32), P(c_i32)]),
    "cannon_solver_solve": (c_i32, [VP, c_f64, P(c_i32)]),
    "cannon_integrate": (c_i32, [VP, c_f64]),
    "cannon_world_step": (c_i32, [VP, c_f64, c_i32]),
    "cannon_world_profile": (c_i32, [VP, P(Profile)]),
    "cannon_world_step_profiled": (c_i32, [VP, c_f64, c_i32]),
    "cannon_world_step_async": (c_i32, [VP, c_f64, c_i32]),
    "cannon_ctx_sync": (c_i32, [VP]),
    "cannon_world_get_contacts": (c_i32, [VP, P(ContactsSoA), P(c_i32)]),
    "cannon_world_enable_contact_events": (c_i32, [VP, c_i32]),
    "cannon_world_get_contact_events": (c_i32, [VP, c_i32, P(c_i32), P(c_i32), P(c_i32), P(c_i32), P(c_i32), P(c_i32)]),
    "cannon_world_get_rows": (c_i32, [VP, c_i32, P(c_i32), P(c_i32), P(c_i32), P(c_f64), P(c_f64), P(c_f64), P(c_i32)]),
    "cannon_world_update_bodies": (c_i32, [VP, c_i32, c_i32, P(c_f32), P(c_f32), P(c_f32), P(c_f32), P(c_f32), P(c_f32)]),
    "cannon_world_set_inv_inertia": (c_i32, [VP, c_i32, c_i32, P(c_f32)]),
    "cannon_world_set_stepnumber": (c_i32, [VP, c_i64]),
    "cannon_world_update_sleep_states": (c_i32, [VP, c_i32, c_i32, P(c_i32)]),
    "cannon_world_set_hinge_motor": (c_i32, [VP, c_i32, c_i32, c_f64, c_f64]),
    "cannon_ray_options_default": (None, [P(RayOptions)]),
    "cannon_world_raycast": (c_i32, [VP, c_i32, P(c_f32), P(c_f32), P(RayOptions), P(c_u8), P(RayHitsSoA), P(c_i32)]),
    "cannon_world_aabb_query": (c_i32, [VP, P(c_f32), P(c_f32), P(c_i32), c_i32, P(c_i32)]),
    "cannon_batch_create": (c_i32, [P(c_i32), c_i32, P(WorldDesc), c_i32, c_i32, P(VP)]),
    "cannon_batch_destroy": (None, [VP]),
    "cannon_batch_last_error": (C.c_char_p, [VP]),
    "cannon_batch_set_materials": (c_i32, [VP, c_i32, P(c_f64), P(c_f64), c_i32, P(ContactMaterialPOD)]),
    "cannon_batch_set_shapes": (c_i32, [VP, c_i32, P(ShapeDesc)]),
    "cannon_batch_set_bodies": (c_i32, [VP, P(BodiesSoA)]),
    "cannon_batch_set_constraints": (c_i32, [VP, c_i32, P(ConstraintDesc)]),
    "cannon_batch_step": (c_i32, [VP, c_f64, c_i32]),
    "cannon_batch_stats": (c_i32, [VP, P(BatchStats)]),
    "cannon_batch_get_bodies": (c_i32, [VP, P(BodiesSoA)]),
    "cannon_batch_shard": (c_i32, [VP, c_i32, P(c_i32), P(c_i32), P(VP)]),
}


class CannonError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cannon error {code}: {msg}")
        self.code = code


def bind(path: str) -> C.CDLL:
    """Load a library implementing include/cannon_cuda.h and attach the prototypes.

    Raises OSError if the file is missing or a declared symbol is not exported.
    """
    if not os.path.exists(path):
        raise OSError(f"{path} not found")
    lib = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError -> missing symbol
        fn.restype = res
        fn.argtypes = args
    lib._path = path
    return lib


def ptr(arr, ctype):
    """Pointer to a C-contiguous numpy array (or NULL for None)."""
    if arr is None:
        return None
    assert arr.flags["C_CONTIGUOUS"]
    return arr.ctypes.data_as(P(ctype))
