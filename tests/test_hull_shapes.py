"""Cone / Capsule / CapsuleLathe / SizedPlane / LatheShape.

The reference builds these as ConvexPolyhedron subclasses (lib/rigid_body_shapes/{cone,capsule,capsule_lathe,lathe,
sized_plane}.dart); their ShapeType only selects the resolver and which shape a resolver sees first
(lib/world/narrow_phase.dart:116-238,336-489,706-710). CPU tests pin the host-side constructors (vertex / face counts
and known vertices derived from the source) and the oracle's dispatch; the GPU tests are bit-exact parity."""
import math

import numpy as np
import pytest

import parity
from cannon_physics_b200 import _ffi as F
from cannon_physics_b200 import api, scenes
from cannon_physics_b200.engine import DeviceWorld, SceneSpec

IDENT = np.array([0, 0, 0, 1], np.float32)


def f32(x):
    return np.float32(x)


# ---------------------------------------------------------------------------------------------------------------------
# constructors (source-derived known answers)
# ---------------------------------------------------------------------------------------------------------------------
def test_cone_constructor_follows_cone_dart():
    c = api.Cone(radius=0.5, height=2.0, numSegments=6)
    assert c.type == 6 and c.uniqueAxes is not None  # cone.dart:58 passes axes => face normals are SAT axes
    assert c.vertices.shape == (7, 3)  # apex + 6 base vertices
    np.testing.assert_array_equal(c.vertices[0], [0, 1, 0])
    np.testing.assert_array_equal(c.vertices[1], [f32(-0.0), -1, 0.5])
    theta = ((2 * math.pi) / 6) * 2
    np.testing.assert_array_equal(c.vertices[3], np.array([-0.5 * math.sin(theta), -1.0, 0.5 * math.cos(theta)], np.float32))
    assert c.faces[0] == [0, 2, 1] and c.faces[4] == [0, 6, 5] and c.faces[5] == [0, 1, 6]  # cone.dart:53,56
    assert c.faces[6] == [1, 2, 3, 4, 5, 6]  # the base
    with pytest.raises(ValueError):
        api.Cone(radius=-1)


def test_capsule_constructor_follows_capsule_dart():
    ns, nh = 8, 4
    c = api.Capsule(radiusTop=0.5, radiusBottom=0.5, height=1.0, numSegments=ns, numHeightSegments=nh)
    assert c.type == 5 and c.uniqueAxes is None  # init(vertices, faces): no axes (capsule.dart:168)
    assert c.vertices.shape == (2 * ns + 2 * (ns + 1) * (nh + 1), 3)
    assert len(c.faces) == ns + 4 * ns * nh
    assert c.faces[0] == [0, 1, 3, 2] and c.faces[ns - 1] == [2 * (ns - 1), 2 * (ns - 1) + 1, 1, 0]
    start1 = 2 * ns
    start2 = start1 + (ns + 1) * (nh + 1)
    assert c.faces[ns] == [1 + start1, 0 + start1, (ns + 1) + 1 + start1]        # [a1, b1, d1] of (iy=0, ix=0)
    assert c.faces[ns + 2] == [1 + start2, 0 + start2, (ns + 1) + 1 + start2]    # [a2, b2, d2]
    # first cap vertex: iy = 0, ix = 0 => ut = 1, v = 0 (capsule.dart:106-112)
    st, ct = math.sin(math.pi / 2), math.cos(math.pi / 2)
    want = np.array([-0.5 * math.cos(math.pi + 2 * math.pi) * st, 0.5 - 0.5 * ct, 0.5 * math.sin(math.pi + 2 * math.pi) * st], np.float32)
    np.testing.assert_array_equal(c.vertices[start1], want)
    # the poles are the farthest points: bounding radius = height/2 + radiusTop
    assert abs(float(np.sqrt((c.vertices.astype(np.float64) ** 2).sum(1)).max()) - 1.0) < 1e-6


def test_sized_plane_and_lathe_constructors():
    p = api.SizedPlane(4.0, 2.0)
    assert p.type == 7 and p.uniqueAxes is None
    np.testing.assert_array_equal(p.vertices, np.array([[-2, 0, -1], [2, 0, -1], [2, 0, 1], [-2, 0, 1]], np.float32))
    assert p.faces == [[3, 2, 1, 0]]
    lathe = api.LatheShape([(0, -0.5), (0.5, 0), (0, 0.5)], numSegments=6)
    assert lathe.type == 3 and lathe.vertices.shape == (7 * 3, 3) and len(lathe.faces) == 6 * 2 * 2
    np.testing.assert_array_equal(lathe.vertices[0], [0, 0.5, 0])       # j runs from the last profile point down (lathe.dart:35)
    np.testing.assert_array_equal(lathe.vertices[1], [0, 0, 0.5])
    assert lathe.faces[0] == [1, 4, 2] and lathe.faces[1] == [5, 2, 4]  # j = 1, i = 0: a=1 b=4 c=5 d=2
    cl = api.CapsuleLathe(radiusTop=0.5, radiusBottom=0.5, height=1.0, numSegments=8, numHeightSegments=4)
    assert cl.type == 5 and len(cl.points) == 2 * (4 - 1) + 4  # capsule_lathe.dart:27-47
    np.testing.assert_array_equal(cl.points[0], [0, 1.0])
    np.testing.assert_array_equal(cl.points[-1], [0, -1.0])


# ---------------------------------------------------------------------------------------------------------------------
# scenes
# ---------------------------------------------------------------------------------------------------------------------
def _mixed_spec(ground="plane", solver=None, seed=3):
    """One of every hull shape dropped on a ground (plane | sized plane | heightfield) in a loose pile."""
    hulls = [api.Box((0.4, 0.3, 0.5)), api.Cylinder(0.4, 0.5, 0.9, 8), api.Cone(0.5, 1.0, 8), api.Capsule(0.3, 0.3, 0.6, 6, 2),
             api.CapsuleLathe(0.3, 0.3, 0.5, 6, 3), api.SizedPlane(1.2, 0.9), api.Sphere(0.4),
             api.ConvexPolyhedron([(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)], [(0, 3, 2), (0, 1, 3), (0, 2, 1), (1, 2, 3)]),
             api.LatheShape([(0, -0.4), (0.45, 0), (0, 0.4)], numSegments=5)]
    shapes = []
    if ground == "plane":
        shapes.append(dict(type=F.SHAPE_PLANE))
        gq, gp = scenes.GROUND_QUAT, (0, 0, 0)
    elif ground == "sized":
        shapes.append(api.SizedPlane(30.0, 30.0)._desc())
        gq, gp = IDENT, (0, 0, 0)
    else:
        rng = np.random.default_rng(7)
        hf = 0.15 * rng.random((14, 14))
        shapes.append(dict(type=F.SHAPE_HEIGHTFIELD, hf_data=hf, hf_element_size=1))
        gq, gp = scenes.GROUND_QUAT, (-6.5, 0, 6.5)
    shapes += [h._desc() for h in hulls]
    rng = np.random.default_rng(seed)
    n = 1 + 3 * len(hulls)
    pos = np.zeros((n, 3), np.float32)
    quat = np.tile(IDENT, (n, 1))
    shape = np.zeros(n, np.int32)
    mass = np.ones(n)
    pos[0], quat[0], mass[0] = gp, gq, 0.0
    k = 1
    for layer in range(3):
        order = rng.permutation(len(hulls))
        for j, h in enumerate(order):
            pos[k] = (1.1 * (j % 3) - 1.1 + 0.1 * rng.random(), 1.0 + 1.3 * layer + 0.05 * j, 1.1 * (j // 3) - 1.1 + 0.1 * rng.random())
            q = rng.normal(size=4)
            quat[k] = (q / np.linalg.norm(q)).astype(np.float32)
            shape[k] = 1 + h
            k += 1
    desc = dict(gravity=(0, -10, 0))
    if solver is not None:
        desc["solver_kind"] = solver
    return SceneSpec(desc=desc, shapes=shapes, bodies=dict(position=pos, quaternion=quat, mass=mass, shape=shape), n_bodies=n, name=f"hull shapes on {ground}")


def _run(world, steps, dt=1 / 60):
    for _ in range(steps):
        world.step(dt)
    return world.get_bodies(("position", "velocity"))


def test_oracle_new_hull_types_rest_on_the_ground(oracle_lib):
    ref = DeviceWorld(oracle_lib, _mixed_spec("plane"))
    out = _run(ref, 240)
    assert np.isfinite(out["position"]).all()
    assert (out["position"][1:, 1] > -0.05).all()  # nothing fell through the plane
    assert len(ref.get_contacts()["body_i"]) > 20


def test_oracle_convex_never_meets_a_sized_plane_but_the_others_do(oracle_lib):
    # narrow_phase.dart:448: the key "convexSizedPlane" can never equal a lower-cased name, so getCollisionType
    # (:476-479) finds no resolver for a plain ConvexPolyhedron against a SizedPlane; box / cone / sphere do collide
    tetra = api.ConvexPolyhedron([(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)], [(0, 3, 2), (0, 1, 3), (0, 2, 1), (1, 2, 3)])
    shapes = [api.SizedPlane(20.0, 20.0)._desc(), tetra._desc(), api.Box((0.3, 0.3, 0.3))._desc(), api.Cone(0.4, 0.8, 8)._desc(), api.Sphere(0.3)._desc()]
    pos = np.array([[0, 0, 0], [-3, 1, 0], [-1, 1, 0], [1, 1, 0], [3, 1, 0]], np.float32)
    spec = SceneSpec(desc=dict(gravity=(0, -10, 0)), shapes=shapes,
                     bodies=dict(position=pos, quaternion=np.tile(IDENT, (5, 1)), mass=np.array([0, 1, 1, 1, 1.0]), shape=np.arange(5, dtype=np.int32)), n_bodies=5)
    ref = DeviceWorld(oracle_lib, spec)
    out = _run(ref, 120)
    assert out["position"][1, 1] < -5.0                  # the tetrahedron fell through
    assert (out["position"][2:, 1] > -0.1).all()         # the others rest on the quad


def test_oracle_rays_ignore_the_new_types(oracle_lib):
    # ray_class.dart:101-123 has no handler for cone / capsule / sizedPlane: a ray through them reports the box behind
    shapes = [api.Cone(1.0, 2.0, 8)._desc(), api.Capsule(0.5, 0.5, 1.0, 6, 2)._desc(), api.Box((0.5, 0.5, 0.5))._desc()]
    pos = np.array([[0, 0, 0], [2, 0, 0], [4, 0, 0]], np.float32)
    spec = SceneSpec(desc=dict(gravity=(0, 0, 0)), shapes=shapes,
                     bodies=dict(position=pos, quaternion=np.tile(IDENT, (3, 1)), mass=np.zeros(3), shape=np.arange(3, dtype=np.int32)), n_bodies=3)
    ref = DeviceWorld(oracle_lib, spec)
    ref.step(1 / 60)
    hits = ref.raycast(np.array([[-5, 0, 0]], np.float32), np.array([[10, 0, 0]], np.float32), mode=F.RAY_CLOSEST)
    assert hits["has_hit"][0] == 1 and hits["body"][0] == 2
    assert abs(hits["distance"][0] - 8.5) < 1e-5


# ---------------------------------------------------------------------------------------------------------------------
# GPU parity (bit-exact)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("ground", ["plane", "sized", "heightfield"])
def test_hull_shapes_staged_parity(cuda_lib, oracle_lib, ground):
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, _mixed_spec(ground))
    seen = 0
    for s in range(150):
        seen = max(seen, parity.staged_step(dev, ref, 1 / 60, f"{ground} step {s}")[1])
    assert seen > 25


@pytest.mark.gpu
@pytest.mark.parametrize("solver", [F.SOLVER_REFERENCE_ORDER, F.SOLVER_COLORED])
def test_hull_shapes_fused_parity(cuda_lib, oracle_lib, solver):
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, _mixed_spec("heightfield", solver=solver, seed=5))
    for s in range(0, 240, 40):
        dev.step(1 / 60, 40)
        ref.step(1 / 60, 40)
        parity.assert_same_state(dev, ref, f"fused step {s + 40}")
    assert len(dev.get_contacts()["body_i"]) == len(ref.get_contacts()["body_i"]) > 20


@pytest.mark.gpu
def test_rays_ignore_the_new_types_on_the_device(cuda_lib, oracle_lib):
    spec = _mixed_spec("plane")
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, spec)
    for w in (dev, ref):
        w.step(1 / 60, 90)
    rng = np.random.default_rng(11)
    frm = np.column_stack([rng.uniform(-3, 3, 64), np.full(64, 6.0), rng.uniform(-3, 3, 64)]).astype(np.float32)
    to = frm + np.array([0, -8, 0], np.float32)
    for mode in (F.RAY_CLOSEST, F.RAY_ANY):
        a, b = dev.raycast(frm, to, mode=mode), ref.raycast(frm, to, mode=mode)
        for k in ("has_hit", "body", "distance", "hit_point_world", "hit_normal_world"):
            assert np.array_equal(a[k], b[k]), (mode, k)
