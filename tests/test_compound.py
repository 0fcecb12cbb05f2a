"""Compound bodies: Body.addShape(shape, offset, orientation), lib/objects/rigid_body.dart:348-377.

What must follow the reference: Body.updateAABB / updateBoundingRadius / updateMassProperties over all shapes
(rigid_body.dart:395-447,587-609), the shape-pair loops of Narrowphase.getContacts with the shapes' world poses
(narrow_phase.dart:669-721), ri / rj relative to the BODY position, Ray.intersectBody over the shapes
(ray_class.dart:226-243). CPU tests are source-derived known answers on the oracle; GPU tests are bit-exact parity."""
import math

import numpy as np
import pytest

import parity
from cannon_physics_b200 import _ffi as F
from cannon_physics_b200 import api, scenes
from cannon_physics_b200.engine import DeviceWorld, SceneSpec

IDENT = np.array([0, 0, 0, 1], np.float32)


def _spec(shapes, bodies, **desc):
    """bodies: list of dict(pos, mass, quat=None, inst=[(shape index, offset, orientation), ...])"""
    n = len(bodies)
    first, shape, off, ori = [0], [], [], []
    for b in bodies:
        for (s, o, q) in b["inst"]:
            shape.append(s)
            off.append((0, 0, 0) if o is None else o)
            ori.append(IDENT if q is None else q)
        first.append(len(shape))
    arrays = dict(position=np.array([b["pos"] for b in bodies], np.float32), quaternion=np.array([b.get("quat", IDENT) for b in bodies], np.float32),
                  mass=np.array([b["mass"] for b in bodies], np.float64))
    return SceneSpec(desc=dict(dict(gravity=(0, -10, 0)), **desc), shapes=[s._desc() if hasattr(s, "_desc") else s for s in shapes], bodies=arrays, n_bodies=n,
                     body_shapes=dict(first=np.array(first, np.int32), shape=np.array(shape, np.int32), offset=np.array(off, np.float32), orientation=np.array(ori, np.float32)))


def _contacts(world):
    world.set_dt(1 / 60)
    p = world.broadphase_pairs()
    return p, world.narrowphase_contacts(*p)


def test_dumbbell_on_the_plane_two_contacts_relative_to_the_body(oracle_lib):
    # two radius-0.5 spheres at x = -1 / +1 of a body whose origin is 0.4 above the ground plane
    spec = _spec([dict(type=F.SHAPE_PLANE), api.Sphere(0.5)],
                 [dict(pos=(0, 0, 0), mass=0, quat=scenes.GROUND_QUAT, inst=[(0, None, None)]),
                  dict(pos=(0, 0.4, 0), mass=2, inst=[(1, (-1, 0, 0), None), (1, (1, 0, 0), None)])])
    w = DeviceWorld(oracle_lib, spec)
    (p1, p2), c = _contacts(w)
    assert len(p1) == 1 and len(c["body_i"]) == 2 and c["per_pair_count"].tolist() == [2]
    assert set(zip(c["body_i"].tolist(), c["body_j"].tolist())) == {(1, 0)}
    # ri = radius * (-n) + xi - body.position: (-1 | +1, -0.5, 0); spherePlane's rj is the sphere centre projected on the plane
    np.testing.assert_allclose(c["ri"], [[-1, -0.5, 0], [1, -0.5, 0]], atol=1e-6)
    np.testing.assert_allclose(c["rj"], [[-1, 0, 0], [1, 0, 0]], atol=1e-6)
    b = w.get_bodies(("bounding_radius", "aabb", "inv_inertia"))
    assert abs(b["bounding_radius"][1] - 1.5) < 1e-12  # offset.length + r, rigid_body.dart:404-409
    np.testing.assert_allclose(b["aabb"][1], [-1.5, -0.1, -0.5, 1.5, 0.9, 0.5], atol=1e-6)
    # updateMassProperties: box inertia of the AABB half extents (1.5, 0.5, 0.5), mass 2
    want = 1.0 / np.array([2 / 12 * (1 + 1), 2 / 12 * (9 + 1), 2 / 12 * (1 + 9)])
    np.testing.assert_allclose(b["inv_inertia"][1], want, rtol=1e-6)


def test_shape_orientation_and_offset_reach_the_resolver(oracle_lib):
    # a thin box shape (half extents 1, 0.1, 0.1) rotated 90 degrees about z inside the body and lifted by 1: it stands
    # upright, its lower end at body.y - 0 -> touches the plane when the body origin is at y = 0 ... place it at 0.05 below
    qz = np.array([0, 0, math.sin(math.pi / 4), math.cos(math.pi / 4)], np.float32)
    spec = _spec([dict(type=F.SHAPE_PLANE), api.Box((1.0, 0.1, 0.1))],
                 [dict(pos=(0, 0, 0), mass=0, quat=scenes.GROUND_QUAT, inst=[(0, None, None)]),
                  dict(pos=(0, -0.05, 0), mass=1, inst=[(1, (0, 1, 0), qz)])])
    w = DeviceWorld(oracle_lib, spec)
    (_, _), c = _contacts(w)
    assert len(c["body_i"]) == 4                      # the four lower corners of the upright box
    assert np.allclose(c["ni"][:, 1], 1, atol=1e-6) or np.allclose(c["ni"][:, 1], -1, atol=1e-6)
    # corners relative to the BODY origin: y = 1 - 1 = 0 (offset + rotated half extent), x/z = +-0.1
    hull_side = c["rj"] if c["body_j"][0] == 1 else c["ri"]
    np.testing.assert_allclose(np.sort(np.abs(hull_side[:, [0, 2]]).ravel()), [0.1] * 8, atol=1e-6)
    np.testing.assert_allclose(hull_side[:, 1], 0.0, atol=1e-6)


def test_api_add_shape_builds_the_table_and_rays_see_every_shape(oracle_lib):
    world = api.World(gravity=(0, 0, 0), _lib=oracle_lib)
    body = api.Body(mass=1, position=(0, 0, 0))
    body.addShape(api.Sphere(0.5), offset=(-2, 0, 0))
    body.addShape(api.Box((0.5, 0.5, 0.5)), offset=(2, 0, 0))
    world.addBody(body)
    world.step(1 / 60)
    res = api.RaycastResult()
    assert world.raycastClosest((-2, 5, 0), (-2, -5, 0), result=res) and abs(res.distance - 4.5) < 1e-5
    assert res.shape is body.shapes[0]  # RaycastResult.shape: the shape that was hit, not the body's first
    assert world.raycastClosest((2, 5, 0), (2, -5, 0), result=res) and abs(res.distance - 4.5) < 1e-5
    assert res.shape is body.shapes[1]
    assert not world.raycastClosest((0, 5, 0), (0, -5, 0), result=res)  # between the two shapes
    # Body.removeShape (rigid_body.dart:371-391): the sphere goes, the box stays where its offset put it
    body.removeShape(body.shapes[0])
    world.step(1 / 60)
    assert not world.raycastClosest((-2, 5, 0), (-2, -5, 0), result=res)
    assert world.raycastClosest((2, 5, 0), (2, -5, 0), result=res) and res.shape is body.shapes[0]


def _table(k):
    """A small table: top plate + four legs (boxes), plus a rotated cylinder rail; all offsets / orientations non-trivial."""
    s = math.sin(math.pi / 4)
    legs = [(1, (sx * 0.5, -0.4, sz * 0.3), None) for sx in (-1, 1) for sz in (-1, 1)]
    return [(0, (0, 0, 0), None)] + legs + [(2, (0, 0.25, 0), (0, 0, s, s))]


def _pile_spec(ground="plane", solver=None, seed=1, n_obj=10):
    rng = np.random.default_rng(seed)
    shapes = [api.Box((0.7, 0.08, 0.45)), api.Box((0.08, 0.35, 0.08)), api.Cylinder(0.1, 0.1, 1.0, 8), api.Sphere(0.3), api.Box((0.3, 0.3, 0.3)), api.Cone(0.3, 0.6, 8)]
    shapes = [s._desc() for s in shapes]
    if ground == "plane":
        shapes.append(dict(type=F.SHAPE_PLANE))
        ground_body = dict(pos=(0, 0, 0), mass=0, quat=scenes.GROUND_QUAT, inst=[(len(shapes) - 1, None, None)])
    else:
        shapes.append(dict(type=F.SHAPE_HEIGHTFIELD, hf_data=0.2 * np.random.default_rng(3).random((14, 14)), hf_element_size=1))
        # a heightfield that is itself an offset, rotated shape of its body
        ground_body = dict(pos=(0, 0, 0), mass=0, inst=[(len(shapes) - 1, (-6.5, 0, 6.5), scenes.GROUND_QUAT)])
    bodies = [ground_body]
    for k in range(n_obj):
        q = rng.normal(size=4)
        q = (q / np.linalg.norm(q)).astype(np.float32)
        pos = (1.6 * (k % 3) - 1.6 + 0.1 * rng.random(), 1.0 + 0.9 * (k // 3) + 0.3 * rng.random(), 1.6 * ((k // 3) % 2) - 0.8 + 0.1 * rng.random())
        kind = k % 4
        if kind == 0:
            inst = _table(k)
        elif kind == 1:   # dumbbell: two spheres and a bar
            inst = [(3, (-0.5, 0, 0), None), (3, (0.5, 0, 0), None), (2, None, (0, 0, math.sin(math.pi / 4), math.cos(math.pi / 4)))]
        elif kind == 2:   # plain single shape through the table
            inst = [(4, None, None)]
        else:             # L shape with a cone on top
            inst = [(4, (0, 0, 0), None), (4, (0.6, 0, 0), None), (5, (0, 0.6, 0), None)]
        bodies.append(dict(pos=pos, mass=1.0 + 0.5 * kind, quat=q, inst=inst))
    bodies.append(dict(pos=(5, 5, 5), mass=1, inst=[]))  # a body without shapes
    desc = {}
    if solver is not None:
        desc["solver_kind"] = solver
    spec = _spec(shapes, bodies, **desc)
    spec.name = f"compound pile on {ground}"
    return spec


def test_oracle_compound_pile_settles(oracle_lib):
    w = DeviceWorld(oracle_lib, _pile_spec("plane"))
    seen = 0
    for _ in range(200):
        w.step(1 / 60)
        seen = max(seen, len(w.get_contacts()["body_i"]))
    out = w.get_bodies(("position", "velocity"))
    assert np.isfinite(out["position"]).all() and seen > 20
    assert (out["position"][1:-1, 1] > -0.3).all() and (out["position"][1:-1, 1] < 6).all()


@pytest.mark.gpu
@pytest.mark.parametrize("ground", ["plane", "heightfield"])
def test_compound_staged_parity(cuda_lib, oracle_lib, ground):
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, _pile_spec(ground))
    parity.assert_same_state(dev, ref, "upload", fields=("bounding_radius", "inv_inertia"))
    seen = 0
    for s in range(150):
        seen = max(seen, parity.staged_step(dev, ref, 1 / 60, f"compound on {ground} step {s}")[1])
        if s % 50 == 0:
            parity.assert_same_state(dev, ref, f"aabb step {s}", fields=("aabb",))
    assert seen > 20


@pytest.mark.gpu
@pytest.mark.parametrize("solver", [F.SOLVER_REFERENCE_ORDER, F.SOLVER_COLORED, F.SOLVER_SPLIT])
def test_compound_fused_parity(cuda_lib, oracle_lib, solver):
    spec = _pile_spec("heightfield", solver=solver, seed=6)
    spec.desc["broadphase_kind"] = F.BP_SAP
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, spec)
    for s in range(0, 240, 40):
        dev.step(1 / 60, 40)
        ref.step(1 / 60, 40)
        parity.assert_same_state(dev, ref, f"fused step {s + 40}")
    assert len(dev.get_contacts()["body_i"]) == len(ref.get_contacts()["body_i"]) > 10


@pytest.mark.gpu
def test_compound_rays_and_events_parity(cuda_lib, oracle_lib):
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, _pile_spec("plane", seed=8))
    for w in (dev, ref):
        w.enable_contact_events(True)
    n_begin = 0
    for s in range(60):
        dev.step(1 / 60)
        ref.step(1 / 60)
        (ba, ea), (bb, eb) = dev.get_contact_events(), ref.get_contact_events()
        assert np.array_equal(ba, bb) and np.array_equal(ea, eb), f"events differ at step {s}"
        n_begin += len(ba)
    assert n_begin > 5
    rng = np.random.default_rng(5)
    frm = np.column_stack([rng.uniform(-2.5, 2.5, 96), np.full(96, 6.0), rng.uniform(-1.5, 1.5, 96)]).astype(np.float32)
    to = frm + np.array([0.3, -8, -0.2], np.float32)
    for mode in (F.RAY_CLOSEST, F.RAY_ANY, F.RAY_ALL):
        a, b = dev.raycast(frm, to, mode=mode), ref.raycast(frm, to, mode=mode)
        for k in ("has_hit", "body", "distance", "hit_point_world", "hit_normal_world", "hit_face_index", "shape_ordinal"):
            assert np.array_equal(a[k], b[k]), (mode, k)
    assert a["n_hits"] > 20 and a["shape_ordinal"].max() > 0


def _batch_spec(n_worlds=5, solver=None):
    """n_worlds copies (slightly perturbed) of: ground plane, a table, a dumbbell, a plain box - world-major, with world ids."""
    shapes = [api.Box((0.7, 0.08, 0.45)), api.Box((0.08, 0.35, 0.08)), api.Cylinder(0.1, 0.1, 1.0, 8), api.Sphere(0.3), api.Box((0.3, 0.3, 0.3)), dict(type=F.SHAPE_PLANE)]
    shapes = [s._desc() if hasattr(s, "_desc") else s for s in shapes]
    rng = np.random.default_rng(12)
    bodies, wid = [], []
    for w in range(n_worlds):
        bodies.append(dict(pos=(0, 0, 0), mass=0, quat=scenes.GROUND_QUAT, inst=[(5, None, None)]))
        bodies.append(dict(pos=(0.05 * w, 0.9, 0), mass=2.0, inst=_table(0)))
        bodies.append(dict(pos=(0.1, 1.8 + 0.02 * w, 0.05), mass=1.0, inst=[(3, (-0.5, 0, 0), None), (3, (0.5, 0, 0), None)],
                           quat=(rng.normal(size=4) / 2).astype(np.float32)))
        bodies.append(dict(pos=(-0.2, 2.6, 0.1 * w), mass=1.0, inst=[(4, None, None)]))
        wid += [w] * 4
    for b in bodies:
        if "quat" in b and b["mass"] > 0:
            q = np.asarray(b["quat"], np.float64)
            b["quat"] = (q / np.linalg.norm(q)).astype(np.float32)
    desc = dict(n_worlds=n_worlds)
    if solver is not None:
        desc["solver_kind"] = solver
    spec = _spec(shapes, bodies, **desc)
    spec.bodies["world_id"] = np.array(wid, np.int32)
    return spec


def test_batch_of_compound_worlds_is_shard_independent(oracle_lib):
    """cannon_batch_set_body_shapes cuts the shape table at the shard boundaries (host glue shared by both libraries)."""
    from cannon_physics_b200 import engine
    fields = ("position", "quaternion", "velocity", "angular_velocity")
    for solver in (F.SOLVER_REFERENCE_ORDER, F.SOLVER_COLORED):
        spec = _batch_spec(5, solver)
        whole = engine.DeviceWorld(oracle_lib, spec)
        whole.step(1 / 60, 40)
        ref = whole.get_bodies(fields)
        assert len(whole.get_contacts()["body_i"]) > 10
        for devices in ((0,), (0, 0), (0, 0, 0)):
            b = engine.DeviceBatch(oracle_lib, spec, devices=devices)
            b.step(1 / 60, 40)
            got = b.get_bodies(fields)
            for k in fields:
                assert np.array_equal(got[k], ref[k]), (devices, k)
            b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("solver", [F.SOLVER_REFERENCE_ORDER, F.SOLVER_COLORED])
def test_batch_of_compound_worlds_parity(cuda_lib, oracle_lib, solver):
    spec = _batch_spec(24, solver)
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, spec)
    for s in range(0, 120, 40):
        dev.step(1 / 60, 40)
        ref.step(1 / 60, 40)
        parity.assert_same_state(dev, ref, f"compound batch step {s + 40}")
