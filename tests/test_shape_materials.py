"""Shape.material (lib/rigid_body_shapes/shape.dart:48) as the reference uses it: the contact material of a shape pair prefers the
shapes' materials (narrow_phase.dart:692-696); createContactEquation takes `shape.material ?? body.material` in resolver order
(:517-521); World.internalStep overrides the restitution by the two BODY materials when both exist (world_class.dart:556-560);
createFrictionEquationsFromContact pairs the shapes in PAIR order (c.si = rsi) with the bodies in RESOLVER order (c.bi), so a pair
whose shapes arrive swapped mixes one body's shape material with the other body's body material (:533-542)."""
import numpy as np
import pytest

import parity
from cannon_physics_b200 import _ffi as F
from cannon_physics_b200 import api, scenes
from cannon_physics_b200.engine import DeviceWorld, SceneSpec

IDENT = np.array([0, 0, 0, 1], np.float32)
# material table: 0 = M_s, 1 = M_p, 2 = M_b
FRICTION = np.array([0.8, 0.5, 0.25])
RESTITUTION = np.array([0.5, 0.9, 0.6])


def _spec(shapes, pos, mass, body_material, quat=None, cms=(), **desc):
    n = len(pos)
    q = np.tile(IDENT, (n, 1)) if quat is None else np.asarray(quat, np.float32)
    return SceneSpec(desc=dict(dict(gravity=(0, -10, 0)), **desc), shapes=shapes,
                     bodies=dict(position=np.asarray(pos, np.float32), quaternion=q, mass=np.asarray(mass, np.float64), shape=np.arange(n, dtype=np.int32),
                                 material=np.asarray(body_material, np.int32)), n_bodies=n,
                     material_friction=FRICTION, material_restitution=RESTITUTION, contact_materials=list(cms))


def _contacts(world):
    world.set_dt(1 / 60)
    return world.narrowphase_contacts(*world.broadphase_pairs())


def test_shape_material_beats_the_body_material_in_resolver_order(oracle_lib):
    # sphere (shape material M_s, no body material) resting in a plane whose BODY carries M_p
    shapes = [dict(type=F.SHAPE_PLANE), dict(api.Sphere(0.5)._desc(), material=0)]
    w = DeviceWorld(oracle_lib, _spec(shapes, [[0, 0, 0], [0, 0.45, 0]], [0, 1], [1, -1], quat=[scenes.GROUND_QUAT, IDENT]))
    c = _contacts(w)
    assert len(c["body_i"]) == 1 and (c["body_i"][0], c["body_j"][0]) == (1, 0)
    assert abs(c["restitution"][0] - 0.5 * 0.9) < 1e-15   # M_s (shape) x M_p (the plane's body material)
    assert abs(c["friction"][0] - 0.8 * 0.5) < 1e-15
    # with a body material on the sphere as well, World.internalStep overrides the restitution by the two body materials
    w = DeviceWorld(oracle_lib, _spec(shapes, [[0, 0, 0], [0, 0.45, 0]], [0, 1], [1, 2], quat=[scenes.GROUND_QUAT, IDENT]))
    assert abs(_contacts(w)["restitution"][0] - 0.5 * 0.9) < 1e-15      # createContactEquation: still the shape's
    w.step(1 / 60)
    c = w.get_contacts()
    assert abs(c["restitution"][0] - 0.6 * 0.9) < 1e-15                 # after the step: M_b x M_p
    assert abs(c["friction"][0] - 0.8 * 0.5) < 1e-15                    # the friction keeps the shape's material


def test_swapped_pairs_mix_shape_and_body_materials(oracle_lib):
    # pair (box body 1, sphere body 0) reaches sphereBox with its shapes swapped: c.bi = the sphere's body, c.si = the BOX shape
    sphere = dict(api.Sphere(0.5)._desc(), material=0)       # M_s on the sphere shape, no body material
    box = api.Box((0.5, 0.5, 0.5))._desc()                   # no shape material; the box BODY carries M_b
    w = DeviceWorld(oracle_lib, _spec([sphere, box], [[0, 0, 0], [0.9, 0, 0]], [1, 1], [-1, 2], gravity=(0, 0, 0)))
    c = _contacts(w)
    assert len(c["body_i"]) == 1 and (c["body_i"][0], c["body_j"][0]) == (0, 1)
    # restitution: resolver order, shape ?? body on both sides: M_s x M_b
    assert abs(c["restitution"][0] - 0.5 * 0.6) < 1e-15
    # friction: matA = rsi (box shape: none) ?? c.bi (sphere body: none) -> null -> the contact material's friction stays
    assert abs(c["friction"][0] - 0.3) < 1e-15


def _pile_spec(solver=None):
    """A small mixed pile in which every combination occurs: shape materials on some shapes, body materials on some bodies, a shape
    contact material, a compound body whose two shapes differ, a heightfield shape with its own material."""
    rng = np.random.default_rng(21)
    hf = 0.15 * rng.random((10, 10))
    shapes = [dict(type=F.SHAPE_HEIGHTFIELD, hf_data=hf, hf_element_size=1, material=1),
              dict(api.Sphere(0.35)._desc(), material=0), api.Sphere(0.35)._desc(),
              dict(api.Box((0.3, 0.3, 0.3))._desc(), material=2), api.Box((0.3, 0.3, 0.3))._desc(),
              dict(api.Cylinder(0.3, 0.3, 0.6, 8)._desc(), material=0)]
    bodies = [dict(pos=(-4.5, 0, 4.5), quat=scenes.GROUND_QUAT, mass=0, mat=-1, inst=[(0, None, None)])]
    k = 0
    for layer in range(2):
        for ix in range(3):
            for iz in range(3):
                s = 1 + (k % 5)
                inst = [(s, None, None)] if k % 4 else [(1, (-0.3, 0, 0), None), (4, (0.3, 0, 0), None)]
                bodies.append(dict(pos=(0.9 * ix - 0.9 + 0.05 * rng.random(), 0.8 + 0.9 * layer, 0.9 * iz - 0.9 + 0.05 * rng.random()), mass=1.0,
                                   mat=[-1, 1, 2][k % 3], inst=inst, quat=IDENT))
                k += 1
    first, shape, off = [0], [], []
    for b in bodies:
        for (s, o, _) in b["inst"]:
            shape.append(s)
            off.append((0, 0, 0) if o is None else o)
        first.append(len(shape))
    n = len(bodies)
    cms = [dict(material_a=0, material_b=1, friction=0.05, restitution=0.2), dict(material_a=1, material_b=2, friction=0.6, restitution=0.1)]
    desc = dict(gravity=(0, -10, 0))
    if solver is not None:
        desc["solver_kind"] = solver
    return SceneSpec(desc=desc, shapes=shapes,
                     bodies=dict(position=np.array([b["pos"] for b in bodies], np.float32), quaternion=np.array([b["quat"] for b in bodies], np.float32),
                                 mass=np.array([b["mass"] for b in bodies], np.float64), material=np.array([b["mat"] for b in bodies], np.int32)),
                     n_bodies=n, material_friction=FRICTION, material_restitution=RESTITUTION, contact_materials=cms,
                     body_shapes=dict(first=np.array(first, np.int32), shape=np.array(shape, np.int32), offset=np.array(off, np.float32), orientation=None),
                     name="shape materials")


def test_oracle_shape_material_pile_uses_several_material_values(oracle_lib):
    w = DeviceWorld(oracle_lib, _pile_spec())
    seen_f, seen_r = set(), set()
    for _ in range(90):
        w.step(1 / 60)
        c = w.get_contacts()
        seen_f |= set(np.round(c["friction"], 6).tolist())
        seen_r |= set(np.round(c["restitution"], 6).tolist())
    assert len(seen_f) >= 4 and len(seen_r) >= 3, (seen_f, seen_r)


@pytest.mark.gpu
@pytest.mark.parametrize("solver", [F.SOLVER_REFERENCE_ORDER, F.SOLVER_COLORED])
def test_shape_materials_parity(cuda_lib, oracle_lib, solver):
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, _pile_spec(solver))
    for s in range(90):
        parity.staged_step(dev, ref, 1 / 60, f"shape materials step {s}")   # compares restitution / friction per contact
    ca, cb = dev.get_contacts(), ref.get_contacts()
    parity.assert_same_contacts(ca, cb, "after the step (restitution overridden by the body materials)")
    for s in range(0, 80, 40):
        dev.step(1 / 60, 40)
        ref.step(1 / 60, 40)
        parity.assert_same_state(dev, ref, f"fused step {s + 40}")
