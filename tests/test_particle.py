"""Particle shape: sphereParticle / planeParticle / boxParticle / particleConvex /
heightfieldParticle, lib/world/narrow_phase.dart:1258,1805,1731,2179,2343.

particleConvex measures the penetration against ConvexPolyhedron.worldVertices / worldFaceNormals, which the reference
computes on the first penetration a hull ever sees and never refreshes (convex_polyhedron.dart:101-103,603,645;
narrow_phase.dart:2207-2212). The oracle and the device keep that state with the shape table; the history test below
pins it. CPU tests are source-derived known answers on the oracle, the GPU tests are bit-exact parity."""
import numpy as np
import pytest

import parity
from cannon_physics_b200 import _ffi as F
from cannon_physics_b200 import api, scenes
from cannon_physics_b200.engine import DeviceWorld, SceneSpec

IDENT = np.array([0, 0, 0, 1], np.float32)


def _spec(shapes, pos, mass, quat=None, **desc):
    n = len(pos)
    q = np.tile(IDENT, (n, 1)) if quat is None else np.asarray(quat, np.float32)
    return SceneSpec(desc=dict(dict(gravity=(0, 0, 0)), **desc), shapes=[s._desc() if hasattr(s, "_desc") else s for s in shapes],
                     bodies=dict(position=np.asarray(pos, np.float32), quaternion=q, mass=np.asarray(mass, np.float64), shape=np.arange(n, dtype=np.int32)), n_bodies=n)


def _contacts(world):
    world.set_dt(1 / 60)
    p = world.broadphase_pairs()
    return world.narrowphase_contacts(*p)


def test_sphere_particle_contact(oracle_lib):
    # particle 0.3 from the centre of a radius-0.5 sphere: one contact, the particle's body first, rj on the sphere surface
    w = DeviceWorld(oracle_lib, _spec([api.Sphere(0.5), api.Particle()], [[0, 0, 0], [0.3, 0, 0]], [1, 1]))
    c = _contacts(w)
    assert len(c["body_i"]) == 1 and c["body_i"][0] == 1 and c["body_j"][0] == 0
    np.testing.assert_array_equal(c["ni"][0], [-1, 0, 0])
    np.testing.assert_array_equal(c["ri"][0], [0, 0, 0])
    np.testing.assert_array_equal(c["rj"][0], [0.5, 0, 0])
    # outside the radius: the prologue's bounding test (radius + 0) already rejects it
    w = DeviceWorld(oracle_lib, _spec([api.Sphere(0.5), api.Particle()], [[0, 0, 0], [0.6, 0, 0]], [1, 1]))
    assert len(_contacts(w)["body_i"]) == 0


def test_plane_particle_contact(oracle_lib):
    # ground plane (normal +y after the usual rotation); particle 0.1 below it
    w = DeviceWorld(oracle_lib, _spec([dict(type=F.SHAPE_PLANE), api.Particle()], [[0, 0, 0], [0.5, -0.1, 0.25]], [0, 1], quat=[scenes.GROUND_QUAT, IDENT]))
    c = _contacts(w)
    assert len(c["body_i"]) == 1 and c["body_i"][0] == 1 and c["body_j"][0] == 0
    assert abs(c["ni"][0][1] + 1) < 1e-6 and abs(c["ni"][0][0]) < 1e-6
    np.testing.assert_allclose(c["rj"][0], [0.5, 0, 0.25], atol=1e-6)  # projected on the plane; the plane position is not subtracted
    w = DeviceWorld(oracle_lib, _spec([dict(type=F.SHAPE_PLANE), api.Particle()], [[0, 0, 0], [0.5, 0.1, 0.25]], [0, 1], quat=[scenes.GROUND_QUAT, IDENT]))
    assert len(_contacts(w)["body_i"]) == 0


def test_particle_in_box_uses_the_pose_of_the_first_penetration(oracle_lib):
    box = api.Box((0.5, 0.5, 0.5))
    w = DeviceWorld(oracle_lib, _spec([box, api.Particle()], [[0, 0, 0], [0.4, 0.1, 0.0]], [1, 1]))
    c = _contacts(w)
    assert len(c["body_i"]) == 1 and c["body_i"][0] == 1 and c["body_j"][0] == 0
    np.testing.assert_array_equal(c["ni"][0], [-1, 0, 0])       # nearest face +x, negated
    np.testing.assert_allclose(c["rj"][0], [0.5, 0.1, 0.0], atol=1e-6)  # the particle projected on that face
    # move box and particle together by +10 in x: pointIsInside (local frame) still holds, but worldVertices are never
    # recomputed (narrow_phase.dart:2207 only runs once), so the penetration is measured against the box at x = 0
    w.update_bodies(0, 2, position=np.array([[10, 0, 0], [10.4, 0.1, 0]], np.float32))
    c2 = _contacts(w)
    assert len(c2["body_i"]) == 1
    # frozen box at the origin, particle at (10.4, 0.1, 0): the smallest |penetration| is now the +y face (0.4), so the
    # contact flips to that face; a box that followed its body would answer (0.5, 0.1, 0) with normal -x as before
    np.testing.assert_array_equal(c2["ni"][0], [0, -1, 0])
    np.testing.assert_allclose(c2["rj"][0], [0.4, 0.5, 0.0], atol=1e-5)
    # a fresh shape table forgets the frozen pose
    w2 = DeviceWorld(oracle_lib, _spec([box, api.Particle()], [[10, 0, 0], [10.4, 0.1, 0.0]], [1, 1]))
    np.testing.assert_allclose(_contacts(w2)["rj"][0], [0.5, 0.1, 0.0], atol=1e-5)


def _pile_spec(ground, solver=None, seed=2, n_part=40):
    """Particles raining on a ground plus a few hull / sphere bodies."""
    rng = np.random.default_rng(seed)
    shapes = []
    if ground == "plane":
        shapes.append(dict(type=F.SHAPE_PLANE))
        gp, gq = (0, 0, 0), scenes.GROUND_QUAT
    else:
        hf = 0.2 * np.random.default_rng(9).random((12, 12))
        shapes.append(dict(type=F.SHAPE_HEIGHTFIELD, hf_data=hf, hf_element_size=1))
        gp, gq = (-5.5, 0, 5.5), scenes.GROUND_QUAT
    solids = [api.Box((0.8, 0.4, 0.8)), api.Sphere(0.7), api.Cylinder(0.6, 0.6, 0.8, 8), api.Cone(0.7, 1.0, 8)]
    shapes += [s._desc() for s in solids] + [api.Particle()._desc()]
    n = 1 + len(solids) + n_part
    pos = np.zeros((n, 3), np.float32)
    quat = np.tile(IDENT, (n, 1))
    mass = np.ones(n)
    shape = np.zeros(n, np.int32)
    pos[0], quat[0], mass[0] = gp, gq, 0.0
    for k in range(len(solids)):
        pos[1 + k] = (2.0 * (k % 2) - 1.0, 1.2, 2.0 * (k // 2) - 1.0)
        shape[1 + k] = 1 + k
        mass[1 + k] = 5.0
    for k in range(n_part):
        pos[1 + len(solids) + k] = (rng.uniform(-2, 2), rng.uniform(0.3, 3.0), rng.uniform(-2, 2))
        shape[1 + len(solids) + k] = 1 + len(solids)
        mass[1 + len(solids) + k] = 0.2
    desc = dict(gravity=(0, -10, 0))
    if solver is not None:
        desc["solver_kind"] = solver
    return SceneSpec(desc=desc, shapes=shapes, bodies=dict(position=pos, quaternion=quat, mass=mass, shape=shape), n_bodies=n, name=f"particles on {ground}")


def test_oracle_particles_land_on_the_plane(oracle_lib):
    w = DeviceWorld(oracle_lib, _pile_spec("plane"))
    seen = 0
    for _ in range(120):
        w.step(1 / 60)
        seen = max(seen, len(w.get_contacts()["body_i"]))
    out = w.get_bodies(("position",))
    assert np.isfinite(out["position"]).all() and seen > 10
    part = out["position"][5:]
    assert (part[:, 1] > -0.2).mean() > 0.6  # most particles are held by the plane (those inside hulls get the reference's stale answers)


@pytest.mark.gpu
@pytest.mark.parametrize("ground", ["plane", "heightfield"])
def test_particle_staged_parity(cuda_lib, oracle_lib, ground):
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, _pile_spec(ground))
    seen = 0
    for s in range(150):
        seen = max(seen, parity.staged_step(dev, ref, 1 / 60, f"particles on {ground} step {s}")[1])
    assert seen > 10


@pytest.mark.gpu
@pytest.mark.parametrize("solver", [F.SOLVER_REFERENCE_ORDER, F.SOLVER_COLORED])
def test_particle_fused_parity(cuda_lib, oracle_lib, solver):
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, _pile_spec("heightfield", solver=solver, seed=4))
    for s in range(0, 200, 40):
        dev.step(1 / 60, 40)
        ref.step(1 / 60, 40)
        parity.assert_same_state(dev, ref, f"fused step {s + 40}")


@pytest.mark.gpu
def test_particle_in_hull_history_matches_on_the_device(cuda_lib, oracle_lib):
    spec = _spec([api.Box((0.5, 0.5, 0.5)), api.Particle(), api.Cylinder(0.5, 0.5, 1.0, 8), api.Particle()],
                 [[0, 0, 0], [0.4, 0.1, 0.0], [3, 0, 0], [3.1, 0.2, 0.1]], [1, 1, 1, 1])
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, spec)
    moves = [None, np.array([[10, 0, 0], [10.4, 0.1, 0], [3, 1, 0], [3.1, 1.2, 0.1]], np.float32),
             np.array([[0, 0, 0], [0.1, 0.3, 0.2], [3, 0, 0], [2.9, -0.2, 0.1]], np.float32)]
    for k, mv in enumerate(moves):
        if mv is not None:
            for w in (dev, ref):
                w.update_bodies(0, 4, position=mv)
        ca, cb = _contacts(dev), _contacts(ref)
        parity.assert_same_contacts(ca, cb, f"history {k}")
        assert len(ca["body_i"]) == 2
