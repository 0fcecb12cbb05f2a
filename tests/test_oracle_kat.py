"""Source-derived known answers for the CPU oracle. The reference has no tests or golden
vectors and cannot run here, so these pin the restatement to hand-derived values of the Dart semantics
(float32 storage, float64 arithmetic, no FMA). PARITY UNPINNED with respect to a running reference."""
import math

import numpy as np

from cannon_physics_b200 import _ffi as F
from cannon_physics_b200 import engine, scenes
from cannon_physics_b200.engine import SceneSpec


def _world(oracle_lib, shapes, bodies, n, **desc):
    d = dict(gravity=(0, -10, 0))
    d.update(desc)
    return engine.DeviceWorld(oracle_lib, SceneSpec(desc=d, shapes=shapes, bodies=bodies, n_bodies=n))


def test_free_fall_two_steps(oracle_lib):
    # (i) g=(0,-10,0), m=1, linearDamping 0.01, dt=1/60, y0=10
    b = {"position": np.array([[0, 10, 0]], np.float32), "mass": np.array([1.0]), "shape": np.array([0], np.int32)}
    w = _world(oracle_lib, [dict(type=F.SHAPE_SPHERE, radius=0.5)], b, 1)
    assert math.pow(0.99, 1 / 60) == 0.99983250843072091
    w.step(1 / 60)
    s = w.get_bodies(("position", "velocity"))
    assert float(s["velocity"][0, 1]) == -0.1666666716337204
    assert float(s["position"][0, 1]) == 9.9972219467163086
    w.step(1 / 60)
    s = w.get_bodies(("position", "velocity"))
    assert float(s["velocity"][0, 1]) == -0.33330541849136353
    assert float(s["position"][0, 1]) == 9.9916667938232422


def test_ground_plane_quaternion_and_infinite_aabb(oracle_lib):
    # (iii) setFromEuler(-pi/2,0,0) -> f32 (-0.70710677,0,0,0.70710677); world normal y = 0.99999994 => AABB infinite (§5.9-13)
    q = scenes.GROUND_QUAT
    assert q.tolist() == [np.float32(-0.70710677), 0.0, 0.0, np.float32(0.70710677)]
    b = {"quaternion": q[None, :].copy(), "shape": np.array([0], np.int32)}
    w = _world(oracle_lib, [dict(type=F.SHAPE_PLANE)], b, 1)
    s = w.get_bodies(("aabb", "bounding_radius"))
    assert np.all(np.isinf(s["aabb"])) and np.all(s["aabb"][0, :3] < 0) and np.all(s["aabb"][0, 3:] > 0)
    assert math.isinf(s["bounding_radius"][0])


def test_sphere_mass_properties(oracle_lib):
    # (iv) sphere r=0.5 m=1: AABB-box inertia 1/12*m*(1+1) = 1/6 -> f32 0.1666666716337204 -> invInertia 6.0
    b = {"position": np.zeros((1, 3), np.float32), "mass": np.array([1.0]), "shape": np.array([0], np.int32)}
    w = _world(oracle_lib, [dict(type=F.SHAPE_SPHERE, radius=0.5)], b, 1)
    s = w.get_bodies(("inv_inertia", "inv_mass", "inv_inertia_world", "bounding_radius"))
    assert s["inv_mass"][0] == 1.0 and s["bounding_radius"][0] == 0.5
    expected = np.float32(1.0 / float(np.float32(1.0 / 6.0)))
    assert np.all(s["inv_inertia"][0] == expected) and abs(float(expected) - 6.0) < 1e-6
    assert np.allclose(np.diag(s["inv_inertia_world"][0].reshape(3, 3)), 6.0, atol=1e-6)


def test_sphere_on_plane_contact_and_rows(oracle_lib):
    # (v) one contact, ni=(0,-1,0), ri=(0,-r,0); friction bound 0.3*10*m; row order [f1,f2,n]; SPOOK (ii)
    r = 0.5
    b = {"position": np.array([[0, 0, 0], [0, 0.4, 0]], np.float32), "quaternion": np.stack([scenes.GROUND_QUAT, [0, 0, 0, 1]]).astype(np.float32),
         "mass": np.array([0.0, 1.0]), "shape": np.array([0, 1], np.int32)}
    w = _world(oracle_lib, [dict(type=F.SHAPE_PLANE), dict(type=F.SHAPE_SPHERE, radius=r)], b, 2)
    w.set_dt(1 / 60)
    p1, p2 = w.broadphase_pairs()
    assert p1.tolist() == [1] and p2.tolist() == [0]  # (bodies[i], bodies[j<i])
    c = w.narrowphase_contacts(p1, p2)
    assert c["per_pair_count"].tolist() == [1]
    assert c["body_i"].tolist() == [1] and c["body_j"].tolist() == [0]  # sphere (type 0) before plane (type 1)
    assert c["ni"][0].tolist() == [0.0, -1.0, 0.0]
    assert c["ri"][0].tolist() == [0.0, -r, 0.0]
    assert c["friction"][0] == 0.3 and c["restitution"][0] == 0.0
    w.solver_solve(1 / 60)
    rows = w.get_rows()
    assert len(rows["B"]) == 3
    # default contact SPOOK (k=1e7, d=3, h=1/60)
    h, k, d = 1 / 60, 1e7, 3.0
    a = 4.0 / (h * (1 + 4 * d))
    bb = 4.0 * d / (1 + 4 * d)
    eps = 4.0 / (h * h * k * (1 + 4 * d))
    assert (a, bb, eps) == (18.46153846153846, 0.92307692307692313, 1.1076923076923077e-4)
    # contact row: g = n.(xj+rj-xi-ri) = -(0 - 0.4 + 0.5) = -0.1 (f32 arithmetic); no velocity, no force => B = -g*a
    g = float(np.float32(-1.0)) * float(np.float32(np.float32(np.float32(0.0) - np.float32(0.4)) - np.float32(-0.5)))
    assert rows["B"][2] == -g * a
    # friction rows are bounded by +-mu*|g|*m_red = 3.0 and have zero B here
    assert rows["B"][0] == 0.0 and rows["B"][1] == 0.0
    # C = invMass + rixn.I.rixn + eps with rixn = 0 for a centred normal
    assert rows["invC"][2] == 1.0 / (1.0 + eps)


def test_box_on_plane_four_contacts(oracle_lib):
    # (vi) axis-aligned box resting on the plane: the 4 lower vertices satisfy n.rel <= 0
    b = {"position": np.array([[0, 0, 0], [0, 0.49, 0]], np.float32), "quaternion": np.stack([scenes.GROUND_QUAT, [0, 0, 0, 1]]).astype(np.float32),
         "mass": np.array([0.0, 1.0]), "shape": np.array([0, 1], np.int32)}
    w = _world(oracle_lib, [dict(type=F.SHAPE_PLANE), dict(type=F.SHAPE_BOX, half_extents=(0.5, 0.5, 0.5))], b, 2)
    p1, p2 = w.broadphase_pairs()
    c = w.narrowphase_contacts(p1, p2)
    assert c["per_pair_count"].tolist() == [4]
    assert c["body_i"].tolist() == [0] * 4 and c["body_j"].tolist() == [1] * 4  # plane (1) before box (2)
    assert np.all(c["rj"][:, 1] == -0.5)


def test_sphere_sphere_always_one_contact_when_bounding_spheres_touch(oracle_lib):
    # §5.9-8: sphereSphere has no distance test of its own; equal types arrive swapped (bi = p2[k], §5.9-7)
    b = {"position": np.array([[0, 0, 0], [0.9, 0, 0]], np.float32), "mass": np.array([1.0, 1.0]), "shape": np.array([0, 0], np.int32)}
    w = _world(oracle_lib, [dict(type=F.SHAPE_SPHERE, radius=0.5)], b, 2, gravity=(0, 0, 0))
    p1, p2 = w.broadphase_pairs()
    assert (p1.tolist(), p2.tolist()) == ([1], [0])
    c = w.narrowphase_contacts(p1, p2)
    assert c["body_i"].tolist() == [0] and c["body_j"].tolist() == [1]
    assert c["ni"][0].tolist() == [1.0, 0.0, 0.0]
    b["position"][1, 0] = 1.0  # exactly touching: norm2 < (ra+rb)^2 is false => no pair
    w2 = _world(oracle_lib, [dict(type=F.SHAPE_SPHERE, radius=0.5)], b, 2, gravity=(0, 0, 0))
    assert len(w2.broadphase_pairs()[0]) == 0


def test_box_box_stack_contacts(oracle_lib):
    # box resting on a box: convexConvex clips the incident face against the reference face: 4 contacts
    b = {"position": np.array([[0, 0.5, 0], [0, 1.49, 0]], np.float32), "mass": np.array([0.0, 1.0]), "shape": np.array([0, 0], np.int32)}
    w = _world(oracle_lib, [dict(type=F.SHAPE_BOX, half_extents=(0.5, 0.5, 0.5))], b, 2)
    p1, p2 = w.broadphase_pairs()
    c = w.narrowphase_contacts(p1, p2)
    assert c["per_pair_count"].tolist() == [4]
    assert np.all(np.abs(c["ni"][:, 1]) == 1.0)


def _events_world(lib):
    """Two spheres over a plane: body 0 = plane, 1 = sphere resting just above it, 2 = sphere dropped on sphere 1."""
    from cannon_physics_b200.scenes import GROUND_QUAT
    b = {"position": np.array([[0, 0, 0], [0, 0.6, 0], [0, 2.5, 0]], np.float32),
         "quaternion": np.array([GROUND_QUAT, [0, 0, 0, 1], [0, 0, 0, 1]], np.float32),
         "mass": np.array([0.0, 1.0, 1.0]), "shape": np.array([0, 1, 1], np.int32),
         "type": np.array([F.BODY_STATIC, F.BODY_DYNAMIC, F.BODY_DYNAMIC], np.int32)}
    shapes = [dict(type=F.SHAPE_PLANE), dict(type=F.SHAPE_SPHERE, radius=0.5)]
    return engine.DeviceWorld(lib, SceneSpec(desc=dict(gravity=(0, -10, 0)), shapes=shapes, bodies=b, n_bodies=3))


def test_contact_events_follow_overlap_keeper(oracle_lib):
    # world_class.dart:606,703-730 + overlap_keeper.dart: a pair is announced once when its first ContactEquation
    # appears, stays silent while it keeps one, and is announced again (endContact) in the step that has none
    w = _events_world(oracle_lib)
    w.enable_contact_events(True)
    seen_begin, seen_end, in_contact = [], [], set()
    for step in range(90):
        w.step(1 / 60)
        begin, end = w.get_contact_events()
        c = w.get_contacts()
        now = {(min(int(a), int(b)), max(int(a), int(b))) for a, b in zip(c["body_i"], c["body_j"])}
        assert {tuple(p) for p in begin.tolist()} == now - in_contact, step
        assert {tuple(p) for p in end.tolist()} == in_contact - now, step
        assert begin.tolist() == sorted(begin.tolist()) and end.tolist() == sorted(end.tolist())
        assert all(a < b for a, b in begin.tolist() + end.tolist())
        in_contact = now
        seen_begin += [tuple(p) for p in begin.tolist()]
        seen_end += [tuple(p) for p in end.tolist()]
    assert (0, 1) in seen_begin and (1, 2) in seen_begin  # sphere 1 lands on the plane, sphere 2 lands on sphere 1
    # re-enabling starts from an empty previous set: everything in contact is announced again
    w.enable_contact_events(True)
    w.step(1 / 60)
    begin, end = w.get_contact_events()
    assert {tuple(p) for p in begin.tolist()} == {(min(int(a), int(b)), max(int(a), int(b))) for a, b in
                                                   zip(w.get_contacts()["body_i"], w.get_contacts()["body_j"])}
    assert len(end) == 0
