"""Ray casts and AABB queries: source-derived known answers on the CPU oracle, CUDA == oracle on the GPU."""
import numpy as np
import pytest

from cannon_physics_b200 import _ffi as F
from cannon_physics_b200 import engine, scenes


def _scene():
    """A unit sphere at the origin, a box of half-extent 0.5 at x = 4, a ground plane at y = -2 (normal +y), a rotated
    cylinder at x = -4."""
    n = 4
    b = {
        "position": np.array([[0, 0, 0], [4, 0, 0], [0, -2, 0], [-4, 0, 0]], np.float32),
        "quaternion": np.array([[0, 0, 0, 1], [0, 0, 0, 1], scenes.GROUND_QUAT, [0.38268343, 0, 0, 0.92387953]], np.float32),
        "mass": np.array([1.0, 1.0, 0.0, 1.0]), "shape": np.array([0, 1, 2, 3], np.int32),
        "collision_filter_group": np.array([1, 2, 1, 1], np.int32),
    }
    shapes = [dict(type=F.SHAPE_SPHERE, radius=1.0), dict(type=F.SHAPE_BOX, half_extents=(0.5, 0.5, 0.5)), dict(type=F.SHAPE_PLANE),
              dict(type=F.SHAPE_CYLINDER, radius_top=0.5, radius_bottom=0.5, height=2.0, num_segments=8)]
    return engine.SceneSpec(desc=dict(gravity=(0, 0, 0)), shapes=shapes, bodies=b, n_bodies=n, name="rays")


def test_raycast_known_answers_on_the_oracle(oracle_lib):
    w = engine.DeviceWorld(oracle_lib, _scene())
    down = lambda x: ([x, 5, 0], [x, -5, 0])
    # closest: the sphere's top (ray_class.dart:411-458: d1 = (5 - 1) / 10), normal = (p - centre) normalised
    r = w.raycast(*map(lambda v: [v], down(0)))
    assert r["has_hit"][0] and r["body"][0] == 0 and r["distance"][0] == pytest.approx(4.0, abs=1e-6)
    assert np.allclose(r["hit_point_world"][0], (0, 1, 0), atol=1e-6) and np.allclose(r["hit_normal_world"][0], (0, 1, 0), atol=1e-6)
    assert r["hit_face_index"][0] == -1  # the plane behind reported last, faces only exist on hulls
    # the box: top face, distance 4.5; hitFaceIndex of CLOSEST is the LAST report of the sequence (:664) = the plane's -1
    r = w.raycast([[4, 5, 0]], [[4, -5, 0]])
    assert r["body"][0] == 1 and r["distance"][0] == pytest.approx(4.5, abs=1e-6) and np.allclose(r["hit_normal_world"][0], (0, 1, 0), atol=1e-6)
    assert r["hit_face_index"][0] == -1
    # ... and without the plane in the way of the sequence: a horizontal ray, face index of the box's -x face
    r = w.raycast([[0, 0, 5]], [[0, 0, -5]], mode=F.RAY_ALL, skip_backfaces=False)
    assert r["n_hits"] == 2 and list(r["body"]) == [0, 0] and np.allclose(r["distance"], (4.0, 6.0), atol=1e-6)  # entry and exit of the sphere
    r = w.raycast([[4, 0, 5]], [[4, 0, -5]], mode=F.RAY_CLOSEST)
    assert r["body"][0] == 1 and r["hit_face_index"][0] >= 0 and np.allclose(r["hit_normal_world"][0], (0, 0, 1), atol=1e-6)
    # the ground plane (plane.dart: normal = q * (0, 0, 1)): hit at y = -2
    r = w.raycast([[10, 5, 0]], [[10, -5, 0]])
    assert r["body"][0] == 2 and np.allclose(r["hit_point_world"][0], (10, -2, 0), atol=1e-5) and r["distance"][0] == pytest.approx(7.0, abs=1e-5)
    # any: the first candidate in body order that reports (the sphere), then stop
    r = w.raycast([[0, 5, 0]], [[0, -5, 0]], mode=F.RAY_ANY)
    assert r["has_hit"][0] and r["body"][0] == 0
    # collision filter (ray_class.dart:218-225): a ray of group 4 / mask ~2 does not see the box (group 2)
    r = w.raycast([[4, 5, 0]], [[4, -1, 0]], collision_filter_mask=~2)
    assert not r["has_hit"][0] and r["body"][0] == -1 and r["distance"][0] == -1.0
    # a miss, and a ray that ends before the sphere
    r = w.raycast([[0, 5, 3], [0, 5, 0]], [[0, -1, 3], [0, 2, 0]])
    assert list(r["has_hit"]) == [False, False]
    # aabbQuery (naive_broadphase.dart:39-56). The ground plane's world normal is (0, 0.99999994, 0) in float, not an axis, so
    # its AABB is infinite in every direction (plane.dart:44-69): it answers every query
    assert list(w.aabb_query((-0.5, -0.5, -0.5), (0.5, 0.5, 0.5))) == [0, 2]
    assert list(w.aabb_query((3.4, 1, -0.1), (3.6, 3, 0.1))) == [2] and list(w.aabb_query((3.4, -3, -0.1), (3.6, 3, 0.1))) == [1, 2]


def test_world_api_raycasts(oracle_lib):
    from cannon_physics_b200 import api
    w = api.World(gravity=(0, -10, 0), _lib=oracle_lib)
    g = api.Body(mass=0, shape=api.Plane())
    g.quaternion[:] = scenes.GROUND_QUAT
    s = api.Body(mass=1, shape=api.Sphere(0.5), position=(0, 3, 0))
    w.addBody(g)
    w.addBody(s)
    res = api.RaycastResult()
    assert w.raycastClosest((0, 10, 0), (0, -10, 0), result=res) and res.body is s and res.distance == pytest.approx(6.5, abs=1e-6)
    hits = []
    assert w.raycastAll((0, 10, 0), (0, -10, 0), {"skipBackfaces": False}, hits.append)
    assert [h.body for h in hits] == [g, s, s] and w.raycastAny((0, 10, 0), (0, -10, 0), result=res)
    for _ in range(30):
        w.step(1 / 60)  # the sphere falls: rays see the current poses
    assert w.raycastClosest((0, 10, 0), (0, -10, 0), result=res) and res.body is s and res.distance > 6.6
    assert w.aabbQuery((-1, 0, -1), (1, 4, 1)) == [g, s]
    assert not w.raycastClosest((5, 10, 0), (5, 1, 0), result=res) and res.body is None


def _random_rays(rng, n, extent):
    a = rng.uniform(-extent, extent, (n, 3)).astype(np.float32)
    b = rng.uniform(-extent, extent, (n, 3)).astype(np.float32)
    a[:, 1] = rng.uniform(0.0, extent, n)
    b[:, 1] = rng.uniform(-1.0, extent / 2, n)
    return a, b


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c2 box stacks", "mixed pile on a plane (spheres, boxes, cylinders)", "spheres in a container"])
def test_cuda_raycasts_equal_the_oracle(cuda_lib, oracle_lib, name):
    mk = {"c2 box stacks": lambda: scenes.box_stacks(4, 5, grid=2),
          "mixed pile on a plane (spheres, boxes, cylinders)": lambda: scenes.mixed_pile_on_heightfield(5, 5, 3, with_heightfield=False, grid_cells=(8, 4, 8)),
          "spheres in a container": lambda: scenes.sphere_container(5, 5, 3, extent=4.0)}[name]
    dev, ref = engine.DeviceWorld(cuda_lib, mk()), engine.DeviceWorld(oracle_lib, mk())
    rng = np.random.default_rng(11)
    for phase in range(3):
        dev.step(1 / 60, 40)
        ref.step(1 / 60, 40)
        a, b = _random_rays(rng, 600, 6.0)
        for mode in (F.RAY_CLOSEST, F.RAY_ANY, F.RAY_ALL):
            for skip in (True, False):
                x = dev.raycast(a, b, mode=mode, skip_backfaces=skip)
                y = ref.raycast(a, b, mode=mode, skip_backfaces=skip)
                assert x["n_hits"] == y["n_hits"], (name, phase, mode, skip)
                for k in ("has_hit", "ray", "body", "hit_face_index", "distance", "hit_point_world", "hit_normal_world", "shape_ordinal"):
                    assert np.array_equal(x[k], y[k]), (name, phase, mode, skip, k)
        assert x["n_hits"] > 10
        lo, hi = np.array([-1.5, -0.5, -1.5], np.float32), np.array([1.0, 2.5, 1.5], np.float32)
        assert np.array_equal(dev.aabb_query(lo, hi), ref.aabb_query(lo, hi))


@pytest.mark.gpu
def test_cuda_raycast_refuses_heightfields_and_bad_modes(cuda_lib):
    w = engine.DeviceWorld(cuda_lib, scenes.mixed_pile_on_heightfield(3, 3, 2, hf_samples=33, grid_cells=(8, 4, 8)))
    with pytest.raises(F.CannonError) as e:
        w.raycast([[0, 5, 0]], [[0, -5, 0]])
    assert e.value.code == F.E_UNSUPPORTED
