"""Trimesh: sphereTrimesh and planeTrimesh, lib/world/narrow_phase.dart:1438-1691,1916-1980, over
lib/rigid_body_shapes/trimesh.dart. The Dart port walks every triangle in index order (its octree query is unused) and runs
the triangle-face test inside the per-corner loop, so a face contact is reported three times - reproduced. Pairs that only
the reference's unfinished resolvers would handle (box / convex / particle / trimesh against a trimesh) make the step
return CANNON_E_UNSUPPORTED. CPU tests: source-derived known answers on the oracle; GPU tests: bit-exact parity."""
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
    return SceneSpec(desc=dict(dict(gravity=(0, -10, 0)), **desc), shapes=[s._desc() if hasattr(s, "_desc") else s for s in shapes],
                     bodies=dict(position=np.asarray(pos, np.float32), quaternion=q, mass=np.asarray(mass, np.float64), shape=np.arange(n, dtype=np.int32)), n_bodies=n)


def _contacts(world):
    world.set_dt(1 / 60)
    return world.narrowphase_contacts(*world.broadphase_pairs())


def _grid_mesh(n=6, size=1.0, amp=0.15, seed=4):
    """A bumpy n x n vertex sheet in the xz plane, two triangles per cell, wound so the normals point up."""
    rng = np.random.default_rng(seed)
    h = amp * rng.random((n, n))
    verts = np.array([[(i - (n - 1) / 2) * size, h[i, j], (j - (n - 1) / 2) * size] for i in range(n) for j in range(n)], np.float64)
    idx = []
    for i in range(n - 1):
        for j in range(n - 1):
            a, b, c, d = i * n + j, (i + 1) * n + j, (i + 1) * n + j + 1, i * n + j + 1
            idx += [a, d, b, b, d, c]
    return api.Trimesh(verts, idx)


def test_sphere_over_a_triangle_reports_the_face_three_times(oracle_lib):
    tri = api.Trimesh([(-2, 0, -2), (-2, 0, 4), (4, 0, -2)], [0, 1, 2])
    w = DeviceWorld(oracle_lib, _spec([tri, api.Sphere(0.5)], [[0, 0, 0], [0.2, 0.3, 0.1]], [0, 1]))
    c = _contacts(w)
    # centre 0.3 above the interior: no vertex or edge within the radius; the face test (:1573-1603) runs once per corner j
    assert len(c["body_i"]) == 3 and set(c["body_i"].tolist()) == {1} and set(c["body_j"].tolist()) == {0}
    for k in range(3):
        np.testing.assert_allclose(c["ni"][k], [0, -1, 0], atol=1e-6)
        np.testing.assert_allclose(c["ri"][k], [0, -0.5, 0], atol=1e-6)
        np.testing.assert_allclose(c["rj"][k], [0.2, 0, 0.1], atol=1e-6)
    # near a corner: the vertex contact of both... one triangle: corner j = 0 only, plus the two edges that meet there
    w = DeviceWorld(oracle_lib, _spec([tri, api.Sphere(0.5)], [[0, 0, 0], [-2.1, 0.2, -2.1]], [0, 1]))
    c = _contacts(w)
    assert len(c["body_i"]) == 1  # only the vertex: the projection on both edges falls outside (positionAlongEdge tests)
    np.testing.assert_allclose(c["rj"][0], [-2, 0, -2], atol=1e-6)


def test_plane_trimesh_contacts_are_the_vertices_below_the_plane(oracle_lib):
    torus = api.Trimesh.createTorus(1.0, 0.4, 8, 6)
    assert torus.vertices.shape == (9 * 7, 3) and len(torus.indices) == 8 * 6 * 6
    # torus axis along z, standing on the ground plane (normal +y): with 6 tubular segments the lowest ring is at y = -1.4 sin 60
    w = DeviceWorld(oracle_lib, _spec([dict(type=F.SHAPE_PLANE), torus], [[0, 0, 0], [0, 1.1, 0]], [0, 1], quat=[scenes.GROUND_QUAT, IDENT]))
    c = _contacts(w)
    below = (torus.vertices[:, 1].astype(np.float32) + np.float32(1.1)) <= 1e-7
    assert len(c["body_i"]) == int(below.sum()) > 0
    assert set(c["body_i"].tolist()) == {0} and set(c["body_j"].tolist()) == {1}
    np.testing.assert_allclose(c["ni"][:, 1], 1, atol=1e-6)
    b = w.get_bodies(("bounding_radius", "aabb"))
    assert abs(b["bounding_radius"][1] - 1.4) < 1e-6
    hy = 1.4 * np.sin(np.pi / 3)
    np.testing.assert_allclose(b["aabb"][1], [-1.4, 1.1 - hy, -0.4, 1.4, 1.1 + hy, 0.4], atol=1e-5)


def test_unfinished_trimesh_pairs_are_refused(oracle_lib):
    mesh = _grid_mesh()
    w = DeviceWorld(oracle_lib, _spec([mesh, api.Box((0.3, 0.3, 0.3))], [[0, 0, 0], [0, 0.3, 0]], [0, 1]))
    with pytest.raises(F.CannonError) as e:
        w.step(1 / 60)
    assert e.value.code == F.E_UNSUPPORTED


def _terrain_spec(solver=None, seed=2, scale=(1, 1, 1)):
    mesh = _grid_mesh(7, 1.0, 0.2)
    mesh.setScale(scale)
    torus = api.Trimesh.createTorus(0.6, 0.25, 6, 5)
    shapes = [mesh, dict(type=F.SHAPE_PLANE), torus, api.Sphere(0.3), api.Sphere(0.45)]
    rng = np.random.default_rng(seed)
    n = 3 + 14
    pos = np.zeros((n, 3), np.float32)
    quat = np.tile(IDENT, (n, 1))
    mass = np.ones(n)
    shape = np.zeros(n, np.int32)
    pos[0], mass[0], shape[0] = (0, 0.5, 0), 0, 0                       # static terrain mesh half a metre above ...
    pos[1], mass[1], shape[1], quat[1] = (0, 0, 0), 0, 1, scenes.GROUND_QUAT  # ... the ground plane
    pos[2], shape[2] = (12, 1.0, 0), 2                                  # a dynamic torus mesh that only meets the plane and the spheres
    q = rng.normal(size=4)
    quat[2] = (q / np.linalg.norm(q)).astype(np.float32)
    for k in range(14):
        pos[3 + k] = (rng.uniform(-2.5, 2.5), rng.uniform(1.0, 3.5), rng.uniform(-2.5, 2.5)) if k < 11 else (12 + 0.3 * (k - 12), 2.0 + 0.8 * (k - 11), 0.1 * (k - 12))
        shape[3 + k] = 3 + (k % 2)
    desc = dict(gravity=(0, -10, 0))
    if solver is not None:
        desc["solver_kind"] = solver
    return SceneSpec(desc=desc, shapes=[s._desc() if hasattr(s, "_desc") else s for s in shapes],
                     bodies=dict(position=pos, quaternion=quat, mass=mass, shape=shape), n_bodies=n, name="spheres on a trimesh terrain")


def test_oracle_spheres_rest_on_the_mesh(oracle_lib):
    w = DeviceWorld(oracle_lib, _terrain_spec())
    seen = 0
    for _ in range(150):
        w.step(1 / 60)
        seen = max(seen, len(w.get_contacts()["body_i"]))
    out = w.get_bodies(("position",))
    assert np.isfinite(out["position"]).all() and seen > 15
    inside = np.abs(out["position"][3:14, [0, 2]]).max(axis=1) < 2.9
    assert (out["position"][3:14, 1][inside] > 0.4).all()  # spheres over the sheet stay on it


@pytest.mark.gpu
@pytest.mark.parametrize("scale", [(1, 1, 1), (1.5, 0.5, 1.25)])
def test_trimesh_staged_parity(cuda_lib, oracle_lib, scale):
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, _terrain_spec(scale=scale))
    parity.assert_same_state(dev, ref, "upload", fields=("bounding_radius", "inv_inertia"))
    seen = 0
    for s in range(120):
        seen = max(seen, parity.staged_step(dev, ref, 1 / 60, f"trimesh step {s}")[1])
    parity.assert_same_state(dev, ref, "aabb", fields=("aabb",))
    assert seen > 15


@pytest.mark.gpu
@pytest.mark.parametrize("solver", [F.SOLVER_REFERENCE_ORDER, F.SOLVER_COLORED])
def test_trimesh_fused_parity(cuda_lib, oracle_lib, solver):
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, _terrain_spec(solver=solver, seed=7))
    for s in range(0, 200, 40):
        dev.step(1 / 60, 40)
        ref.step(1 / 60, 40)
        parity.assert_same_state(dev, ref, f"fused step {s + 40}")


@pytest.mark.gpu
def test_unfinished_trimesh_pairs_are_refused_on_the_device(cuda_lib):
    w = DeviceWorld(cuda_lib, _spec([_grid_mesh(), api.Box((0.3, 0.3, 0.3))], [[0, 0, 0], [0, 0.3, 0]], [0, 1]))
    with pytest.raises(F.CannonError) as e:
        w.step(1 / 60)
    assert e.value.code == F.E_UNSUPPORTED
    w2 = DeviceWorld(cuda_lib, _spec([_grid_mesh(), api.Sphere(0.3)], [[0, 0, 0], [0, 0.3, 0]], [0, 1]))
    with pytest.raises(F.CannonError):
        w2.raycast(np.array([[0, 5, 0]], np.float32), np.array([[0, -5, 0]], np.float32))
