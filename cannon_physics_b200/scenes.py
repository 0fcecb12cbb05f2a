"""Seeded synthetic scenes for the BASELINE.json configurations.

Generator = splitmix64(seed) -> uniform f64 in [0,1) -> cast to f32 at store, so the same scene can be
rebuilt bit-identically anywhere (the reference demos use unseeded ``math.Random()``).
Every generator takes size parameters so the parity tests can run reduced instances of the same recipe.
"""
from __future__ import annotations

import math

import numpy as np

from . import _ffi as F
from .engine import SceneSpec

_MASK = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & _MASK

    def next_u64(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & _MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
        return z ^ (z >> 31)

    def uniform(self, n: int) -> np.ndarray:
        """n uniform f64 in [0,1) (vectorised: the stream is the sequential splitmix64 stream)."""
        idx = np.arange(1, n + 1, dtype=np.uint64)
        with np.errstate(over="ignore"):
            z = np.uint64(self.s) + idx * np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        self.s = (self.s + n * 0x9E3779B97F4A7C15) & _MASK
        return (z >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def quat_from_euler(x: float, y: float, z: float) -> np.ndarray:
    """Quat.setFromEuler, order XYZ (lib/math/quaternion.dart:191-203), rounded to f32 at store."""
    c1, c2, c3 = math.cos(x / 2), math.cos(y / 2), math.cos(z / 2)
    s1, s2, s3 = math.sin(x / 2), math.sin(y / 2), math.sin(z / 2)
    return np.array([s1 * c2 * c3 + c1 * s2 * s3, c1 * s2 * c3 - s1 * c2 * s3, c1 * c2 * s3 + s1 * s2 * c3,
                     c1 * c2 * c3 - s1 * s2 * s3], dtype=np.float32)


GROUND_QUAT = quat_from_euler(-math.pi / 2, 0.0, 0.0)


def _base_bodies(n: int):
    q = np.zeros((n, 4), np.float32)
    q[:, 3] = 1.0
    return {
        "position": np.zeros((n, 3), np.float32), "quaternion": q,
        "velocity": np.zeros((n, 3), np.float32), "angular_velocity": np.zeros((n, 3), np.float32),
        "mass": np.zeros(n, np.float64), "shape": np.zeros(n, np.int32), "material": np.full(n, -1, np.int32),
    }


def _random_unit_quats(rng: SplitMix64, n: int) -> np.ndarray:
    u = rng.uniform(3 * n).reshape(n, 3)
    u1, u2, u3 = u[:, 0], u[:, 1], u[:, 2]
    q = np.stack([np.sqrt(1 - u1) * np.sin(2 * np.pi * u2), np.sqrt(1 - u1) * np.cos(2 * np.pi * u2),
                  np.sqrt(u1) * np.sin(2 * np.pi * u3), np.sqrt(u1) * np.cos(2 * np.pi * u3)], axis=1)
    return q.astype(np.float32)


def spheres_on_plane(nx=10, ny=10, nz=10, seed=1, radius=0.25, spacing=0.6, y0=1.0, jitter=0.05,
                     broadphase=F.BP_NAIVE, iterations=10, solver=F.SOLVER_REFERENCE_ORDER) -> SceneSpec:
    """config 1: nx*ny*nz spheres dropped on a plane (NaiveBroadphase + GSSolver 10 it, dt=1/60)."""
    n = nx * ny * nz + 1
    rng = SplitMix64(seed)
    b = _base_bodies(n)
    b["quaternion"][0] = GROUND_QUAT
    b["shape"][0] = 0
    ii, jj, kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    jit = (rng.uniform(3 * (n - 1)).reshape(n - 1, 3) * 2 - 1) * jitter
    pos = np.stack([(ii.ravel() - (nx - 1) / 2) * spacing, y0 + jj.ravel() * spacing, (kk.ravel() - (nz - 1) / 2) * spacing], axis=1) + jit
    b["position"][1:] = pos.astype(np.float32)
    b["mass"][1:] = 1.0
    b["shape"][1:] = 1
    return SceneSpec(
        desc=dict(gravity=(0, -10, 0), broadphase_kind=broadphase, solver_iterations=iterations, solver_kind=solver),
        shapes=[dict(type=F.SHAPE_PLANE), dict(type=F.SHAPE_SPHERE, radius=radius)],
        bodies=b, n_bodies=n, name=f"c1_spheres_on_plane_{nx}x{ny}x{nz}")


def box_stacks(n_stacks=250, height=20, seed=2, he=0.5, grid=16, pitch=2.0, gap=0.02, jitter=0.01,
               broadphase=F.BP_SAP, iterations=20, solver=F.SOLVER_REFERENCE_ORDER) -> SceneSpec:
    """config 2: `height`-high box stacks with friction/restitution, SAPBroadphase axis x, GSSolver 20 it."""
    n = n_stacks * height + 1
    rng = SplitMix64(seed)
    b = _base_bodies(n)
    b["quaternion"][0] = GROUND_QUAT
    b["shape"][0] = 0
    b["material"][:] = 0
    s = np.arange(n_stacks)
    gx, gz = s % grid, s // grid
    jit = (rng.uniform(2 * n_stacks * height).reshape(n_stacks, height, 2) * 2 - 1) * jitter
    k = np.arange(height)
    px = (gx[:, None] - (grid - 1) / 2) * pitch + jit[:, :, 0]
    pz = (gz[:, None] - (grid - 1) / 2) * pitch + jit[:, :, 1]
    py = np.broadcast_to(he + k[None, :] * (2 * he + gap), (n_stacks, height))
    b["position"][1:] = np.stack([px.ravel(), py.ravel(), pz.ravel()], axis=1).astype(np.float32)
    b["mass"][1:] = 1.0
    b["shape"][1:] = 1
    return SceneSpec(
        desc=dict(gravity=(0, -10, 0), broadphase_kind=broadphase, sap_axis=0, solver_iterations=iterations, solver_kind=solver),
        shapes=[dict(type=F.SHAPE_PLANE), dict(type=F.SHAPE_BOX, half_extents=(he, he, he))],
        bodies=b, n_bodies=n,
        material_friction=np.array([-1.0]), material_restitution=np.array([-1.0]),
        contact_materials=[dict(material_a=0, material_b=0, friction=0.3, restitution=0.2)],
        name=f"c2_box_stacks_{n_stacks}x{height}")


def heightfield_data(nsamp: int) -> np.ndarray:
    """h(i,j) = cos(2*pi*i/n)*cos(2*pi*j/n)+2, border rows/cols = 3 (examples/lib/examples/heightfield.dart:46-60)."""
    i = np.arange(nsamp, dtype=np.float64)
    h = np.cos(2 * np.pi * i[:, None] / nsamp) * np.cos(2 * np.pi * i[None, :] / nsamp) + 2.0
    h[0, :] = 3.0
    h[-1, :] = 3.0
    h[:, 0] = 3.0
    h[:, -1] = 3.0
    return h


def mixed_pile_on_heightfield(nx=100, nz=100, layers=10, seed=3, hf_samples=257, pitch=2.4, layer_pitch=0.8,
                              iterations=10, solver=F.SOLVER_COLORED, grid_cells=(128, 16, 128), with_heightfield=True,
                              kinds=("sphere", "box", "cylinder")) -> SceneSpec:
    """config 3: nx*nz*layers mixed sphere/box/cylinder pile on a heightfield, GridBroadphase, 10 it.

    The lattice is centred over the heightfield (which spans hf_samples-1 units); pitch is shrunk when
    the requested lattice would not fit so reduced instances keep every body above terrain.
    """
    n_dyn = nx * nz * layers
    n = n_dyn + 1
    rng = SplitMix64(seed)
    b = _base_bodies(n)
    size = float(hf_samples - 1)
    hf = heightfield_data(hf_samples)
    half = size / 2.0
    if with_heightfield:
        # body at (-size/2, -4, size/2), rotated -pi/2 about x (examples/lib/examples/heightfield.dart:69-74)
        b["position"][0] = (-half, -4.0, half)
        b["quaternion"][0] = GROUND_QUAT
        shapes = [dict(type=F.SHAPE_HEIGHTFIELD, hf_data=hf, hf_element_size=1)]
    else:
        b["position"][0] = (0.0, -1.0, 0.0)
        b["quaternion"][0] = GROUND_QUAT
        shapes = [dict(type=F.SHAPE_PLANE)]
    b["shape"][0] = 0
    shape_ids = {}
    for kname in kinds:
        shape_ids[kname] = len(shapes)
        if kname == "sphere":
            shapes.append(dict(type=F.SHAPE_SPHERE, radius=0.25))
        elif kname == "box":
            shapes.append(dict(type=F.SHAPE_BOX, half_extents=(0.25, 0.25, 0.25)))
        else:
            shapes.append(dict(type=F.SHAPE_CYLINDER, radius_top=0.25, radius_bottom=0.25, height=0.5, num_segments=8))
    pitch = min(pitch, (size - 4.0) / max(nx, nz))
    ii, kk, ll = np.meshgrid(np.arange(nx), np.arange(nz), np.arange(layers), indexing="ij")
    x = (ii.ravel() - (nx - 1) / 2) * pitch
    z = (kk.ravel() - (nz - 1) / 2) * pitch
    if with_heightfield:
        # local heightfield coords: lx = x + half, ly = half - z (rotation -pi/2 about x maps local y -> -z)
        lx = np.clip(np.rint(x + half).astype(int), 0, hf_samples - 1)
        ly = np.clip(np.rint(half - z).astype(int), 0, hf_samples - 1)
        ground = hf[lx, ly] - 4.0
    else:
        ground = np.full(n_dyn, -1.0)
    y = ground + 1.0 + ll.ravel() * layer_pitch
    jit = (rng.uniform(2 * n_dyn).reshape(n_dyn, 2) * 2 - 1) * 0.05
    b["position"][1:] = np.stack([x + jit[:, 0], y, z + jit[:, 1]], axis=1).astype(np.float32)
    kind_idx = np.arange(n_dyn) % len(kinds)
    sid = np.array([shape_ids[k] for k in kinds], dtype=np.int32)[kind_idx]
    b["shape"][1:] = sid
    quats = _random_unit_quats(rng, n_dyn)
    is_sphere = np.array([k == "sphere" for k in kinds])[kind_idx]
    quats[is_sphere] = (0, 0, 0, 1)
    b["quaternion"][1:] = quats
    b["mass"][1:] = 1.0
    lo = b["position"][1:].min(axis=0) - 1.0
    hi = b["position"][1:].max(axis=0) + 1.0
    lo[1] = min(lo[1], -6.0)
    return SceneSpec(
        desc=dict(gravity=(0, -10, 0), broadphase_kind=F.BP_GRID, grid_min=lo, grid_max=hi, grid_nx=grid_cells[0],
                  grid_ny=grid_cells[1], grid_nz=grid_cells[2], solver_iterations=iterations, solver_kind=solver),
        shapes=shapes, bodies=b, n_bodies=n, name=f"c3_mixed_pile_{nx}x{nz}x{layers}")


def chain_worlds(n_worlds=4096, chains=7, links=9, seed=4, iterations=10, solver=F.SOLVER_REFERENCE_ORDER, top_y=None) -> SceneSpec:
    """config 4: n_worlds independent worlds of 1 plane + chains*links boxes hanging as jointed chains.

    Links are joined alternately by two corner PointToPointConstraints (examples/lib/examples/constraints.dart:135-146)
    and a HingeConstraint (axis x); the top link of each chain has mass 0. Seed is seed + world index.
    """
    per = 1 + chains * links
    n = n_worlds * per
    b = _base_bodies(n)
    b["world_id"] = np.repeat(np.arange(n_worlds, dtype=np.int32), per)
    hx, hy, hz = 0.25, 0.25, 0.05
    space = 0.1 * hy
    # hang low enough that the last links of every chain rest on the ground plane (joint rows + contacts)
    if top_y is None:
        top_y = (links - 2) * (2 * hy + 2 * space)
    cons = []
    for w in range(n_worlds):
        rng = SplitMix64(seed + w)
        base = w * per
        b["quaternion"][base] = GROUND_QUAT
        b["shape"][base] = 0
        anchors = (rng.uniform(2 * chains).reshape(chains, 2) * 2 - 1) * 0.5
        for c in range(chains):
            ax = (c - (chains - 1) / 2) * 1.5 + anchors[c, 0]
            az = anchors[c, 1]
            prev = -1
            for l in range(links):
                idx = base + 1 + c * links + l
                b["position"][idx] = (ax, top_y - l * (2 * hy + 2 * space), az)
                b["mass"][idx] = 0.0 if l == 0 else 0.3
                b["shape"][idx] = 1
                if l > 0:
                    if l % 2 == 1:
                        cons.append(dict(type=F.CONSTRAINT_POINT_TO_POINT, body_a=idx, body_b=prev,
                                         pivot_a=(hx, hy + space, 0), pivot_b=(hx, -hy - space, 0)))
                        cons.append(dict(type=F.CONSTRAINT_POINT_TO_POINT, body_a=idx, body_b=prev,
                                         pivot_a=(-hx, hy + space, 0), pivot_b=(-hx, -hy - space, 0)))
                    else:
                        cons.append(dict(type=F.CONSTRAINT_HINGE, body_a=idx, body_b=prev,
                                         pivot_a=(0, hy + space, 0), pivot_b=(0, -hy - space, 0),
                                         axis_a=(1, 0, 0), axis_b=(1, 0, 0)))
                prev = idx
    return SceneSpec(
        desc=dict(gravity=(0, -10, 0), broadphase_kind=F.BP_NAIVE, solver_iterations=iterations, solver_kind=solver,
                  n_worlds=n_worlds),
        shapes=[dict(type=F.SHAPE_PLANE), dict(type=F.SHAPE_BOX, half_extents=(hx, hy, hz))],
        bodies=b, n_bodies=n, constraints=cons, name=f"c4_chain_worlds_{n_worlds}x{per}")


def constraint_zoo(seed=7, iterations=10, solver=F.SOLVER_REFERENCE_ORDER, groups=4) -> SceneSpec:
    """every joint type of lib/constraints over a ground plane. Each group holds
      * a DistanceConstraint pendulum (static anchor sphere + dynamic sphere; second pendulum with the default distance),
      * two boxes under a LockConstraint (identity poses) and two more with rotated poses (Body.vectorToLocalFrame quirk),
      * a ragdoll-like limb: static box + three boxes joined by ConeTwistConstraints (examples/lib/examples/ragdoll.dart:310-404),
      * a PointToPoint + Hinge pair, so the old joints run next to the new ones,
      * three Springs applied in the postStep slot (lib/objects/spring.dart).
    """
    per = 2 + 2 + 4 + 4 + 3
    n = 1 + groups * per
    rng = SplitMix64(seed)
    b = _base_bodies(n)
    b["quaternion"][0] = GROUND_QUAT
    b["shape"][0] = 0
    cons = []
    springs = []
    hx = 0.25
    for g in range(groups):
        o = 1 + g * per
        gx = (g - (groups - 1) / 2) * 6.0
        jit = (rng.uniform(8) * 2 - 1) * 0.1
        # distance pendulums
        b["position"][o] = (gx, 5.0, 0); b["shape"][o] = 2; b["mass"][o] = 0.0
        b["position"][o + 1] = (gx + 1.5 + jit[0], 4.0, jit[1]); b["shape"][o + 1] = 2; b["mass"][o + 1] = 1.0
        cons.append(dict(type=F.CONSTRAINT_DISTANCE, body_a=o, body_b=o + 1, distance=1.5))
        b["position"][o + 2] = (gx, 5.0, 2.0); b["shape"][o + 2] = 2; b["mass"][o + 2] = 0.0
        b["position"][o + 3] = (gx + 1.0, 4.5 + jit[2], 2.0); b["shape"][o + 3] = 2; b["mass"][o + 3] = 0.7
        cons.append(dict(type=F.CONSTRAINT_DISTANCE, body_a=o + 3, body_b=o + 2, max_force=5e4))  # default distance
        # locks: axis-aligned pair, then a pair with rotated poses
        b["position"][o + 4] = (gx - 1.0, 1.5, -2.0); b["shape"][o + 4] = 1; b["mass"][o + 4] = 1.0
        b["position"][o + 5] = (gx - 0.4, 1.6, -2.0); b["shape"][o + 5] = 1; b["mass"][o + 5] = 2.0
        cons.append(dict(type=F.CONSTRAINT_LOCK, body_a=o + 4, body_b=o + 5))
        b["position"][o + 6] = (gx + 1.0, 1.5, -2.0); b["shape"][o + 6] = 1; b["mass"][o + 6] = 1.0
        b["quaternion"][o + 6] = quat_from_euler(0.3 + jit[3], -0.2, 0.5)
        b["position"][o + 7] = (gx + 1.6, 1.7, -2.1); b["shape"][o + 7] = 1; b["mass"][o + 7] = 1.0
        b["quaternion"][o + 7] = quat_from_euler(-0.4, 0.6 + jit[4], 0.1)
        cons.append(dict(type=F.CONSTRAINT_LOCK, body_a=o + 6, body_b=o + 7, max_force=1e5))
        # cone-twist limb hanging from a static box
        b["position"][o + 8] = (gx, 4.0, -4.0); b["shape"][o + 8] = 1; b["mass"][o + 8] = 0.0
        for k in range(3):
            idx = o + 9 + k
            b["position"][idx] = (gx + 0.05 * k + jit[5] * k, 4.0 - 0.7 * (k + 1), -4.0)
            b["shape"][idx] = 1
            b["mass"][idx] = 0.8
            cons.append(dict(type=F.CONSTRAINT_CONE_TWIST, body_a=idx - 1, body_b=idx, pivot_a=(0, -0.35, 0), pivot_b=(0, 0.35, 0),
                             axis_a=(0, 1, 0), axis_b=(0, 1, 0), angle=math.pi / 4 if k < 2 else math.pi / 8, twist_angle=math.pi / 8,
                             collide_connected=0 if k == 1 else 1))
        b["velocity"][o + 11] = (1.5, 0, 0.8 + jit[6])  # kick the last limb so the cone limits come into play
        # old joints
        b["position"][o + 12] = (gx, 3.0, 4.0); b["shape"][o + 12] = 1; b["mass"][o + 12] = 0.0
        b["position"][o + 13] = (gx, 2.3, 4.0); b["shape"][o + 13] = 1; b["mass"][o + 13] = 0.5
        b["position"][o + 14] = (gx, 1.6, 4.0 + jit[7]); b["shape"][o + 14] = 1; b["mass"][o + 14] = 0.5
        cons.append(dict(type=F.CONSTRAINT_POINT_TO_POINT, body_a=o + 13, body_b=o + 12, pivot_a=(0, 0.35, 0), pivot_b=(0, -0.35, 0)))
        cons.append(dict(type=F.CONSTRAINT_HINGE, body_a=o + 14, body_b=o + 13, pivot_a=(0, 0.35, 0), pivot_b=(0, -0.35, 0),
                         axis_a=(1, 0, 0), axis_b=(1, 0, 0)))
        # springs (lib/objects/spring.dart, examples/lib/examples/spring.dart): the p2p link is also tied to its anchor by an
        # off-centre spring, the two locked pairs are tied together, and one body carries two springs (accumulation order)
        springs.append(dict(body_a=o + 12, body_b=o + 14, rest_length=1.0, stiffness=50.0, damping=1.0, local_anchor_a=(0.25, 0, 0),
                            local_anchor_b=(-0.25, 0.25, 0)))
        springs.append(dict(body_a=o + 5, body_b=o + 6, rest_length=1.2, stiffness=80.0, damping=2.0, local_anchor_b=(0, 0.25, 0)))
        springs.append(dict(body_a=o + 11, body_b=o + 6, rest_length=2.0, stiffness=20.0, damping=0.5))
    return SceneSpec(
        desc=dict(gravity=(0, -10, 0), broadphase_kind=F.BP_NAIVE, solver_iterations=iterations, solver_kind=solver),
        shapes=[dict(type=F.SHAPE_PLANE), dict(type=F.SHAPE_BOX, half_extents=(hx, hx, hx)), dict(type=F.SHAPE_SPHERE, radius=0.2)],
        bodies=b, n_bodies=n, constraints=cons, springs=springs, name=f"constraint_zoo_{groups}")


def sphere_container(nx=160, nz=160, ny=40, n_spheres=None, seed=5, radius=0.25, pitch=0.6, extent=100.0,
                     iterations=10, solver=F.SOLVER_COLORED, allow_sleep=True, broadphase=F.BP_NAIVE) -> SceneSpec:
    """config 5: granular sphere pile in a 5-plane container (floor + 4 walls, container.dart:62-100), sleeping on."""
    n_dyn = n_spheres if n_spheres is not None else nx * ny * nz
    n = n_dyn + 5
    rng = SplitMix64(seed)
    b = _base_bodies(n)
    half = extent / 2
    b["quaternion"][0] = GROUND_QUAT
    b["quaternion"][1] = quat_from_euler(0, math.pi / 2, 0)
    b["position"][1] = (-half, 0, 0)
    b["quaternion"][2] = quat_from_euler(0, -math.pi / 2, 0)
    b["position"][2] = (half, 0, 0)
    b["quaternion"][3] = quat_from_euler(0, 0, 0)
    b["position"][3] = (0, 0, -half)
    b["quaternion"][4] = quat_from_euler(0, math.pi, 0)
    b["position"][4] = (0, 0, half)
    b["shape"][:5] = 0
    site = np.arange(n_dyn)
    ix, iz, iy = site % nx, (site // nx) % nz, site // (nx * nz)
    jit = (rng.uniform(3 * n_dyn).reshape(n_dyn, 3) * 2 - 1) * 0.05
    pos = np.stack([(ix - (nx - 1) / 2) * pitch, radius + 0.1 + iy * pitch, (iz - (nz - 1) / 2) * pitch], axis=1) + jit
    b["position"][5:] = pos.astype(np.float32)
    b["mass"][5:] = 1.0
    b["shape"][5:] = 1
    return SceneSpec(
        desc=dict(gravity=(0, -10, 0), broadphase_kind=broadphase, solver_iterations=iterations, solver_kind=solver,
                  allow_sleep=1 if allow_sleep else 0),
        shapes=[dict(type=F.SHAPE_PLANE), dict(type=F.SHAPE_SPHERE, radius=radius)],
        bodies=b, n_bodies=n, name=f"c5_sphere_container_{n_dyn}")
