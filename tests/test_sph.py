"""SPHSystem, lib/objects/sph_system.dart: a World.subsystem updated after gravity and before the
broadphase (world_class.dart:472-475). Reproduced as written, including pressures[j] / densities[j] indexed by the
position in the neighbour list (:131-133,142). CPU: known answers computed here in float64 from the formulas of the
source; GPU: bit-exact parity with the oracle."""
import math

import numpy as np
import pytest

import parity
from cannon_physics_b200 import _ffi as F
from cannon_physics_b200 import api, scenes
from cannon_physics_b200.engine import DeviceWorld, SceneSpec

IDENT = np.array([0, 0, 0, 1], np.float32)


def _w(h, r):
    return (315.0 / (64.0 * math.pi * h ** 9)) * (h * h - r * r) ** 3


def test_two_particles_density_pressure_and_forces(oracle_lib):
    # two unit-mass particles 0.5 apart, h = 1, no gravity, at rest: the force is the pressure term only
    h, cs, rho0, eps, m = 1.0, 2.0, 1.5, 1e-5, 1.0
    spec = SceneSpec(desc=dict(gravity=(0, 0, 0)), shapes=[api.Particle()._desc()],
                     bodies=dict(position=np.array([[0, 0, 0], [0.5, 0, 0]], np.float32), mass=np.array([m, m]), shape=np.zeros(2, np.int32)), n_bodies=2,
                     sph_systems=[dict(particles=[0, 1], density=rho0, smoothing_radius=h, speed_of_sound=cs)])
    w = DeviceWorld(oracle_lib, spec)
    w.set_dt(1 / 60)
    w.apply_gravity()  # gravity (none) + subsystems
    f = w.get_bodies(("force",))["force"]
    dens = m * _w(h, 0.5) + m * _w(h, 0.0)
    pres = cs * cs * (dens - rho0)
    pij = -m * (pres / (dens * dens + eps) + pres / (dens * dens + eps))
    grad = 945.0 / (32.0 * math.pi * h ** 9) * (h * h - 0.25) ** 2
    # particle 0: rVec = p0 - p1 = (-0.5, 0, 0); the self term has rVec = 0
    fx0 = m * (pij * grad * -0.5)
    np.testing.assert_allclose(f[0], [fx0, 0, 0], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(f[1], [-fx0, 0, 0], rtol=1e-6, atol=1e-9)
    assert f[0][0] * np.sign(pres) > 0 or pres == 0  # over-dense -> pushed apart (particle 0 towards -x when pres < 0 ... sign follows pij)


def _block_spec(nx=6, ny=5, nz=6, solver=None, seed=1, h=0.6):
    """A block of SPH particles dropped into a box of planes, plus a second small system and a few rigid spheres."""
    rng = np.random.default_rng(seed)
    shapes = [dict(type=F.SHAPE_PLANE), api.Particle()._desc(), api.Sphere(0.25)._desc()]
    pos, quat, mass, shape = [[0, 0, 0]], [scenes.GROUND_QUAT], [0.0], [0]
    s = math.sin(math.pi / 4)
    for p, q in (((-1.6, 0, 0), (0, s, 0, s)), ((1.6, 0, 0), (0, -s, 0, s)), ((0, 0, -1.6), (0, 0, 0, 1)), ((0, 0, 1.6), (0, 1, 0, 0))):
        pos.append(list(p)); quat.append(q); mass.append(0.0); shape.append(0)
    first = len(pos)
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                pos.append([0.3 * (i - (nx - 1) / 2) + 0.02 * rng.random(), 0.4 + 0.3 * j, 0.3 * (k - (nz - 1) / 2) + 0.02 * rng.random()])
                quat.append(IDENT); mass.append(0.02); shape.append(1)
    n_fluid = len(pos) - first
    for k in range(4):
        pos.append([0.5 * k - 0.75, 2.6, 0.1 * k]); quat.append(IDENT); mass.append(0.5); shape.append(2)
    n = len(pos)
    desc = dict(gravity=(0, -10, 0))
    if solver is not None:
        desc["solver_kind"] = solver
    fluid = list(range(first, first + n_fluid))
    return SceneSpec(desc=desc, shapes=shapes, bodies=dict(position=np.array(pos, np.float32), quaternion=np.array(quat, np.float32), mass=np.array(mass), shape=np.array(shape, np.int32),
                                                          linear_damping=np.full(n, 0.1)), n_bodies=n,
                     sph_systems=[dict(particles=fluid[: n_fluid - 20], density=1.0, smoothing_radius=h, speed_of_sound=0.5, viscosity=0.03),
                                  dict(particles=fluid[n_fluid - 20:][::-1], density=0.5, smoothing_radius=0.5, speed_of_sound=1.0)], name="sph block")


def test_oracle_fluid_block_stays_finite_and_spreads(oracle_lib):
    w = DeviceWorld(oracle_lib, _block_spec(4, 3, 4))
    p0 = w.get_bodies(("position",))["position"].copy()
    for _ in range(60):
        w.step(1 / 60)
    p1 = w.get_bodies(("position",))["position"]
    assert np.isfinite(p1).all()
    assert np.abs(p1[5:] - p0[5:]).max() > 0.05


def test_api_subsystems_list(oracle_lib):
    world = api.World(gravity=(0, 0, 0), _lib=oracle_lib)
    sph = api.SPHSystem()
    sph.smoothingRadius, sph.density = 1.0, 1.5
    a = api.Body(mass=1, shape=api.Particle(), position=(0, 0, 0))
    b = api.Body(mass=1, shape=api.Particle(), position=(0.5, 0, 0))
    for body in (a, b):
        world.addBody(body)
        sph.add(body)
    world.subsystems.append(sph)
    world.step(1 / 60)
    assert a.velocity[0] != 0 and abs(a.velocity[0] + b.velocity[0]) < 1e-7  # equal and opposite pressure forces


@pytest.mark.gpu
def test_sph_forces_parity(cuda_lib, oracle_lib):
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, _block_spec())
    for w in (dev, ref):
        w.set_dt(1 / 60)
        w.apply_gravity()
    parity.assert_same_state(dev, ref, "forces after gravity + SPH", fields=("force",))
    assert np.abs(dev.get_bodies(("force",))["force"][5:185]).max() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("solver", [F.SOLVER_REFERENCE_ORDER, F.SOLVER_COLORED])
def test_sph_fused_parity(cuda_lib, oracle_lib, solver):
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, _block_spec(solver=solver, seed=3))
    for s in range(0, 120, 30):
        dev.step(1 / 60, 30)
        ref.step(1 / 60, 30)
        parity.assert_same_state(dev, ref, f"sph step {s + 30}")
