"""GPU parity tests proper: the CUDA library (through the C ABI) against the CPU oracle and the committed
golden fixtures. Bar: bit-exact pair lists, per-pair contact counts, contact geometry, solver rows and body
state in REFERENCE_ORDER mode (integer / index work and f32-stored f64 arithmetic are reproduced exactly, so
the 1e-5 relative tolerance north_star allows for one-step state is met with tolerance 0)."""
import os

import numpy as np
import pytest

import parity
from cannon_physics_b200 import _ffi as F
from cannon_physics_b200 import engine, scenes

pytestmark = pytest.mark.gpu
REF = F.SOLVER_REFERENCE_ORDER
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_native_library_is_the_cuda_backend(cuda_lib):
    assert cuda_lib.cannon_backend() == b"cuda"
    assert os.path.basename(cuda_lib._path) == "libcannon_cuda.so"


STAGED = {
    "c1 spheres on plane, Naive": (lambda: scenes.spheres_on_plane(4, 4, 4), 90),
    "c1 dense spheres, bounding boxes": (lambda: _with(scenes.spheres_on_plane(4, 4, 4, spacing=0.45), use_bounding_boxes=1), 40),
    "c2 box stacks, SAP x": (lambda: scenes.box_stacks(4, 5, grid=2), 60),
    "c2 box stacks, SAP z": (lambda: _with(scenes.box_stacks(6, 4, grid=3), sap_axis=2), 40),
    "c3 mixed pile on plane, Grid": (lambda: scenes.mixed_pile_on_heightfield(4, 4, 3, with_heightfield=False, solver=REF, grid_cells=(8, 4, 8)), 90),
    "c3 mixed pile on heightfield, Grid": (lambda: scenes.mixed_pile_on_heightfield(4, 4, 3, hf_samples=33, solver=REF, grid_cells=(8, 4, 8)), 90),
    "c4 jointed chain worlds (batch)": (lambda: scenes.chain_worlds(3, chains=2, links=5), 80),
    "c5 container with sleeping": (lambda: scenes.sphere_container(5, 5, 3, extent=4.0, solver=REF), 150),
    "8f joints: distance, lock, cone-twist": (lambda: scenes.constraint_zoo(groups=2), 150),
    # bodies fall asleep while their springs still see a relative velocity: sleepTick must come after the postStep slot
    # (world_class.dart:685-699), or the spring forces of the next step differ
    "springs + sleeping (sleepTick after postStep)": (lambda: _sleepy(scenes.constraint_zoo(groups=2)), 120),
}


def _sleepy(spec):
    spec.desc.update(allow_sleep=1)
    n = spec.n_bodies
    spec.bodies["allow_sleep"] = np.ones(n, np.uint8)
    spec.bodies["sleep_speed_limit"] = np.full(n, 1.5)
    spec.bodies["sleep_time_limit"] = np.full(n, 0.1)
    return spec


SPLIT = {
    "split: spheres on plane": (lambda: _with(scenes.spheres_on_plane(4, 4, 3), solver_kind=F.SOLVER_SPLIT), 80),
    "split: box stacks SAP": (lambda: _with(scenes.box_stacks(6, 4, grid=3), solver_kind=F.SOLVER_SPLIT), 60),
    "split: jointed chains": (lambda: _with(scenes.chain_worlds(1, chains=3, links=5), solver_kind=F.SOLVER_SPLIT), 80),
    "split: sleeping container": (lambda: _with(scenes.sphere_container(5, 5, 3, extent=4.0, solver=REF), solver_kind=F.SOLVER_SPLIT), 150),
    "split: distance, lock, cone-twist": (lambda: _with(scenes.constraint_zoo(groups=2), solver_kind=F.SOLVER_SPLIT), 100),
}


@pytest.mark.parametrize("name", list(SPLIT))
def test_split_solver_parity(cuda_lib, oracle_lib, name):
    # SplitSolver mode: islands + per-island GS in descending creation-id order; same arithmetic => bit-exact vs oracle
    mk, steps = SPLIT[name]
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, mk())
    for s in range(steps):
        parity.staged_step(dev, ref, 1 / 60, f"{name} step {s}")
        assert dev.profile()["n_islands"] == ref.profile()["n_islands"], f"{name} step {s}: island count"
    assert dev.profile()["n_islands"] > 0


def _with(spec, **desc):
    spec.desc.update(desc)
    return spec


@pytest.mark.parametrize("name", list(STAGED))
def test_staged_parity_every_stage_every_step(cuda_lib, oracle_lib, name):
    mk, steps = STAGED[name]
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, mk())
    seen = np.zeros(3, np.int64)
    for s in range(steps):
        seen = np.maximum(seen, parity.staged_step(dev, ref, 1 / 60, f"{name} step {s}"))
    assert seen[0] > 0 and seen[2] > 0, "scene never produced pairs / rows"


def _colored(mk):
    return lambda: _with(mk(), solver_kind=F.SOLVER_COLORED)


@pytest.mark.parametrize("name", list(STAGED))
def test_colored_solver_staged_parity(cuda_lib, oracle_lib, name):
    """The throughput solver (COLORED) is the path bench.py times. It performs GSSolver's per-row arithmetic in the colour
    order that include/cannon_cuda.h specifies; the oracle restates that order sequentially, so pairs, contacts, rows
    (B, invC, lambda, colour) and body state must agree bit for bit after every stage of every step."""
    mk, steps = STAGED[name]
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, _colored(mk)())
    seen = np.zeros(3, np.int64)
    for s in range(steps):
        seen = np.maximum(seen, parity.staged_step(dev, ref, 1 / 60, f"colored {name} step {s}", compare_levels=True))
    assert seen[0] > 0 and seen[2] > 0, "scene never produced pairs / rows"


FUSED = {
    "c1 600 steps": (lambda: scenes.spheres_on_plane(5, 5, 5), 600, 50),
    "c2 stacks": (lambda: scenes.box_stacks(9, 6, grid=3), 120, 20),
    "c3 heightfield": (lambda: scenes.mixed_pile_on_heightfield(6, 6, 3, hf_samples=33, solver=REF, grid_cells=(8, 4, 8)), 120, 20),
    # 768 bodies settling on the terrain: every sphere / box / cylinder meets several pillars (pruned SAT axes, pillar
    # table, conservative sphere-pillar rejection: all must leave pairs, contacts, rows and state bit-identical)
    "c3 heightfield, 768 bodies": (lambda: scenes.mixed_pile_on_heightfield(16, 16, 3, hf_samples=65, solver=REF, grid_cells=(16, 4, 16)), 150, 50),
    "c4 batch": (lambda: scenes.chain_worlds(8, chains=3, links=6), 120, 20),
    "c5 sleeping": (lambda: scenes.sphere_container(6, 6, 4, extent=5.0, solver=REF), 240, 40),
    "8f joints": (lambda: scenes.constraint_zoo(groups=3), 240, 30),
}


@pytest.mark.parametrize("name", list(FUSED))
def test_fused_step_parity(cuda_lib, oracle_lib, name):
    mk, steps, chunk = FUSED[name]
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, mk())
    for s in range(0, steps, chunk):
        dev.step(1 / 60, chunk)  # `chunk` steps enqueued without host round trips
        ref.step(1 / 60, chunk)
        parity.assert_same_state(dev, ref, f"{name} after {s + chunk} steps")
        pa, pb = dev.profile(), ref.profile()
        assert (pa["n_pairs"], pa["n_contacts"], pa["n_rows"], pa["iterations_done"]) == (pb["n_pairs"], pb["n_contacts"], pb["n_rows"], pb["iterations_done"])
    assert dev.get_time() == ref.get_time()


@pytest.mark.parametrize("name", list(FUSED))
def test_colored_solver_fused_step_parity(cuda_lib, oracle_lib, name):
    # the same through the fused device-resident step (graph replay) in COLORED mode: bit-exact against the oracle
    mk, steps, chunk = FUSED[name]
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, _colored(mk)())
    for s in range(0, steps, chunk):
        dev.step(1 / 60, chunk)
        ref.step(1 / 60, chunk)
        parity.assert_same_state(dev, ref, f"colored {name} after {s + chunk} steps")
        pa, pb = dev.profile(), ref.profile()
        assert (pa["n_pairs"], pa["n_contacts"], pa["n_rows"], pa["iterations_done"], pa["n_levels"]) == \
               (pb["n_pairs"], pb["n_contacts"], pb["n_rows"], pb["iterations_done"], pb["n_levels"])


# BASELINE configurations at their stated sizes (c3 at 1/16 of the lattice on the full 257-sample field: the oracle needs
# ~7 s per step at 100k bodies). REF = the reference's insertion order, COLORED = the solver bench.py times.
FULL = {
    "c1 full: 1000 spheres on a plane, Naive, 600 steps": (lambda: scenes.spheres_on_plane(10, 10, 10), 600, 100),
    "c2 full: 250 x 20 box stacks, SAP, 20 it, 30 steps": (lambda: scenes.box_stacks(250, 20), 30, 10),
    "c3: 6250 bodies on the 257-sample heightfield, Grid 128x16x128, 150 steps": (lambda: scenes.mixed_pile_on_heightfield(25, 25, 10, solver=REF), 150, 50),
    "c4: 64 worlds x 64 bodies, P2P + hinge chains, 120 steps": (lambda: scenes.chain_worlds(64), 120, 40),
}


@pytest.mark.parametrize("kind", ["reference", "colored"])
@pytest.mark.parametrize("name", list(FULL))
def test_full_size_configs_fused_parity(cuda_lib, oracle_lib, name, kind):
    mk, steps, chunk = FULL[name]
    spec = _with(mk(), solver_kind=REF if kind == "reference" else F.SOLVER_COLORED)
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, spec)
    most = 0
    for s in range(0, steps, chunk):
        dev.step(1 / 60, chunk)
        ref.step(1 / 60, chunk)
        parity.assert_same_state(dev, ref, f"{name} [{kind}] after {s + chunk} steps")
        pa, pb = dev.profile(), ref.profile()
        assert (pa["n_pairs"], pa["n_contacts"], pa["n_rows"], pa["iterations_done"]) == (pb["n_pairs"], pb["n_contacts"], pb["n_rows"], pb["iterations_done"])
        most = max(most, pa["n_contacts"])
    assert most > 0
    # the refitted AABBs themselves (Body.updateAABB, rigid_body.dart:415-447), not only the pair lists they lead to
    a, b = dev.get_bodies(("aabb",)), ref.get_bodies(("aabb",))
    assert np.array_equal(a["aabb"], b["aabb"]), f"{name}: AABB arrays differ"


@pytest.mark.parametrize("opts", [dict(quat_normalize_fast=1), dict(quat_normalize_skip=3), dict(quat_normalize_fast=1, quat_normalize_skip=1),
                                  dict(has_friction_gravity=1, friction_gravity=(0, -2, 0)), dict(has_friction_gravity=1, friction_gravity=(3, 0, -4))])
def test_world_options_quat_normalize_and_friction_gravity(cuda_lib, oracle_lib, opts):
    # World(quatNormalizeFast / quatNormalizeSkip / frictionGravity): world_class.dart:135-144,668, quaternion.dart:171-185,
    # narrow_phase.dart:547-550
    spec = _with(scenes.mixed_pile_on_heightfield(5, 5, 3, with_heightfield=False, solver=REF, grid_cells=(8, 4, 8)), **opts)
    spec.bodies["angular_velocity"][1:] = (0.7, -1.3, 0.4)  # tumbling boxes and cylinders: the normalisation branch matters
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, spec)
    for s in range(90):
        parity.staged_step(dev, ref, 1 / 60, f"{opts} step {s}")
    a, b = dev.get_bodies(("aabb",)), ref.get_bodies(("aabb",))
    assert np.array_equal(a["aabb"], b["aabb"])


@pytest.mark.parametrize("name", ["c1_small", "c2_small", "c3_plane_small", "c3_hf_small", "c4_small", "c5_small", "joints_small",
                                  "c2_colored_small", "c3_hf_colored_small", "c4_colored_small", "c2_quatfast_small",
                                  "hulls_small", "particles_small", "compound_small", "compound_colored_small", "trimesh_small", "sph_small"])
def test_cuda_matches_golden_fixture(cuda_lib, name):
    from make_golden import CASES, run_case
    got = run_case(cuda_lib, *CASES[name])
    ref = np.load(os.path.join(GOLDEN, name + ".npz"))
    for k in ref.files:
        assert np.array_equal(got[k], ref[k]), (name, k)


def test_edge_cases_empty_and_degenerate_worlds(cuda_lib, oracle_lib):
    from cannon_physics_b200.engine import SceneSpec
    # empty world, a world with one shapeless body, two coincident spheres (zero-length normal)
    for bodies, n, shapes in [
        ({}, 0, []),
        ({"mass": np.array([1.0]), "shape": np.array([-1], np.int32)}, 1, []),
        ({"mass": np.array([1.0, 1.0]), "shape": np.array([0, 0], np.int32)}, 2, [dict(type=F.SHAPE_SPHERE, radius=0.5)]),
    ]:
        spec = SceneSpec(desc=dict(gravity=(0, -10, 0)), shapes=shapes, bodies=bodies, n_bodies=n)
        dev, ref = parity.make_pair(cuda_lib, oracle_lib, spec)
        for s in range(3):
            parity.staged_step(dev, ref, 1 / 60, f"edge n={n} step {s}")


def test_triggers_disabled_response_masks_and_kinematic(cuda_lib, oracle_lib):
    spec = scenes.spheres_on_plane(3, 3, 2, spacing=0.45, y0=0.3)
    n = spec.n_bodies
    b = spec.bodies
    b["is_trigger"] = np.zeros(n, np.uint8); b["is_trigger"][3] = 1
    b["collision_response"] = np.ones(n, np.uint8); b["collision_response"][4] = 0
    b["collision_filter_group"] = np.ones(n, np.int32); b["collision_filter_group"][5] = 2
    b["collision_filter_mask"] = np.full(n, -1, np.int32); b["collision_filter_mask"][6] = 2
    b["type"] = np.full(n, -1, np.int32); b["type"][7] = F.BODY_KINEMATIC
    b["velocity"][7] = (0.5, 0, 0)
    b["fixed_rotation"] = np.zeros(n, np.uint8); b["fixed_rotation"][8] = 1
    b["linear_factor"] = np.ones((n, 3), np.float32); b["linear_factor"][9] = (1, 1, 0)
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, spec)
    for s in range(40):
        parity.staged_step(dev, ref, 1 / 60, f"flags step {s}")


def test_user_mutations_between_steps(cuda_lib, oracle_lib):
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, scenes.box_stacks(4, 3, grid=2))
    rng = np.random.default_rng(0)
    n = dev.n
    for s in range(30):
        f = rng.normal(size=(n, 3)).astype(np.float32) * 5
        t = rng.normal(size=(n, 3)).astype(np.float32)
        for w in (dev, ref):
            w.update_bodies(0, n, force=f, torque=t)
            w.step(1 / 60)
        parity.assert_same_state(dev, ref, f"mutations step {s}")


def test_hinge_motor_and_collide_connected(cuda_lib, oracle_lib):
    spec = scenes.chain_worlds(2, chains=2, links=4)
    for c in spec.constraints:
        if c["type"] == F.CONSTRAINT_HINGE:
            c.update(motor_enabled=1, motor_target_velocity=2.0, motor_max_force=5.0, collide_connected=0)
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, spec)
    for s in range(60):
        parity.staged_step(dev, ref, 1 / 60, f"motor step {s}")


@pytest.mark.parametrize("kind", ["reference", "colored"])
def test_live_setters_hinge_motor_sleep_state_inertia_stepnumber(cuda_lib, oracle_lib, kind):
    """The between-step setters a live world offers without a rebuild: HingeConstraint motor fields
    (hinge_constraint.dart:56-76; switching the motor on or off changes the set of accepted equations), Body.sleep / wakeUp
    (rigid_body.dart:263-278), Body.invInertia and World.stepnumber. Same calls on both libraries, bit-exact afterwards."""
    spec = _with(scenes.chain_worlds(2, chains=2, links=4), allow_sleep=1, quat_normalize_skip=2,
                 solver_kind=REF if kind == "reference" else F.SOLVER_COLORED)
    hinges = [k for k, c in enumerate(spec.constraints) if c["type"] == F.CONSTRAINT_HINGE]
    assert len(hinges) >= 2
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, spec)
    n = spec.n_bodies
    for s in range(50):
        if s == 10:
            for w in (dev, ref):
                w.set_hinge_motor(hinges[0], True, 2.0, 5.0)
                w.set_hinge_motor(hinges[1], True, -1.0, 50.0)
        if s == 20:
            inv = dev.get_bodies(("inv_inertia",))["inv_inertia"].reshape(n, 3).copy()
            inv[3] = (0.5, 2.0, 1.0)
            inv[4] = (1.0, 1.0, 1.0)   # isotropic: the world inertia is still refreshed (forced update)
            sl = dev.get_bodies(("sleep_state",))["sleep_state"].copy()
            sl[5] = F.SLEEPING if hasattr(F, "SLEEPING") else 2
            for w in (dev, ref):
                w.set_inv_inertia(3, inv[3:5])
                w.update_sleep_states(0, sl)
                w.set_stepnumber(7)
        if s == 30:
            for w in (dev, ref):
                w.set_hinge_motor(hinges[0], False, 0.0, 5.0)
                w.update_sleep_states(5, np.zeros(1, np.int32))
        for w in (dev, ref):
            w.step(1 / 60, 1)
        parity.assert_same_state(dev, ref, f"setters step {s}")
        if s == 20:
            a, b = dev.get_bodies(("inv_inertia", "inv_inertia_world")), ref.get_bodies(("inv_inertia", "inv_inertia_world"))
            assert np.array_equal(a["inv_inertia"], b["inv_inertia"]) and np.array_equal(a["inv_inertia_world"], b["inv_inertia_world"])
    assert dev.get_time() == ref.get_time() and dev.get_time()[1] == 7 + 30


def test_large_pair_set_equals_brute_force_and_is_deterministic(cuda_lib):
    # size-independent property at a size the oracle cannot finish quickly: the hashed-grid broadphase must
    # return exactly the Naive set {(i,j<i): |xj-xi|^2 < (ri+rj)^2} in i-major / j-ascending order
    spec = scenes.sphere_container(n_spheres=40000, pitch=0.45, extent=40.0, allow_sleep=False)
    w = engine.DeviceWorld(cuda_lib, spec)
    p1, p2 = w.broadphase_pairs()
    assert np.all(p2 < p1)
    key = p1.astype(np.int64) << 32 | p2
    assert np.all(np.diff(key) > 0)
    pos = spec.bodies["position"]
    dyn = np.arange(5, spec.n_bodies)
    from scipy.spatial import cKDTree
    tree = cKDTree(pos[dyn].astype(np.float64))
    cand = tree.query_pairs(0.5 + 1e-3, output_type="ndarray")
    a, b = dyn[cand[:, 0]], dyn[cand[:, 1]]
    r = (pos[b] - pos[a]).astype(np.float32).astype(np.float64)  # f32 difference, f64 norm (broadphase.dart:77-86)
    ok = (r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2]) < 0.25
    hi, lo = np.maximum(a[ok], b[ok]), np.minimum(a[ok], b[ok])
    expect = set((hi.astype(np.int64) << 32 | lo).tolist())
    sphere_pairs = key[p2 >= 5]
    assert set(sphere_pairs.tolist()) == expect
    # every sphere pairs with every plane (infinite bounding radius)
    assert np.count_nonzero(p2 < 5) - 0 == 5 * len(dyn) + 0 * 10 or np.count_nonzero(p2 < 5) == 5 * len(dyn)
    q1, q2 = w.broadphase_pairs()
    assert np.array_equal(p1, q1) and np.array_equal(p2, q2)


def test_colored_solver_statistical_agreement_and_determinism(cuda_lib):
    # throughput mode: different row order => statistical agreement only (rest penetration, no energy blow-up)
    def settle(kind):
        w = engine.DeviceWorld(cuda_lib, scenes.sphere_container(8, 8, 4, extent=6.0, solver=kind, allow_sleep=False))
        w.step(1 / 60, 360)
        return w.get_bodies(("position", "velocity"))
    a, b, c = settle(F.SOLVER_COLORED), settle(REF), settle(F.SOLVER_COLORED)
    d, e = settle(F.SOLVER_COLORED_F32), settle(F.SOLVER_COLORED_F32)
    assert np.array_equal(a["position"], c["position"]), "colored mode must be run-to-run deterministic"
    assert np.array_equal(d["position"], e["position"]), "the reduced-precision sweep must be run-to-run deterministic"
    # same colour order, f32 + FMA instead of the reference arithmetic: a granular pile is chaotic, so after 360 steps the
    # two agree statistically (mean height, no blow-up), not per body
    assert abs(d["position"][5:, 1].mean() - a["position"][5:, 1].mean()) < 0.05
    for s in (a, b, d):
        assert np.abs(s["velocity"][5:]).max() < 1.5 and np.abs(s["velocity"][5:]).mean() < 0.05
        assert s["position"][5:, 1].min() > 0.2  # rest penetration below 0.05
    assert abs(a["position"][5:, 1].mean() - b["position"][5:, 1].mean()) < 0.05
    assert abs(np.sort(a["position"][5:, 1])[-1] - np.sort(b["position"][5:, 1])[-1]) < 0.15


def test_oversize_hulls_take_the_sequential_sat_path(cuda_lib, oracle_lib):
    # 40-segment cylinders have 42 faces / 41 unique edges: larger than the tile kernel's scratch (32), so these
    # tasks run through the sequential SAT kernels; small hulls in the same scene still use the tile kernel
    from cannon_physics_b200.engine import SceneSpec
    n = 7
    b = {"position": np.array([[0, 0, 0], [0, 0.6, 0], [0.2, 1.7, 0.1], [0.1, 2.9, 0], [1.3, 0.6, 0], [1.2, 1.8, 0.2], [-1.0, 0.5, 0.3]], np.float32),
         "quaternion": np.tile(np.array([0, 0, 0, 1], np.float32), (n, 1)), "mass": np.array([0, 1, 1, 1, 1, 1, 1.0]),
         "shape": np.array([0, 1, 2, 1, 2, 3, 4], np.int32)}
    b["quaternion"][0] = scenes.GROUND_QUAT
    tetra = dict(type=F.SHAPE_CONVEX, vertices=[(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)], faces=[(0, 3, 2), (0, 1, 3), (0, 2, 1), (1, 2, 3)])
    shapes = [dict(type=F.SHAPE_PLANE), dict(type=F.SHAPE_CYLINDER, radius_top=0.5, radius_bottom=0.5, height=1.0, num_segments=40),
              dict(type=F.SHAPE_BOX, half_extents=(0.5, 0.5, 0.5)), dict(type=F.SHAPE_CYLINDER, radius_top=0.4, radius_bottom=0.5, height=1.0, num_segments=8), tetra]
    spec = SceneSpec(desc=dict(gravity=(0, -10, 0)), shapes=shapes, bodies=b, n_bodies=n)
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, spec)
    seen = 0
    for s in range(80):
        seen = max(seen, parity.staged_step(dev, ref, 1 / 60, f"oversize step {s}")[1])
    assert seen > 8


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c2 sap stacks", "c5 sleeping", "c4 joints", "c3 heightfield"])
def test_graph_replay_equals_eager_launches(cuda_lib, name):
    """cannon_world_step replays a captured CUDA graph of one step; the eager launch sequence must give the same bits
    (time / stepnumber live on the device, cooperative kernels size their own barriers)."""
    import os
    mk = {"c2 sap stacks": lambda: scenes.box_stacks(6, 5, grid=3),
          "c5 sleeping": lambda: scenes.sphere_container(6, 6, 3, extent=4.0, solver=F.SOLVER_REFERENCE_ORDER),
          "c4 joints": lambda: scenes.chain_worlds(6, chains=2, links=5),
          "c3 heightfield": lambda: scenes.mixed_pile_on_heightfield(6, 6, 3, solver=F.SOLVER_REFERENCE_ORDER)}[name]
    a = engine.DeviceWorld(cuda_lib, mk())
    b = engine.DeviceWorld(cuda_lib, mk())
    a.step(1 / 60, 70)  # step 1 eager, 69 replays
    a.step(1 / 60, 50)
    os.environ["CANNON_NO_GRAPH"] = "1"
    try:
        b.step(1 / 60, 70)
        b.step(1 / 60, 50)
    finally:
        del os.environ["CANNON_NO_GRAPH"]
    parity.assert_same_state(a, b, name)
    pa, pb = a.profile(), b.profile()
    for k in ("n_pairs", "n_contacts", "n_rows", "steps", "contact_iters_total", "kernel_launches"):
        if k != "kernel_launches":
            assert pa[k] == pb[k], (k, pa[k], pb[k])
    assert a.get_time() == b.get_time()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["ring (warp per world)", "staged (CANNON_GW_NO_RING)", "staged fallback: a world too big for the ring"])
def test_batch_colored_world_kernel_equals_grid_sweep(cuda_lib, variant):
    """A colored batch is swept world by world (k_gs_world_ring: one warp per world, or k_gs_world: one CTA per world when
    the device-side check finds a world that does not fit the ring's tables); the grid-wide staged sweep over the same
    colours must give the same bits: units of one colour touch disjoint bodies, so the order inside a colour cannot matter."""
    big = variant.startswith("staged fallback")
    def mk():
        spec = scenes.chain_worlds(6, chains=12, links=9) if big else scenes.chain_worlds(40, chains=2, links=6)  # 109 > 96 bodies
        spec.desc["solver_kind"] = F.SOLVER_COLORED_F32
        return spec
    if "NO_RING" in variant:
        os.environ["CANNON_GW_NO_RING"] = "1"  # read when the world is created
    try:
        a = engine.DeviceWorld(cuda_lib, mk())
    finally:
        os.environ.pop("CANNON_GW_NO_RING", None)
    b = engine.DeviceWorld(cuda_lib, mk())
    os.environ["CANNON_GS_NO_WORLD_KERNEL"] = "1"
    os.environ["CANNON_NO_GRAPH"] = "1"
    try:
        b.step(1 / 60, 90)
    finally:
        del os.environ["CANNON_GS_NO_WORLD_KERNEL"]
        del os.environ["CANNON_NO_GRAPH"]
    a.step(1 / 60, 90)
    parity.assert_same_state(a, b, "colored batch")
    assert a.profile()["n_rows"] == b.profile()["n_rows"] > 0
    # the colored order converges differently along a chain than the reference's list order (10 sweeps do not
    # converge either), so only sanity is asserted against it: the chains hold together and stay near their anchors
    pa = a.get_bodies(("position", "velocity"))
    assert np.all(np.isfinite(pa["position"])) and np.abs(pa["position"]).max() < 20 and np.abs(pa["velocity"]).max() < 30


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["spheres bouncing on a plane", "mixed pile on heightfield", "jointed chains (batch)"])
def test_contact_events_match_the_overlap_keeper(cuda_lib, oracle_lib, name):
    """beginContact / endContact lists (world_class.dart:703-730, overlap_keeper.dart) from the device
    pair sets equal the oracle's sorted-list OverlapKeeper, step by step, in the same order."""
    mk = {"spheres bouncing on a plane": lambda: _with(scenes.spheres_on_plane(5, 5, 4, spacing=0.55), default_contact_material=dict(restitution=0.6)),
          "mixed pile on heightfield": lambda: scenes.mixed_pile_on_heightfield(8, 8, 3, hf_samples=33, solver=REF, grid_cells=(8, 4, 8)),
          "jointed chains (batch)": lambda: scenes.chain_worlds(4, chains=3, links=6)}[name]
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, mk())
    dev.enable_contact_events(True)
    ref.enable_contact_events(True)
    n_begin = n_end = 0
    for s in range(150):
        dev.step(1 / 60)
        ref.step(1 / 60)
        (ba, ea), (bb, eb) = dev.get_contact_events(), ref.get_contact_events()
        assert np.array_equal(ba, bb) and np.array_equal(ea, eb), f"{name}: events differ at step {s}"
        n_begin += len(ba)
        n_end += len(ea)
    assert n_begin > 0 and n_end > 0, (n_begin, n_end)
    parity.assert_same_state(dev, ref, name)
    # a multi-step call keeps the events of its last step
    dev.step(1 / 60, 3)
    ref.step(1 / 60, 3)
    (ba, ea), (bb, eb) = dev.get_contact_events(), ref.get_contact_events()
    assert np.array_equal(ba, bb) and np.array_equal(ea, eb)


@pytest.mark.gpu
def test_world_api_dispatches_contact_events(cuda_lib):
    from cannon_physics_b200 import api
    world = api.World(gravity=(0, -10, 0), _lib=cuda_lib)
    ground = api.Body(mass=0, shape=api.Plane())
    ground.quaternion[:] = scenes.GROUND_QUAT
    ball = api.Body(mass=1, shape=api.Sphere(0.5), position=(0, 1.0, 0))
    world.addBody(ground)
    world.addBody(ball)
    heard = []
    world.addEventListener("beginContact", lambda e: heard.append((e["type"], e["bodyA"] is ground, e["bodyB"] is ball)))
    for _ in range(60):
        world.step(1 / 60)
    assert heard and heard[0] == ("beginContact", True, True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["events_bouncing", "events_heightfield"])
def test_cuda_contact_events_match_golden_fixture(cuda_lib, name):
    from make_golden import EVENT_CASES, run_events
    got = run_events(cuda_lib, *EVENT_CASES[name])
    ref = np.load(os.path.join(GOLDEN, name + ".npz"))
    for k in ref.files:
        assert np.array_equal(got[k], ref[k]), f"{name}: {k} differs from the fixture"


@pytest.mark.gpu
def test_dataflow_sweep_equals_level_sweep_on_the_100k_pile(cuda_lib):
    """The COLORED solver's production kernel (k_gs_exact: per-body progress counters, no grid barrier between colours)
    against the plain level-by-level sweep of the same colours (k_gs, CANNON_GS_NO_DATAFLOW) on the full bench workload -
    the settled 100 001-body pile of config 3 (~4.4e5 contacts, 1.3e6 rows), a size the oracle needs 10 s per step for.
    Units of one colour touch disjoint bodies and every unit runs after its predecessors on both bodies, so the two must
    agree bit for bit; any missed dependency or stale read in the dataflow would show up here."""
    import bench
    spec, _, state = bench.build_spec("c3", 1.0, 0, 1)
    assert "settled" in state
    a = engine.DeviceWorld(cuda_lib, spec)
    os.environ["CANNON_GS_NO_DATAFLOW"] = "1"  # read when the world is created
    try:
        b = engine.DeviceWorld(cuda_lib, spec)
    finally:
        del os.environ["CANNON_GS_NO_DATAFLOW"]
    for s in range(3):
        a.step(1 / 60, 4)
        b.step(1 / 60, 4)
        parity.assert_same_state(a, b, f"100k pile after {4 * (s + 1)} steps")
        pa, pb = a.profile(), b.profile()
        assert (pa["n_pairs"], pa["n_contacts"], pa["n_rows"], pa["n_levels"], pa["iterations_done"]) == \
               (pb["n_pairs"], pb["n_contacts"], pb["n_rows"], pb["n_levels"], pb["iterations_done"])
    assert pa["n_contacts"] > 300000
    ra, rb = a.get_rows(), b.get_rows()
    for k in ("body_i", "body_j", "B", "invC", "lambda", "level"):
        assert np.array_equal(ra[k], rb[k]), k


@pytest.mark.gpu
def test_cannon_batch_two_contexts_one_host_thread(cuda_lib, oracle_lib):
    """cannon_batch_* on the device: the batch split into two shards, each on its own context / stream (both on GPU 0 here;
    one per GPU on a multi-GPU box), stepped concurrently from this one thread (step_async on every shard, then ctx_sync).
    Per-world results equal the unsharded batch and the oracle bit for bit; the statistics add up."""
    import torch
    spec = _with(scenes.chain_worlds(12, chains=3, links=6), solver_kind=F.SOLVER_COLORED)
    whole, ref = parity.make_pair(cuda_lib, oracle_lib, spec)
    devices = (0, 1) if torch.cuda.device_count() > 1 else (0, 0)
    b = engine.DeviceBatch(cuda_lib, spec, devices=devices)
    for _ in range(3):
        whole.step(1 / 60, 20)
        ref.step(1 / 60, 20)
        b.step(1 / 60, 20)
        got, want = b.get_bodies(parity.STATE), ref.get_bodies(parity.STATE)
        for k in parity.STATE:
            assert np.array_equal(got[k], want[k]), k
        parity.assert_same_state(whole, ref, "unsharded batch")
    st = b.stats()
    assert st["steps"] == 60 and st["n_contacts"] == whole.profile()["n_contacts"] and st["step_call_ms_max"] > 0
    b.close()


@pytest.mark.gpu
def test_just_test_pairs_contact_events_parity(cuda_lib, oracle_lib):
    from test_oracle_properties import _just_test_spec
    dev, ref = parity.make_pair(cuda_lib, oracle_lib, _just_test_spec())
    dev.enable_contact_events(True)
    ref.enable_contact_events(True)
    nb = 0
    for s in range(240):
        dev.step(1 / 60)
        ref.step(1 / 60)
        (ba, ea), (bb, eb) = dev.get_contact_events(), ref.get_contact_events()
        assert np.array_equal(ba, bb) and np.array_equal(ea, eb), f"events differ at step {s}: {ba.tolist()} {bb.tolist()} / {ea.tolist()} {eb.tolist()}"
        nb += len(ba)
    assert nb >= 4
    parity.assert_same_state(dev, ref, "justTest scene")
