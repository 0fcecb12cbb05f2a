"""CPU-only checks of the host layer: ABI export + layout, scene generators, API flattening, sharding."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from cannon_physics_b200 import _ffi as F
from cannon_physics_b200 import engine, scenes
from cannon_physics_b200.batch import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    import re
    text = open(os.path.join(ROOT, "include", "cannon_cuda.h")).read()
    return sorted(set(re.findall(r"\b(cannon_[a-z_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _header_symbols() == sorted(F.PROTOTYPES)


@pytest.mark.parametrize("which", ["cuda", "oracle"])
def test_library_exports_every_declared_symbol(which, oracle_lib):
    path = os.path.join(ROOT, "cannon_physics_b200", "libcannon_cuda.so") if which == "cuda" else os.path.join(ROOT, "oracle", "libcannon_oracle.so")
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if l.strip()}
    missing = [s for s in F.PROTOTYPES if s not in exported]
    assert not missing, missing


def test_cuda_library_loads_without_gpu_and_reports_nogpu(cuda_lib):
    # no compute calls here: loading, version and defaults must work on a CPU-only box
    assert cuda_lib.cannon_version() == 1
    assert cuda_lib.cannon_backend() == b"cuda"
    import torch
    if not torch.cuda.is_available():
        h = F.VP()
        assert cuda_lib.cannon_ctx_create(0, C.byref(h)) == F.E_NOGPU  # no CPU fallback, fails loudly


@pytest.mark.parametrize("which", ["cuda", "oracle"])
def test_struct_layouts_match_c(which, cuda_lib, oracle_lib):
    lib = cuda_lib if which == "cuda" else oracle_lib
    d = F.WorldDesc()
    lib.cannon_world_desc_default(C.byref(d))
    assert (d.solver_iterations, d.solver_tolerance, d.broadphase_kind, d.n_worlds) == (10, 1e-7, F.BP_NAIVE, 1)
    assert (d.grid_nx, d.grid_ny, d.grid_nz) == (10, 10, 10) and list(d.grid_min) == [100.0] * 3 and list(d.grid_max) == [-100.0] * 3
    cm = d.default_contact_material
    assert (cm.friction, cm.restitution, cm.contact_equation_stiffness, cm.contact_equation_relaxation) == (0.3, 0.0, 1e7, 3.0)
    assert (cm.friction_equation_stiffness, cm.friction_equation_relaxation) == (1e7, 3.0)
    s = F.ShapeDesc()
    lib.cannon_shape_desc_default(C.byref(s))
    assert (s.type, s.collision_response, s.collision_filter_group, s.collision_filter_mask) == (0, 1, -1, -1)
    assert (s.radius, s.radius_top, s.radius_bottom, s.height, s.num_segments, s.hf_element_size) == (1.0, 1.0, 1.0, 1.0, 8, 1)


def test_splitmix64_known_values_and_vectorised_stream():
    r = scenes.SplitMix64(0)
    assert [r.next_u64() for _ in range(3)] == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]
    a, b = scenes.SplitMix64(12345), scenes.SplitMix64(12345)
    seq = np.array([(a.next_u64() >> 11) / (1 << 53) for _ in range(64)])
    assert np.array_equal(seq, b.uniform(64)) and a.s == b.s


def test_scene_generators_are_deterministic_and_shaped():
    s1, s2 = scenes.spheres_on_plane(3, 3, 3), scenes.spheres_on_plane(3, 3, 3)
    assert s1.n_bodies == 28 and np.array_equal(s1.bodies["position"], s2.bodies["position"])
    c2 = scenes.box_stacks(4, 5, grid=2)
    assert c2.n_bodies == 21 and c2.desc["broadphase_kind"] == F.BP_SAP and c2.contact_materials[0]["restitution"] == 0.2
    c3 = scenes.mixed_pile_on_heightfield(4, 4, 2, hf_samples=33)
    assert c3.n_bodies == 33 and c3.shapes[0]["hf_data"].shape == (33, 33) and c3.desc["broadphase_kind"] == F.BP_GRID
    assert np.all(c3.shapes[0]["hf_data"][0, :] == 3.0)
    c4 = scenes.chain_worlds(3, chains=2, links=4)
    assert c4.n_bodies == 3 * 9 and len(c4.constraints) == 3 * 2 * (2 + 1 + 2) and np.all(np.diff(c4.bodies["world_id"]) >= 0)
    c5 = scenes.sphere_container(n_spheres=100)
    assert c5.n_bodies == 105 and c5.desc["allow_sleep"] == 1


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 4096):
        for ws in (1, 2, 3, 8):
            parts = [shard_range(n, r, ws) for r in range(ws)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(ws - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 4, 4)


def test_api_world_flattens_like_the_reference_objects(oracle_lib):
    from cannon_physics_b200 import api
    stone = api.Material(name="stone")
    world = api.World(gravity=(0, -10, 0), broadphase=api.SAPBroadphase(axisIndex=0), solver=api.GSSolver(), _lib=oracle_lib)
    world.solver.iterations = 20
    world.addContactMaterial(api.ContactMaterial(stone, stone, friction=0.3, restitution=0.2))
    ground = api.Body(mass=0, material=stone, shape=api.Plane(), quaternion=api.Quaternion.setFromEuler(-np.pi / 2, 0, 0))
    world.addBody(ground)
    shape = api.Box((0.5, 0.5, 0.5))
    boxes = [api.Body(mass=1, material=stone, shape=shape, position=(0, 0.5 + 1.02 * k, 0)) for k in range(3)]
    for b in boxes:
        world.addBody(b)
    spec = world._spec()
    assert spec.n_bodies == 4 and len(spec.shapes) == 2 and spec.desc["solver_iterations"] == 20 and spec.desc["broadphase_kind"] == F.BP_SAP
    assert spec.bodies["type"].tolist() == [F.BODY_STATIC, 0, 0, 0] and spec.bodies["material"].tolist() == [0] * 4
    for _ in range(30):
        world.step(1 / 60)
    assert world.stepnumber == 30 and abs(world.time - 0.5) < 1e-12
    ys = [float(b.position[1]) for b in boxes]
    assert 0.45 < ys[0] < 0.55 and ys[0] < ys[1] < ys[2]  # the stack stays a stack
    assert len(world.contacts["body_i"]) >= 8
    boxes[2].addShape(api.Sphere(0.4), offset=(0, 0.9, 0))  # a compound body: the world is rebuilt with a shape table
    spec = world._spec()
    assert spec.body_shapes is not None and spec.body_shapes["first"].tolist() == [0, 1, 2, 3, 5] and spec.body_shapes["shape"].tolist() == [0, 1, 1, 1, 2]
    world.step(1 / 60)
    assert world.stepnumber == 31


def test_world_api_contact_event_listeners(oracle_lib):
    """Host mirror of EventTarget for the world-level contact events (world_class.dart:703-730), driven by the checker
    library on the CPU: begin once when the ball lands, nothing while it rests, `endContact` when it is lifted away."""
    from cannon_physics_b200 import api, scenes
    world = api.World(gravity=(0, -10, 0), _lib=oracle_lib)
    ground = api.Body(mass=0, shape=api.Plane())
    ground.quaternion[:] = scenes.GROUND_QUAT
    ball = api.Body(mass=1, shape=api.Sphere(0.5), position=(0, 0.8, 0))
    world.addBody(ground)
    world.addBody(ball)
    heard = []
    world.addEventListener("beginContact", lambda e: heard.append(("begin", e["bodyA"] is ground, e["bodyB"] is ball)))
    world.addEventListener("endContact", lambda e: heard.append(("end", e["bodyA"] is ground, e["bodyB"] is ball)))
    assert world.hasAnyEventListener("beginContact") and not world.hasAnyEventListener("collide")
    for _ in range(60):
        world.step(1 / 60)
    assert heard == [("begin", True, True)]
    ball.position[:] = (0, 5, 0)
    ball.velocity[:] = 0
    world.markDirty()
    world.step(1 / 60, nsteps=2)  # listeners hear every step of a multi-step call
    assert heard == [("begin", True, True), ("end", True, True)]
    import pytest
    with pytest.raises(api.CannonError):
        world.addEventListener("collide", lambda e: None)


def test_profiles_readme_matches_the_evidence_files():
    """profiles/README.md is generated (profiles/make_readme.py) from the committed bench lines: regenerating it must not
    change a byte, so the numbers quoted there are the numbers in profiles/*.json."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    readme = os.path.join(root, "profiles", "README.md")
    before = open(readme).read()
    subprocess.check_call([sys.executable, os.path.join(root, "profiles", "make_readme.py")], stdout=subprocess.DEVNULL)
    after = open(readme).read()
    assert before == after
    assert "@@" not in after


def test_world_api_remove_body_and_constraint(oracle_lib):
    """World.removeBody / removeConstraint / clearForces (world_class.dart:234-236,303-320,773-781) in the host mirror: a
    structural change rebuilds the device world from the current body state, the remaining bodies are re-indexed and keep
    moving as if nothing else had happened."""
    import pytest
    from cannon_physics_b200 import api, scenes

    def build(with_extra):
        w = api.World(gravity=(0, -10, 0), _lib=oracle_lib)
        g = api.Body(mass=0, shape=api.Plane())
        g.quaternion[:] = scenes.GROUND_QUAT
        a = api.Body(mass=1, shape=api.Sphere(0.5), position=(0, 3, 0))
        b = api.Body(mass=1, shape=api.Sphere(0.5), position=(5, 4, 0))   # far away: never touches `a`
        c = api.Body(mass=1, shape=api.Box((0.25, 0.25, 0.25)), position=(5, 6, 0))
        for body in (g, a) + ((b, c) if with_extra else ()):
            w.addBody(body)
        return w, g, a, b, c

    w, g, a, b, c = build(True)
    joint = api.DistanceConstraint(b, c)
    w.addConstraint(joint)
    for _ in range(30):
        w.step(1 / 60, sync=False)   # the Body objects are stale until a structural change pulls the state
    with pytest.raises(api.CannonError):
        w.removeBody(c)
        w.step(1 / 60)               # the joint still references `c`
    w.addBody(c)
    w.removeConstraint(joint)
    w.removeBody(b)
    w.removeBody(c)
    assert [body.index for body in w.bodies] == [0, 1] and b.index == -1 and b.world is None
    w.clearForces()
    for _ in range(60):
        w.step(1 / 60)
    ref, _, a_ref, _, _ = build(False)
    for _ in range(90):
        ref.step(1 / 60)
    # `a` never interacted with the removed bodies: same trajectory as in a world that never had them
    assert np.array_equal(a.position, a_ref.position) and np.array_equal(a.velocity, a_ref.velocity)


def test_world_api_mutators_after_unsynced_steps_do_not_roll_the_world_back(oracle_lib):
    """After step(sync=False) the device is ahead of the Body objects. Every mutator (applyForce, applyImpulse, wakeUp,
    sleep, clearForces) pulls the device state before it writes, so the next upload can never restore a stale pose;
    markDirty() - whose in-place edits have already happened - refuses instead."""
    from cannon_physics_b200 import api

    def build():
        w = api.World(gravity=(0, -10, 0), _lib=oracle_lib)
        a = api.Body(mass=1, shape=api.Sphere(0.5), position=(0, 10, 0))
        b = api.Body(mass=2, shape=api.Box((0.3, 0.2, 0.1)), position=(3, 10, 0), angularVelocity=(1, 2, 3))
        w.addBody(a)
        w.addBody(b)
        return w, a, b

    ref, ra, rb = build()
    for _ in range(11):
        ref.step(1 / 60)
    for mutate in ("applyForce", "applyImpulse", "wakeUp", "clearForces", "addContactMaterial"):
        w, a, b = build()
        for _ in range(10):
            w.step(1 / 60, sync=False)
        if mutate == "applyForce":
            a.applyForce((0, 0, 0))
        elif mutate == "applyImpulse":
            a.applyImpulse((0, 0, 0))
        elif mutate == "wakeUp":
            a.wakeUp()
        elif mutate == "clearForces":
            w.clearForces()
        else:
            w.addContactMaterial(api.ContactMaterial(api.Material(), api.Material(), friction=0.1))
        w.step(1 / 60)
        assert w.stepnumber == 11, mutate
        for x, y in ((a, ra), (b, rb)):
            assert np.array_equal(x.position, y.position) and np.array_equal(x.velocity, y.velocity), mutate
            assert np.array_equal(x.quaternion, y.quaternion) and np.array_equal(x.angularVelocity, y.angularVelocity), mutate
    w, a, b = build()
    w.step(1 / 60, sync=False)
    with pytest.raises(api.CannonError):
        w.markDirty()
    w.sync()
    a.position[1] = 20.0
    w.markDirty()
    w.step(1 / 60)
    assert a.position[1] > 19.9


def test_world_api_rebuild_keeps_stepnumber_inertia_and_sleep_timers(oracle_lib):
    """A structural change (addBody far away) rebuilds the device world. What the reference would not touch stays put:
    stepnumber (the quatNormalizeSkip phase), the tumbling box's invInertia (computed once at construction,
    rigid_body.dart:85,362 - not from the pose at rebuild time) and timeLastSleepy; body.sleep() / wakeUp() do not rebuild."""
    from cannon_physics_b200 import api

    def build():
        w = api.World(gravity=(0, 0, 0), quatNormalizeSkip=3, allowSleep=True, _lib=oracle_lib)
        box = api.Body(mass=2, shape=api.Box((0.5, 0.2, 0.1)), position=(0, 0, 0), angularVelocity=(1, 2, 3), angularDamping=0.0)
        w.addBody(box)
        return w, box

    ref, rbox = build()
    for _ in range(60):
        ref.step(1 / 60)
    w, box = build()
    for _ in range(20):
        w.step(1 / 60, sync=False)
    w.addBody(api.Body(mass=1, shape=api.Sphere(0.5), position=(100, 0, 0)))
    dev_before = None
    for k in range(40):
        w.step(1 / 60, sync=False)
        if k == 0:
            dev_before = w._dev
            assert w.stepnumber == 21
    w.sync()
    assert w.stepnumber == 60
    assert np.array_equal(box.quaternion, rbox.quaternion) and np.array_equal(box.angularVelocity, rbox.angularVelocity)
    box.sleep()
    w.step(1 / 60)
    assert w._dev is dev_before and box.sleepState == api.BodySleepStates.sleeping and not box.angularVelocity.any()
    box.wakeUp()
    w.step(1 / 60)
    assert w._dev is dev_before and box.sleepState != api.BodySleepStates.sleeping  # at rest: awake -> sleepy (rigid_body.dart:290)


def test_world_api_hinge_motor_setters_reach_a_live_world(oracle_lib):
    """HingeConstraint.enableMotor / setMotorSpeed / setMotorMaxForce / disableMotor (hinge_constraint.dart:56-76) after the
    first step: effective from the next step, without a rebuild, identical to a world that was built with the motor on."""
    from cannon_physics_b200 import api

    def build(motor_from_start):
        w = api.World(gravity=(0, 0, 0), _lib=oracle_lib)
        base = api.Body(mass=0, shape=api.Box((0.5, 0.5, 0.5)), position=(0, 0, 0))
        wheel = api.Body(mass=1, shape=api.Sphere(0.5), position=(2, 0, 0), angularDamping=0.0)
        w.addBody(base)
        w.addBody(wheel)
        h = api.HingeConstraint(base, wheel, pivotA=(2, 0, 0), pivotB=(0, 0, 0), axisA=(0, 0, 1), axisB=(0, 0, 1))
        if motor_from_start:
            h.enableMotor()
            h.setMotorSpeed(2.0)
            h.setMotorMaxForce(50.0)
        w.addConstraint(h)
        return w, wheel, h

    a, wa, ha = build(False)
    a.step(1 / 60)            # the motor is off: the wheel stays (numerically almost) at rest
    assert np.abs(wa.angularVelocity).max() < 1e-9
    dev = a._dev
    ha.enableMotor()
    ha.setMotorSpeed(2.0)
    ha.setMotorMaxForce(50.0)
    # the same history, but the motor fields reach the device through a full rebuild of the world (descriptor path)
    b, wb, hb = build(False)
    b.step(1 / 60)
    hb.motorEnabled, hb.motorTargetVelocity, hb.motorMaxForce = True, 2.0, 50.0
    b._structure_dirty = True
    for _ in range(30):
        a.step(1 / 60)
        b.step(1 / 60)
    assert a._dev is dev
    assert abs(wa.angularVelocity[2]) > 1.0
    assert np.array_equal(wa.angularVelocity, wb.angularVelocity) and np.array_equal(wa.quaternion, wb.quaternion)
    ha.disableMotor()
    before = wa.angularVelocity.copy()
    for _ in range(10):
        a.step(1 / 60)
    assert a._dev is dev and np.allclose(wa.angularVelocity, before, atol=1e-5)  # free spinning again


def test_world_api_spring_constraint_is_a_bounded_distance_row(oracle_lib):
    """SpringConstraint (spring_constraint.dart:7-56): rest length = the distance at construction, one bidirectional
    equation with |force| <= stiffness. A stiff one holds a pendulum like a DistanceConstraint of that length; a soft one
    (bound below the bob's weight) lets it sag."""
    from cannon_physics_b200 import api

    def run(make):
        w = api.World(gravity=(0, -10, 0), _lib=oracle_lib)
        anchor = api.Body(mass=0, shape=api.Sphere(0.1), position=(0, 5, 0))
        bob = api.Body(mass=1, shape=api.Sphere(0.1), position=(0, 3, 0), linearDamping=0.0)
        w.addBody(anchor)
        w.addBody(bob)
        w.addConstraint(make(anchor, bob))
        for _ in range(60):
            w.step(1 / 60)
        return bob.position.copy()

    stiff = run(lambda a, b: api.SpringConstraint(a, b, stiffness=1e6))
    dist = run(lambda a, b: api.DistanceConstraint(a, b, 2.0, 1e6))
    soft = run(lambda a, b: api.SpringConstraint(a, b, stiffness=0.05))
    assert np.array_equal(stiff, dist) and abs(stiff[1] - 3.0) < 0.05
    # the bound acts on the multiplier, an impulse per step (gs_solver.dart:88-98): 0.05 N s per 1/60 s = 3 N against a
    # 10 N weight - the bob falls, only slowed down
    assert soft[1] < 2.0

def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the CUDA arm) needs no GPU: one JSON line with the
    same metric / unit / config keys, `impl: reference`, a cpu_baseline describing the run and a zero-copy e2e record.
    (Run at 1/16 of the bodies here to keep the CPU suite short; the default is the full 100 001-body settled state.)"""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.check_output([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                                   "--scale", "0.0625"], stderr=subprocess.DEVNULL, text=True)
    j = json.loads(out.strip().splitlines()[-1])
    assert j["impl"] == "reference" and j["metric"] == "body-steps/s" and j["unit"] == "body-steps/s" and j["higher_is_better"] is True
    assert j["n_gpus"] == 1 and j["steps"] == 2 and j["warmup"] == 1 and j["value"] > 0
    assert j["config"]["workload"].startswith("c3") and j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["value"] == j["value"]
    assert j["config"]["same_config"] is True and j["dtype"] == "f64"
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_reference_arm_uses_the_settled_state_and_never_maps_the_cuda_library():
    """Both bench arms start config 3 from the committed settled snapshot (same 100 001 bodies), and the reference arm gets
    its host modules without running the package __init__, so libcannon_cuda.so is never mapped into that process."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); import bench\n"
        "spec, label, state = bench.build_spec('c3', 1.0, 0, 1, solver='reference')\n"
        "st = np.load(bench.SETTLED)\n"
        "assert spec.n_bodies == 100001 and 'settled' in state and np.array_equal(spec.bodies['position'], st['position'])\n"
        "assert np.array_equal(spec.bodies['velocity'], st['velocity']) and float(np.abs(st['velocity']).max()) > 0\n"
        "lib = bench.oracle_lib(); assert lib.cannon_backend() == b'oracle'\n"
        "maps = open('/proc/self/maps').read()\n"
        "assert 'libcannon_oracle.so' in maps and 'libcannon_cuda' not in maps, 'reference arm mapped the product library'\n"
        "print('ok')\n" % root)
    out = subprocess.check_output([sys.executable, "-c", code], text=True)
    assert out.strip().endswith("ok")


def test_world_api_lock_constraint_survives_a_rebuild(oracle_lib):
    """LockConstraint derives its pivots and frame vectors from the poses its constructor sees (lock_constraint.dart:29-43).
    A structural change later rebuilds the device world: the constraint must come back as it was made (the recorded poses
    travel in cannon_constraint_desc.ctor_*), not re-lock the bodies in the pose they have drifted to."""
    from cannon_physics_b200 import api

    def build():
        w = api.World(gravity=(0, -10, 0), _lib=oracle_lib)
        a = api.Body(mass=0, shape=api.Box((0.25, 0.25, 0.25)), position=(0, 5, 0))
        b = api.Body(mass=1, shape=api.Box((0.25, 0.25, 0.25)), position=(1, 5, 0), quaternion=(0, 0.38268343, 0, 0.92387953), angularDamping=0.0)
        w.addBody(a)
        w.addBody(b)
        w.addConstraint(api.LockConstraint(a, b, maxForce=50.0))  # soft: the hanging box sags and swings
        return w, b

    ref, rb = build()
    for _ in range(60):
        ref.step(1 / 60)
    w, b = build()
    for _ in range(25):
        w.step(1 / 60, sync=False)
    w.addBody(api.Body(mass=1, shape=api.Sphere(0.2), position=(50, 0, 0)))  # rebuild while the box is displaced
    for _ in range(35):
        w.step(1 / 60, sync=False)
    w.sync()
    assert np.array_equal(b.position, rb.position) and np.array_equal(b.quaternion, rb.quaternion)


def test_dart_binding_matches_the_header():
    """dart/cannon_cuda_bindings.dart cannot be compiled here (no SDK): check it against include/cannon_cuda.h textually -
    every exported symbol is looked up, and every struct has the header's fields in the header's order (camelCase)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "cannon_cuda.h")).read(), flags=re.S)
    d = open(os.path.join(root, "dart", "cannon_cuda_bindings.dart")).read()
    syms = set(re.findall(r"\b(cannon_[a-z0-9_]+)\s*\(", h))
    assert syms == set(re.findall(r"'(cannon_[a-z0-9_]+)'", d))

    def camel(s):
        p = s.split("_")
        return p[0] + "".join(x.capitalize() for x in p[1:])

    n_structs = 0
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\} (\w+);", h, flags=re.S):
        fields = []
        for line in m.group(2).split(";"):
            line = line.strip()
            if not line:
                continue
            decl = re.sub(r"^(const\s+)?\w+\s*\**\s*", "", line)
            for nm in decl.split(","):
                nm = re.sub(r"\[.*\]", "", nm.strip().lstrip("*").strip())
                if nm:
                    fields.append(camel(nm))
        cls = "".join(x.capitalize() for x in m.group(3).split("_"))
        dm = re.search(r"final class " + cls + r" extends Struct \{(.*?)\n\}", d, flags=re.S)
        assert dm, f"no Dart struct for {m.group(3)}"
        assert re.findall(r"external (?:[\w<>]+) (\w+);", dm.group(1)) == fields, m.group(3)
        n_structs += 1
    assert n_structs >= 12
