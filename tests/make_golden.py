"""Generates tests/golden/*.npz from the CPU oracle (the reference itself cannot run here: no Dart SDK,
SURVEY.md §8c). The fixtures freeze the oracle's behaviour so (a) oracle regressions are caught on CPU and
(b) the CUDA path is compared against committed vectors on the GPU box, not only against a live oracle.

    python tests/make_golden.py        # rewrites the fixtures
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cannon_physics_b200 import _ffi as F  # noqa: E402
from cannon_physics_b200 import engine, scenes  # noqa: E402

REF = F.SOLVER_REFERENCE_ORDER
CASES = {
    "c1_small": (lambda: scenes.spheres_on_plane(4, 4, 4), 90),
    "c2_small": (lambda: scenes.box_stacks(4, 5, grid=2), 60),
    "c3_plane_small": (lambda: scenes.mixed_pile_on_heightfield(4, 4, 3, with_heightfield=False, solver=REF, grid_cells=(8, 4, 8)), 90),
    "c3_hf_small": (lambda: scenes.mixed_pile_on_heightfield(4, 4, 3, hf_samples=33, solver=REF, grid_cells=(8, 4, 8)), 90),
    "c4_small": (lambda: scenes.chain_worlds(3, chains=2, links=4), 60),
    "c5_small": (lambda: scenes.sphere_container(6, 6, 4, extent=5.0, solver=REF), 120),
    "joints_small": (lambda: scenes.constraint_zoo(groups=2), 120),
    # COLORED (the throughput solver bench.py times): GSSolver's arithmetic in the colour order of include/cannon_cuda.h
    "c2_colored_small": (lambda: _with(scenes.box_stacks(4, 5, grid=2), solver_kind=F.SOLVER_COLORED), 60),
    "c3_hf_colored_small": (lambda: scenes.mixed_pile_on_heightfield(4, 4, 3, hf_samples=33, solver=F.SOLVER_COLORED, grid_cells=(8, 4, 8)), 90),
    "c4_colored_small": (lambda: _with(scenes.chain_worlds(3, chains=2, links=4), solver_kind=F.SOLVER_COLORED), 60),
    # World(quatNormalizeFast: true, quatNormalizeSkip: 2, frictionGravity: ...) (world_class.dart:135-144,668; quaternion.dart:171-185)
    "c2_quatfast_small": (lambda: _with(scenes.mixed_pile_on_heightfield(4, 4, 3, with_heightfield=False, solver=REF, grid_cells=(8, 4, 8)),
                                        quat_normalize_fast=1, quat_normalize_skip=2, has_friction_gravity=1, friction_gravity=(0, -3, 0)), 90),
    # SURVEY 8f rank 4 (the scene builders live next to their tests): hull subclasses, Particle, compound bodies, Trimesh, SPH
    "hulls_small": (lambda: _t("test_hull_shapes")._mixed_spec("heightfield"), 90),
    "particles_small": (lambda: _t("test_particle")._pile_spec("heightfield", n_part=24), 90),
    "compound_small": (lambda: _t("test_compound")._pile_spec("heightfield", n_obj=8), 90),
    "compound_colored_small": (lambda: _t("test_compound")._pile_spec("plane", solver=F.SOLVER_COLORED, n_obj=8), 90),
    "trimesh_small": (lambda: _t("test_trimesh")._terrain_spec(scale=(1.25, 0.75, 1.0)), 80),
    "sph_small": (lambda: _t("test_sph")._block_spec(4, 3, 4), 60),
}
RANK4_CASES = ["hulls_small", "particles_small", "compound_small", "compound_colored_small", "trimesh_small", "sph_small"]


def _t(module):
    import importlib
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    return importlib.import_module(module)


def _with(spec, **desc):
    spec.desc.update(desc)
    return spec


def run_case(lib, make_spec, steps):
    w = engine.DeviceWorld(lib, make_spec())
    out = {}
    w.step(1 / 60, steps - 1)
    # last step through the staged entry points so pair lists / contacts are part of the fixture
    w.set_dt(1 / 60)
    w.apply_gravity()
    p1, p2 = w.broadphase_pairs()
    c = w.narrowphase_contacts(p1, p2)
    w.solver_solve(1 / 60)
    rows = w.get_rows()
    w.integrate(1 / 60)
    st = w.get_bodies(("position", "quaternion", "velocity", "angular_velocity", "sleep_state"))
    out.update(p1=p1, p2=p2, per_pair_count=c["per_pair_count"], c_bi=c["body_i"], c_bj=c["body_j"], c_ri=c["ri"], c_rj=c["rj"], c_ni=c["ni"],
               row_B=rows["B"], row_invC=rows["invC"], row_lambda=rows["lambda"], **st)
    return out


def _bouncing():
    spec = scenes.spheres_on_plane(5, 5, 4, spacing=0.55)
    spec.desc["default_contact_material"] = dict(restitution=0.6)
    return spec


# world-level contact events (world_class.dart:703-730): per step, the beginContact / endContact pair lists
EVENT_CASES = {
    "events_bouncing": (_bouncing, 120),
    "events_heightfield": (lambda: scenes.mixed_pile_on_heightfield(6, 6, 3, hf_samples=33, solver=REF, grid_cells=(8, 4, 8)), 120),
}


def run_events(lib, make_spec, steps):
    """(step, a, b) rows of every beginContact / endContact event of the run, in dispatch order."""
    w = engine.DeviceWorld(lib, make_spec())
    w.enable_contact_events(True)
    begin, end = [], []
    for s in range(steps):
        w.step(1 / 60)
        b, e = w.get_contact_events()
        begin += [(s, int(x), int(y)) for x, y in b]
        end += [(s, int(x), int(y)) for x, y in e]
    return {"begin": np.array(begin, np.int32).reshape(-1, 3), "end": np.array(end, np.int32).reshape(-1, 3)}


if __name__ == "__main__":
    lib = F.bind(os.path.join(ROOT, "oracle", "libcannon_oracle.so"))
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    only = sys.argv[1:]  # optional: fixture names to (re)write; default all
    for name, (mk, steps) in CASES.items():
        if only and name not in only:
            continue
        res = run_case(lib, mk, steps)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **res)
        print(name, {k: v.shape for k, v in res.items() if k in ("p1", "c_bi", "row_B")})
    for name, (mk, steps) in EVENT_CASES.items():
        if only and name not in only:
            continue
        res = run_events(lib, mk, steps)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **res)
        print(name, {k: v.shape for k, v in res.items()})
