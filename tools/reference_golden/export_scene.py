"""Write a scene of cannon_physics_b200/scenes.py as JSON for the reference-side runner (golden_dump.dart). See README.md."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def jsonable(v):
    if isinstance(v, np.ndarray):
        return v.tolist()
    if isinstance(v, (np.floating, np.integer)):
        return v.item()
    if isinstance(v, dict):
        return {k: jsonable(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [jsonable(x) for x in v]
    return v


def export(spec, steps, dt=1 / 60, checkpoints=None):
    return {
        "name": spec.name, "dt": dt, "steps": steps, "checkpoints": checkpoints or sorted({1, 2, steps // 2, steps}),
        "desc": jsonable(spec.desc), "shapes": jsonable(spec.shapes), "n_bodies": spec.n_bodies,
        "bodies": jsonable(spec.bodies), "material_friction": jsonable(spec.material_friction),
        "material_restitution": jsonable(spec.material_restitution), "contact_materials": jsonable(spec.contact_materials),
        "constraints": jsonable(spec.constraints), "springs": jsonable(spec.springs),
        # Body.addShape table (first / shape / offset / orientation) and the SPHSystem subsystems
        "body_shapes": jsonable(spec.body_shapes), "sph_systems": jsonable(spec.sph_systems),
    }


def main():
    import make_golden
    from cannon_physics_b200 import _ffi as F
    ap = argparse.ArgumentParser()
    ap.add_argument("case", nargs="?")
    ap.add_argument("out", nargs="?")
    ap.add_argument("--list", action="store_true")
    ap.add_argument("--broadphase", choices=["keep", "naive"], default="keep")
    a = ap.parse_args()
    if a.list or not a.case:
        print("\n".join(make_golden.CASES))
        return
    mk, steps = make_golden.CASES[a.case]
    spec = mk()
    if a.broadphase == "naive":
        spec.desc["broadphase_kind"] = F.BP_NAIVE
    if spec.desc.get("n_worlds", 1) > 1:
        raise SystemExit("batches have no reference counterpart: export one world of the batch instead")
    with open(a.out, "w") as f:
        json.dump(export(spec, steps), f)
    print(f"{a.case}: {spec.n_bodies} bodies, {steps} steps -> {a.out}")


if __name__ == "__main__":
    main()
