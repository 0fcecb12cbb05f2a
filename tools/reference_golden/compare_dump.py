"""Compare a dump of the real reference (golden_dump.dart) with the CPU oracle on the same scene. See README.md."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cannon_physics_b200 import _ffi as F  # noqa: E402
from cannon_physics_b200 import engine  # noqa: E402

FIELDS = ("position", "quaternion", "velocity", "angular_velocity", "sleep_state")


def spec_from_json(s):
    b = {}
    for k, v in s["bodies"].items():
        b[k] = np.asarray(v)
    shapes = []
    for sh in s["shapes"]:
        sh = dict(sh)
        for k in ("vertices", "hf_data"):
            if k in sh:
                sh[k] = np.asarray(sh[k])
        shapes.append(sh)
    return engine.SceneSpec(desc=s["desc"], shapes=shapes, bodies=b, n_bodies=s["n_bodies"],
                            material_friction=None if s["material_friction"] is None else np.asarray(s["material_friction"]),
                            material_restitution=None if s["material_restitution"] is None else np.asarray(s["material_restitution"]),
                            contact_materials=s["contact_materials"], constraints=s["constraints"], springs=s["springs"], name=s["name"],
                            body_shapes=None if s.get("body_shapes") is None else {k: (None if v is None else np.asarray(v)) for k, v in s["body_shapes"].items()},
                            sph_systems=s.get("sph_systems") or [])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene")
    ap.add_argument("dump")
    ap.add_argument("--save", help="write the dump as an .npz fixture")
    a = ap.parse_args()
    scene, dump = json.load(open(a.scene)), json.load(open(a.dump))
    lib = F.bind(os.path.join(ROOT, "oracle", "libcannon_oracle.so"))
    w = engine.DeviceWorld(lib, spec_from_json(scene))
    ok = True
    fixture = {"contacts_per_step": np.asarray(dump["contacts_per_step"], np.int32)}
    for step in range(1, scene["steps"] + 1):
        w.step(scene["dt"], 1)
        nc = w.profile()["n_contacts"]
        if nc != dump["contacts_per_step"][step - 1]:
            print(f"step {step}: contact count oracle {nc} != reference {dump['contacts_per_step'][step - 1]}")
            ok = False
        cp = dump["checkpoints"].get(str(step))
        if cp is None:
            continue
        st = w.get_bodies(FIELDS)
        for k in FIELDS:
            ref = np.asarray(cp[k], dtype=st[k].dtype).reshape(st[k].shape)
            fixture[f"{k}_{step}"] = ref
            if np.array_equal(ref, st[k]):
                print(f"step {step}: {k} bit-exact")
                continue
            ok = False
            err = np.abs(ref.astype(np.float64) - st[k]) / np.maximum(np.abs(ref.astype(np.float64)), 1e-6)
            print(f"step {step}: {k} differs in {(ref != st[k]).sum()} entries, max relative error {err.max():.3e}")
    if a.save:
        np.savez_compressed(a.save, **fixture)
    print("ORACLE PINNED: identical to the reference on this scene" if ok else "MISMATCH (see above)")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
