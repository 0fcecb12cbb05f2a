#!/usr/bin/env python
"""Writes the settled state of BASELINE config 3 that bench.py starts from (both arms).

The c3 recipe (scenes.mixed_pile_on_heightfield, seed 3) starts as a lattice hanging 1 m above the terrain; the
headline regime is the *pile* (~4e5 contacts), which the lattice reaches after ~250 steps. This script steps the
recipe on the GPU and saves position / quaternion / velocity / angular velocity / sleep state of all 100 001 bodies:

    gpurun -- python tools/make_settled.py --steps 250 --out gpurun_out/c3_settled.npz
    cp gpurun_out/c3_settled.npz bench_data/c3_settled.npz

Any trajectory of the engine is a valid starting state (both bench arms load the same file); the file also records
the library version, the solver kind that produced it and the contact count of the last step.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=250)
    ap.add_argument("--side", type=int, default=100)
    ap.add_argument("--layers", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "c3_settled.npz"))
    args = ap.parse_args()

    import cannon_physics_b200 as cp
    from cannon_physics_b200 import _ffi as F
    from cannon_physics_b200 import engine, scenes

    spec = scenes.mixed_pile_on_heightfield(args.side, args.side, args.layers, seed=3, solver=F.SOLVER_COLORED)
    w = engine.DeviceWorld(cp.lib, spec)
    curve = []
    done = 0
    while done < args.steps:
        k = min(25, args.steps - done)
        w.step(1.0 / 60.0, k)
        done += k
        p = w.profile()
        curve.append((done, p["n_pairs"], p["n_contacts"], p["step_call_ms"] / k))
        print(f"step {done}: pairs={p['n_pairs']} contacts={p['n_contacts']} ms/step={p['step_call_ms'] / k:.3f}", flush=True)
    st = w.get_bodies(("position", "quaternion", "velocity", "angular_velocity", "sleep_state"))
    t, sn = w.get_time()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    np.savez_compressed(args.out, position=st["position"], quaternion=st["quaternion"], velocity=st["velocity"],
                        angular_velocity=st["angular_velocity"], sleep_state=st["sleep_state"],
                        steps=np.int64(args.steps), time=np.float64(t), side=np.int64(args.side), layers=np.int64(args.layers),
                        seed=np.int64(3), curve=np.asarray(curve, dtype=np.float64),
                        abi_version=np.int64(cp.lib.cannon_version()))
    print(f"wrote {args.out}: {spec.n_bodies} bodies after {args.steps} steps, {curve[-1][2]} contacts")


if __name__ == "__main__":
    main()
