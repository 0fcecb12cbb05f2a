"""Profiling aid (not a test): ms per step over the course of a run (chunks of 20 steps) and the stage split at each chunk end.

usage (GPU box): python tools/step_curve.py [config] [chunks]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cannon_physics_b200 as cp  # noqa: E402
from cannon_physics_b200 import engine  # noqa: E402

config = sys.argv[1] if len(sys.argv) > 1 else "c3"
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 11
spec, label = bench.build_spec(config, 1.0, 0, 1)
w = engine.DeviceWorld(cp.lib, spec, device=0)
for c in range(chunks):
    w.step(1 / 60, 20)
    p = w.profile()
    print(f"steps {c * 20:4d}-{c * 20 + 19:4d}: {p['step_call_ms'] / 20:6.3f} ms/step | last: bp {p['broadphase']:.3f} np {p['narrowphase']:.3f} "
          f"solve {p['solve']:.3f} (sched {p['schedule_ms']:.3f} gs {p['gs_ms']:.3f}) int {p['integrate']:.3f} | pairs {p['n_pairs']} contacts {p['n_contacts']} levels {p['n_levels']}")
