"""Profiling aid: narrowphase task / contact statistics per resolver type for a bench configuration."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cannon_physics_b200 as cp  # noqa: E402
from cannon_physics_b200 import engine  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 220
config = sys.argv[2] if len(sys.argv) > 2 else "c3"
spec, label, _state = bench.build_spec(config, 1.0, 0, 1)
w = engine.DeviceWorld(cp.lib, spec, device=0)
w.step(1 / 60, steps)
prof = w.profile()
print({k: prof[k] for k in ("n_pairs", "n_tasks", "n_contacts", "n_rows", "n_levels", "broadphase", "narrowphase", "solve")})
print("tasks by type", list(prof["n_tasks_by_type"]))
c = w.get_contacts()
shape_type = np.array([s["type"] for s in spec.shapes])
bt = shape_type[np.asarray(spec.bodies["shape"])]
if bt is not None:
    ti, tj = bt[c["body_i"]], bt[c["body_j"]]
    key = np.minimum(ti, tj) * 100 + np.maximum(ti, tj)
    u, n = np.unique(key, return_counts=True)
    print("contacts by shape-type pair", dict(zip(u.tolist(), n.tolist())))
    pk = c["body_i"].astype(np.int64) * (1 << 32) + c["body_j"]
    for k in u:
        m = key == k
        print(" pair type", k, "contacts", int(m.sum()), "distinct body pairs", len(np.unique(pk[m])))
