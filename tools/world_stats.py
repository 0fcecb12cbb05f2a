"""Profiling aid (not a test): per-world unit / row / colour statistics of a colored batch (sizes k_gs_world_ring's tables).

usage (GPU box): python tools/world_stats.py [steps] [n_worlds]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cannon_physics_b200 as cp  # noqa: E402
from cannon_physics_b200 import _ffi as F  # noqa: E402
from cannon_physics_b200 import engine, scenes  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 512
spec = scenes.chain_worlds(nw)
spec.desc["solver_kind"] = F.SOLVER_COLORED
w = engine.DeviceWorld(cp.lib, spec, device=0)
w.step(1 / 60, steps)
rows = w.get_rows()
per_world = 64
wd = rows["body_i"] // per_world
key = rows["body_i"].astype(np.int64) * (1 << 32) + rows["body_j"]
lv = rows["level"]
newunit = np.ones(len(lv), bool)
newunit[1:] = (key[1:] != key[:-1]) | (lv[1:] != lv[:-1])
print("rows per world: mean %.0f max %d" % (np.bincount(wd).mean(), np.bincount(wd).max()))
print("units per world: mean %.0f max %d" % (np.bincount(wd[newunit]).mean(), np.bincount(wd[newunit]).max()))
sizes = np.diff(np.flatnonzero(np.append(newunit, True)))
print("rows per unit: max %d, histogram %s" % (sizes.max(), np.bincount(sizes).tolist()))
wl = wd.astype(np.int64) * 1024 + lv
print("rows per (world, colour): mean %.1f max %d; colours: max %d" % (np.bincount(wl)[np.bincount(wl) > 0].mean(), np.bincount(wl).max(), lv.max() + 1))
