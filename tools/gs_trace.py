"""Profiling aid (not a test): per-phase work / barrier-wait cycles of the colored sweep (CANNON_GS_TRACE).

usage (GPU box): python tools/gs_trace.py [steps] [config]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
path = os.path.join(ROOT, "gpurun_out", "gs_trace.bin")
os.makedirs(os.path.dirname(path), exist_ok=True)
os.environ["CANNON_GS_TRACE"] = path
import bench  # noqa: E402
import cannon_physics_b200 as cp  # noqa: E402
from cannon_physics_b200 import engine  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 120
config = sys.argv[2] if len(sys.argv) > 2 else "c3"
spec, label, state = bench.build_spec(config, 1.0, 0, 1, solver=sys.argv[3] if len(sys.argv) > 3 else "auto")
w = engine.DeviceWorld(cp.lib, spec, device=0)
w.step(1 / 60, steps)
prof = w.profile()
print({k: prof[k] for k in ("n_contacts", "n_rows", "n_levels", "iterations_done", "gs_ms", "schedule_ms", "solve")})
raw = np.fromfile(path, dtype=np.int64)
nsm = (len(raw)) // (4 * 64 * 2 + 64)
t = raw[: nsm * 4 * 64 * 2].reshape(-1, 64, 2)
raw_tail_base = None
grid = int(os.environ.get('GS_TRACE_CTAS', '148'))
t = t[:grid]
raw_tail_base = grid * 64 * 2
print("CTAs", t.shape[0])
nph = min(64, prof["n_levels"] * prof["iterations_done"])
ghz = 1.9
cntw, cntr = t[:, :, 0] >> 32, t[:, :, 1] >> 32  # windows / row steps per CTA and phase (k_gs_exact)
t = t & 0xffffffff
print("phase  work(us): mean   max | wait(us): min  mean | phase total(us) of CTA0 | windows: sum max/CTA | row steps: sum max/CTA")
for ph in range(nph):
    wk, wt = t[:, ph, 0] / ghz / 1e3, t[:, ph, 1] / ghz / 1e3
    print(f"{ph:4d}  {wk.mean():8.2f} {wk.max():8.2f} | {wt.min():8.2f} {wt.mean():8.2f} | {wk[0] + wt[0]:8.2f} | {cntw[:, ph].sum():6d} {cntw[:, ph].max():4d} | {cntr[:, ph].sum():7d} {cntr[:, ph].max():5d}")
tot = (t[:, :nph, 0] + t[:, :nph, 1]).sum(1) / ghz / 1e3
print("sum over traced phases (us): CTA mean", tot.mean(), " => per phase", tot.mean() / nph)
rows = w.get_rows()
lv = rows["level"]
key = rows["body_i"].astype(np.int64) * (1 << 32) + rows["body_j"]
newunit = np.ones(len(lv), bool)
newunit[1:] = (key[1:] != key[:-1]) | (lv[1:] != lv[:-1])
print("rows per level ", np.bincount(lv).tolist())
print("units per level", np.bincount(lv[newunit]).tolist())
sizes = np.diff(np.flatnonzero(np.append(newunit, True)))
print("rows per unit histogram", np.bincount(sizes).tolist())

ncta = t.shape[0]
tk = raw[raw_tail_base: raw_tail_base + ncta * 16].reshape(ncta, 4, 4)
print("warp 0 of the first CTAs, iteration 1 colour 0: per task (wait us, solve us, of which body-lambda gather us, units*1000+rows)")
for c in range(0, min(ncta, 148), 12):
    print(c, [(round(a / ghz / 1e3, 2), round(b / ghz / 1e3, 2), round(cc / ghz / 1e3, 2), int(d)) for a, b, cc, d in tk[c] if d > 0])
