"""Profiling aid: run c3 to the contact-rich regime normally, then time one more step under the current CANNON_NP_DEBUG."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["CANNON_NO_GRAPH"] = "1"  # the debug switch is read per launch
import bench  # noqa: E402
import cannon_physics_b200 as cp  # noqa: E402
from cannon_physics_b200 import engine  # noqa: E402

spec, label = bench.build_spec("c3", 1.0, 0, 1)
w = engine.DeviceWorld(cp.lib, spec, device=0)
w.step(1 / 60, int(sys.argv[1]) if len(sys.argv) > 1 else 220)
state = w.get_bodies()
for d in ("0", "1", "2", "3", "0"):
    os.environ["CANNON_NP_DEBUG"] = d
    w.update_bodies(0, len(state["position"]), position=state["position"], quaternion=state["quaternion"], velocity=state["velocity"], angular_velocity=state["angular_velocity"])
    w.step(1 / 60, 1)
    prof = w.profile()
    print(d, {k: round(prof[k], 3) for k in ("broadphase", "narrowphase", "solve", "gs_ms")}, prof["n_contacts"], flush=True)

