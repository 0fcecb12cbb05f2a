"""Small colored-solver runs for compute-sanitizer racecheck (ring kernel, staged sweep, split SAT kernel); not a test."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cannon_physics_b200 as cp  # noqa: E402
from cannon_physics_b200 import _ffi as F  # noqa: E402
from cannon_physics_b200 import engine, scenes  # noqa: E402

for name, spec, steps in (("ring: 6 jointed worlds", scenes.chain_worlds(6, chains=2, links=6), 40),
                          ("k_gs_fast + SAT launches: 4x4x3 pile on a heightfield", scenes.mixed_pile_on_heightfield(4, 4, 3, hf_samples=33, grid_cells=(8, 4, 8)), 60)):
    spec.desc["solver_kind"] = F.SOLVER_COLORED
    w = engine.DeviceWorld(cp.lib, spec, device=0)
    w.step(1 / 60, steps)
    p = w.profile()
    print(name, "steps", steps, "contacts", p["n_contacts"], "rows", p["n_rows"], "levels", p["n_levels"], flush=True)
