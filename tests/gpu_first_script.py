import sys, os, time, traceback
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(),'tests'))
import numpy as np
import cannon_physics_b200 as cp
from cannon_physics_b200 import _ffi, engine, scenes
import parity
oracle = _ffi.bind('oracle/libcannon_oracle.so')
F=_ffi
def run(name, spec, steps, staged=True):
    try:
        dev, ref = parity.make_pair(cp.lib, oracle, spec)
        t=time.time()
        for s in range(steps):
            if staged:
                r = parity.staged_step(dev, ref, 1/60, f"{name} step {s}")
            else:
                dev.step(1/60); ref.step(1/60)
                parity.assert_same_state(dev, ref, f"{name} step {s}")
                pa, pb = dev.profile(), ref.profile()
                assert pa['n_pairs']==pb['n_pairs'] and pa['n_contacts']==pb['n_contacts'] and pa['n_rows']==pb['n_rows'], (pa,pb)
                r=(pa['n_pairs'],pa['n_contacts'],pa['n_rows'],pa['n_levels'])
        print(f"OK   {name}: {steps} steps, last (pairs,contacts,rows..)={r} {time.time()-t:.1f}s", flush=True)
    except Exception as e:
        print(f"FAIL {name}: {type(e).__name__}: {str(e)[:600]}", flush=True)
        traceback.print_exc(limit=2)
run("c1 staged", scenes.spheres_on_plane(4,4,4), 90)
run("c1 fused", scenes.spheres_on_plane(5,5,5), 120, staged=False)
run("c2 staged", scenes.box_stacks(4, 5, grid=2), 60)
run("c2 fused", scenes.box_stacks(9, 6, grid=3), 60, staged=False)
run("c3 plane staged", scenes.mixed_pile_on_heightfield(4,4,3, with_heightfield=False, solver=F.SOLVER_REFERENCE_ORDER, grid_cells=(8,4,8)), 90)
run("c3 hf staged", scenes.mixed_pile_on_heightfield(4,4,3, hf_samples=33, solver=F.SOLVER_REFERENCE_ORDER, grid_cells=(8,4,8)), 90)
run("c4 staged", scenes.chain_worlds(3, chains=2, links=4), 60)
run("c5 fused", scenes.sphere_container(6,6,4, extent=5.0, solver=F.SOLVER_REFERENCE_ORDER), 120, staged=False)
