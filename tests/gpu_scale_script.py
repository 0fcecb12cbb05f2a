import sys, os, time, json
sys.path.insert(0, os.getcwd())
import numpy as np
import cannon_physics_b200 as cp
from cannon_physics_b200 import _ffi as F, engine, scenes
def run(name, spec, chunks, per=20):
    t=time.time()
    try:
        w = engine.DeviceWorld(cp.lib, spec)
        print(f"{name}: n={spec.n_bodies} setup {time.time()-t:.1f}s", flush=True)
        for c in range(chunks):
            w.step(1/60, per)
            p = w.profile()
            b = w.get_bodies(("position","velocity"))
            y = b["position"][:,1]
            print(f"  steps={p['steps']} ms/step={p['step_call_ms']/per:.3f} pairs={p['n_pairs']} contacts={p['n_contacts']} rows={p['n_rows']} levels={p['n_levels']} it={p['iterations_done']} | bp={p['broadphase']:.3f} np={p['narrowphase']:.3f} solve={p['solve']:.3f} (sched={p['schedule_ms']:.3f} gs={p['gs_ms']:.3f}) int={p['integrate']:.3f} | ymin={y[1:].min():.2f} ymax={y[1:].max():.2f} vmax={np.abs(b['velocity']).max():.2f} launches={p['kernel_launches']} tasks={p['n_tasks']} {p['n_tasks_by_type']}", flush=True)
    except Exception as e:
        print(f"FAIL {name}: {e}", flush=True)
which = sys.argv[1] if len(sys.argv)>1 else "all"
if which in ("all","c3s"):
    run("c3 10k colored", scenes.mixed_pile_on_heightfield(32,32,10), 6)
    run("c3 10k reforder", scenes.mixed_pile_on_heightfield(32,32,10, solver=F.SOLVER_REFERENCE_ORDER), 6)
if which in ("all","c3"):
    run("c3 100k colored", scenes.mixed_pile_on_heightfield(100,100,10), 8, per=25)
if which in ("all","c1"):
    run("c1 1000", scenes.spheres_on_plane(10,10,10), 6, per=100)
if which in ("all","c2"):
    run("c2 5000", scenes.box_stacks(250,20), 4, per=25)
if which in ("all","c4"):
    run("c4 512 worlds", scenes.chain_worlds(512), 4, per=50)
if which in ("all","c5"):
    run("c5 200k", scenes.sphere_container(n_spheres=200000), 4, per=25)
