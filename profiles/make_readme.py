#!/usr/bin/env python
"""Fills profiles/README.tmpl.md with the numbers of the committed evidence files -> profiles/README.md."""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    with open(os.path.join(HERE, name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def sci(v):
    m, e = f"{v:.2e}".split("e")
    return f"{m} × 10^{int(e)}"


b = load("r01_bench_c3.json")
ls, rf, e2e, cpu = b["last_step"], b["roofline"], b["e2e"], b["cpu_baseline"]
head = f"""| quantity | value |
|---|---|
| steps/s, ms/step (200 timed steps from the initial lattice, device time by CUDA events) | {b['steps_per_s']:.1f} steps/s, {b['ms_per_step']:.2f} ms |
| body-steps/s (`value`) | {sci(b['value'])} |
| contact-iters/s | {sci(b['contact_iters_per_s'])} |
| e2e through the C ABI with host buffers, same 200 steps on a fresh world (H2D force+torque {e2e['h2d_bytes_per_step'] / 1e6:.1f} MB, D2H position+quaternion {e2e['d2h_bytes_per_step'] / 1e6:.1f} MB per step) | {sci(e2e['value'])} body-steps/s |
| CPU oracle, {cpu['cores']} host core ({cpu['sample'].split(' (')[0]}) | {sci(cpu['value'])} body-steps/s |
| last step: pairs / contacts / rows / colours | {ls['pairs']} / {ls['contacts']} / {ls['rows']} / {ls['levels']} |
| last step stage times (ms): broadphase / narrowphase / solve (schedule, sweeps) / integrate | {ls['broadphase_ms']:.2f} / {ls['narrowphase_ms']:.2f} / {ls['solve_ms']:.2f} ({ls['schedule_ms']:.2f}, {ls['gs_ms']:.3f}) / {ls['integrate_ms']:.3f} |
| kernels launched per step | {b['gpu_launches'] // b['steps']} (one CUDA-graph replay) |"""
roof = (f"algorithmic bytes (SURVEY §8d: 408 B × {ls['contacts']} contacts × {ls['iterations']} iterations) = {rf['algorithmic_bytes_per_launch'] / 1e9:.3f} GB in "
        f"{rf['launch_ms']:.3f} ms (CUDA events in the timed run) = **{rf['achieved'] / 1e3:.2f} TB/s = {rf['frac']:.3f} of the measured {rf['peak']:.0f} GB/s**; "
        f"ncu DRAM traffic {rf['traffic'] / 1e9:.3f} GB per launch = {rf['traffic'] / 1e9 / rf['launch_ms']:.2f} TB/s (the packed rows move 84 B per row-iteration: "
        f"every row byte is read from DRAM exactly once per iteration, no re-reads).")
shares = []
for line in open(os.path.join(HERE, "r01_launches.md")):
    m = re.match(r"\| `(?:void )?([^`]+)` \| (\d+) \| ([\d.]+) \| ([\d.]+) % \|", line)
    if m and len(shares) < 6:
        shares.append(f"`{m.group(1)}` {m.group(4)} %")
rows = ["| config | solver | ms/step | steps/s | body-steps/s | contact-iters/s | CPU oracle (1 core) body-steps/s |", "|---|---|---|---|---|---|---|"]
solvers = ["reference order (bit-exact)", "colored", "reference order", "colored", "colored", "reference order", "colored (ring kernel)", "colored"]
names = {"c1": "c1 1000 spheres on a plane, Naive, 600 steps", "c2": "c2 250 × 20 box stacks, SAP, 20 it, 300 steps", "c3": "c3 100 k mixed pile on heightfield, Grid",
         "c4": "c4 4096 × 64-body jointed worlds, 1 GPU", "c5": "c5 1 M spheres in a container, sleeping on"}
tab = [json.loads(l) for l in open(os.path.join(HERE, "r01_table.jsonl")) if l.strip()]
for t, sv in zip(tab, solvers):
    c = t["config"]["workload"][:2]
    cb = t.get("cpu_baseline")
    rows.append(f"| {names[c]} | {sv} | {t['ms_per_step']:.2f} | {t['steps_per_s']:.0f} | {sci(t['value'])} | {sci(t['contact_iters_per_s'])} | {sci(cb['value']) if cb else ''} |")
sc = []
for n in (2, 4, 8):
    for name, what in ((f"r01_scale{n}_c3.json", f"{n} × c3 replicas (weak)"), (f"r01_scale{n}_c4_reference.json", f"c4 sharded over {n} GPUs, reference order (strong)"),
                       (f"r01_scale{n}_c4_colored.json", f"c4 sharded over {n} GPUs, colored (strong)")):
        p = os.path.join(HERE, name)
        if os.path.exists(p):
            j = load(name)
            sc.append(f"| {what} | {j['steps']} | {j['ms_per_step']:.2f} | {sci(j['value'])} | " + (sci(j['e2e']['value']) if j.get("e2e") else "") + " |")
scale2 = ("Multi-GPU checks (`gpurun --gpus N`, one process per GPU, no collective on the step path, `r01_scaleN_*.json`; 100 timed steps, so the c3 lines\n"
          "average over the cheaper early steps):\n\n| run | steps | ms/step | body-steps/s | e2e body-steps/s |\n|---|---|---|---|---|\n" + "\n".join(sc) + "\n\n"
          "Replicas scale linearly (no shared resource). The *fixed* 4096-world batch does not: 2.44 ms on one GPU, 1.52 ms on two, 1.09 ms on four — a\n"
          "step has a latency floor of ≈ 0.5 ms (73 dependent launches, the cooperative colouring and sweep phases; config 1 with 1000 bodies takes\n"
          "0.52 ms), and a quarter of the batch per GPU is within 2× of it. Batches that grow with the GPU count (4096 worlds *per* GPU) keep the one-GPU\n"
          "rate per GPU; lowering the floor (fewer, fused launches for small batches) is what strong scaling needs next. One GPU with a quarter / an eighth\n"
          "of the batch (`bench.py --config c4 --scale 0.25 | 0.125`): 1.02 / 0.81 ms per step; at 512 worlds broadphase 0.22, narrowphase 0.12, solve 0.57 ms\n"
          "(colouring 0.12, ring sweep 0.28 = the serial chain of one world: 10 iterations × ≈ 14 colour segments × ≈ 2 µs).") if sc else ""
c4 = [t for t, sv in zip(tab, solvers) if sv.startswith("colored (ring")][0]
c5 = tab[-1]
# ---- round 2 -----------------------------------------------------------------------------------------------------------
b2, ref2 = load("r02_bench_c3.json"), load("r02_bench_c3_reference.json")
rf2, e2, cp2, st2, ls2 = b2["roofline"], b2["e2e"], b2["cpu_baseline"], b2["stage_ms"], b2["last_step"]
tr2 = json.load(open(os.path.join(HERE, "r02_ncu_traffic.json")))["k_gs_exact"]
r02 = [f"""| quantity | value |
|---|---|
| driver command `python bench.py --gpus 1 --steps 20 --warmup 5`, settled pile (start state `bench_data/c3_settled.npz`) | {b2['steps_per_s']:.1f} steps/s, {b2['ms_per_step']:.3f} ms per step |
| body-steps/s (`value`, device time) / through the C ABI with host buffers (`e2e`) | {sci(b2['value'])} / {sci(e2['value'])} |
| contact-iters/s | {sci(b2['contact_iters_per_s'])} |
| last step: pairs / contacts / rows / colours / iterations | {ls2['pairs']} / {ls2['contacts']} / {ls2['rows']} / {ls2['levels']} / {ls2['iterations']} |
| stage times, ms (eager pass, CUDA events): broadphase / narrowphase / solve (colouring, sweep) / integrate | {st2['broadphase']:.3f} / {st2['narrowphase']:.3f} / {st2['solve']:.3f} ({st2['schedule']:.3f}, {st2['gs']:.3f}) / {st2['integrate']:.3f} |
| kernels launched per step | {b2['gpu_launches'] // b2['steps']} (one CUDA-graph replay; 76 before the single-pass scan) |
| reference arm (`--impl reference`: the oracle in reference order on the SAME 100 001-body settled state, 1 core) | {ref2['ms_per_step'] / 1e3:.2f} s per step, {sci(ref2['value'])} body-steps/s |
| `cpu_baseline` inside the product line (oracle in the colour order, {cp2['sample'].split(' oracle')[0]} steps) | {sci(cp2['value'])} body-steps/s |
| e2e ratio product / reference arm | {e2['value'] / ref2['value']:.0f} x |
| arithmetic of the timed sweep | f64 on f32-stored vectors, no FMA: bit-exact against the oracle (`dtype: "{b2['dtype']}"`) |""",
       f"""Roofline of `k_gs_exact` (one launch = {ls2['iterations']} iterations x {ls2['levels']} colours): algorithmic bytes (SURVEY 8d: 408 B x contact-iterations) =
{rf2['algorithmic_bytes_per_launch'] / 1e9:.3f} GB in {rf2['launch_ms']:.3f} ms (CUDA events, eager pass of the same run) = **{rf2['achieved'] / 1e3:.2f} TB/s = {rf2['frac']:.3f} of the measured
{rf2['peak']:.0f} GB/s**. ncu (`r02_k_gs_exact.md`): DRAM traffic {tr2['dram_bytes'] / 1e9:.3f} GB per launch in {tr2['seconds'] * 1e3:.3f} ms = {tr2['dram_bytes'] / tr2['seconds'] / 1e12:.2f} TB/s
({tr2['dram_bytes'] / tr2['seconds'] / 1e9 / rf2['peak']:.3f}): less than the algorithmic figure because the body deltas stay in L2; every row byte is read once per iteration."""]
sc2 = ["| GPUs | c3 replicas: body-steps/s (ms/step) | e2e body-steps/s | c4, 4096 worlds: 1 GPU ms | sharded ms | speed-up |", "|---|---|---|---|---|---|"]
for n in (1, 2, 4, 8):
    j = b2 if n == 1 else load(f"r02_scale{n}.json")
    c = j["c4"]
    sc2.append(f"| {n} | {sci(j['value'])} ({j['ms_per_step']:.3f}) | {sci(j['e2e']['value'])} | {c['ms_per_step_1gpu']:.3f} | "
               + (f"{c['ms_per_step_sharded']:.3f} | {c['speedup']:.2f} x |" if "speedup" in c else "- | - |"))
sh, fu = load("r02_c4_shard512.json"), load("r02_c4_full.json")
c4txt = (f"Config 4 on one GPU, exact COLORED solver (`bench.py --config c4 [--scale 0.125]`): 4096 worlds {fu['ms_per_step']:.2f} ms per step "
         f"(broadphase {fu['stage_ms']['broadphase']:.2f}, narrowphase {fu['stage_ms']['narrowphase']:.2f}, solve {fu['stage_ms']['solve']:.2f} of which the per-world sweep "
         f"{fu['stage_ms']['gs']:.2f}); one 512-world shard {sh['ms_per_step']:.2f} ms (sweep {sh['stage_ms']['gs']:.2f}). The grid-wide level sweep needed 5.06 / 1.90 ms.")
shares2 = []
for line in open(os.path.join(HERE, "r02_launches.md")):
    m = re.match(r"\| `([^`]+)` \| (\d+) \| ([\d.]+) \| ([\d.]+) % \|", line)
    if m and len(shares2) < 8:
        shares2.append(f"`{m.group(1)}` {m.group(4)} %")
rows2 = ["| config | solver | ms/step | steps/s | body-steps/s | contact-iters/s | contacts / colours | sweep: algorithmic bytes / time vs HBM peak |", "|---|---|---|---|---|---|---|---|"]
for t in (json.loads(l) for l in open(os.path.join(HERE, "r02_table.jsonl")) if l.strip()):
    wl = t["config"]["workload"]
    rows2.append(f"| {names[wl[:2]]} | {wl.split('solver=')[1]} | {t['ms_per_step']:.2f} | {t['steps_per_s']:.0f} | {sci(t['value'])} | {sci(t['contact_iters_per_s'])} | "
                 f"{t['last_step']['contacts']} / {t['last_step']['levels']} | {t['roofline']['frac']:.3f} (`{t['roofline']['kernel']}`) |")
s = open(os.path.join(HERE, "README.tmpl.md")).read()
s = s.replace("@@R02TABLE@@", "\n".join(rows2))
for k, v in (("@@R02HEAD@@", r02[0]), ("@@R02ROOF@@", r02[1]), ("@@R02SCALE@@", "\n".join(sc2)), ("@@R02C4@@", c4txt), ("@@R02SHARES@@", ", ".join(shares2) + ".")):
    s = s.replace(k, v)
for k, v in (("@@HEADLINE@@", head), ("@@ROOFLINE@@", roof), ("@@SHARES@@", ", ".join(shares) + "."), ("@@TABLE@@", "\n".join(rows)), ("@@SCALE2@@", scale2),
             ("@@C3MS@@", f"{b['ms_per_step']:.2f}"), ("@@C4MS@@", f"{c4['ms_per_step']:.2f}"), ("@@C5MS@@", f"{c5['ms_per_step']:.1f}")):
    s = s.replace(k, v)
open(os.path.join(HERE, "README.md"), "w").write(s)
print("profiles/README.md written")
