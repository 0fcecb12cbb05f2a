#!/usr/bin/env python
"""bench.py — body-steps/s of the cannon_physics step path on B200 (BASELINE.json metric).

A "step" is one World.step(dt) over the synthetic scene. The default workload is BASELINE config 3 (100k mixed
sphere/box/cylinder pile on a 257x257 heightfield, GridBroadphase, 10 solver iterations) in its *settled* regime: both
arms start from bench_data/c3_settled.npz, the state 250 steps after the recipe's hanging lattice (tools/make_settled.py;
~4.3e5 contacts), so the timed steps are pile steps whatever --steps / --warmup are. With --gpus N every rank steps its
own replica of that world (a single large world does not shard): scaling "weak", no data-path collective,
NCCL only reduces the statistics. The line also carries `c4`: the strong-scaling record of BASELINE config 4 (4096
independent 64-body jointed worlds): rank 0 steps all worlds alone, then every rank steps its shard.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference] [--config c3|c1|c2|c4|c5]

`--impl reference` times the reference's CPU algorithm — its restatement in oracle/ (the reference is single-isolate Dart
and cannot run here) — on the same bodies, the same settled state and the same configuration, on one host
core. That arm never imports the package __init__ (which loads libcannon_cuda.so).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

DT = 1.0 / 60.0
METRIC = "body-steps/s"
SETTLED = os.path.join(ROOT, "bench_data", "c3_settled.npz")
STATE_FIELDS = ("position", "quaternion", "velocity", "angular_velocity", "sleep_state")


def host_modules():
    """_ffi / engine / scenes / batch WITHOUT running cannon_physics_b200/__init__.py (which loads the CUDA library):
    the reference arm must not map libcannon_cuda.so. In the product arm the package is already imported and these are
    just its submodules."""
    if "cannon_physics_b200" not in sys.modules:
        pkg = types.ModuleType("cannon_physics_b200")
        pkg.__path__ = [os.path.join(ROOT, "cannon_physics_b200")]
        sys.modules["cannon_physics_b200"] = pkg
    mods = {n: importlib.import_module("cannon_physics_b200." + n) for n in ("_ffi", "engine", "scenes", "batch")}
    return mods["_ffi"], mods["engine"], mods["scenes"], mods["batch"]


def load_settled(spec):
    """Replace the recipe's initial lattice by the committed settled state (same bodies, same order)."""
    if not os.path.exists(SETTLED):
        raise SystemExit(f"{SETTLED} is missing: run tools/make_settled.py on a GPU box (bench.py does not fall back to the free-fall start)")
    st = np.load(SETTLED)
    assert st["position"].shape[0] == spec.n_bodies, "settled snapshot does not match the c3 recipe"
    for k in STATE_FIELDS:
        spec.bodies[k] = np.ascontiguousarray(st[k])
    return int(st["steps"])


def build_spec(config: str, scale: float, rank: int, world_size: int, solver: str = "auto", lattice: bool = False):
    F, engine, scenes, batch = host_modules()
    kinds = {"colored": F.SOLVER_COLORED, "reference": F.SOLVER_REFERENCE_ORDER, "colored_f32": F.SOLVER_COLORED_F32, "split": F.SOLVER_SPLIT}
    state = "recipe start"
    if config == "c3":
        side = max(4, int(round(100 * scale ** 0.5)))
        # replicas: every rank steps the same world (seed 3) so that all ranks time the same settled state
        spec = scenes.mixed_pile_on_heightfield(side, side, 10, seed=3, solver=F.SOLVER_COLORED)
        label = f"c3: {side}x{side}x10 mixed sphere/box/cylinder pile on a 257x257 heightfield, GridBroadphase 128x16x128, 10 it"
        if side == 100 and not lattice:
            n0 = load_settled(spec)
            state = f"settled pile: bench_data/c3_settled.npz ({n0} steps after the lattice start)"
    elif config == "c1":
        spec = scenes.spheres_on_plane(10, 10, 10, seed=1 + rank)
        label = "c1: 1000 spheres on a plane, NaiveBroadphase, GS 10 it"
    elif config == "c2":
        spec = scenes.box_stacks(250, 20, seed=2 + rank)
        label = "c2: 250 x 20-high box stacks, SAPBroadphase, GS 20 it"
    elif config == "c4":
        total = max(world_size, int(round(4096 * scale)))
        b, e = batch.shard_range(total, rank, world_size)
        spec = scenes.chain_worlds(e - b, seed=4 + b, solver=F.SOLVER_COLORED)
        label = f"c4: {total} independent 64-body jointed worlds sharded over {world_size} GPU(s), NaiveBroadphase, GS 10 it"
    elif config == "c5":
        n = int(round(1_000_000 * scale))
        spec = scenes.sphere_container(n_spheres=n, seed=5 + rank, solver=F.SOLVER_COLORED)
        label = f"c5: {n}-sphere pile in a 5-plane container, sleeping on, GS 10 it"
    else:
        raise SystemExit(f"unknown config {config}")
    if solver != "auto":
        spec.desc["solver_kind"] = kinds[solver]
    names = {v: k for k, v in kinds.items()}
    label += f", solver={names[spec.desc['solver_kind']]}"
    return spec, label, state


def solver_dtype(spec, F) -> str:
    # the arithmetic type of the sweep: COLORED / REFERENCE_ORDER / SPLIT compute in f64 on f32-stored vectors exactly like the
    # Dart VM; COLORED_F32 packs the rows to f32 and uses FMA
    return "f32" if spec.desc.get("solver_kind") == F.SOLVER_COLORED_F32 else "f64"


def n_dynamic(spec) -> int:
    return int(np.count_nonzero(spec.bodies["mass"] > 0))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        self.stop_flag = True
        self.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def oracle_lib():
    """CPU restatement: used ONLY for cpu_baseline / --impl reference (never on the product path)."""
    F, _, _, _ = host_modules()
    path = os.path.join(ROOT, "oracle", "libcannon_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return F.bind(path)


def time_oracle(spec, steps: int, warmup: int, budget_s: float):
    """Body-steps/s of the oracle on `spec` (one core); stops early when the time budget is spent."""
    _, engine, _, _ = host_modules()
    w = engine.DeviceWorld(oracle_lib(), spec)
    nd = n_dynamic(spec)
    for _ in range(warmup):
        w.step(DT, 1)
    ci0 = w.profile()["contact_iters_total"]
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        w.step(DT, 1)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    el = time.perf_counter() - t0
    prof = w.profile()
    prof["contact_iters_timed"] = prof["contact_iters_total"] - ci0
    return nd * done / el, done, el, prof


def run_reference(args, rank: int, world_size: int):
    """The reference's own algorithm (GSSolver insertion order, f64 on f32 stores) on the SAME workload: same bodies, same
    settled state, same broadphase and iteration count. One host core: the reference is single-isolate."""
    if rank != 0:
        return
    F, _, _, _ = host_modules()
    scale = args.scale
    if args.config in ("c4", "c5") and not args.full_reference:
        scale = min(scale, 0.02)  # 82 worlds / 20k spheres: bounded samples of configurations this arm is not the headline for
    spec, label, state = build_spec(args.config, scale, 0, 1, solver="reference", lattice=args.lattice)
    # a settled c3 step costs ~7 s on one core: K timed steps after at most 2 warm-up steps stay within a few minutes
    wu = min(args.warmup, 2)
    value, done, el, prof = time_oracle(spec, args.steps, wu, budget_s=args.reference_budget)
    same = scale == args.scale
    sample = (f"{done} step(s) of the full workload ({spec.n_bodies} bodies, {state}), reference-order GSSolver, {el:.1f} s" if same else
              f"{done} step(s) of the recipe reduced to {spec.n_bodies} bodies ({state})")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "body-steps/s", "n_gpus": args.gpus, "steps": done, "warmup": wu,
        "ms_per_step": 1000.0 * el / max(done, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": label, "state": state, "bodies_per_gpu": spec.n_bodies, "dt": DT, "same_config": same},
        "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "contact_iters_per_s": prof["contact_iters_timed"] / el if el > 0 else 0.0,
        "last_step": {"pairs": prof["n_pairs"], "contacts": prof["n_contacts"], "rows": prof["n_rows"], "iterations": prof["iterations_done"]},
        "note": "the Dart reference cannot run here (no SDK) and cannot build this config at all (SURVEY.md §5.9-2,4); this is its CPU "
                "restatement (oracle/, kind 'port') on one core, the reference being single-isolate",
    }
    print(json.dumps(line), flush=True)


def c4_record(cp, engine, args, rank, world_size, local_rank, dist, torch, dev_t):
    """Strong scaling of BASELINE config 4: 4096 independent 64-body jointed worlds. Rank 0 steps all of them alone; then
    every rank steps its contiguous shard (no data-path collective); time = max over ranks."""
    from cannon_physics_b200.batch import reduce_stats
    total = max(world_size, int(round(4096 * args.c4_scale)))
    K, W = args.c4_steps, 20
    out = {"worlds": total, "bodies_per_world": 64, "steps": K, "warmup": W, "n_gpus": world_size}

    def run(spec):
        w = engine.DeviceWorld(cp.lib, spec, device=local_rank)
        w.step(DT, W)
        torch.cuda.synchronize()
        w.step(DT, K)
        p = w.profile()
        ms = p["step_call_ms"] / K
        last = {"contacts": p["n_contacts"], "rows": p["n_rows"], "levels": p["n_levels"], "iterations": p["iterations_done"]}
        del w
        return ms, last

    if world_size > 1:
        dist.barrier()
    if rank == 0:
        spec, label, _ = build_spec("c4", args.c4_scale, 0, 1)
        out["ms_per_step_1gpu"], out["last_step_1gpu"] = run(spec)
        out["workload"] = label
    if world_size > 1:
        dist.barrier()
        spec, _, _ = build_spec("c4", args.c4_scale, rank, world_size)
        ms, _ = run(spec)
        st = reduce_stats({"worlds": spec.desc["n_worlds"]}, ms, device=dev_t)
        out["ms_per_step_sharded"] = st["elapsed_ms"]
        if rank == 0:
            out["speedup"] = out["ms_per_step_1gpu"] / out["ms_per_step_sharded"]
    if rank == 0:
        ms = out.get("ms_per_step_sharded", out["ms_per_step_1gpu"])
        out["world_steps_per_s"] = total / (ms / 1000.0)
        out["body_steps_per_s"] = total * 63 / (ms / 1000.0)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", default="c3")
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the named configuration's body count")
    ap.add_argument("--solver", default="auto", choices=["auto", "colored", "reference", "colored_f32", "split"],
                    help="override the solver kind (auto = colored: GSSolver arithmetic in colour order, bit-exact against the oracle)")
    ap.add_argument("--lattice", action="store_true", help="c3 from the recipe's hanging lattice instead of the settled snapshot")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="skip the config-4 strong-scaling record")
    ap.add_argument("--c4-scale", type=float, default=1.0)
    ap.add_argument("--c4-steps", type=int, default=100)
    ap.add_argument("--reference-budget", type=float, default=420.0, help="wall-clock budget of the reference arm's timed steps, seconds")
    ap.add_argument("--full-reference", action="store_true", help="reference arm at full size also for c4 / c5")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world_size)
        return

    import torch  # plumbing only: process group + barrier; the step path never touches torch
    import torch.distributed as dist

    import cannon_physics_b200 as cp
    from cannon_physics_b200 import _ffi as F
    from cannon_physics_b200 import engine
    from cannon_physics_b200.batch import reduce_stats

    dev_t = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev_t)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev_t)

    spec, label, state = build_spec(args.config, args.scale, rank, world_size, solver=args.solver, lattice=args.lattice)
    nd = n_dynamic(spec)
    world = engine.DeviceWorld(cp.lib, spec, device=local_rank)

    W = max(args.warmup, 3)
    K = args.steps
    world.step(DT, W)
    launches0 = world.profile()["kernel_launches"]
    ci0 = world.profile()["contact_iters_total"]

    sampler = ClockSampler(local_rank)
    sampler.start()
    if world_size > 1:
        dist.barrier()
    torch.cuda.synchronize()
    world.step(DT, K)  # K steps enqueued back to back (graph replays); device time by CUDA events on the library's stream
    torch.cuda.synchronize()
    if world_size > 1:
        dist.barrier()
    prof = world.profile()
    elapsed_ms = prof["step_call_ms"]
    launches = prof["kernel_launches"] - launches0
    contact_iters = prof["contact_iters_total"] - ci0

    stats = reduce_stats({"bodies": nd, "body_steps": nd * K, "contact_iters": contact_iters, "contacts": prof["n_contacts"],
                          "rows": prof["n_rows"], "steps": K}, elapsed_ms, device=dev_t)
    total_ms = stats["elapsed_ms"]
    value = stats["body_steps"] / (total_ms / 1000.0)

    # roofline of the dominant kernel (the Gauss-Seidel sweep, K5). Its launch duration is measured live over a second pass of
    # K steps in which every step is launched eagerly between its own CUDA events on the library's stream
    # (cannon_world_step_profiled); algorithmic bytes = SURVEY.md §8d: 408 B per contact-iteration + 200 B per joint-row-iteration.
    world.step_profiled(DT, K)
    pp = world.profile()
    clocks = sampler.result()
    ci_prof = pp["contact_iters_total"] - prof["contact_iters_total"]
    C_last, I_last, R_last = pp["n_contacts"], pp["iterations_done"], pp["n_rows"]
    joint_rows = max(0, R_last - 3 * C_last)
    alg_bytes = 408.0 * ci_prof + 200.0 * joint_rows * max(I_last, 1) * K
    gs_ms = pp["sum_gs"] / max(pp["sum_steps"], 1)
    peak, peak_src = measured_peak()
    achieved = alg_bytes / K / (gs_ms / 1000.0) / 1e9 if gs_ms > 0 else 0.0
    kind = spec.desc.get("solver_kind")
    if kind == F.SOLVER_COLORED_F32:
        gs_kernel = "k_gs_world_ring" if spec.desc.get("n_worlds", 1) > 1 else "k_gs_fast"
    elif kind == F.SOLVER_COLORED:
        gs_kernel = "k_gs_world_exact" if spec.desc.get("n_worlds", 1) > 1 else "k_gs_exact"
    else:
        gs_kernel = "k_gs"
    step_ms_prof = pp["sum_step_ms"] / max(pp["sum_steps"], 1)
    roofline = {"bound": "hbm", "kernel": gs_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None,  # dram__bytes per launch comes from an ncu capture (profiles/), never from inside a timed run
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes / K, "launch_ms": gs_ms,
                "share_of_step": gs_ms / step_ms_prof if step_ms_prof > 0 else None,
                "how": f"{K} eager steps after the timed region, CUDA events around every launch on the library's stream"}
    stages = {k: pp["sum_" + k] / max(pp["sum_steps"], 1) for k in ("broadphase", "narrowphase", "solve", "schedule", "gs", "integrate", "step_ms")}

    # end to end through the public C ABI with host buffers: per step upload force/torque, step, download poses
    e2e = None
    if not args.no_e2e:
        n = spec.n_bodies
        # pinned host buffers for the per-step inputs (force, torque) and outputs (position, quaternion)
        pin = lambda shape: torch.zeros(shape, dtype=torch.float32, pin_memory=True).numpy()
        force, torque = pin((n, 3)), pin((n, 3))
        out = {"position": pin((n, 3)), "quaternion": pin((n, 4))}
        ke = K
        # same trajectory segment as the device-timed run: a fresh world from the same state, W untimed warm-up steps, K timed steps
        world_e = engine.DeviceWorld(cp.lib, spec, device=local_rank)
        for _ in range(W):
            world_e.update_bodies(0, n, force=force, torque=torque)
            world_e.step(DT, 1)
            world_e.get_bodies(("position", "quaternion"), out=out)
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(ke):
            world_e.update_bodies(0, n, force=force, torque=torque)
            world_e.step(DT, 1)
            poses = world_e.get_bodies(("position", "quaternion"), out=out)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        del world_e
        e2e_stats = reduce_stats({"body_steps": nd * ke}, el * 1000.0, device=dev_t)
        e2e = {"value": e2e_stats["body_steps"] / (e2e_stats["elapsed_ms"] / 1000.0), "unit": "body-steps/s",
               "h2d_bytes_per_step": int(force.nbytes + torque.nbytes), "d2h_bytes_per_step": int(poses["position"].nbytes + poses["quaternion"].nbytes),
               "steps": ke}

    c4 = None
    if args.config == "c3" and not args.no_c4:
        c4 = c4_record(cp, engine, args, rank, world_size, local_rank, dist, torch, dev_t)

    cpu_baseline = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        # the oracle (same colour order, same arithmetic) on the state the timed run started from, bounded to ~25 s on one core
        ospec, _, _ = build_spec(args.config, args.scale, rank, world_size, solver=args.solver, lattice=args.lattice)
        v, done, el, _ = time_oracle(ospec, steps=50, warmup=0, budget_s=25.0)
        cpu_baseline = {"value": v, "unit": "body-steps/s", "cores": 1, "kind": "port",
                        "sample": f"{done} oracle step(s) of the same workload from the same start state ({el:.1f} s, 1 core; the Dart reference is single-isolate)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "body-steps/s", "n_gpus": world_size, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong" if args.config == "c4" else "weak",
            "vs_baseline": None, "dtype": solver_dtype(spec, F), "data": "synthetic",
            "config": {"workload": label, "state": state, "bodies_per_gpu": spec.n_bodies, "dt": DT,
                       "l2": "body + row state exceeds the 126 MB L2 at this size; no explicit flush" if spec.n_bodies >= 50000 else
                             "working set smaller than L2 (latency-bound configuration, see DESIGN.md)"},
            "steps_per_s": K / (total_ms / 1000.0),
            "contact_iters_per_s": stats["contact_iters"] / (total_ms / 1000.0),
            "last_step": {"pairs": prof["n_pairs"], "contacts": prof["n_contacts"], "rows": prof["n_rows"], "levels": prof["n_levels"],
                          "iterations": prof["iterations_done"]},
            "stage_ms": stages,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "c4": c4, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
