#!/usr/bin/env python
"""bench.py — body-steps/s of the cannon_physics step path on B200 (BASELINE.json metric).

A "step" is one World.step(dt) over the synthetic scene; the default workload is BASELINE config 3
(100k mixed sphere/box/cylinder pile on a heightfield, GridBroadphase, 10 solver iterations). With
--gpus N every rank steps its own independent world (batch of N worlds, no cross-GPU traffic on the step
path; NCCL only reduces the statistics), so scaling is "weak". `--config c4` runs the 4096 x 64-body
jointed worlds sharded across the ranks instead.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference] [--config c3|c1|c2|c4|c5]

`--impl reference` times the CPU restatement oracle (the reference is single-isolate Dart and cannot run
here) on a bounded instance of the same recipe on one host core.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

DT = 1.0 / 60.0
METRIC = "body-steps/s"


def build_spec(config: str, scale: float, rank: int, world_size: int):
    from cannon_physics_b200 import _ffi as F
    from cannon_physics_b200 import scenes
    from cannon_physics_b200.batch import shard_range
    if config == "c3":
        side = max(4, int(round(100 * scale ** 0.5)))
        spec = scenes.mixed_pile_on_heightfield(side, side, 10, seed=3 + rank, solver=F.SOLVER_COLORED)
        label = f"c3: {side}x{side}x10 mixed sphere/box/cylinder pile on a 257x257 heightfield, GridBroadphase 128x16x128, 10 it, colored GS"
    elif config == "c1":
        spec = scenes.spheres_on_plane(10, 10, 10, seed=1 + rank)
        label = "c1: 1000 spheres on a plane, NaiveBroadphase, reference-order GS 10 it"
    elif config == "c2":
        spec = scenes.box_stacks(250, 20, seed=2 + rank)
        label = "c2: 250 x 20-high box stacks, SAPBroadphase, reference-order GS 20 it"
    elif config == "c4":
        total = max(world_size, int(round(4096 * scale)))
        b, e = shard_range(total, rank, world_size)
        spec = scenes.chain_worlds(e - b, seed=4 + b)
        label = f"c4: {total} independent 64-body jointed worlds sharded over {world_size} GPU(s), NaiveBroadphase, GS 10 it"
    elif config == "c5":
        n = int(round(1_000_000 * scale))
        spec = scenes.sphere_container(n_spheres=n, seed=5 + rank, solver=F.SOLVER_COLORED)
        label = f"c5: {n}-sphere pile in a 5-plane container, sleeping on, colored GS 10 it"
    else:
        raise SystemExit(f"unknown config {config}")
    return spec, label


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
            time.sleep(0.2)

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


def ncu_traffic():
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("k_gs_dram_bytes_per_launch")
        except Exception:
            return None
    return None


def oracle_lib():
    """CPU restatement: used ONLY for cpu_baseline / --impl reference (never on the product path)."""
    from cannon_physics_b200 import _ffi
    path = os.path.join(ROOT, "oracle", "libcannon_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return _ffi.bind(path)


def time_oracle(spec, steps: int, warmup: int, budget_s: float):
    """Body-steps/s of the oracle on `spec`; stops early when the time budget is spent."""
    from cannon_physics_b200 import engine
    w = engine.DeviceWorld(oracle_lib(), spec)
    nd = n_dynamic(spec)
    for _ in range(warmup):
        w.step(DT, 1)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        w.step(DT, 1)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    el = time.perf_counter() - t0
    prof = w.profile()
    return nd * done / el, done, el, prof


def run_reference(args, rank: int, world_size: int):
    if rank != 0:
        return
    from cannon_physics_b200 import _ffi as F
    from cannon_physics_b200 import scenes
    # bounded instance of the same recipe (per-body cost of the oracle is ~linear in bodies with the grid broadphase)
    if args.config == "c3":
        spec = scenes.mixed_pile_on_heightfield(25, 25, 10, seed=3, solver=F.SOLVER_REFERENCE_ORDER)
        sample = "c3 recipe at 25x25x10 = 6250 bodies (1/16 of the 100k lattice), reference-order GS, same heightfield"
    else:
        spec, _ = build_spec(args.config, min(args.scale, 0.02 if args.config in ("c4", "c5") else 1.0), 0, 1)
        sample = f"{args.config} recipe, reduced instance {spec.name}"
    value, done, el, prof = time_oracle(spec, args.steps, args.warmup, budget_s=150.0)
    _, label = build_spec(args.config, args.scale, 0, 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "body-steps/s", "n_gpus": args.gpus, "steps": done, "warmup": args.warmup,
        "ms_per_step": 1000.0 * el / max(done, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": label},
        "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "contact_iters_per_s": prof["contact_iters_total"] / el if el > 0 else 0.0,
        "note": "the Dart reference cannot run here (no SDK) and cannot build this config at all (SURVEY.md §5.9-2,4); "
                "this is its CPU restatement on one core, the reference being single-isolate",
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", default="c3")
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the named configuration's body count")
    ap.add_argument("--solver", default="auto", choices=["auto", "colored", "reference"],
                    help="override the solver kind of the configuration (reference = GSSolver equation order, bit-reproducible)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
    from cannon_physics_b200 import engine
    from cannon_physics_b200.batch import reduce_stats

    dev_t = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev_t)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev_t)

    spec, label = build_spec(args.config, args.scale, rank, world_size)
    if args.solver != "auto":
        from cannon_physics_b200 import _ffi as F
        spec.desc["solver_kind"] = F.SOLVER_COLORED if args.solver == "colored" else F.SOLVER_REFERENCE_ORDER
        label += f" [solver={args.solver}]"
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
    world.step(DT, K)  # K steps enqueued back to back; device time by CUDA events on the library's stream
    torch.cuda.synchronize()
    if world_size > 1:
        dist.barrier()
    prof = world.profile()
    clocks = sampler.result()
    elapsed_ms = prof["step_call_ms"]
    launches = prof["kernel_launches"] - launches0
    contact_iters = prof["contact_iters_total"] - ci0

    stats = reduce_stats({"bodies": nd, "body_steps": nd * K, "contact_iters": contact_iters, "contacts": prof["n_contacts"],
                          "rows": prof["n_rows"], "steps": K}, elapsed_ms, device=dev_t)
    total_ms = stats["elapsed_ms"]
    value = stats["body_steps"] / (total_ms / 1000.0)

    # roofline of the dominant kernel (Gauss-Seidel sweeps, K5): algorithmic bytes of SURVEY.md §8d
    C_last, I_last, R_last = prof["n_contacts"], prof["iterations_done"], prof["n_rows"]
    joint_rows = max(0, R_last - 3 * C_last)
    alg_bytes = (408.0 * C_last + 200.0 * joint_rows) * max(I_last, 1)
    gs_ms = prof["gs_ms"]
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (gs_ms / 1000.0) / 1e9 if gs_ms > 0 else 0.0
    from cannon_physics_b200 import _ffi as _F
    colored = spec.desc.get("solver_kind") == _F.SOLVER_COLORED
    gs_kernel = ("k_gs_world_ring" if spec.desc.get("n_worlds", 1) > 1 else "k_gs_fast") if colored else "k_gs"
    # the ncu DRAM figure in profiles/ncu_traffic.json was captured for k_gs_fast on the default workload only
    roofline = {"bound": "hbm", "kernel": gs_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic() if (gs_kernel == "k_gs_fast" and args.config == "c3" and args.scale == 1.0) else None,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": gs_ms,
                "share_of_step": gs_ms / (total_ms / K) if total_ms > 0 else None}

    # end to end through the public C ABI with host buffers: per step upload force/torque, step, download poses
    e2e = None
    if not args.no_e2e:
        n = spec.n_bodies
        # pinned host buffers for the per-step inputs (force, torque) and outputs (position, quaternion)
        pin = lambda shape: torch.zeros(shape, dtype=torch.float32, pin_memory=True).numpy()
        force, torque = pin((n, 3)), pin((n, 3))
        out = {"position": pin((n, 3)), "quaternion": pin((n, 4))}
        ke = K
        # same trajectory segment as the device-timed run: a fresh world, W untimed warm-up steps, K timed steps
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

    cpu_baseline = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        # the oracle continues from the device state at the end of the run (contact-rich regime), bounded to ~20 s
        from cannon_physics_b200 import _ffi as F
        st = world.get_bodies(("position", "quaternion", "velocity", "angular_velocity", "sleep_state"))
        ospec, _ = build_spec(args.config, args.scale, rank, world_size)
        for k in ("position", "quaternion", "velocity", "angular_velocity", "sleep_state"):
            ospec.bodies[k] = st[k]
        ospec.desc = dict(ospec.desc, solver_kind=F.SOLVER_REFERENCE_ORDER)
        v, done, el, _ = time_oracle(ospec, steps=50, warmup=0, budget_s=20.0)
        cpu_baseline = {"value": v, "unit": "body-steps/s", "cores": 1, "kind": "port",
                        "sample": f"{done} oracle step(s) continuing from the device state after the timed run ({el:.1f} s, 1 core; "
                                  f"the Dart reference is single-isolate)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "body-steps/s", "n_gpus": world_size, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong" if args.config == "c4" else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": label, "bodies_per_gpu": spec.n_bodies, "dt": DT,
                       "l2": "body + row state exceeds the 126 MB L2 at this size; no explicit flush" if spec.n_bodies >= 50000 else
                             "working set smaller than L2 (latency-bound configuration, see DESIGN.md)"},
            "steps_per_s": K / (total_ms / 1000.0),
            "contact_iters_per_s": stats["contact_iters"] / (total_ms / 1000.0),
            "last_step": {"pairs": prof["n_pairs"], "contacts": C_last, "rows": R_last, "levels": prof["n_levels"], "iterations": I_last,
                          "broadphase_ms": prof["broadphase"], "narrowphase_ms": prof["narrowphase"], "solve_ms": prof["solve"],
                          "schedule_ms": prof["schedule_ms"], "gs_ms": gs_ms, "integrate_ms": prof["integrate"]},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
