"""world_size-2 gloo test of the batch sharding + statistics reduction used by bench.py --gpus N (CPU only).
The worlds of a batch are independent, so per-world results must not depend on how the batch is sharded."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from cannon_physics_b200 import _ffi, engine, scenes
    from cannon_physics_b200.batch import reduce_stats, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    lib = _ffi.bind(os.path.join(ROOT, "oracle", "libcannon_oracle.so"))  # host-logic test: the checker stands in for a GPU
    total = 5
    b, e = shard_range(total, rank, world_size)
    w = engine.DeviceWorld(lib, scenes.chain_worlds(e - b, chains=2, links=4, seed=4 + b))
    w.step(1 / 60, 30)
    st = w.get_bodies(("position",))
    np.save(os.path.join(out_dir, f"pos_{rank}.npy"), st["position"])
    stats = reduce_stats({"bodies": w.n, "body_steps": w.n * 30, "steps": 30}, elapsed_ms=10.0 * (rank + 1))
    if rank == 0:
        np.save(os.path.join(out_dir, "stats.npy"), np.array([stats["bodies"], stats["body_steps"], stats["elapsed_ms"]]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_batch_matches_single_process(tmp_path, oracle_lib):
    from cannon_physics_b200 import engine, scenes
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    whole = engine.DeviceWorld(oracle_lib, scenes.chain_worlds(5, chains=2, links=4))
    whole.step(1 / 60, 30)
    ref = whole.get_bodies(("position",))["position"]
    got = np.concatenate([np.load(tmp_path / "pos_0.npy"), np.load(tmp_path / "pos_1.npy")])
    assert np.array_equal(got, ref)  # bit-identical per world for any shard count
    bodies, body_steps, elapsed = np.load(tmp_path / "stats.npy")
    assert bodies == whole.n and body_steps == whole.n * 30 and elapsed == 20.0  # sum of work, max of time


def test_cannon_batch_abi_shards_from_one_thread(oracle_lib):
    """cannon_batch_* (include/cannon_cuda.h): one host thread, G shards. The entry points are host glue over the per-world
    ABI and live in both libraries, so the checker library exercises the whole path on the CPU: sharding of bodies and
    constraints (global -> shard-local indices), stepping, gathered download, statistics. Per-world results must be
    bit-identical for every shard count, and equal to one cannon_world holding the whole batch."""
    from cannon_physics_b200 import _ffi as F, engine, scenes
    fields = ("position", "quaternion", "velocity", "angular_velocity", "sleep_state", "world_id")
    for solver in (F.SOLVER_REFERENCE_ORDER, F.SOLVER_COLORED):  # COLORED hashes world-local unit keys: shard-independent too
      spec = scenes.chain_worlds(7, chains=2, links=4, solver=solver)
      whole = engine.DeviceWorld(oracle_lib, spec)
      whole.step(1 / 60, 25)
      ref = whole.get_bodies(fields)
      for devices in ((0,), (0, 0), (0, 0, 0)):
          b = engine.DeviceBatch(oracle_lib, spec, devices=devices)
          b.step(1 / 60, 20)
          b.step(1 / 60, 5)
          got = b.get_bodies(fields)
          for k in fields:
              assert np.array_equal(got[k], ref[k]), (devices, k)
          st = b.stats()
          assert (st["n_gpus"], st["n_worlds"], st["steps"]) == (len(devices), 7, 25) and sum(st["gpu_worlds"]) == 7
          assert st["n_contacts"] == whole.profile()["n_contacts"] and st["n_rows"] == whole.profile()["n_rows"]
          first, n, h = b.shard(len(devices) - 1)
          assert first + n == 7 and h.value
          b.close()
    # a constraint across two worlds cannot be sharded
    bad = scenes.chain_worlds(2, chains=2, links=4)
    bad.constraints[0] = dict(bad.constraints[0], body_b=bad.n_bodies - 1)
    try:
        engine.DeviceBatch(oracle_lib, bad, devices=(0, 0))
        raise AssertionError("expected CANNON_E_INVALID")
    except F.CannonError as e:
        assert e.code == F.E_INVALID and "one world" in str(e)
