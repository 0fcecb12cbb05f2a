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
