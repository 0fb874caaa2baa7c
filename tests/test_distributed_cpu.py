"""CPU, world_size 2, gloo: the multi-GPU plumbing (weight broadcast, utterance sharding, result gather)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import b200tts  # noqa: F401
    from b200tts import distributed
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    state = {"a.weight": rng.standard_normal((3, 5, 7), dtype=np.float32), "b": rng.standard_normal(130, dtype=np.float32),
             "s": np.float32(2.5).reshape(())} if rank == 0 else None
    manifest, flat = distributed.broadcast_packed(state, src=0)
    got = {n: flat[o:o + int(np.prod(s, dtype=np.int64))].reshape(s).numpy().copy() for n, s, o in manifest}
    shards = distributed.shard_utterances([1126 ** 2, 752 ** 2, 1502 ** 2, 900 ** 2, 1000 ** 2], world)
    res = distributed.gather_objects({"rank": rank, "mine": shards[rank]}, dst=0)
    q.put((rank, {k: v.tolist() for k, v in got.items()}, shards, res))
    dist.destroy_process_group()


def test_broadcast_shard_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=120) for _ in procs], key=lambda o: o[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    want = {"a.weight": rng.standard_normal((3, 5, 7), dtype=np.float32), "b": rng.standard_normal(130, dtype=np.float32)}
    for rank, got, shards, res in outs:
        np.testing.assert_array_equal(np.array(got["a.weight"], dtype=np.float32), want["a.weight"])
        np.testing.assert_array_equal(np.array(got["b"], dtype=np.float32), want["b"])
        assert got["s"] == 2.5
        assert sorted(shards[0] + shards[1]) == [0, 1, 2, 3, 4] and shards == outs[0][2]
    assert outs[0][3] is not None and [r["rank"] for r in outs[0][3]] == [0, 1]
    assert outs[1][3] is None


def test_shard_balance_single_process():
    import b200tts  # noqa: F401
    from b200tts import distributed
    costs = [float(n) ** 2 for n in np.random.default_rng(1).integers(752, 1502, 64)]
    shards = distributed.shard_utterances(costs, 8)
    loads = [sum(costs[i] for i in s) for s in shards]
    assert sorted(i for s in shards for i in s) == list(range(64))
    assert max(loads) / (sum(loads) / 8) < 1.05
    m, flat = distributed.pack_state({"x": np.arange(5, dtype=np.float32)})
    assert m == [("x", [5], 0)] and flat.shape == (64,)
