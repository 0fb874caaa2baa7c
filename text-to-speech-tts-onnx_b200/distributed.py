"""Multi-GPU plumbing. The hot path shards by utterance with NO data-path collective: the only
exchange is the one-time weight broadcast from rank 0 over NCCL/NVLink, and an optional gather of results.
One process per GPU (torchrun); works with a single process too (world size 1, no process group needed)."""
import json

import numpy as np
import torch
import torch.distributed as dist


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def pack_state(state: dict):
    """dict name -> fp32 ndarray  ==>  (manifest [(name, shape, offset)], flat fp32 ndarray); offsets 64-float aligned."""
    manifest, off = [], 0
    for k, v in state.items():
        v = np.asarray(v, dtype=np.float32)
        manifest.append((k, list(v.shape), off))
        off += (v.size + 63) // 64 * 64
    flat = np.zeros(off, dtype=np.float32)
    for (k, shape, o) in manifest:
        flat[o:o + int(np.prod(shape, dtype=np.int64))] = np.asarray(state[k], dtype=np.float32).reshape(-1)
    return manifest, flat


def broadcast_packed(state, src: int = 0, device=None):
    """Rank `src` provides `state`; every rank returns (manifest, flat tensor on `device`). Backend-agnostic
    (NCCL on GPUs, gloo in the CPU tests)."""
    world, rank = _world(), _rank()
    if world == 1:
        manifest, flat = pack_state(state)
        t = torch.from_numpy(flat)
        return manifest, (t.to(device) if device is not None else t)
    if rank == src:
        manifest, flat = pack_state(state)
        meta = [json.dumps(manifest)]
    else:
        manifest, flat, meta = None, None, [None]
    dist.broadcast_object_list(meta, src=src)
    manifest = json.loads(meta[0])
    total = max((o + int(np.prod(s, dtype=np.int64)) for _, s, o in manifest), default=0)
    total = (total + 63) // 64 * 64
    if rank == src:
        t = torch.from_numpy(flat)
        t = t.to(device) if device is not None else t
    else:
        t = torch.empty(total, dtype=torch.float32, device=device if device is not None else "cpu")
    dist.broadcast(t, src=src)
    return [tuple(m) for m in manifest], t


def load_state_broadcast(engine, prefix: str, state, src: int = 0):
    """Broadcast `state` (given on rank `src`) and load it into this rank's engine from device memory."""
    device = torch.device("cuda", engine.device)
    manifest, flat = broadcast_packed(state, src=src, device=device)
    torch.cuda.synchronize(device)
    base = flat.data_ptr()
    for name, shape, off in manifest:
        engine.load_tensor_device(f"{prefix}.{name}", base + 4 * off, shape)
    return len(manifest)


def shard_utterances(costs, world: int):
    """Deal utterances to ranks so that the per-rank cost (e.g. N^2 + c*N per utterance) is balanced: sort by cost,
    descending, and give each to the currently lightest rank (LPT). Returns a list of index lists, one per rank;
    deterministic, identical on every rank."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: (loads[j], j))
        out[r].append(i)
        loads[r] += costs[i]
    return [sorted(s) for s in out]


def gather_objects(obj, dst: int = 0):
    """Host-side gather of small per-rank results (PCM lengths, timings)."""
    world, rank = _world(), _rank()
    if world == 1:
        return [obj]
    out = [None] * world if rank == dst else None
    dist.gather_object(obj, out, dst=dst)
    return out
