"""Sharding of one self-play generation across the GPUs of a box (one process per GPU).

The reference scales by launching NUM_WORKER_TASKS independent single-threaded worker
processes that never talk to each other (cpp/src/OTHWorker.cpp:12-13,44-49) and meets the
controller only through files.  Games are independent units here too: game `g` of a
generation is owned by rank `g % world`, every rank plays its games on its own GPU with no
data-path collective, and the only exchanges are the two tiny ones SURVEY.md 8(e) names:

  * broadcast_module_state -- the new network's parameters + buffers, rank 0 -> all, once per
    generation (replaces every worker polling the shared filesystem for the traced .pt,
    cpp/src/selfplay/GridWorker.hpp:35-55);
  * gather_sample_counts   -- one int64 per rank, so that every rank knows the row range of its
    samples inside the generation's arrays (replaces the controller counting rows file by
    file, scripts/othello_controller.py:84-100).

The collectives run over torch.distributed: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
import numpy as np


def shard_plan(num_games, rank, world):
    """Games of a generation owned by `rank`: stream ids rank, rank + world, ...
    Returns (first_game, stride, count) for Engine.begin_iteration / set_game_stride."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    if num_games < 0:
        raise ValueError("num_games must be non-negative")
    count = (num_games - rank + world - 1) // world if num_games > rank else 0
    return rank, world, count


def owner_of(game_id, world):
    return game_id % world


def flatten_module_state(module):
    """Parameters then buffers of a torch module, in module order, as one flat fp32 vector."""
    import torch
    tensors = [p for p in module.parameters()] + [b for b in module.buffers()]
    if not tensors:
        return torch.zeros(0), tensors
    flat = torch.cat([t.detach().reshape(-1).to(torch.float32) for t in tensors])
    return flat, tensors


def load_flat_state(tensors, flat):
    import torch
    at = 0
    with torch.no_grad():
        for t in tensors:
            n = t.numel()
            t.copy_(flat[at:at + n].reshape(t.shape).to(t.dtype))
            at += n
    if at != flat.numel():
        raise ValueError(f"flat state has {flat.numel()} elements, module needs {at}")


def broadcast_module_state(module, src=0, group=None):
    """Makes every rank's `module` equal to rank `src`'s: one broadcast of the flat state.
    Returns the number of bytes broadcast.  Integer buffers (BatchNorm's num_batches_tracked)
    travel as fp32, exact below 2^24."""
    import torch.distributed as dist
    flat, tensors = flatten_module_state(module)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src, group=group)
        load_flat_state(tensors, flat)
    return flat.numel() * 4


def gather_sample_counts(n_samples, device=None, group=None):
    """All-gather of one int64 per rank; returns the list of per-rank sample counts."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [int(n_samples)]
    mine = torch.tensor([int(n_samples)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, mine, group=group)
    return [int(t.item()) for t in out]


def row_offsets(counts):
    """First row of each rank's block when the generation's samples are laid out rank-major."""
    return [int(x) for x in np.concatenate([[0], np.cumsum(counts)[:-1]])]


def merge_shards(per_rank_games, world):
    """Interleaves per-rank lists of per-game items back into game-id order
    (rank r's k-th game is game r + k * world).  Used to compare a sharded run with an unsharded one."""
    total = sum(len(x) for x in per_rank_games)
    out = [None] * total
    for r, games in enumerate(per_rank_games):
        for k, item in enumerate(games):
            out[r + k * world] = item
    return out
