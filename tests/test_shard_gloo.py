"""CPU tests of the multi-GPU host logic (sprl_b200/shard.py) with the gloo backend,
world_size 2 and 3: the game -> rank partition, the weight broadcast that opens a
generation, the sample-count all-gather, and the merge back into game order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sprl_b200 import shard
from sprl_b200.network import make_network


def test_shard_plan_partitions_games():
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 8, 1000):
            seen = []
            for r in range(world):
                first, stride, count = shard.shard_plan(n, r, world)
                assert first == r and stride == world
                ids = [first + k * stride for k in range(count)]
                assert all(shard.owner_of(g, world) == r for g in ids)
                seen += ids
            assert sorted(seen) == list(range(n))
    with pytest.raises(ValueError):
        shard.shard_plan(10, 2, 2)
    with pytest.raises(ValueError):
        shard.shard_plan(-1, 0, 1)


def test_merge_shards_restores_game_order():
    world, n = 3, 11
    per_rank = [[g for g in range(n) if g % world == r] for r in range(world)]
    assert shard.merge_shards(per_rank, world) == list(range(n))
    assert shard.row_offsets([5, 0, 7]) == [0, 5, 5]


def test_flat_state_round_trip():
    a, b = make_network("othello", 1), make_network("othello", 2)
    flat, _ = shard.flatten_module_state(a)
    # 162,949 parameters + BatchNorm running stats and counters
    assert sum(p.numel() for p in a.parameters()) == 162949 and flat.numel() > 162949
    _, tensors_b = shard.flatten_module_state(b)
    shard.load_flat_state(tensors_b, flat)
    for x, y in zip(a.state_dict().values(), b.state_dict().values()):
        assert torch.equal(x, y)
    with pytest.raises(ValueError):
        shard.load_flat_state(tensors_b, torch.cat([flat, flat[:1]]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        # every rank starts from different weights; after the broadcast all equal rank 0's
        net = make_network("othello", seed=100 + rank)
        nbytes = shard.broadcast_module_state(net, src=0)
        want = make_network("othello", seed=100)
        same = all(torch.equal(x, y) for x, y in zip(net.state_dict().values(), want.state_dict().values()))
        # pretend generation: this rank "plays" its shard; samples per game depend on the game id only
        first, stride, count = shard.shard_plan(37, rank, world)
        games = [first + k * stride for k in range(count)]
        n_samples = sum(8 * (50 + g % 11) for g in games)
        counts = shard.gather_sample_counts(n_samples)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), same=same, nbytes=nbytes, games=games, counts=counts,
                 offsets=shard.row_offsets(counts), mine=n_samples)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_generation_exchange_over_gloo(world, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    all_games = []
    for r, z in enumerate(res):
        assert bool(z["same"]), "weights differ from rank 0's after the broadcast"
        assert int(z["nbytes"]) == int(res[0]["nbytes"]) > 4 * 162949
        assert list(z["counts"]) == [int(x["mine"]) for x in res]           # every rank sees every count
        assert int(z["offsets"][r]) == sum(int(x["mine"]) for x in res[:r])
        all_games.append(list(z["games"]))
    assert shard.merge_shards(all_games, world) == list(range(37))
