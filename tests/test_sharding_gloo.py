"""world_size-2 gloo test of the N>1 path: block partition of pairs, per-rank tracking, final gather.
On CPU the per-rank compute is the oracle (the checker); on the GPU box bench.py uses the same
partition/gather code with the CUDA path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lego_slam_b200 import sharding


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sharding.shard_sizes(n, world)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_pairs, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lego_slam_b200 import synth
    from oracle import binding as ob
    lo, hi = sharding.shard_range(n_pairs, rank, world)
    kps, sus = [], []
    for pair in range(lo, hi):
        L, R, kp1, kp2, _ = synth.stereo_case(94, 310, 20, seed=100 + pair, min_dist=8)
        o, s, _ = ob.track(L, R, kp1, kp2, ob.make_params(levels=3))
        kps.append(o)
        sus.append(s)
    kp_local = torch.from_numpy(np.stack(kps)) if kps else torch.zeros((0, 20, 2))
    su_local = torch.from_numpy(np.stack(sus)) if sus else torch.zeros((0, 20), dtype=torch.uint8)
    kp_full, su_full = sharding.gather_results(kp_local, su_local, n_pairs)
    if rank == 0:
        np.save(os.path.join(out_dir, "kp.npy"), kp_full.numpy())
        np.save(os.path.join(out_dir, "su.npy"), su_full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_tracking_equals_single_process(tmp_path):
    from lego_slam_b200 import synth
    from oracle import binding as ob
    n_pairs, world = 5, 2  # uneven on purpose: shards of 3 and 2
    mp.spawn(_worker, args=(world, _free_port(), n_pairs, str(tmp_path)), nprocs=world, join=True)
    kp = np.load(tmp_path / "kp.npy")
    su = np.load(tmp_path / "su.npy")
    assert kp.shape == (n_pairs, 20, 2) and su.shape == (n_pairs, 20)
    for pair in range(n_pairs):
        L, R, kp1, kp2, _ = synth.stereo_case(94, 310, 20, seed=100 + pair, min_dist=8)
        o, s, _ = ob.track(L, R, kp1, kp2, ob.make_params(levels=3))
        assert np.array_equal(o.view(np.uint32), kp[pair].view(np.uint32))
        assert np.array_equal(s, su[pair])
