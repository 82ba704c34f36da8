"""N>1 path on CPU: world_size-2 gloo run of the pose sharding + result gather logic.  The
per-rank "compute" is the CPU oracle here (this is a test), so the gathered arrays must equal
the single-process result on the full batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fcl_b200.poses import random_poses
    from fcl_b200.sharding import all_gather_contacts, all_gather_records, shard_range
    from oracle import pyoracle as O

    g = os.path.join(ROOT, "tests", "golden")
    env, rob = O.Model.from_npz(os.path.join(g, "env.npz")), O.Model.from_npz(os.path.join(g, "rob.npz"))
    s, e = shard_range(n, rank, world)
    P = random_poses(e - s, seed=3, start=s)  # every rank generates only its own block of the global batch
    d = O.distance_batch(env, rob, P, None, True, 2)
    dist_all = all_gather_records(torch.from_numpy(d["min_distance"]), n)
    pts_all = all_gather_records(torch.from_numpy(np.concatenate([d["p1"], d["p2"]], axis=1)), n)
    c = O.collide_batch(env, rob, P, None, 20, True)
    counts, offsets, con = all_gather_contacts(torch.from_numpy(c["counts"]),
                                               torch.from_numpy(c["contacts"].view(np.uint8).copy()), n)
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), dist=dist_all.numpy(), pts=pts_all.numpy(),
                 counts=counts.numpy(), offsets=offsets.numpy(), con=con.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [600, 601])  # even and ragged shards
def test_two_rank_gloo_gather_matches_single_process(tmp_path, oracle, oracle_env_rob, n):
    from fcl_b200.poses import random_poses

    port = _free_port()
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npz"))
    env, rob = oracle_env_rob
    P = random_poses(n, seed=3)
    d = oracle.distance_batch(env, rob, P, None, True, 2)
    assert np.array_equal(got["dist"], d["min_distance"])
    assert np.array_equal(got["pts"], np.concatenate([d["p1"], d["p2"]], axis=1))
    c = oracle.collide_batch(env, rob, P, None, 20, True)
    assert np.array_equal(got["counts"], c["counts"])
    assert np.array_equal(got["offsets"], c["offsets"])
    assert got["con"].tobytes() == c["contacts"].tobytes()


def test_shard_ranges_cover_and_balance():
    from fcl_b200.sharding import shard_range, shard_sizes

    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            s = shard_sizes(n, w)
            assert max(s) - min(s) <= 1 and sum(s) == n
