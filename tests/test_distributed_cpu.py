"""CPU, world_size=2, gloo: the pose-sharding + final-gather logic of the multi-GPU driver (the data path itself has no
collective).  Rendezvous on 127.0.0.1."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffdock_b200.distributed import assign_balanced, shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 5, 8, 40, 41):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_poses, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from diffdock_b200.distributed import sample_sharded
    from diffdock_b200.synthetic import make_pose_list
    poses = make_pose_list(n_poses, n_res=12, n_atoms=5, seed=3)

    def fake_sampler(block):     # stands in for sampling(): marks every pose with a value that depends on the pose only
        for d in block:
            d['ligand'].pos = d['ligand'].pos * 2 + 1
        return block

    out = sample_sharded(poses, fake_sampler)
    expect = torch.stack([p['ligand'].pos * 2 + 1 for p in make_pose_list(n_poses, n_res=12, n_atoms=5, seed=3)])
    ret[rank] = bool(torch.equal(out, expect))
    dist.destroy_process_group()


def test_two_rank_sharded_sampling_gathers_all_poses():
    for n_poses, port in ((5, 29611), (4, 29612), (1, 29613)):     # uneven, even, and an empty block on rank 1
        mgr = mp.Manager()
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, n_poses, ret), nprocs=2, join=True)
        assert ret[0] and ret[1]


def test_assign_balanced_is_deterministic_and_balanced():
    import random
    rnd = random.Random(0)
    costs = [rnd.randint(200, 600) * rnd.randint(15, 50) for _ in range(64)]
    for w in (1, 2, 4, 8):
        parts = assign_balanced(costs, w)
        assert sorted(i for p in parts for i in p) == list(range(64)) and parts == assign_balanced(costs, w)
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) <= 1.05 * sum(costs) / w + max(costs) * 0.25
    assert assign_balanced([1.0, 1.0], 4) == [[0], [1], [], []]


def _worker_complexes(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from diffdock_b200.distributed import sample_complexes_sharded
    sizes = [(3, 5), (4, 7), (2, 9), (5, 4), (3, 6)]            # (poses, atoms) per complex
    shapes = [(p, a, 3) for p, a in sizes]
    calls = []

    def sample_one(i):                                           # value depends on the complex only
        calls.append(i)
        p, a = sizes[i]
        return torch.arange(p * a * 3, dtype=torch.float32).reshape(p, a, 3) + 1000 * i

    out = sample_complexes_sharded(len(sizes), [p * a for p, a in sizes], shapes, sample_one)
    ok = all(torch.equal(out[i], torch.arange(p * a * 3, dtype=torch.float32).reshape(p, a, 3) + 1000 * i)
             for i, (p, a) in enumerate(sizes))
    ret[rank] = (ok, sorted(calls))
    dist.destroy_process_group()


def test_two_rank_complex_sharding_gathers_every_complex():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_complexes, args=(2, 29621, ret), nprocs=2, join=True)
    assert ret[0][0] and ret[1][0]
    assert sorted(ret[0][1] + ret[1][1]) == [0, 1, 2, 3, 4] and ret[0][1] and ret[1][1]      # disjoint, both ranks worked
