"""Multi-GPU driver: poses / complexes are independent units, so each rank samples a contiguous
block with no collective inside the step loop; ONE all_gather of the final ligand coordinates at the end
(NCCL over NVLink on GPUs; gloo in the CPU tests).  The reference samples the N poses of a complex in one process
(inference.py:236-262: N deep copies -> utils/sampling.py:sampling); its only data parallelism is PyG DataParallel
over complexes (utils/utils.py:279), so this module has no reference counterpart to mirror."""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous, size-balanced block [lo, hi) of ``n_items`` for ``rank`` (first ``n_items % world`` ranks get one more)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_positions(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """local [n_local, n_atoms, 3] on every rank -> [n_total, n_atoms, 3] on every rank (shards may differ by one pose)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_max = (n_total + world - 1) // world
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    out = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, r, world)
        out.append(bufs[r][:hi - lo])
    return torch.cat(out, 0)


def sample_sharded(data_list: Sequence, sampler: Callable[[List], List], group=None, device=None,
                   dtype=torch.float32) -> torch.Tensor:
    """Runs ``sampler(local_block)`` (e.g. a partial of diffdock_b200.sampling.sampling) on this rank's block of poses of
    one complex and returns the final coordinates of ALL poses, [len(data_list), n_atoms, 3], on every rank.
    ``device`` / ``dtype`` fix where the gathered tensor lives on EVERY rank (default: the current CUDA device under NCCL,
    the CPU otherwise): a rank whose block is empty (fewer poses than ranks) must still enter the collective with a
    buffer on the same kind of device and of the same dtype as its peers."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if device is None:
        nccl = dist.is_initialized() and dist.get_backend(group) == 'nccl'
        device = torch.device('cuda', torch.cuda.current_device()) if nccl else torch.device('cpu')
    lo, hi = shard_bounds(len(data_list), rank, world)
    block = list(data_list[lo:hi])
    done = sampler(block) if block else []
    n_atoms = data_list[0]['ligand'].pos.shape[0]
    if done:
        local = torch.stack([d['ligand'].pos for d in done]).to(device=device, dtype=dtype)
    else:
        local = torch.zeros((0, n_atoms, 3), device=device, dtype=dtype)
    if world == 1:
        return local
    return all_gather_positions(local, len(data_list), group)


# ---------------------------------------------------------------------------------------------------------------------
# Level-1 partitioning of SURVEY.md section 8(e): whole complexes over GPUs (inference.py:224 walks them one after another)
def assign_balanced(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of items with the given costs (N_r * N_l * poses per complex) to ``world``
    ranks; deterministic (ties broken by index), every rank gets its items in ascending index order."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda q: (load[q], q))
        out[r].append(i)
        load[r] += float(costs[i])
    return [sorted(v) for v in out]


def gather_ragged(local: Sequence[torch.Tensor], owner: Sequence[int], shapes: Sequence[Sequence[int]], group=None,
                  device=None, dtype=torch.float32) -> List[torch.Tensor]:
    """``local[j]`` = result of the j-th item this rank owns (items in ascending index order); ``owner[i]`` / ``shapes[i]`` are
    known on every rank.  ONE all_gather of a packed, padded buffer; returns all items, in index order, on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if device is None:
        nccl = dist.is_initialized() and dist.get_backend(group) == 'nccl'
        device = torch.device('cuda', torch.cuda.current_device()) if nccl else torch.device('cpu')
    numel = [int(torch.Size(sh).numel()) for sh in shapes]
    per_rank = [sum(numel[i] for i in range(len(owner)) if owner[i] == r) for r in range(world)]
    mine = [i for i in range(len(owner)) if owner[i] == rank]
    assert len(mine) == len(local), (len(mine), len(local))
    flat = torch.zeros(max(max(per_rank), 1), device=device, dtype=dtype)
    off = 0
    for i, t in zip(mine, local):
        assert tuple(t.shape) == tuple(shapes[i]), (tuple(t.shape), tuple(shapes[i]))
        flat[off:off + numel[i]] = t.reshape(-1).to(device=device, dtype=dtype)
        off += numel[i]
    if world == 1:
        bufs = [flat]
    else:
        bufs = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(bufs, flat, group=group)
    offs = [0] * world
    out = []
    for i in range(len(owner)):
        r = owner[i]
        out.append(bufs[r][offs[r]:offs[r] + numel[i]].reshape(tuple(shapes[i])))
        offs[r] += numel[i]
    return out


def sample_complexes_sharded(n_complexes: int, costs: Sequence[float], shapes: Sequence[Sequence[int]],
                             sample_one: Callable[[int], torch.Tensor], group=None, device=None) -> List[torch.Tensor]:
    """Config-5-style job: ``n_complexes`` independent complexes, each sampled as ONE batch of all its poses by
    ``sample_one(i) -> [n_poses, n_atoms, 3]`` on the rank that owns it (so the batch a pose is scored in - which the default
    centre convolution depends on, models/cg_model.py:374 - never depends on the number of GPUs), then one collective
    returning every complex's final coordinates on every rank.  With ``rng='philox'`` noise keyed by (complex, pose) the
    gathered result is the same for any world size up to the fp32 summation order of the scatter atomics."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    parts = assign_balanced(costs, world)
    owner = [0] * n_complexes
    for r, items in enumerate(parts):
        for i in items:
            owner[i] = r
    local = [sample_one(i) for i in parts[rank]]
    return gather_ragged(local, owner, shapes, group=group, device=device)
