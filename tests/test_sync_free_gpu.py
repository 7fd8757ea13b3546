"""GPU: the sync-free forward (device-built neighbour lists, device-side edge counts, reverse cross graph as a permutation,
fused edge embedding; diffdock_b200/cg_model.py:_forward_sync_free) against the host-sized forward of the same model and
against the CPU oracle; the CUDA-graph sampler against the eager sampler; counter-based (Philox) noise."""
import copy
import ctypes as C
from functools import partial

import numpy as np
import pytest
import torch

from tests.parity_helpers import make_model_pair, rel_err

pytestmark = pytest.mark.gpu


def _args(**over):
    from diffdock_b200.synthetic import default_model_args
    kw = dict(ns=16, nv=4, sh_lmax=2, num_conv_layers=3, distance_embed_dim=16, cross_distance_embed_dim=16, sigma_embed_dim=16)
    kw.update(over)
    return default_model_args(**kw)


def _batch(n_poses, seed, t, args, n_res=60, n_atoms=12, device='cuda:0', far=()):
    from diffdock_b200.synthetic import make_pose_list
    from diffdock_b200.hetero import collate
    from diffdock_b200.diffusion_utils import set_time
    poses = make_pose_list(n_poses, n_res=n_res, n_atoms=n_atoms, seed=seed, tr_sigma_max=args.tr_sigma_max * t)
    for i in far:
        poses[i]['ligand'].pos = poses[i]['ligand'].pos + 80.0
    g = collate(poses).to(device)
    set_time(g, None, t, t, t, n_poses, False, device)
    return poses, g


@pytest.mark.parametrize("over,t,far", [({}, 0.5, ()), ({}, 1.0, ()), ({'smooth_edges': True}, 0.3, ()),
                                        ({'dynamic_max_cross': False, 'cross_max_distance': 25.0}, 0.5, ()),
                                        ({'num_prot_emb_layers': 1}, 0.4, ()), ({}, 0.05, (1,)),
                                        ({'differentiate_convolutions': False}, 0.6, ()), ({'no_torsion': True}, 0.5, ())])
def test_sync_free_forward_matches_host_sized_and_oracle(built_lib, over, t, far):
    args = _args(**over)
    o, p = make_model_pair(args, seed=3)
    assert p.sync_free_capable()
    poses, g = _batch(3, 11, t, args, far=far)
    got = p(g)
    p2 = copy.deepcopy(p)
    p2._sync_free = False                                     # the exactly-sized path with host-side counts
    _, g2 = _batch(3, 11, t, args, far=far)
    ref = p2(g2)
    torch.cuda.synchronize()
    # a pose displaced by 80 A makes the fp32 centroid / distance arithmetic the dominant error (see tests/test_model_gpu.py)
    tol = 3e-4 if far else 2e-5
    for a, b in zip(got[:3], ref[:3]):
        assert a.shape == b.shape
        if a.numel():
            assert rel_err(a, b) < tol
    from oracle.diffusion import set_time as o_set_time
    from diffdock_b200.hetero import collate
    g_cpu = collate(poses)
    o_set_time(g_cpu, t, t, t, 3, 'cpu')
    with torch.no_grad():
        oref = o(g_cpu)
    for a, b in zip(got[:3], oref[:3]):
        if b.numel():
            assert rel_err(a, b) < (3e-4 if far else 1e-4)


def test_full_width_sync_free_vs_oracle(built_lib):
    """DiffDock-L widths (ns=48, nv=10, 64-dim embeddings: the edge-embedding kernel's main instantiation)."""
    args = _args(ns=48, nv=10, num_conv_layers=4, distance_embed_dim=64, cross_distance_embed_dim=64, sigma_embed_dim=64)
    o, p = make_model_pair(args, seed=5)
    assert p.sync_free_capable()
    poses, g = _batch(2, 21, 0.5, args, n_res=90, n_atoms=15)
    got = p(g)
    from oracle.diffusion import set_time as o_set_time
    from diffdock_b200.hetero import collate
    g_cpu = collate(poses)
    o_set_time(g_cpu, 0.5, 0.5, 0.5, 2, 'cpu')
    with torch.no_grad():
        oref = o(g_cpu)
    for a, b in zip(got[:3], oref[:3]):
        assert rel_err(a, b) < 1e-4


def test_shared_receptor_collate_matches_general_collate(built_lib):
    from diffdock_b200.hetero import collate, collate_shared_receptor
    from diffdock_b200.diffusion_utils import set_time
    args = _args()
    o, p = make_model_pair(args, seed=7)
    poses, g = _batch(4, 31, 0.5, args)
    gs = collate_shared_receptor([q.clone() for q in poses], 'cuda:0')
    assert getattr(gs['receptor'], '_unique', None) == (60, poses[0]['receptor', 'receptor'].num_edges, 4)
    set_time(gs, None, 0.5, 0.5, 0.5, 4, False, 'cuda:0')
    a, b = p(g), p(gs)
    for x, y in zip(a[:3], b[:3]):
        assert rel_err(x, y) < 5e-5            # library GEMMs of different heights + scatter order: not bit-equal
    assert torch.equal(gs['receptor'].x, g['receptor'].x) and torch.equal(gs['receptor', 'receptor'].edge_index,
                                                                          g['receptor', 'receptor'].edge_index)
    # layer-0 receptor<-receptor messages computed for ONE copy and added to all (needs the sampler's promise of a uniform t)
    gs._uniform_t = True
    calls = []
    orig = p.conv_layers[0].accumulate_group
    p.conv_layers[0].accumulate_group = lambda *a_, **k_: (calls.append(a_[1][0].shape[0]), orig(*a_, **k_))[1]
    c2 = p(gs)
    assert calls == [poses[0]['receptor', 'receptor'].num_edges]            # one copy's edges, once
    for x, y in zip(a[:3], c2[:3]):
        assert rel_err(y, x) < 5e-5


def _sample(p, args, poses, **kw):
    from diffdock_b200.diffusion_utils import get_t_schedule, t_to_sigma
    from diffdock_b200.sampling import sampling
    sched = get_t_schedule('expbeta', 6)
    out, _ = sampling([q.clone() for q in poses], p, 6, sched, sched, sched, 'cuda:0', partial(t_to_sigma, args=args), args,
                      batch_size=kw.pop('batch_size', len(poses)), no_final_step_noise=True, **kw)
    torch.cuda.synchronize()
    return torch.stack([d['ligand'].pos for d in out]).cpu()


def test_cuda_graph_sampler_matches_eager_sampler(built_lib):
    from diffdock_b200.synthetic import make_pose_list
    args = _args()
    _, p = make_model_pair(args, seed=9)
    poses = make_pose_list(4, n_res=60, n_atoms=12, seed=41, tr_sigma_max=args.tr_sigma_max)
    eager = _sample(p, args, poses, rng='philox', seed=123, cuda_graph=False)
    graphed = _sample(p, args, poses, rng='philox', seed=123, cuda_graph=True)
    assert torch.isfinite(graphed).all()
    assert float((eager - graphed).abs().max()) < 2e-3      # 6 chained steps; scatter order differs run to run
    ode_e = _sample(p, args, poses, ode=True, cuda_graph=False)
    ode_g = _sample(p, args, poses, ode=True, cuda_graph=True)
    assert float((ode_e - ode_g).abs().max()) < 2e-3
    other = _sample(p, args, poses, rng='philox', seed=124, cuda_graph=True)
    assert float((other - graphed).abs().max()) > 1e-2     # another seed, another trajectory


def test_philox_noise_independent_of_batch_split(built_lib):
    """Per-pose counter-based streams: with a batch-composition-independent model (fixed_center_conv=True; the default
    indexes the ligand features with graph ids, models/cg_model.py:374, so its scores depend on the batch) a pose gets the
    same trajectory whether it is sampled in a batch of 6 or in batches of 2 - what sharding over GPUs relies on."""
    from diffdock_b200.synthetic import make_pose_list
    args = _args(fixed_center_conv=True)
    _, p = make_model_pair(args, seed=13)
    poses = make_pose_list(6, n_res=60, n_atoms=12, seed=51, tr_sigma_max=args.tr_sigma_max)
    keys = (7 << 32) + torch.arange(6)
    whole = _sample(p, args, poses, rng='philox', seed=5, pose_keys=keys)
    parts = _sample(p, args, poses, rng='philox', seed=5, pose_keys=keys, batch_size=2)
    assert float((whole - parts).abs().max()) < 2e-3
    tail = _sample(p, args, poses[4:], rng='philox', seed=5, pose_keys=keys[4:])
    assert float((whole[4:] - tail).abs().max()) < 2e-3


def test_philox_known_answer_and_moments(built_lib):
    """Philox4x32-10 against the Random123 known-answer vectors, and the first moments of the normals."""
    lib = built_lib
    raw = torch.zeros(4, dtype=torch.int32, device='cuda')
    z = torch.zeros(4, device='cuda')
    vp = lambda t: C.c_void_p(t.data_ptr())
    # counter = (0, 0, 0, 0), key = (0, 0)
    assert lib.ddb200_philox_probe(C.c_uint64(0), 0, 0, 0, 1, vp(z), vp(raw), None) == 0
    torch.cuda.synchronize()
    assert [int(v) & 0xffffffff for v in raw.tolist()] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    # counter = ff..f, key = ff..f
    assert lib.ddb200_philox_probe(C.c_uint64(0xffffffffffffffff), -1, 0xffffffff, 0xffffffff, 1, vp(z), vp(raw), None) == 0
    torch.cuda.synchronize()
    assert [int(v) & 0xffffffff for v in raw.tolist()] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    n = 50000
    zz = torch.zeros(4 * n, device='cuda')
    assert lib.ddb200_philox_probe(C.c_uint64(99), 12345, 3, 0, n, vp(zz), None, None) == 0
    torch.cuda.synchronize()
    assert abs(float(zz.mean())) < 0.01 and abs(float(zz.var()) - 1.0) < 0.02 and float(zz.abs().max()) < 6.5
    assert abs(float((zz[0::4] * zz[1::4]).mean())) < 0.01


def test_complex_sharding_result_independent_of_world_size(built_lib):
    """Level-1 partitioning (whole complexes per rank) emulated on one GPU: the jobs of a 1-rank and of a 3-rank
    assignment are executed one after another and must give the same coordinates for every complex - each complex is always
    one batch (so even the batch-dependent default centre convolution cannot tell the difference) and its noise is keyed by
    (complex, pose, step)."""
    from diffdock_b200.distributed import assign_balanced
    from diffdock_b200.synthetic import make_pose_list
    args = _args()
    _, p = make_model_pair(args, seed=23)
    sizes = [(40, 9), (70, 14), (55, 11), (30, 8)]
    costs = [r * a for r, a in sizes]

    def sample_one(i):
        poses = make_pose_list(3, n_res=sizes[i][0], n_atoms=sizes[i][1], seed=70 + i, tr_sigma_max=args.tr_sigma_max,
                               share_receptor=True)
        return _sample(p, args, poses, rng='philox', seed=9, pose_keys=(i << 32) + torch.arange(3))

    results = {}
    for world in (1, 3):
        out = {}
        for rank_items in assign_balanced(costs, world):
            for i in rank_items:                      # what rank r would run, in its order
                out[i] = sample_one(i)
        results[world] = out
    for i in range(len(sizes)):
        assert float((results[1][i] - results[3][i]).abs().max()) < 2e-3
