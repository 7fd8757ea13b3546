"""CPU: the prior sampler ``diffdock_b200.sampling.randomize_position`` against the poses the UNMODIFIED reference function
(utils/sampling.py:16-58) produced from the same seeds (tests/golden/ref_prior.pt, tests/golden/make_golden_prior.py):
same random streams in the same order, so the comparison is bit for bit."""
import copy
import random

import numpy as np
import pytest
import torch

from tests.parity_helpers import load_golden


@pytest.mark.parametrize('i', range(6))
def test_randomize_position_matches_reference(i):
    from diffdock_b200.hetero import graph_from_dict
    from diffdock_b200.sampling import randomize_position
    c = load_golden('ref_prior.pt')[i]
    g = graph_from_dict(c['complex'])
    poses = [copy.deepcopy(g) for _ in range(c['n'])]
    np.random.seed(c['seed']); random.seed(c['seed']); torch.manual_seed(c['seed'])
    randomize_position(poses, **c['kw'])
    for p, want in zip(poses, c['pos_out']):
        assert p['ligand'].pos.dtype == torch.float32 and torch.equal(p['ligand'].pos, want), \
            float((p['ligand'].pos - want).abs().max())


def test_randomize_position_on_shared_receptor_copies():
    """pose_copies of one complex (shared receptor store): every pose still gets its own coordinates, and the result equals
    the deep-copy path."""
    from diffdock_b200.hetero import graph_from_dict
    from diffdock_b200.inputs import pose_copies
    from diffdock_b200.sampling import randomize_position
    c = load_golden('ref_prior.pt')[3]
    g = graph_from_dict(c['complex'])
    poses = pose_copies(g, c['n'])
    np.random.seed(c['seed']); random.seed(c['seed']); torch.manual_seed(c['seed'])
    randomize_position(poses, **c['kw'])
    for p, want in zip(poses, c['pos_out']):
        assert torch.equal(p['ligand'].pos, want)
    assert not torch.equal(poses[0]['ligand'].pos, poses[1]['ligand'].pos)


def test_rank_poses_matches_the_reference_epilogue():
    """inference.py:274-283 restated on arrays: + original_center, first column of a multi-threshold head, argsort descending."""
    from diffdock_b200.hetero import HeteroGraph
    from diffdock_b200.sampling import rank_poses
    g = torch.Generator().manual_seed(0)
    poses = []
    for _ in range(5):
        h = HeteroGraph()
        h['ligand'].pos = torch.randn(7, 3, generator=g)
        poses.append(h)
    centre = torch.tensor([[10.0, -2.0, 3.5]])
    conf = torch.randn(5, 2, generator=g)
    got_pos, got_conf, order = rank_poses(poses, conf, centre, rmsd_classification_cutoff=[2.0, 5.0])
    want_pos = np.asarray([p['ligand'].pos.cpu().numpy() + centre.cpu().numpy() for p in poses])
    c = conf[:, 0].cpu().numpy()
    ro = np.argsort(c)[::-1]
    assert np.array_equal(order, ro) and np.array_equal(got_conf, c[ro]) and np.array_equal(got_pos, want_pos[ro])
    p2, c2, o2 = rank_poses(poses, None, centre)
    assert c2 is None and o2 is None and np.array_equal(p2, want_pos)
    p3, c3, _ = rank_poses(poses, conf[:, 1], centre, rmsd_classification_cutoff=2.0)
    assert np.array_equal(c3, np.sort(conf[:, 1].numpy())[::-1])
