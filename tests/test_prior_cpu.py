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
