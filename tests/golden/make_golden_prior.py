"""Golden vectors of the prior sampler: runs the UNMODIFIED ``randomize_position`` (utils/sampling.py:16-58) from
/root/reference on seeded synthetic complexes and stores the poses before / after in tests/golden/ref_prior.pt.

    cd /tmp/tables && python /root/repo/tests/golden/make_golden_prior.py 2>/dev/null    # cwd holds utils/so3.py's .npy caches
                                                                                 # (recomputed there when absent: minutes)

Seeds: np.random.seed / random.seed / torch.manual_seed are set to the case's ``seed`` immediately before the call; the test
does the same before calling diffdock_b200.sampling.randomize_position."""
import copy
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install()
import utils.sampling as r_sampling            # noqa: E402
from diffdock_b200.hetero import graph_to_dict   # noqa: E402
from diffdock_b200.synthetic import make_complex  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'ref_prior.pt')

CASES = [dict(seed=0, n=4, kw=dict(no_torsion=False, no_random=False, tr_sigma_max=19.0)),
         dict(seed=1, n=3, kw=dict(no_torsion=True, no_random=False, tr_sigma_max=19.0)),
         dict(seed=2, n=3, kw=dict(no_torsion=False, no_random=True, tr_sigma_max=19.0)),
         dict(seed=3, n=3, kw=dict(no_torsion=False, no_random=False, tr_sigma_max=19.0, initial_noise_std_proportion=1.46)),
         dict(seed=4, n=3, kw=dict(no_torsion=False, no_random=False, tr_sigma_max=19.0, choose_residue=True)),
         dict(seed=5, n=2, kw=dict(no_torsion=False, no_random=False, tr_sigma_max=19.0, pocket_knowledge=True, pocket_cutoff=9))]

if __name__ == '__main__':
    fx = []
    for c in CASES:
        g = make_complex(n_res=80, n_atoms=14 + c['seed'], seed=10 + c['seed'])
        if c['kw'].get('pocket_knowledge'):
            g['ligand'].orig_pos = [(g['ligand'].pos + torch.tensor([[4.0, -3.0, 2.0]])).numpy().astype(np.float64)]
            g.original_center = torch.tensor([[1.0, 2.0, 3.0]])
        poses = [copy.deepcopy(g) for _ in range(c['n'])]
        np.random.seed(c['seed']); random.seed(c['seed']); torch.manual_seed(c['seed'])
        r_sampling.randomize_position(poses, **c['kw'])
        g['receptor'].x = g['receptor'].x[:, :1].clone()          # the prior does not read the features: keep the fixture small
        if 'side_chain_vecs' in g['receptor']:
            del g['receptor'].__dict__['side_chain_vecs']
        fx.append({'seed': c['seed'], 'n': c['n'], 'kw': c['kw'], 'complex': graph_to_dict(g),
                   'pos_out': [p['ligand'].pos.clone() for p in poses]})
        print(c['kw'], float(poses[0]['ligand'].pos.abs().max()))
    torch.save(fx, OUT)
    print(OUT, os.path.getsize(OUT) // 1024, 'KiB')
