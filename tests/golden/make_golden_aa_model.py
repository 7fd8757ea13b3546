"""Golden vectors for the all-atom SCORE model: runs the UNMODIFIED reference
models/aa_model.py (AAModel.forward, score mode) from /root/reference in the authoring container, with the third-party
packages supplied by oracle/ref_shims.py, and stores inputs/outputs as tests/golden/ref_aa_model.pt.

    cd /tmp/tables && python /root/repo/tests/golden/make_golden_aa_model.py 2>/dev/null   # cwd holds utils/so3.py's .npy caches
"""
import copy
import os
import sys
from functools import partial

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install()
sys.path.insert(0, '/root/reference')
import utils.diffusion_utils as r_du        # noqa: E402
from utils import torus as r_torus          # noqa: E402
import models.aa_model as r_aa              # noqa: E402

from diffdock_b200.hetero import collate, graph_to_dict   # noqa: E402
from diffdock_b200.synthetic import default_model_args, make_pose_list   # noqa: E402
from tests.parity_helpers import rand_bn_    # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
_tab = np.load(os.path.join(ROOT, 'diffdock_b200', 'tables', 'score_norm_tables.npz'))
r_torus.score_norm_ = _tab['torus_score_norm']      # the stored Monte-Carlo torus table instance (the import re-drew it)


def model_case(lmax, seed, n_poses=3, t=0.4, lm=True, **over):
    a = default_model_args(ns=6, nv=3, sh_lmax=lmax, num_conv_layers=3, distance_embed_dim=8,
                           cross_distance_embed_dim=8, sigma_embed_dim=8, **over)
    kw = dict(sigma_embed_dim=8, sh_lmax=lmax, ns=6, nv=3, num_conv_layers=3, lig_max_radius=a.max_radius,
              rec_max_radius=a.rec_max_radius, cross_max_distance=a.cross_max_distance,
              center_max_distance=a.center_max_distance, distance_embed_dim=8, cross_distance_embed_dim=8,
              dynamic_max_cross=True, lm_embedding_type='precomputed' if lm else None, embed_also_ligand=True,
              num_prot_emb_layers=a.num_prot_emb_layers, differentiate_convolutions=a.differentiate_convolutions)
    torch.manual_seed(seed)
    model = r_aa.AAModel(partial(r_du.t_to_sigma, args=a), torch.device('cpu'),
                         r_du.get_timestep_embedding('sinusoidal', 8, a.embedding_scale), **kw).eval()
    gg = torch.Generator().manual_seed(seed + 1)
    for m in model.modules():
        if m.__class__.__name__ == 'BatchNorm':
            rand_bn_(m, gg)
    poses = make_pose_list(n_poses, n_res=20, n_atoms=9, seed=seed + 2, tr_sigma_max=a.tr_sigma_max * t,
                           lm_dim=16 if lm else 0, all_atoms=True)
    if lm:   # shrink the LM embedding (1280 -> 16) to keep the fixture small: patch the encoder's input Linear
        torch.manual_seed(seed + 5)
        model.rec_node_embedding.additional_features_dim = 16
        model.rec_node_embedding.additional_features_embedder = torch.nn.Linear(16 + 6, 6)
    batch = collate(copy.deepcopy(poses))
    r_du.set_time(batch, t, t, t, t, n_poses, True, 'cpu')
    with torch.no_grad():
        tr, rot, tor, _ = model(batch)
    print('tr', tr[0], 'tor', tor[:3])
    return dict(args=vars(a), kw=kw, lm_dim=16 if lm else 0, state=model.state_dict(),
                poses=[graph_to_dict(p) for p in poses], t=t, tr=tr, rot=rot, tor=tor)


cases = [model_case(2, 50),
         model_case(1, 51, t=0.9, differentiate_convolutions=False),      # faster + multigroup crashes in the reference
         model_case(2, 52, t=0.15, num_prot_emb_layers=1, lm=False)]
torch.save(cases, os.path.join(OUT, 'ref_aa_model.pt'))
print('ref_aa_model.pt', os.path.getsize(os.path.join(OUT, 'ref_aa_model.pt')) // 1024, 'KiB')
