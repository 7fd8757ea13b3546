"""Golden vectors for the ALL-ATOM confidence model: runs the UNMODIFIED reference
models/old_aa_model.py (AAOldModel, confidence_mode) from /root/reference in the authoring container, with the third-party
packages supplied by oracle/ref_shims.py, and stores inputs/outputs as tests/golden/ref_confidence_aa.pt.

    cd /tmp && python /root/repo/tests/golden/make_golden_confidence_aa.py
"""
import copy
import os
import sys
from functools import partial

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install()
sys.path.insert(0, '/root/reference')
import models.old_aa_model as r_aa           # noqa: E402
import utils.diffusion_utils as r_du        # noqa: E402

from diffdock_b200.hetero import collate, graph_to_dict   # noqa: E402
from diffdock_b200.synthetic import default_model_args, make_pose_list   # noqa: E402
from tests.parity_helpers import rand_bn_    # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def case(seed, num_conv_layers, dynamic, n_poses=3, lm=True, smooth=False, affinity=False):
    a = default_model_args()
    kw = dict(sigma_embed_dim=8, sh_lmax=2, ns=6, nv=3, num_conv_layers=num_conv_layers, lig_max_radius=5.0,
              rec_max_radius=30.0, cross_max_distance=25.0, distance_embed_dim=8, cross_distance_embed_dim=8,
              dynamic_max_cross=dynamic, smooth_edges=smooth, lm_embedding_type='esm' if lm else None,
              confidence_mode=True, use_old_atom_encoder=True, num_confidence_outputs=1, affinity_prediction=affinity)
    torch.manual_seed(seed)
    model = r_aa.AAOldModel(partial(r_du.t_to_sigma, args=a), torch.device('cpu'),
                            r_du.get_timestep_embedding('sinusoidal', 8, a.embedding_scale), **kw).eval()
    gg = torch.Generator().manual_seed(seed + 1)
    for m in model.modules():
        if m.__class__.__name__ in ('BatchNorm', 'BatchNorm1d'):
            rand_bn_(m, gg)
    poses = make_pose_list(n_poses, n_res=20, n_atoms=9, seed=seed + 2, tr_sigma_max=1.5, lm_dim=16 if lm else 0,
                           all_atoms=True)
    if lm:   # shrink the LM embedding (1280 -> 16) to keep the fixture small
        torch.manual_seed(seed + 5)
        model.rec_node_embedding.lm_embedding_dim = 16
        model.rec_node_embedding.lm_embedding_layer = torch.nn.Linear(16 + 6, 6)
    batch = collate(copy.deepcopy(poses))
    r_du.set_time(batch, 0, 0, 0, 0, n_poses, True, 'cpu')           # utils/sampling.py:215,222 with all_atoms
    with torch.no_grad():
        conf = model(batch)
    print('confidence', conf)
    return dict(kw=kw, lm_dim=16 if lm else 0, state=model.state_dict(), poses=[graph_to_dict(p) for p in poses],
                confidence=conf)


cases = [case(40, 3, False), case(41, 2, True, smooth=True), case(42, 4, False, lm=False, affinity=True)]
torch.save(cases, os.path.join(OUT, 'ref_confidence_aa.pt'))
print('ref_confidence_aa.pt', os.path.getsize(os.path.join(OUT, 'ref_confidence_aa.pt')) // 1024, 'KiB')
