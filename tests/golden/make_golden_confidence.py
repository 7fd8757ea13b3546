"""Golden vectors for the confidence model: runs the UNMODIFIED reference
models/old_cg_model.py (CGOldModel, confidence_mode) from /root/reference in the authoring container, with the
third-party packages supplied by oracle/ref_shims.py, and stores inputs/outputs as tests/golden/ref_confidence.pt.

    cd /tmp/tables && python /root/repo/tests/golden/make_golden_confidence.py
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
import models.old_cg_model as r_old          # noqa: E402
import utils.diffusion_utils as r_du        # noqa: E402

from diffdock_b200.hetero import collate, graph_to_dict   # noqa: E402
from diffdock_b200.synthetic import default_model_args, make_pose_list   # noqa: E402
from tests.parity_helpers import rand_bn_    # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def case(seed, num_conv_layers, dynamic, n_poses=3, lm=True, smooth=False):
    a = default_model_args()
    kw = dict(sigma_embed_dim=8, sh_lmax=2, ns=6, nv=3, num_conv_layers=num_conv_layers, lig_max_radius=5.0,
              rec_max_radius=30.0, cross_max_distance=25.0, distance_embed_dim=8, cross_distance_embed_dim=8,
              dynamic_max_cross=dynamic, smooth_edges=smooth, lm_embedding_type='esm' if lm else None,
              confidence_mode=True, use_old_atom_encoder=True, num_confidence_outputs=1)
    torch.manual_seed(seed)
    model = r_old.CGOldModel(partial(r_du.t_to_sigma, args=a), torch.device('cpu'),
                             r_du.get_timestep_embedding('sinusoidal', 8, a.embedding_scale), **kw).eval()
    gg = torch.Generator().manual_seed(seed + 1)
    for m in model.modules():
        if m.__class__.__name__ in ('BatchNorm', 'BatchNorm1d'):
            rand_bn_(m, gg)
    poses = make_pose_list(n_poses, n_res=24, n_atoms=9, seed=seed + 2, tr_sigma_max=1.5, lm_dim=16 if lm else 0)
    if lm:   # shrink the LM embedding (1280 -> 16) to keep the fixture small
        torch.manual_seed(seed + 5)
        model.rec_node_embedding.lm_embedding_dim = 16
        model.rec_node_embedding.lm_embedding_layer = torch.nn.Linear(16 + 6, 6)
    batch = collate(copy.deepcopy(poses))
    r_du.set_time(batch, 0, 0, 0, 0, n_poses, False, 'cpu')          # utils/sampling.py:215,222
    with torch.no_grad():
        conf = model(batch)
    print('confidence', conf)
    return dict(kw=kw, lm_dim=16 if lm else 0, state=model.state_dict(), poses=[graph_to_dict(p) for p in poses],
                confidence=conf)


cases = [case(30, 3, False), case(31, 2, True, smooth=True), case(32, 4, False, lm=False)]
torch.save(cases, os.path.join(OUT, 'ref_confidence.pt'))
print('ref_confidence.pt', os.path.getsize(os.path.join(OUT, 'ref_confidence.pt')) // 1024, 'KiB')


# ------------------------------------------------------------------------------------------------ sampling + confidence
# utils/sampling.py run unmodified: 3 reverse-diffusion steps of the score model of ref_cg_model.pt[0], then the confidence
# model of case 0 above on the final poses (confidence_data_list path, with and without confidence crop_beyond).
import numpy as np                           # noqa: E402
import utils.sampling as r_sampling         # noqa: E402
import models.cg_model as r_cg              # noqa: E402
from argparse import Namespace              # noqa: E402
from diffdock_b200.hetero import graph_from_dict   # noqa: E402

score_case = torch.load(os.path.join(OUT, 'ref_cg_model.pt'), weights_only=False)[0]
sa = Namespace(**score_case['args'])
score = r_cg.CGModel(partial(r_du.t_to_sigma, args=sa), torch.device('cpu'),
                     r_du.get_timestep_embedding('sinusoidal', 8, sa.embedding_scale), **score_case['kw']).eval()
score.rec_node_embedding.additional_features_dim = 16
score.rec_node_embedding.additional_features_embedder = torch.nn.Linear(16 + 6, 6)
score.load_state_dict(score_case['state'], strict=True)
poses = [graph_from_dict(d) for d in score_case['poses']]

ca = default_model_args()
ckw = cases[0]['kw']
conf = r_old.CGOldModel(partial(r_du.t_to_sigma, args=ca), torch.device('cpu'),
                        r_du.get_timestep_embedding('sinusoidal', 8, ca.embedding_scale), **ckw).eval()
conf.rec_node_embedding.lm_embedding_dim = 16
conf.rec_node_embedding.lm_embedding_layer = torch.nn.Linear(16 + 6, 6)
conf.load_state_dict(cases[0]['state'], strict=True)

sched = np.array([0.30, 0.18, 0.07])
runs = []
for crop in (None, 9.0):
    cargs = Namespace(all_atoms=False, crop_beyond=crop)
    torch.manual_seed(77)
    margs = copy.deepcopy(sa)
    out_list, c = r_sampling.sampling(data_list=copy.deepcopy(poses), model=score, inference_steps=3, tr_schedule=sched,
                                      rot_schedule=sched, tor_schedule=sched, device=torch.device('cpu'),
                                      t_to_sigma=partial(r_du.t_to_sigma, args=sa), model_args=margs, batch_size=3,
                                      no_final_step_noise=True, confidence_model=conf,
                                      confidence_data_list=copy.deepcopy(poses), confidence_model_args=cargs)
    print('crop', crop, 'confidence', c)
    runs.append(dict(crop_beyond=crop, confidence=c, final_pos=[d['ligand'].pos.clone() for d in out_list]))
torch.save(dict(score_case=0, confidence_case=0, seed=77, schedule=sched, runs=runs),
           os.path.join(OUT, 'ref_sampling_confidence.pt'))
print('ref_sampling_confidence.pt written')
