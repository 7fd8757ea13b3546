"""Golden-vector generator: runs the UNMODIFIED reference code from /root/reference in the authoring container and
stores its inputs/outputs as fixtures under tests/golden/.  /root/reference does not exist on the GPU box, so tests only
ever read the fixtures.

    cd /tmp/tables && python /root/repo/tests/golden/make_golden.py      # cwd holds utils/so3.py's .npy caches

Third-party packages the reference imports but that cannot be installed here (e3nn, torch_scatter, torch_cluster,
torch_geometric, rdkit, ...) are supplied by oracle/ref_shims.py; everything under models/ and utils/ is the
reference's own code.  Fixtures:
  ref_layers.pt         models/layers.py (GaussianSmearing, AtomEncoder), utils/diffusion_utils.py (sinusoidal_embedding,
                        get_t_schedule, t_to_sigma), utils/geometry.py (axis_angle_to_matrix, Kabsch batch)
  ref_faster_tp.pt      models/tensor_layers.py FasterTensorProduct (self-contained arithmetic: pins l<=1 CG + norms)
  ref_conv_layer.pt     models/tensor_layers.py TensorProductConvLayer (1 and 4 edge groups; fctp and faster)
  ref_cg_model.pt       models/cg_model.py CGModel.forward, score mode (lmax 2 and 1)
  ref_conformer.pt      utils/diffusion_utils.py modify_conformer_batch (torsion.py:75-90 + geometry.py:246-276)
  ref_sampling.pt       utils/sampling.py sampling(): 4-step trajectory, seeded noise, default-yaml temperatures
"""
import copy
import os
import sys
from argparse import Namespace
from functools import partial

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install()
OUT = os.path.join(ROOT, 'tests', 'golden')
torch.set_num_threads(4)

import models.layers as r_layers            # noqa: E402
import models.tensor_layers as r_tl         # noqa: E402
import utils.diffusion_utils as r_du        # noqa: E402
import utils.geometry as r_geo              # noqa: E402
from utils import torus as r_torus          # noqa: E402
import models.cg_model as r_cg              # noqa: E402
import utils.sampling as r_sampling         # noqa: E402

from diffdock_b200.hetero import collate, graph_to_dict   # noqa: E402
from diffdock_b200.synthetic import default_model_args, make_pose_list   # noqa: E402
from tests.parity_helpers import rand_bn_    # noqa: E402

# use the stored Monte-Carlo torus table instance (the import above re-drew it)
_tab = np.load(os.path.join(ROOT, 'diffdock_b200', 'tables', 'score_norm_tables.npz'))
r_torus.score_norm_ = _tab['torus_score_norm']
from utils import so3 as r_so3              # noqa: E402
assert np.array_equal(np.nan_to_num(r_so3._exp_score_norms), np.nan_to_num(_tab['so3_exp_score_norms']))


def save(name, obj):
    torch.save(obj, os.path.join(OUT, name))
    print(name, os.path.getsize(os.path.join(OUT, name)) // 1024, 'KiB')


# ------------------------------------------------------------------------------------------------ layers / geometry
g = torch.Generator().manual_seed(0)
gs = r_layers.GaussianSmearing(0.0, 5.0, 16)
d = torch.rand(40, generator=g) * 6
torch.manual_seed(0)
enc = r_layers.AtomEncoder(8, ([5, 3, 7], 0), sigma_embed_dim=4, lm_embedding_dim=6)
xe = torch.cat([torch.stack([torch.randint(0, k, (9,), generator=g) for k in (5, 3, 7)], 1).float(),
                torch.randn(9, 10, generator=g)], 1)
aa = torch.randn(12, 3, generator=g)
aa[0] *= 1e-8
A, Bm = torch.randn(5, 11, 3, generator=g), torch.randn(5, 11, 3, generator=g)
R, t = r_geo.rigid_transform_Kabsch_3D_torch_batch(A, Bm)
targs = Namespace(tr_sigma_min=0.1, tr_sigma_max=19.0, rot_sigma_min=0.03, rot_sigma_max=1.55, tor_sigma_min=0.0314,
                  tor_sigma_max=3.14)
save('ref_layers.pt', dict(
    gs_in=d, gs_out=gs(d), gs_coeff=gs.coeff,
    enc_state=enc.state_dict(), enc_in=xe, enc_out=enc(xe).detach(),
    sin_in=torch.tensor([0.0, 0.05, 0.5, 1.0]), sin_out=r_du.sinusoidal_embedding(1000 * torch.tensor([0.0, 0.05, 0.5, 1.0]), 16),
    sched20=r_du.get_t_schedule('expbeta', 20), t2s=[float(v) for v in r_du.t_to_sigma(0.3, 0.3, 0.3, targs)],
    aa_in=aa, aa_out=r_geo.axis_angle_to_matrix(aa), kabsch_A=A, kabsch_B=Bm, kabsch_R=R, kabsch_t=t))

# ------------------------------------------------------------------------------------------------ FasterTensorProduct
seq = r_tl.get_irrep_seq(6, 3, False, False)
cases = []
for i in range(4):
    ins, outs = seq[i], seq[min(i + 1, 3)]
    tp = r_tl.FasterTensorProduct(ins, '1x0e+1x1o', outs)
    E = 7
    x = torch.randn(E, r_tl.irrep_to_size(ins), generator=g)
    v = torch.randn(E, 3, generator=g)
    v = v / v.norm(dim=-1, keepdim=True)
    sh = torch.cat([torch.ones(E, 1), np.sqrt(3.0) * v], 1).float()
    w = torch.randn(E, tp.weight_numel, generator=g)
    cases.append(dict(in_irreps=ins, out_irreps=outs, x=x, sh=sh, vec=v, w=w, out=tp(x, sh, w)))
save('ref_faster_tp.pt', cases)


# ------------------------------------------------------------------------------------------------ conv layer
def conv_case(ins, shs, outs, groups, faster, seed, residual=True, out_nodes=None, reduce='mean'):
    torch.manual_seed(seed)
    layer = r_tl.TensorProductConvLayer(ins, shs, outs, n_edge_features=12, hidden_features=12, residual=residual,
                                        faster=faster, edge_groups=groups).eval()
    gg = torch.Generator().manual_seed(seed + 1)
    rand_bn_(layer.batch_norm, gg)
    N, E = 10, 60
    x = torch.randn(N, r_tl.irrep_to_size(ins), generator=gg)
    nt = out_nodes or N
    ei = torch.stack([torch.randint(0, nt, (E,), generator=gg), torch.randint(0, N, (E,), generator=gg)])
    vec = torch.randn(E, 3, generator=gg)
    from e3nn import o3
    sh = o3.spherical_harmonics(o3.Irreps(shs), vec, normalize=True, normalization='component')
    ea = torch.randn(E, 12, generator=gg)
    ea_in = [ea[:10], ea[10:35], ea[35:36], ea[36:]] if groups == 4 else ea
    ew = torch.rand(E, 1, generator=gg)
    with torch.no_grad():
        out = layer(x, ei, ea_in, sh, out_nodes=out_nodes, reduce=reduce, edge_weight=ew)
    return dict(in_irreps=ins, sh_irreps=shs, out_irreps=outs, groups=groups, faster=faster, residual=residual,
                out_nodes=out_nodes, reduce=reduce, state=layer.state_dict(), x=x, edge_index=ei, vec=vec, sh=sh,
                edge_attr=ea, group_cuts=[10, 35, 36], edge_weight=ew, out=out)


sh2, sh1 = '1x0e+1x1o+1x2e', '1x0e+1x1o'
save('ref_conv_layer.pt', [
    conv_case(seq[3], sh2, seq[3], 1, False, 1),
    conv_case(seq[2], sh2, seq[3], 4, False, 2),
    conv_case(seq[3], sh1, seq[3], 1, True, 3),   # reference: faster + multigroup crashes (tensor_layers.py:199)
    conv_case(seq[1], sh1, seq[2], 1, True, 4),
    conv_case(seq[3], sh2, '2x1o + 2x1e', 1, False, 5, residual=False, out_nodes=3),
    conv_case(seq[0], sh2, seq[1], 1, False, 6, reduce='sum'),
])


# ------------------------------------------------------------------------------------------------ CGModel
def model_case(lmax, seed, n_poses=3, t=0.4, lm=True, **over):
    a = default_model_args(ns=6, nv=3, sh_lmax=lmax, num_conv_layers=3, distance_embed_dim=8,
                           cross_distance_embed_dim=8, sigma_embed_dim=8, **over)
    kw = dict(sigma_embed_dim=8, sh_lmax=lmax, ns=6, nv=3, num_conv_layers=3, lig_max_radius=a.max_radius,
              rec_max_radius=a.rec_max_radius, cross_max_distance=a.cross_max_distance,
              center_max_distance=a.center_max_distance, distance_embed_dim=8, cross_distance_embed_dim=8,
              dynamic_max_cross=True, lm_embedding_type='precomputed' if lm else None, embed_also_ligand=True,
              num_prot_emb_layers=a.num_prot_emb_layers,
              differentiate_convolutions=a.differentiate_convolutions)
    torch.manual_seed(seed)
    model = r_cg.CGModel(partial(r_du.t_to_sigma, args=a), torch.device('cpu'),
                         r_du.get_timestep_embedding('sinusoidal', 8, a.embedding_scale), **kw).eval()
    gg = torch.Generator().manual_seed(seed + 1)
    for m in model.modules():
        if m.__class__.__name__ == 'BatchNorm':
            rand_bn_(m, gg)
    poses = make_pose_list(n_poses, n_res=24, n_atoms=9, seed=seed + 2, tr_sigma_max=a.tr_sigma_max * t,
                           lm_dim=16 if lm else 0)
    if lm:   # shrink the LM embedding (1280 -> 16) to keep the fixture small: patch the encoder's input Linear
        torch.manual_seed(seed + 5)
        model.rec_node_embedding.additional_features_dim = 16
        model.rec_node_embedding.additional_features_embedder = torch.nn.Linear(16 + 6, 6)
    batch = collate(copy.deepcopy(poses))
    r_du.set_time(batch, t, t, t, t, n_poses, False, 'cpu')
    with torch.no_grad():
        tr, rot, tor, _ = model(batch)
    return dict(args=vars(a), kw=kw, lm_dim=16 if lm else 0, state=model.state_dict(),
                poses=[graph_to_dict(p) for p in poses], t=t, tr=tr, rot=rot, tor=tor), model, a, poses


c2, m2, a2, p2 = model_case(2, 10)
c1, m1, a1, p1 = model_case(1, 11, t=0.9, differentiate_convolutions=False)
c3, _, _, _ = model_case(2, 12, t=0.1, num_prot_emb_layers=1)
save('ref_cg_model.pt', [c2, c1, c3])

# ------------------------------------------------------------------------------------------------ conformer update
batch = collate(copy.deepcopy(p2))
mask_rotate = torch.from_numpy(p2[0]['ligand'].mask_rotate[0])
nb = int(mask_rotate.shape[0])
gg = torch.Generator().manual_seed(20)
tr_u, rot_u, tor_u = torch.randn(3, 3, generator=gg), 0.3 * torch.randn(3, 3, generator=gg), torch.randn(3 * nb, generator=gg)
new_pos = r_du.modify_conformer_batch(batch['ligand'].pos, batch, tr_u, rot_u, tor_u, mask_rotate)
rigid_pos = r_du.modify_conformer_batch(batch['ligand'].pos, batch, tr_u, rot_u, None, mask_rotate)
save('ref_conformer.pt', dict(poses=[graph_to_dict(p) for p in p2], tr=tr_u, rot=rot_u, tor=tor_u, new_pos=new_pos,
                              rigid_pos=rigid_pos))

# ------------------------------------------------------------------------------------------------ sampling loop
steps = 4
sched = r_du.get_t_schedule('expbeta', steps)
data_list = copy.deepcopy(p2)
torch.manual_seed(123)
margs = Namespace(**vars(a2))
margs.crop_beyond = None
out_list, _ = r_sampling.sampling(data_list=data_list, model=m2, inference_steps=steps, tr_schedule=sched,
                                  rot_schedule=sched, tor_schedule=sched, device=torch.device('cpu'),
                                  t_to_sigma=partial(r_du.t_to_sigma, args=a2), model_args=margs, batch_size=3,
                                  no_final_step_noise=True,
                                  temp_sampling=[1.170050527854316, 2.06391612594481, 7.044261621607846],
                                  temp_psi=[0.727287304570729, 0.9022615585677628, 0.5946212391366862],
                                  temp_sigma_data=[0.9299802531572672, 0.7464326999906034, 0.6943254174849822])
save('ref_sampling.pt', dict(model_case=0, steps=steps, seed=123, schedule=sched,
                             final_pos=[d['ligand'].pos.clone() for d in out_list]))

# same run with per-step receptor cropping (utils/sampling.py:104-109 -> utils/utils.py:388-413)
import utils.utils as r_utils                # noqa: E402
kept = []
_orig_crop = r_utils.crop_beyond


def _spy(graph, cutoff, all_atoms):
    _orig_crop(graph, cutoff, all_atoms)
    kept.append(int(graph['receptor'].pos.shape[0]))


r_sampling.crop_beyond = _spy
data_list = copy.deepcopy(p2)
torch.manual_seed(321)
margs.crop_beyond = 7.0
sched_full = sched
sched = np.array([0.30, 0.22, 0.15, 0.08])     # late, small-sigma steps: the cut-off 3*sigma_tr + 7 A crops partially
out_list, _ = r_sampling.sampling(data_list=data_list, model=m2, inference_steps=steps, tr_schedule=sched,
                                  rot_schedule=sched, tor_schedule=sched, device=torch.device('cpu'),
                                  t_to_sigma=partial(r_du.t_to_sigma, args=a2), model_args=margs, batch_size=3,
                                  no_final_step_noise=True, temp_sampling=[1.17, 2.06, 7.04],
                                  temp_psi=[0.73, 0.90, 0.59], temp_sigma_data=[0.93, 0.75, 0.69])
print('residues kept per (step, pose):', kept)
save('ref_sampling_crop.pt', dict(model_case=0, steps=steps, seed=321, schedule=sched, crop_beyond=7.0, kept=kept,
                                  final_pos=[d['ligand'].pos.clone() for d in out_list]))
print('done')
