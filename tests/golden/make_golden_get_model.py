"""Golden vectors of the model factory: runs the UNMODIFIED ``get_model`` (utils/utils.py:172-281) from /root/reference with its
four model classes replaced by recorders and stores, per argument namespace, the class it picked and the keywords it passed
(tests/golden/ref_get_model.pt).

    cd /tmp/tables && python /root/repo/tests/golden/make_golden_get_model.py 2>/dev/null
"""
import os
import sys
from argparse import Namespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install()
import utils.utils as U   # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'ref_get_model.pt')


def recorder(name):
    class R:
        def __init__(self, **kw):
            self.name, self.kw = name, kw

        def to(self, device):
            return self
    return R


for n in ('CGModel', 'AAModel', 'CGOldModel', 'AAOldModel'):
    setattr(U, n, recorder(n))

BASE = dict(no_torsion=False, num_conv_layers=6, max_radius=5.0, scale_by_sigma=True, sigma_embed_dim=64, ns=48, nv=10,
            distance_embed_dim=64, cross_distance_embed_dim=64, no_batch_norm=False, dropout=0.1, use_second_order_repr=False,
            cross_max_distance=80, dynamic_max_cross=True, esm_embeddings_path=None)
CASES = [
    ('v1.0-style score yml (few flags)', dict(BASE), dict(old=False)),
    ('DiffDock-L score yml', dict(BASE, embedding_type='sinusoidal', embedding_scale=1000, sh_lmax=1, no_differentiate_convolutions=False,
                                  tp_weights_layers=2, num_prot_emb_layers=3, reduce_pseudoscalars=True, embed_also_ligand=True,
                                  not_fixed_center_conv=False, norm_by_sigma=True, smooth_edges=True, odd_parity=True,
                                  pdbbind_esm_embeddings_path='data/esm2.pt', no_aminoacid_identities=False,
                                  include_miscellaneous_atoms=False, depthwise_convolution=False, parallel=1,
                                  parallel_aggregators='mean max', affinity_prediction=False), dict(old=False)),
    ('all-atom score model, fourier embedding, esm model name', dict(BASE, all_atoms=True, embedding_type='fourier', embedding_scale=30,
                                                                     esm_embeddings_model='esm2_t33', atom_confidence_loss_weight=0.5,
                                                                     atom_rmsd_classification_cutoff=[1.0, 2.0], sidechain_loss_weight=0.0,
                                                                     backbone_loss_weight=0.2, not_fixed_center_conv=True), dict(old=False)),
    ('new-style confidence model', dict(BASE, rmsd_classification_cutoff=[2.0, 5.0], moad_esm_embeddings_path=None,
                                        no_differentiate_convolutions=True), dict(old=False, confidence_mode=True)),
    ('old confidence model (CG)', dict(BASE, rmsd_classification_cutoff=2.0, esm_embeddings_path='emb', use_old_atom_encoder=False,
                                       not_fixed_center_conv=True), dict(old=True, confidence_mode=True)),
    ('old confidence model (all atoms)', dict(BASE, all_atoms=True, rmsd_classification_cutoff=[2.0], include_miscellaneous_atoms=True,
                                              no_aminoacid_identities=True), dict(old=True, confidence_mode=True)),
]

if __name__ == '__main__':
    fx = []
    t = torch.tensor([0.0, 0.3, 1.0])
    for label, a, kw in CASES:
        m = U.get_model(Namespace(**a), torch.device('cpu'), t_to_sigma='T2S', no_parallel=True, **kw)
        k = dict(m.kw)
        assert k.pop('t_to_sigma') == 'T2S' and k.pop('device') == torch.device('cpu')
        emb = k.pop('timestep_emb_func')
        fx.append({'label': label, 'args': a, 'call': kw, 'class': m.name, 'kwargs': k, 'emb_of_t': emb(t)})
        print(label, m.name, len(k))
    torch.save(fx, OUT)
    print(OUT, os.path.getsize(OUT) // 1024, 'KiB')
