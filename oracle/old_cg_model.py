"""Oracle restatement of models/old_cg_model.py (CGOldModel) in CONFIDENCE MODE - the ranking model sampling() calls once per
batch (utils/sampling.py:208-227).  TEST INFRASTRUCTURE.

Same constructor keywords and state_dict keys as the reference class (models/old_cg_model.py:19-201) for the supported
subset: confidence_mode=True, use_old_atom_encoder=True (the reference's new AtomEncoder does not accept the
``lm_embedding_type`` keyword this class passes, models/old_cg_model.py:63-66 vs models/layers.py:33), no miscellaneous
atoms, one noise schedule.  Wiring differences to the score model (models/cg_model.py): one OldTensorProductConvLayer per
edge type and layer (no shared multi-group layer), residual=False, the receptor is not updated in the last layer
(:271), ligand->receptor messages reuse the ligand-centred edge attributes and spherical harmonics (:275-276)."""
import numpy as np
import torch
from torch import nn
import torch.nn.functional as F

from . import e3nn_lite as o3
from .graph_ops import radius, radius_graph, scatter
from .layers import GaussianSmearing, OldAtomEncoder
from .tensor_layers import OldTensorProductConvLayer
from .cg_model import LIG_FEATURE_DIMS, REC_RESIDUE_FEATURE_DIMS


def _mlp(i, h, o, dropout):
    return nn.Sequential(nn.Linear(i, h), nn.ReLU(), nn.Dropout(dropout), nn.Linear(h, o))


class CGOldModel(nn.Module):
    def __init__(self, t_to_sigma, device, timestep_emb_func, in_lig_edge_features=4, sigma_embed_dim=32, sh_lmax=2,
                 ns=16, nv=4, num_conv_layers=2, lig_max_radius=5, rec_max_radius=30, cross_max_distance=250,
                 center_max_distance=30, distance_embed_dim=32, cross_distance_embed_dim=32, no_torsion=False,
                 scale_by_sigma=True, norm_by_sigma=True, use_second_order_repr=False, batch_norm=True,
                 dynamic_max_cross=False, dropout=0.0, smooth_edges=False, odd_parity=False,
                 separate_noise_schedule=False, lm_embedding_type=None, confidence_mode=False, confidence_dropout=0,
                 confidence_no_batchnorm=False, asyncronous_noise_schedule=False, affinity_prediction=False, parallel=1,
                 parallel_aggregators="mean max min std", num_confidence_outputs=1, fixed_center_conv=False,
                 no_aminoacid_identities=False, include_miscellaneous_atoms=False, use_old_atom_encoder=False,
                 lm_embedding_dim=1280):
        super().__init__()
        assert parallel == 1, "not implemented"
        assert confidence_mode and use_old_atom_encoder and not include_miscellaneous_atoms, "oracle subset"
        assert not (separate_noise_schedule or asyncronous_noise_schedule or use_second_order_repr), "oracle subset"
        self.t_to_sigma, self.device, self.timestep_emb_func = t_to_sigma, device, timestep_emb_func
        self.in_lig_edge_features, self.sigma_embed_dim = in_lig_edge_features, sigma_embed_dim
        self.lig_max_radius, self.rec_max_radius = lig_max_radius, rec_max_radius
        self.cross_max_distance, self.dynamic_max_cross = cross_max_distance, dynamic_max_cross
        self.sh_irreps = o3.Irreps.spherical_harmonics(lmax=sh_lmax)
        self.ns, self.nv, self.smooth_edges = ns, nv, smooth_edges
        self.confidence_mode, self.num_conv_layers = confidence_mode, num_conv_layers
        self.affinity_prediction, self.no_aminoacid_identities = affinity_prediction, no_aminoacid_identities
        kw = dict(lm_embedding_dim=lm_embedding_dim) if lm_embedding_type is not None else {}
        self.lig_node_embedding = OldAtomEncoder(ns, LIG_FEATURE_DIMS, sigma_embed_dim)
        self.lig_edge_embedding = _mlp(in_lig_edge_features + sigma_embed_dim + distance_embed_dim, ns, ns, dropout)
        self.rec_node_embedding = OldAtomEncoder(ns, REC_RESIDUE_FEATURE_DIMS, sigma_embed_dim,
                                                 lm_embedding_type=lm_embedding_type, **kw)
        self.rec_edge_embedding = _mlp(sigma_embed_dim + distance_embed_dim, ns, ns, dropout)
        self.cross_edge_embedding = _mlp(sigma_embed_dim + cross_distance_embed_dim, ns, ns, dropout)
        self.lig_distance_expansion = GaussianSmearing(0.0, lig_max_radius, distance_embed_dim)
        self.rec_distance_expansion = GaussianSmearing(0.0, rec_max_radius, distance_embed_dim)
        self.cross_distance_expansion = GaussianSmearing(0.0, cross_max_distance, cross_distance_embed_dim)
        seq = [f'{ns}x0e', f'{ns}x0e + {nv}x1o', f'{ns}x0e + {nv}x1o + {nv}x1e',
               f'{ns}x0e + {nv}x1o + {nv}x1e + {ns}x0o']
        lig, rec, l2r, r2l = [], [], [], []
        for i in range(num_conv_layers):
            p = dict(in_irreps=seq[min(i, 3)], sh_irreps=self.sh_irreps, out_irreps=seq[min(i + 1, 3)],
                     n_edge_features=3 * ns, hidden_features=3 * ns, residual=False, batch_norm=batch_norm,
                     dropout=dropout)
            lig.append(OldTensorProductConvLayer(**p))           # creation order as in the reference (:118-125)
            rec.append(OldTensorProductConvLayer(**p))
            l2r.append(OldTensorProductConvLayer(**p))
            r2l.append(OldTensorProductConvLayer(**p))
        self.lig_conv_layers, self.rec_conv_layers = nn.ModuleList(lig), nn.ModuleList(rec)
        self.lig_to_rec_conv_layers, self.rec_to_lig_conv_layers = nn.ModuleList(l2r), nn.ModuleList(r2l)
        bn = (lambda: nn.Identity()) if confidence_no_batchnorm else (lambda: nn.BatchNorm1d(ns))
        self.confidence_predictor = nn.Sequential(
            nn.Linear(2 * ns if num_conv_layers >= 3 else ns, ns), bn(), nn.ReLU(), nn.Dropout(confidence_dropout),
            nn.Linear(ns, ns), bn(), nn.ReLU(), nn.Dropout(confidence_dropout),
            nn.Linear(ns, 2 if affinity_prediction else 1))

    def _sh(self, vec):
        return o3.spherical_harmonics(self.sh_irreps, vec, normalize=True, normalization='component')

    def get_edge_weight(self, edge_vec, max_norm):                      # old_cg_model.py:353-359
        if self.smooth_edges:
            nn_ = torch.clip(edge_vec.norm(dim=-1) * np.pi / max_norm, max=np.pi)
            return 0.5 * (torch.cos(nn_) + 1.0).unsqueeze(-1)
        return 1.0

    def build_lig_conv_graph(self, data):                               # :361-391
        lig, ll = data['ligand'], data['ligand', 'ligand']
        lig.node_sigma_emb = self.timestep_emb_func(lig.node_t['tr'])
        radius_edges = radius_graph(lig.pos, self.lig_max_radius, lig.batch)
        edge_index = torch.cat([ll.edge_index, radius_edges], 1).long()
        edge_attr = torch.cat([ll.edge_attr, torch.zeros(radius_edges.shape[-1], self.in_lig_edge_features)], 0)
        edge_attr = torch.cat([edge_attr, lig.node_sigma_emb[edge_index[0]]], 1)
        node_attr = torch.cat([lig.x, lig.node_sigma_emb], 1)
        src, dst = edge_index
        vec = lig.pos[dst] - lig.pos[src]
        edge_attr = torch.cat([edge_attr, self.lig_distance_expansion(vec.norm(dim=-1))], 1)
        return node_attr, edge_index, edge_attr, self._sh(vec), self.get_edge_weight(vec, self.lig_max_radius)

    def build_rec_conv_graph(self, data):                               # :393-414
        rec = data['receptor']
        rec.node_sigma_emb = self.timestep_emb_func(rec.node_t['tr'])
        node_attr = torch.cat([rec.x, rec.node_sigma_emb], 1)
        edge_index = data['receptor', 'receptor'].edge_index.long()
        src, dst = edge_index
        vec = rec.pos[dst] - rec.pos[src]
        edge_attr = torch.cat([rec.node_sigma_emb[src], self.rec_distance_expansion(vec.norm(dim=-1))], 1)
        return node_attr, edge_index, edge_attr, self._sh(vec), self.get_edge_weight(vec, self.rec_max_radius)

    def build_cross_conv_graph(self, data, cutoff):                     # :439-461
        lig, rec = data['ligand'], data['receptor']
        if torch.is_tensor(cutoff):
            edge_index = radius(rec.pos / cutoff[rec.batch], lig.pos / cutoff[lig.batch], 1, rec.batch, lig.batch,
                                max_num_neighbors=10000)
        else:
            edge_index = radius(rec.pos, lig.pos, cutoff, rec.batch, lig.batch, max_num_neighbors=10000)
        src, dst = edge_index
        vec = rec.pos[dst] - lig.pos[src]
        edge_attr = torch.cat([lig.node_sigma_emb[src], self.cross_distance_expansion(vec.norm(dim=-1))], 1)
        cutoff_d = cutoff[lig.batch[src]].squeeze() if torch.is_tensor(cutoff) else cutoff
        return edge_index, edge_attr, self._sh(vec), self.get_edge_weight(vec, cutoff_d)

    def forward(self, data):                                            # :203-301
        if self.no_aminoacid_identities:
            data['receptor'].x = data['receptor'].x * 0
        tr_sigma = data.complex_t['tr']                                 # confidence mode: times are passed as they are
        ns = self.ns
        lig_node, lig_ei, lig_ea, lig_sh, lig_ew = self.build_lig_conv_graph(data)
        lig_src, lig_dst = lig_ei
        lig_node, lig_ea = self.lig_node_embedding(lig_node), self.lig_edge_embedding(lig_ea)
        rec_node, rec_ei, rec_ea, rec_sh, rec_ew = self.build_rec_conv_graph(data)
        rec_src, rec_dst = rec_ei
        rec_node, rec_ea = self.rec_node_embedding(rec_node), self.rec_edge_embedding(rec_ea)
        cutoff = (tr_sigma * 3 + 20).unsqueeze(1) if self.dynamic_max_cross else self.cross_max_distance
        lr_ei, lr_ea, lr_sh, lr_ew = self.build_cross_conv_graph(data, cutoff)
        cross_lig, cross_rec = lr_ei
        lr_ea = self.cross_edge_embedding(lr_ea)
        L = len(self.lig_conv_layers)
        for l in range(L):
            ea_ = torch.cat([lig_ea, lig_node[lig_src, :ns], lig_node[lig_dst, :ns]], -1)
            lig_intra = self.lig_conv_layers[l](lig_node, lig_ei, ea_, lig_sh, edge_weight=lig_ew)
            ea_ = torch.cat([lr_ea, lig_node[cross_lig, :ns], rec_node[cross_rec, :ns]], -1)
            lig_inter = self.rec_to_lig_conv_layers[l](rec_node, lr_ei, ea_, lr_sh, out_nodes=lig_node.shape[0],
                                                       edge_weight=lr_ew)
            if l != L - 1:
                ea_ = torch.cat([rec_ea, rec_node[rec_src, :ns], rec_node[rec_dst, :ns]], -1)
                rec_intra = self.rec_conv_layers[l](rec_node, rec_ei, ea_, rec_sh, edge_weight=rec_ew)
                ea_ = torch.cat([lr_ea, lig_node[cross_lig, :ns], rec_node[cross_rec, :ns]], -1)
                rec_inter = self.lig_to_rec_conv_layers[l](lig_node, torch.flip(lr_ei, dims=[0]), ea_, lr_sh,
                                                           out_nodes=rec_node.shape[0], edge_weight=lr_ew)
            lig_node = F.pad(lig_node, (0, lig_intra.shape[-1] - lig_node.shape[-1])) + lig_intra + lig_inter
            if l != L - 1:
                rec_node = F.pad(rec_node, (0, rec_intra.shape[-1] - rec_node.shape[-1])) + rec_intra + rec_inter
        scal = torch.cat([lig_node[:, :ns], lig_node[:, -ns:]], 1) if self.num_conv_layers >= 3 else lig_node[:, :ns]
        pooled = scatter(scal, data['ligand'].batch, dim=0, dim_size=data.num_graphs, reduce='mean')
        return self.confidence_predictor(pooled).squeeze(dim=-1)
