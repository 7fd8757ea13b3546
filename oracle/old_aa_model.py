"""Oracle restatement of models/old_aa_model.py (AAOldModel) in CONFIDENCE MODE - the all-atom ranking model that
inference.py:192,209 selects when the confidence model's parameters say ``all_atoms`` and that utils/sampling.py:208-227 calls
once per batch of final poses.  TEST INFRASTRUCTURE.

Same constructor keywords and state_dict keys as the reference class (models/old_aa_model.py:21-150) for the supported subset:
confidence_mode=True, use_old_atom_encoder=True (the reference's new AtomEncoder does not accept the ``lm_embedding_type``
keyword this class passes, models/old_aa_model.py:71 vs models/layers.py:33), one noise schedule, parallel=1.
Three node types (ligand atoms, receptor residues, receptor atoms), nine OldTensorProductConvLayer per interaction layer
(:105-121): ligand<-ligand, ligand<-residue, ligand<-atom, atom<-atom, atom<-ligand, atom<-residue, residue<-residue,
residue<-ligand, residue<-atom; the messages of the reversed directions reuse the forward edge attributes and spherical
harmonics (:246-266, SURVEY.md hazard C.7); atoms and residues are not updated in the last layer (:241)."""
import numpy as np
import torch
from torch import nn
import torch.nn.functional as F

from . import e3nn_lite as o3
from .graph_ops import radius, radius_graph, scatter
from .layers import GaussianSmearing, OldAtomEncoder
from .tensor_layers import OldTensorProductConvLayer
from .cg_model import LIG_FEATURE_DIMS, REC_RESIDUE_FEATURE_DIMS

REC_ATOM_FEATURE_DIMS = ([38, 119, 23, 38], 0)       # datasets/process_mols.py:78-83


def _mlp(i, h, o, dropout):
    return nn.Sequential(nn.Linear(i, h), nn.ReLU(), nn.Dropout(dropout), nn.Linear(h, o))


class AAOldModel(nn.Module):
    def __init__(self, t_to_sigma, device, timestep_emb_func, in_lig_edge_features=4, sigma_embed_dim=32, sh_lmax=2,
                 ns=16, nv=4, num_conv_layers=2, lig_max_radius=5, rec_max_radius=30, cross_max_distance=250,
                 center_max_distance=30, distance_embed_dim=32, cross_distance_embed_dim=32, no_torsion=False,
                 scale_by_sigma=True, norm_by_sigma=True, use_second_order_repr=False, batch_norm=True,
                 dynamic_max_cross=False, dropout=0.0, smooth_edges=False, odd_parity=False,
                 separate_noise_schedule=False, lm_embedding_type=False, confidence_mode=False, confidence_dropout=0,
                 confidence_no_batchnorm=False, asyncronous_noise_schedule=False, affinity_prediction=False, parallel=1,
                 parallel_aggregators="mean max min std", num_confidence_outputs=1, fixed_center_conv=False,
                 no_aminoacid_identities=False, include_miscellaneous_atoms=False, use_old_atom_encoder=False,
                 lm_embedding_dim=1280):
        super().__init__()
        assert parallel == 1 and confidence_mode and use_old_atom_encoder, "oracle subset"
        assert not (separate_noise_schedule or asyncronous_noise_schedule or use_second_order_repr), "oracle subset"
        lm_embedding_type = lm_embedding_type or None
        self.t_to_sigma, self.device, self.timestep_emb_func = t_to_sigma, device, timestep_emb_func
        self.in_lig_edge_features, self.sigma_embed_dim = in_lig_edge_features, sigma_embed_dim
        self.lig_max_radius, self.rec_max_radius = lig_max_radius, rec_max_radius
        self.cross_max_distance, self.dynamic_max_cross = cross_max_distance, dynamic_max_cross
        self.sh_irreps = o3.Irreps.spherical_harmonics(lmax=sh_lmax)
        self.ns, self.nv, self.smooth_edges = ns, nv, smooth_edges
        self.confidence_mode, self.num_conv_layers = confidence_mode, num_conv_layers
        self.affinity_prediction, self.no_aminoacid_identities = affinity_prediction, no_aminoacid_identities
        S, D, Dx = sigma_embed_dim, distance_embed_dim, cross_distance_embed_dim
        kw = dict(lm_embedding_dim=lm_embedding_dim) if lm_embedding_type is not None else {}
        self.lig_node_embedding = OldAtomEncoder(ns, LIG_FEATURE_DIMS, S)
        self.lig_edge_embedding = _mlp(in_lig_edge_features + S + D, ns, ns, dropout)
        self.rec_node_embedding = OldAtomEncoder(ns, REC_RESIDUE_FEATURE_DIMS, S, lm_embedding_type=lm_embedding_type, **kw)
        self.rec_edge_embedding = _mlp(S + D, ns, ns, dropout)
        self.atom_node_embedding = OldAtomEncoder(ns, REC_ATOM_FEATURE_DIMS, S)
        self.atom_edge_embedding = _mlp(S + D, ns, ns, dropout)
        self.lr_edge_embedding = _mlp(S + Dx, ns, ns, dropout)
        self.ar_edge_embedding = _mlp(S + D, ns, ns, dropout)
        self.la_edge_embedding = _mlp(S + Dx, ns, ns, dropout)
        self.lig_distance_expansion = GaussianSmearing(0.0, lig_max_radius, D)
        self.rec_distance_expansion = GaussianSmearing(0.0, rec_max_radius, D)
        self.cross_distance_expansion = GaussianSmearing(0.0, cross_max_distance, Dx)
        seq = [f'{ns}x0e', f'{ns}x0e + {nv}x1o', f'{ns}x0e + {nv}x1o + {nv}x1e',
               f'{ns}x0e + {nv}x1o + {nv}x1e + {ns}x0o']
        convs = []
        for i in range(num_conv_layers):
            p = dict(in_irreps=seq[min(i, 3)], sh_irreps=self.sh_irreps, out_irreps=seq[min(i + 1, 3)],
                     n_edge_features=3 * ns, residual=False, batch_norm=batch_norm, dropout=dropout)
            convs += [OldTensorProductConvLayer(**p) for _ in range(9)]       # 3 intra & 6 inter per layer (:119-120)
        self.conv_layers = nn.ModuleList(convs)
        bn = (lambda: nn.Identity()) if confidence_no_batchnorm else (lambda: nn.BatchNorm1d(ns))
        out_dim = (num_confidence_outputs + 1) if affinity_prediction else num_confidence_outputs
        self.confidence_predictor = nn.Sequential(
            nn.Linear(2 * ns if num_conv_layers >= 3 else ns, ns), bn(), nn.ReLU(), nn.Dropout(confidence_dropout),
            nn.Linear(ns, ns), bn(), nn.ReLU(), nn.Dropout(confidence_dropout), nn.Linear(ns, out_dim))

    def _sh(self, vec):
        return o3.spherical_harmonics(self.sh_irreps, vec, normalize=True, normalization='component')

    def get_edge_weight(self, edge_vec, max_norm):                      # models/old_aa_model.py:352-356
        if self.smooth_edges:
            nn_ = torch.clip(edge_vec.norm(dim=-1) * np.pi / max_norm, max=np.pi)
            return 0.5 * (torch.cos(nn_) + 1.0).unsqueeze(-1)
        return 1.0

    def build_lig_conv_graph(self, data):                               # :358-398
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

    def _static_graph(self, data, nt, expansion, max_r):                # :400-445 (residues: rec expansion, atoms: lig expansion)
        st = data[nt]
        st.node_sigma_emb = self.timestep_emb_func(st.node_t['tr'])
        node_attr = torch.cat([st.x, st.node_sigma_emb], 1)
        edge_index = data[nt, nt].edge_index.long()
        src, dst = edge_index
        vec = st.pos[dst] - st.pos[src]
        edge_attr = torch.cat([st.node_sigma_emb[src], expansion(vec.norm(dim=-1))], 1)
        return node_attr, edge_index, edge_attr, self._sh(vec), self.get_edge_weight(vec, max_r)

    def build_cross_conv_graph(self, data, cutoff):                     # :447-491
        lig, rec, atom = data['ligand'], data['receptor'], data['atom']
        if torch.is_tensor(cutoff):
            lr = radius(rec.pos / cutoff[rec.batch], lig.pos / cutoff[lig.batch], 1, rec.batch, lig.batch,
                        max_num_neighbors=10000)
        else:
            lr = radius(rec.pos, lig.pos, cutoff, rec.batch, lig.batch, max_num_neighbors=10000)
        lr_vec = rec.pos[lr[1]] - lig.pos[lr[0]]
        lr_attr = torch.cat([lig.node_sigma_emb[lr[0]], self.cross_distance_expansion(lr_vec.norm(dim=-1))], 1)
        cutoff_d = cutoff[lig.batch[lr[0]]].squeeze() if torch.is_tensor(cutoff) else cutoff
        lr_w = self.get_edge_weight(lr_vec, cutoff_d)
        la = radius(atom.pos, lig.pos, self.lig_max_radius, atom.batch, lig.batch, max_num_neighbors=10000)
        la_vec = atom.pos[la[1]] - lig.pos[la[0]]
        la_attr = torch.cat([lig.node_sigma_emb[la[0]], self.cross_distance_expansion(la_vec.norm(dim=-1))], 1)
        la_w = self.get_edge_weight(la_vec, self.lig_max_radius)
        ar = data['atom', 'receptor'].edge_index.long()
        ar_vec = rec.pos[ar[1]] - atom.pos[ar[0]]
        ar_attr = torch.cat([atom.node_sigma_emb[ar[0]], self.rec_distance_expansion(ar_vec.norm(dim=-1))], 1)
        return (lr, lr_attr, self._sh(lr_vec), lr_w, la, la_attr, self._sh(la_vec), la_w, ar, ar_attr, self._sh(ar_vec), 1)

    def forward(self, data):                                            # :202-286
        if self.no_aminoacid_identities:
            data['receptor'].x = data['receptor'].x * 0
        tr_sigma = data.complex_t['tr']                                 # confidence mode: times are passed as they are (:209)
        ns, L, C = self.ns, self.num_conv_layers, self.conv_layers
        lig, lig_ei, lig_ea, lig_sh, lig_w = self.build_lig_conv_graph(data)
        lig, lig_ea = self.lig_node_embedding(lig), self.lig_edge_embedding(lig_ea)
        rec, rec_ei, rec_ea, rec_sh, rec_w = self._static_graph(data, 'receptor', self.rec_distance_expansion, self.rec_max_radius)
        rec, rec_ea = self.rec_node_embedding(rec), self.rec_edge_embedding(rec_ea)
        atom, at_ei, at_ea, at_sh, at_w = self._static_graph(data, 'atom', self.lig_distance_expansion, self.lig_max_radius)
        atom, at_ea = self.atom_node_embedding(atom), self.atom_edge_embedding(at_ea)
        cutoff = (tr_sigma * 3 + 20).unsqueeze(1) if self.dynamic_max_cross else self.cross_max_distance
        lr, lr_ea, lr_sh, lr_w, la, la_ea, la_sh, la_w, ar, ar_ea, ar_sh, ar_w = self.build_cross_conv_graph(data, cutoff)
        lr_ea, la_ea, ar_ea = self.lr_edge_embedding(lr_ea), self.la_edge_embedding(la_ea), self.ar_edge_embedding(ar_ea)
        cat = lambda e, a, b: torch.cat([e, a[:, :ns], b[:, :ns]], -1)
        for l in range(L):
            lig_up = C[9 * l](lig, lig_ei, cat(lig_ea, lig[lig_ei[0]], lig[lig_ei[1]]), lig_sh, edge_weight=lig_w)
            lr_up = C[9 * l + 1](rec, lr, cat(lr_ea, lig[lr[0]], rec[lr[1]]), lr_sh, out_nodes=lig.shape[0], edge_weight=lr_w)
            la_up = C[9 * l + 2](atom, la, cat(la_ea, lig[la[0]], atom[la[1]]), la_sh, out_nodes=lig.shape[0], edge_weight=la_w)
            if l != L - 1:
                at_up = C[9 * l + 3](atom, at_ei, cat(at_ea, atom[at_ei[0]], atom[at_ei[1]]), at_sh, edge_weight=at_w)
                al_up = C[9 * l + 4](lig, torch.flip(la, dims=[0]), cat(la_ea, atom[la[1]], lig[la[0]]), la_sh,
                                     out_nodes=atom.shape[0], edge_weight=la_w)
                ar_up = C[9 * l + 5](rec, ar, cat(ar_ea, atom[ar[0]], rec[ar[1]]), ar_sh, out_nodes=atom.shape[0], edge_weight=ar_w)
                rec_up = C[9 * l + 6](rec, rec_ei, cat(rec_ea, rec[rec_ei[0]], rec[rec_ei[1]]), rec_sh, edge_weight=rec_w)
                rl_up = C[9 * l + 7](lig, torch.flip(lr, dims=[0]), cat(lr_ea, rec[lr[1]], lig[lr[0]]), lr_sh,
                                     out_nodes=rec.shape[0], edge_weight=lr_w)
                ra_up = C[9 * l + 8](atom, torch.flip(ar, dims=[0]), cat(ar_ea, rec[ar[1]], atom[ar[0]]), ar_sh,
                                     out_nodes=rec.shape[0], edge_weight=ar_w)
            lig = F.pad(lig, (0, lig_up.shape[-1] - lig.shape[-1])) + lig_up + la_up + lr_up
            if l != L - 1:
                atom = F.pad(atom, (0, at_up.shape[-1] - atom.shape[-1])) + at_up + al_up + ar_up
                rec = F.pad(rec, (0, rec_up.shape[-1] - rec.shape[-1])) + rec_up + ra_up + rl_up
        scal = torch.cat([lig[:, :ns], lig[:, -ns:]], 1) if L >= 3 else lig[:, :ns]
        pooled = scatter(scal, data['ligand'].batch, dim=0, dim_size=data.num_graphs, reduce='mean')
        return self.confidence_predictor(pooled).squeeze(dim=-1)
