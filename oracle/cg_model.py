"""Oracle restatement of models/cg_model.py (CGModel, score mode).  TEST INFRASTRUCTURE.

Same constructor keywords and state_dict keys as the reference class (models/cg_model.py:20-255) for
the supported subset: score mode (confidence_mode=False), lm_embedding_type in {None,'precomputed'},
no misc atoms / side-chain head / depthwise convolution."""
import numpy as np
import torch
from torch import nn

from . import e3nn_lite as o3
from .graph_ops import radius, radius_graph
from .layers import AtomEncoder, GaussianSmearing
from .tables import so3_score_norm, torus_score_norm
from .tensor_layers import TensorProductConvLayer, get_irrep_seq

LIG_FEATURE_DIMS = ([119, 4, 12, 12, 8, 10, 6, 6, 2, 8, 2, 2, 2, 2, 2, 2], 0)   # datasets/process_mols.py:59-76
REC_RESIDUE_FEATURE_DIMS = ([38], 0)                                           # datasets/process_mols.py:85-87


def _mlp(i, h, o, dropout):
    return nn.Sequential(nn.Linear(i, h), nn.ReLU(), nn.Dropout(dropout), nn.Linear(h, o))


class CGModel(nn.Module):
    def __init__(self, t_to_sigma, device, timestep_emb_func, in_lig_edge_features=4, sigma_embed_dim=32, sh_lmax=2,
                 ns=16, nv=4, num_conv_layers=2, lig_max_radius=5, rec_max_radius=30, cross_max_distance=250,
                 center_max_distance=30, distance_embed_dim=32, cross_distance_embed_dim=32, no_torsion=False,
                 scale_by_sigma=True, norm_by_sigma=True, use_second_order_repr=False, batch_norm=True,
                 dynamic_max_cross=False, dropout=0.0, smooth_edges=False, odd_parity=False,
                 separate_noise_schedule=False, lm_embedding_type=None, confidence_mode=False,
                 differentiate_convolutions=True, tp_weights_layers=2, num_prot_emb_layers=0,
                 reduce_pseudoscalars=False, embed_also_ligand=False, fixed_center_conv=False,
                 no_aminoacid_identities=False, **unused):
        super().__init__()
        assert not confidence_mode and not separate_noise_schedule
        assert lm_embedding_type in (None, 'precomputed')
        self.t_to_sigma, self.device, self.timestep_emb_func = t_to_sigma, device, timestep_emb_func
        self.in_lig_edge_features, self.sigma_embed_dim = in_lig_edge_features, sigma_embed_dim
        self.lig_max_radius, self.rec_max_radius = lig_max_radius, rec_max_radius
        self.cross_max_distance, self.dynamic_max_cross = cross_max_distance, dynamic_max_cross
        self.sh_irreps = o3.Irreps.spherical_harmonics(lmax=sh_lmax)
        self.ns, self.nv = ns, nv
        self.scale_by_sigma, self.no_torsion = scale_by_sigma, no_torsion
        self.smooth_edges, self.odd_parity = smooth_edges, odd_parity
        self.fixed_center_conv, self.no_aminoacid_identities = fixed_center_conv, no_aminoacid_identities
        self.differentiate_convolutions = differentiate_convolutions
        self.embed_also_ligand = embed_also_ligand
        lm_dim = 1280 if lm_embedding_type == 'precomputed' else 0
        S, D, Dx = sigma_embed_dim, distance_embed_dim, cross_distance_embed_dim

        self.lig_node_embedding = AtomEncoder(ns, LIG_FEATURE_DIMS, S)
        self.lig_edge_embedding = _mlp(in_lig_edge_features + S + D, ns, ns, dropout)
        self.rec_node_embedding = AtomEncoder(ns, REC_RESIDUE_FEATURE_DIMS, 0, lm_embedding_dim=lm_dim)
        self.rec_edge_embedding = _mlp(D, ns, ns, dropout)
        self.rec_sigma_embedding = _mlp(S, ns, ns, dropout)
        self.cross_edge_embedding = _mlp(S + Dx, ns, ns, dropout)
        self.lig_distance_expansion = GaussianSmearing(0.0, lig_max_radius, D)
        self.rec_distance_expansion = GaussianSmearing(0.0, rec_max_radius, D)
        self.cross_distance_expansion = GaussianSmearing(0.0, cross_max_distance, Dx)

        seq = get_irrep_seq(ns, nv, use_second_order_repr, reduce_pseudoscalars)
        faster = sh_lmax == 1 and not use_second_order_repr

        def conv(i, groups):
            return TensorProductConvLayer(in_irreps=seq[min(i, len(seq) - 1)], sh_irreps=self.sh_irreps,
                                          out_irreps=seq[min(i + 1, len(seq) - 1)], n_edge_features=3 * ns,
                                          hidden_features=3 * ns, residual=True, batch_norm=batch_norm,
                                          dropout=dropout, faster=faster, tp_weights_layers=tp_weights_layers,
                                          edge_groups=groups)

        self.rec_emb_layers = nn.ModuleList([conv(i, 1) for i in range(num_prot_emb_layers)])
        if embed_also_ligand:
            self.lig_emb_layers = nn.ModuleList([conv(i, 1) for i in range(num_prot_emb_layers)])
        last = num_prot_emb_layers + num_conv_layers - 1
        self.conv_layers = nn.ModuleList([
            conv(i, 1 if not differentiate_convolutions else (2 if i == last else 4))
            for i in range(num_prot_emb_layers, num_prot_emb_layers + num_conv_layers)])

        self.center_distance_expansion = GaussianSmearing(0.0, center_max_distance, D)
        self.center_edge_embedding = _mlp(D + S, ns, ns, dropout)
        self.final_conv = TensorProductConvLayer(in_irreps=self.conv_layers[-1].out_irreps, sh_irreps=self.sh_irreps,
                                                 out_irreps='2x1o + 2x1e' if not odd_parity else '1x1o + 1x1e',
                                                 n_edge_features=2 * ns, residual=False, dropout=dropout,
                                                 batch_norm=batch_norm)
        self.tr_final_layer = nn.Sequential(nn.Linear(1 + S, ns), nn.Dropout(dropout), nn.ReLU(), nn.Linear(ns, 1))
        self.rot_final_layer = nn.Sequential(nn.Linear(1 + S, ns), nn.Dropout(dropout), nn.ReLU(), nn.Linear(ns, 1))
        if not no_torsion:
            self.final_edge_embedding = _mlp(D, ns, ns, dropout)
            self.final_tp_tor = o3.FullTensorProduct(self.sh_irreps, "2e")
            self.tor_bond_conv = TensorProductConvLayer(in_irreps=self.conv_layers[-1].out_irreps,
                                                        sh_irreps=self.final_tp_tor.irreps_out,
                                                        out_irreps=f'{ns}x0o + {ns}x0e' if not odd_parity else f'{ns}x0o',
                                                        n_edge_features=3 * ns, residual=False, dropout=dropout,
                                                        batch_norm=batch_norm)
            self.tor_final_layer = nn.Sequential(nn.Linear(2 * ns if not odd_parity else ns, ns, bias=False), nn.Tanh(),
                                                 nn.Dropout(dropout), nn.Linear(ns, 1, bias=False))

    # ------------------------------------------------------------------------------------------------
    def _dtype(self):
        return self.tr_final_layer[0].weight.dtype

    def _temb(self, t):
        return self.timestep_emb_func(t).to(self._dtype())

    def _sh(self, vec):
        return o3.spherical_harmonics(self.sh_irreps, vec, normalize=True, normalization='component')

    def get_edge_weight(self, edge_vec, max_norm):                      # cg_model.py:459-465
        if self.smooth_edges:
            nn_ = torch.clip(edge_vec.norm(dim=-1) * np.pi / max_norm, max=np.pi)
            return 0.5 * (torch.cos(nn_) + 1.0).unsqueeze(-1)
        return 1.0

    def build_lig_conv_graph(self, data):                               # cg_model.py:467-497
        lig = data['ligand']
        lig.node_sigma_emb = self._temb(lig.node_t['tr'])
        radius_edges = radius_graph(lig.pos, self.lig_max_radius, lig.batch)
        ll = data['ligand', 'ligand']
        edge_index = torch.cat([ll.edge_index, radius_edges], 1).long()
        edge_attr = torch.cat([ll.edge_attr.to(self._dtype()),
                               torch.zeros(radius_edges.shape[-1], self.in_lig_edge_features, dtype=self._dtype(),
                                           device=lig.x.device)], 0)
        edge_attr = torch.cat([edge_attr, lig.node_sigma_emb[edge_index[0]]], 1)
        node_attr = torch.cat([lig.x.to(self._dtype()), lig.node_sigma_emb], 1)
        src, dst = edge_index
        vec = (lig.pos[dst] - lig.pos[src]).to(self._dtype())
        edge_attr = torch.cat([edge_attr, self.lig_distance_expansion(vec.norm(dim=-1))], 1)
        return node_attr, edge_index, edge_attr, self._sh(vec), self.get_edge_weight(vec, self.lig_max_radius)

    def build_rec_conv_graph(self, data):                               # cg_model.py:499-514
        rec = data['receptor']
        src, dst = data['receptor', 'receptor'].edge_index
        vec = (rec.pos[dst.long()] - rec.pos[src.long()]).to(self._dtype())
        return (rec.x.to(self._dtype()), self.rec_distance_expansion(vec.norm(dim=-1)), self._sh(vec),
                self.get_edge_weight(vec, self.rec_max_radius))

    def build_cross_conv_graph(self, data, cutoff):                     # cg_model.py:539-562
        lig, rec = data['ligand'], data['receptor']
        if torch.is_tensor(cutoff):
            edge_index = radius(rec.pos / cutoff[rec.batch], lig.pos / cutoff[lig.batch], 1, rec.batch, lig.batch,
                                max_num_neighbors=10000)
        else:
            edge_index = radius(rec.pos, lig.pos, cutoff, rec.batch, lig.batch, max_num_neighbors=10000)
        src, dst = edge_index
        vec = (rec.pos[dst] - lig.pos[src]).to(self._dtype())
        edge_attr = torch.cat([lig.node_sigma_emb[src], self.cross_distance_expansion(vec.norm(dim=-1))], 1)
        cutoff_d = cutoff[lig.batch[src]].squeeze() if torch.is_tensor(cutoff) else cutoff
        return edge_index, edge_attr, self._sh(vec), self._sh(-vec), self.get_edge_weight(vec, cutoff_d)

    def build_center_conv_graph(self, data):                            # cg_model.py:610-623
        lig = data['ligand']
        edge_index = torch.stack([lig.batch, torch.arange(len(lig.batch), device=lig.x.device)], 0)
        center = torch.zeros((data.num_graphs, 3), dtype=lig.pos.dtype, device=lig.x.device)
        center.index_add_(0, lig.batch, lig.pos)
        center = center / torch.bincount(lig.batch).unsqueeze(1)
        vec = (lig.pos[edge_index[1]] - center[edge_index[0]]).to(self._dtype())
        edge_attr = torch.cat([self.center_distance_expansion(vec.norm(dim=-1)), lig.node_sigma_emb[edge_index[1]]], 1)
        return edge_index, edge_attr, self._sh(vec)

    def build_bond_conv_graph(self, data):                              # cg_model.py:625-639
        lig = data['ligand']
        bonds = data['ligand', 'ligand'].edge_index[:, lig.edge_mask].long()
        bond_pos = (lig.pos[bonds[0]] + lig.pos[bonds[1]]) / 2
        edge_index = radius(lig.pos, bond_pos, self.lig_max_radius, batch_x=lig.batch, batch_y=lig.batch[bonds[0]])
        vec = (lig.pos[edge_index[1]] - bond_pos[edge_index[0]]).to(self._dtype())
        edge_attr = self.final_edge_embedding(self.lig_distance_expansion(vec.norm(dim=-1)))
        return bonds, edge_index, edge_attr, self._sh(vec), self.get_edge_weight(vec, self.lig_max_radius)

    def ligand_embedding(self, data):                                   # cg_model.py:257-270
        node, ei, ea, sh, ew = self.build_lig_conv_graph(data)
        node, ea = self.lig_node_embedding(node), self.lig_edge_embedding(ea)
        assert self.embed_also_ligand, "otherwise reimplement padding"
        for layer in self.lig_emb_layers:
            ea_ = torch.cat([ea, node[ei[0], :self.ns], node[ei[1], :self.ns]], -1)
            node = layer(node, ei, ea_, sh, edge_weight=ew)
        return node, ei, ea, sh, ew

    def embedding(self, data):                                          # cg_model.py:272-306
        rec, rr = data['receptor'], data['receptor', 'receptor']
        if not hasattr(rec, 'rec_node_attr'):
            node, ea, sh, ew = self.build_rec_conv_graph(data)
            node, ea = self.rec_node_embedding(node), self.rec_edge_embedding(ea)
            for layer in self.rec_emb_layers:
                ea_ = torch.cat([ea, node[rr.edge_index[0], :self.ns], node[rr.edge_index[1], :self.ns]], -1)
                node = layer(node, rr.edge_index, ea_, sh, edge_weight=ew)
            rec.rec_node_attr, rr.rec_edge_attr, rr.edge_sh, rr.edge_weight = node, ea, sh, ew
        sig = self.rec_sigma_embedding(self._temb(data.complex_t['tr']))
        rec_node = rec.rec_node_attr + 0
        rec_node[:, :self.ns] = rec_node[:, :self.ns] + sig[rec.batch]
        rec_ea = rr.rec_edge_attr + sig[rec.batch[rr.edge_index[0]]]
        return self.ligand_embedding(data) + (rec_node, rr.edge_index, rec_ea, rr.edge_sh, rr.edge_weight)

    def forward(self, data):                                            # cg_model.py:308-424
        if self.no_aminoacid_identities:
            data['receptor'].x = data['receptor'].x * 0
        tr_sigma, rot_sigma, tor_sigma = self.t_to_sigma(*[data.complex_t[k] for k in ('tr', 'rot', 'tor')])
        (lig_node, lig_ei, lig_ea, lig_sh, lig_ew,
         rec_node, rec_ei, rec_ea, rec_sh, rec_ew) = self.embedding(data)

        cutoff = (tr_sigma * 3 + 20).unsqueeze(1) if self.dynamic_max_cross else self.cross_max_distance
        lr_ei, lr_ea, lr_sh, rev_sh, lr_ew = self.build_cross_conv_graph(data, cutoff)
        lr_ea = self.cross_edge_embedding(lr_ea)

        n_lig = len(lig_node)
        node = torch.cat([lig_node, rec_node], 0)
        lr_ei = torch.stack([lr_ei[0], lr_ei[1] + n_lig], 0)
        ei = torch.cat([lig_ei, lr_ei, rec_ei + n_lig, torch.flip(lr_ei, dims=[0])], 1)
        ea = torch.cat([lig_ea, lr_ea, rec_ea, lr_ea], 0)
        sh = torch.cat([lig_sh, lr_sh, rec_sh, rev_sh], 0)
        ew = (torch.cat([lig_ew, lr_ew, rec_ew, lr_ew], 0) if torch.is_tensor(lig_ew)
              else torch.ones((ei.shape[1], 1), dtype=node.dtype, device=ei.device))
        s1 = lig_ei.shape[1]
        s2 = s1 + lr_ei.shape[1]
        s3 = s2 + rec_ei.shape[1]

        L = len(self.conv_layers)
        for l, layer in enumerate(self.conv_layers):
            if l < L - 1:
                ea_ = torch.cat([ea, node[ei[0], :self.ns], node[ei[1], :self.ns]], -1)
                if self.differentiate_convolutions:
                    ea_ = [ea_[:s1], ea_[s1:s2], ea_[s2:s3], ea_[s3:]]
                node = layer(node, ei, ea_, sh, edge_weight=ew)
            else:   # last layer: only edges whose target is a ligand atom
                ea_ = torch.cat([ea[:s2], node[ei[0, :s2], :self.ns], node[ei[1, :s2], :self.ns]], -1)
                if self.differentiate_convolutions:
                    ea_ = [ea_[:s1], ea_[s1:s2]]
                node = layer(node, ei[:, :s2], ea_, sh[:s2], edge_weight=ew[:s2])
        lig_node = node[:n_lig]
        return self._heads(data, lig_node, tr_sigma, rot_sigma, tor_sigma)

    def _heads(self, data, lig_node, tr_sigma, rot_sigma, tor_sigma):   # cg_model.py:368-424 (shared with aa_model.py:443-508)
        # translation / rotation head
        c_ei, c_ea, c_sh = self.build_center_conv_graph(data)
        c_ea = self.center_edge_embedding(c_ea)
        idx = c_ei[1] if self.fixed_center_conv else c_ei[0]
        c_ea = torch.cat([c_ea, lig_node[idx, :self.ns]], -1)
        g = self.final_conv(lig_node, c_ei, c_ea, c_sh, out_nodes=data.num_graphs)
        tr = g[:, :3] + (g[:, 6:9] if not self.odd_parity else 0)
        rot = g[:, 3:6] + (g[:, 9:] if not self.odd_parity else 0)
        data.graph_sigma_emb = self._temb(data.complex_t['tr'])
        tr_norm = torch.linalg.vector_norm(tr, dim=1).unsqueeze(1)
        tr = tr / tr_norm * self.tr_final_layer(torch.cat([tr_norm, data.graph_sigma_emb], 1))
        rot_norm = torch.linalg.vector_norm(rot, dim=1).unsqueeze(1)
        rot = rot / rot_norm * self.rot_final_layer(torch.cat([rot_norm, data.graph_sigma_emb], 1))
        if self.scale_by_sigma:
            tr = tr / tr_sigma.unsqueeze(1).to(tr.dtype)
            rot = rot * so3_score_norm(rot_sigma.cpu()).unsqueeze(1).to(rot.device).to(rot.dtype)

        lig = data['ligand']
        if self.no_torsion or lig.edge_mask.sum() == 0:
            return tr, rot, torch.empty(0, device=tr.device), None

        # torsion head
        bonds, t_ei, t_ea, t_sh, t_ew = self.build_bond_conv_graph(data)
        bond_vec = (lig.pos[bonds[1]] - lig.pos[bonds[0]]).to(self._dtype())
        bond_attr = lig_node[bonds[0]] + lig_node[bonds[1]]
        bonds_sh = o3.spherical_harmonics("2e", bond_vec, normalize=True, normalization='component')
        t_sh = self.final_tp_tor(t_sh, bonds_sh[t_ei[0]])
        t_ea = torch.cat([t_ea, lig_node[t_ei[1], :self.ns], bond_attr[t_ei[0], :self.ns]], -1)
        tor = self.tor_bond_conv(lig_node, t_ei, t_ea, t_sh, out_nodes=int(lig.edge_mask.sum()), reduce='mean',
                                 edge_weight=t_ew)
        tor = self.tor_final_layer(tor).squeeze(1)
        edge_sigma = tor_sigma[lig.batch][data['ligand', 'ligand'].edge_index[0]][lig.edge_mask]
        if self.scale_by_sigma:
            tor = tor * torch.sqrt(torch.as_tensor(torus_score_norm(edge_sigma.cpu().numpy())).float()
                                   .to(tor.device)).to(tor.dtype)
        return tr, rot, tor, None
