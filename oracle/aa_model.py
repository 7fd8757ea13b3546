"""Oracle restatement of models/aa_model.py (AAModel, the all-atom score model).  TEST INFRASTRUCTURE.

Same constructor keywords and state_dict keys as the reference class (models/aa_model.py:22-273) for the supported subset:
score mode, lm_embedding_type in {None, 'precomputed'}, one noise schedule, parallel=1.  Compared with the coarse-grained
model (oracle/cg_model.py) there is a third node type (receptor atoms) and the joint graph has nine edge groups
(aa_model.py:401-417) - ligand<-ligand, ligand<-residue, ligand<-atom, residue<-residue, residue<-ligand, residue<-atom,
atom<-atom, atom<-ligand, atom<-residue - three in the last layer (:429-430).  Unlike the coarse-grained model, the reversed
groups reuse the forward direction's spherical harmonics (:405-406) and the ligand-atom distances go through the LIGAND
distance expansion (:613) into an MLP sized for the cross expansion (:108) - both reproduced.  The receptor part (residue and
atom embeddings, their edge attributes, optional protein-embedding layers over four groups, :275-325) is cached on ``data``."""
import numpy as np
import torch
from torch import nn

from . import e3nn_lite as o3
from .cg_model import CGModel, LIG_FEATURE_DIMS, REC_RESIDUE_FEATURE_DIMS, _mlp
from .graph_ops import radius
from .layers import AtomEncoder
from .tensor_layers import TensorProductConvLayer, get_irrep_seq

REC_ATOM_FEATURE_DIMS = ([38, 119, 23, 38], 0)       # datasets/process_mols.py:78-83


class AAModel(CGModel):
    def __init__(self, t_to_sigma, device, timestep_emb_func, in_lig_edge_features=4, sigma_embed_dim=32, sh_lmax=2,
                 ns=16, nv=4, num_conv_layers=2, lig_max_radius=5, rec_max_radius=30, cross_max_distance=250,
                 center_max_distance=30, distance_embed_dim=32, cross_distance_embed_dim=32, no_torsion=False,
                 scale_by_sigma=True, norm_by_sigma=True, use_second_order_repr=False, batch_norm=True,
                 dynamic_max_cross=False, dropout=0.0, smooth_edges=False, odd_parity=False,
                 separate_noise_schedule=False, lm_embedding_type=None, confidence_mode=False,
                 differentiate_convolutions=True, tp_weights_layers=2, num_prot_emb_layers=0,
                 reduce_pseudoscalars=False, embed_also_ligand=False, fixed_center_conv=False,
                 no_aminoacid_identities=False, **unused):
        super().__init__(t_to_sigma, device, timestep_emb_func, in_lig_edge_features=in_lig_edge_features,
                         sigma_embed_dim=sigma_embed_dim, sh_lmax=sh_lmax, ns=ns, nv=nv, num_conv_layers=num_conv_layers,
                         lig_max_radius=lig_max_radius, rec_max_radius=rec_max_radius,
                         cross_max_distance=cross_max_distance, center_max_distance=center_max_distance,
                         distance_embed_dim=distance_embed_dim, cross_distance_embed_dim=cross_distance_embed_dim,
                         no_torsion=no_torsion, scale_by_sigma=scale_by_sigma, use_second_order_repr=use_second_order_repr,
                         batch_norm=batch_norm, dynamic_max_cross=dynamic_max_cross, dropout=dropout,
                         smooth_edges=smooth_edges, odd_parity=odd_parity, lm_embedding_type=lm_embedding_type,
                         differentiate_convolutions=differentiate_convolutions, tp_weights_layers=tp_weights_layers,
                         num_prot_emb_layers=num_prot_emb_layers, reduce_pseudoscalars=reduce_pseudoscalars,
                         embed_also_ligand=embed_also_ligand, fixed_center_conv=fixed_center_conv,
                         no_aminoacid_identities=no_aminoacid_identities)
        S, D, Dx = sigma_embed_dim, distance_embed_dim, cross_distance_embed_dim
        del self.cross_edge_embedding
        self.atom_node_embedding = AtomEncoder(ns, REC_ATOM_FEATURE_DIMS, 0)
        self.atom_edge_embedding = _mlp(D, ns, ns, dropout)
        self.lr_edge_embedding = _mlp(S + Dx, ns, ns, dropout)
        self.ar_edge_embedding = _mlp(D, ns, ns, dropout)
        self.la_edge_embedding = _mlp(S + Dx, ns, ns, dropout)
        seq = get_irrep_seq(ns, nv, use_second_order_repr, reduce_pseudoscalars)
        faster = sh_lmax == 1 and not use_second_order_repr

        def conv(i, groups):
            return TensorProductConvLayer(in_irreps=seq[min(i, len(seq) - 1)], sh_irreps=self.sh_irreps,
                                          out_irreps=seq[min(i + 1, len(seq) - 1)], n_edge_features=3 * ns,
                                          hidden_features=3 * ns, residual=True, batch_norm=batch_norm, dropout=dropout,
                                          faster=faster, tp_weights_layers=tp_weights_layers, edge_groups=groups)

        d = differentiate_convolutions
        self.rec_emb_layers = nn.ModuleList([conv(i, 4 if d else 1) for i in range(num_prot_emb_layers)])
        last = num_prot_emb_layers + num_conv_layers - 1
        self.conv_layers = nn.ModuleList([conv(i, 1 if not d else (3 if i == last else 9))
                                          for i in range(num_prot_emb_layers, num_prot_emb_layers + num_conv_layers)])

    # ---------------------------------------------------------------------------------------------------------
    def _plain_graph(self, data, nt, expansion, max_r):                 # aa_model.py:558-586
        st = data[nt]
        ei = data[nt, nt].edge_index.long()
        vec = (st.pos[ei[1]] - st.pos[ei[0]]).to(self._dtype())
        return st.x, expansion(vec.norm(dim=-1)), self._sh(vec), self.get_edge_weight(vec, max_r)

    def build_cross_lig_conv_graph(self, data, cutoff):                 # :588-623
        lig, rec, atom = data['ligand'], data['receptor'], data['atom']
        if torch.is_tensor(cutoff):
            lr = radius(rec.pos / cutoff[rec.batch], lig.pos / cutoff[lig.batch], 1, rec.batch, lig.batch,
                        max_num_neighbors=10000)
        else:
            lr = radius(rec.pos, lig.pos, cutoff, rec.batch, lig.batch, max_num_neighbors=10000)
        lr_vec = (rec.pos[lr[1]] - lig.pos[lr[0]]).to(self._dtype())
        lr_attr = torch.cat([lig.node_sigma_emb[lr[0]], self.cross_distance_expansion(lr_vec.norm(dim=-1))], 1)
        cutoff_d = cutoff[lig.batch[lr[0]]].squeeze() if torch.is_tensor(cutoff) else cutoff
        lr_w = self.get_edge_weight(lr_vec, cutoff_d)
        la = radius(atom.pos, lig.pos, self.lig_max_radius, atom.batch, lig.batch, max_num_neighbors=10000)
        la_vec = (atom.pos[la[1]] - lig.pos[la[0]]).to(self._dtype())
        la_attr = torch.cat([lig.node_sigma_emb[la[0]], self.lig_distance_expansion(la_vec.norm(dim=-1))], 1)     # :613
        return lr, lr_attr, self._sh(lr_vec), lr_w, la, la_attr, self._sh(la_vec), self.get_edge_weight(la_vec, self.lig_max_radius)

    def embedding(self, data):                                          # :275-362
        rec, atom = data['receptor'], data['atom']
        rr, aa, ar = data['receptor', 'receptor'], data['atom', 'atom'], data['atom', 'receptor']
        ns = self.ns
        if not hasattr(rec, 'rec_node_attr'):
            r_node, r_ea, r_sh, r_w = self._plain_graph(data, 'receptor', self.rec_distance_expansion, self.rec_max_radius)
            r_node, r_ea = self.rec_node_embedding(r_node), self.rec_edge_embedding(r_ea)
            a_node, a_ea, a_sh, a_w = self._plain_graph(data, 'atom', self.lig_distance_expansion, self.lig_max_radius)
            a_node, a_ea = self.atom_node_embedding(a_node), self.atom_edge_embedding(a_ea)
            ar_ei0 = ar.edge_index.long()
            ar_vec = (rec.pos[ar_ei0[1]] - atom.pos[ar_ei0[0]]).to(self._dtype())
            ar_ea, ar_sh = self.ar_edge_embedding(self.rec_distance_expansion(ar_vec.norm(dim=-1))), self._sh(ar_vec)
            n_rec = len(r_node)
            node = torch.cat([r_node, a_node], 0)
            ar_ei = torch.stack([ar_ei0[0] + n_rec, ar_ei0[1]], 0)
            ei = torch.cat([rr.edge_index.long(), ar_ei, aa.edge_index.long() + n_rec, torch.flip(ar_ei, dims=[0])], 1)
            ea = torch.cat([r_ea, ar_ea, a_ea, ar_ea], 0)
            sh = torch.cat([r_sh, ar_sh, a_sh, ar_sh], 0)
            ew = (torch.cat([r_w, torch.ones_like(r_w[:1]).expand(ar_ei.shape[1], 1), a_w,
                             torch.ones_like(r_w[:1]).expand(ar_ei.shape[1], 1)], 0)
                  if torch.is_tensor(r_w) else torch.ones((ei.shape[1], 1), dtype=node.dtype))
            s1 = rr.edge_index.shape[1]
            s2 = s1 + ar_ei.shape[1]
            s3 = s2 + aa.edge_index.shape[1]
            for layer in self.rec_emb_layers:
                ea_ = torch.cat([ea, node[ei[0], :ns], node[ei[1], :ns]], -1)
                if self.differentiate_convolutions:
                    ea_ = [ea_[:s1], ea_[s1:s2], ea_[s2:s3], ea_[s3:]]
                node = layer(node, ei, ea_, sh, edge_weight=ew)
            rec.rec_node_attr, rr.rec_edge_attr, rr.edge_sh, rr.edge_weight = node[:n_rec], r_ea, r_sh, r_w
            atom.atom_node_attr, aa.atom_edge_attr, aa.edge_sh, aa.edge_weight = node[n_rec:], a_ea, a_sh, a_w
            ar.edge_attr, ar.edge_sh, ar.edge_weight = ar_ea, ar_sh, 1
        sig = self.rec_sigma_embedding(self._temb(data.complex_t['tr']))
        rec_node = rec.rec_node_attr + 0
        rec_node[:, :ns] = rec_node[:, :ns] + sig[rec.batch]
        rec_ea = rr.rec_edge_attr + sig[rec.batch[rr.edge_index[0]]]
        atom_node = atom.atom_node_attr + 0
        atom_node[:, :ns] = atom_node[:, :ns] + sig[atom.batch]
        atom_ea = aa.atom_edge_attr + sig[atom.batch[aa.edge_index[0]]]
        ar_ea = ar.edge_attr + sig[atom.batch[ar.edge_index[0]]]
        return self.ligand_embedding(data) + (rec_node, rr.edge_index.long(), rec_ea, rr.edge_sh, rr.edge_weight,
                                              atom_node, aa.edge_index.long(), atom_ea, aa.edge_sh, aa.edge_weight,
                                              ar.edge_index.long(), ar_ea, ar.edge_sh, ar.edge_weight)

    def forward(self, data):                                            # :364-508
        if self.no_aminoacid_identities:
            data['receptor'].x = data['receptor'].x * 0
        tr_sigma, rot_sigma, tor_sigma = self.t_to_sigma(*[data.complex_t[k] for k in ('tr', 'rot', 'tor')])
        (lig, lig_ei, lig_ea, lig_sh, lig_w, rec, rec_ei, rec_ea, rec_sh, rec_w,
         atom, at_ei, at_ea, at_sh, at_w, ar_ei, ar_ea, ar_sh, ar_w) = self.embedding(data)
        cutoff = (tr_sigma * 3 + 20).unsqueeze(1) if self.dynamic_max_cross else self.cross_max_distance
        lr_ei, lr_ea, lr_sh, lr_w, la_ei, la_ea, la_sh, la_w = self.build_cross_lig_conv_graph(data, cutoff)
        lr_ea, la_ea = self.lr_edge_embedding(lr_ea), self.la_edge_embedding(la_ea)
        n_lig, n_rec, ns = len(lig), len(rec), self.ns
        node = torch.cat([lig, rec, atom], 0)
        rec_ei = rec_ei + n_lig
        at_ei = at_ei + n_lig + n_rec
        lr_ei = torch.stack([lr_ei[0], lr_ei[1] + n_lig], 0)
        la_ei = torch.stack([la_ei[0], la_ei[1] + n_lig + n_rec], 0)
        ar_ei = torch.stack([ar_ei[0] + n_lig + n_rec, ar_ei[1] + n_lig], 0)
        flip = lambda e: torch.flip(e, dims=[0])
        parts_ei = [lig_ei, lr_ei, la_ei, rec_ei, flip(lr_ei), flip(ar_ei), at_ei, flip(la_ei), ar_ei]
        parts_ea = [lig_ea, lr_ea, la_ea, rec_ea, lr_ea, ar_ea, at_ea, la_ea, ar_ea]
        parts_sh = [lig_sh, lr_sh, la_sh, rec_sh, lr_sh, ar_sh, at_sh, la_sh, ar_sh]
        ei, ea, sh = torch.cat(parts_ei, 1), torch.cat(parts_ea, 0), torch.cat(parts_sh, 0)
        if torch.is_tensor(lig_w):
            one = lambda n: torch.ones((n, 1), dtype=node.dtype)
            arw = one(ar_ei.shape[1])
            ew = torch.cat([lig_w, lr_w, la_w, rec_w, lr_w, arw, at_w, la_w, arw], 0)
        else:
            ew = torch.ones((ei.shape[1], 1), dtype=node.dtype)
        cuts = np.cumsum([p.shape[0] for p in parts_ea]).tolist()
        L = len(self.conv_layers)
        for l, layer in enumerate(self.conv_layers):
            if l < L - 1:
                ea_ = torch.cat([ea, node[ei[0], :ns], node[ei[1], :ns]], -1)
                if self.differentiate_convolutions:
                    ea_ = [ea_[a:b] for a, b in zip([0] + cuts[:-1], cuts)]
                node = layer(node, ei, ea_, sh, edge_weight=ew)
            else:       # last layer: only the three groups that end on ligand atoms
                s3 = cuts[2]
                ea_ = torch.cat([ea[:s3], node[ei[0, :s3], :ns], node[ei[1, :s3], :ns]], -1)
                if self.differentiate_convolutions:
                    ea_ = [ea_[:cuts[0]], ea_[cuts[0]:cuts[1]], ea_[cuts[1]:s3]]
                node = layer(node, ei[:, :s3], ea_, sh[:s3], edge_weight=ew[:s3])
        return self._heads(data, node[:n_lig], tr_sigma, rot_sigma, tor_sigma)
