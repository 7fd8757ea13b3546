"""Seeded synthetic protein-ligand complexes with the input layout of the reference's featuriser
(datasets/process_mols.py:59-87,161-241,279-301; utils/torsion.py:15-45), as specified in
SURVEY.md section 8(d).  Host-side (numpy) workload generator for tests and bench - there is no
dataset or RDKit/ProDy/ESM in this image.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch

from .hetero import HeteroGraph

# vocabulary sizes: datasets/process_mols.py:59-76 (ligand atoms), :85-87 (residues)
LIG_FEATURE_DIMS = ([119, 4, 12, 12, 8, 10, 6, 6, 2, 8, 2, 2, 2, 2, 2, 2], 0)
REC_RESIDUE_FEATURE_DIMS = ([38], 0)
REC_ATOM_FEATURE_DIMS = ([38, 119, 23, 38], 0)   # datasets/process_mols.py:78-83 (residue type, atomic number, two atom-name classes)
LM_EMBED_DIM = 1280  # ESM2-650M, models/cg_model.py:73-74


def default_model_args(**over):
    """'DiffDock-L-shaped' hyper-parameters (CFG-L2 of SURVEY.md section 8); the released
    model_parameters.yml is not in the reference tree."""
    a = dict(ns=48, nv=10, sh_lmax=2, num_conv_layers=6, num_prot_emb_layers=0, max_radius=5.0,
             rec_max_radius=30.0, cross_max_distance=80.0, center_max_distance=30.0, distance_embed_dim=64,
             cross_distance_embed_dim=64, sigma_embed_dim=64, embedding_type='sinusoidal', embedding_scale=1000,
             dynamic_max_cross=True, no_torsion=False, scale_by_sigma=True, use_second_order_repr=False,
             no_batch_norm=False, dropout=0.0, c_alpha_max_neighbors=24, receptor_radius=15.0,
             tr_sigma_min=0.1, tr_sigma_max=19.0, rot_sigma_min=0.03, rot_sigma_max=1.55,
             tor_sigma_min=0.0314, tor_sigma_max=3.14, all_atoms=False, crop_beyond=None,
             differentiate_convolutions=True, tp_weights_layers=2, reduce_pseudoscalars=False,
             embed_also_ligand=True, smooth_edges=False, odd_parity=False, fixed_center_conv=False)
    a.update(over)
    return SimpleNamespace(**a)


def _receptor(n_res, rng, max_neighbors, radius):
    R = (3.0 * n_res / (4.0 * math.pi * 0.0075)) ** (1.0 / 3.0)
    v = rng.normal(size=(n_res, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    pos = v * (R * rng.uniform(size=(n_res, 1)) ** (1.0 / 3.0))
    pos -= pos.mean(0, keepdims=True)
    src, dst = [], []
    for s in range(0, n_res, 512):  # blocked kNN, O(N^2) but small
        d = np.linalg.norm(pos[s:s + 512, None, :] - pos[None, :, :], axis=-1)
        for i in range(d.shape[0]):
            d[i, s + i] = np.inf
        k = min(max_neighbors, n_res - 1)
        nb = np.argpartition(d, k - 1, axis=1)[:, :k]
        for i in range(d.shape[0]):
            row = nb[i][d[i, nb[i]] < radius]
            row = row[np.argsort(d[i, row], kind='stable')]
            src.extend(row.tolist())            # neighbour
            dst.extend([s + i] * len(row))      # centre   (process_mols.py:192 stores [neighbour, centre])
    edge_index = np.stack([np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64)], 0)
    x = np.concatenate([rng.integers(0, 38, size=(n_res, 1)).astype(np.float32),
                        rng.normal(size=(n_res, LM_EMBED_DIM)).astype(np.float32)], 1)
    return pos.astype(np.float32), edge_index, x, R


def _ligand(n_atoms, rng):
    pos = np.zeros((n_atoms, 3))
    parent = np.zeros(n_atoms, dtype=np.int64)
    for i in range(1, n_atoms):
        for _ in range(200):
            p = i - 1 if rng.uniform() < 0.8 else int(rng.integers(0, i))
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            cand = pos[p] + 1.5 * d
            if np.all(np.linalg.norm(pos[:i] - cand, axis=1) >= 1.2):
                break
        pos[i], parent[i] = cand, p
    # bonds stored as consecutive (u->v, v->u) pairs like the reference's RDKit loop (process_mols.py:284-292)
    ei = []
    for i in range(1, n_atoms):
        ei += [(parent[i], i), (i, parent[i])]
    edge_index = np.asarray(ei, dtype=np.int64).T.reshape(2, -1)
    btype = rng.integers(0, 4, size=n_atoms - 1)
    edge_attr = np.zeros((2 * (n_atoms - 1), 4), dtype=np.float32)
    edge_attr[np.arange(0, 2 * (n_atoms - 1), 2), btype] = 1
    edge_attr[np.arange(1, 2 * (n_atoms - 1), 2), btype] = 1
    x = np.stack([rng.integers(0, d, size=n_atoms) for d in LIG_FEATURE_DIMS[0]], 1).astype(np.int64)
    # rotatable-bond masks, semantics of utils/torsion.py:15-45 on a tree: removing a bond always disconnects;
    # keep it iff the smaller side has >= 2 atoms; mark the directed edge u->v whose head v lies in that side.
    adj = [[] for _ in range(n_atoms)]
    for i in range(1, n_atoms):
        adj[i].append(int(parent[i]))
        adj[int(parent[i])].append(i)
    edge_mask = np.zeros(edge_index.shape[1], dtype=bool)
    rows = []
    for e in range(0, edge_index.shape[1], 2):
        a, b = int(edge_index[0, e]), int(edge_index[1, e])
        seen = {a}
        stack = [a]
        while stack:  # component of a without using bond (a,b)
            n = stack.pop()
            for m in adj[n]:
                if (n == a and m == b) or (n == b and m == a) or m in seen:
                    continue
                seen.add(m)
                stack.append(m)
        side_a = seen
        side_b = set(range(n_atoms)) - side_a
        small = side_a if len(side_a) <= len(side_b) else side_b
        if len(small) > 1:
            m = np.zeros(n_atoms, dtype=bool)
            m[list(small)] = True
            if a in small:      # edge e is a->b with a in the rotating side: mark the reverse edge b->a
                edge_mask[e + 1] = True
            else:
                edge_mask[e] = True
            rows.append((e + 1 if a in small else e, m))
    rows.sort(key=lambda t: t[0])
    mask_rotate = np.stack([m for _, m in rows], 0) if rows else np.zeros((0, n_atoms), dtype=bool)
    return pos.astype(np.float32), edge_index, edge_attr, x, edge_mask, mask_rotate


def _random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    r, i, j, k = q
    return np.array([[1 - 2 * (j * j + k * k), 2 * (i * j - k * r), 2 * (i * k + j * r)],
                     [2 * (i * j + k * r), 1 - 2 * (i * i + k * k), 2 * (j * k - i * r)],
                     [2 * (i * k - j * r), 2 * (j * k + i * r), 1 - 2 * (i * i + j * j)]])


def _receptor_atoms(rpos, res_type, rng, atom_cutoff, atom_max_neighbors):
    """All-atom receptor nodes in the layout of datasets/process_mols.py:203-239: 3-7 atoms scattered ~1.5 A around every
    residue position; ('atom','atom') contact edges stored as [neighbour, centre] (all atoms within ``atom_cutoff``, the
    nearest ``atom_max_neighbors`` if there are more, the nearest one if there is none); ('atom','receptor') edges
    [atom, its residue]; features = (residue type, atomic number, atom-name class 2, atom-name class 3)."""
    n_res = rpos.shape[0]
    counts = rng.integers(3, 8, size=n_res)
    res_of = np.repeat(np.arange(n_res), counts)
    apos = (rpos[res_of] + rng.normal(scale=1.5, size=(len(res_of), 3))).astype(np.float32)
    dims = REC_ATOM_FEATURE_DIMS[0]
    ax = np.stack([res_type[res_of].astype(np.int64)] + [rng.integers(0, d, size=len(res_of)) for d in dims[1:]], 1).astype(np.float32)
    src, dst = [], []
    for s in range(0, len(apos), 512):
        d = np.linalg.norm(apos[s:s + 512, None, :] - apos[None, :, :], axis=-1)
        for i in range(d.shape[0]):
            row = d[i].copy()
            row[s + i] = np.inf
            nb = np.where(row < atom_cutoff)[0]
            if len(nb) > atom_max_neighbors:
                nb = np.argsort(row, kind='stable')[:atom_max_neighbors]
            if len(nb) == 0:
                nb = np.argsort(row, kind='stable')[:1]
            src.extend(nb.tolist())
            dst.extend([s + i] * len(nb))
    aa_edge = np.stack([np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64)], 0)
    ar_edge = np.stack([np.arange(len(res_of), dtype=np.int64), res_of.astype(np.int64)], 0)
    return apos, ax, aa_edge, ar_edge


def make_complex(n_res=200, n_atoms=20, seed=0, max_neighbors=24, rec_radius=15.0, name=None,
                 lm_dim=LM_EMBED_DIM, all_atoms=False, atom_cutoff=5.0, atom_max_neighbors=8) -> HeteroGraph:
    """One synthetic complex (CPU tensors).  Ligand is placed half-way between receptor centre and surface.
    ``all_atoms``: also the receptor-atom nodes and edges the all-atom models read (models/old_aa_model.py:424-486)."""
    rng = np.random.default_rng(seed)
    rpos, redge, rx, R = _receptor(n_res, rng, max_neighbors, rec_radius)
    if lm_dim != LM_EMBED_DIM:
        rx = rx[:, :1 + lm_dim]
    lpos, ledge, lattr, lx, emask, mrot = _ligand(n_atoms, rng)
    d = rng.normal(size=3)
    d /= np.linalg.norm(d)
    lpos = (lpos - lpos.mean(0)) @ _random_rotation(rng).T + 0.5 * R * d
    g = HeteroGraph()
    g['name'] = name or f"synth_r{n_res}_l{n_atoms}_s{seed}"
    lig = g['ligand']
    lig.x = torch.from_numpy(lx)
    lig.pos = torch.from_numpy(lpos.astype(np.float32))
    lig.edge_mask = torch.from_numpy(emask)
    lig.mask_rotate = [mrot]
    ll = g['ligand', 'ligand']
    ll.edge_index = torch.from_numpy(ledge)
    ll.edge_attr = torch.from_numpy(lattr)
    rec = g['receptor']
    rec.x = torch.from_numpy(rx)
    rec.pos = torch.from_numpy(rpos)
    # carried (and cropped) by utils/utils.py:crop_beyond; not read by the coarse-grained score model
    rec.side_chain_vecs = torch.from_numpy(np.random.default_rng(seed + 7919).normal(size=(n_res, 4, 3)).astype(np.float32))
    g['receptor', 'receptor'].edge_index = torch.from_numpy(redge)
    if all_atoms:
        apos, ax, aa_edge, ar_edge = _receptor_atoms(rpos, rx[:, 0], np.random.default_rng(seed + 104729), atom_cutoff,
                                                     atom_max_neighbors)
        g['atom'].x = torch.from_numpy(ax)
        g['atom'].pos = torch.from_numpy(apos)
        g['atom', 'atom'].edge_index = torch.from_numpy(aa_edge)
        g['atom', 'receptor'].edge_index = torch.from_numpy(ar_edge)
    return g


def randomize_pose(g: HeteroGraph, seed, tr_sigma_max, no_torsion=False) -> HeteroGraph:
    """Prior sample in the style of utils/sampling.py:16-58 (uniform torsions, uniform rotation about the
    ligand centroid, placed at the receptor centre + N(0, tr_sigma_max)); numpy RNG, seeded."""
    rng = np.random.default_rng(seed)
    g = g.clone()
    pos = g['ligand'].pos.double().numpy().copy()
    if not no_torsion:
        ei = g['ligand', 'ligand'].edge_index.numpy().T[g['ligand'].edge_mask.numpy()]
        mrot = g['ligand'].mask_rotate[0]
        for r, (u, v) in enumerate(ei):
            ang = rng.uniform(-math.pi, math.pi)
            ax = pos[u] - pos[v]
            ax /= np.linalg.norm(ax)
            K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
            Rm = np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)
            pos[mrot[r]] = (pos[mrot[r]] - pos[v]) @ Rm.T + pos[v]
    centre = g['receptor'].pos.double().numpy().mean(0)
    pos = (pos - pos.mean(0)) @ _random_rotation(rng).T + centre + rng.normal(scale=tr_sigma_max, size=(1, 3))
    g['ligand'].pos = torch.from_numpy(pos.astype(np.float32))
    return g


def make_pose_list(n_poses, n_res=200, n_atoms=20, seed=0, tr_sigma_max=19.0, share_receptor=False, **kw):
    """``n_poses`` prior samples of one synthetic complex.  Default: every pose is a deep copy of the complex, like
    inference.py:236-239; ``share_receptor=True`` lets the poses share the receptor tensors (same storage), which keeps a
    many-complex workload (BASELINE config 5) small on the host."""
    base = make_complex(n_res, n_atoms, seed, **kw)
    poses = [randomize_pose(base, seed * 100003 + 17 * p + 1, tr_sigma_max) for p in range(n_poses)]
    if share_receptor:
        for g in poses[1:]:
            g._nodes['receptor'] = poses[0]._nodes['receptor']
            g._edges[('receptor', 'receptor')] = poses[0]._edges[('receptor', 'receptor')]
    return poses


def config5_sizes(n_complexes=64, seed=0):
    """BASELINE config 5 / SURVEY 8(d): N_r ~ U(200, 600), N_l ~ U(15, 50) per complex (seeded)."""
    rng = np.random.default_rng(seed)
    return [(int(rng.integers(200, 601)), int(rng.integers(15, 51))) for _ in range(n_complexes)]
