"""Oracle restatement of the INPUT SIDE of the hot path: complex graph construction from
parsed arrays.  TEST INFRASTRUCTURE (see oracle/__init__.py) - numpy / plain loops, each function citing the reference
lines it follows (gcorso/DiffDock @ b4704d9).

Pinned by tests/golden/ref_inputs.pt, which tests/golden/make_golden_inputs.py produced by running the UNMODIFIED
``new_extract_receptor_structure`` / ``get_lig_graph`` (datasets/process_mols.py), ``get_transformation_mask``
(utils/torsion.py) and the ESM chain re-ordering of datasets/pdbbind.py on synthetic arrays (parsers - RDKit, ProDy,
Biopython - replaced by array inputs; they are outside section 8).

Not restated: ``side_chain_vecs`` (chi angles, process_mols.py:163-165) - read only by the side-chain head, which is out of
scope - and the all-atom branch (process_mols.py:203-239): in this tree ``get_moad_atom_feats`` (:244-247) returns an empty
array, so the reference's own inference-time all-atom featurisation yields no atoms.
"""
import numpy as np

# index of the residue's three-letter name in allowable_features['possible_amino_acids'] (process_mols.py:47-49) via
# aa_short2long (datasets/constants.py:37-40); anything else maps to the last entry 'misc' (safe_index, :120-125)
RESIDUE_ORDER = 'ARNDCQEGHILKMFPSTWYV'
MISC_RESIDUE = 37
N_BOND_TYPES = 4          # process_mols.py:57: SINGLE, DOUBLE, TRIPLE, AROMATIC


def _fma32(a, b, c):
    """fp32 fused multiply-add, emulated in extended precision: the product of two fp32 numbers is exact in 48 bits and the
    sum with a third fits numpy's longdouble (64-bit significand on x86) in all but astronomically rare alignments, so the
    single rounding to fp32 is the FMA's (plain fp64 would round twice)."""
    wide = np.longdouble if np.finfo(np.longdouble).nmant >= 63 else np.float64
    return (a.astype(wide) * b.astype(wide) + c.astype(wide)).astype(np.float32)


def cdist_sq_f32(x):
    """Squared distances as ``torch.cdist(x, x)`` forms them for fp32 ``x`` [N, 3] (process_mols.py:176 calls it on the
    C-alpha coordinates).  More than 25 points -> ATen's ``_euclidean_dist``: ``[-2 x_i, |x_i|^2, 1] . [x_j, 1, |x_j|^2]``
    by an sgemm with K = 5 - one FMA chain in k order - then clamp_min(0); otherwise the direct form sum (a - b)^2.
    tests/test_inputs_cpu.py checks both forms bit for bit against the installed torch."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = x.shape[0]
    if n > 25:
        nrm = ((x[:, 0] * x[:, 0] + x[:, 1] * x[:, 1]) + x[:, 2] * x[:, 2]).astype(np.float32)
        a = np.concatenate([-2.0 * x, nrm[:, None], np.ones((n, 1), np.float32)], 1).astype(np.float32)
        b = np.concatenate([x, np.ones((n, 1), np.float32), nrm[:, None]], 1).astype(np.float32)
        acc = np.zeros((n, n), np.float32)
        for k in range(5):
            acc = _fma32(a[:, k:k + 1], b[None, :, k], acc)
        return np.maximum(acc, np.float32(0))
    d = x[:, None, :] - x[None, :, :]
    acc = d[..., 0] * d[..., 0]
    acc = _fma32(d[..., 1], d[..., 1], acc)
    return _fma32(d[..., 2], d[..., 2], acc)


def cdist_f32(x):
    """``torch.cdist(x, x)``: the correctly rounded (IEEE) fp32 square root of ``cdist_sq_f32``.  torch's CPU kernels take
    the root with a vectorised routine that is 1 ulp off for ~0.7 % of the arguments (measured: torch 2.11, AVX-512 build) -
    a property of the host build, not of the algorithm, so neither the oracle nor the CUDA kernel copies it; it can only
    move a pair whose distance lies within one ulp of the cut-off."""
    return np.sqrt(cdist_sq_f32(x))


def contact_graph(coords, cutoff, max_neighbors=None):
    """process_mols.py:176-192 (and :206-224 for atoms): per centre the neighbours within ``cutoff`` (index order), the
    ``max_neighbors`` nearest if there are more, the nearest other point if there are none -> edge_index [2, E] int64,
    rows [neighbour, centre]."""
    dist = cdist_f32(coords)
    cut = np.float32(cutoff)
    k = max_neighbors if max_neighbors else 1000
    nbr, ctr = [], []
    for i in range(dist.shape[0]):
        dst = [j for j in np.where(dist[i] < cut)[0].tolist() if j != i]
        if len(dst) > k:
            order = sorted(range(dist.shape[0]), key=lambda j: (dist[i, j], j))
            order.remove(i)
            dst = order[:k]
        if len(dst) == 0 and dist.shape[0] > 1:
            order = sorted(range(dist.shape[0]), key=lambda j: (dist[i, j], j))
            order.remove(i)
            dst = order[:1]
        nbr += dst
        ctr += [i] * len(dst)
    return np.asarray([nbr, ctr], dtype=np.int64).reshape(2, -1)


def residue_features(seq):
    """process_mols.py:194-196: one categorical column per residue."""
    idx = [RESIDUE_ORDER.index(c) if c in RESIDUE_ORDER else MISC_RESIDUE for c in seq]
    return np.asarray(idx, dtype=np.float32)[:, None]


def extract_receptor_structure(seq, all_coords, neighbor_cutoff=20, max_neighbors=None, lm_embeddings=None):
    """new_extract_receptor_structure (process_mols.py:161-202) for ``all_atoms=False, knn_only_graph=False``.
    ``all_coords`` [N, n_atoms_per_residue, 3] (column 1 = C-alpha); ``lm_embeddings``: list of per-chain [L_c, D] arrays.
    Returns dict(x [N, 1 (+D)], pos [N, 3] fp32, edge_index [2, E])."""
    ca = np.asarray(all_coords)[:, 1, :].astype(np.float32)
    if len(ca) > 3000:
        raise ValueError(f'The receptor is too large {len(ca)}')
    ei = contact_graph(ca, neighbor_cutoff, max_neighbors)
    x = residue_features(seq)
    if lm_embeddings is not None:
        x = np.concatenate([x, np.concatenate([np.asarray(e) for e in lm_embeddings], 0).astype(np.float32)], 1)
    return {'x': x, 'pos': ca, 'edge_index': ei}


def lig_graph(atom_feats, bond_begin, bond_end, bond_type, pos=None):
    """get_lig_graph (process_mols.py:279-301) with the RDKit molecule replaced by its arrays: every bond listed in both
    directions (u->v, v->u consecutively), one-hot bond type (UNSPECIFIED -> class 0)."""
    row, col, typ = [], [], []
    for u, v, t in zip(bond_begin, bond_end, bond_type):
        row += [u, v]
        col += [v, u]
        typ += 2 * [t if 0 <= t < N_BOND_TYPES else 0]
    ea = np.zeros((len(typ), N_BOND_TYPES), np.float32)
    ea[np.arange(len(typ)), np.asarray(typ, dtype=np.int64)] = 1.0
    out = {'x': np.asarray(atom_feats), 'edge_index': np.asarray([row, col], dtype=np.int64).reshape(2, -1), 'edge_attr': ea}
    if pos is not None:
        out['pos'] = np.asarray(pos, dtype=np.float32)
    return out


def transformation_mask(edge_index, n_atoms):
    """get_transformation_mask (utils/torsion.py:15-45) without networkx: for each bond (edges 2k, 2k+1 are its two
    directions) remove it; if the molecule falls apart and the smallest part ``l`` (first of the smallest in component
    order = order of the lowest atom index) has more than one atom, ``l`` is attached to edge 2k+1 when edges[2k, 0] lies in
    ``l``, else to edge 2k (:27-32) - i.e. to the direction u->v whose head v is in the rotating part.
    Returns (mask_edges [E] bool, mask_rotate [n_rotatable, n_atoms] bool)."""
    edges = np.asarray(edge_index).T
    adj = [set() for _ in range(n_atoms)]
    for u, v in edges.tolist():
        adj[u].add(v)
        adj[v].add(u)

    def components(skip):
        seen, comps = set(), []
        for s in range(n_atoms):
            if s in seen:
                continue
            comp, stack = {s}, [s]
            while stack:
                a = stack.pop()
                for b in adj[a]:
                    if (a, b) == skip or (b, a) == skip or b in comp:
                        continue
                    comp.add(b)
                    stack.append(b)
            seen |= comp
            comps.append(comp)
        return comps

    to_rotate = []
    for i in range(0, edges.shape[0], 2):
        assert edges[i, 0] == edges[i + 1, 1]
        comps = components((int(edges[i, 0]), int(edges[i, 1])))
        if len(comps) > 1:
            small = sorted(comps, key=len)[0]         # sorted() is stable: first of the smallest (torsion.py:25)
            if len(small) > 1:
                if int(edges[i, 0]) in small:
                    to_rotate += [[], sorted(small)]
                else:
                    to_rotate += [sorted(small), []]
                continue
        to_rotate += [[], []]
    mask_edges = np.asarray([len(l) > 0 for l in to_rotate], dtype=bool)
    mask_rotate = np.zeros((int(mask_edges.sum()), n_atoms), dtype=bool)
    idx = 0
    for i in range(edges.shape[0]):
        if mask_edges[i]:
            mask_rotate[idx][np.asarray(to_rotate[i], dtype=int)] = True
            idx += 1
    return mask_edges, mask_rotate


def centre_complex(rec_pos, lig_pos):
    """InferenceDataset.get (utils/inference_utils.py:229-236): receptor and ligand are mean-centred separately (fp32
    means); returns (rec_pos, lig_pos, original_center [1, 3])."""
    import torch
    rp, lp = torch.as_tensor(rec_pos, dtype=torch.float32), torch.as_tensor(lig_pos, dtype=torch.float32)
    pc = torch.mean(rp, dim=0, keepdim=True)
    lc = torch.mean(lp, dim=0, keepdim=True)
    return (rp - pc).numpy(), (lp - lc).numpy(), pc.numpy()


def chain_embeddings(id_to_embeddings, names):
    """datasets/pdbbind.py:215-230: the cache of datasets/esm_embeddings_to_pt.py maps '<name>_chain_<k>' -> [L, D];
    per complex the chains are collected in dictionary order and re-ordered by k."""
    emb, idx = {}, {}
    for key, e in id_to_embeddings.items():
        name = key.split('_chain_')[0]
        if name in names:
            emb.setdefault(name, []).append(e)
            idx.setdefault(name, []).append(int(key.split('_chain_')[1]))
    out = []
    for name in names:
        order = np.argsort(idx.get(name, []))
        out.append([emb[name][i] for i in order])
    return out
