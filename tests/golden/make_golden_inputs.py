"""Golden vectors of the INPUT SIDE: runs the UNMODIFIED reference functions
    datasets/process_mols.py  new_extract_receptor_structure (:161-202), get_lig_graph (:279-301)
    utils/torsion.py          get_transformation_mask (:15-45)
    datasets/pdbbind.py       the ESM chain collection / re-ordering lines :217-230 (executed from the file's own source)
on synthetic arrays in the authoring container and stores inputs + outputs in tests/golden/ref_inputs.pt.

    python tests/golden/make_golden_inputs.py

The parsers the reference feeds these functions from (RDKit, ProDy, Biopython) are not installed and outside section 8:
oracle/ref_shims.py makes the modules importable, the RDKit molecule is replaced by a minimal stand-in exposing the five
methods get_lig_graph calls, ``lig_atom_featurizer`` by the given feature matrix, torch_geometric's ``to_networkx`` by a
three-line DiGraph builder.  Everything that decides the fixture's content is the reference's own code.
"""
import ast
import os
import sys

import networkx as nx
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install()
import datasets.process_mols as pm      # noqa: E402
import utils.torsion as r_torsion       # noqa: E402
from diffdock_b200.hetero import HeteroGraph   # noqa: E402
from diffdock_b200.synthetic import _ligand    # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'ref_inputs.pt')


# ---- stand-ins for the third-party objects -----------------------------------------------------------------------------
class _Bond:
    def __init__(self, u, v, t):
        self.u, self.v, self.t = u, v, t

    def GetBeginAtomIdx(self):
        return self.u

    def GetEndAtomIdx(self):
        return self.v

    def GetBondType(self):
        return self.t


class _Conf:
    def __init__(self, pos):
        self.pos = pos

    def GetPositions(self):
        return self.pos


class _Mol:
    def __init__(self, bonds, pos):
        self.bonds, self.pos = bonds, pos

    def GetBonds(self):
        return self.bonds

    def GetNumConformers(self):
        return 1

    def GetConformer(self):
        return _Conf(self.pos)


def _to_networkx(data, to_undirected=False):
    g = nx.DiGraph()
    g.add_nodes_from(range(data['ligand'].x.shape[0]))
    g.add_edges_from(data['ligand', 'ligand'].edge_index.T.tolist())
    return g


HeteroGraph.to_homogeneous = lambda self: self
r_torsion.to_networkx = _to_networkx
BOND_KEYS = list(pm.bonds.keys())         # the placeholder objects standing for BT.SINGLE / DOUBLE / TRIPLE / AROMATIC


def receptor_case(n_res, cutoff, max_neighbors, seed, lm_dim=8, chains=2):
    rng = np.random.default_rng(seed)
    R = (3.0 * n_res / (4.0 * np.pi * 0.0075)) ** (1.0 / 3.0)
    v = rng.normal(size=(n_res, 3))
    ca = v / np.linalg.norm(v, axis=1, keepdims=True) * (R * rng.uniform(size=(n_res, 1)) ** (1.0 / 3.0)) + 30.0
    all_coords = np.full((n_res, 14, 3), np.nan)
    all_coords[:, 1] = ca
    all_coords[:, 0] = ca + rng.normal(size=(n_res, 3))
    all_coords[:, 2] = ca + rng.normal(size=(n_res, 3))
    letters = 'ARNDCQEGHILKMFPSTWYVXU'
    seq = ''.join(letters[i] for i in rng.integers(0, len(letters), size=n_res))
    cuts = [0] + sorted(rng.choice(np.arange(1, n_res), size=chains - 1, replace=False).tolist()) + [n_res]
    lm = [rng.normal(size=(cuts[c + 1] - cuts[c], lm_dim)).astype(np.float32) for c in range(chains)]
    g = HeteroGraph()
    pm.new_extract_receptor_structure(seq, all_coords.astype(np.float32), g, neighbor_cutoff=cutoff,
                                      max_neighbors=max_neighbors, lm_embeddings=lm)
    return {'seq': seq, 'all_coords': torch.from_numpy(all_coords.astype(np.float32)), 'lm': [torch.from_numpy(e) for e in lm],
            'cutoff': cutoff, 'max_neighbors': max_neighbors,
            'x': g['receptor'].x.clone(), 'pos': g['receptor'].pos.clone(),
            'edge_index': g['receptor', 'receptor'].edge_index.clone()}


def ligand_case(n_atoms, seed, ring=False):
    rng = np.random.default_rng(seed)
    pos, ei, _, x, _, _ = _ligand(n_atoms, rng)
    begin, end = ei[0, 0::2].tolist(), ei[1, 0::2].tolist()
    if ring and n_atoms > 6:        # close one ring: its bonds are not rotatable
        begin.append(0)
        end.append(5)
    btype = rng.integers(0, 5, size=len(begin)).tolist()            # 4 = UNSPECIFIED
    unspecified = object()
    bonds = [_Bond(u, v, BOND_KEYS[t] if t < 4 else unspecified) for u, v, t in zip(begin, end, btype)]
    pm.lig_atom_featurizer = lambda mol: torch.from_numpy(x)
    # BT.UNSPECIFIED is a fresh placeholder on every access in the shimmed module, so `!=` is always true there; give the
    # comparison the reference's meaning for the stand-in
    class _BT:
        UNSPECIFIED = unspecified
    pm.BT = _BT
    g = HeteroGraph()
    pm.get_lig_graph(_Mol(bonds, pos.astype(np.float64)), g)
    mask_edges, mask_rotate = r_torsion.get_transformation_mask(g)
    return {'atom_feats': torch.from_numpy(x), 'bond_begin': begin, 'bond_end': end, 'bond_type': btype,
            'pos_in': torch.from_numpy(pos.astype(np.float64)),
            'x': g['ligand'].x.clone(), 'pos': g['ligand'].pos.clone(), 'edge_index': g['ligand', 'ligand'].edge_index.clone(),
            'edge_attr': g['ligand', 'ligand'].edge_attr.clone(), 'mask_edges': torch.from_numpy(mask_edges),
            'mask_rotate': torch.from_numpy(mask_rotate)}


def esm_case(seed):
    """Executes datasets/pdbbind.py:217-230 (the statements between `id_to_embeddings = torch.load(...)` and the `else:`)
    straight from the reference file on a synthetic cache."""
    rng = np.random.default_rng(seed)
    names = ['6abc', '1xyz', '9q9q']
    cache = {}
    for name, n_chains in zip(names, (3, 1, 12)):
        for k in rng.permutation(n_chains).tolist():
            cache[f'{name}_chain_{k}'] = torch.from_numpy(rng.normal(size=(int(rng.integers(3, 9)), 4)).astype(np.float32))
    cache['other_chain_0'] = torch.zeros(2, 4)
    src = open('/root/reference/datasets/pdbbind.py').read().splitlines()
    lo = next(i for i, l in enumerate(src) if 'chain_embeddings_dictlist = defaultdict(list)' in l)
    hi = next(i for i, l in enumerate(src) if i > lo and l.strip() == 'else:')
    block = '\n'.join(l[12:] for l in src[lo:hi])
    ast.parse(block)
    from collections import defaultdict
    env = {'defaultdict': defaultdict, 'np': np, 'id_to_embeddings': cache, 'complex_names_all': names}
    exec(block, env)
    return {'cache': cache, 'names': names, 'chains': env['lm_embeddings_chains_all']}


if __name__ == '__main__':
    fx = {'receptor': [receptor_case(300, 15.0, 24, 0), receptor_case(1500, 15.0, 24, 1), receptor_case(120, 6.0, 10, 2),
                       receptor_case(20, 9.0, 4, 3), receptor_case(200, 30.0, None, 4)],
          'ligand': [ligand_case(12, 0), ligand_case(40, 1), ligand_case(25, 2, ring=True)],
          'esm': esm_case(0)}
    torch.save(fx, OUT)
    print(OUT, os.path.getsize(OUT) // 1024, 'KiB')
    for c in fx['receptor']:
        print('receptor', len(c['seq']), 'edges', c['edge_index'].shape[1])
    for c in fx['ligand']:
        print('ligand', c['x'].shape[0], 'rotatable', int(c['mask_edges'].sum()))
