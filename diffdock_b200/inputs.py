"""Input side of the hot path: complex graphs from parsed arrays, on the device.

What the reference does between its parsers (RDKit / ProDy / Biopython / ESM - all outside section 8) and ``sampling()``:

  datasets/process_mols.py:161-202  new_extract_receptor_structure  residue features, C-alpha contact graph (cdist + a Python
                                                                     loop over residues: O(N^2) host memory, seconds)
  datasets/process_mols.py:279-301  get_lig_graph                   bond list -> edge_index / one-hot edge_attr
  utils/torsion.py:15-45            get_transformation_mask         rotatable-bond masks (networkx)
  utils/inference_utils.py:181-242  InferenceDataset.get            centring, ``original_center``, ``success``
  datasets/pdbbind.py:215-230       ESM cache lookup                '<name>_chain_<k>' -> per-chain embeddings, ordered by k
  inference.py:236-239              N x copy.deepcopy(complex)      one host copy of the receptor (with its 1280-wide LM
                                                                     embedding) per pose, each uploaded by the sampler

Here the same functions take the parsers' OUTPUT ARRAYS (sequence string, coordinates, feature matrix, bond list) and
build the graph with the tensors already on the GPU: the contact graph is one kernel pair (``ddb200_contact_count/_fill``,
the reference's edge list, bit for bit up to the order of exact-distance ties), a complex is packed into ONE pinned host buffer (one H2D copy, or a file that is
read straight into pinned memory), and the N poses handed to ``sampling()`` are N light graph objects that share the device
copy of the receptor - no deep copies, no per-pose upload.

No CPU fallback for the device parts: ``contact_graph`` needs the CUDA library and a CUDA tensor.
"""
from __future__ import annotations

import json
import struct
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .hetero import HeteroGraph, Store
from .ops import _need_cuda, _ptr, _stream

# index of a residue in allowable_features['possible_amino_acids'] (datasets/process_mols.py:47-49) through aa_short2long
# (datasets/constants.py:37-40); unknown letters map to the last entry, 'misc' (safe_index, process_mols.py:120-125)
_RESIDUE_LETTERS = 'ARNDCQEGHILKMFPSTWYV'
_MISC_RESIDUE = 37
N_BOND_TYPES = 4                  # process_mols.py:57
MAX_RESIDUES = 3000               # process_mols.py:169


# ---------------------------------------------------------------------------------------------------------------- receptor
def contact_graph(pos: torch.Tensor, cutoff: float, max_neighbors: Optional[int] = None, knn_only: bool = False) -> torch.Tensor:
    """Edge list [2, E] int64 (rows [neighbour, centre]) of datasets/process_mols.py:176-192: per centre the other points
    within ``cutoff`` in index order, the ``max_neighbors`` nearest (by distance) when there are more, the nearest one
    when there are none.  ``pos`` [N, 3] fp32 CUDA.  Squared distances are formed in torch.cdist's fp32 operation order
    (bit-identical); exact-distance ties, which np.argsort leaves unspecified, are broken by index."""
    _need_cuda(pos)
    assert pos.dtype == torch.float32 and pos.dim() == 2 and pos.shape[1] == 3
    pos = pos.contiguous()
    n = pos.shape[0]
    k = int(max_neighbors) if max_neighbors else 1000
    count = torch.empty(n, dtype=torch.int32, device=pos.device)
    lib = _lib.lib()
    _lib.check(lib.ddb200_contact_count(_ptr(pos), n, float(cutoff), k, int(bool(knn_only)), _ptr(count), _stream()),
               'ddb200_contact_count')
    incl = torch.cumsum(count, 0, dtype=torch.int32)
    total = int(incl[-1].item()) if n else 0            # preprocessing, once per complex: the size is read back
    row_start = (incl - count).contiguous()
    ei = torch.empty((2, total), dtype=torch.int32, device=pos.device)
    if total:
        _lib.check(lib.ddb200_contact_fill(_ptr(pos), n, float(cutoff), k, int(bool(knn_only)), _ptr(row_start), _ptr(ei[0]),
                                           _ptr(ei[1]), _stream()), 'ddb200_contact_fill')
    return ei.long()


def residue_features(seq: str) -> torch.Tensor:
    """[N, 1] fp32 categorical residue column (process_mols.py:194-196)."""
    idx = [_RESIDUE_LETTERS.index(c) if c in _RESIDUE_LETTERS else _MISC_RESIDUE for c in seq]
    return torch.tensor(idx, dtype=torch.float32).unsqueeze(1)


def new_extract_receptor_structure(seq, all_coords, complex_graph, neighbor_cutoff=20, max_neighbors=None, lm_embeddings=None,
                                   knn_only_graph=False, all_atoms=False, atom_cutoff=None, atom_max_neighbors=None,
                                   device='cuda'):
    """Same name, arguments and effect on ``complex_graph`` as datasets/process_mols.py:161-202, with the receptor tensors
    created on ``device``.  ``all_coords`` [N, atoms_per_residue, 3] (column 1 = C-alpha), ``lm_embeddings`` a list of
    per-chain [L_c, D] arrays.  ``side_chain_vecs`` is not produced (read only by the side-chain head, out of scope);
    ``all_atoms`` is refused: in the reference tree ``get_moad_atom_feats`` (:244-247) returns an empty feature array, so its
    own all-atom branch yields no atoms."""
    if all_atoms:
        raise NotImplementedError("all-atom featurisation: the reference's get_moad_atom_feats is an empty stub")
    coords = torch.as_tensor(np.asarray(all_coords)[:, 1, :], dtype=torch.float32)
    if len(coords) > MAX_RESIDUES:
        raise ValueError(f'The receptor is too large {len(coords)}')
    dev = torch.device(device)
    pos = coords.to(dev)
    k = max_neighbors if max_neighbors else (32 if knn_only_graph else None)
    edge_index = contact_graph(pos, neighbor_cutoff, k, knn_only=knn_only_graph)
    x = residue_features(seq)
    if lm_embeddings is not None:
        lm = torch.cat([torch.as_tensor(np.asarray(e), dtype=torch.float32) for e in lm_embeddings], 0)
        x = torch.cat([x, lm], 1)
    complex_graph['receptor'].x = x.to(dev)
    complex_graph['receptor'].pos = pos
    complex_graph['receptor', 'rec_contact', 'receptor'].edge_index = edge_index
    return


# ------------------------------------------------------------------------------------------------------------------ ligand
def get_lig_graph(atom_feats, bond_begin: Sequence[int], bond_end: Sequence[int], bond_type: Sequence[int], complex_graph,
                  pos=None):
    """datasets/process_mols.py:279-301 with the RDKit molecule replaced by what the function reads from it: the atom
    feature matrix (lig_atom_featurizer's output), the bond list (begin, end, type index per process_mols.py:57; anything
    outside 0..3 = UNSPECIFIED -> class 0) and the conformer's coordinates."""
    b, e = torch.as_tensor(bond_begin, dtype=torch.long), torch.as_tensor(bond_end, dtype=torch.long)
    t = torch.as_tensor(bond_type, dtype=torch.long)
    t = torch.where((t >= 0) & (t < N_BOND_TYPES), t, torch.zeros_like(t))
    row = torch.stack([b, e], 1).reshape(-1)
    col = torch.stack([e, b], 1).reshape(-1)
    edge_type = t.repeat_interleave(2)
    complex_graph['ligand'].x = torch.as_tensor(atom_feats)
    complex_graph['ligand', 'lig_bond', 'ligand'].edge_index = torch.stack([row, col], 0)
    complex_graph['ligand', 'lig_bond', 'ligand'].edge_attr = torch.nn.functional.one_hot(
        edge_type, num_classes=N_BOND_TYPES).to(torch.float)
    if pos is not None:
        complex_graph['ligand'].pos = torch.as_tensor(np.asarray(pos)).float()
    return


def get_transformation_mask(pyg_data):
    """utils/torsion.py:15-45: which directed bond edges are rotatable and which atoms each of them moves.  Integer host
    work on a graph of tens of atoms; connected components by depth-first search instead of networkx (component order =
    order of the lowest atom index, as networkx yields them for nodes added in index order)."""
    edges = pyg_data['ligand', 'ligand'].edge_index.T.cpu().numpy()
    n = int(pyg_data['ligand'].x.shape[0])
    adj: List[List[int]] = [[] for _ in range(n)]
    for u, v in edges.tolist():
        adj[u].append(v)
    for u, v in edges.tolist():              # to_undirected
        if u not in adj[v]:
            adj[v].append(u)
    to_rotate: List[List[int]] = []
    for i in range(0, edges.shape[0], 2):
        assert edges[i, 0] == edges[i + 1, 1]
        a, b = int(edges[i, 0]), int(edges[i, 1])
        comps, seen = [], np.zeros(n, dtype=bool)
        for s in range(n):
            if seen[s]:
                continue
            comp, stack = [s], [s]
            seen[s] = True
            while stack:
                p = stack.pop()
                for q in adj[p]:
                    if seen[q] or (p == a and q == b) or (p == b and q == a):
                        continue
                    seen[q] = True
                    comp.append(q)
                    stack.append(q)
            comps.append(comp)
        placed = False
        if len(comps) > 1:
            small = sorted(comps, key=len)[0]
            if len(small) > 1:
                to_rotate += ([[], small] if a in small else [small, []])
                placed = True
        if not placed:
            to_rotate += [[], []]
    mask_edges = np.asarray([len(l) > 0 for l in to_rotate], dtype=bool)
    mask_rotate = np.zeros((int(mask_edges.sum()), n), dtype=bool)
    idx = 0
    for i in range(edges.shape[0]):
        if mask_edges[i]:
            mask_rotate[idx][np.asarray(to_rotate[i], dtype=int)] = True
            idx += 1
    return mask_edges, mask_rotate


# ----------------------------------------------------------------------------------------------------------------- complex
def build_complex(name, seq, all_coords, atom_feats, bond_begin, bond_end, bond_type, lig_pos, lm_embeddings=None,
                  receptor_radius=30, c_alpha_max_neighbors=None, knn_only_graph=False, device='cuda') -> HeteroGraph:
    """InferenceDataset.get (utils/inference_utils.py:181-242) from parsed arrays: ligand graph + masks
    (get_lig_graph_with_matching's graph part, process_mols.py:323-384 without conformer matching), receptor graph,
    separate mean-centring of receptor and ligand, ``original_center``, ``success``.  Receptor tensors live on ``device``;
    the ligand stays on the host until the sampler batches the poses."""
    g = HeteroGraph()
    g['name'] = name
    get_lig_graph(atom_feats, bond_begin, bond_end, bond_type, g, pos=lig_pos)
    edge_mask, mask_rotate = get_transformation_mask(g)
    g['ligand'].edge_mask = torch.tensor(edge_mask)
    g['ligand'].mask_rotate = mask_rotate
    new_extract_receptor_structure(seq, all_coords, g, neighbor_cutoff=receptor_radius, max_neighbors=c_alpha_max_neighbors,
                                   lm_embeddings=lm_embeddings, knn_only_graph=knn_only_graph, device=device)
    protein_center = torch.mean(g['receptor'].pos, dim=0, keepdim=True)
    g['receptor'].pos = g['receptor'].pos - protein_center
    ligand_center = torch.mean(g['ligand'].pos, dim=0, keepdim=True)
    g['ligand'].pos = g['ligand'].pos - ligand_center
    g.original_center = protein_center
    g['success'] = True
    return g


def pose_copies(complex_graph: HeteroGraph, n: int) -> List[HeteroGraph]:
    """The ``[copy.deepcopy(orig_complex_graph) for _ in range(N)]`` of inference.py:236-239 without the copies: N graph
    objects with their own ligand store (the sampler rewrites ``pos`` per pose) that SHARE the receptor stores - on the
    device when ``complex_graph`` came from ``build_complex`` / ``PackedComplex.to`` - so ``sampling()`` uploads nothing
    for the receptor and embeds it once."""
    out = []
    for _ in range(n):
        h = HeteroGraph()
        for k, st in complex_graph._nodes.items():
            h._nodes[k] = Store(**st.__dict__) if k == 'ligand' else st
        for k, st in complex_graph._edges.items():
            h._edges[k] = st
        h._globals.update(complex_graph._globals)
        out.append(h)
    return out


# ---------------------------------------------------------------------------------------------------------- ESM embedding cache
class EsmCache:
    """The dictionary written by datasets/esm_embeddings_to_pt.py ('<complex>_chain_<k>' -> [L_k, 1280] layer-33
    representations) with the lookup of datasets/pdbbind.py:215-230: a complex's chains, ordered by chain index."""

    def __init__(self, id_to_embeddings: Dict[str, torch.Tensor]):
        self._chains: Dict[str, List] = {}
        for key, emb in id_to_embeddings.items():
            name, _, idx = key.partition('_chain_')
            self._chains.setdefault(name, []).append((int(idx), len(self._chains.get(name, ())), emb))

    @classmethod
    def load(cls, path):
        return cls(torch.load(path, map_location='cpu'))

    def chains(self, name: str) -> List[torch.Tensor]:
        items = self._chains.get(name, [])
        order = np.argsort([i for i, _, _ in items])          # the reference's np.argsort (pdbbind.py:228)
        return [items[j][2] for j in order]

    def __contains__(self, name):
        return name in self._chains


# ------------------------------------------------------------------------------------------------------------ packed complexes
_MAGIC = b'DDB2PACK'
_ALIGN = 256


class PackedComplex:
    """One complex as ONE contiguous byte buffer + a small header: every tensor attribute of every node / edge store, 256-byte
    aligned.  ``pack`` gathers them into pinned host memory, ``to(device)`` is a single H2D copy after which the graph's
    tensors are views into the device buffer; ``save`` / ``load`` move the same bytes to and from a file (the load reads
    straight into pinned memory).  Replaces the per-attribute ``.to(device)`` of torch_geometric's Batch (utils/sampling.py:80)
    and the pickled HeteroData caches of datasets/pdbbind.py:262-268 for the inference path."""

    def __init__(self, header: dict, buffer: torch.Tensor):
        self.header, self.buffer = header, buffer

    @staticmethod
    def pack(g: HeteroGraph, pin: bool = True) -> 'PackedComplex':
        entries, off = [], 0
        tensors = []

        def add(kind, key, attr, v):
            nonlocal off
            t = v.detach().cpu().contiguous()
            nbytes = t.numel() * t.element_size()
            entries.append({'kind': kind, 'key': list(key) if isinstance(key, tuple) else key, 'attr': attr,
                            'dtype': str(t.dtype).replace('torch.', ''), 'shape': list(t.shape), 'offset': off, 'nbytes': nbytes})
            tensors.append((off, t))
            off += (nbytes + _ALIGN - 1) // _ALIGN * _ALIGN

        extras = {'nodes': {}, 'edges': {}, 'globals': {}}
        for k, st in g._nodes.items():
            for a, v in st.__dict__.items():
                if a.startswith('_'):
                    continue
                if torch.is_tensor(v):
                    add('node', k, a, v)
                elif isinstance(v, np.ndarray):
                    add('node_np', k, a, torch.from_numpy(np.ascontiguousarray(v)))
                elif isinstance(v, list) and v and all(isinstance(q, np.ndarray) for q in v):
                    for q in v:        # e.g. mask_rotate = [array] (the list a PyG batch makes of it)
                        add('node_np_list', k, a, torch.from_numpy(np.ascontiguousarray(q)))
                else:
                    extras['nodes'].setdefault(k, {})[a] = v
        for k, st in g._edges.items():
            for a, v in st.__dict__.items():
                if torch.is_tensor(v):
                    add('edge', k, a, v)
        for a, v in g._globals.items():
            if torch.is_tensor(v):
                add('global', '', a, v)
            elif isinstance(v, (str, int, float, bool)) or v is None:
                extras['globals'][a] = v
        buf = torch.empty(max(off, 1), dtype=torch.uint8, pin_memory=pin and torch.cuda.is_available())
        for o, t in tensors:
            n = t.numel() * t.element_size()
            if n:
                buf[o:o + n] = t.reshape(-1).view(torch.uint8)
        return PackedComplex({'entries': entries, 'extras': extras, 'nbytes': off}, buf)

    def to(self, device, non_blocking: bool = True, host_keys: Sequence[str] = ('ligand',)) -> HeteroGraph:
        """HeteroGraph whose tensors are views of ONE device buffer (single copy).  Stores named in ``host_keys`` (default: the
        ligand, which the sampler re-batches per pose) stay views of the host buffer."""
        dev_buf = self.buffer.to(device, non_blocking=non_blocking)
        g = HeteroGraph()

        def view(e, buf):
            dt = getattr(torch, e['dtype'])
            n = e['nbytes']
            flat = buf[e['offset']:e['offset'] + n]
            if dt == torch.bool:
                return flat.view(torch.bool).reshape(e['shape'])
            return flat.view(dt).reshape(e['shape'])

        for e in self.header['entries']:
            key = tuple(e['key']) if isinstance(e['key'], list) else e['key']
            on_host = (key in host_keys) or (isinstance(key, tuple) and all(k in host_keys for k in key))
            buf = self.buffer if on_host else dev_buf
            if e['kind'] == 'node':
                setattr(g[key], e['attr'], view(e, buf))
            elif e['kind'] == 'node_np':
                setattr(g[key], e['attr'], view(e, self.buffer).numpy())
            elif e['kind'] == 'node_np_list':
                if e['attr'] not in g[key]:
                    setattr(g[key], e['attr'], [])
                getattr(g[key], e['attr']).append(view(e, self.buffer).numpy())
            elif e['kind'] == 'edge':
                setattr(g[key], e['attr'], view(e, buf))
            else:
                g._globals[e['attr']] = view(e, buf)
        for k, d in self.header['extras']['nodes'].items():
            for a, v in d.items():
                setattr(g[k], a, v)
        g._globals.update(self.header['extras']['globals'])
        return g                              # the views keep the device buffer alive

    # -- file format: magic | u64 header length | JSON header | padding to 256 | payload ------------------------------------
    def save(self, path):
        hdr = json.dumps(self.header).encode()
        with open(path, 'wb') as f:
            f.write(_MAGIC)
            f.write(struct.pack('<Q', len(hdr)))
            f.write(hdr)
            pad = (-(len(_MAGIC) + 8 + len(hdr))) % _ALIGN
            f.write(b'\0' * pad)
            f.write(self.buffer.numpy().tobytes())

    @staticmethod
    def load(path, pin: bool = True) -> 'PackedComplex':
        with open(path, 'rb') as f:
            if f.read(len(_MAGIC)) != _MAGIC:
                raise ValueError(f'{path}: not a packed complex')
            (n,) = struct.unpack('<Q', f.read(8))
            header = json.loads(f.read(n).decode())
            f.seek((len(_MAGIC) + 8 + n + _ALIGN - 1) // _ALIGN * _ALIGN)
            buf = torch.empty(max(header['nbytes'], 1), dtype=torch.uint8, pin_memory=pin and torch.cuda.is_available())
            got = f.readinto(memoryview(buf.numpy()))
            if got < header['nbytes']:
                raise ValueError(f'{path}: truncated payload ({got} of {header["nbytes"]} bytes)')
        return PackedComplex(header, buf)
