"""CPU: input side.  The oracle restatement (oracle/inputs.py) and the product's host-side
functions (diffdock_b200/inputs.py: ligand graph, rotatable-bond masks, ESM cache lookup, packed complexes) against
tests/golden/ref_inputs.pt, which tests/golden/make_golden_inputs.py recorded from the UNMODIFIED reference functions
(datasets/process_mols.py:161-202,279-301, utils/torsion.py:15-45, datasets/pdbbind.py:217-230).  Integer / index results
are compared bit for bit."""
import numpy as np
import pytest
import torch

from tests.parity_helpers import canonical_contact_edges, load_golden


@pytest.fixture(scope='module')
def fx():
    return load_golden('ref_inputs.pt')


def test_cdist_restatement_is_torch_cdist():
    """The arithmetic the contact graph depends on.  oracle.inputs.cdist_sq_f32 (and with it the CUDA kernel, which follows
    the same operation order) equals, bit for bit, the matrix product ATen's _euclidean_dist forms in the > 25-point regime;
    the distances themselves equal torch.cdist up to the 1-ulp sloppiness of torch's own vectorised CPU square root."""
    from oracle.inputs import cdist_f32, cdist_sq_f32
    g = torch.Generator().manual_seed(0)
    for n in (26, 300, 1100):
        x = (torch.randn(n, 3, generator=g) * 20 + 30).float()
        nrm = x.pow(2).sum(-1, keepdim=True)
        one = torch.ones_like(nrm)
        sq = torch.cat([x.mul(-2), nrm, one], -1).matmul(torch.cat([x, one, nrm], -1).mT).clamp_min(0)
        assert torch.equal(sq.sqrt(), torch.cdist(x, x))                 # this IS torch.cdist's recipe
        assert np.array_equal(cdist_sq_f32(x.numpy()), sq.numpy()), n
        d, t = cdist_f32(x.numpy()), torch.cdist(x, x).numpy()
        ulp = np.spacing(np.maximum(d, t))
        assert np.all(np.abs(d - t) <= ulp) and np.mean(d == t) > 0.98
    for n in (2, 7, 25):          # direct form (<= 25 points): torch takes an exact root there, so the distances agree bit for bit
        x = (torch.randn(n, 3, generator=g) * 20 + 30).float()
        assert np.array_equal(cdist_f32(x.numpy()), torch.cdist(x, x).numpy()), n


@pytest.mark.parametrize('i', range(5))
def test_oracle_receptor_matches_reference(fx, i):
    from oracle.inputs import extract_receptor_structure
    c = fx['receptor'][i]
    got = extract_receptor_structure(c['seq'], c['all_coords'].numpy(), c['cutoff'], c['max_neighbors'],
                                     [e.numpy() for e in c['lm']])
    want = canonical_contact_edges(c['edge_index'].numpy(), c['pos'].numpy())
    assert np.array_equal(got['edge_index'], want)
    assert int((want != c['edge_index'].numpy()).any(0).sum()) <= 2          # only exact-distance ties were re-ordered
    assert np.array_equal(got['x'], c['x'].numpy()) and np.array_equal(got['pos'], c['pos'].numpy())


@pytest.mark.parametrize('i', range(3))
def test_ligand_graph_and_masks_match_reference(fx, i):
    from oracle.inputs import lig_graph, transformation_mask
    from diffdock_b200.hetero import HeteroGraph
    from diffdock_b200.inputs import get_lig_graph, get_transformation_mask
    c = fx['ligand'][i]
    o = lig_graph(c['atom_feats'].numpy(), c['bond_begin'], c['bond_end'], c['bond_type'], c['pos_in'].numpy())
    me, mr = transformation_mask(o['edge_index'], o['x'].shape[0])
    g = HeteroGraph()
    get_lig_graph(c['atom_feats'], c['bond_begin'], c['bond_end'], c['bond_type'], g, pos=c['pos_in'].numpy())
    pe, pr = get_transformation_mask(g)
    for ei, ea, x, pos, m_e, m_r in ((o['edge_index'], o['edge_attr'], o['x'], o['pos'], me, mr),
                                     (g['ligand', 'ligand'].edge_index.numpy(), g['ligand', 'ligand'].edge_attr.numpy(),
                                      g['ligand'].x.numpy(), g['ligand'].pos.numpy(), pe, pr)):
        assert np.array_equal(ei, c['edge_index'].numpy()) and np.array_equal(ea, c['edge_attr'].numpy())
        assert np.array_equal(x, c['x'].numpy()) and np.array_equal(pos, c['pos'].numpy())
        assert np.array_equal(m_e, c['mask_edges'].numpy()) and np.array_equal(m_r, c['mask_rotate'].numpy())
    assert g['ligand', 'ligand'].edge_attr.dtype == torch.float32 and g['ligand', 'ligand'].edge_index.dtype == torch.int64


def test_esm_cache_lookup_matches_reference(fx):
    from oracle.inputs import chain_embeddings
    from diffdock_b200.inputs import EsmCache
    c = fx['esm']
    want = c['chains']
    got_o = chain_embeddings(c['cache'], c['names'])
    cache = EsmCache(c['cache'])
    for name, w, o in zip(c['names'], want, got_o):
        p = cache.chains(name)
        assert len(w) == len(o) == len(p)
        for a, b, d in zip(w, o, p):
            assert torch.equal(a, b) and torch.equal(a, d)
    assert 'other' in cache and 'absent' not in cache and cache.chains('absent') == []


def test_residue_features_and_size_limit():
    from diffdock_b200.inputs import residue_features, new_extract_receptor_structure
    from diffdock_b200.hetero import HeteroGraph
    x = residue_features('AVXU?')
    assert x.tolist() == [[0.0], [19.0], [37.0], [37.0], [37.0]]
    with pytest.raises(ValueError, match='too large'):
        new_extract_receptor_structure('A' * 3001, np.zeros((3001, 3, 3), np.float32), HeteroGraph(), device='cpu')
    with pytest.raises(NotImplementedError):
        new_extract_receptor_structure('A', np.zeros((1, 3, 3), np.float32), HeteroGraph(), all_atoms=True)


def test_contact_graph_refuses_cpu_tensors():
    from diffdock_b200.inputs import contact_graph
    with pytest.raises(RuntimeError):
        contact_graph(torch.zeros(4, 3), 5.0, 3)


def test_packed_complex_round_trip(tmp_path):
    """pack -> (file) -> unpack gives back every tensor bit for bit, with the declared alignment, and pose_copies shares the
    receptor stores."""
    from diffdock_b200.inputs import PackedComplex, pose_copies
    from diffdock_b200.synthetic import make_complex
    g = make_complex(n_res=60, n_atoms=11, seed=3)
    pk = PackedComplex.pack(g, pin=False)
    assert all(e['offset'] % 256 == 0 for e in pk.header['entries'])
    path = tmp_path / 'c.ddpk'
    pk.save(path)
    for src in (pk, PackedComplex.load(path, pin=False)):
        h = src.to('cpu')
        for k, st in g._nodes.items():
            for a, v in st.__dict__.items():
                if a.startswith('_'):
                    continue
                w = getattr(h[k], a)
                if torch.is_tensor(v):
                    assert torch.equal(v, w) and v.dtype == w.dtype, (k, a)
                elif isinstance(v, np.ndarray):
                    assert np.array_equal(v, w) and v.dtype == w.dtype, (k, a)
                elif isinstance(v, list) and v and isinstance(v[0], np.ndarray):
                    assert len(v) == len(w) and all(np.array_equal(p, q) and p.dtype == q.dtype for p, q in zip(v, w)), (k, a)
                else:
                    assert v == w, (k, a)
        for k, st in g._edges.items():
            for a, v in st.__dict__.items():
                assert torch.equal(v, getattr(h[k], a)), (k, a)
        assert h['name'] == g['name']
    poses = pose_copies(h, 3)
    assert poses[0]['receptor'] is poses[2]['receptor'] and poses[0]['ligand'] is not poses[1]['ligand']
    poses[0]['ligand'].pos = poses[0]['ligand'].pos + 1.0
    assert not torch.equal(poses[0]['ligand'].pos, poses[1]['ligand'].pos)
    with open(path, 'r+b') as f:
        f.truncate(path.stat().st_size - 100)
    with pytest.raises(ValueError, match='truncated'):
        PackedComplex.load(path, pin=False)


def test_pose_copies_take_the_shared_receptor_collate():
    """The sampler's shared-receptor collate recognises ``pose_copies`` by pointer identity (no tensor comparison) and gives the
    batch the general collate of deep copies gives."""
    import copy
    from diffdock_b200.hetero import collate, collate_shared_receptor
    from diffdock_b200.inputs import pose_copies
    from diffdock_b200.synthetic import make_complex
    g = make_complex(n_res=50, n_atoms=9, seed=7)
    light = pose_copies(g, 4)
    deep = [copy.deepcopy(g) for _ in range(4)]
    for i in range(4):
        light[i]['ligand'].pos = light[i]['ligand'].pos + float(i)
        deep[i]['ligand'].pos = deep[i]['ligand'].pos + float(i)
    a, b = collate_shared_receptor(light, 'cpu'), collate(deep)
    assert a['receptor']._unique == (50, g['receptor', 'receptor'].edge_index.shape[1], 4)
    for key in ('ligand', 'receptor'):
        for attr in ('x', 'pos', 'batch'):
            assert torch.equal(getattr(a[key], attr), getattr(b[key], attr)), (key, attr)
    for key in (('ligand', 'ligand'), ('receptor', 'receptor')):
        assert torch.equal(a[key].edge_index, b[key].edge_index), key
    assert a.num_graphs == 4
