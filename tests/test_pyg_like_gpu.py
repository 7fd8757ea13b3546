"""GPU: the product CGModel driven the way the reference drives it, with a batch object that is NOT diffdock_b200's
HeteroGraph.  torch_geometric is not installed here, so the batch is a minimal stand-in with PyG's storage behaviour
(2-tuple keys resolving to 3-tuple edge types, attribute stores, ``num_graphs``); the loop body is a restatement of
utils/sampling.py:96-131 (set_time -> model(batch)[:3]) with the reference's own ``set_time`` arithmetic
(utils/diffusion_utils.py:146-168)."""
import pytest
import torch

from tests.parity_helpers import make_model_pair, rel_err

pytestmark = pytest.mark.gpu


class _Storage:
    """Attribute bag like torch_geometric.data.storage.BaseStorage (attribute and item access, no other behaviour)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __getitem__(self, k):
        return self.__dict__[k]

    def __setitem__(self, k, v):
        self.__dict__[k] = v

    def __contains__(self, k):
        return k in self.__dict__

    @property
    def num_nodes(self):
        return self.__dict__['pos'].shape[0] if 'pos' in self.__dict__ else self.__dict__['x'].shape[0]

    @property
    def num_edges(self):
        return self.__dict__['edge_index'].shape[1]


class MiniPyGBatch:
    """HeteroDataBatch look-alike: node stores by name, edge stores by 3-tuple, 2-tuple keys resolve to the unique edge type
    with those end points (datasets/process_mols.py:202,294-295), globals as attributes."""

    def __init__(self, nodes, edges, num_graphs, **globals_):
        object.__setattr__(self, '_node_store', nodes)
        object.__setattr__(self, '_edge_store', edges)
        object.__setattr__(self, '_glob', dict(num_graphs=num_graphs, **globals_))

    def __getitem__(self, key):
        if isinstance(key, tuple):
            if len(key) == 2:
                hits = [k for k in self._edge_store if k[0] == key[0] and k[-1] == key[1]]
                assert len(hits) == 1
                key = hits[0]
            return self._edge_store[key]
        if key in self._node_store:
            return self._node_store[key]
        return self._glob[key]

    def __getattr__(self, name):
        g = object.__getattribute__(self, '_glob')
        if name in g:
            return g[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        self._glob[name] = value


def _from_hetero(g, device):
    """Re-pack a collated diffdock_b200 batch into the stand-in (tensors moved to `device`)."""
    mv = lambda v: v.to(device) if torch.is_tensor(v) else v
    nodes = {nt: _Storage(**{k: mv(v) for k, v in g[nt].__dict__.items() if not k.startswith('_')}) for nt in g.node_types}
    names = {('ligand', 'ligand'): ('ligand', 'lig_bond', 'ligand'), ('receptor', 'receptor'): ('receptor', 'rec_contact', 'receptor')}
    edges = {names[et]: _Storage(**{k: mv(v) for k, v in g[et].__dict__.items() if not k.startswith('_')}) for et in g.edge_types}
    return MiniPyGBatch(nodes, edges, g.num_graphs, name=g['name'])


def _reference_set_time(complex_graphs, t, t_tr, t_rot, t_tor, batchsize, all_atoms, device):
    """utils/diffusion_utils.py:146-168 restated for the coarse-grained case."""
    for nt in ('ligand', 'receptor'):
        n = complex_graphs[nt].num_nodes
        complex_graphs[nt].node_t = {'tr': t_tr * torch.ones(n).to(device), 'rot': t_rot * torch.ones(n).to(device),
                                     'tor': t_tor * torch.ones(n).to(device)}
    complex_graphs.complex_t = {'tr': t_tr * torch.ones(batchsize).to(device), 'rot': t_rot * torch.ones(batchsize).to(device),
                                'tor': t_tor * torch.ones(batchsize).to(device)}


@pytest.mark.parametrize("sync_free", [True, False])
def test_model_accepts_pyg_like_batch_in_the_reference_loop(built_lib, sync_free):
    from diffdock_b200.synthetic import default_model_args, make_pose_list
    from diffdock_b200.hetero import collate
    from diffdock_b200.diffusion_utils import get_t_schedule
    from oracle.diffusion import set_time as o_set_time
    args = default_model_args(ns=16, nv=4, sh_lmax=2, num_conv_layers=3, distance_embed_dim=16, cross_distance_embed_dim=16,
                              sigma_embed_dim=16)
    o, p = make_model_pair(args, seed=17)
    p._sync_free = None if sync_free else False
    poses = make_pose_list(3, n_res=50, n_atoms=10, seed=61, tr_sigma_max=args.tr_sigma_max)
    batch = _from_hetero(collate(poses), 'cuda:0')
    g_cpu = collate(poses)
    sched = get_t_schedule('expbeta', 4)
    for t_idx in range(3):                                   # utils/sampling.py:96-116
        t_tr, t_rot, t_tor = sched[t_idx], sched[t_idx], sched[t_idx]
        _reference_set_time(batch, None, t_tr, t_rot, t_tor, 3, False, 'cuda:0')
        with torch.no_grad():
            tr_score, rot_score, tor_score = p(batch)[:3]
        o_set_time(g_cpu, t_tr, t_tr, t_tr, 3, 'cpu')
        with torch.no_grad():
            ref = o(g_cpu)
        for a, b in zip((tr_score, rot_score, tor_score), ref[:3]):
            assert rel_err(a, b) < 1e-4
        # the caches the reference leaves on the batch (models/cg_model.py:292-295,385,474)
        assert 'rec_node_attr' in batch['receptor'] and 'rec_edge_attr' in batch['receptor', 'receptor']
        assert 'node_sigma_emb' in batch['ligand'] and batch.graph_sigma_emb.shape[0] == 3
