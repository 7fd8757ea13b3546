"""Import shims that let the UNMODIFIED reference modules under /root/reference be imported in the authoring
container, where e3nn / torch_scatter / torch_cluster / torch_geometric / rdkit / prody / esm are not installed.
TEST INFRASTRUCTURE - used only by tests/golden/make_golden.py (fixture generation); never at test run time on the
GPU box (where /root/reference does not exist) and never by the product.

* Arithmetic third-party packages are provided by the oracle's restatements (oracle/e3nn_lite.py,
  oracle/graph_ops.py), so the reference's own wiring code runs on top of them.
* torch_geometric's container/loader classes are mapped onto diffdock_b200.hetero.
* Pure I/O packages (rdkit, Bio, prody, esm, ...) become permissive placeholder modules: importable, never called.
"""
import importlib.abc
import importlib.machinery
import sys
import types

import torch

from . import e3nn_lite, graph_ops

REFERENCE_ROOT = '/root/reference'
_PLACEHOLDER_ROOTS = ('rdkit', 'Bio', 'prody', 'esm', 'wandb', 'spyrmsd', 'openbabel', 'graph_tool')


class _Anything:
    """Callable, attribute-ful placeholder."""

    def __init__(self, name='placeholder'):
        self.__name = name

    def __call__(self, *a, **k):
        return _Anything(self.__name + '()')

    def __getattr__(self, k):
        if k.startswith('__') and k.endswith('__'):
            raise AttributeError(k)
        return _Anything(f'{self.__name}.{k}')

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _PlaceholderModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, k):
        if k.startswith('__') and k.endswith('__'):
            raise AttributeError(k)
        v = _Anything(f'{self.__name__}.{k}')
        setattr(self, k, v)
        return v


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split('.')[0] in _PLACEHOLDER_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _PlaceholderModule(spec.name)

    def exec_module(self, module):
        pass


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []
    sys.modules[name] = m
    return m


def install():
    """Idempotently install the shims and put the reference tree on sys.path."""
    if getattr(install, '_done', False):
        return
    install._done = True
    sys.meta_path.append(_Finder())

    # ---- e3nn ---------------------------------------------------------------------------------
    class _Linear(torch.nn.Module):  # o3.Linear: only constructed when sidechain_pred/depthwise (out of scope)
        def __init__(self, *a, **k):
            raise NotImplementedError("o3.Linear is outside the restated subset")

    o3 = _mod('e3nn.o3', Irreps=e3nn_lite.Irreps, Irrep=e3nn_lite.Irrep,
              spherical_harmonics=e3nn_lite.spherical_harmonics,
              FullyConnectedTensorProduct=e3nn_lite.FullyConnectedTensorProduct,
              FullTensorProduct=e3nn_lite.FullTensorProduct, TensorProduct=e3nn_lite.TensorProduct,
              wigner_3j=e3nn_lite.wigner_3j, Linear=_Linear)
    nn = _mod('e3nn.nn', BatchNorm=e3nn_lite.BatchNorm)
    _mod('e3nn', o3=o3, nn=nn)

    # ---- torch_scatter / torch_cluster --------------------------------------------------------
    _mod('torch_scatter', scatter=graph_ops.scatter, scatter_mean=graph_ops.scatter_mean,
         scatter_add=lambda s, i, dim=0, dim_size=None: graph_ops.scatter(s, i, dim, dim_size, 'sum'),
         scatter_max=_Anything('scatter_max'), scatter_min=_Anything('scatter_min'), scatter_std=_Anything('scatter_std'))
    _mod('torch_cluster', radius=graph_ops.radius, radius_graph=graph_ops.radius_graph, knn_graph=_Anything('knn_graph'),
         knn=_Anything('knn'))

    # ---- torch_geometric ----------------------------------------------------------------------
    from diffdock_b200 import hetero

    class Batch:
        @staticmethod
        def from_data_list(dl):
            return hetero.collate(dl)

    class DataLoader:
        """torch_geometric.loader.DataLoader for lists of HeteroGraph: sequential mini-batches."""

        def __init__(self, dataset, batch_size=1, shuffle=False, **kw):
            assert not shuffle
            self.dataset, self.batch_size = dataset, batch_size

        def __iter__(self):
            for i in range(0, len(self.dataset), self.batch_size):
                yield hetero.collate(self.dataset[i:i + self.batch_size])

        def __len__(self):
            return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    data = _mod('torch_geometric.data', Batch=Batch, Data=type('Data', (), {}), HeteroData=hetero.HeteroGraph,
                Dataset=type('Dataset', (), {}))
    loader = _mod('torch_geometric.loader', DataLoader=DataLoader, DataListLoader=DataLoader)
    _mod('torch_geometric.loader.dataloader', Collater=_Anything('Collater'))
    def subgraph(subset, edge_index, edge_attr=None, relabel_nodes=False, num_nodes=None):
        """torch_geometric.utils.subgraph for a boolean node mask: edges with both end points kept, optionally relabelled."""
        assert subset.dtype == torch.bool
        ok = subset[edge_index[0]] & subset[edge_index[1]]
        ei = edge_index[:, ok]
        if relabel_nodes:
            ei = (torch.cumsum(subset.long(), 0) - 1)[ei]
        return ei, (edge_attr[ok] if edge_attr is not None else None)

    utils = _mod('torch_geometric.utils', to_networkx=_Anything('to_networkx'), subgraph=subgraph,
                 degree=_Anything('degree'))
    dp = _mod('torch_geometric.nn.data_parallel', DataParallel=_Anything('DataParallel'))
    tnn = _mod('torch_geometric.nn', data_parallel=dp, DataParallel=_Anything('DataParallel'))
    tr = _mod('torch_geometric.transforms', BaseTransform=type('BaseTransform', (), {}))
    _mod('torch_geometric', data=data, loader=loader, utils=utils, nn=tnn, transforms=tr)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
