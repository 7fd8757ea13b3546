"""Oracle restatement of models/tensor_layers.py (the equivariant graph convolution).
TEST INFRASTRUCTURE (see oracle/__init__.py)."""
import math

import torch
import torch.nn.functional as F
from torch import nn

from . import e3nn_lite as o3
from .graph_ops import scatter
from .layers import fc_block


def get_irrep_seq(ns, nv, use_second_order_repr, reduce_pseudoscalars):
    """models/tensor_layers.py:17-32."""
    last = nv if reduce_pseudoscalars else ns
    if use_second_order_repr:
        return [f'{ns}x0e',
                f'{ns}x0e + {nv}x1o + {nv}x2e',
                f'{ns}x0e + {nv}x1o + {nv}x2e + {nv}x1e + {nv}x2o',
                f'{ns}x0e + {nv}x1o + {nv}x2e + {nv}x1e + {nv}x2o + {last}x0o']
    return [f'{ns}x0e',
            f'{ns}x0e + {nv}x1o',
            f'{ns}x0e + {nv}x1o + {nv}x1e',
            f'{ns}x0e + {nv}x1o + {nv}x1e + {last}x0o']


class FasterTensorProduct(nn.Module):
    """models/tensor_layers.py:44-122: closed-form l<=1 tensor product with sh = 0e+1o.  Weight layout: four
    [fan_in, mul_out] blocks in the order 0e,1o,1e,0o, fan-in rows in the append order below, scale 1/sqrt(fan_in)."""
    KEYS = ('0e', '1o', '1e', '0o')

    def __init__(self, in_irreps, sh_irreps, out_irreps, **kw):
        super().__init__()
        assert o3.Irreps(sh_irreps) == o3.Irreps('1x0e+1x1o')
        self.in_irreps, self.out_irreps = o3.Irreps(in_irreps), o3.Irreps(out_irreps)
        im = {k: 0 for k in self.KEYS}
        om = {k: 0 for k in self.KEYS}
        for m, ir in self.in_irreps:
            im[str(ir)] = m
        for m, ir in self.out_irreps:
            om[str(ir)] = m
        self.weight_shapes = {'0e': (im['0e'] + im['1o'], om['0e']),
                              '1o': (im['0e'] + im['1o'] + im['1e'], om['1o']),
                              '1e': (im['1o'] + im['1e'] + im['0o'], om['1e']),
                              '0o': (im['1e'] + im['0o'], om['0o'])}
        self.weight_numel = sum(a * b for a, b in self.weight_shapes.values())

    def forward(self, x, sh, weight):
        f = {}
        for (m, ir), sl in zip(self.in_irreps, self.in_irreps.slices()):
            v = x[..., sl]
            f[str(ir)] = v.reshape(*v.shape[:-1], m, 3) if ir[0] == 1 else v
        y0, y1 = sh[..., 0:1], sh[..., 1:4]           # [E,1], [E,3]
        mid = {k: [] for k in self.KEYS}
        if '0e' in f:
            mid['0e'].append(f['0e'] * y0)
            mid['1o'].append(f['0e'][..., None] * y1[..., None, :])
        if '1o' in f:
            mid['0e'].append((f['1o'] * y1[..., None, :]).sum(-1) / math.sqrt(3))
            mid['1o'].append(f['1o'] * y0[..., None])
            mid['1e'].append(torch.linalg.cross(f['1o'], y1[..., None, :].expand_as(f['1o']), dim=-1) / math.sqrt(2))
        if '1e' in f:
            mid['1o'].append(torch.linalg.cross(f['1e'], y1[..., None, :].expand_as(f['1e']), dim=-1) / math.sqrt(2))
            mid['1e'].append(f['1e'] * y0[..., None])
            mid['0o'].append((f['1e'] * y1[..., None, :]).sum(-1) / math.sqrt(3))
        if '0o' in f:
            mid['1e'].append(f['0o'][..., None] * y1[..., None, :])
            mid['0o'].append(f['0o'] * y0)
        w, start = {}, 0
        for k in self.KEYS:
            a, b = self.weight_shapes[k]
            w[k] = weight[..., start:start + a * b].reshape(*weight.shape[:-1], a, b) / math.sqrt(a) if a * b else None
            start += a * b
        res = {}
        for k in ('0e', '0o'):
            if mid[k]:
                res[k] = torch.einsum('...u,...uw->...w', torch.cat(mid[k], -1), w[k])
        for k in ('1o', '1e'):
            if mid[k]:
                r = torch.einsum('...uk,...uw->...wk', torch.cat(mid[k], -2), w[k])
                res[k] = r.reshape(*r.shape[:-2], -1)
        return torch.cat([res[str(ir)] for _, ir in self.out_irreps], -1)


class TensorProductConvLayer(nn.Module):
    """models/tensor_layers.py:234-335 (depthwise=False).  forward = conv (:125-231) -> BatchNorm -> residual."""

    def __init__(self, in_irreps, sh_irreps, out_irreps, n_edge_features, residual=True, batch_norm=True,
                 dropout=0.0, hidden_features=None, faster=False, edge_groups=1, tp_weights_layers=2,
                 activation='relu', depthwise=False):
        super().__init__()
        assert not depthwise, "depthwise_convolution is outside the hot-path scope"
        self.in_irreps, self.out_irreps, self.sh_irreps = in_irreps, out_irreps, sh_irreps
        self.residual, self.edge_groups = residual, edge_groups
        self.out_size = o3.Irreps(out_irreps).dim
        hidden_features = n_edge_features if hidden_features is None else hidden_features
        self.tp = (FasterTensorProduct(in_irreps, sh_irreps, out_irreps) if faster
                   else o3.FullyConnectedTensorProduct(in_irreps, sh_irreps, out_irreps, shared_weights=False))
        mk = lambda: fc_block(n_edge_features, hidden_features, self.tp.weight_numel, tp_weights_layers, dropout,
                              activation)
        self.fc = mk() if edge_groups == 1 else nn.ModuleList([mk() for _ in range(edge_groups)])
        self.batch_norm = o3.BatchNorm(out_irreps) if batch_norm else None

    def forward(self, node_attr, edge_index, edge_attr, edge_sh, out_nodes=None, reduce='mean', edge_weight=1.0):
        if edge_index.shape[1] == 0 and node_attr.shape[0] == 0:
            raise ValueError("No edges and no nodes")
        dt = node_attr.dtype
        if edge_index.shape[1] == 0:
            out = torch.zeros((node_attr.shape[0], self.out_size), dtype=dt, device=node_attr.device)
        else:
            tgt, src = edge_index[0], edge_index[1]   # gather row 1, scatter onto row 0 (hazard C.1)
            n_out = out_nodes or node_attr.shape[0]
            n_out = int(n_out)
            if self.edge_groups == 1:                  # tp_scatter_simple, :125-145
                assert torch.is_tensor(edge_attr)
                w = self.fc(edge_attr).to(dt) * edge_weight
                out = scatter(self.tp(node_attr[src], edge_sh, w), tgt, dim=0, dim_size=n_out, reduce=reduce)
            else:                                      # tp_scatter_multigroup, :148-231
                assert isinstance(edge_attr, list) and reduce in ('mean', 'sum')
                assert sum(a.shape[0] for a in edge_attr) == edge_index.shape[1]
                out = torch.zeros((n_out, self.out_size), dtype=dt, device=node_attr.device)
                cnt = torch.zeros(n_out, dtype=dt, device=node_attr.device)
                s = 0
                for g, ea in enumerate(edge_attr):
                    e = s + ea.shape[0]
                    fc = self.fc[g] if isinstance(self.fc, nn.ModuleList) else self.fc
                    w = fc(ea)
                    w = w * (edge_weight[s:e] if hasattr(edge_weight, '__getitem__') else edge_weight)
                    out = out + scatter(self.tp(node_attr[src[s:e]], edge_sh[s:e], w), tgt[s:e], dim=0,
                                        dim_size=n_out, reduce='sum')
                    cnt = cnt + torch.bincount(tgt[s:e], minlength=n_out)
                    s = e
                if reduce == 'mean':
                    out = out / torch.clamp(cnt, torch.finfo(dt).eps)[:, None]
            if self.batch_norm:
                out = self.batch_norm(out)
        if self.residual:
            out = out + F.pad(node_attr, (0, out.shape[-1] - node_attr.shape[-1]))
        return out.to(dt)


class OldTensorProductConvLayer(nn.Module):
    """models/tensor_layers.py:338-380: one radial MLP, tensor product + scatter (the reference's 100 000-edge chunking
    shares that MLP, so it only bounds memory), then the residual BEFORE the BatchNorm (the new layer adds it after)."""

    def __init__(self, in_irreps, sh_irreps, out_irreps, n_edge_features, residual=True, batch_norm=True, dropout=0.0,
                 hidden_features=None):
        super().__init__()
        self.in_irreps, self.out_irreps, self.sh_irreps, self.residual = in_irreps, out_irreps, sh_irreps, residual
        hidden_features = n_edge_features if hidden_features is None else hidden_features
        self.tp = o3.FullyConnectedTensorProduct(in_irreps, sh_irreps, out_irreps, shared_weights=False)
        self.fc = nn.Sequential(nn.Linear(n_edge_features, hidden_features), nn.ReLU(), nn.Dropout(dropout),
                                nn.Linear(hidden_features, self.tp.weight_numel))
        self.batch_norm = o3.BatchNorm(out_irreps) if batch_norm else None

    def forward(self, node_attr, edge_index, edge_attr, edge_sh, out_nodes=None, reduce='mean', edge_weight=1.0):
        tgt, src = edge_index[0], edge_index[1]
        n_out = int(out_nodes or node_attr.shape[0])
        w = self.fc(edge_attr) * edge_weight
        summed = scatter(self.tp(node_attr[src], edge_sh, w), tgt, dim=0, dim_size=n_out, reduce='sum')
        if reduce == 'mean':     # tp_scatter_multigroup, :227-229
            cnt = torch.bincount(tgt, minlength=n_out).to(summed.dtype)
            summed = summed / torch.clamp(cnt, torch.finfo(summed.dtype).eps)[:, None]
        out = summed
        if self.residual:
            out = out + F.pad(node_attr, (0, out.shape[-1] - node_attr.shape[-1]))
        if self.batch_norm:
            out = self.batch_norm(out)
        return out.to(node_attr.dtype)
