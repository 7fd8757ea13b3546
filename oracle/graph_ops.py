"""torch_scatter / torch_cluster semantics restated in plain PyTorch.  TEST INFRASTRUCTURE.

torch-scatter 2.1.0 and torch-cluster 1.6.0 (requirements.txt:19,21) are un-vendored native
wheels.  Reference call sites: models/tensor_layers.py:144,220 (scatter sum/mean),
models/cg_model.py:365 (scatter_mean), models/cg_model.py:477 (radius_graph),
models/cg_model.py:543-548,630 (radius).
"""
import torch


def scatter(src, index, dim=0, dim_size=None, reduce='sum'):
    """torch_scatter.scatter for dim=0: 'sum'/'add' = index_add; 'mean' = sum / count.clamp(min=1)."""
    assert dim == 0
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    dim_size = int(dim_size)
    out = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    out.index_add_(0, index, src)
    if reduce in ('sum', 'add'):
        return out
    if reduce == 'mean':
        cnt = torch.bincount(index, minlength=dim_size).clamp(min=1).to(src.dtype)
        return out / cnt.reshape(-1, *([1] * (src.dim() - 1)))
    raise NotImplementedError(reduce)


def scatter_mean(src, index, dim=0, dim_size=None):
    return scatter(src, index, dim=dim, dim_size=dim_size, reduce='mean')


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32):
    """torch_cluster.radius: for every y_j the points x_i of the same batch element with
    ||x_i - y_j||^2 < r^2 (strict; the CUDA kernel compares squared distances), at most
    ``max_num_neighbors`` of them - the first ones in x-index order, as the CUDA kernel keeps.
    Returns [2, E]: row 0 = index into y, row 1 = index into x, sorted by (y, x)."""
    if batch_x is None:
        batch_x = torch.zeros(x.shape[0], dtype=torch.long, device=x.device)
    if batch_y is None:
        batch_y = torch.zeros(y.shape[0], dtype=torch.long, device=y.device)
    rows, cols = [], []
    r2 = float(r) * float(r)
    # blocked over y to bound memory
    step = max(1, 2_000_000 // max(1, x.shape[0]))
    for s in range(0, y.shape[0], step):
        yy = y[s:s + step]
        d = x[None, :, :] - yy[:, None, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]   # fixed evaluation order
        ok = (d2 < r2) & (batch_y[s:s + step, None] == batch_x[None, :])
        if max_num_neighbors < x.shape[0]:
            rank = torch.cumsum(ok.to(torch.int64), dim=1)
            ok = ok & (rank <= max_num_neighbors)
        yi, xi = torch.nonzero(ok, as_tuple=True)
        rows.append(yi + s)
        cols.append(xi)
    if not rows:
        return torch.zeros(2, 0, dtype=torch.long, device=x.device)
    return torch.stack([torch.cat(rows), torch.cat(cols)], 0)


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow='source_to_target'):
    """torch_cluster.radius_graph = radius(x, x, r, batch, batch, cap (+1 if no loops)), rows swapped to
    [neighbour, centre] for flow='source_to_target', self-loops removed."""
    assert flow == 'source_to_target'
    ei = radius(x, x, r, batch, batch, max_num_neighbors if loop else max_num_neighbors + 1)
    row, col = ei[1], ei[0]
    if not loop:
        m = row != col
        row, col = row[m], col[m]
    return torch.stack([row, col], 0)
