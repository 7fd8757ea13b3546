"""Tensor-level wrappers over the C ABI (device pointers + current CUDA stream).  CUDA tensors only.

What each wrapper stands in for in the reference (details per entry point in include/diffdock_b200.h):
  tpconv_accumulate / tpconv_finalize   gather + o3.spherical_harmonics + tensor product + torch_scatter.scatter + bincount and
                                        the mean / BatchNorm / residual epilogue, models/tensor_layers.py:139-144,204-229,327-332
  radius                                torch_cluster.radius / radius_graph, models/cg_model.py:477,543-548,630
  segment_ptr                           the CSR row pointer of a sorted ``batch`` vector (PyG ``Batch.ptr``)
  pose_update                           utils/sampling.py:133-191 + utils/diffusion_utils.py:60-78 (modify_conformer_batch) +
                                        utils/torsion.py:75-90 + utils/geometry.py:246-276 (Kabsch)"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .tp_table import TpTable


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("diffdock_b200 ops run on CUDA tensors only (no CPU fallback)")


class _Profile:
    """Optional live accounting used by bench.py: CUDA-event pairs around every tensor-product conv launch (on the
    launching stream) with that launch's ALGORITHMIC bytes, and a count of all kernels
    launched through this module."""

    def __init__(self):
        self.reset(False)

    def reset(self, enabled=False):
        self.enabled, self.pairs, self.bytes, self.all_launches = enabled, [], 0, 0
        self.fused_pairs, self.fused_bytes, self.fused_flops = [], 0, 0

    def summary(self):
        torch.cuda.synchronize()
        return {'launches': len(self.pairs), 'ms': sum(a.elapsed_time(b) for a, b in self.pairs), 'bytes': self.bytes,
                'all_launches': self.all_launches, 'fused_launches': len(self.fused_pairs),
                'fused_ms': sum(a.elapsed_time(b) for a, b in self.fused_pairs), 'fused_bytes': self.fused_bytes,
                'fused_flops': self.fused_flops}


PROFILE = _Profile()


class TpHandle:
    """Device-resident tensor-product table (ddb200_tp_table)."""

    def __init__(self, table: TpTable):
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA device required")
        self.table = table
        h = C.c_void_p()
        ib = np.ascontiguousarray(table.iblob, dtype=np.int32)
        fb = np.ascontiguousarray(table.fblob, dtype=np.float32)
        rc = _lib.lib().ddb200_tp_table_create(ib.ctypes.data_as(C.c_void_p), len(ib), fb.ctypes.data_as(C.c_void_p),
                                               len(fb), C.byref(h))
        _lib.check(rc, 'ddb200_tp_table_create')
        self._h = h

    def __deepcopy__(self, memo):
        return TpHandle(self.table)

    def info(self, what):
        return _lib.lib().ddb200_tp_table_info(self._h, what)

    def __del__(self):
        try:
            if getattr(self, '_h', None):
                _lib.lib().ddb200_tp_table_destroy(self._h)
        except Exception:
            pass


def tpconv_accumulate(h: TpHandle, x, edge_src, edge_dst, geo, w, sum_buf, cnt_buf=None, edge_weight=None,
                      count_node_bytes=True):
    """sum_buf[edge_dst[e]] += TP(x[edge_src[e]], Y(geo[e]), w[e]) (* edge_weight[e]);  cnt_buf[edge_dst[e]] += 1."""
    _need_cuda(x, edge_src, edge_dst, geo, w, sum_buf)
    E = edge_src.shape[0]
    if E == 0:
        return
    t = h.table
    assert x.dtype == torch.float32 and w.dtype == torch.float32 and geo.dtype == torch.float32
    assert edge_src.dtype == torch.int32 and edge_dst.dtype == torch.int32
    assert x.stride(1) == 1 and w.stride(1) == 1 and geo.is_contiguous() and sum_buf.is_contiguous()
    assert edge_src.is_contiguous() and edge_dst.is_contiguous()
    assert x.shape[1] == t.d_in and w.shape[1] >= t.weight_numel_padded and sum_buf.shape[1] == t.d_out
    assert geo.shape[0] == E and w.shape[0] == E and geo.shape[1] == (3 if t.sh_lmax >= 0 else t.d_sh)
    if edge_weight is not None:
        edge_weight = edge_weight.reshape(-1).contiguous().float()
        assert edge_weight.shape[0] == E
    prof = PROFILE.enabled
    if prof:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = _lib.lib().ddb200_tpconv_accumulate(h._h, _ptr(x), x.stride(0), _ptr(edge_src), _ptr(edge_dst), _ptr(geo),
                                             _ptr(edge_weight), _ptr(w), w.stride(0), E, _ptr(sum_buf), _ptr(cnt_buf),
                                             _stream())
    if prof:
        e1.record()
        PROFILE.pairs.append((e0, e1))
        PROFILE.bytes += E * (4 * t.weight_numel + 12 + 4) + (4 * E if edge_weight is not None else 0)
        if count_node_bytes:   # node tensors are compulsory traffic once per (layer, edge set), not per edge block
            PROFILE.bytes += 4 * (sum_buf.shape[0] + 1) + 4 * x.shape[0] * t.d_in + 4 * sum_buf.shape[0] * t.d_out
    PROFILE.all_launches += 1
    _lib.check(rc, 'ddb200_tpconv_accumulate')


def tpconv_finalize(sum_buf, cnt_buf, mean, bn_scale=None, bn_shift=None, residual=None, out=None):
    _need_cuda(sum_buf)
    n, d = sum_buf.shape
    if out is None:
        out = torch.empty_like(sum_buf)
    res_stride = residual.stride(0) if residual is not None else 0
    res_dim = residual.shape[1] if residual is not None else 0
    if residual is not None:
        assert residual.stride(1) == 1 and residual.shape[0] == n and res_dim <= d
    rc = _lib.lib().ddb200_tpconv_finalize(_ptr(sum_buf), _ptr(cnt_buf), n, d, 1 if mean else 0, _ptr(bn_scale),
                                           _ptr(bn_shift), _ptr(residual), res_stride, res_dim, _ptr(out), _stream())
    _lib.check(rc, 'ddb200_tpconv_finalize')
    PROFILE.all_launches += 1
    return out


def segment_ptr(batch, num_graphs):
    """CSR offsets [B+1] (int32) of a sorted batch vector."""
    counts = torch.bincount(batch, minlength=num_graphs)
    ptr = torch.zeros(num_graphs + 1, dtype=torch.int32, device=batch.device)
    ptr[1:] = torch.cumsum(counts, 0)
    return ptr


def radius(x, y, x_ptr, y_batch, r=1.0, r_per_graph=None, max_num_neighbors=32, exclude_self=False):
    """Neighbour pairs (row = index into y, col = index into x), int32, sorted by (row, col).
    Semantics of torch_cluster.radius(x, y, r, batch_x, batch_y, max_num_neighbors)."""
    _need_cuda(x, y, x_ptr, y_batch)
    x, y = x.float().contiguous(), y.float().contiguous()
    yb = y_batch.to(torch.int32).contiguous()
    n_y = y.shape[0]
    if r_per_graph is not None:
        r_per_graph = r_per_graph.reshape(-1).float().contiguous()
    count = torch.empty(n_y, dtype=torch.int32, device=x.device)
    L = _lib.lib()
    rc = L.ddb200_radius_count(_ptr(x), _ptr(y), _ptr(x_ptr), _ptr(yb), _ptr(r_per_graph), float(r), n_y,
                               int(max_num_neighbors), int(exclude_self), _ptr(count), _stream())
    _lib.check(rc, 'ddb200_radius_count')
    incl = torch.cumsum(count, 0, dtype=torch.int32)
    row_start = (incl - count).contiguous()
    n_edges = int(incl[-1].item()) if n_y else 0      # host sync: the edge count sizes the output buffers
    row = torch.empty(n_edges, dtype=torch.int32, device=x.device)
    col = torch.empty(n_edges, dtype=torch.int32, device=x.device)
    if n_edges:
        rc = L.ddb200_radius_fill(_ptr(x), _ptr(y), _ptr(x_ptr), _ptr(yb), _ptr(r_per_graph), float(r), n_y,
                                  int(max_num_neighbors), int(exclude_self), _ptr(row_start), _ptr(row), _ptr(col),
                                  _stream())
        _lib.check(rc, 'ddb200_radius_fill')
    PROFILE.all_launches += 2
    return row, col, count


def pose_update(pos, n_poses, bond_u, bond_v, mask_rotate_u8, tr_score, rot_score, tor_score, coef, tr_z=None,
                rot_z=None, tor_z=None, use_torsion=True):
    """New ligand coordinates [n_poses * n_atoms, 3] after one reverse-diffusion step (ddb200_pose_update)."""
    _need_cuda(pos, tr_score, rot_score)
    pos = pos.float().contiguous()
    n_atoms = pos.shape[0] // n_poses
    n_bonds = int(bond_u.shape[0]) if bond_u is not None else 0
    f = lambda t: t.float().contiguous() if t is not None else None
    tr_score, rot_score, tor_score, tr_z, rot_z, tor_z = map(f, (tr_score, rot_score, tor_score, tr_z, rot_z, tor_z))
    out = torch.empty_like(pos)
    c = (C.c_float * 6)(*[float(v) for v in coef])
    rc = _lib.lib().ddb200_pose_update(_ptr(pos), n_poses, n_atoms, n_bonds, _ptr(bond_u), _ptr(bond_v),
                                       _ptr(mask_rotate_u8), _ptr(tr_score), _ptr(rot_score), _ptr(tor_score),
                                       _ptr(tr_z), _ptr(rot_z), _ptr(tor_z), C.cast(c, C.c_void_p),
                                       1 if use_torsion else 0, _ptr(out), _stream())
    _lib.check(rc, 'ddb200_pose_update')
    PROFILE.all_launches += 1
    return out
