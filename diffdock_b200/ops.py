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
        self.fused_pairs, self.fused_bytes, self.fused_flops, self.fused_alg_flops = [], 0, 0, 0

    def summary(self):
        torch.cuda.synchronize()
        return {'launches': len(self.pairs), 'ms': sum(a.elapsed_time(b) for a, b in self.pairs), 'bytes': self.bytes,
                'all_launches': self.all_launches, 'fused_launches': len(self.fused_pairs),
                'fused_ms': sum(a.elapsed_time(b) for a, b in self.fused_pairs), 'fused_bytes': self.fused_bytes,
                'fused_flops': self.fused_flops, 'fused_alg_flops': self.fused_alg_flops}


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


# ---------------------------------------------------------------------------------------------------------------------
# Sync-free graph construction: upper-bound buffers + the live edge count in device memory (no .item())
def radius_count(x, y, x_ptr, y_batch32, r=1.0, r_per_graph=None, max_num_neighbors=32, exclude_self=False):
    """count[j] = neighbours of y_j among the x of its complex (ddb200_radius_count); int32, no host sync."""
    _need_cuda(x, y, x_ptr, y_batch32)
    assert x.dtype == torch.float32 and y.dtype == torch.float32 and x.is_contiguous() and y.is_contiguous()
    assert y_batch32.dtype == torch.int32 and x_ptr.dtype == torch.int32
    n_y = y.shape[0]
    count = torch.empty(n_y, dtype=torch.int32, device=x.device)
    rc = _lib.lib().ddb200_radius_count(_ptr(x), _ptr(y), _ptr(x_ptr), _ptr(y_batch32), _ptr(r_per_graph), float(r), n_y,
                                        int(max_num_neighbors), int(exclude_self), _ptr(count), _stream())
    _lib.check(rc, 'ddb200_radius_count')
    PROFILE.all_launches += 1
    return count


def graph_fill(x, y, x_ptr, y_batch32, row_start, capacity, r=1.0, r_per_graph=None, max_num_neighbors=32,
               exclude_self=False, pre_ptr=None, pre_col=None, want_vec=True, want_eid=False, slot_out=None, slot_in=None,
               y_ptr=None, slot_ld=0, want_perm=False, row_offset=0, col_offset=0, fill_row=None):
    """Fill pass into buffers of ``capacity`` edges (ddb200_graph_fill).  Returns (row, col, vec | None, eid | None,
    perm | None); entries beyond the live count keep their initial value: ``fill_row`` for row (None = uninitialised),
    0 for col / perm, (1, 0, 0) for vec, -1 for eid - valid operands for padded library ops."""
    dev = x.device
    n_y = y.shape[0]
    cap = max(int(capacity), 1)
    init = fill_row is not None
    row = torch.full((cap,), int(fill_row), dtype=torch.int32, device=dev) if init else torch.empty(cap, dtype=torch.int32, device=dev)
    col = torch.zeros(cap, dtype=torch.int32, device=dev) if init else torch.empty(cap, dtype=torch.int32, device=dev)
    vec = eid = perm = None
    if want_vec:
        vec = torch.empty((cap, 3), dtype=torch.float32, device=dev)
        if init:
            vec.zero_()
            vec[:, 0] = 1.0
    if want_eid:
        eid = torch.full((cap,), -1, dtype=torch.int32, device=dev)
    if want_perm:
        perm = torch.zeros(cap, dtype=torch.int32, device=dev) if init else torch.empty(cap, dtype=torch.int32, device=dev)
    rc = _lib.lib().ddb200_graph_fill(_ptr(x), _ptr(y), _ptr(x_ptr), _ptr(y_batch32), _ptr(r_per_graph), float(r), n_y,
                                      int(max_num_neighbors), int(exclude_self), _ptr(row_start), _ptr(pre_ptr),
                                      _ptr(pre_col), _ptr(row), _ptr(col), _ptr(vec), _ptr(eid), _ptr(slot_out),
                                      _ptr(slot_in), _ptr(y_ptr), int(slot_ld), _ptr(perm), int(row_offset),
                                      int(col_offset), _stream())
    _lib.check(rc, 'ddb200_graph_fill')
    PROFILE.all_launches += 1
    return row, col, vec, eid, perm


EDGE_EMBED_SHAPES = {(64, 48), (32, 48), (64, 32), (32, 32), (64, 24), (32, 24), (64, 16), (32, 16), (16, 16), (8, 16),
                     (16, 24), (8, 24)}


def edge_embed(edge_vec, edge_row, u, w1_rbf, w2, b2, rbf_offset, rbf_coeff, n_edges_dev, out=None):
    """out[e] = W2 relu(u[edge_row[e]] + W1_rbf rbf(|edge_vec[e]|)) + b2 for e < *n_edges_dev (ddb200_edge_embed)."""
    _need_cuda(edge_vec, edge_row, u, w1_rbf, w2, b2)
    cap, ns, D = edge_vec.shape[0], u.shape[1], w1_rbf.shape[1]
    assert (D, ns) in EDGE_EMBED_SHAPES
    for t in (edge_vec, u, w1_rbf, w2, b2, rbf_offset):
        assert t.dtype == torch.float32 and t.is_contiguous()
    assert edge_row.dtype == torch.int32 and w1_rbf.shape[0] == ns and tuple(w2.shape) == (ns, ns) and rbf_offset.shape[0] == D
    if out is None:
        out = torch.empty((cap, ns), dtype=torch.float32, device=edge_vec.device)
    rc = _lib.lib().ddb200_edge_embed(_ptr(edge_vec), _ptr(edge_row), _ptr(u), _ptr(w1_rbf), _ptr(w2), _ptr(b2), D, ns,
                                      _ptr(rbf_offset), float(rbf_coeff), cap, _ptr(n_edges_dev), _ptr(out), _stream())
    _lib.check(rc, 'ddb200_edge_embed')
    PROFILE.all_launches += 1
    return out


def pose_update_dev(pos, n_poses, bond_u, bond_v, mask_rotate_u8, tr_score, rot_score, tor_score, coef_table, step_dev=None,
                    tr_z=None, rot_z=None, tor_z=None, seed=0, pose_key=None, use_torsion=True, out=None):
    """ddb200_pose_update_dev: SDE coefficients from a device table row, optional in-kernel Philox noise; ``out`` may be
    ``pos`` itself (in place)."""
    _need_cuda(pos, tr_score, rot_score, coef_table)
    assert pos.dtype == torch.float32 and pos.is_contiguous() and coef_table.dtype == torch.float32 and coef_table.is_contiguous()
    n_atoms = pos.shape[0] // n_poses
    n_bonds = int(bond_u.shape[0]) if bond_u is not None else 0
    f = lambda t: t.float().contiguous() if t is not None else None
    tr_score, rot_score, tor_score, tr_z, rot_z, tor_z = map(f, (tr_score, rot_score, tor_score, tr_z, rot_z, tor_z))
    if out is None:
        out = torch.empty_like(pos)
    if pose_key is not None:
        assert pose_key.dtype == torch.int64 and pose_key.is_cuda and pose_key.shape[0] >= n_poses
    if step_dev is not None:
        assert step_dev.dtype == torch.int32 and step_dev.is_cuda
    rc = _lib.lib().ddb200_pose_update_dev(_ptr(pos), n_poses, n_atoms, n_bonds, _ptr(bond_u), _ptr(bond_v),
                                           _ptr(mask_rotate_u8), _ptr(tr_score), _ptr(rot_score), _ptr(tor_score),
                                           _ptr(tr_z), _ptr(rot_z), _ptr(tor_z), _ptr(coef_table), _ptr(step_dev),
                                           C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), _ptr(pose_key),
                                           1 if use_torsion else 0, _ptr(out), _stream())
    _lib.check(rc, 'ddb200_pose_update_dev')
    PROFILE.all_launches += 1
    return out


def csr_sort_by_target(tgt32, n_rows, want_row_ptr=False):
    """Stable device-side sort of an edge list by target: (tgt_sorted int32, perm int64, row_ptr int32 | None);
    ddb200_csr_sort_by_target with a torch-allocated workspace.  No host synchronisation."""
    _need_cuda(tgt32)
    assert tgt32.dtype == torch.int32 and tgt32.is_contiguous()
    n = tgt32.shape[0]
    dev = tgt32.device
    L = _lib.lib()
    need = C.c_size_t(0)
    _lib.check(L.ddb200_csr_sort_by_target(None, n, int(n_rows), None, None, None, None, C.byref(need), _stream()),
               'ddb200_csr_sort_by_target(size)')
    ws = torch.empty(max(int(need.value), 1), dtype=torch.uint8, device=dev)
    out_t = torch.empty_like(tgt32)
    perm = torch.empty_like(tgt32)
    rp = torch.empty(int(n_rows) + 1, dtype=torch.int32, device=dev) if want_row_ptr else None
    have = C.c_size_t(ws.numel())
    _lib.check(L.ddb200_csr_sort_by_target(_ptr(tgt32), n, int(n_rows), _ptr(out_t), _ptr(perm), _ptr(rp), _ptr(ws),
                                           C.byref(have), _stream()), 'ddb200_csr_sort_by_target')
    PROFILE.all_launches += 2 + (1 if want_row_ptr else 0)
    return out_t, perm.long(), rp
