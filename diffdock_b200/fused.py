"""Host side of the fully fused convolution kernel (csrc/fused_conv.cu): per (layer, edge group) plan.
Replaces, per edge group, the reference's edge_attr_ assembly (models/cg_model.py:342-349), the radial FCBlock
(models/layers.py:10-17 at models/tensor_layers.py:140,211) and the tensor product + scatter (models/tensor_layers.py:139-144,204-221).

The fused kernel computes, for a tile of 128 edges, the radial MLP on the tcgen05 tensor cores and contracts the resulting
per-edge tensor-product weights with the edge's irreps *straight out of tensor memory* - the ``[E, weight_numel]`` weight
tensor (11-28 KB per edge) never exists in HBM.  To make that possible the weight columns are cut into N tiles that hold
whole rows ``u`` of one path block ``[mul_in, mul_out]``:

    (mul_out, 2l_out+1) = (48, 1): 4 rows x 48 columns = 192        (10, 3): 16 rows x 10 columns = 160
    (16, 1): 8 x 16 = 128                                            (4, 3): 16 x 4 = 64

so that a consumer thread (one edge = one TMEM lane) knows at compile time which register of its accumulator every
TMEM column feeds.  This module builds, from a ``TpTable`` and the radial MLP's second Linear:
  * the tile table (int32 [T, 8]) and the Clebsch-Gordan term tables of every path,
  * the pre-split, pre-swizzled bf16 operand images of W2 per tile (rows permuted into tile order, zero padded), with the
    bias folded in as two extra K columns (hi, lo) that multiply constant-one columns of the activation operand.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .ops import PROFILE, _need_cuda, _ptr, _stream

from .irreps import real_cg
from .radial import BK, BN
from .tp_table import TpTable

# (mul_out, d_out) -> (consumer kind id, rows per tile)
CONSUMER_KINDS = {(48, 1): (0, 4), (10, 3): (1, 16), (16, 1): (2, 8), (4, 3): (3, 16)}
MAX_K = 144
ENABLED = os.environ.get('DDB200_FUSED_CONV', '1') != '0'


def supported(table: TpTable, hidden: int, k1: int) -> bool:
    if table.sh_lmax < 0:
        return False
    if 3 * hidden + 2 > ((3 * hidden + BK - 1) // BK) * BK or 3 * k1 + 2 > ((3 * k1 + BK - 1) // BK) * BK:
        return False      # no spare K columns for the folded bias
    if (3 * hidden + BK - 1) // BK > 7 or (3 * k1 + BK - 1) // BK > 7:
        return False
    for p in table.paths:
        if (p.mul_out, 2 * p.l_out + 1) not in CONSUMER_KINDS:
            return False
        if (2 * p.l_in + 1) not in (1, 3):
            return False
    return True


def _split_images(w_rows: torch.Tensor, bias_rows: torch.Tensor, K: int):
    """w_rows [T, 256, K] fp32 (zero rows where padded), bias_rows [T, 256] -> bf16 images [T, n_kb, 256, 8, 8]:
    columns [hi | lo | hi | bias_hi, bias_lo | 0...], 128B-swizzled."""
    T = w_rows.shape[0]
    n_kb = (3 * K + BK - 1) // BK
    dev = w_rows.device
    hi = w_rows.to(torch.bfloat16)
    lo = (w_rows - hi.float()).to(torch.bfloat16)
    bhi = bias_rows.to(torch.bfloat16)
    blo = (bias_rows - bhi.float()).to(torch.bfloat16)
    bp = torch.zeros((T, BN, n_kb * BK), dtype=torch.bfloat16, device=dev)
    bp[:, :, :K], bp[:, :, K:2 * K], bp[:, :, 2 * K:3 * K] = hi, lo, hi
    bp[:, :, 3 * K], bp[:, :, 3 * K + 1] = bhi, blo          # x constant-one activation columns
    img = bp.reshape(T, BN, n_kb, 8, 8).permute(0, 2, 1, 3, 4).contiguous()
    rows = torch.arange(BN, device=dev) % 8
    src_chunk = torch.arange(8, device=dev)[None, :] ^ rows[:, None]
    return torch.gather(img, 3, src_chunk[None, None, :, :, None].expand(T, n_kb, BN, 8, 8)).contiguous()


class FusedPlan:
    """Device-resident plan of one (layer, edge group)."""

    def __init__(self, table: TpTable, w1: torch.Tensor, b1: torch.Tensor, w2_ref: torch.Tensor, b2_ref: torch.Tensor):
        """w1 [H, K1], b1 [H]; w2_ref [weight_numel, H], b2_ref [weight_numel] in the REFERENCE weight-row order."""
        dev = w2_ref.device
        H, K1 = w1.shape
        assert supported(table, H, K1)
        self.table, self.hidden, self.k1 = table, H, K1
        paths = sorted(table.paths, key=lambda p: (p.i_out, p.w_ref_off))
        tiles, row_src, mt_off = [], [], {}
        ment_i, term_y, term_v = [], [], []
        for pi, p in enumerate(paths):          # Clebsch-Gordan terms per path: (begin, count) for every (i, k)
            d_in, d_out = 2 * p.l_in + 1, 2 * p.l_out + 1
            C = real_cg(p.l_in, p.l_sh, p.l_out)
            mt_off[id(p)] = len(ment_i) // 2
            for i in range(d_in):
                for k in range(d_out):
                    js = [j for j in range(2 * p.l_sh + 1) if C[i, j, k] != 0.0]
                    ment_i += [len(term_y), len(js)]
                    for j in js:
                        term_y.append(p.sh_off + j)
                        term_v.append(p.coef * C[i, j, k])
        group_prev = None
        path_prev = None
        for p in paths:
            d_in, d_out = 2 * p.l_in + 1, 2 * p.l_out + 1
            kind, rows_per = CONSUMER_KINDS[(p.mul_out, d_out)]
            for u0 in range(0, p.mul_in, rows_per):
                nrow = min(rows_per, p.mul_in - u0)
                n_mma = rows_per * p.mul_out      # always the full tile: no stale TMEM columns are ever read
                first = group_prev != p.i_out
                group_prev = p.i_out
                new_path = path_prev is not p        # flag 4: the consumer rebuilds its C.Y matrix
                path_prev = p
                tiles.append([kind, n_mma, p.in_off + u0 * d_in, nrow, d_in, p.out_off,
                              (1 if first else 0) | (4 if new_path else 0), mt_off[id(p)]])
                src = np.full(BN, -1, dtype=np.int64)
                cols = p.w_ref_off + (u0 * p.mul_out) + np.arange(nrow * p.mul_out)
                src[:nrow * p.mul_out] = cols
                row_src.append(src)
        for i in range(len(tiles)):          # last tile of an accumulator run
            if i == len(tiles) - 1 or tiles[i + 1][6] & 1:
                tiles[i][6] |= 2
        self.n_tiles = len(tiles)
        src = torch.as_tensor(np.stack(row_src), device=dev)                      # [T, 256] -> reference weight row or -1
        ok = src >= 0
        w_rows = torch.zeros((self.n_tiles, BN, H), dtype=torch.float32, device=dev)
        b_rows = torch.zeros((self.n_tiles, BN), dtype=torch.float32, device=dev)
        w_rows[ok] = w2_ref.detach().float()[src[ok]]
        b_rows[ok] = b2_ref.detach().float()[src[ok]]
        self.w2_images = _split_images(w_rows, b_rows, H)
        w1p = torch.zeros((1, BN, K1), dtype=torch.float32, device=dev)
        b1p = torch.zeros((1, BN), dtype=torch.float32, device=dev)
        w1p[0, :H], b1p[0, :H] = w1.detach().float(), b1.detach().float()
        self.w1_images = _split_images(w1p, b1p, K1)
        self.tiles = torch.as_tensor(np.asarray(tiles, dtype=np.int32), device=dev).contiguous()
        self.ment = torch.as_tensor(np.asarray(ment_i, dtype=np.int32), device=dev).contiguous()
        self.term_y = torch.as_tensor(np.asarray(term_y if term_y else [0], dtype=np.int32), device=dev).contiguous()
        self.term_v = torch.as_tensor(np.asarray(term_v if term_v else [0.0], dtype=np.float32), device=dev).contiguous()
        self.n_ment = len(ment_i) // 2
        self.n_terms = len(term_y)
        # bf16 tensor-core FLOPs issued per 128-edge tile (split-bf16 x3, K padded to 64s, full-width N tiles)
        n_kb, n_kb1 = (3 * H + BK - 1) // BK, (3 * K1 + BK - 1) // BK
        n1 = (H + 15) // 16 * 16
        self.mma_flops_per_tile = 2 * 128 * (n1 * n_kb1 * BK + sum(t[1] for t in tiles) * n_kb * BK)


def fused_conv(plan: FusedPlan, edge_attr, node, ns, tgt32, src32, x, edge_vec, sum_buf, cnt_buf, edge_weight=None):
    """sum_buf[tgt] += TP(x[src], Y(vec), radial_mlp(...)) for one CSR-sorted edge group, in ONE kernel."""
    _need_cuda(edge_attr, x, edge_vec, sum_buf)
    E, ne = edge_attr.shape
    if E == 0:
        return
    t = plan.table
    assert edge_attr.dtype == torch.float32 and edge_attr.stride(1) == 1 and x.stride(1) == 1 and edge_vec.is_contiguous()
    assert tgt32.dtype == torch.int32 and src32.dtype == torch.int32 and tgt32.is_contiguous() and src32.is_contiguous()
    assert ne + 2 * ns == plan.k1 and x.shape[1] == t.d_in and sum_buf.shape[1] == t.d_out and sum_buf.is_contiguous()
    if edge_weight is not None:
        edge_weight = edge_weight.reshape(-1).contiguous().float()
    prof = PROFILE.enabled
    if prof:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = _lib.lib().ddb200_fused_conv(_ptr(edge_attr), edge_attr.stride(0), ne, _ptr(node) if ns else C.c_void_p(0),
                                      node.stride(0) if ns else 0, ns, _ptr(tgt32), _ptr(src32), _ptr(plan.w1_images),
                                      plan.hidden, _ptr(plan.w2_images), _ptr(plan.tiles), plan.n_tiles, _ptr(plan.ment),
                                      plan.n_ment, _ptr(plan.term_y), _ptr(plan.term_v), plan.n_terms, _ptr(x),
                                      x.stride(0), _ptr(edge_vec), _ptr(edge_weight), t.sh_lmax, E, _ptr(sum_buf),
                                      t.d_out, _ptr(cnt_buf), _stream())
    if prof:
        e1.record()
        PROFILE.fused_pairs.append((e0, e1))
        PROFILE.fused_bytes += E * (4 * t.weight_numel + 12 + 4) + 4 * (sum_buf.shape[0] + 1) + \
            4 * x.shape[0] * t.d_in + 4 * sum_buf.shape[0] * t.d_out
        PROFILE.fused_flops += ((E + 127) // 128) * plan.mma_flops_per_tile
    PROFILE.all_launches += 1
    _lib.check(rc, 'ddb200_fused_conv')
