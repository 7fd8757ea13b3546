"""CPU self-consistency of the restated third-party semantics (no reference fixtures exist for these)."""
import math

import numpy as np
import pytest
import torch

from oracle import e3nn_lite as o3
from oracle.graph_ops import radius, radius_graph, scatter


def _rot(seed):
    g = torch.Generator().manual_seed(seed)
    q, r = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))
    q = q * torch.sign(torch.diagonal(r))
    return q * torch.linalg.det(q)


def _D(l, R):
    """Representation matrix of rotation R on the real l-irrep, recovered from the spherical harmonics themselves."""
    g = torch.Generator().manual_seed(99)
    v = torch.randn(40, 3, generator=g, dtype=torch.float64)
    Y, YR = o3.spherical_harmonics(l, v), o3.spherical_harmonics(l, v @ R.T)
    return torch.linalg.lstsq(Y, YR).solution.T      # YR = Y @ D^T


def test_closed_forms():
    s3, s6 = math.sqrt(3), math.sqrt(6)
    assert torch.allclose(o3.wigner_3j(1, 1, 0)[:, :, 0], torch.eye(3, dtype=torch.float64) / s3)
    eps = torch.zeros(3, 3, 3, dtype=torch.float64)
    for (i, j, k), s in {(0, 1, 2): 1, (1, 2, 0): 1, (2, 0, 1): 1, (0, 2, 1): -1, (2, 1, 0): -1, (1, 0, 2): -1}.items():
        eps[i, j, k] = s
    assert torch.allclose(o3.wigner_3j(1, 1, 1), eps / s6)
    for l in (1, 2, 3):
        assert torch.allclose(o3.wigner_3j(l, l, 0)[:, :, 0], torch.eye(2 * l + 1, dtype=torch.float64) / math.sqrt(2 * l + 1))
        assert torch.allclose(o3.wigner_3j(0, l, l)[0], torch.eye(2 * l + 1, dtype=torch.float64) / math.sqrt(2 * l + 1))


@pytest.mark.parametrize("ls", [(2, 2, 4), (1, 2, 3), (2, 2, 3)])
def test_high_l_blocks_induce_representations(ls):
    """l=3,4 blocks (torsion head, FullTensorProduct(sh, 2e)): D_l3 := (2 l3+1) C^T (D_l1 x D_l2) C must be an orthogonal
    matrix and a group homomorphism - true iff C intertwines l1 x l2 with a (2 l3+1)-dimensional irrep."""
    l1, l2, l3 = ls
    C = o3.wigner_3j(l1, l2, l3)

    def D3(R):
        return (2 * l3 + 1) * torch.einsum('ijk,ia,jb,abc->kc', C, _D(l1, R), _D(l2, R), C)

    Ra, Rb = _rot(1), _rot(2)
    I = torch.eye(2 * l3 + 1, dtype=torch.float64)
    assert torch.allclose(D3(Ra) @ D3(Ra).T, I, atol=1e-10)
    assert torch.allclose(D3(Ra) @ D3(Rb), D3(Ra @ Rb), atol=1e-10)


@pytest.mark.parametrize("ls", [(1, 1, 2), (1, 2, 1), (2, 2, 2), (2, 1, 1), (1, 2, 2), (2, 2, 1), (1, 1, 1), (0, 2, 2)])
def test_wigner_blocks_are_invariant(ls):
    R = _rot(3)
    l1, l2, l3 = ls
    C = o3.wigner_3j(l1, l2, l3)
    D1, D2, D3 = _D(l1, R), _D(l2, R), _D(l3, R)
    assert torch.allclose(torch.einsum('ijk,ia,jb,kc->abc', C, D1, D2, D3), C, atol=1e-10)
    assert abs(float(C.norm()) - 1) < 1e-12


def test_sh_component_normalisation_and_zero_vector():
    v = torch.randn(1000, 3, dtype=torch.float64)
    Y = o3.spherical_harmonics([0, 1, 2], v)
    assert torch.allclose((Y[:, 1:4] ** 2).sum(-1), torch.full((1000,), 3.0, dtype=torch.float64))
    assert torch.allclose((Y[:, 4:9] ** 2).sum(-1), torch.full((1000,), 5.0, dtype=torch.float64))
    Z = o3.spherical_harmonics([0, 1, 2], torch.zeros(1, 3))
    assert Z[0, 0] == 1 and torch.all(Z[0, 1:4] == 0)


def test_fctp_equivariance_and_bn_residual_layer():
    from oracle.tensor_layers import TensorProductConvLayer, get_irrep_seq
    torch.manual_seed(0)
    seq = get_irrep_seq(6, 3, False, False)
    layer = TensorProductConvLayer(seq[3], '1x0e+1x1o+1x2e', seq[3], 12, hidden_features=12).double().eval()
    N, E = 8, 50
    x = torch.randn(N, 30, dtype=torch.float64)
    ei = torch.randint(0, N, (2, E))
    vec = torch.randn(E, 3, dtype=torch.float64)
    ea = torch.randn(E, 12, dtype=torch.float64)
    R = _rot(5)

    def rotate_feats(f):  # irreps 6x0e + 3x1o + 3x1e + 6x0o ; proper rotation acts as R on both 1o and 1e
        out = f.clone()
        out[:, 6:15] = (f[:, 6:15].reshape(-1, 3, 3) @ R.T).reshape(-1, 9)
        out[:, 15:24] = (f[:, 15:24].reshape(-1, 3, 3) @ R.T).reshape(-1, 9)
        return out

    sh = lambda v: o3.spherical_harmonics(o3.Irreps.spherical_harmonics(2), v)
    with torch.no_grad():
        a = rotate_feats(layer(x, ei, ea, sh(vec)))
        b = layer(rotate_feats(x), ei, ea, sh(vec @ R.T))
    assert torch.allclose(a, b, atol=1e-10)
    # invariance to edge permutation
    perm = torch.randperm(E)
    with torch.no_grad():
        c = layer(x, ei[:, perm], ea[perm], sh(vec)[perm])
    assert torch.allclose(layer(x, ei, ea, sh(vec)), c, atol=1e-10)


def test_scatter_and_radius_semantics():
    src = torch.tensor([[1.0], [2.0], [4.0]])
    idx = torch.tensor([0, 0, 2])
    assert scatter(src, idx, 0, 4, 'sum').flatten().tolist() == [3, 0, 4, 0]
    assert scatter(src, idx, 0, 4, 'mean').flatten().tolist() == [1.5, 0, 4, 0]     # empty segment -> 0
    x = torch.tensor([[0., 0, 0], [1, 0, 0], [2, 0, 0], [0, 0, 0.5]])
    bx = torch.tensor([0, 0, 0, 1])
    y = torch.tensor([[0.1, 0, 0], [0, 0, 0.4]])
    by = torch.tensor([0, 1])
    assert radius(x, y, 1.0, bx, by).tolist() == [[0, 0, 1], [0, 1, 3]]           # strict <, same batch only
    assert radius(x, y, 1.0, bx, by, max_num_neighbors=1).tolist() == [[0, 1], [0, 3]]
    rg = radius_graph(x[:3], 1.5, torch.zeros(3, dtype=torch.long))
    assert sorted(map(tuple, rg.T.tolist())) == [(0, 1), (1, 0), (1, 2), (2, 1)]     # [neighbour, centre], no loops


def test_batchnorm_eval_formula():
    bn = o3.BatchNorm('2x0e + 1x1o + 1x0o').eval()
    with torch.no_grad():
        bn.running_mean.copy_(torch.tensor([1.0, 2.0]))
        bn.running_var.copy_(torch.tensor([4.0, 9.0, 16.0, 25.0]))
        bn.weight.copy_(torch.tensor([1.0, 2.0, 3.0, 4.0]))
        bn.bias.copy_(torch.tensor([0.5, -0.5]))
    x = torch.ones(1, 6)
    y = bn(x)[0]
    e = 1e-5
    exp = [(1 - 1) / math.sqrt(4 + e) * 1 + 0.5, (1 - 2) / math.sqrt(9 + e) * 2 - 0.5] + [3 / math.sqrt(16 + e)] * 3 + \
          [4 / math.sqrt(25 + e)]    # 0o is NOT treated as a scalar: no mean, no bias
    assert torch.allclose(y, torch.tensor(exp), atol=1e-6)
