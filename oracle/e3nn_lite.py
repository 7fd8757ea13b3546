"""Restatement of the e3nn 0.5.x subset used by the DiffDock score model.  TEST INFRASTRUCTURE.

e3nn (requirements.txt:7 ``e3nn==0.5.0``) is an un-vendored third-party dependency of the
reference and is not installable here; this file restates its *published algorithm*:

* ``Irrep`` / ``Irreps``                      - e3nn/o3/_irreps.py
* ``wigner_3j``                               - e3nn/o3/_wigner.py (``_so3_clebsch_gordan``,
                                                ``_su2_clebsch_gordan``, ``change_basis_real_to_complex``)
* ``spherical_harmonics``                     - e3nn/o3/_spherical_harmonics.py (component normalisation)
* ``FullyConnectedTensorProduct`` ('uvw')     - e3nn/o3/_tensor_product/_sub.py + _codegen.py
* ``FullTensorProduct`` ('uvuv')              - same
* ``BatchNorm`` (eval)                        - e3nn/nn/_batchnorm.py

Reference call sites: models/tensor_layers.py:9-11,274-299,307 and models/cg_model.py:3,5,47,
240,411,494,511,556-557,622,636.  Parity of the l=2 CG signs and of ``Irreps.sort`` order is
UNPINNED (see oracle/__init__.py); everything with l<=1 is pinned against the reference's
in-tree ``FasterTensorProduct``.
"""
from __future__ import annotations

import math
from fractions import Fraction
from functools import lru_cache
from math import factorial

import numpy as np
import torch


# ----------------------------------------------------------------------------- irreps
class Irrep(tuple):
    """(l, p) with p=+1 'e' / -1 'o'.  Tuple subclass like e3nn's, so ``ir[0]`` is l."""

    def __new__(cls, l, p=None):
        if p is None:
            if isinstance(l, Irrep):
                return l
            if isinstance(l, str):
                s = l.strip()
                return tuple.__new__(cls, (int(s[:-1]), {'e': 1, 'o': -1}[s[-1]]))
            if isinstance(l, tuple):
                l, p = l
        assert p in (1, -1) and l >= 0
        return tuple.__new__(cls, (int(l), int(p)))

    @property
    def l(self):
        return self[0]

    @property
    def p(self):
        return self[1]

    @property
    def dim(self):
        return 2 * self[0] + 1

    def is_scalar(self):
        return self[0] == 0 and self[1] == 1

    def __repr__(self):
        return f"{self[0]}{'e' if self[1] == 1 else 'o'}"

    __str__ = __repr__

    def __mul__(self, other):
        other = Irrep(other)
        p = self.p * other.p
        return [Irrep(l, p) for l in range(abs(self.l - other.l), self.l + other.l + 1)]


class MulIr(tuple):
    def __new__(cls, mul, ir):
        return tuple.__new__(cls, (int(mul), Irrep(ir)))

    @property
    def mul(self):
        return self[0]

    @property
    def ir(self):
        return self[1]

    @property
    def dim(self):
        return self[0] * self[1].dim

    def __repr__(self):
        return f"{self[0]}x{self[1]}"


class Irreps(tuple):
    def __new__(cls, irreps=None):
        if isinstance(irreps, Irreps):
            return irreps
        out = []
        if isinstance(irreps, Irrep):
            out.append(MulIr(1, irreps))
        elif isinstance(irreps, str):
            for term in irreps.split('+'):
                term = term.strip()
                if not term:
                    continue
                if 'x' in term:
                    m, ir = term.split('x')
                    out.append(MulIr(int(m), Irrep(ir)))
                else:
                    out.append(MulIr(1, Irrep(term)))
        elif irreps is not None:
            for item in irreps:
                if isinstance(item, (str, Irrep)):
                    out.append(MulIr(1, Irrep(item)))
                else:
                    m, ir = item
                    out.append(MulIr(m, Irrep(ir)))
        return tuple.__new__(cls, out)

    @staticmethod
    def spherical_harmonics(lmax, p=-1):
        return Irreps([(1, (l, p ** l)) for l in range(lmax + 1)])

    @property
    def dim(self):
        return sum(mi.dim for mi in self)

    @property
    def num_irreps(self):
        return sum(mi.mul for mi in self)

    def slices(self):
        s, i = [], 0
        for mi in self:
            s.append(slice(i, i + mi.dim))
            i += mi.dim
        return s

    def __contains__(self, ir):
        ir = Irrep(ir)
        return any(mi.ir == ir for mi in self)

    def __add__(self, other):
        return Irreps(list(self) + list(Irreps(other)))

    def __eq__(self, other):
        try:
            return tuple(self) == tuple(Irreps(other))
        except Exception:
            return False

    def __hash__(self):
        return tuple.__hash__(self)

    def __repr__(self):
        return '+'.join(repr(mi) for mi in self)

    def sort(self):
        """e3nn Irreps.sort: stable sort on (ir, original index); returns (irreps, p, inv)."""
        order = sorted(range(len(self)), key=lambda i: (tuple(self[i].ir), i))
        inv = tuple(order)
        p = [0] * len(self)
        for new, old in enumerate(order):
            p[old] = new
        return Irreps([self[i] for i in order]), tuple(p), inv

    def simplify(self):
        out = []
        for mi in self:
            if out and out[-1][1] == mi.ir:
                out[-1] = (out[-1][0] + mi.mul, mi.ir)
            elif mi.mul > 0:
                out.append((mi.mul, mi.ir))
        return Irreps(out)


# ----------------------------------------------------------------------------- Wigner 3j
def _su2_cg_coeff(j1, m1, j2, m2, j3, m3):
    if m3 != m1 + m2:
        return 0.0
    vmin = int(max(-j1 + j2 + m3, -j1 + m1, 0))
    vmax = int(min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3))

    def f(n):
        return factorial(round(n))

    C = ((2.0 * j3 + 1.0) * Fraction(
        f(j3 + j1 - j2) * f(j3 - j1 + j2) * f(j1 + j2 - j3) * f(j3 + m3) * f(j3 - m3),
        f(j1 + j2 + j3 + 1) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2))) ** 0.5
    S = 0
    for v in range(vmin, vmax + 1):
        S += (-1) ** int(v + j2 + m2) * Fraction(
            f(j2 + j3 + m1 - v) * f(j1 - m1 + v),
            f(v) * f(j3 - j1 + j2 - v) * f(j3 + m3 - v) * f(v + j1 - j2 - m3))
    return float(C * S)


def _su2_cg(j1, j2, j3):
    mat = np.zeros((2 * j1 + 1, 2 * j2 + 1, 2 * j3 + 1))
    if abs(j1 - j2) <= j3 <= j1 + j2:
        for m1 in range(-j1, j1 + 1):
            for m2 in range(-j2, j2 + 1):
                if abs(m1 + m2) <= j3:
                    mat[j1 + m1, j2 + m2, j3 + m1 + m2] = _su2_cg_coeff(j1, m1, j2, m2, j3, m1 + m2)
    return mat


def _real_to_complex(l):
    q = np.zeros((2 * l + 1, 2 * l + 1), dtype=np.complex128)
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = 1 / math.sqrt(2)
        q[l + m, l - abs(m)] = -1j / math.sqrt(2)
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m / math.sqrt(2)
        q[l + m, l - abs(m)] = 1j * (-1) ** m / math.sqrt(2)
    return (-1j) ** l * q


@lru_cache(maxsize=None)
def _wigner_3j_np(l1, l2, l3):
    assert abs(l2 - l3) <= l1 <= l2 + l3
    Q1, Q2, Q3 = _real_to_complex(l1), _real_to_complex(l2), _real_to_complex(l3)
    C = _su2_cg(l1, l2, l3).astype(np.complex128)
    C = np.einsum('ij,kl,mn,ikn->jlm', Q1, Q2, np.conj(Q3.T), C)
    assert np.abs(C.imag).max() < 1e-9
    C = C.real
    return C / np.linalg.norm(C)


def wigner_3j(l1, l2, l3, dtype=torch.float64):
    """Real-basis Wigner 3j symbol, Frobenius-normalised, shape [2l1+1, 2l2+1, 2l3+1]."""
    return torch.from_numpy(_wigner_3j_np(int(l1), int(l2), int(l3)).copy()).to(dtype)


# ----------------------------------------------------------------------------- spherical harmonics
def _sh_l(l, x, y, z):
    """Component-normalised real SH polynomial of a UNIT vector (e3nn basis/ordering)."""
    if l == 0:
        return [torch.ones_like(x)]
    if l == 1:
        s = math.sqrt(3.0)
        return [s * x, s * y, s * z]
    if l == 2:
        s3, s5 = math.sqrt(3.0), math.sqrt(5.0)
        return [s5 * s3 * x * z,
                s5 * s3 * x * y,
                s5 * (y * y - 0.5 * (x * x + z * z)),
                s5 * s3 * y * z,
                s5 * (s3 / 2.0) * (z * z - x * x)]
    raise NotImplementedError("oracle SH restated for l<=2 only (the path uses sh_lmax in {1,2} and '2e')")


def spherical_harmonics(l, x, normalize=True, normalization='component'):
    """o3.spherical_harmonics(irreps|str|int|list, x, normalize, normalization='component')."""
    assert normalization == 'component'
    if isinstance(l, int):
        ls = [l]
    elif isinstance(l, (list, tuple)) and not isinstance(l, Irreps) and all(isinstance(v, int) for v in l):
        ls = list(l)
    else:
        irreps = Irreps(l)
        ls = []
        for mi in irreps:
            assert mi.ir.p == (-1) ** mi.ir.l or True
            ls.extend([mi.ir.l] * mi.mul)
    if normalize:
        x = torch.nn.functional.normalize(x, dim=-1)  # x / max(||x||, 1e-12)
    xx, yy, zz = x[..., 0], x[..., 1], x[..., 2]
    out = []
    for li in ls:
        out.extend(_sh_l(li, xx, yy, zz))
    return torch.stack(out, dim=-1)


# ----------------------------------------------------------------------------- tensor products
class Instruction:
    __slots__ = ('i_in1', 'i_in2', 'i_out', 'mode', 'has_weight', 'path_weight', 'path_shape', 'w_offset')

    def __init__(self, i_in1, i_in2, i_out, mode, has_weight, path_weight, path_shape, w_offset):
        self.i_in1, self.i_in2, self.i_out, self.mode = i_in1, i_in2, i_out, mode
        self.has_weight, self.path_weight, self.path_shape, self.w_offset = has_weight, path_weight, path_shape, w_offset


class TensorProduct(torch.nn.Module):
    """e3nn o3.TensorProduct restricted to modes 'uvw' and 'uvuv', external (per-sample) weights,
    irrep_normalization='component', path_normalization='element'."""

    def __init__(self, irreps_in1, irreps_in2, irreps_out, instructions, shared_weights=False,
                 internal_weights=False, **kw):
        super().__init__()
        assert not internal_weights
        self.irreps_in1, self.irreps_in2, self.irreps_out = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        self.shared_weights = shared_weights
        fan = {}
        for (i1, i2, io, mode, hw) in instructions:
            m1, m2 = self.irreps_in1[i1].mul, self.irreps_in2[i2].mul
            x = {'uvw': m1 * m2, 'uvu': m2, 'uvv': m1, 'uuw': m1, 'uuu': 1, 'uvuv': 1}[mode]
            fan[io] = fan.get(io, 0) + x
        self.instructions, off = [], 0
        for (i1, i2, io, mode, hw) in instructions:
            m1, m2, mo = self.irreps_in1[i1].mul, self.irreps_in2[i2].mul, self.irreps_out[io].mul
            alpha = self.irreps_out[io].ir.dim / fan[io]
            shape = {'uvw': (m1, m2, mo), 'uvuv': (m1, m2)}[mode] if hw else ()
            self.instructions.append(Instruction(i1, i2, io, mode, hw, math.sqrt(alpha), shape, off))
            off += int(np.prod(shape)) if hw else 0
        self.weight_numel = off
        # e3nn registers these two buffers on every TensorProduct (they show up in checkpoints)
        self.register_buffer('weight', torch.zeros(0))
        self.register_buffer('output_mask', torch.ones(self.irreps_out.dim))

    def forward(self, x1, x2, weight=None):
        dt = x1.dtype
        lead = x1.shape[:-1]
        x1 = x1.reshape(-1, x1.shape[-1])
        x2 = x2.reshape(-1, x2.shape[-1])
        Z = x1.shape[0]
        if weight is not None:
            weight = weight.reshape(-1, weight.shape[-1])
        s1, s2, so = self.irreps_in1.slices(), self.irreps_in2.slices(), self.irreps_out.slices()
        out = [None] * len(self.irreps_out)
        for ins in self.instructions:
            mi1, mi2, mio = self.irreps_in1[ins.i_in1], self.irreps_in2[ins.i_in2], self.irreps_out[ins.i_out]
            a = x1[:, s1[ins.i_in1]].reshape(Z, mi1.mul, mi1.ir.dim)
            b = x2[:, s2[ins.i_in2]].reshape(Z, mi2.mul, mi2.ir.dim)
            C = wigner_3j(mi1.ir.l, mi2.ir.l, mio.ir.l, dtype=dt).to(x1.device)
            xx = torch.einsum('zui,zvj->zuvij', a, b)
            if ins.mode == 'uvw':
                w = weight[:, ins.w_offset:ins.w_offset + int(np.prod(ins.path_shape))].reshape(Z, *ins.path_shape)
                r = torch.einsum('zuvw,ijk,zuvij->zwk', w, C, xx)
            elif ins.mode == 'uvuv':
                assert not ins.has_weight
                r = torch.einsum('ijk,zuvij->zuvk', C, xx)
            else:
                raise NotImplementedError(ins.mode)
            r = ins.path_weight * r.reshape(Z, mio.dim)
            out[ins.i_out] = r if out[ins.i_out] is None else out[ins.i_out] + r
        out = [o if o is not None else x1.new_zeros(Z, self.irreps_out[i].dim) for i, o in enumerate(out)]
        return torch.cat(out, dim=-1).reshape(*lead, self.irreps_out.dim)


class FullyConnectedTensorProduct(TensorProduct):
    def __init__(self, irreps_in1, irreps_in2, irreps_out, shared_weights=False, **kw):
        i1, i2, io = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        instr = [(a, b, c, 'uvw', True)
                 for a, mia in enumerate(i1) for b, mib in enumerate(i2) for c, mic in enumerate(io)
                 if mic.ir in mia.ir * mib.ir]
        super().__init__(i1, i2, io, instr, shared_weights=shared_weights, **kw)


class FullTensorProduct(TensorProduct):
    def __init__(self, irreps_in1, irreps_in2, filter_ir_out=None, **kw):
        i1, i2 = Irreps(irreps_in1).simplify(), Irreps(irreps_in2).simplify()
        out, instr = [], []
        for a, mia in enumerate(i1):
            for b, mib in enumerate(i2):
                for ir_out in mia.ir * mib.ir:
                    if filter_ir_out is not None and ir_out not in filter_ir_out:
                        continue
                    instr.append((a, b, len(out), 'uvuv', False))
                    out.append((mia.mul * mib.mul, ir_out))
        out = Irreps(out)
        out, p, _ = out.sort()
        instr = [(a, b, p[c], m, hw) for (a, b, c, m, hw) in instr]
        super().__init__(i1, i2, out, instr, **kw)


# ----------------------------------------------------------------------------- batch norm
class BatchNorm(torch.nn.Module):
    """e3nn.nn.BatchNorm (eval-mode arithmetic; training statistics are out of scope)."""

    def __init__(self, irreps, eps=1e-5, momentum=0.1, affine=True, reduce='mean', instance=False,
                 normalization='component'):
        super().__init__()
        self.irreps = Irreps(irreps)
        self.eps = eps
        self.affine = affine
        ns = sum(mi.mul for mi in self.irreps if mi.ir.is_scalar())
        nf = self.irreps.num_irreps
        self.register_buffer('running_mean', torch.zeros(ns))
        self.register_buffer('running_var', torch.ones(nf))
        if affine:
            self.weight = torch.nn.Parameter(torch.ones(nf))
            self.bias = torch.nn.Parameter(torch.zeros(ns))

    def forward(self, x):
        assert not self.training, "oracle BatchNorm restates eval mode only"
        N = x.shape[0]
        out, ix, irm, iw = [], 0, 0, 0
        for mi in self.irreps:
            f = x[:, ix:ix + mi.dim].reshape(N, mi.mul, mi.ir.dim)
            ix += mi.dim
            if mi.ir.is_scalar():
                f = f - self.running_mean[irm:irm + mi.mul].reshape(1, -1, 1)
            nrm = (self.running_var[iw:iw + mi.mul] + self.eps).pow(-0.5)
            if self.affine:
                nrm = nrm * self.weight[iw:iw + mi.mul]
            f = f * nrm.reshape(1, -1, 1)
            if self.affine and mi.ir.is_scalar():
                f = f + self.bias[irm:irm + mi.mul].reshape(1, -1, 1)
            if mi.ir.is_scalar():
                irm += mi.mul
            iw += mi.mul
            out.append(f.reshape(N, -1))
        return torch.cat(out, dim=-1)
