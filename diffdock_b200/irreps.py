"""Irreps bookkeeping and real-basis Clebsch-Gordan blocks for the tensor-product table generator.

Product-side implementation (independent of ``oracle/``): parses e3nn-style irreps strings such as
``'48x0e + 10x1o'`` (models/tensor_layers.py:17-32) and builds the real Wigner-3j blocks
``C[l1,l2,l3]`` in the convention of e3nn 0.5 (Frobenius norm 1, real basis in which the l=1 irrep is the
Cartesian (x, y, z) vector), which is the convention the reference's weights are trained in
(models/tensor_layers.py:299, models/cg_model.py:240).

Method: complex SU(2) Clebsch-Gordan coefficients from the Racah closed form, conjugated into the real basis
with the unitary Q_l below, normalised.  Everything is float64 numpy; results are cached.
"""
from __future__ import annotations

import math
import re
from functools import lru_cache
from typing import List, Tuple

import numpy as np

_TERM = re.compile(r'^\s*(?:(\d+)\s*x\s*)?(\d+)\s*([eo])\s*$')


def parse_irreps(spec) -> List[Tuple[int, int, int]]:
    """-> list of (mul, l, parity) with parity +1 ('e') or -1 ('o').  Accepts a string, or any iterable of
    (mul, (l, p)) pairs (e.g. an e3nn Irreps object)."""
    if isinstance(spec, str):
        out = []
        for term in spec.split('+'):
            m = _TERM.match(term)
            if not m:
                raise ValueError(f"bad irreps term {term!r}")
            out.append((int(m.group(1) or 1), int(m.group(2)), 1 if m.group(3) == 'e' else -1))
        return out
    out = []
    for item in spec:
        mul, ir = item
        l, p = ir
        out.append((int(mul), int(l), int(p)))
    return out


def irreps_dim(irreps) -> int:
    return sum(m * (2 * l + 1) for m, l, _ in irreps)


def irreps_offsets(irreps) -> List[int]:
    offs, o = [], 0
    for m, l, _ in irreps:
        offs.append(o)
        o += m * (2 * l + 1)
    return offs


def sh_irreps(lmax) -> List[Tuple[int, int, int]]:
    return [(1, l, (-1) ** l) for l in range(lmax + 1)]


def irreps_str(irreps) -> str:
    return ' + '.join(f"{m}x{l}{'e' if p == 1 else 'o'}" for m, l, p in irreps)


# ------------------------------------------------------------------------------------------------
def _fact(n: int) -> int:
    return math.factorial(n)


def _cg_complex(j1: int, j2: int, j3: int) -> np.ndarray:
    """<j1 m1 j2 m2 | j3 m3> for integer spins, indexed [j1+m1, j2+m2, j3+m3] (Racah)."""
    out = np.zeros((2 * j1 + 1, 2 * j2 + 1, 2 * j3 + 1))
    pref0 = (2 * j3 + 1) * _fact(j3 + j1 - j2) * _fact(j3 - j1 + j2) * _fact(j1 + j2 - j3) / _fact(j1 + j2 + j3 + 1)
    for m1 in range(-j1, j1 + 1):
        for m2 in range(-j2, j2 + 1):
            m3 = m1 + m2
            if abs(m3) > j3:
                continue
            pref = pref0 * _fact(j3 + m3) * _fact(j3 - m3) / (
                _fact(j1 - m1) * _fact(j1 + m1) * _fact(j2 - m2) * _fact(j2 + m2))
            s = 0.0
            for v in range(max(-j1 + j2 + m3, -j1 + m1, 0), min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3) + 1):
                s += (-1) ** (v + j2 + m2) * _fact(j2 + j3 + m1 - v) * _fact(j1 - m1 + v) / (
                    _fact(v) * _fact(j3 - j1 + j2 - v) * _fact(j3 + m3 - v) * _fact(v + j1 - j2 - m3))
            out[j1 + m1, j2 + m2, j3 + m3] = math.sqrt(pref) * s
    return out


def _q_real(l: int) -> np.ndarray:
    """Unitary taking real-basis components to complex (m = -l..l) components."""
    q = np.zeros((2 * l + 1, 2 * l + 1), dtype=np.complex128)
    r2 = 1.0 / math.sqrt(2.0)
    for m in range(1, l + 1):
        q[l - m, l + m] = r2
        q[l - m, l - m] = -1j * r2
        q[l + m, l + m] = (-1) ** m * r2
        q[l + m, l - m] = 1j * (-1) ** m * r2
    q[l, l] = 1.0
    return (-1j) ** l * q


@lru_cache(maxsize=None)
def real_cg(l1: int, l2: int, l3: int) -> np.ndarray:
    """Real Wigner-3j block [2l1+1, 2l2+1, 2l3+1], Frobenius norm 1."""
    if not (abs(l1 - l2) <= l3 <= l1 + l2):
        raise ValueError("triangle rule violated")
    c = np.einsum('ia,kb,nc,ikn->abc', _q_real(l1), _q_real(l2), np.conj(_q_real(l3)), _cg_complex(l1, l2, l3))
    if np.abs(c.imag).max() > 1e-9:
        raise AssertionError("CG block is not real")
    c = c.real
    c = np.where(np.abs(c) < 1e-14, 0.0, c)
    return c / np.linalg.norm(c)
