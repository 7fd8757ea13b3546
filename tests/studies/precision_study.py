#!/usr/bin/env python
"""CPU study: score error of the radial-MLP operand formats the fused kernel could use on the tensor cores.

The oracle CGModel (fp32) is evaluated once as it is and once per scheme with every Linear of the convolution layers'
radial MLPs (models/tensor_layers.py:140,211 FCBlock) replaced by an emulation of the tensor-core arithmetic:
operands quantised as the kernel would, products and sums in fp64 (the MMA accumulates in fp32; its rounding is the same for
all schemes and is not what distinguishes them).

    bf16x3      x_hi.W_hi + x_hi.W_lo + x_lo.W_hi, all parts bf16             (round-2 kernel, 3 f16-rate products)
    f16x1       fp16(x).fp16(W)                                               (1 product)
    f16+fp8     fp16(x).fp16(W) + e4m3(x_lo).e4m3(W) + e4m3(x).e4m3(W_lo)     (1 f16-rate + 2 fp8-rate products = 2 units)
                with the power-of-two scales of diffdock_b200/fused.py
Prints max|a-b| / max|b| of tr / rot / tor scores versus the plain oracle.  Runs in a minute on CPU.  Test infrastructure (it
evaluates the oracle, so it lives under tests/): python tests/studies/precision_study.py
"""
import argparse
import os
import sys

import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def _e4m3(t):
    return t.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).to(torch.float64)


def _pow2_ceil_exp(v):
    import math
    return int(math.ceil(math.log2(max(float(v), 1e-30))))


class EmuLinear(nn.Module):
    def __init__(self, lin: nn.Linear, scheme: str):
        super().__init__()
        self.lin, self.scheme = lin, scheme

    def forward(self, x):
        W, b = self.lin.weight.detach(), self.lin.bias.detach()
        s = self.scheme
        xd, Wd = x.double(), W.double()
        if s == 'bf16x3':
            xh = x.to(torch.bfloat16).double(); xl = (xd - xh).float().to(torch.bfloat16).double()
            Wh = W.to(torch.bfloat16).double(); Wl = (Wd - Wh).float().to(torch.bfloat16).double()
            y = xh @ Wh.T + xh @ Wl.T + xl @ Wh.T
        elif s == 'f16x1':
            y = x.half().double() @ W.half().double().T
        elif s == 'tf32':
            def tf(t):
                return (t.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32).double()
            y = tf(x.contiguous()) @ tf(W.contiguous()).T
        elif s.startswith('f16+fp8'):
            m = _pow2_ceil_exp(max(W.abs().max(), b.abs().max()))
            k = 15 - m                          # main scale: |W| 2^k <= 2^15 (fp16 normal range)
            ka, kb = 8, 19 - m                  # x_lo 2^ka ~ |x| 2^-4 ; W_lo 2^kb <= 2^7
            sb, sa = k - ka, k - kb             # partner scales so that every product carries 2^k
            xh = x.half().double()
            xl8 = _e4m3(((xd - xh) * 2.0 ** ka).float())
            x8 = _e4m3((xd * 2.0 ** sa).float())
            Wh = (W * 2.0 ** k).half().double()
            W8 = _e4m3((Wd * 2.0 ** sb).float())
            Wl8 = _e4m3(((Wd * 2.0 ** k - Wh) * 2.0 ** (kb - k)).float())
            y = (xh @ Wh.T + xl8 @ W8.T + x8 @ Wl8.T) * 2.0 ** (-k)
        else:
            raise ValueError(s)
        return (y + b.double()).float()


def patch(model, scheme):
    n = 0
    for mod in model.modules():
        if mod.__class__.__name__ in ('TensorProductConvLayer', 'OldTensorProductConvLayer'):
            fcs = mod.fc if isinstance(mod.fc, nn.ModuleList) else [mod.fc]
            for fc in fcs:
                for i, sub in enumerate(fc):
                    if isinstance(sub, nn.Linear):
                        fc[i] = EmuLinear(sub, scheme)
                        n += 1
                    elif isinstance(sub, EmuLinear):
                        sub.scheme = scheme
                        n += 1
    return n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ns', type=int, default=48)
    ap.add_argument('--nv', type=int, default=10)
    ap.add_argument('--layers', type=int, default=6)
    ap.add_argument('--res', type=int, default=120)
    ap.add_argument('--atoms', type=int, default=18)
    ap.add_argument('--poses', type=int, default=2)
    ap.add_argument('--seeds', type=int, default=2)
    ap.add_argument('--wscale', type=float, default=1.0, help='multiply the radial-MLP weights (range robustness)')
    a = ap.parse_args()
    from tests.parity_helpers import make_model_pair, rel_err
    from diffdock_b200.synthetic import default_model_args, make_pose_list
    from diffdock_b200.hetero import collate
    from oracle.diffusion import set_time
    for seed in range(a.seeds):
        for t in (1.0, 0.5, 0.05):
            args = default_model_args(ns=a.ns, nv=a.nv, sh_lmax=2, num_conv_layers=a.layers, distance_embed_dim=64,
                                      cross_distance_embed_dim=64, sigma_embed_dim=64)
            o, _ = make_model_pair(args, seed, product=False)
            if a.wscale != 1.0:
                with torch.no_grad():
                    for mod in o.modules():
                        if mod.__class__.__name__ == 'TensorProductConvLayer':
                            for prm in mod.fc.parameters():
                                prm.mul_(a.wscale)
            poses = make_pose_list(a.poses, n_res=a.res, n_atoms=a.atoms, seed=seed + 3, tr_sigma_max=args.tr_sigma_max * t)

            def run():
                g = collate(poses)
                set_time(g, t, t, t, a.poses, 'cpu')
                with torch.no_grad():
                    return o(g)
            ref = run()
            for scheme in ('bf16x3', 'f16+fp8', 'f16x1', 'tf32'):
                patch(o, scheme)
                got = run()
                print(f'seed {seed} t {t:4.2f} {scheme:8s} tr {rel_err(got[0], ref[0]):.2e} rot {rel_err(got[1], ref[1]):.2e} '
                      f'tor {rel_err(got[2], ref[2]):.2e}', flush=True)


if __name__ == '__main__':
    main()
