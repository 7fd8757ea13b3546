#!/usr/bin/env python
"""Micro-benchmark of the fused tensor-product conv kernel alone (BASELINE config 4: TP-conv HBM-roofline scan).
Random node features / weights, receptor-like degree distribution; reports achieved algorithmic GB/s (SURVEY 8(d) formula)
per (stage_floats, warps, stages) configuration.   python tools/bench_tpconv.py [--edges 200000] [--sweep]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(E, n_nodes, deg, stage_floats, warps, stages, lmax=2, layer=3, reps=5):
    from diffdock_b200 import ops
    from diffdock_b200.tensor_layers import get_irrep_seq
    from diffdock_b200.tp_table import build_table
    for k, v in (('DDB200_TPCONV_WARPS', warps), ('DDB200_TPCONV_STAGES', stages)):
        if v:
            os.environ[k] = str(v)
        else:
            os.environ.pop(k, None)
    seq = get_irrep_seq(48, 10, False, False)
    sh = '1x0e+1x1o+1x2e' if lmax == 2 else '1x0e+1x1o'
    t = build_table(seq[min(layer, 3)], sh, seq[min(layer + 1, 3)], 'fctp', stage_floats=stage_floats)
    h = ops.TpHandle(t)
    g = torch.Generator(device='cuda').manual_seed(0)
    x = torch.randn(n_nodes, t.d_in, device='cuda', generator=g)
    dst = (torch.arange(E, device='cuda') // deg).clamp_max(n_nodes - 1).int()
    src = torch.randint(0, n_nodes, (E,), device='cuda', generator=g).int()
    vec = torch.randn(E, 3, device='cuda', generator=g)
    w = torch.randn(E, t.weight_numel_padded, device='cuda', generator=g)
    out = torch.zeros(n_nodes, t.d_out, device='cuda')
    cnt = torch.zeros(n_nodes, device='cuda')
    for _ in range(2):
        ops.tpconv_accumulate(h, x, src, dst, vec, w, out, cnt)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    times = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.tpconv_accumulate(h, x, src, dst, vec, w, out, cnt)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sorted(times)[len(times) // 2]
    nbytes = E * (4 * t.weight_numel + 16) + 4 * (n_nodes + 1) + 4 * n_nodes * t.d_in + 4 * n_nodes * t.d_out
    return {'E': E, 'stage_floats': t.stage_floats, 'warps': h.info(6), 'stages': h.info(7), 'smem': h.info(5),
            'chunks_per_edge': t.n_chunks, 'ms': round(ms, 4), 'GBps': round(nbytes / ms / 1e6, 1)}


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--edges', type=int, default=200000)
    ap.add_argument('--nodes', type=int, default=48000)
    ap.add_argument('--deg', type=int, default=24)
    ap.add_argument('--sweep', action='store_true')
    ap.add_argument('--scan', action='store_true', help='receptor/ligand size scan (BASELINE config 4)')
    a = ap.parse_args()
    peak = 6566.7
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    if a.sweep:
        for sf in (384, 512, 768, 1024, 1536):
            for warps, stages in ((0, 0), (0, 3), (0, 2), (12, 0)):
                try:
                    r = run(a.edges, a.nodes, a.deg, sf, warps, stages)
                    r['frac'] = round(r['GBps'] / peak, 3)
                    print(json.dumps(r), flush=True)
                except Exception as ex:
                    print('fail', sf, warps, stages, ex, flush=True)
    elif a.scan:
        for n_r in (500, 1000, 2000, 3000, 5000):
            for n_l in (20, 40, 80):
                n_nodes = 32 * (n_r + n_l)
                E = min(32 * (24 * n_r), 400000)          # receptor contact edges of a 32-pose batch, capped by memory
                r = run(E, n_nodes, 24, 768, 0, 0)          # 3 KB chunks: the default of build_table / bench.py
                r.update(n_res=n_r, n_lig=n_l, frac=round(r['GBps'] / peak, 3))
                print(json.dumps(r), flush=True)
    else:
        r = run(a.edges, a.nodes, a.deg, 512, 0, 0)
        r['frac'] = round(r['GBps'] / peak, 3)
        print(json.dumps(r))
