#!/usr/bin/env python
"""Micro-benchmark of the fully fused convolution kernel alone (csrc/fused_conv.cu) on receptor-like edges of the full-width
156 -> 156 layer: bf16 tcgen05 TFLOP/s issued, and - with DDB200_FUSED_DEBUG=1 - the clock breakdown of its warp roles.
    python tools/bench_fused.py [--edges 400000] [--layer 3]        DDB200_FUSED_CTA_PAIR=0|1 selects the kernel variant"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

NAMES = {0: 'mma_role', 1: 'mma_wait_acc_free', 2: 'mma_wait_B', 3: 'mma_wait_A2', 5: 'prod_wait_stage_free', 6: 'relay_wait_full', 7: 'relay_arrive',
         8: 'cons_tile_loop', 9: 'cons_wait_hidden', 10: 'a0_build', 16: 'cons_M_build', 17: 'cons_body_kind0', 18: 'cons_body_other', 19: 'n_kind0', 20: 'n_other',
         21: 'cons_flush', 22: 'cons_z', 23: 'cons_wait_acc', 24: 'cons_fma', 11: 'unit_total', 12: 'units'}

def scan():
    """BASELINE config 4 for the kernel the model runs: receptor sizes 500-5000 x ligand sizes 20-80, 32 poses, full-width
    156->156 layer, receptor contact edges (24 per residue) + cross edges at the mid-schedule cut-off are emulated by E
    CSR-sorted edges over 32 (N_r + N_l) nodes.  Prints one JSON line per shape: ms, issued bf16 TFLOP/s, and the
    SURVEY 8(d) equivalent HBM rate (algorithmic bytes of the un-fused formulation / time) as a fraction of the measured peak."""
    from diffdock_b200 import fused
    from diffdock_b200.tensor_layers import get_irrep_seq
    from diffdock_b200.tp_table import build_table
    peak = 6566.7
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    ns = 48
    seq = get_irrep_seq(ns, 10, False, False)
    t = build_table(seq[3], '1x0e+1x1o+1x2e', seq[3], 'fctp')
    g = torch.Generator(device='cuda').manual_seed(0)
    H = K1 = 3 * ns
    plan = fused.FusedPlan(t, torch.randn(H, K1, device='cuda', generator=g) / K1 ** 0.5, torch.randn(H, device='cuda', generator=g) * 0.1,
                           torch.randn(t.weight_numel, H, device='cuda', generator=g) / H ** 0.5,
                           torch.randn(t.weight_numel, device='cuda', generator=g) * 0.1)
    for n_r in (500, 1000, 2000, 3000, 5000):
        for n_l in (20, 40, 80):
            N = 32 * (n_r + n_l)
            E = 32 * 24 * n_r + 32 * n_l * min(n_r, 300)        # contacts + ~300 residues within the cut-off per atom
            x = torch.randn(N, t.d_in, device='cuda', generator=g)
            tgt = torch.sort(torch.randint(0, N, (E,), device='cuda', generator=g)).values.int()
            src = torch.randint(0, N, (E,), device='cuda', generator=g).int()
            vec = torch.randn(E, 3, device='cuda', generator=g)
            ea = torch.randn(E, ns, device='cuda', generator=g)
            out, cnt = torch.zeros(N, t.d_out, device='cuda'), torch.zeros(N, device='cuda')
            ts = []
            for i in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fused.fused_conv(plan, ea, x, ns, tgt, src, x, vec, out, cnt)
                e1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            nbytes = E * (4 * t.weight_numel + 16) + 4 * (N + 1) + 4 * N * t.d_in + 4 * N * t.d_out
            print(json.dumps({'n_res': n_r, 'n_lig': n_l, 'nodes': N, 'E': E, 'ms': round(ms, 3),
                              'bf16_issued_TFLOPs': round(((E + 127) // 128) * plan.mma_flops_per_tile / ms / 1e9, 1),
                              'algorithmic_TFLOPs': round(E * plan.alg_flops_per_edge / ms / 1e9, 1),
                              'equivalent_GBps': round(nbytes / ms / 1e6, 1), 'frac_of_hbm_peak': round(nbytes / ms / 1e6 / peak, 3)}),
                  flush=True)
            del x, tgt, src, vec, ea, out, cnt


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--scan', action='store_true', help='BASELINE config 4 size scan of the fused kernel')
    ap.add_argument('--edges', type=int, default=400000)
    ap.add_argument('--nodes', type=int, default=48000)
    ap.add_argument('--deg', type=int, default=24)
    ap.add_argument('--layer', type=int, default=3)
    a = ap.parse_args()
    if a.scan:
        scan()
        sys.exit(0)
    from diffdock_b200 import _lib, fused
    from diffdock_b200.tensor_layers import get_irrep_seq
    from diffdock_b200.tp_table import build_table
    ns = 48
    seq = get_irrep_seq(ns, 10, False, False)
    t = build_table(seq[min(a.layer, 3)], '1x0e+1x1o+1x2e', seq[min(a.layer + 1, 3)], 'fctp')
    g = torch.Generator(device='cuda').manual_seed(0)
    H, K1 = 3 * ns, 3 * ns
    w1 = torch.randn(H, K1, device='cuda', generator=g) / K1 ** 0.5
    b1 = torch.randn(H, device='cuda', generator=g) * 0.1
    w2 = torch.randn(t.weight_numel, H, device='cuda', generator=g) / H ** 0.5
    b2 = torch.randn(t.weight_numel, device='cuda', generator=g) * 0.1
    plan = fused.FusedPlan(t, w1, b1, w2, b2)
    E, N = a.edges, a.nodes
    x = torch.randn(N, t.d_in, device='cuda', generator=g)
    tgt = (torch.arange(E, device='cuda') // a.deg).clamp_max(N - 1).int()
    src = torch.randint(0, N, (E,), device='cuda', generator=g).int()
    vec = torch.randn(E, 3, device='cuda', generator=g)
    ea = torch.randn(E, ns, device='cuda', generator=g)
    out = torch.zeros(N, t.d_out, device='cuda')
    cnt = torch.zeros(N, device='cuda')
    run = lambda: fused.fused_conv(plan, ea, x, ns, tgt, src, x, vec, out, cnt)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    dbg = (C.c_uint64 * 32)()
    have_dbg = _lib.lib().ddb200_fused_debug_read(dbg) == 0
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    flops = ((E + 127) // 128) * plan.mma_flops_per_tile
    r = {'E': E, 'tiles': plan.n_tiles, 'ms': round(ms, 3), 'bf16_issued_TFLOPs': round(flops / ms / 1e9, 1),
         'pair': os.environ.get('DDB200_FUSED_CTA_PAIR', '1')}
    if have_dbg and _lib.lib().ddb200_fused_debug_read(dbg) == 0:
        pair = os.environ.get('DDB200_FUSED_CTA_PAIR', '1') != '0'
        units = max(int(dbg[12]), 1)         # counted once per CTA: a pair unit counts twice
        one_cta = {0, 1, 2, 3, 6, 7}         # counters only the leader (MMA role) or only the peer (relay) adds to
        r['clk_per_cta_unit'] = {NAMES[i]: int(dbg[i]) * (2 if (pair and i in one_cta) else 1) // units for i in NAMES if i != 12}
        r['units'] = units // 5
        r['issue_clk_per_mma'] = round(dbg[13] / max(dbg[14], 1), 1)
        r['sm_clock_ghz_in_kernel'] = round(dbg[25] / max(dbg[26], 1), 3)     # clock64 ticks per globaltimer ns, CTA 0
    print(json.dumps(r), flush=True)
