#!/usr/bin/env python
"""Input side: time the receptor contact graph on the GPU (ddb200_contact_count/_fill) next to
the reference formulation on the host (torch.cdist + the Python loop of datasets/process_mols.py:176-192, restated in
oracle/inputs.py), and the packed-complex upload next to the per-attribute upload.  One JSON line per size."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def reference_loop(coords, cutoff, k):
    """datasets/process_mols.py:176-192 as the reference runs it (torch.cdist + np.where / np.argsort per residue)."""
    distances = torch.cdist(coords, coords)
    src_list, dst_list = [], []
    for i in range(len(coords)):
        dst = list(np.where(distances[i, :] < cutoff)[0])
        dst.remove(i)
        if len(dst) > k:
            dst = list(np.argsort(distances[i, :]))[1: k + 1]
        if len(dst) == 0:
            dst = list(np.argsort(distances[i, :]))[1:2]
        src_list.extend([i] * len(dst))
        dst_list.extend(dst)
    return torch.from_numpy(np.asarray([dst_list, src_list]))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--sizes', type=int, nargs='+', default=[500, 1500, 3000])
    ap.add_argument('--atoms', type=int, default=12000, help='an all-atom-sized point set (device only; the host loop needs minutes)')
    a = ap.parse_args()
    from diffdock_b200.inputs import PackedComplex, contact_graph
    from diffdock_b200.synthetic import make_complex
    for n in a.sizes + [a.atoms]:
        rng = np.random.default_rng(n)
        R = (3.0 * n / (4.0 * np.pi * (0.0075 if n <= 3000 else 0.06))) ** (1.0 / 3.0)
        v = rng.normal(size=(n, 3))
        pos = torch.from_numpy((v / np.linalg.norm(v, axis=1, keepdims=True) * (R * rng.uniform(size=(n, 1)) ** (1 / 3.0))).astype(np.float32))
        cutoff, k = (15.0, 24) if n <= 3000 else (5.0, 8)
        dev = pos.cuda()
        for _ in range(2):
            ei = contact_graph(dev, cutoff, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ei = contact_graph(dev, cutoff, k)
        e1.record()
        torch.cuda.synchronize()
        r = {'points': n, 'cutoff': cutoff, 'max_neighbors': k, 'edges': int(ei.shape[1]), 'gpu_ms': round(e0.elapsed_time(e1) / 10, 3)}
        if n <= 3000:
            t0 = time.perf_counter()
            ref = reference_loop(pos, cutoff, k)
            r['host_reference_ms'] = round((time.perf_counter() - t0) * 1e3, 1)
            r['same_edge_set'] = bool(ref.shape == ei.shape and torch.equal(torch.sort(ref[0] * n + ref[1]).values,
                                                                             torch.sort(ei.cpu()[0] * n + ei.cpu()[1]).values))
        print(json.dumps(r), flush=True)
    g = make_complex(n_res=1500, n_atoms=40, seed=0)
    pk = PackedComplex.pack(g)
    for name, fn in (('packed_one_copy', lambda: pk.to('cuda')), ('per_attribute', lambda: g.clone().to('cuda'))):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        print(json.dumps({'upload': name, 'bytes': int(pk.header['nbytes']), 'ms': round((time.perf_counter() - t0) / 5 * 1e3, 3)}), flush=True)
