#!/usr/bin/env python
"""Host-side anatomy of BASELINE config 5 on one GPU: per complex the wall time of collate/H2D, per-batch constants, graph
capture (incl. the eager step a new shape needs) and replays, plus the cudaMalloc segments the caching allocator had to add.
    python tools/profile_config5.py [--n 24] [--order index|desc]"""
import argparse
import json
import os
import sys
import time
from functools import partial

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=24)
    ap.add_argument('--order', default='index')
    ap.add_argument('--sync', type=int, default=1, help='1: synchronise around every phase (phase times); 0: only total wall')
    a = ap.parse_args()
    import bench
    from diffdock_b200 import sampling as S
    from diffdock_b200.cg_model import CGModel
    from diffdock_b200.diffusion_utils import get_t_schedule, get_timestep_embedding, t_to_sigma
    from diffdock_b200.synthetic import config5_sizes, default_model_args, make_pose_list
    dev = torch.device('cuda:0')
    args = default_model_args()
    t2s = partial(t_to_sigma, args=args)
    torch.manual_seed(0)
    model = CGModel(t2s, dev, get_timestep_embedding('sinusoidal', args.sigma_embed_dim, args.embedding_scale),
                    **bench.model_kwargs(args)).eval().to(dev)
    sched = get_t_schedule('expbeta', 20)
    sizes = config5_sizes(64, seed=0)[:a.n]
    order = list(range(len(sizes)))
    if a.order == 'desc':
        order.sort(key=lambda i: -sizes[i][0] * sizes[i][1])
    data = {i: make_pose_list(40, n_res=sizes[i][0], n_atoms=sizes[i][1], seed=1000 + i, tr_sigma_max=args.tr_sigma_max,
                              share_receptor=True) for i in order}
    marks = {}

    def timed(name, fn):
        def w(*x, **k):
            if not a.sync or torch.cuda.is_current_stream_capturing():
                return fn(*x, **k)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn(*x, **k)
            torch.cuda.synchronize()
            marks[name] = marks.get(name, 0.0) + time.perf_counter() - t0
            return r
        return w
    orig_static = model._static
    model._static = timed('static', orig_static)
    S._collate_any = timed('collate', S._collate_any)
    S.GraphedSteps.__init__ = timed('capture', S.GraphedSteps.__init__)
    S.GraphedSteps.run = timed('replays', S.GraphedSteps.run)
    rows = []
    torch.cuda.synchronize()
    T0 = time.perf_counter()
    for i in order:
        marks.clear()
        seg0 = torch.cuda.memory_stats().get('segment.all.allocated', 0)
        t0 = time.perf_counter()
        keys = (i << 32) + torch.arange(40, dtype=torch.int64)
        out, _ = S.sampling(data[i], model, 20, sched, sched, sched, dev, t2s, args, batch_size=40, no_final_step_noise=True,
                            rng='philox', seed=2024, pose_keys=keys, **bench.TEMPS)
        res = torch.stack([d['ligand'].pos for d in out])
        if a.sync:
            torch.cuda.synchronize()
        rows.append({'i': i, 'n_res': sizes[i][0], 'n_atoms': sizes[i][1], 'total_ms': round((time.perf_counter() - t0) * 1e3, 1),
                     **{k: round(v * 1e3, 1) for k, v in marks.items()},
                     'new_segments': torch.cuda.memory_stats().get('segment.all.allocated', 0) - seg0})
    torch.cuda.synchronize()
    wall = time.perf_counter() - T0
    for r in rows:
        print(json.dumps(r))
    print(json.dumps({'order': a.order, 'sync': a.sync, 'complexes': len(order), 'wall_s': round(wall, 3),
                      'reserved_GB': round(torch.cuda.memory_reserved() / 2 ** 30, 2)}))
