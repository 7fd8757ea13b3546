#!/usr/bin/env python
"""Where one sampling() call spends its wall time (host side): collate/H2D, per-batch constants, graph capture, 20 replays,
D2H.  python tools/profile_sampling.py [--n-res 400 --n-atoms 30 --poses 40]"""
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
    ap.add_argument('--n-res', type=int, default=400)
    ap.add_argument('--n-atoms', type=int, default=30)
    ap.add_argument('--poses', type=int, default=40)
    ap.add_argument('--share', type=int, default=1)
    a = ap.parse_args()
    import bench
    from diffdock_b200 import sampling as S
    from diffdock_b200.cg_model import CGModel
    from diffdock_b200.diffusion_utils import get_t_schedule, get_timestep_embedding, t_to_sigma
    from diffdock_b200.synthetic import default_model_args, make_pose_list
    dev = torch.device('cuda:0')
    args = default_model_args()
    t2s = partial(t_to_sigma, args=args)
    torch.manual_seed(0)
    model = CGModel(t2s, dev, get_timestep_embedding('sinusoidal', args.sigma_embed_dim, args.embedding_scale),
                    **bench.model_kwargs(args)).eval().to(dev)
    sched = get_t_schedule('expbeta', 20)
    out = {}
    for rep in range(4):
        poses = make_pose_list(a.poses, n_res=a.n_res, n_atoms=a.n_atoms, seed=5 + rep, tr_sigma_max=args.tr_sigma_max,
                               share_receptor=bool(a.share))
        marks = {}
        orig_collate, orig_init, orig_run = S._collate_any, S.GraphedSteps.__init__, S.GraphedSteps.run

        def timed(name, fn):
            def w(*x, **k):
                if torch.cuda.is_current_stream_capturing():      # inside the step capture: no timing, no sync
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
        S._collate_any = timed('collate+h2d', orig_collate)
        S.GraphedSteps.__init__ = timed('static+capture', orig_init)
        S.GraphedSteps.run = timed('replays', orig_run)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res, _ = S.sampling(poses, model, 20, sched, sched, sched, dev, t2s, args, batch_size=a.poses, no_final_step_noise=True,
                            rng='philox', seed=1, **bench.TEMPS)
        fin = torch.stack([d['ligand'].pos for d in res]).cpu()
        torch.cuda.synchronize()
        marks['total'] = time.perf_counter() - t0
        S._collate_any, S.GraphedSteps.__init__, S.GraphedSteps.run = orig_collate, orig_init, orig_run
        del model._static
        out[f'rep{rep}'] = {k: round(v * 1e3, 1) for k, v in marks.items()}
    print(json.dumps(out))
