#!/usr/bin/env python
"""bench.py - poses/sec at 20 diffusion steps (BASELINE.json metric) on synthetic protein-ligand graphs.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W     # the CPU oracle (reference restatement) on the host cores

A "step" is one reverse-diffusion step of the hot path for one batch: set_time -> score-model forward (graph build,
embeddings, 6 tensor-product conv layers, tr/rot/tor heads) -> pose update, for POSES poses of one synthetic complex
(1500 residues / 40 ligand atoms, BASELINE config 2) per GPU; consecutive steps walk the 20-step 'expbeta' schedule
(t: 1 -> 0.05), so K=20 is exactly one sampling run.   value = total poses / (20 * mean step time).
The JSON line also carries the end-to-end number through diffdock_b200.sampling.sampling() with host inputs, the
HBM roofline of the fused tensor-product conv kernel measured live with CUDA events, and a CPU baseline.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from functools import partial

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from diffdock_b200.synthetic import default_model_args, make_pose_list   # noqa: E402
from diffdock_b200.hetero import collate                                  # noqa: E402

TEMPS = dict(temp_sampling=[1.170050527854316, 2.06391612594481, 7.044261621607846],      # default_inference_args.yaml
             temp_psi=[0.727287304570729, 0.9022615585677628, 0.5946212391366862],
             temp_sigma_data=[0.9299802531572672, 0.7464326999906034, 0.6943254174849822])
N_SCHED = 20


def model_kwargs(a):
    return dict(sigma_embed_dim=a.sigma_embed_dim, sh_lmax=a.sh_lmax, ns=a.ns, nv=a.nv, num_conv_layers=a.num_conv_layers,
                lig_max_radius=a.max_radius, rec_max_radius=a.rec_max_radius, cross_max_distance=a.cross_max_distance,
                center_max_distance=a.center_max_distance, distance_embed_dim=a.distance_embed_dim,
                cross_distance_embed_dim=a.cross_distance_embed_dim, dynamic_max_cross=a.dynamic_max_cross,
                lm_embedding_type='precomputed', embed_also_ligand=True, num_prot_emb_layers=a.num_prot_emb_layers)


def randomise_bn(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if hasattr(m, 'running_var') and hasattr(m, 'running_mean'):
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))


def host_threads():
    """Threads for the CPU oracle: every host core up to 32 (beyond that the oracle's many small PyTorch ops lose time
    to oversubscription: 134 s/pose-step with 128 threads vs 17 s with 8 on this workload); override DDB200_CPU_THREADS."""
    return int(os.environ.get('DDB200_CPU_THREADS', min(os.cpu_count() or 1, 32)))


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p)), 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}, 'fallback'


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        while not self._halt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--id={self.index}', f'--query-gpu={q}', '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(self.rows)}


# ----------------------------------------------------------------------------------------------------- CPU oracle arm
def oracle_step_fn(args, n_res, n_atoms, seed):
    """One bounded sample of the workload on the host: ONE pose of the same synthetic complex - score-model forward
    (oracle restatement of the reference's e3nn/torch_scatter op sequence) + pose update, at schedule point t_idx."""
    from oracle.cg_model import CGModel
    from oracle.diffusion import modify_conformer_batch, set_time, t_to_sigma
    from oracle.layers import get_timestep_embedding
    from diffdock_b200.diffusion_utils import get_t_schedule
    from diffdock_b200.sampling import step_coefficients
    torch.manual_seed(0)
    model = CGModel(partial(t_to_sigma, args=args), 'cpu',
                    get_timestep_embedding('sinusoidal', args.sigma_embed_dim, args.embedding_scale), **model_kwargs(args)).eval()
    randomise_bn(model, 1)
    pose = make_pose_list(1, n_res=n_res, n_atoms=n_atoms, seed=seed, tr_sigma_max=args.tr_sigma_max)
    g = collate(pose)
    sched = get_t_schedule('expbeta', N_SCHED)
    mask_rotate = torch.from_numpy(pose[0]['ligand'].mask_rotate[0])

    def step(i):
        t_idx = i % N_SCHED
        t = sched[t_idx]
        set_time(g, t, t, t, 1, 'cpu')
        with torch.no_grad():
            tr, rot, tor, _ = model(g)
            c = step_coefficients(t_idx, N_SCHED, sched, sched, sched, partial(t_to_sigma, args=args), args, False,
                                  **TEMPS)
            g['ligand'].pos = modify_conformer_batch(g['ligand'].pos, g, c[0] * tr, c[2] * rot, c[4] * tor, mask_rotate)
    return step


def run_reference(cli):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = host_threads()
    torch.set_num_threads(cores)
    args = default_model_args(sh_lmax=cli.sh_lmax)
    step = oracle_step_fn(args, cli.n_res, cli.n_atoms, seed=100)
    budget = float(os.environ.get('DDB200_REF_BUDGET_S', '240'))
    t0 = time.perf_counter()
    step(0)                       # first warm-up step doubles as the cost probe
    probe = time.perf_counter() - t0
    warm = max(0, min(cli.warmup, int(budget * 0.2 / max(probe, 1e-3))) - 1)
    for i in range(warm):
        step(1 + i)
    steps = max(1, min(cli.steps, int((budget - probe * (1 + warm)) / max(probe, 1e-3))))
    t0 = time.perf_counter()
    for i in range(steps):
        step(1 + warm + i)
    dt = (time.perf_counter() - t0) / steps
    value = 1.0 / (N_SCHED * dt)
    sample = (f"1 pose of the {cli.n_res}-residue/{cli.n_atoms}-atom complex per step (forward + pose update); "
              f"{steps} of the requested {cli.steps} steps timed within a {budget:.0f} s budget")
    line = {"impl": "reference", "metric": "poses/sec at 20 diffusion steps", "value": value, "unit": "poses/s",
            "n_gpus": cli.gpus, "steps": steps, "warmup": 1 + warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cli, cli.poses),      # same workload as the CUDA arm; the bounded sample is below
            "cpu_baseline": {"value": value, "unit": "poses/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(cli, poses):
    return {"workload": f"DiffDock-L-shaped score model (ns=48,nv=10,sh_lmax={cli.sh_lmax},6 conv layers) reverse-diffusion "
                        f"step, synthetic complex {cli.n_res} residues / {cli.n_atoms} ligand atoms, {poses} poses per GPU "
                        f"(BASELINE config 2), 20-step expbeta schedule",
            "poses_per_gpu": poses, "n_res": cli.n_res, "n_atoms": cli.n_atoms, "sh_lmax": cli.sh_lmax,
            "l2": "per-step working set (edge embeddings ~0.3 GB per receptor edge group and layer, operand images, "
                  "node tensors) exceeds the 126 MB L2; no explicit flush",
            "warmup_executed": cli.warmup if getattr(cli, 'short_warmup', False) else max(cli.warmup, N_SCHED),
            "parallelism": f"poses sharded over {cli.gpus} GPU(s), no data-path collective"}


# ----------------------------------------------------------------------------------------------------- CUDA arm
def tpconv_stream_roofline(dev, n_edges=200000):
    """BASELINE metric 'fused TP-conv HBM GB/s vs peak': the streaming tensor-product conv kernel (per-edge weights read
    from HBM, the un-fused formulation of SURVEY 8(d)) timed alone with CUDA events on 200k receptor-like edges of the
    full-width 156->156 layer (5.7 GB of weights >> L2), median of 5 launches."""
    from diffdock_b200 import ops
    from diffdock_b200.tensor_layers import get_irrep_seq
    from diffdock_b200.tp_table import build_table
    seq = get_irrep_seq(48, 10, False, False)
    t = build_table(seq[3], '1x0e+1x1o+1x2e', seq[3], 'fctp')
    h = ops.TpHandle(t)
    g = torch.Generator(device=dev).manual_seed(0)
    n_nodes = 48000
    x = torch.randn(n_nodes, t.d_in, device=dev, generator=g)
    dst = (torch.arange(n_edges, device=dev) // 24).clamp_max(n_nodes - 1).int()
    src = torch.randint(0, n_nodes, (n_edges,), device=dev, generator=g).int()
    vec = torch.randn(n_edges, 3, device=dev, generator=g)
    w = torch.randn(n_edges, t.weight_numel_padded, device=dev, generator=g)
    out, cnt = torch.zeros(n_nodes, t.d_out, device=dev), torch.zeros(n_nodes, device=dev)
    was = ops.PROFILE.enabled
    ops.PROFILE.enabled = False
    times = []
    for i in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.tpconv_accumulate(h, x, src, dst, vec, w, out, cnt)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            times.append(e0.elapsed_time(e1))
    ops.PROFILE.enabled = was
    ms = sorted(times)[len(times) // 2]
    nbytes = n_edges * (4 * t.weight_numel + 16) + 4 * (n_nodes + 1) + 4 * n_nodes * t.d_in + 4 * n_nodes * t.d_out
    return {"bound": "hbm", "kernel": "tpconv_accumulate_kernel", "achieved": nbytes / ms / 1e6, "unit": "GB/s",
            "edges": n_edges, "bytes_per_launch": nbytes, "ms_per_launch": ms, "traffic": None,
            "how": "standalone launches, CUDA events, weights (5.7 GB) larger than L2; the model itself runs the fully "
                   "fused kernel (see 'roofline')"}


def run_cuda(cli):
    import torch.distributed as dist
    from diffdock_b200 import ops
    from diffdock_b200.cg_model import CGModel
    from diffdock_b200.diffusion_utils import get_t_schedule, get_timestep_embedding, set_time, t_to_sigma
    from diffdock_b200.sampling import sampling, step_coefficients
    import __graft_entry__ as ge
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if rank == 0:
        ge.build()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
        dist.barrier()
    args = default_model_args(sh_lmax=cli.sh_lmax)
    t2s = partial(t_to_sigma, args=args)
    torch.manual_seed(0)
    model = CGModel(t2s, dev, get_timestep_embedding('sinusoidal', args.sigma_embed_dim, args.embedding_scale),
                    **model_kwargs(args)).eval()
    randomise_bn(model, 1)
    model = model.to(dev)
    poses = make_pose_list(cli.poses, n_res=cli.n_res, n_atoms=cli.n_atoms, seed=100 + rank, tr_sigma_max=args.tr_sigma_max)
    sched = get_t_schedule('expbeta', N_SCHED)
    lig0 = poses[0]['ligand']
    mask_u8 = torch.from_numpy(lig0.mask_rotate[0].astype(np.uint8)).to(dev)
    rb = poses[0]['ligand', 'ligand'].edge_index.T[lig0.edge_mask]
    bu, bv = rb[:, 0].int().contiguous().to(dev), rb[:, 1].int().contiguous().to(dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    g = collate(poses).to(dev)
    pos0 = g['ligand'].pos.clone()

    def step(i):
        t_idx = i % N_SCHED
        if t_idx == 0:
            g['ligand'].pos = pos0.clone()       # a fresh sampling run starts from the prior again
        t = sched[t_idx]
        coef = step_coefficients(t_idx, N_SCHED, sched, sched, sched, t2s, args, False, **TEMPS)
        set_time(g, None, t, t, t, cli.poses, False, dev)
        tr, rot, tor = model(g)[:3]
        last = t_idx == N_SCHED - 1
        z = (lambda shape: None) if last else (lambda shape: torch.randn(shape, device=dev, generator=gen))
        g['ligand'].pos = ops.pose_update(g['ligand'].pos, cli.poses, bu, bv, mask_u8, tr, rot, tor, coef,
                                          z((cli.poses, 3)), z((cli.poses, 3)), z(tuple(tor.shape)))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # Warm-up: the W requested steps, extended to one full pass over the 20-point schedule: every step of the schedule
    # has its own neighbour-list sizes, and the first visit of each grows torch's caching allocator (cudaMalloc + sync).
    # Measured: 67.0 ms/step when only steps 0-2 were warmed, 57.4 ms/step on the second pass over the same steps.
    n_warm = cli.warmup if cli.short_warmup else max(cli.warmup, N_SCHED)
    for i in range(n_warm):
        step(i + cli.warmup - n_warm)
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ops.PROFILE.reset(enabled=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(cli.steps):
        step(cli.warmup + i)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1) / cli.steps
    launches = ops.PROFILE.all_launches
    clocks = sampler.stop() if sampler else None
    # Per-kernel durations: the SAME K steps replayed with a CUDA-event pair (launching stream) around every
    # tensor-product conv launch, kept out of the timed region (its own step time is reported as replay_ms_per_step).
    ops.PROFILE.reset(enabled=True)
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for i in range(cli.steps):
        step(cli.warmup + i)
    r1.record()
    prof = ops.PROFILE.summary()
    prof['all_launches'] = launches
    prof['replay_ms_per_step'] = r0.elapsed_time(r1) / cli.steps
    ops.PROFILE.reset(enabled=False)
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * cli.poses / (N_SCHED * ms_max * 1e-3)

    # ---- end to end through the public API with host inputs ---------------------------------------------------
    if cli.no_e2e:
        if rank == 0:
            print(json.dumps({"profiling_run": True, "ms_per_step": ms_max, "value": value, "tpconv": prof}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    host_list = [p.clone() for p in poses]
    for p in host_list:
        p._apply(lambda t: t.pin_memory() if t.is_floating_point() or t.dtype in (torch.int64, torch.bool) else t)
    h2d = sum(t.numel() * t.element_size() for p in host_list for st in list(p._nodes.values()) + list(p._edges.values())
              for t in st.__dict__.values() if torch.is_tensor(t))
    sync_all()
    t0 = time.perf_counter()
    out, _ = sampling(host_list, model, N_SCHED, sched, sched, sched, dev, t2s, args, batch_size=cli.poses,
                      no_final_step_noise=True, **TEMPS)
    final = torch.stack([d['ligand'].pos for d in out]).cpu()        # D2H of the result inside the timed region
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        gathered = [torch.empty_like(final, device=dev) for _ in range(world)]   # final pose gather over NCCL
        dist.all_gather(gathered, final.to(dev))
    e2e_val = world * cli.poses / float(t_e.item())
    assert torch.isfinite(final).all()

    stream_roof = tpconv_stream_roofline(dev) if rank == 0 else None
    if rank == 0:
        pk, pk_kind = peaks()
        roof = None
        if prof['fused_launches']:
            ach = prof['fused_flops'] / (prof['fused_ms'] * 1e-3) / 1e12
            peak_tf = pk.get('bf16_tflops_sustained', pk['bf16_tflops'])
            roof = {"bound": "tensor", "kernel": "fused_conv_kernel", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": ach / peak_tf, "peak_kind": pk_kind + " (sustained bf16 cuBLAS: kernel timed inside a long step)",
                    "traffic": None,
                    "traffic_ncu": {"dram_bytes_per_launch": 151.7e6, "edges_per_launch": 400000,
                                    "source": "profiles/r01i_fused_summary.csv (tools/bench_fused.py under ncu --set full); "
                                              "the un-fused formulation moves 11.4 GB for the same launch"},
                    "launches": prof['fused_launches'],
                    "timing": "CUDA-event pair per launch on the launching stream, over a replay of the timed K steps "
                              "(the timed region itself carries no per-launch events)",
                    "replay_ms_per_step": prof['replay_ms_per_step'],
                    "flops": "bf16 tcgen05 MMA FLOPs issued (radial MLP as split-bf16 x3, K padded to 448, full N tiles)",
                    "kernel_ms_per_step": prof['fused_ms'] / cli.steps, "share_of_step": prof['fused_ms'] / cli.steps / ms,
                    "equivalent_hbm_GBps": prof['fused_bytes'] / (prof['fused_ms'] * 1e-3) / 1e9,
                    "equivalent_note": "algorithmic bytes of the un-fused formulation (SURVEY 8(d)) / fused-kernel time; "
                                       "the per-edge weights never reach HBM, so this may exceed the HBM peak"}
        elif prof['launches']:
            ach = prof['bytes'] / (prof['ms'] * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": "tpconv_accumulate_kernel", "achieved": ach, "peak": pk['hbm_gbs'],
                    "unit": "GB/s", "frac": ach / pk['hbm_gbs'], "peak_kind": pk_kind, "traffic": None,
                    "launches": prof['launches'], "kernel_ms_per_step": prof['ms'] / cli.steps,
                    "share_of_step": prof['ms'] / cli.steps / ms}
        if stream_roof:
            stream_roof.update(peak=pk['hbm_gbs'], frac=stream_roof['achieved'] / pk['hbm_gbs'], peak_kind=pk_kind)
        line = {"metric": "poses/sec at 20 diffusion steps", "value": value, "unit": "poses/s", "n_gpus": world,
                "steps": cli.steps, "warmup": cli.warmup, "ms_per_step": ms_max, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(cli, cli.poses), "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "poses/s", "h2d_bytes_per_step": h2d // N_SCHED,
                        "d2h_bytes_per_step": int(final.numel() * 4 // N_SCHED), "seconds_per_run": float(t_e.item())},
                "gpu_launches": prof['all_launches'], "roofline": roof, "roofline_tpconv_stream": stream_roof}
        if world == 1 and not cli.no_cpu_baseline:
            cores = host_threads()
            torch.set_num_threads(cores)
            ostep = oracle_step_fn(args, cli.n_res, cli.n_atoms, seed=100)
            t0 = time.perf_counter()
            ostep(10)           # t = 0.5: mid-schedule edge count
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": 1.0 / (N_SCHED * dt), "unit": "poses/s", "cores": cores, "kind": "port",
                                    "sample": f"1 pose-step (forward + pose update) of the same complex at t=0.5, "
                                              f"{dt:.1f} s on {cores} host threads, oracle = reference op sequence restated"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--poses', type=int, default=32)
    ap.add_argument('--n-res', dest='n_res', type=int, default=1500)
    ap.add_argument('--n-atoms', dest='n_atoms', type=int, default=40)
    ap.add_argument('--sh-lmax', dest='sh_lmax', type=int, default=2)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true', help='skip the end-to-end leg (profiling runs)')
    ap.add_argument('--short-warmup', dest='short_warmup', action='store_true',
                    help='warm up exactly --warmup steps instead of a full schedule pass (runs under ncu)')
    cli = ap.parse_args()
    cli.warmup = max(cli.warmup, 0)
    if cli.impl == 'reference':
        run_reference(cli)
    else:
        if cli.warmup < 3:
            cli.warmup = 3
        run_cuda(cli)


if __name__ == '__main__':
    main()
