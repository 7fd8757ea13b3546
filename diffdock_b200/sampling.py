"""Reverse-diffusion sampler: drop-in for ``utils/sampling.py:sampling`` (same arguments and return value).

Per step the reference runs ~40 small PyTorch ops, a Python loop over rotatable bonds with host-sync asserts and a
batched cuSOLVER SVD (utils/sampling.py:133-191, utils/diffusion_utils.py:60-78, utils/torsion.py:75-90,
utils/geometry.py:246-276).  Here a step is: ``set_time`` (device fills) -> score model -> ONE pose-update kernel
(``ddb200_pose_update``) that forms the three perturbations from host-computed SDE coefficients, moves the ligand
rigidly, applies the torsion updates sequentially and Kabsch-aligns - no host synchronisation inside the loop apart
from the neighbour-list sizes in the score model.
"""
from __future__ import annotations

import copy
import math
import os

import numpy as np
import torch

from . import ops
from .diffusion_utils import set_time
from .hetero import collate, collate_shared_receptor


def randomize_position(data_list, no_torsion, no_random, tr_sigma_max, pocket_knowledge=False, pocket_cutoff=7,
                       initial_noise_std_proportion=-1.0, choose_residue=False):
    """Prior sample of every pose, in place: drop-in for ``utils/sampling.py:16-58`` (called by inference.py:237 right before
    ``sampling``).  Same arguments, same random streams in the same order - numpy's global generator for the torsion angles
    (all poses first, :33-40) and, through scipy's ``Rotation.random``, for the rotations; Python's ``random`` / torch's
    global generator for the translation (:45-57) - so a seeded run reproduces the reference's poses bit for bit.
    Host arithmetic like the reference's (a few hundred flops per pose); what changes is that the poses may be the
    light-weight ``inputs.pose_copies`` of a device-resident complex: receptor statistics are then taken on the device once
    and only three floats come back."""
    import random as _random
    from scipy.spatial.transform import Rotation as R
    rec0 = data_list[0]['receptor']
    lig_dev = data_list[0]['ligand'].pos.device
    center_pocket = rec0.pos.mean(dim=0).to(lig_dev)
    if pocket_knowledge:
        cpx = data_list[0]
        ref_lig = torch.from_numpy(cpx['ligand'].orig_pos[0]).float() - cpx.original_center.cpu()
        d = torch.cdist(rec0.pos.cpu(), ref_lig)
        label = torch.any(d < pocket_cutoff, dim=1)
        if torch.any(label):
            center_pocket = rec0.pos.cpu()[label].mean(dim=0)
        else:
            print("No pocket residue below minimum distance ", pocket_cutoff, "taking closest at", torch.min(d))
            center_pocket = rec0.pos.cpu()[torch.argmin(torch.min(d, dim=1)[0])]
    if not no_torsion:
        for g in data_list:
            lig = g['ligand']
            mask = lig.mask_rotate[0] if isinstance(lig.mask_rotate, list) else lig.mask_rotate
            updates = np.random.uniform(low=-np.pi, high=np.pi, size=int(lig.edge_mask.sum()))
            bonds = g['ligand', 'ligand'].edge_index.T[lig.edge_mask].cpu().numpy()
            pos = lig.pos.cpu().numpy().copy()
            for k, (u, v) in enumerate(bonds):              # utils/torsion.py:48-72: fp64 rotation, fp32 coordinates
                if updates[k] == 0:
                    continue
                axis = pos[u] - pos[v]
                rot = R.from_rotvec(axis * updates[k] / np.linalg.norm(axis)).as_matrix()
                pos[mask[k]] = (pos[mask[k]] - pos[v]) @ rot.T + pos[v]
            lig.pos = torch.from_numpy(pos.astype(np.float32))
    rec_sq = None
    for g in data_list:
        lig = g['ligand']
        centre = torch.mean(lig.pos, dim=0, keepdim=True)
        rot = torch.from_numpy(R.random().as_matrix()).float()
        lig.pos = (lig.pos - centre) @ rot.T + center_pocket
        if not no_random:
            rpos = g['receptor'].pos
            if choose_residue:
                idx = _random.randint(0, len(rpos) - 1)
                tr_update = torch.normal(mean=rpos[idx:idx + 1].cpu(), std=0.01)
            elif initial_noise_std_proportion >= 0.0:
                if rec_sq is None or g['receptor'] is not rec0:
                    rec_sq = torch.sqrt(torch.mean(torch.sum(rpos ** 2, dim=1))).cpu()
                tr_update = torch.normal(mean=0, std=rec_sq * initial_noise_std_proportion / 1.73, size=(1, 3))   # fp32 product
            else:
                tr_update = torch.normal(mean=0, std=-initial_noise_std_proportion * tr_sigma_max, size=(1, 3))
            lig.pos = lig.pos + tr_update


def rank_poses(data_list, confidence, original_center, rmsd_classification_cutoff=None):
    """The epilogue of a docking run, inference.py:274-283: absolute ligand coordinates of every pose (the graphs are centred
    on the receptor: + ``original_center``) ordered by decreasing confidence.  ``confidence``: what ``sampling`` returned (None:
    poses stay in sampling order); a list-valued ``rmsd_classification_cutoff`` (multi-threshold confidence head) ranks by the
    first output column, as the reference does.  Returns (ligand_pos [N, n_atoms, 3] float array, confidence [N] or None,
    re_order [N] or None).  ONE device->host copy for all poses instead of one per pose."""
    pos = torch.stack([g['ligand'].pos for g in data_list])
    ligand_pos = pos.cpu().numpy() + torch.as_tensor(original_center).cpu().numpy()
    if confidence is None:
        return ligand_pos, None, None
    if isinstance(rmsd_classification_cutoff, list):
        confidence = confidence[:, 0]
    confidence = confidence.cpu().numpy()
    re_order = np.argsort(confidence)[::-1]
    return ligand_pos[re_order], confidence[re_order], re_order


def is_iterable(arr):
    try:
        iter(arr)
        return True
    except TypeError:
        return False


def _triple(v):
    return list(v) if is_iterable(v) else [v] * 3


def _nan_guard(tr, rot, tor):
    """utils/sampling.py:117-131 without the host round trip: scores are only touched when a NaN shows up in the
    per-pose mean of the translation score."""
    cond = torch.isnan(tr.mean(dim=-1)).any()

    def fix(s):
        if s is None or s.numel() == 0:
            return s
        eps = 0.01 * torch.nanmean(s.abs())
        s = torch.where(cond & torch.isnan(s), eps, s)
        s = torch.where(cond & (s == float('inf')), eps, s)
        return torch.where(cond & (s == float('-inf')), -eps, s)

    return fix(tr), fix(rot), fix(tor)


def step_coefficients(t_idx, inference_steps, tr_schedule, rot_schedule, tor_schedule, t_to_sigma, model_args, ode,
                      temp_sampling, temp_psi, temp_sigma_data):
    """Host scalars (a, c) per degree of freedom such that  perturbation = a * score + c * z
    (utils/sampling.py:97-102,133-186)."""
    last = t_idx == inference_steps - 1
    ts, tp, tsd = _triple(temp_sampling), _triple(temp_psi), _triple(temp_sigma_data)
    out = []
    sig = t_to_sigma(tr_schedule[t_idx], rot_schedule[t_idx], tor_schedule[t_idx])
    lims = [(model_args.tr_sigma_min, model_args.tr_sigma_max), (model_args.rot_sigma_min, model_args.rot_sigma_max),
            (model_args.tor_sigma_min, model_args.tor_sigma_max)]
    for k, sched in enumerate((tr_schedule, rot_schedule, tor_schedule)):
        dt = float(sched[t_idx] - sched[t_idx + 1]) if not last else float(sched[t_idx])
        s_min, s_max = lims[k]
        sigma = float(sig[k])
        g = sigma * math.sqrt(2 * math.log(s_max / s_min))
        if ode:
            a, c = 0.5 * g * g * dt, 0.0
        else:
            a, c = g * g * dt, g * math.sqrt(dt)
        if ts[k] != 1.0:       # low-temperature sampling, :173-186 (uses the SDE form even when ode is set)
            sigma_data = math.exp(tsd[k] * math.log(s_max) + (1 - tsd[k]) * math.log(s_min))
            lam = (sigma_data + sigma) / (sigma_data + sigma / ts[k])
            a, c = g * g * dt * (lam + ts[k] * tp[k] / 2), g * math.sqrt(dt * (1 + tp[k]))
        out += [a, c]
    return out


def crop_receptor(g, cutoff):
    """Device-side ``crop_beyond`` (utils/utils.py:388-413, called per step at utils/sampling.py:104-109): a view of the
    batch whose receptor keeps only the residues within ``cutoff`` of some ligand atom of the same complex, contact edges
    restricted and relabelled.  The reference deep-copies the batch, splits it into a Python list, crops each complex
    and re-collates; here one neighbour-count kernel + index compaction does the same on the device.  The ligand store is
    shared with ``g`` (the sampler keeps updating ``g['ligand'].pos``); like the reference's fresh Batch, the returned
    graph carries no cached receptor embeddings."""
    from .hetero import HeteroGraph, Store
    if not isinstance(g, HeteroGraph):
        raise NotImplementedError("device-side crop_beyond is written against diffdock_b200.hetero.HeteroGraph batches")
    lig, rec, rr = g['ligand'], g['receptor'], g['receptor', 'receptor']
    B = g.num_graphs
    lig_ptr = ops.segment_ptr(lig.batch, B)
    _, _, count = ops.radius(lig.pos, rec.pos, lig_ptr, rec.batch, r=float(cutoff), max_num_neighbors=1)
    keep = count > 0
    out = HeteroGraph()
    out._nodes['ligand'] = lig
    for k, st in g._nodes.items():
        if k not in ('ligand', 'receptor'):
            out._nodes[k] = st
    new = Store()
    for k, v in rec.__dict__.items():
        if k in ('rec_node_attr', 'ptr', 'node_t') or k.startswith('_'):
            continue
        if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == keep.shape[0]:
            setattr(new, k, v[keep])
        else:
            setattr(new, k, v)
    out._nodes['receptor'] = new
    ei = rr.edge_index
    ok = keep[ei[0]] & keep[ei[1]]
    relabel = torch.cumsum(keep.long(), 0) - 1
    out._edges[('receptor', 'receptor')] = Store(edge_index=relabel[ei[:, ok]])
    for et, st in g._edges.items():
        if et != ('receptor', 'receptor'):
            out._edges[et] = st
    for k, v in g._globals.items():
        out._globals[k] = v
    return out


_GRAPH_POOLS = {}


def _graph_pool(device):
    """(memory pool, capture stream, warm-up stream) shared by the step graphs of one device.  torch releases a graph pool when
    the last graph captured into it dies, so a one-kernel anchor graph keeps it alive for the life of the process.  The two
    streams are persistent as well: the caching allocator keeps one block cache per stream, so a fresh side stream per batch
    (the usual warm-up recipe) made every batch cudaMalloc its working set again."""
    dev = torch.device(device)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _GRAPH_POOLS:
        pool = torch.cuda.graph_pool_handle()
        anchor = torch.cuda.CUDAGraph()
        scratch = torch.zeros(8, device=dev)
        torch.cuda.synchronize(dev)
        with torch.cuda.graph(anchor, pool=pool):
            scratch.add_(1.0)
        _GRAPH_POOLS[key] = (pool, anchor, scratch, torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
    e = _GRAPH_POOLS[key]
    return e[0], e[3], e[4]


class GraphedSteps:
    """All reverse-diffusion steps of one batch as replays of ONE CUDA graph.

    The sync-free score model (diffdock_b200.CGModel._forward_sync_free) has static shapes for a given batch, the SDE
    coefficients and schedule times of every step sit in device tables indexed by a device-side step counter, and the pose
    update runs in place - so a step needs no host value at all: the 20 steps of utils/sampling.py:96-191 become 20 graph
    launches (about 500 kernel launches each) with the host idle."""

    def __init__(self, model, g, b, coef_rows, t_rows, bond_u, bond_v, mask_u8, use_torsion, device, draw_noise,
                 philox=None, warmup=1, consume_warmup=False):
        """``consume_warmup``: when a batch of new shapes needs an eager step before the capture, that step IS step 0 of the
        run (``steps_done`` = 1 afterwards) instead of being thrown away - ``run(n)`` then replays the remaining n - 1."""
        self.g, self.b, self.device = g, b, device
        lig = g['ligand']
        self.pos = lig.pos = lig.pos.float().contiguous().clone()         # static buffer, updated in place
        self.coef = torch.tensor(coef_rows, dtype=torch.float32, device=device).contiguous()        # [steps, 6]
        self.times = torch.tensor(t_rows, dtype=torch.float32, device=device).contiguous()          # [steps, 3]
        self.step = torch.zeros(1, dtype=torch.int32, device=device)
        n_lig, n_rec = lig.num_nodes, g['receptor'].num_nodes
        names = ('tr', 'rot', 'tor')

        def one_step():
            t = self.times.index_select(0, self.step.long())[0]                                     # [3] on the device
            for nt, n in (('ligand', n_lig), ('receptor', n_rec)):
                g[nt].node_t = {k: t[i].expand(n) for i, k in enumerate(names)}
            g.complex_t = {k: t[i].expand(b) for i, k in enumerate(names)}
            g._uniform_t = True                      # every graph of the batch is at the same diffusion time
            tr, rot, tor = model(g)[:3]
            tr, rot, tor = _nan_guard(tr, rot, tor)
            has_tor = use_torsion and tor is not None and tor.numel() > 0
            tr_z = rot_z = tor_z = None
            if draw_noise and philox is None:
                tr_z = torch.normal(mean=0, std=1, size=(b, 3), device=device)
                rot_z = torch.normal(mean=0, std=1, size=(b, 3), device=device)
                if has_tor:
                    tor_z = torch.normal(mean=0, std=1, size=tuple(tor.shape), device=device)
            ops.pose_update_dev(self.pos, b, bond_u, bond_v, mask_u8, tr, rot, tor if has_tor else None, self.coef,
                                step_dev=self.step, tr_z=tr_z, rot_z=rot_z, tor_z=tor_z,
                                seed=philox[0] if philox else 0, pose_key=philox[1] if philox else None,
                                use_torsion=has_tor, out=self.pos)
            self.step.add_(1)

        pos0 = self.pos.clone()
        # Outside the capture: the per-batch constants (receptor embedding, static CSR; they read sizes back to the host)
        # and, the first time a model meets a batch of these shapes, one eager step: lazy library handles, table uploads,
        # and above all lazy module loading - a kernel variant (cuBLAS picks them by shape, the conv kernel has a single-CTA
        # and a CTA-pair form) that runs for the first time INSIDE a capture invalidates it.
        if hasattr(model, '_static'):
            model._static(g)
        sig = (b, n_lig, n_rec, int(g['ligand', 'ligand'].edge_index.shape[1]), int(g['receptor', 'receptor'].edge_index.shape[1]),
               int(bond_u.shape[0]) if bond_u is not None else 0, draw_noise, philox is not None)
        seen = getattr(model, '_graph_warmed_shapes', None)
        if seen is None:
            seen = set()
            try:
                model._graph_warmed_shapes = seen
            except Exception:
                pass
        n_warm = 0 if sig in seen else warmup
        seen.add(sig)
        self.steps_done = 0
        pool, cap_stream, side = _graph_pool(device)
        cur = torch.cuda.current_stream(device)
        if n_warm:
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(n_warm):
                    one_step()
            cur.wait_stream(side)
            try:
                model._graph_warmed = True
            except Exception:
                pass
        # One memory pool per device shared by all step graphs of this process: a sampling() call captures a new graph per
        # batch (shapes differ from complex to complex); with a private pool each capture would cudaMalloc its whole
        # footprint again (hundreds of ms for a 1500-residue x 40-pose batch) - the blocks of a finished batch's graph are
        # reused instead.  The capture is opened with capture_begin / capture_end on a persistent stream rather than with the
        # torch.cuda.graph context manager, which synchronises the device, runs gc.collect() and empties the allocator cache
        # on entry: measured on BASELINE config 5 that was 50 ms ... 1.9 s per batch (the collector walks every pose graph
        # the process holds; the emptied cache is re-allocated by the next batch) against 0.15-0.4 s of replays.
        self.graph = torch.cuda.CUDAGraph()
        cap_stream.wait_stream(cur)
        with torch.cuda.stream(cap_stream):
            self.graph.capture_begin(pool=pool)
            try:
                one_step()
            finally:
                self.graph.capture_end()
        cur.wait_stream(cap_stream)
        # the capture itself executes nothing.  An eager step that ran before it is either step 0 of this run (same kernels,
        # same device-side step counter and noise streams as a replay) or is undone.
        if n_warm == 1 and consume_warmup:
            self.steps_done = 1
        elif n_warm:
            self.pos.copy_(pos0)
            self.step.zero_()

    def run(self, n_steps):
        for _ in range(n_steps - self.steps_done):
            self.graph.replay()
        self.steps_done = n_steps
        return self.pos


def _collate_any(items, device):
    """diffdock_b200 graphs: shared-receptor collate; torch_geometric HeteroData (what inference.py passes, utils/sampling.py:80):
    PyG's own ``Batch.from_data_list`` - the score model only needs the attribute contract of SURVEY.md section 8(b)."""
    from .hetero import HeteroGraph
    if isinstance(items[0], HeteroGraph):
        return collate_shared_receptor(items, device)
    try:
        from torch_geometric.data import Batch
    except ImportError as e:
        raise TypeError(f"cannot batch {type(items[0]).__name__} objects: pass diffdock_b200.hetero.HeteroGraph items, or "
                        f"install torch_geometric for HeteroData lists") from e
    return Batch.from_data_list(items).to(device)


def _use_cuda_graph(model, model_args, noise_fn, visualization_list, N, batch_size, cuda_graph):
    if cuda_graph is False or os.environ.get('DDB200_CUDA_GRAPH', '1') == '0':
        return False
    ok = (hasattr(model, 'sync_free_capable') and model.sync_free_capable() and noise_fn is None
          and visualization_list is None and getattr(model_args, 'crop_beyond', None) is None)
    if cuda_graph is True and not ok:
        raise RuntimeError("cuda_graph=True needs the sync-free model path, no noise_fn / visualization / crop_beyond")
    return ok


def _eager_steps(g, b, model, inference_steps, tr_schedule, rot_schedule, tor_schedule, t_schedule, t_to_sigma, model_args,
                 coef_rows, device, bond_u, bond_v, mask_u8, use_torsion, ode, no_random, no_final_step_noise, noise_fn,
                 n_noise, philox, visualization_list, data_list, batch_id, batch_size, n):
    """The step loop launched op by op (utils/sampling.py:96-191): injected noise, per-step receptor cropping, visualisation,
    or a score model whose shapes are outside the sync-free path."""
    coef_dev = torch.tensor(coef_rows, dtype=torch.float32, device=device) if philox else None
    for t_idx in range(inference_steps):
        t_tr, t_rot, t_tor = tr_schedule[t_idx], rot_schedule[t_idx], tor_schedule[t_idx]
        if getattr(model_args, 'crop_beyond', None) is not None:
            tr_sigma = float(t_to_sigma(t_tr, t_rot, t_tor)[0])
            mod = crop_receptor(g, tr_sigma * 3 + model_args.crop_beyond)
        else:
            mod = g
        set_time(mod, t_schedule[t_idx] if t_schedule is not None else None, t_tr, t_rot, t_tor, b,
                 bool(getattr(model_args, 'all_atoms', False)), device)
        mod._uniform_t = True                        # set_time gives every graph of the batch the same diffusion time
        tr_score, rot_score, tor_score = model(mod)[:3]
        tr_score, rot_score, tor_score = _nan_guard(tr_score, rot_score, tor_score)
        has_tor = use_torsion and tor_score.numel() > 0
        if philox:        # in-kernel counter-based noise: the same draws as the graphed path
            step_dev = torch.full((1,), t_idx, dtype=torch.int32, device=device)
            g['ligand'].pos = ops.pose_update_dev(
                g['ligand'].pos.float().contiguous(), b, bond_u, bond_v, mask_u8, tr_score, rot_score,
                tor_score if has_tor else None, coef_dev, step_dev=step_dev, seed=philox[0], pose_key=philox[1],
                use_torsion=has_tor)
        else:
            zero = no_random or (no_final_step_noise and t_idx == inference_steps - 1)
            tr_z = rot_z = tor_z = None
            if not ode and not zero:
                draw = (lambda kind, shape: noise_fn(kind, shape).to(device)) if noise_fn is not None else \
                    (lambda kind, shape: torch.normal(mean=0, std=1, size=shape, device=device))
                tr_z = draw('tr', (n_noise, 3))
                rot_z = draw('rot', (n_noise, 3))
                if use_torsion:
                    tor_z = draw('tor', tuple(tor_score.shape))
            coef = list(coef_rows[t_idx])
            g['ligand'].pos = ops.pose_update(g['ligand'].pos, b, bond_u, bond_v, mask_u8, tr_score, rot_score,
                                              tor_score if has_tor else None, coef, tr_z, rot_z, tor_z, use_torsion=has_tor)
        if visualization_list is not None:
            for idx_b in range(b):
                visualization_list[batch_id * batch_size + idx_b].add(
                    (g['ligand'].pos[idx_b * n:n * (idx_b + 1)].detach().cpu()
                     + data_list[batch_id * batch_size + idx_b].original_center.detach().cpu()), part=1, order=t_idx + 2)


@torch.no_grad()
def sampling(data_list, model, inference_steps, tr_schedule, rot_schedule, tor_schedule, device, t_to_sigma, model_args,
             no_random=False, ode=False, visualization_list=None, confidence_model=None, confidence_data_list=None,
             confidence_model_args=None, t_schedule=None, batch_size=32, no_final_step_noise=False, pivot=None,
             return_full_trajectory=False, temp_sampling=1.0, temp_psi=0.0, temp_sigma_data=0.5, return_features=False,
             noise_fn=None, cuda_graph=None, rng=None, seed=0, pose_keys=None):
    """Same arguments and return value as ``utils/sampling.py:sampling``.  Extensions (all optional):
    ``noise_fn(kind, shape) -> tensor`` (kind in 'tr','rot','tor') replaces the device RNG - used by the injected-noise parity
    tests; otherwise torch.normal is drawn on ``device`` in the reference's order.
    ``cuda_graph``: None = capture the step in a CUDA graph whenever possible, True = require it, False = eager steps.
    ``rng='philox'``: noise from counter-based Philox streams keyed by (``seed``, ``pose_keys[i]``) and indexed by the step,
    so the result for a pose does not depend on batch composition or on how poses are sharded over GPUs;
    ``pose_keys`` [len(data_list)] int64 = (complex id << 32) | pose id, default 0..N-1."""
    assert not (return_full_trajectory or return_features or pivot), "Not implemented yet in new inference version"
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError("diffdock_b200.sampling runs on a CUDA device only (no CPU fallback)")
    N = len(data_list)
    lig0 = data_list[0]['ligand']
    mask_rotate = np.asarray(lig0.mask_rotate[0] if isinstance(lig0.mask_rotate, list) else lig0.mask_rotate)
    mask_u8 = torch.from_numpy(mask_rotate.astype(np.uint8)).contiguous().to(device)
    ei0 = data_list[0]['ligand', 'ligand'].edge_index
    rot_bonds = ei0.T[lig0.edge_mask.cpu()] if ei0.numel() else ei0.T
    bond_u = rot_bonds[:, 0].to(torch.int32).contiguous().to(device)
    bond_v = rot_bonds[:, 1].to(torch.int32).contiguous().to(device)
    use_torsion = not model_args.no_torsion
    confidence = [] if confidence_model is not None else None
    conf_batches = None
    if confidence_model is not None and confidence_data_list is not None:
        conf_batches = [confidence_data_list[i:i + batch_size] for i in range(0, len(confidence_data_list), batch_size)]

    philox = rng == 'philox'
    assert rng in (None, 'philox'), "rng: None (torch.normal in the reference's order) or 'philox'"
    if philox:
        assert noise_fn is None
        keys_all = torch.arange(N, dtype=torch.int64) if pose_keys is None else torch.as_tensor(pose_keys, dtype=torch.int64)
    graphed = _use_cuda_graph(model, model_args, noise_fn, visualization_list, N, batch_size, cuda_graph)

    for batch_id, b0 in enumerate(range(0, N, batch_size)):
        g = _collate_any(data_list[b0:b0 + batch_size], device)
        b = g.num_graphs
        n = len(g['ligand'].pos) // b
        keys = keys_all[b0:b0 + b].to(device) if philox else None
        coef_rows, t_rows = [], []
        for t_idx in range(inference_steps):
            coef = step_coefficients(t_idx, inference_steps, tr_schedule, rot_schedule, tor_schedule, t_to_sigma,
                                     model_args, ode, temp_sampling, temp_psi, temp_sigma_data)
            if ode or no_random or (no_final_step_noise and t_idx == inference_steps - 1):
                coef[1] = coef[3] = coef[5] = 0.0          # no noise in this step (utils/sampling.py:136-145,158-161)
            coef_rows.append(coef)
            t_rows.append([float(tr_schedule[t_idx]), float(rot_schedule[t_idx]), float(tor_schedule[t_idx])])
        if graphed and t_schedule is None and b > 0:
            steps = GraphedSteps(model, g, b, coef_rows, t_rows, bond_u, bond_v, mask_u8, use_torsion, device,
                                 draw_noise=not (ode or no_random), philox=(seed, keys) if philox else None,
                                 consume_warmup=True)
            steps.run(inference_steps)
        else:
            _eager_steps(g, b, model, inference_steps, tr_schedule, rot_schedule, tor_schedule, t_schedule, t_to_sigma,
                         model_args, coef_rows, device, bond_u, bond_v, mask_u8, use_torsion, ode, no_random,
                         no_final_step_noise, noise_fn, min(batch_size, N), (seed, keys) if philox else None,
                         visualization_list, data_list, batch_id, batch_size, n)
        for i in range(b):
            data_list[b0 + i]['ligand'].pos = g['ligand'].pos[i * n:n * (i + 1)]
        if confidence_model is not None:
            if conf_batches is not None:
                cg = collate(copy.deepcopy(conf_batches[batch_id]))
                cg['ligand'].pos = g['ligand'].pos.cpu()
                cg = cg.to(device)
                if getattr(confidence_model_args, 'crop_beyond', None) is not None:     # utils/sampling.py:213-217
                    cg = crop_receptor(cg, confidence_model_args.crop_beyond)
                set_time(cg, 0, 0, 0, 0, b, confidence_model_args.all_atoms, device)
                out = confidence_model(cg)
            else:
                out = confidence_model(g)
            confidence.append(out[0] if type(out) is tuple else out)
    if confidence_model is not None:
        confidence = torch.nan_to_num(torch.cat(confidence, dim=0), nan=-1000)
    return data_list, confidence
