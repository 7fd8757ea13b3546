"""Noise schedules and time embeddings of the sampler (utils/diffusion_utils.py:28-32,99-143,146-168)."""
import math

import numpy as np
import torch
import torch.nn.functional as F


def t_to_sigma(t_tr, t_rot, t_tor, args):
    """Geometric interpolation sigma_min^(1-t) * sigma_max^t per degree of freedom (utils/diffusion_utils.py:28-32)."""
    tr = args.tr_sigma_min ** (1 - t_tr) * args.tr_sigma_max ** t_tr
    rot = args.rot_sigma_min ** (1 - t_rot) * args.rot_sigma_max ** t_rot
    tor = args.tor_sigma_min ** (1 - t_tor) * args.tor_sigma_max ** t_tor
    return tr, rot, tor


def sinusoidal_embedding(timesteps, embedding_dim, max_positions=10000):
    assert timesteps.dim() == 1
    half = embedding_dim // 2
    rate = math.log(max_positions) / (half - 1)
    freq = torch.exp(torch.arange(half, dtype=torch.float32, device=timesteps.device) * -rate)
    arg = timesteps.float()[:, None] * freq[None, :]
    emb = torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1), mode='constant')
    return emb


class GaussianFourierProjection(torch.nn.Module):
    """Gaussian Fourier embedding of the noise level (utils/diffusion_utils.py:113-125): a fixed random projection W drawn at
    construction (non-trainable parameter, so that it follows ``.to(device)`` and sits in the module's state_dict)."""

    def __init__(self, embedding_size=256, scale=1.0):
        super().__init__()
        self.W = torch.nn.Parameter(torch.randn(embedding_size // 2) * scale, requires_grad=False)

    def forward(self, x):
        x_proj = x[:, None] * self.W[None, :] * 2 * np.pi
        return torch.cat([torch.sin(x_proj), torch.cos(x_proj)], dim=-1)


def get_timestep_embedding(embedding_type, embedding_dim, embedding_scale=10000):
    """utils/diffusion_utils.py:128-135."""
    if embedding_type == 'sinusoidal':
        return lambda x: sinusoidal_embedding(embedding_scale * x, embedding_dim)
    if embedding_type == 'fourier':
        return GaussianFourierProjection(embedding_size=embedding_dim, scale=embedding_scale)
    raise NotImplementedError(embedding_type)


def get_t_schedule(sigma_schedule='expbeta', inference_steps=20, inf_sched_alpha=1, inf_sched_beta=1, t_max=1):
    """'expbeta' schedule (utils/diffusion_utils.py:138-143): beta-quantiles of a linear grid, 1 -> 1/steps for a=b=1."""
    if sigma_schedule != 'expbeta':
        raise Exception()
    from scipy.stats import beta
    hi = beta.cdf(t_max, a=inf_sched_alpha, b=inf_sched_beta)
    grid = np.linspace(hi, 0, inference_steps + 1)[:-1]
    return beta.ppf(grid, a=inf_sched_alpha, b=inf_sched_beta)


def set_time(complex_graphs, t, t_tr, t_rot, t_tor, batchsize, all_atoms, device, include_miscellaneous_atoms=False):
    """Writes node_t / complex_t like utils/diffusion_utils.py:146-168 - one fill per tensor, created on the device."""
    assert not include_miscellaneous_atoms, "miscellaneous atoms are outside the hot-path scope"
    for nt in ('ligand', 'receptor') + (('atom',) if all_atoms else ()):
        n = complex_graphs[nt].num_nodes
        complex_graphs[nt].node_t = {'tr': torch.full((n,), float(t_tr), device=device),
                                     'rot': torch.full((n,), float(t_rot), device=device),
                                     'tor': torch.full((n,), float(t_tor), device=device)}
    complex_graphs.complex_t = {'tr': torch.full((batchsize,), float(t_tr), device=device),
                                'rot': torch.full((batchsize,), float(t_rot), device=device),
                                'tor': torch.full((batchsize,), float(t_tor), device=device)}
