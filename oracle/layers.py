"""Oracle restatement of models/layers.py and of the time embeddings in utils/diffusion_utils.py.
TEST INFRASTRUCTURE (see oracle/__init__.py)."""
import math

import numpy as np
import torch
from torch import nn


def fc_block(in_dim, hidden_dim, out_dim, layers, dropout=0.0, activation='relu'):
    """models/layers.py:10-17 (FCBlock): Linear-act-Dropout [x(layers-2)] ... Linear.  Index layout of the
    nn.Sequential (.0, .3, .6 ...) is kept so state_dicts are interchangeable."""
    act = {'relu': nn.ReLU, 'silu': nn.SiLU}[activation]
    assert layers >= 2
    seq = [nn.Linear(in_dim, hidden_dim), act(), nn.Dropout(dropout)]
    for _ in range(layers - 2):
        seq += [nn.Linear(hidden_dim, hidden_dim), act(), nn.Dropout(dropout)]
    seq += [nn.Linear(hidden_dim, out_dim)]
    return nn.Sequential(*seq)


class GaussianSmearing(nn.Module):
    """models/layers.py:20-30.  coeff is a Python float taken from the fp32 linspace (hazard C.10)."""

    def __init__(self, start=0.0, stop=5.0, num_gaussians=50):
        super().__init__()
        offset = torch.linspace(start, stop, num_gaussians)
        self.coeff = -0.5 / (offset[1] - offset[0]).item() ** 2
        self.register_buffer('offset', offset)

    def forward(self, dist):
        d = dist.reshape(-1, 1) - self.offset.reshape(1, -1)
        return torch.exp(self.coeff * d * d)


class AtomEncoder(nn.Module):
    """models/layers.py:33-67: sum of categorical embeddings, then Linear on [emb | scalar/LM/sigma feats]."""

    def __init__(self, emb_dim, feature_dims, sigma_embed_dim, lm_embedding_dim=0):
        super().__init__()
        self.atom_embedding_list = nn.ModuleList()
        self.num_categorical_features = len(feature_dims[0])
        self.additional_features_dim = feature_dims[1] + sigma_embed_dim + lm_embedding_dim
        for dim in feature_dims[0]:
            emb = nn.Embedding(dim, emb_dim)
            nn.init.xavier_uniform_(emb.weight.data)
            self.atom_embedding_list.append(emb)
        if self.additional_features_dim > 0:
            self.additional_features_embedder = nn.Linear(self.additional_features_dim + emb_dim, emb_dim)

    def forward(self, x):
        assert x.shape[1] == self.num_categorical_features + self.additional_features_dim
        e = 0
        for i in range(self.num_categorical_features):
            e = e + self.atom_embedding_list[i](x[:, i].long())
        if self.additional_features_dim > 0:
            e = self.additional_features_embedder(torch.cat([e, x[:, self.num_categorical_features:].to(e.dtype)], 1))
        return e


class OldAtomEncoder(nn.Module):
    """models/layers.py:70-117: categorical embeddings + Linear(scalar features incl. sigma embedding), then an optional
    Linear([emb | 1280 LM columns]) - the encoder of the confidence model (use_old_atom_encoder)."""

    def __init__(self, emb_dim, feature_dims, sigma_embed_dim, lm_embedding_type=None, lm_embedding_dim=1280):
        super().__init__()
        self.atom_embedding_list = nn.ModuleList()
        self.num_categorical_features = len(feature_dims[0])
        self.num_scalar_features = feature_dims[1] + sigma_embed_dim
        self.lm_embedding_type = lm_embedding_type
        for dim in feature_dims[0]:
            emb = nn.Embedding(dim, emb_dim)
            nn.init.xavier_uniform_(emb.weight.data)
            self.atom_embedding_list.append(emb)
        if self.num_scalar_features > 0:
            self.linear = nn.Linear(self.num_scalar_features, emb_dim)
        if lm_embedding_type is not None:
            if lm_embedding_type != 'esm':
                raise ValueError('LM Embedding type was not correctly determined. LM embedding type: ', lm_embedding_type)
            self.lm_embedding_dim = lm_embedding_dim
            self.lm_embedding_layer = nn.Linear(self.lm_embedding_dim + emb_dim, emb_dim)

    def forward(self, x):
        nc, nsf = self.num_categorical_features, self.num_scalar_features
        assert x.shape[1] == nc + nsf + (self.lm_embedding_dim if self.lm_embedding_type is not None else 0)
        out = 0
        for i in range(nc):
            out = out + self.atom_embedding_list[i](x[:, i].long())
        if nsf > 0:
            out = out + self.linear(x[:, nc:nc + nsf])
        if self.lm_embedding_type is not None:
            out = self.lm_embedding_layer(torch.cat([out, x[:, -self.lm_embedding_dim:]], 1))
        return out


def sinusoidal_embedding(timesteps, embedding_dim, max_positions=10000):
    """utils/diffusion_utils.py:99-110."""
    assert timesteps.dim() == 1
    half = embedding_dim // 2
    k = math.log(max_positions) / (half - 1)
    freq = torch.exp(torch.arange(half, dtype=torch.float32, device=timesteps.device) * -k)
    emb = timesteps.float()[:, None] * freq[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], 1)
    if embedding_dim % 2 == 1:
        emb = torch.nn.functional.pad(emb, (0, 1))
    return emb


class GaussianFourierProjection(nn.Module):
    """utils/diffusion_utils.py:113-125."""

    def __init__(self, embedding_size=256, scale=1.0):
        super().__init__()
        self.W = nn.Parameter(torch.randn(embedding_size // 2) * scale, requires_grad=False)

    def forward(self, x):
        x_proj = x[:, None] * self.W[None, :] * 2 * np.pi
        return torch.cat([torch.sin(x_proj), torch.cos(x_proj)], dim=-1)


def get_timestep_embedding(embedding_type, embedding_dim, embedding_scale=10000):
    """utils/diffusion_utils.py:128-135."""
    if embedding_type == 'sinusoidal':
        return lambda x: sinusoidal_embedding(embedding_scale * x, embedding_dim)
    assert embedding_type == 'fourier'
    return GaussianFourierProjection(embedding_size=embedding_dim, scale=embedding_scale)
