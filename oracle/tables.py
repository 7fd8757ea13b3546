"""so3 / torus ``score_norm`` look-ups (utils/so3.py:6,89-93, utils/torus.py:25-26,79-83).  TEST INFRASTRUCTURE.

The two 1-D tables are produced by importing the unmodified reference modules
(tests/golden/make_tables.py) and stored once in ``diffdock_b200/tables/score_norm_tables.npz``;
``torus.score_norm_`` is an unseeded Monte-Carlo estimate in the reference (utils/torus.py:66-76), so oracle and
product must share one table instance - both read the same data file."""
import os

import numpy as np
import torch

_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'diffdock_b200', 'tables',
                     'score_norm_tables.npz')
_T = None


def tables():
    global _T
    if _T is None:
        z = np.load(_PATH)
        _T = {'so3': z['so3_exp_score_norms'].astype(np.float64), 'torus': z['torus_score_norm'].astype(np.float64)}
    return _T


SO3_MIN_EPS, SO3_MAX_EPS, SO3_N_EPS = 0.0005, 4, 2000
TORUS_SIGMA_MIN, TORUS_SIGMA_MAX, TORUS_SIGMA_N = 3e-3, 2, 5000


def so3_score_norm(eps):
    """utils/so3.py:89-93 (eps: CPU tensor)."""
    eps = eps.numpy()
    idx = (np.log10(eps) - np.log10(SO3_MIN_EPS)) / (np.log10(SO3_MAX_EPS) - np.log10(SO3_MIN_EPS)) * SO3_N_EPS
    idx = np.clip(np.around(idx).astype(int), a_min=0, a_max=SO3_N_EPS - 1)
    return torch.from_numpy(tables()['so3'][idx]).float()


def torus_score_norm(sigma):
    """utils/torus.py:79-83 (sigma: numpy array)."""
    s = np.log(sigma / np.pi)
    s = (s - np.log(TORUS_SIGMA_MIN)) / (np.log(TORUS_SIGMA_MAX) - np.log(TORUS_SIGMA_MIN)) * TORUS_SIGMA_N
    s = np.round(np.clip(s, 0, TORUS_SIGMA_N)).astype(int)
    return tables()['torus'][s]
