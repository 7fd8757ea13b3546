"""Generates diffdock_b200/tables/score_norm_tables.npz by importing the UNMODIFIED reference modules
utils/so3.py and utils/torus.py (run in the authoring container; /root/reference is not on the GPU box).

    cd /tmp/some_scratch_dir && python /root/repo/tests/golden/make_tables.py

The modules cache multi-hundred-MB .npy files in the current directory and take ~15 minutes on 8 cores.
utils/torus.py:66-76 estimates ``score_norm_`` by unseeded Monte Carlo; the numpy seed is fixed here so that the
stored instance is reproducible.  Oracle and product read the same stored instance.
"""
import os
import sys

import numpy as np

sys.path.insert(0, '/root/reference')
np.random.seed(0)
from utils import so3    # noqa: E402

np.random.seed(0)
from utils import torus  # noqa: E402

out = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'diffdock_b200',
                   'tables', 'score_norm_tables.npz')
np.savez(out, so3_exp_score_norms=np.asarray(so3._exp_score_norms), torus_score_norm=np.asarray(torus.score_norm_))
print('wrote', out)
