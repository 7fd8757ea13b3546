"""Golden vector for the 'fourier' time embedding: the UNMODIFIED utils/diffusion_utils.py:get_timestep_embedding('fourier')
of /root/reference (GaussianFourierProjection, :113-135), seeded, stored as tests/golden/ref_fourier.pt.
    cd /tmp && python /root/repo/tests/golden/make_golden_fourier.py 2>/dev/null"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install()
sys.path.insert(0, '/root/reference')
import utils.diffusion_utils as r_du        # noqa: E402

torch.manual_seed(11)
emb = r_du.get_timestep_embedding('fourier', 32, 1000)
x = torch.rand(9, generator=torch.Generator().manual_seed(12))
torch.save(dict(seed=11, dim=32, scale=1000, W=emb.W.detach().clone(), x=x, out=emb(x).detach()),
           os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_fourier.pt'))
print('ref_fourier.pt written')
