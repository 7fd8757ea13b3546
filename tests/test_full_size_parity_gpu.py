"""GPU vs CPU oracle at the BENCHMARKED size: one pose of the 1500-residue / 40-atom CFG-L2 complex (48 000-node batches in
bench.py are 32-40 copies of exactly this graph), product scores against the oracle's, tolerance 1e-4 (north_star).
The oracle forward of this size takes ~20-40 s on the CPU."""
import pytest
import torch

from tests.parity_helpers import make_model_pair, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("t", [0.5])
def test_one_full_size_pose_matches_oracle(built_lib, t):
    from diffdock_b200.synthetic import default_model_args, make_pose_list
    from diffdock_b200.hetero import collate
    from diffdock_b200.diffusion_utils import set_time
    from oracle.diffusion import set_time as o_set_time
    args = default_model_args()                       # ns=48, nv=10, sh_lmax=2, 6 conv layers, 64-dim embeddings
    o, p = make_model_pair(args, seed=0)
    poses = make_pose_list(1, n_res=1500, n_atoms=40, seed=100, tr_sigma_max=args.tr_sigma_max)
    g = collate(poses).to('cuda:0')
    set_time(g, None, t, t, t, 1, False, 'cuda:0')
    got = p(g)
    torch.cuda.synchronize()
    g_cpu = collate(poses)
    o_set_time(g_cpu, t, t, t, 1, 'cpu')
    torch.set_num_threads(min(torch.get_num_threads(), 32))
    with torch.no_grad():
        ref = o(g_cpu)
    errs = [rel_err(a, b) for a, b in zip(got[:3], ref[:3]) if b.numel()]
    assert max(errs) < 1e-4, errs
