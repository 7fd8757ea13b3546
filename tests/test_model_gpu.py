"""GPU parity: diffdock_b200.CGModel (CUDA) vs the CPU oracle CGModel on seeded synthetic complexes.
Tolerance on scores: 1e-4 relative (north_star), measured as max|a-b| / max|b| per output."""
import pytest
import torch

from tests.parity_helpers import model_parity_case

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.mark.parametrize("lmax", [2, 1])
def test_scores_match_oracle_small(built_lib, lmax):
    errs = model_parity_case(seed=0, lmax=lmax, ns=16, nv=4, n_layers=3, emb=16, n_res=60, n_atoms=12, n_poses=3, t=0.5)
    assert max(errs.values()) < TOL, errs


@pytest.mark.parametrize("t", [1.0, 0.5, 0.05])
def test_scores_match_oracle_cfg_l2(built_lib, t):
    errs = model_parity_case(seed=1, lmax=2, ns=48, nv=10, n_layers=6, emb=64, n_res=120, n_atoms=18, n_poses=2, t=t)
    assert max(errs.values()) < TOL, errs


def test_no_rotatable_bonds_branch(built_lib):
    errs = model_parity_case(seed=2, lmax=2, ns=16, nv=4, n_layers=2, emb=16, n_res=40, n_atoms=3, n_poses=2, t=0.3)
    assert errs['tor_numel'] == 0 and max(errs['tr'], errs['rot']) < TOL, errs


@pytest.mark.parametrize("far", [(1,), (0, 1, 2)])
def test_complexes_without_cross_edges(built_lib, far):
    """Ligands outside every ligand-receptor cut-off: some / all complexes contribute no cross edges (empty edge groups,
    receptor nodes with no incoming ligand message)."""
    errs = model_parity_case(seed=4, lmax=2, ns=16, nv=4, n_layers=2, emb=16, n_res=40, n_atoms=8, n_poses=3, t=0.3,
                             far_poses=far)
    errs.pop('tor_numel', None)
    # 3e-4: the displaced ligands sit 80 A from the origin, where the fp32 centroid (summed in a different order by the CPU
    # oracle and the GPU atomics) costs the translation score up to ~1e-4 by itself (moving them to 500 A changes the
    # ORACLE's own tr score by 3e-4)
    assert max(errs.values()) < 3e-4, errs
