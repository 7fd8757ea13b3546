"""CPU: the oracle restatement vs fixtures recorded from the UNMODIFIED reference code (tests/golden/make_golden.py).
This is what pins the oracle: same inputs, same state_dict, fp32, tolerance 1e-5 relative
(different op order only)."""
import copy
from argparse import Namespace
from functools import partial

import numpy as np
import pytest
import torch

from diffdock_b200.hetero import collate, graph_from_dict
from tests.parity_helpers import golden_model, load_golden, rel_err

TOL = 1e-5


def test_layers_geometry_schedule():
    from oracle import diffusion as od, layers as ol
    g = load_golden('ref_layers.pt')
    gs = ol.GaussianSmearing(0.0, 5.0, 16)
    assert gs.coeff == g['gs_coeff'] and rel_err(gs(g['gs_in']), g['gs_out']) < 1e-6
    enc = ol.AtomEncoder(8, ([5, 3, 7], 0), sigma_embed_dim=4, lm_embedding_dim=6)
    enc.load_state_dict(g['enc_state'])
    assert rel_err(enc(g['enc_in']), g['enc_out']) < 1e-6
    assert rel_err(ol.sinusoidal_embedding(1000 * g['sin_in'], 16), g['sin_out']) < 1e-6
    assert np.allclose(od.get_t_schedule(20), g['sched20'], rtol=0, atol=1e-15)
    a = Namespace(tr_sigma_min=0.1, tr_sigma_max=19.0, rot_sigma_min=0.03, rot_sigma_max=1.55, tor_sigma_min=0.0314,
                  tor_sigma_max=3.14)
    assert np.allclose(od.t_to_sigma(0.3, 0.3, 0.3, a), g['t2s'], rtol=1e-15)
    assert rel_err(od.axis_angle_to_matrix(g['aa_in']), g['aa_out']) < 1e-6
    R, t = od.kabsch_batch(g['kabsch_A'], g['kabsch_B'])
    assert rel_err(R, g['kabsch_R']) < 1e-5 and rel_err(t, g['kabsch_t']) < 1e-5


def test_faster_tensor_product_pins_cg_convention():
    """Reference FasterTensorProduct (self-contained arithmetic) == oracle FasterTensorProduct == e3nn-lite FCTP with the
    weight rows permuted: pins the l<=1 Clebsch-Gordan signs/normalisation of the restated e3nn recipe."""
    from oracle import e3nn_lite as o3
    from oracle.tensor_layers import FasterTensorProduct
    from diffdock_b200.tp_table import build_table
    for c in load_golden('ref_faster_tp.pt'):
        tp = FasterTensorProduct(c['in_irreps'], '1x0e+1x1o', c['out_irreps'])
        assert rel_err(tp(c['x'], c['sh'], c['w']), c['out']) < 1e-6
        # same numbers from the generic Clebsch-Gordan route (FCTP lmax=1) after mapping the weight layout
        tf, tc = build_table(c['in_irreps'], '1x0e+1x1o', c['out_irreps'], 'faster'), \
            build_table(c['in_irreps'], '1x0e+1x1o', c['out_irreps'], 'fctp')
        key = lambda p: (p.i_in, p.i_sh, p.i_out)
        fpaths = {key(p): p for p in tf.paths}
        w_fctp = torch.zeros(c['w'].shape[0], tc.weight_numel)
        for p in tc.paths:
            q = fpaths[key(p)]
            n = p.mul_in * p.mul_out
            w_fctp[:, p.w_ref_off:p.w_ref_off + n] = c['w'][:, q.w_ref_off:q.w_ref_off + n]
        fctp = o3.FullyConnectedTensorProduct(c['in_irreps'], '1x0e+1x1o', c['out_irreps'])
        assert rel_err(fctp(c['x'], c['sh'], w_fctp), c['out']) < 1e-6


def _layer_from_case(c, cls):
    layer = cls(c['in_irreps'], c['sh_irreps'], c['out_irreps'], n_edge_features=12, hidden_features=12,
                residual=c['residual'], faster=c['faster'], edge_groups=c['groups']).eval()
    layer.load_state_dict({k: v for k, v in c['state'].items() if not k.startswith('tp.')}, strict=False)
    cuts = [0] + c['group_cuts'] + [c['edge_attr'].shape[0]]
    ea = [c['edge_attr'][cuts[i]:cuts[i + 1]] for i in range(4)] if c['groups'] == 4 else c['edge_attr']
    return layer, ea


def test_conv_layer_matches_reference():
    from oracle.tensor_layers import TensorProductConvLayer
    for c in load_golden('ref_conv_layer.pt'):
        layer, ea = _layer_from_case(c, TensorProductConvLayer)
        with torch.no_grad():
            out = layer(c['x'], c['edge_index'], ea, c['sh'], out_nodes=c['out_nodes'], reduce=c['reduce'],
                        edge_weight=c['edge_weight'])
        assert rel_err(out, c['out']) < TOL


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_cg_model_matches_reference(idx):
    from oracle.diffusion import set_time
    case = load_golden('ref_cg_model.pt')[idx]
    m, poses, a = golden_model(case, 'oracle')
    b = collate(poses)
    set_time(b, case['t'], case['t'], case['t'], len(poses), 'cpu')
    with torch.no_grad():
        tr, rot, tor, _ = m(b)
    assert rel_err(tr, case['tr']) < TOL and rel_err(rot, case['rot']) < TOL and rel_err(tor, case['tor']) < TOL


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_confidence_model_matches_reference(idx):
    """oracle/old_cg_model.py vs the reference's models/old_cg_model.py CGOldModel(confidence_mode) (fixture)."""
    from oracle.diffusion import set_time
    from tests.parity_helpers import golden_confidence_model
    case = load_golden('ref_confidence.pt')[idx]
    m, poses = golden_confidence_model(case, 'oracle')
    b = collate(poses)
    set_time(b, 0, 0, 0, len(poses), 'cpu')
    with torch.no_grad():
        conf = m(b)
    assert conf.shape == case['confidence'].shape
    assert (conf - case['confidence']).abs().max() < 1e-5 * max(1.0, float(case['confidence'].abs().max()))


@pytest.mark.parametrize("run", [0, 1])
def test_sampling_with_confidence_model_matches_reference(run):
    """oracle sampling() + confidence model (with / without confidence_model_args.crop_beyond) vs utils/sampling.py:208-227."""
    from oracle.sampling import sampling
    from oracle.diffusion import t_to_sigma
    from tests.parity_helpers import golden_confidence_model
    s = load_golden('ref_sampling_confidence.pt')
    r = s['runs'][run]
    score, poses, a = golden_model(load_golden('ref_cg_model.pt')[s['score_case']], 'oracle')
    conf, _ = golden_confidence_model(load_golden('ref_confidence.pt')[s['confidence_case']], 'oracle')
    torch.manual_seed(s['seed'])
    out, c = sampling(copy.deepcopy(poses), score, len(s['schedule']), s['schedule'], s['schedule'], s['schedule'], 'cpu',
                      partial(t_to_sigma, args=a), a, batch_size=3, no_final_step_noise=True, confidence_model=conf,
                      confidence_data_list=copy.deepcopy(poses),
                      confidence_model_args=Namespace(all_atoms=False, crop_beyond=r['crop_beyond']))
    for d, ref in zip(out, r['final_pos']):
        assert rel_err(d['ligand'].pos, ref) < 1e-4
    assert (c - r['confidence']).abs().max() < 1e-5, (c, r['confidence'])


def test_conformer_update_matches_reference():
    from oracle.diffusion import modify_conformer_batch
    c = load_golden('ref_conformer.pt')
    poses = [graph_from_dict(d) for d in c['poses']]
    b = collate(poses)
    mr = torch.from_numpy(poses[0]['ligand'].mask_rotate[0])
    assert rel_err(modify_conformer_batch(b['ligand'].pos, b, c['tr'], c['rot'], c['tor'], mr), c['new_pos']) < TOL
    assert rel_err(modify_conformer_batch(b['ligand'].pos, b, c['tr'], c['rot'], None, mr), c['rigid_pos']) < 1e-6


def test_sampling_trajectory_matches_reference():
    """4-step reverse diffusion with the reference's torch.normal draws (same seed, same call order) and the default
    low-temperature parameters: final ligand coordinates within 1e-4 relative (|x| ~ tens of Angstrom)."""
    from oracle.diffusion import t_to_sigma
    from oracle.sampling import sampling
    s = load_golden('ref_sampling.pt')
    case = load_golden('ref_cg_model.pt')[s['model_case']]
    m, poses, a = golden_model(case, 'oracle')
    torch.manual_seed(s['seed'])
    out, _ = sampling(copy.deepcopy(poses), m, s['steps'], s['schedule'], s['schedule'], s['schedule'], 'cpu',
                      partial(t_to_sigma, args=a), a, batch_size=3, no_final_step_noise=True,
                      temp_sampling=[1.170050527854316, 2.06391612594481, 7.044261621607846],
                      temp_psi=[0.727287304570729, 0.9022615585677628, 0.5946212391366862],
                      temp_sigma_data=[0.9299802531572672, 0.7464326999906034, 0.6943254174849822])
    for d, ref in zip(out, s['final_pos']):
        assert rel_err(d['ligand'].pos, ref) < 1e-4


def test_sampling_with_crop_beyond_matches_reference():
    """utils/sampling.py:104-109 + utils/utils.py:388-413: per-step receptor cropping (7..17 of 24 residues survive)."""
    from oracle.diffusion import t_to_sigma
    from oracle.sampling import sampling
    s = load_golden('ref_sampling_crop.pt')
    case = load_golden('ref_cg_model.pt')[s['model_case']]
    m, poses, a = golden_model(case, 'oracle')
    a.crop_beyond = s['crop_beyond']
    assert 0 < min(s['kept']) and max(s['kept']) < 24
    torch.manual_seed(s['seed'])
    out, _ = sampling(copy.deepcopy(poses), m, s['steps'], s['schedule'], s['schedule'], s['schedule'], 'cpu',
                      partial(t_to_sigma, args=a), a, batch_size=3, no_final_step_noise=True,
                      temp_sampling=[1.17, 2.06, 7.04], temp_psi=[0.73, 0.90, 0.59], temp_sigma_data=[0.93, 0.75, 0.69])
    for d, ref in zip(out, s['final_pos']):
        assert rel_err(d['ligand'].pos, ref) < 1e-4


def test_to_data_list_inverts_collate():
    case = load_golden('ref_cg_model.pt')[0]
    poses = [graph_from_dict(d) for d in case['poses']]
    back = collate(poses).to_data_list()
    for p, q in zip(poses, back):
        assert torch.equal(p['ligand'].pos, q['ligand'].pos) and torch.equal(p['ligand'].edge_mask, q['ligand'].edge_mask)
        assert torch.equal(p['receptor', 'receptor'].edge_index, q['receptor', 'receptor'].edge_index)
        assert torch.equal(p['ligand', 'ligand'].edge_index, q['ligand', 'ligand'].edge_index)


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_oracle_all_atom_confidence_matches_reference_fixture(idx):
    """oracle/old_aa_model.py vs the unmodified models/old_aa_model.py (tests/golden/make_golden_confidence_aa.py):
    nine convolutions per layer over ligand atoms, residues and receptor atoms, affinity head in case 2."""
    import copy
    from diffdock_b200.hetero import collate
    from oracle.diffusion import set_time
    from tests.parity_helpers import golden_confidence_model, load_golden
    case = load_golden('ref_confidence_aa.pt')[idx]
    m, poses = golden_confidence_model(case, 'oracle', all_atoms=True)
    b = collate(copy.deepcopy(poses))
    set_time(b, 0, 0, 0, len(poses), 'cpu', all_atoms=True)
    with torch.no_grad():
        conf = m(b)
    ref = case['confidence']
    assert conf.shape == ref.shape
    assert float((conf - ref).abs().max()) < 1e-5 * max(1.0, float(ref.abs().max()))


def test_fourier_time_embedding_matches_reference_fixture():
    """utils/diffusion_utils.py:113-135 (GaussianFourierProjection): same RNG draw for W at the same seed, same output -
    oracle and product."""
    from tests.parity_helpers import load_golden
    from oracle.layers import get_timestep_embedding as o_emb
    from diffdock_b200.diffusion_utils import get_timestep_embedding as p_emb
    ref = load_golden('ref_fourier.pt')
    for make in (o_emb, p_emb):
        torch.manual_seed(ref['seed'])
        emb = make('fourier', ref['dim'], ref['scale'])
        assert torch.equal(emb.W, ref['W']) and list(emb.state_dict().keys()) == ['W']
        assert float((emb(ref['x']) - ref['out']).abs().max()) == 0.0


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_oracle_all_atom_score_model_matches_reference_fixture(idx):
    """oracle/aa_model.py vs the unmodified models/aa_model.py (tests/golden/make_golden_aa_model.py): nine edge groups over
    ligand atoms, residues and receptor atoms (lmax 2; lmax 1 with one shared radial MLP; protein embedding layer)."""
    import copy
    from diffdock_b200.hetero import collate
    from oracle.diffusion import set_time
    from tests.parity_helpers import golden_model, load_golden
    case = load_golden('ref_aa_model.pt')[idx]
    m, poses, a = golden_model(case, 'oracle', all_atoms=True)
    b = collate(copy.deepcopy(poses))
    t = case['t']
    set_time(b, t, t, t, len(poses), 'cpu', all_atoms=True)
    with torch.no_grad():
        tr, rot, tor, _ = m(b)
    for got, key in ((tr, 'tr'), (rot, 'rot'), (tor, 'tor')):
        ref = case[key]
        assert got.shape == ref.shape and float((got - ref).abs().max() / ref.abs().max()) < 1e-5
