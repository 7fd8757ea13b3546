"""Shared helpers of the parity tests: seeded random layers/graphs evaluated by the oracle (CPU) and the product (CUDA)."""
import torch


def rand_bn_(bn, gen):
    """Non-trivial eval-mode statistics."""
    with torch.no_grad():
        bn.running_mean.copy_(0.1 * torch.randn(bn.running_mean.shape, generator=gen))
        bn.running_var.copy_(0.5 + torch.rand(bn.running_var.shape, generator=gen))
        bn.weight.copy_(1.0 + 0.2 * torch.randn(bn.weight.shape, generator=gen))
        bn.bias.copy_(0.1 * torch.randn(bn.bias.shape, generator=gen))


def make_layer_pair(in_irreps, sh_irreps, out_irreps, n_edge_features, seed=0, **kw):
    """(oracle layer on CPU, product layer with identical parameters)."""
    from oracle.tensor_layers import TensorProductConvLayer as OLayer
    from diffdock_b200.tensor_layers import TensorProductConvLayer as PLayer
    torch.manual_seed(seed)
    o = OLayer(in_irreps, sh_irreps, out_irreps, n_edge_features, **kw).eval()
    gen = torch.Generator().manual_seed(seed + 1)
    if o.batch_norm is not None:
        rand_bn_(o.batch_norm, gen)
    p = PLayer(in_irreps, sh_irreps, out_irreps, n_edge_features, **kw).eval()
    sd = {k: v for k, v in o.state_dict().items() if not k.startswith('tp.')}
    missing, unexpected = p.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return o, p


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def layer_parity_case(seed=0, n_nodes=64, n_edges=700, ns=48, nv=10, lmax=2, stage=3, groups=1, faster=False,
                      device='cuda:0', reduce='mean', use_vec=True, edge_weight_tensor=False, out_nodes=None,
                      residual=True):
    from oracle import e3nn_lite as o3
    from oracle.tensor_layers import get_irrep_seq
    seq = get_irrep_seq(ns, nv, False, False)
    sh_irreps = str(o3.Irreps.spherical_harmonics(lmax))
    o, p = make_layer_pair(seq[min(stage, 3)], sh_irreps, seq[min(stage + 1, 3)], 3 * ns, seed=seed,
                           hidden_features=3 * ns, edge_groups=groups, faster=faster, residual=residual)
    g = torch.Generator().manual_seed(seed + 2)
    x = torch.randn(n_nodes, o3.Irreps(seq[min(stage, 3)]).dim, generator=g)
    n_tgt = out_nodes or n_nodes
    ei = torch.stack([torch.randint(0, n_tgt, (n_edges,), generator=g), torch.randint(0, n_nodes, (n_edges,), generator=g)])
    vec = torch.randn(n_edges, 3, generator=g)
    ea = torch.randn(n_edges, 3 * ns, generator=g)
    sh = o3.spherical_harmonics(o3.Irreps(sh_irreps), vec, normalize=True, normalization='component')
    ew = torch.rand(n_edges, 1, generator=g) if edge_weight_tensor else 1.0
    if groups > 1:
        cuts = sorted(torch.randint(0, n_edges, (groups - 1,), generator=g).tolist())
        b = [0] + cuts + [n_edges]
        ea_o = [ea[b[i]:b[i + 1]] for i in range(groups)]
    else:
        ea_o = ea
    with torch.no_grad():
        ref = o(x, ei, ea_o, sh, out_nodes=out_nodes, reduce=reduce, edge_weight=ew)
    p = p.to(device)
    dev = lambda t: t.to(device) if torch.is_tensor(t) else t
    ea_p = [dev(a) for a in ea_o] if groups > 1 else dev(ea)
    got = p(dev(x), dev(ei), ea_p, dev(sh), out_nodes=out_nodes, reduce=reduce, edge_weight=dev(ew),
            edge_vec=dev(vec) if use_vec else None)
    torch.cuda.synchronize()
    return rel_err(got, ref)


def make_model_pair(args, seed=0, lm=True, product=True):
    """(oracle CGModel on CPU, product CGModel) sharing one random state_dict (BatchNorm statistics randomised)."""
    from functools import partial
    from oracle.cg_model import CGModel as OModel
    from oracle.layers import get_timestep_embedding as o_emb
    from oracle.diffusion import t_to_sigma as o_t2s
    from diffdock_b200.cg_model import CGModel as PModel
    from diffdock_b200.diffusion_utils import get_timestep_embedding as p_emb, t_to_sigma as p_t2s
    kw = dict(sigma_embed_dim=args.sigma_embed_dim, sh_lmax=args.sh_lmax, ns=args.ns, nv=args.nv,
              num_conv_layers=args.num_conv_layers, lig_max_radius=args.max_radius, rec_max_radius=args.rec_max_radius,
              cross_max_distance=args.cross_max_distance, center_max_distance=args.center_max_distance,
              distance_embed_dim=args.distance_embed_dim, cross_distance_embed_dim=args.cross_distance_embed_dim,
              dynamic_max_cross=args.dynamic_max_cross, lm_embedding_type='precomputed' if lm else None,
              embed_also_ligand=True, num_prot_emb_layers=args.num_prot_emb_layers,
              use_second_order_repr=args.use_second_order_repr, no_torsion=args.no_torsion,
              smooth_edges=args.smooth_edges, fixed_center_conv=args.fixed_center_conv,
              reduce_pseudoscalars=args.reduce_pseudoscalars,
              differentiate_convolutions=args.differentiate_convolutions)
    torch.manual_seed(seed)
    o = OModel(partial(o_t2s, args=args), 'cpu', o_emb('sinusoidal', args.sigma_embed_dim, args.embedding_scale), **kw).eval()
    gen = torch.Generator().manual_seed(seed + 1)
    for m in o.modules():
        if m.__class__.__name__ == 'BatchNorm':
            rand_bn_(m, gen)
    if not product:
        return o, None
    p = PModel(partial(p_t2s, args=args), torch.device('cuda:0'),
               p_emb('sinusoidal', args.sigma_embed_dim, args.embedding_scale), **kw).eval()
    p.load_state_dict(o.state_dict(), strict=True)     # includes e3nn-style tp.* buffers, which must be accepted
    return o, p.to('cuda:0')


def model_parity_case(seed=0, lmax=2, ns=16, nv=4, n_layers=3, emb=16, n_res=60, n_atoms=12, n_poses=3, t=0.5,
                      far_poses=(), run_product=True, **over):
    from diffdock_b200.synthetic import default_model_args, make_pose_list
    from diffdock_b200.hetero import collate
    from oracle.diffusion import set_time
    args = default_model_args(ns=ns, nv=nv, sh_lmax=lmax, num_conv_layers=n_layers, distance_embed_dim=emb,
                              cross_distance_embed_dim=emb, sigma_embed_dim=emb, **over)
    o, p = make_model_pair(args, seed, product=run_product)
    poses = make_pose_list(n_poses, n_res=n_res, n_atoms=n_atoms, seed=seed + 3, tr_sigma_max=args.tr_sigma_max * t)
    for i in far_poses:       # ligand moved out of every cross cut-off: that complex has no ligand-receptor edges
        # (80 A: beyond receptor radius + cut-off for the sizes used, yet small enough that the fp32 centroid sum - whose
        # order differs between CPU and GPU atomics - keeps the 1e-4 tolerance; at 500 A it is borderline)
        poses[i]['ligand'].pos = poses[i]['ligand'].pos + 80.0
    g_cpu = collate(poses)
    set_time(g_cpu, t, t, t, n_poses, 'cpu')
    if not run_product:
        with torch.no_grad():
            return o(g_cpu)
    g_gpu = collate(poses).to('cuda:0')
    set_time(g_gpu, t, t, t, n_poses, 'cuda:0')
    with torch.no_grad():
        ref = o(g_cpu)
    got = p(g_gpu)
    torch.cuda.synchronize()
    errs = {'tr': rel_err(got[0], ref[0]), 'rot': rel_err(got[1], ref[1]), 'tor_numel': int(ref[2].numel())}
    assert got[2].numel() == ref[2].numel()
    if ref[2].numel():
        errs['tor'] = rel_err(got[2], ref[2])
    errs = {k: v for k, v in errs.items()}
    if errs['tor_numel'] > 0:
        errs.pop('tor_numel')
    return errs


# ---------------------------------------------------------------------------------------------- golden fixtures
import os as _os

GOLDEN = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), 'golden')


def load_golden(name):
    return torch.load(_os.path.join(GOLDEN, name), weights_only=False)


def golden_model(case, which, all_atoms=False):
    """Model ('oracle' on CPU | 'product' on cuda:0) + pose list rebuilt from a ref_cg_model.pt case (``all_atoms``: a
    ref_aa_model.pt case, models/aa_model.py)."""
    from argparse import Namespace
    from functools import partial
    from diffdock_b200.hetero import graph_from_dict
    a = Namespace(**case['args'])
    if which == 'oracle':
        if all_atoms:
            from oracle.aa_model import AAModel as CGModel
        else:
            from oracle.cg_model import CGModel
        from oracle.layers import get_timestep_embedding
        from oracle.diffusion import t_to_sigma
        dev = 'cpu'
    else:
        if all_atoms:
            from diffdock_b200.aa_model import AAModel as CGModel
        else:
            from diffdock_b200.cg_model import CGModel
        from diffdock_b200.diffusion_utils import get_timestep_embedding, t_to_sigma
        dev = torch.device('cuda:0')
    m = CGModel(partial(t_to_sigma, args=a), dev, get_timestep_embedding('sinusoidal', 8, a.embedding_scale),
                **case['kw']).eval()
    if case['lm_dim']:   # the fixture shrinks the 1280-wide LM embedding to 16 columns (see make_golden.py)
        m.rec_node_embedding.additional_features_dim = case['lm_dim']
        m.rec_node_embedding.additional_features_embedder = torch.nn.Linear(case['lm_dim'] + 6, 6)
    m.load_state_dict(case['state'], strict=True)
    poses = [graph_from_dict(d) for d in case['poses']]
    return m.to(dev), poses, a


def golden_confidence_model(case, which, all_atoms=False):
    """Confidence model ('oracle' on CPU | 'product' on cuda:0) + pose list rebuilt from a ref_confidence.pt case
    (``all_atoms``: a ref_confidence_aa.pt case, models/old_aa_model.py)."""
    from functools import partial
    from diffdock_b200.hetero import graph_from_dict
    from diffdock_b200.synthetic import default_model_args
    a = default_model_args()
    if which == 'oracle':
        if all_atoms:
            from oracle.old_aa_model import AAOldModel as CGOldModel
        else:
            from oracle.old_cg_model import CGOldModel
        from oracle.layers import get_timestep_embedding
        from oracle.diffusion import t_to_sigma
        dev = 'cpu'
    else:
        if all_atoms:
            from diffdock_b200.old_aa_model import AAOldModel as CGOldModel
        else:
            from diffdock_b200.old_cg_model import CGOldModel
        from diffdock_b200.diffusion_utils import get_timestep_embedding, t_to_sigma
        dev = torch.device('cuda:0')
    kw = dict(case['kw'])
    if case['lm_dim']:
        kw['lm_embedding_dim'] = case['lm_dim']     # the fixture shrinks the 1280-wide LM embedding to 16 columns
    m = CGOldModel(partial(t_to_sigma, args=a), dev, get_timestep_embedding('sinusoidal', 8, a.embedding_scale), **kw).eval()
    m.load_state_dict(case['state'], strict=True)
    return m.to(dev), [graph_from_dict(d) for d in case['poses']]


def canonical_contact_edges(edge_index, coords):
    """Contact-graph edge list [2, E] (rows [neighbour, centre], centre by centre) with TIES made canonical: where a centre's
    neighbours are listed by distance (more hits than max_neighbors: np.argsort order, datasets/process_mols.py:184), runs of
    exactly equal fp32 distance are re-ordered by index.  np.argsort's default sort is not stable and its tie order depends on
    the numpy build (AVX-512 quicksort vs introsort), so that is the strongest order a restatement can reproduce."""
    import numpy as np
    from oracle.inputs import cdist_sq_f32
    ei = np.asarray(edge_index)
    sq = cdist_sq_f32(np.asarray(coords, dtype=np.float32))
    out = ei.copy()
    start = 0
    E = ei.shape[1]
    while start < E:
        end = start
        while end < E and ei[1, end] == ei[1, start]:
            end += 1
        nb = ei[0, start:end]
        if not np.all(np.diff(nb) > 0):            # listed by distance
            i = ei[1, start]
            order = sorted(range(end - start), key=lambda q: (sq[i, nb[q]], nb[q]))
            out[0, start:end] = nb[order]
        start = end
    return out
