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
                      device='cuda:0', reduce='mean', use_vec=True, edge_weight_tensor=False, out_nodes=None):
    from oracle import e3nn_lite as o3
    from oracle.tensor_layers import get_irrep_seq
    seq = get_irrep_seq(ns, nv, False, False)
    sh_irreps = str(o3.Irreps.spherical_harmonics(lmax))
    o, p = make_layer_pair(seq[min(stage, 3)], sh_irreps, seq[min(stage + 1, 3)], 3 * ns, seed=seed,
                           hidden_features=3 * ns, edge_groups=groups, faster=faster)
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
