"""GPU: the indirection features of the fully fused convolution (edge_perm / vec_sign / ea_add / device-side edge count,
include/diffdock_b200.h:ddb200_fused_args) against the same kernel fed with materialised arrays, and against the oracle layer
(so the reverse direction of a bipartite graph, models/cg_model.py:556-557, and the per-call sigma-embedding add, :298-301,
are covered without building the whole model)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(seed, stage=3, n_nodes=300, E=3000, ns=48, nv=10):
    from diffdock_b200 import fused
    from diffdock_b200.tensor_layers import get_irrep_seq
    from diffdock_b200.tp_table import build_table
    seq = get_irrep_seq(ns, nv, False, False)
    t = build_table(seq[min(stage, 3)], '1x0e+1x1o+1x2e', seq[min(stage + 1, 3)], 'fctp')
    g = torch.Generator(device='cuda').manual_seed(seed)
    H = K1 = 3 * ns
    r = lambda *s: torch.randn(*s, device='cuda', generator=g)
    plan = fused.FusedPlan(t, r(H, K1) / K1 ** 0.5, 0.1 * r(H), r(t.weight_numel, H) / H ** 0.5, 0.1 * r(t.weight_numel))
    x = r(n_nodes, t.d_in)
    tgt = torch.sort(torch.randint(0, n_nodes, (E,), device='cuda', generator=g)).values.int()
    src = torch.randint(0, n_nodes, (E,), device='cuda', generator=g).int()
    return fused, t, plan, x, tgt, src, r, g


def _run(fused, plan, t, n_nodes, *a, **kw):
    out = torch.zeros(n_nodes, t.d_out, device='cuda')
    cnt = torch.zeros(n_nodes, device='cuda')
    fused.fused_conv(plan, *a, out, cnt, **kw)
    torch.cuda.synchronize()
    return out, cnt


def _close(a, b, tol=2e-5):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)) < tol


def test_perm_sign_add_match_materialised(built_lib):
    n_nodes, E, ns = 300, 3000, 48
    fused, t, plan, x, tgt, src, r, g = _setup(1, n_nodes=n_nodes, E=E)
    rows = 2 * E                                   # the attribute / vector store is larger than the edge list and permuted
    ea_store, vec_store, ew_store = r(rows, ns), r(rows, 3), torch.rand(rows, device='cuda', generator=g)
    perm = torch.randperm(rows, device='cuda', generator=g)[:E].int()
    add = r(7, ns)
    add_idx = torch.randint(0, 7, (E,), device='cuda', generator=g).int()
    got, gcnt = _run(fused, plan, t, n_nodes, ea_store, x, ns, tgt, src, x, vec_store, edge_weight=ew_store, edge_perm=perm,
                     vec_sign=-1.0, ea_add=add, ea_add_idx=add_idx)
    pl = perm.long()
    ea = (ea_store[pl] + add[add_idx.long()]).contiguous()
    ref, rcnt = _run(fused, plan, t, n_nodes, ea, x, ns, tgt, src, x, (-vec_store[pl]).contiguous(),
                     edge_weight=ew_store[pl].contiguous())
    assert _close(got, ref, 1e-5) and torch.equal(gcnt, rcnt)
    assert torch.equal(rcnt, torch.bincount(tgt.long(), minlength=n_nodes).float())


@pytest.mark.parametrize("live", [0, 1, 127, 129, 1000, 3000])
def test_device_side_edge_count(built_lib, live):
    """Arrays padded to a capacity, the live count in device memory: identical to a launch on the truncated arrays."""
    n_nodes, E, ns = 300, 3000, 48
    fused, t, plan, x, tgt, src, r, g = _setup(2, n_nodes=n_nodes, E=E)
    ea, vec = r(E, ns), r(E, 3)
    tgt_pad = tgt.clone()
    tgt_pad[live:] = 0                             # garbage beyond the live count must never be touched
    n_dev = torch.tensor([live], dtype=torch.int32, device='cuda')
    got, gcnt = _run(fused, plan, t, n_nodes, ea, x, ns, tgt_pad, src, x, vec, n_edges_dev=n_dev)
    if live == 0:
        assert float(got.abs().max()) == 0.0 and float(gcnt.abs().max()) == 0.0
        return
    ref, rcnt = _run(fused, plan, t, n_nodes, ea[:live].contiguous(), x, ns, tgt[:live].contiguous(), src[:live].contiguous(),
                     x, vec[:live].contiguous())
    assert _close(got, ref, 1e-5) and torch.equal(gcnt, rcnt)


@pytest.mark.parametrize("stage", [0, 3])
def test_unsorted_targets_and_long_runs(built_lib, stage):
    """The scatter stage reduces runs of equal targets; unsorted input (runs of length one) and runs longer than a warp
    must give the same sums as an index_add of the per-edge results."""
    n_nodes, E, ns = 50, 2000, 48
    fused, t, plan, x, tgt, src, r, g = _setup(3 + stage, stage=stage, n_nodes=n_nodes, E=E)
    ea, vec = r(E, ns), r(E, 3)
    sorted_out, _ = _run(fused, plan, t, n_nodes, ea, x, ns, tgt, src, x, vec)
    shuffle = torch.randperm(E, device='cuda', generator=g)
    got, cnt = _run(fused, plan, t, n_nodes, ea[shuffle].contiguous(), x, ns, tgt[shuffle].contiguous(),
                    src[shuffle].contiguous(), x, vec[shuffle].contiguous())
    assert _close(got, sorted_out, 2e-5)
    assert torch.equal(cnt, torch.bincount(tgt.long(), minlength=n_nodes).float())
    # every edge on one target (one run per warp).  The radial MLP reads node[tgt, :ns]: give every node the same scalars so
    # that moving the target does not change the per-edge messages, only where they are summed.
    xs = x.clone()
    xs[:, :ns] = xs[0, :ns]
    base, _ = _run(fused, plan, t, n_nodes, ea, xs, ns, tgt, src, xs, vec)
    one = torch.zeros(E, dtype=torch.int32, device='cuda') + 3
    got1, cnt1 = _run(fused, plan, t, n_nodes, ea, xs, ns, one, src, xs, vec)
    assert _close(got1[3], base.sum(0), 5e-5) and float(cnt1[3]) == E and float(got1[:3].abs().max()) == 0.0
