"""GPU unit tests of the sync-free graph construction entry points (include/diffdock_b200.h): ddb200_graph_fill against the
two-pass radius search + host-side assembly it replaces, ddb200_edge_embed against the library MLP, ddb200_csr_sort_by_target
against torch.sort(stable=True)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cloud(seed, sizes_x, sizes_y, spread=6.0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    x = torch.randn(sum(sizes_x), 3, device='cuda', generator=g) * spread
    y = torch.randn(sum(sizes_y), 3, device='cuda', generator=g) * spread
    bx = torch.repeat_interleave(torch.arange(len(sizes_x), device='cuda'), torch.tensor(sizes_x, device='cuda'))
    by = torch.repeat_interleave(torch.arange(len(sizes_y), device='cuda'), torch.tensor(sizes_y, device='cuda'))
    return x, y, bx, by


@pytest.mark.parametrize("per_graph", [False, True])
def test_graph_fill_forward_and_reverse_match_radius_and_sort(built_lib, per_graph):
    from diffdock_b200 import ops
    sizes_x, sizes_y = [50, 0, 37, 64], [7, 5, 9, 3]            # a complex without candidates, ragged sizes
    x, y, bx, by = _cloud(1, sizes_x, sizes_y)
    B = len(sizes_x)
    x_ptr, y_ptr = ops.segment_ptr(bx, B), ops.segment_ptr(by, B)
    rpg = (torch.tensor([9.0, 5.0, 12.0, 7.5], device='cuda') if per_graph else None)
    r = 1.0 if per_graph else 8.0
    row, col, cnt = ops.radius(x, y, x_ptr, by, r=r, r_per_graph=rpg, max_num_neighbors=10000)
    E = row.shape[0]
    assert E > 50
    by32, bx32 = by.int().contiguous(), bx.int().contiguous()
    c = ops.radius_count(x, y, x_ptr, by32, r=r, r_per_graph=rpg, max_num_neighbors=10000)
    assert torch.equal(c, cnt)
    incl = torch.cumsum(c, 0, dtype=torch.int32)
    cap = E + 13
    slot = torch.full((y.shape[0], max(sizes_x)), -1, dtype=torch.int32, device='cuda')
    frow, fcol, fvec, _, _ = ops.graph_fill(x, y, x_ptr, by32, (incl - c).contiguous(), cap, r=r, r_per_graph=rpg,
                                            max_num_neighbors=10000, slot_out=slot, slot_ld=slot.shape[1], col_offset=1000,
                                            fill_row=0)
    assert int(incl[-1]) == E
    assert torch.equal(frow[:E], row) and torch.equal(fcol[:E] - 1000, col)
    assert torch.equal(fvec[:E], x[col.long()] - y[row.long()])
    assert bool((frow[E:] == 0).all()) and bool((fcol[E:] == 0).all())
    # reverse direction: queries = x, candidates = y; same pairs, sorted by x; perm points into the forward list
    cr = ops.radius_count(y, x, y_ptr, bx32, r=r, r_per_graph=rpg, max_num_neighbors=1 << 30)
    incr = torch.cumsum(cr, 0, dtype=torch.int32)
    assert int(incr[-1]) == E
    rrow, rcol, _, _, perm = ops.graph_fill(y, x, y_ptr, bx32, (incr - cr).contiguous(), cap, r=r, r_per_graph=rpg,
                                            max_num_neighbors=1 << 30, want_vec=False, slot_in=slot, y_ptr=x_ptr,
                                            slot_ld=slot.shape[1], want_perm=True, row_offset=500)
    order = torch.sort(col, stable=True).indices                  # what the host-sized path does (cg_model.py)
    assert torch.equal(rrow[:E] - 500, col[order]) and torch.equal(rcol[:E], row[order])
    assert torch.equal(perm[:E].long(), order)


def test_graph_fill_static_edges_first_and_caps(built_lib):
    """Ligand graph: bond edges (CSR by target) listed before the radius hits, self excluded, cap 32 (+ self)."""
    from diffdock_b200 import ops
    sizes = [40, 25]
    x, _, bx, _ = _cloud(2, sizes, [1], spread=1.2)                # dense: the cap binds for many atoms
    n = x.shape[0]
    B = len(sizes)
    ptr = ops.segment_ptr(bx, B)
    g = torch.Generator().manual_seed(3)
    pre_cnt = torch.randint(0, 4, (n,), generator=g)
    pre_ptr = torch.zeros(n + 1, dtype=torch.int32)
    pre_ptr[1:] = torch.cumsum(pre_cnt, 0)
    pre_col = torch.cat([torch.randint(0, sizes[0], (int(pre_cnt[:sizes[0]].sum()),), generator=g),
                         sizes[0] + torch.randint(0, sizes[1], (int(pre_cnt[sizes[0]:].sum()),), generator=g)]).int()
    pre_ptr, pre_col, pre_cnt32 = pre_ptr.cuda(), pre_col.cuda(), pre_cnt.int().cuda()
    bx32 = bx.int().contiguous()
    centre, nbr, cnt = ops.radius(x, x, ptr, bx, r=5.0, max_num_neighbors=33, exclude_self=True)
    assert int(cnt.max()) in (32, 33)          # the cap binds (33 when the atom itself is not among its first 33 hits)
    tot = ops.radius_count(x, x, ptr, bx32, r=5.0, max_num_neighbors=33, exclude_self=True) + pre_cnt32
    incl = torch.cumsum(tot, 0, dtype=torch.int32)
    E = int(incl[-1])
    row, col, vec, eid, _ = ops.graph_fill(x, x, ptr, bx32, (incl - tot).contiguous(), E + 5, r=5.0, max_num_neighbors=33,
                                           exclude_self=True, pre_ptr=pre_ptr, pre_col=pre_col, want_eid=True, fill_row=0)
    # reference assembly: concatenate [static, radius] and sort stably by target
    s_tgt = torch.repeat_interleave(torch.arange(n, device='cuda'), pre_cnt32.long())
    all_tgt = torch.cat([s_tgt, centre.long()])
    all_src = torch.cat([pre_col.long(), nbr.long()])
    all_eid = torch.cat([torch.arange(pre_col.shape[0], device='cuda'), torch.full((centre.shape[0],), -1, device='cuda')])
    order = torch.sort(all_tgt, stable=True).indices
    assert torch.equal(row[:E].long(), all_tgt[order]) and torch.equal(col[:E].long(), all_src[order])
    assert torch.equal(eid[:E].long(), all_eid[order]) and bool((eid[E:] == -1).all())
    assert torch.equal(vec[:E], x[all_src[order]] - x[all_tgt[order]])


@pytest.mark.parametrize("D,ns", [(64, 48), (16, 16), (32, 24)])
def test_edge_embed_matches_library_mlp(built_lib, D, ns):
    from diffdock_b200 import ops
    from diffdock_b200.layers import GaussianSmearing
    g = torch.Generator(device='cuda').manual_seed(D + ns)
    S, n_nodes, cap, live = 2 * D // 2, 70, 5000, 4321
    torch.manual_seed(D)
    mlp = torch.nn.Sequential(torch.nn.Linear(S + D, ns), torch.nn.ReLU(), torch.nn.Dropout(0.0), torch.nn.Linear(ns, ns)).cuda()
    gs = GaussianSmearing(0.0, 80.0, D).cuda()
    sigma = torch.randn(n_nodes, S, device='cuda', generator=g)
    vec = torch.randn(cap, 3, device='cuda', generator=g) * 20
    row = torch.randint(0, n_nodes, (cap,), device='cuda', generator=g).int()
    n_dev = torch.tensor([live], dtype=torch.int32, device='cuda')
    with torch.no_grad():
        ref = mlp(torch.cat([sigma[row.long()], gs(vec.norm(dim=-1))], 1))
        u = torch.addmm(mlp[0].bias, sigma, mlp[0].weight[:, :S].t()).contiguous()
        out = torch.full((cap, ns), float('nan'), device='cuda')
        ops.edge_embed(vec, row, u, mlp[0].weight[:, S:].contiguous(), mlp[3].weight.contiguous(), mlp[3].bias.contiguous(),
                       gs.offset.contiguous(), float(gs.coeff), n_dev, out=out)
    torch.cuda.synchronize()
    err = float((out[:live] - ref[:live]).abs().max() / ref.abs().max())
    assert err < 2e-6, err
    assert bool(torch.isnan(out[live:]).all())                    # rows beyond the live count are never written


def test_csr_sort_by_target_is_stable(built_lib):
    from diffdock_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(9)
    for n, rows in ((0, 5), (1, 1), (1000, 7), (200000, 48000)):
        tgt = torch.randint(0, rows, (n,), device='cuda', generator=g).int()
        st, perm, rp = ops.csr_sort_by_target(tgt, rows, want_row_ptr=True)
        ref_t, ref_p = torch.sort(tgt.long(), stable=True)
        assert torch.equal(st.long(), ref_t) and torch.equal(perm, ref_p)
        cnt = torch.bincount(tgt.long(), minlength=rows)
        assert torch.equal(rp[1:].long() - rp[:-1].long(), cnt) and int(rp[0]) == 0
