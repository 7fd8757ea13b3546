"""Oracle restatement of the sampler-step arithmetic: utils/diffusion_utils.py:28-32,60-78,138-168,
utils/torsion.py:75-90, utils/geometry.py:7-86,246-276.  TEST INFRASTRUCTURE."""
import numpy as np
import torch
from scipy.stats import beta


def t_to_sigma(t_tr, t_rot, t_tor, args):
    """utils/diffusion_utils.py:28-32."""
    return (args.tr_sigma_min ** (1 - t_tr) * args.tr_sigma_max ** t_tr,
            args.rot_sigma_min ** (1 - t_rot) * args.rot_sigma_max ** t_rot,
            args.tor_sigma_min ** (1 - t_tor) * args.tor_sigma_max ** t_tor)


def get_t_schedule(inference_steps, inf_sched_alpha=1, inf_sched_beta=1, t_max=1):
    """utils/diffusion_utils.py:138-143 ('expbeta')."""
    lin_max = beta.cdf(t_max, a=inf_sched_alpha, b=inf_sched_beta)
    c = np.linspace(lin_max, 0, inference_steps + 1)[:-1]
    return beta.ppf(c, a=inf_sched_alpha, b=inf_sched_beta)


def set_time(g, t_tr, t_rot, t_tor, batchsize, device, all_atoms=False):
    """utils/diffusion_utils.py:146-168."""
    for nt in ('ligand', 'receptor') + (('atom',) if all_atoms else ()):
        n = g[nt].num_nodes
        g[nt].node_t = {'tr': t_tr * torch.ones(n).to(device), 'rot': t_rot * torch.ones(n).to(device),
                        'tor': t_tor * torch.ones(n).to(device)}
    g.complex_t = {'tr': t_tr * torch.ones(batchsize).to(device), 'rot': t_rot * torch.ones(batchsize).to(device),
                   'tor': t_tor * torch.ones(batchsize).to(device)}


def axis_angle_to_matrix(aa):
    """utils/geometry.py:7-86 (pytorch3d): axis-angle -> quaternion (small-angle series below 1e-6) -> matrix."""
    ang = torch.norm(aa, p=2, dim=-1, keepdim=True)
    half = 0.5 * ang
    small = ang.abs() < 1e-6
    s = torch.where(small, 0.5 - ang * ang / 48, torch.sin(half) / torch.where(small, torch.ones_like(ang), ang))
    q = torch.cat([torch.cos(half), aa * s], -1)
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def torsion_update_batch(pos, rot_bonds, mask_rotate, torsion_updates):
    """utils/torsion.py:75-90: sequential rotations, each about pos[u]-pos[v] through pos[v] of the atoms in
    mask_rotate[r]; later bonds see already-updated positions."""
    pos = pos + 0
    for r, e in enumerate(rot_bonds):
        u, v = int(e[0]), int(e[1])
        assert not mask_rotate[r, u] and mask_rotate[r, v]
        ax = pos[:, u] - pos[:, v]
        R = axis_angle_to_matrix(ax / torch.linalg.norm(ax, dim=-1, keepdims=True) * torsion_updates[:, r:r + 1])
        m = mask_rotate[r]
        pos[:, m] = torch.bmm(pos[:, m] - pos[:, v:v + 1], R.transpose(1, 2)) + pos[:, v:v + 1]
    return pos


def kabsch_batch(A, B):
    """utils/geometry.py:246-276: R,t with R@A+t ~= B  (A,B: [B,N,3])."""
    A, B = A.permute(0, 2, 1), B.permute(0, 2, 1)
    cA, cB = A.mean(2, keepdim=True), B.mean(2, keepdim=True)
    H = torch.bmm(A - cA, (B - cB).transpose(1, 2))
    U, S, Vt = torch.linalg.svd(H)
    R = torch.bmm(Vt.transpose(1, 2), U.transpose(1, 2))
    SS = torch.diag(torch.tensor([1., 1., -1.], dtype=A.dtype, device=A.device))
    Rm = torch.bmm(Vt.transpose(1, 2) @ SS, U.transpose(1, 2))
    R = torch.where(torch.linalg.det(R)[:, None, None] < 0, Rm, R)
    return R, torch.bmm(-R, cA) + cB


def modify_conformer_batch(orig_pos, data, tr_update, rot_update, torsion_updates, mask_rotate):
    """utils/diffusion_utils.py:60-78."""
    B = data.num_graphs
    N = data['ligand'].num_nodes // B
    M = data['ligand', 'ligand'].num_edges // B
    pos = orig_pos.reshape(B, N, 3) + 0
    edge_index = data['ligand', 'ligand'].edge_index[:, :M]
    edge_mask = data['ligand'].edge_mask[:M]
    center = pos.mean(1, keepdim=True)
    rigid = torch.bmm(pos - center, axis_angle_to_matrix(rot_update).permute(0, 2, 1)) + tr_update.unsqueeze(1) + center
    if torsion_updates is None:
        return rigid.reshape(-1, 3)
    flex = torsion_update_batch(rigid, edge_index.T[edge_mask], mask_rotate, torsion_updates.reshape(B, -1))
    R, t = kabsch_batch(flex, rigid)
    return (torch.bmm(flex, R.transpose(1, 2)) + t.transpose(1, 2)).reshape(-1, 3)


def crop_beyond(g, cutoff):
    """utils/utils.py:388-413 (all_atoms=False): drop the residues farther than ``cutoff`` from every ligand atom, with
    torch_geometric.utils.subgraph(..., relabel_nodes=True) restated for the receptor contact edges."""
    lig, rec = g['ligand'].pos, g['receptor'].pos
    keep = torch.any(torch.sum((lig.unsqueeze(0) - rec.unsqueeze(1)) ** 2, -1) < cutoff ** 2, dim=1)
    st = g['receptor']
    st.pos, st.x = st.pos[keep], st.x[keep]
    if 'side_chain_vecs' in st:
        st.side_chain_vecs = st.side_chain_vecs[keep]
    rr = g['receptor', 'receptor']
    ei = rr.edge_index
    relabel = torch.cumsum(keep.long(), 0) - 1
    ok = keep[ei[0]] & keep[ei[1]]
    rr.edge_index = relabel[ei[:, ok]]
    return g
