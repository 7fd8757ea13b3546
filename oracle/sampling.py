"""Oracle restatement of the reverse-diffusion loop utils/sampling.py:69-231 including the confidence-model call after
the last step (:208-227; no visualisation).  TEST INFRASTRUCTURE."""
import numpy as np
import torch

from diffdock_b200.hetero import collate

import copy

from .diffusion import crop_beyond, modify_conformer_batch, set_time


def _triple(v):
    try:
        iter(v)
        return list(v)
    except TypeError:
        return [v] * 3


def sampling(data_list, model, inference_steps, tr_schedule, rot_schedule, tor_schedule, device, t_to_sigma,
             model_args, no_random=False, ode=False, batch_size=32, no_final_step_noise=False,
             temp_sampling=1.0, temp_psi=0.0, temp_sigma_data=0.5, noise_fn=None, confidence_model=None,
             confidence_data_list=None, confidence_model_args=None):
    """noise_fn(kind, shape) -> tensor replaces torch.normal when given (injected-noise parity runs); otherwise
    torch.normal is called in the reference's order (tr, rot, tor per step), so a shared torch.manual_seed
    reproduces the reference's CPU draws."""
    N = len(data_list)
    mask_rotate = torch.from_numpy(data_list[0]['ligand'].mask_rotate[0]).to(device)
    ts, tp, tsd = _triple(temp_sampling), _triple(temp_psi), _triple(temp_sigma_data)
    a = model_args

    def draw(kind, shape, zero):
        if zero:
            return torch.zeros(shape, device=device)
        if noise_fn is not None:
            return noise_fn(kind, shape).to(device)
        return torch.normal(mean=0, std=1, size=shape, device=device)

    confidence = []
    with torch.no_grad():
        for b0 in range(0, N, batch_size):
            chunk = data_list[b0:b0 + batch_size]
            g = collate(chunk).to(device)
            b = g.num_graphs
            n = len(g['ligand'].pos) // b
            for t_idx in range(inference_steps):
                t_tr, t_rot, t_tor = tr_schedule[t_idx], rot_schedule[t_idx], tor_schedule[t_idx]
                last = t_idx == inference_steps - 1
                dt_tr = tr_schedule[t_idx] - tr_schedule[t_idx + 1] if not last else tr_schedule[t_idx]
                dt_rot = rot_schedule[t_idx] - rot_schedule[t_idx + 1] if not last else rot_schedule[t_idx]
                dt_tor = tor_schedule[t_idx] - tor_schedule[t_idx + 1] if not last else tor_schedule[t_idx]
                tr_sigma, rot_sigma, tor_sigma = t_to_sigma(t_tr, t_rot, t_tor)
                if getattr(a, 'crop_beyond', None) is not None:         # sampling.py:104-109
                    mod = collate([crop_beyond(x, tr_sigma * 3 + a.crop_beyond)
                                   for x in copy.deepcopy(g).to_data_list()])
                else:
                    mod = g
                set_time(mod, t_tr, t_rot, t_tor, b, device)
                tr_score, rot_score, tor_score = model(mod)[:3]
                if torch.isnan(tr_score.mean(-1)).sum() > 0:            # sampling.py:117-131
                    for s in (tr_score, rot_score, tor_score):
                        eps = 0.01 * torch.nanmean(s.abs())
                        s.nan_to_num_(nan=eps, posinf=eps, neginf=-eps)
                tr_g = tr_sigma * torch.sqrt(torch.tensor(2 * np.log(a.tr_sigma_max / a.tr_sigma_min)))
                rot_g = rot_sigma * torch.sqrt(torch.tensor(2 * np.log(a.rot_sigma_max / a.rot_sigma_min)))
                zero = no_random or (no_final_step_noise and last)
                if ode:
                    tr_perturb = 0.5 * tr_g ** 2 * dt_tr * tr_score
                    rot_perturb = 0.5 * rot_score * dt_rot * rot_g ** 2
                else:
                    tr_z = draw('tr', (min(batch_size, N), 3), zero)
                    tr_perturb = tr_g ** 2 * dt_tr * tr_score + tr_g * np.sqrt(dt_tr) * tr_z
                    rot_z = draw('rot', (min(batch_size, N), 3), zero)
                    rot_perturb = rot_score * dt_rot * rot_g ** 2 + rot_g * np.sqrt(dt_rot) * rot_z
                if not a.no_torsion:
                    tor_g = tor_sigma * torch.sqrt(torch.tensor(2 * np.log(a.tor_sigma_max / a.tor_sigma_min)))
                    if ode:
                        tor_perturb = 0.5 * tor_g ** 2 * dt_tor * tor_score
                    else:
                        tor_z = draw('tor', tuple(tor_score.shape), zero)
                        tor_perturb = tor_g ** 2 * dt_tor * tor_score + tor_g * np.sqrt(dt_tor) * tor_z
                else:
                    tor_perturb = None
                if ts[0] != 1.0:                                        # sampling.py:173-176
                    sd = np.exp(tsd[0] * np.log(a.tr_sigma_max) + (1 - tsd[0]) * np.log(a.tr_sigma_min))
                    lam = (sd + tr_sigma) / (sd + tr_sigma / ts[0])
                    tr_perturb = tr_g ** 2 * dt_tr * (lam + ts[0] * tp[0] / 2) * tr_score \
                        + tr_g * np.sqrt(dt_tr * (1 + tp[0])) * tr_z
                if ts[1] != 1.0:
                    sd = np.exp(tsd[1] * np.log(a.rot_sigma_max) + (1 - tsd[1]) * np.log(a.rot_sigma_min))
                    lam = (sd + rot_sigma) / (sd + rot_sigma / ts[1])
                    rot_perturb = rot_g ** 2 * dt_rot * (lam + ts[1] * tp[1] / 2) * rot_score \
                        + rot_g * np.sqrt(dt_rot * (1 + tp[1])) * rot_z
                if ts[2] != 1.0:
                    sd = np.exp(tsd[2] * np.log(a.tor_sigma_max) + (1 - tsd[2]) * np.log(a.tor_sigma_min))
                    lam = (sd + tor_sigma) / (sd + tor_sigma / ts[2])
                    tor_perturb = tor_g ** 2 * dt_tor * (lam + ts[2] * tp[2] / 2) * tor_score \
                        + tor_g * np.sqrt(dt_tor * (1 + tp[2])) * tor_z
                g['ligand'].pos = modify_conformer_batch(g['ligand'].pos, g, tr_perturb, rot_perturb,
                                                         tor_perturb if not a.no_torsion else None, mask_rotate)
            for i in range(b):
                data_list[b0 + i]['ligand'].pos = g['ligand'].pos[i * n:n * (i + 1)]
            if confidence_model is not None:                                     # utils/sampling.py:208-227
                if confidence_data_list is not None:
                    cl = copy.deepcopy(confidence_data_list[b0:b0 + batch_size])
                    cg = collate(cl)
                    cg['ligand'].pos = g['ligand'].pos.cpu()
                    cb = getattr(confidence_model_args, 'crop_beyond', None)
                    if cb is not None:                                           # :213-217, per complex
                        parts = cg.to_data_list()
                        for part in parts:
                            crop_beyond(part, cb)
                        cg = collate(parts)
                    set_time(cg, 0, 0, 0, b, device)
                    out = confidence_model(cg)
                else:
                    out = confidence_model(g)
                confidence.append(out[0] if type(out) is tuple else out)
    if confidence_model is not None:
        return data_list, torch.nan_to_num(torch.cat(confidence, dim=0), nan=-1000)
    return data_list, None
