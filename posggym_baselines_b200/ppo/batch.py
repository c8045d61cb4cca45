"""Per-policy views of a rollout batch: the reference's ``split_batch_by_policy``,
``filter_batch_by_partner``, ``filter_policy_batches_by_partner_dist`` and ``combine_batches``
(/root/reference/posggym_baselines/ppo/utils.py:12-237) on device tensors.  The selections are boolean masks over
the (sequence, agent) dims applied with one index kernel per entry -- no loops over envs or sequences; ``config``
is any object with the reference's ``PPOConfig`` accessors (``num_agents``, ``get_all_policy_ids``,
``get_policy_idx``, ``get_policy_partner_distribution``)."""
from __future__ import annotations

from typing import Dict, List

import torch


def split_batch_by_policy(batch: Dict[str, torch.Tensor], config) -> Dict[str, Dict[str, torch.Tensor]]:
    """ppo/utils.py:12-71: entries of shape ``(L, B, N, ...)`` -> per policy ``(L, B_j, ...)``."""
    pol = batch["policy_idxs"]
    n_agents = config.num_agents
    partner = None
    if n_agents > 1:
        # partner policy ids of every (sequence, agent): shape (L, B, N, N - 1)
        cols = [[j for j in range(n_agents) if j != i] for i in range(n_agents)]
        idx = torch.tensor(cols, device=pol.device)                      # [N, N - 1]
        partner = pol[:, :, idx]                                         # gather along the agent dim
    out = {}
    for policy_id in config.get_all_policy_ids():
        mask = pol[0] == config.get_policy_idx(policy_id)                # [B, N]
        pb = {}
        for k, b in batch.items():
            if k == "initial_lstm_states":
                pb[k] = (b[0][:, mask], b[1][:, mask])
            elif k == "policy_stats":
                pb[k] = b.get(policy_id, {})                             # already split by policy in the worker
            else:
                pb[k] = b[:, mask]
        pb["partner_policy_idxs"] = partner[:, mask] if partner is not None else None
        out[policy_id] = pb
    return out


def filter_batch_by_partner(batch: Dict[str, torch.Tensor], exclude_policy_ids: List[str], config) -> Dict[str, torch.Tensor]:
    """ppo/utils.py:74-128: drop the sequences played with a partner in ``exclude_policy_ids``."""
    if config.num_agents == 1:
        return batch
    assert "partner_policy_idxs" in batch
    partners = batch["partner_policy_idxs"][0]                           # (B, N - 1): constant within a sequence
    exclude = torch.tensor([config.get_policy_idx(p) for p in exclude_policy_ids], device=partners.device,
                           dtype=partners.dtype)
    keep = torch.all(torch.isin(partners, exclude, invert=True), dim=-1)
    out = {}
    for k, b in batch.items():
        if k == "initial_lstm_states":
            out[k] = (b[0][:, keep], b[1][:, keep])
        elif k == "policy_stats":
            out[k] = b
        else:
            out[k] = b[:, keep]
    return out


def filter_policy_batches_by_partner_dist(policy_batches: Dict[str, Dict[str, torch.Tensor]], config):
    """ppo/utils.py:131-166: per policy, drop experience against partners its partner distribution gives 0."""
    all_ids = config.get_all_policy_ids()
    out = {}
    for policy_id, batch in policy_batches.items():
        dist = config.get_policy_partner_distribution(policy_id)
        out[policy_id] = filter_batch_by_partner(batch, [p for p in all_ids if dist.get(p, 0.0) == 0.0], config)
    return out


def combine_batches(batches: List[Dict[str, torch.Tensor]], config=None) -> Dict[str, torch.Tensor]:
    """ppo/utils.py:169-237: concatenate worker batches along the sequence dim; policy statistics are reduced
    by their name (``mean_`` / ``min_`` / ``max_``, else the first worker's value)."""
    out = {}
    for k in batches[0]:
        if k == "initial_lstm_states":
            out[k] = (torch.cat([b[k][0] for b in batches], dim=1), torch.cat([b[k][1] for b in batches], dim=1))
        elif k == "policy_stats":
            ids, names = set(), set()
            for b in batches:
                ids.update(b[k].keys())
                for stats in b[k].values():
                    names.update(stats.keys())
            out[k] = {}
            for pid in ids:
                stats = {}
                for name in names:
                    vals = [b[k][pid][name] for b in batches if pid in b[k] and name in b[k][pid]]
                    if not vals:
                        continue
                    v = torch.tensor(vals)
                    stats[name] = (torch.mean(v) if "mean_" in name else torch.min(v) if "min_" in name
                                   else torch.max(v) if "max_" in name else v[0])
                if stats:
                    out[k][pid] = stats
        else:
            out[k] = torch.cat([b[k] for b in batches], dim=1)
    return out
