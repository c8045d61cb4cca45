"""Generalised advantage estimation in one kernel launch: the backward loop of
``PPOTrainer.update`` (/root/reference/posggym_baselines/ppo/core.py:226-244) that issues about six
tensor kernels per time step.  Bit-identical to the reference's float32 expressions."""
from __future__ import annotations

from typing import Tuple

import torch

from posggym_baselines_b200 import _lib


def compute_gae(rewards: torch.Tensor, dones: torch.Tensor, values: torch.Tensor, gamma: float,
                gae_lambda: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """``rewards`` ``[T, ...]``, ``dones`` / ``values`` ``[T + 1, ...]`` (float32, CUDA; the trailing
    dimensions are the batch).  Returns ``(advantages, returns)`` of shape ``[T, ...]``;
    ``returns = advantages + values[:-1]``."""
    if not rewards.is_cuda:
        raise _lib.PosgB200Error("compute_gae needs CUDA tensors: posggym_baselines_b200 has no CPU fallback")
    t = rewards.shape[0]
    assert dones.shape[0] == t + 1 and values.shape[0] == t + 1 and dones.shape[1:] == rewards.shape[1:]
    r = rewards.to(torch.float32).contiguous()
    d = dones.to(torch.float32).contiguous()
    v = values.to(torch.float32).contiguous()
    adv, ret = torch.empty_like(r), torch.empty_like(r)
    n = r[0].numel() if t > 0 else 0
    with torch.cuda.device(r.device):
        _lib.check(_lib.lib().posg_gae(t, n, _lib.ptr(r), _lib.ptr(d), _lib.ptr(v), float(gamma), float(gae_lambda),
                                       _lib.ptr(adv), _lib.ptr(ret), _lib.stream_ptr()))
    return adv, ret
