"""Restatement of the PPO learner's sequence-chunking helpers (oracle; test infrastructure only).

Follows /root/reference/posggym_baselines/ppo/utils.py: ``get_seq_idxs`` (240-292),
``split_and_pad_batch`` (295-344) and ``split_lstm_state_batch`` (347-383) -- per-env Python loops
over the ``dones`` buffer that build nested lists of step-index tensors.  Pinned by
tests/test_ppo_seq.py against the reference module itself when /root/reference is importable."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn.functional as F


def get_seq_idxs(dones_buf: torch.Tensor, max_seq_len: int) -> Tuple[List[List[torch.Tensor]], List[List[torch.Tensor]]]:
    steps, num_envs = dones_buf.shape
    seq_idxs, overlapping = [], []
    for env in range(num_envs):
        cuts = [int(i) for i in torch.nonzero(dones_buf[:, env]).flatten().tolist()]
        bounds = [0] + cuts + [steps]
        chunks = []
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            for s in range(lo, hi, max_seq_len):           # empty pieces (lo == hi) yield nothing
                chunks.append(torch.arange(s, min(s + max_seq_len, hi)))
        seq_idxs.append(chunks)
        overlapping.append([torch.cat((c, c[-1:] + 1)) for c in chunks])
    return seq_idxs, overlapping


def split_and_pad_batch(batch: torch.Tensor, seq_idxs, max_seq_len: int, padding_value: float = 0.0) -> torch.Tensor:
    seqs = [batch[idx, env] for env, chunks in enumerate(seq_idxs) for idx in chunks]
    out = torch.nn.utils.rnn.pad_sequence(seqs, batch_first=False, padding_value=padding_value)
    if out.shape[0] < max_seq_len:
        pad = [0, 0] * out.dim()
        pad[-1] = max_seq_len - out.shape[0]
        out = F.pad(out, pad, mode="constant", value=0.0)
    return out


def split_lstm_state_batch(lstm_state_batch: torch.Tensor, seq_idxs) -> torch.Tensor:
    firsts = [lstm_state_batch[idx[0], :, env] for env, chunks in enumerate(seq_idxs) for idx in chunks]
    return torch.stack(firsts, dim=1)
